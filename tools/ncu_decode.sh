#!/bin/bash
# ncu captures of the decode program kernel (run under gpurun, one GPU):
#   1. plain run (must exit 0), 2. launch list of bench.py, 3. --set full of one decode_w4_kernel launch (4 layers: the
#   kernel's 128-stage launch replayed ~40 times would take minutes; 16 stages show the same per-stage behaviour).
set -e
python bench.py --steps 2 --warmup 3 --no-gemm --no-cpu-baseline > gpurun_out/r2_ncu_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'decode_w4_kernel|gemv_w4_kernel|gemm_w4|dow_kernel' -c 80 --csv --log-file gpurun_out/r2_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-gemm --no-cpu-baseline > gpurun_out/r2_ncu_list.log 2>&1
python bench.py --steps 2 --warmup 3 --no-gemm --no-cpu-baseline --layers 4 > gpurun_out/r2_ncu_plain4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:decode_w4_kernel -s 3 -c 1 -o gpurun_out/r2_decode_w4 \
    python bench.py --steps 2 --warmup 3 --no-gemm --no-cpu-baseline --layers 4 > gpurun_out/r2_ncu_full.log 2>&1
