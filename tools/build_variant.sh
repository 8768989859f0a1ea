#!/bin/bash
# Development helper: build a second copy of the library with a variant of ONE source file (extra nvcc flags), for
# A/B timing in a single gpurun call.  usage: tools/build_variant.sh NAME [file.cu] [extra nvcc flags...]
# The result is qeft_b200/csrc/variants/libqeft_b200_NAME.so; select it with QEFT_B200_LIB=<path>.
set -e
cd "$(dirname "$0")/../qeft_b200/csrc"
name=$1; shift
src=${1:-decode_w4.cu}; shift || true
mkdir -p variants build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC \
  -Xcompiler -fvisibility=hidden -I ../../include "$@" -c "$src" -o "variants/${src%.cu}_$name.o"
objs=""
for f in *.cu; do
  if [ "$f" == "$src" ]; then objs="$objs variants/${src%.cu}_$name.o"; else objs="$objs build/${f%.cu}.o"; fi
done
nvcc -shared -o "variants/libqeft_b200_$name.so" $objs -cudart static -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a
echo "variants/libqeft_b200_$name.so"
