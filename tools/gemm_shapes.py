"""Per-shape prefill GEMM timing (CUDA events): achieved algorithmic TFLOP/s (2 M N K) against MEASURED_PEAKS.json.

    python tools/gemm_shapes.py [--M 2048] [--iters 20] [--shapes 4096x4096,...] [--bwd]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qeft_b200 import qeft_cuda  # noqa: E402
from qeft_b200.synth import synth_tensors  # noqa: E402


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters     # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--M", type=int, default=2048)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--shapes", default="4096x4096,11008x4096,4096x11008,12288x4096,22016x4096")
    ap.add_argument("--bwd", action="store_true")
    ap.add_argument("--cublas", action="store_true", help="also time torch.matmul on a dense fp16 weight (library reference)")
    args = ap.parse_args()
    peak = 1590.0
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["bf16_tflops"]
    M = args.M
    for shp in args.shapes.split(","):
        N, K = map(int, shp.split("x"))
        t = synth_tensors(N, K, seed=1)
        x = torch.randn(M, K, device="cuda").half()
        y = torch.empty(M, N, device="cuda", dtype=torch.float16)
        flops = 2.0 * M * N * K
        us = timeit(lambda: qeft_cuda.gemm_w4(x, t["qweight"], t["scales"], t["scaled_zeros"], t["oweight"], None, out=y,
                                              pdl=False), args.iters)
        line = {"op": "gemm_w4", "shape": shp, "M": M, "us": round(us, 2), "TFLOPs": round(flops / us / 1e6, 1),
                "frac_measured_bf16_peak": round(flops / us / 1e6 / peak, 3)}
        print(json.dumps(line), flush=True)
        if args.cublas:
            w = qeft_cuda.dequant_w4(t["qweight"], t["scales"], t["scaled_zeros"], t["oweight"])
            us = timeit(lambda: torch.matmul(x, w.t(), out=y), args.iters)
            print(json.dumps({"op": "cublas_dense_fp16", "shape": shp, "M": M, "us": round(us, 2),
                              "TFLOPs": round(flops / us / 1e6, 1)}), flush=True)
        if args.bwd:
            dy = torch.randn(M, N, device="cuda").half()
            dx = torch.empty(M, K, device="cuda", dtype=torch.float16)
            us = timeit(lambda: qeft_cuda.gemm_w4_dx(dy, t["qweight"], t["scales"], t["scaled_zeros"], t["oweight"], K,
                                                     out=dx, pdl=False), args.iters)
            print(json.dumps({"op": "gemm_w4_dx", "shape": shp, "M": M, "us": round(us, 2),
                              "TFLOPs": round(flops / us / 1e6, 1),
                              "frac_measured_bf16_peak": round(flops / us / 1e6 / peak, 3)}), flush=True)
            dow = torch.zeros(N, 128, device="cuda", dtype=torch.float32)
            x_out = x[:, K - 128:].contiguous()          # (the autograd function keeps this compact copy from the forward)
            us = timeit(lambda: qeft_cuda.dow(dy, x_out, 128, out=dow, pdl=False), args.iters)
            print(json.dumps({"op": "dow", "shape": shp, "M": M, "us": round(us, 2),
                              "TFLOPs": round(2.0 * M * N * 128 / us / 1e6, 1)}), flush=True)
        del t
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
