"""Minimal driver for ncu: a few prefill GEMM launches of one shape."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qeft_b200 import qeft_cuda  # noqa: E402
from qeft_b200.synth import synth_tensors  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="4096x4096")
ap.add_argument("--M", type=int, default=2048)
ap.add_argument("--iters", type=int, default=6)
ap.add_argument("--op", default="fwd", choices=["fwd", "dx", "dow"])
args = ap.parse_args()
N, K = map(int, args.shape.split("x"))
t = synth_tensors(N, K, seed=0)
x = torch.randn(args.M, K, device="cuda").half()
dy = torch.randn(args.M, N, device="cuda").half()
for i in range(args.iters):
    if args.op == "fwd":
        y = qeft_cuda.gemm_w4(x, t["qweight"], t["scales"], t["scaled_zeros"], t["oweight"], None, pdl=False)
    elif args.op == "dx":
        y = qeft_cuda.gemm_w4_dx(dy, t["qweight"], t["scales"], t["scaled_zeros"], t["oweight"], K, pdl=False)
    else:
        y = qeft_cuda.dow(dy, x[:, K - 128:].contiguous(), 128)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
