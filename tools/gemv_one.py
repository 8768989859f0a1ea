"""Minimal driver for ncu: a few decode GEMV launches of one shape over rotating weight copies."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qeft_b200 import _lib, qeft_cuda  # noqa: E402
from qeft_b200.synth import synth_tensors  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="4096x4096")
ap.add_argument("--m", type=int, default=1)
ap.add_argument("--iters", type=int, default=24)
ap.add_argument("--copies", type=int, default=24)
args = ap.parse_args()
N, K = map(int, args.shape.split("x"))
layers = [synth_tensors(N, K, seed=i) for i in range(args.copies)]
x = torch.randn(args.m, K, device="cuda").half()
y = torch.empty(args.m, N, device="cuda", dtype=torch.float16)
for i in range(args.iters):
    t = layers[i % args.copies]
    qeft_cuda.gemv_w4(x, t["qweight"], t["scales"], t["scaled_zeros"], t["oweight_interleaved"], args.m, N, K, 128,
                      ow_layout=_lib.OW_INTERLEAVED, out=y, pdl=False)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
