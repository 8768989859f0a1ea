"""In-kernel timeline of the decode GEMV (debug): CTA 0's globaltimer stamps for a chain of launches in a CUDA graph.

    QEFT_GEMV_STAMPS=1 python tools/gemv_timeline.py [--shape 4096x4096]
Stamps: 0 start (after setup barrier), 1 ring filled + partials zeroed, 2 after griddepcontrol.wait, 3 x staged,
4 main loop done, 5 results stored.
"""
import argparse
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qeft_b200 import _lib, qeft_cuda  # noqa: E402
from qeft_b200.synth import synth_tensors  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="4096x4096")
ap.add_argument("--n", type=int, default=12)
ap.add_argument("--no-pdl", action="store_true")
args = ap.parse_args()
N, K = map(int, args.shape.split("x"))
layers = [synth_tensors(N, K, seed=i) for i in range(args.n)]
x = torch.randn(1, K, device="cuda").half()
ys = [torch.empty(1, N, device="cuda", dtype=torch.float16) for _ in range(args.n)]


def run():
    for t, y in zip(layers, ys):
        qeft_cuda.gemv_w4(x, t["qweight"], t["scales"], t["scaled_zeros"], t["oweight_interleaved"], 1, N, K, 128,
                          ow_layout=_lib.OW_INTERLEAVED, out=y, pdl=not args.no_pdl)


s = torch.cuda.Stream()
with torch.cuda.stream(s):
    run()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    run()
torch.cuda.synchronize()
g.replay()
torch.cuda.synchronize()
lib = _lib.load()
lib.qeft_gemv_debug_stamps.restype = C.c_int
total = 3 * args.n
buf = (C.c_ulonglong * (total * 8))()
assert lib.qeft_gemv_debug_stamps(buf, total) == 0
# the graph replay re-uses the slots of the captured launches (slots n .. 2n-1): replayed kernels carry those pointers
rows = [[buf[i * 8 + j] for j in range(7)] for i in range(args.n, 2 * args.n)]
t0 = rows[0][0]
print("launch  start  filled  waited  staged  looped  stored  last-CTA-stored (us, relative to the first launch's start; the last column accumulates over replays)")
for i, r in enumerate(rows):
    print(i, " ".join(f"{(v - t0) / 1e3:8.2f}" for v in r))

# per-CTA view of one mid-chain launch (slot = launch number % 64)
import numpy as np
slot = (args.n + args.n // 2) % 64
cb = (C.c_ulonglong * (512 * 4))()
lib.qeft_gemv_debug_cta_stamps.restype = C.c_int
if lib.qeft_gemv_debug_cta_stamps(cb, slot) == 0:
    a = np.array(cb[:], dtype=np.int64).reshape(512, 4)
    a = a[a[:, 0] > 0]
    base = a[:, 0].min()
    rel = (a - base) / 1e3
    names = ["start", "waited", "staged", "stored"]
    print(f"per-CTA stamps of launch {args.n // 2} ({len(a)} CTAs), us relative to the earliest start:")
    for j, nm in enumerate(names):
        col = rel[:, j]
        print(f"  {nm:7s} min {col.min():7.2f}  p50 {np.median(col):7.2f}  p90 {np.percentile(col, 90):7.2f}  max {col.max():7.2f}")
    d = rel[:, 3] - rel[:, 2]
    print(f"  staged->stored per CTA: min {d.min():.2f} p50 {np.median(d):.2f} max {d.max():.2f}")
    d = rel[:, 2] - rel[:, 1]
    print(f"  waited->staged per CTA: min {d.min():.2f} p50 {np.median(d):.2f} max {d.max():.2f}")
