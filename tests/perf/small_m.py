"""Small token counts (8 <= M <= 256): the tcgen05 GEMM against ceil(M / 8) decode-GEMV launches and the reference's
gemm_4bit (+ outlier F.linear), per Llama-2-7B layer shape.  Decides the host-side dispatch in QuantLinear.forward.

    python tests/perf/small_m.py
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402  (under tests/: the only tree besides bench.py / smoke() that may touch oracle/)
from qeft_b200 import _lib, qeft_cuda  # noqa: E402
from qeft_b200.synth import synth_tensors  # noqa: E402


def timed(fn, iters=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


ref = build_ref.load()
for N, K in ((4096, 4096), (11008, 4096), (4096, 11008)):
    copies = 12
    layers = [synth_tensors(N, K, seed=i, fast=True) for i in range(copies)]      # > L2 in total
    for M in (8, 16, 32, 64, 128, 256):
        x = torch.randn(M, K, device="cuda").half()
        y = torch.empty(M, N, device="cuda", dtype=torch.float16)
        ctr = [0]

        def nxt():
            ctr[0] += 1
            return layers[ctr[0] % copies]

        def gemm():
            t = nxt()
            qeft_cuda.gemm_w4(x, t["qweight"], t["scales"], t["scaled_zeros"], t["oweight"], None, out=y)

        def gemv_chunks():
            t = nxt()
            for r0 in range(0, M, 8):
                m = min(8, M - r0)
                qeft_cuda.gemv_w4(x[r0:r0 + m], t["qweight"], t["scales"], t["scaled_zeros"], t["oweight_interleaved"], m, N, K,
                                  128, ow_layout=_lib.OW_INTERLEAVED, out=y[r0:r0 + m])

        def ref_gemm():
            t = nxt()
            return ref.gemm_4bit(x, t["qweight"], t["scales"], t["scaled_zeros"]) + torch.nn.functional.linear(x[..., -128:], t["oweight"])

        row = {"shape": f"{N}x{K}", "M": M, "gemm_us": round(timed(gemm), 1), "gemv_chunks_us": round(timed(gemv_chunks), 1)}
        if ref is not None:
            row["ref_gemm_us"] = round(timed(ref_gemm), 1)
        print(json.dumps(row), flush=True)
    del layers
    torch.cuda.empty_cache()
