"""The reference's own kernels (oracle/_ref/qeft_cuda_ref.so, unmodified, recompiled for sm_100a) timed beside ours
on the same B200, same packed tensors, same shapes.

    python tests/perf/ref_kernels.py [--iters 100]
One JSON line per shape: decode GEMV (m = 1) and prefill GEMM (M = 2048).  Both sides are timed as plain launches on
the default stream with CUDA events (the reference launches on the legacy default stream and cannot be graph-captured);
weights rotate through > 3x L2 bytes.  `ours_chain_us` is our GEMV inside a CUDA graph with PDL (how decode.py runs it).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402  (under tests/: the only tree besides bench.py / smoke() that may touch oracle/)
from qeft_b200 import _lib, qeft_cuda  # noqa: E402
from qeft_b200.synth import synth_tensors  # noqa: E402


def timed(fn, iters):
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--shapes", default="4096x4096,12288x4096,22016x4096,4096x11008,8192x8192,28672x8192")
    ap.add_argument("--M", type=int, default=2048)
    args = ap.parse_args()
    ref = build_ref.load()
    assert ref is not None, "oracle/_ref/qeft_cuda_ref.so missing: python oracle/build_ref.py (build container)"
    for shp in args.shapes.split(","):
        N, K = map(int, shp.split("x"))
        per = N * (K - 128) // 2 + 4 * N * ((K - 128) // 128) + 2 * N * 128 + 2 * K + 2 * N
        copies = max(2, int(400e6 // per) + 1)
        layers = [synth_tensors(N, K, seed=i, fast=True) for i in range(copies)]
        x = torch.randn(1, K, device="cuda").half()
        ctr = [0]

        def ref_gemv():
            t = layers[ctr[0] % copies]; ctr[0] += 1
            return ref.gemv_4bit_qeft(x, t["qweight"], t["scales"], t["scaled_zeros"], t["oweight_interleaved"], 1, N, K, 128)

        out = torch.empty(1, N, device="cuda", dtype=torch.float16)

        def our_gemv():
            t = layers[ctr[0] % copies]; ctr[0] += 1
            return qeft_cuda.gemv_w4(x, t["qweight"], t["scales"], t["scaled_zeros"], t["oweight_interleaved"], 1, N, K, 128,
                                     ow_layout=_lib.OW_INTERLEAVED, out=out)

        iters = (args.iters // copies + 1) * copies
        t_ref = timed(ref_gemv, iters)
        t_our = timed(our_gemv, iters)
        # ours as decode.py runs it: graph + PDL
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for i in range(copies):
                our_gemv()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(copies):
                t = layers[i]
                qeft_cuda.gemv_w4(x, t["qweight"], t["scales"], t["scaled_zeros"], t["oweight_interleaved"], 1, N, K, 128,
                                  ow_layout=_lib.OW_INTERLEAVED, out=out, pdl=True)
        t_chain = timed(g.replay, max(3, iters // copies)) / copies
        # prefill GEMM, M tokens: reference = gemm_4bit + F.linear on the outlier columns (qlinear.py:264-266)
        M = args.M
        xm = torch.randn(M, K, device="cuda").half()
        t0 = layers[0]

        def ref_gemm():
            y = ref.gemm_4bit(xm, t0["qweight"], t0["scales"], t0["scaled_zeros"])
            return y + torch.nn.functional.linear(xm[..., -128:], t0["oweight"])

        ym = torch.empty(M, N, device="cuda", dtype=torch.float16)

        def our_gemm():
            return qeft_cuda.gemm_w4(xm, t0["qweight"], t0["scales"], t0["scaled_zeros"], t0["oweight"], None, out=ym)

        g_ref = timed(ref_gemm, 20)
        g_our = timed(our_gemm, 20)
        fl = 2.0 * M * N * K
        print(json.dumps({
            "shape": shp,
            "gemv_m1": {"ref_us": round(t_ref, 2), "ours_us": round(t_our, 2), "ours_chain_us": round(t_chain, 2),
                        "ref_GBps": round(per / t_ref / 1e3, 1), "ours_GBps": round(per / t_our / 1e3, 1),
                        "ours_chain_GBps": round(per / t_chain / 1e3, 1), "speedup_stream": round(t_ref / t_our, 2),
                        "speedup_chain": round(t_ref / t_chain, 2)},
            f"gemm_M{M}": {"ref_us": round(g_ref, 1), "ours_us": round(g_our, 1), "ref_TFLOPs": round(fl / g_ref / 1e6, 1),
                           "ours_TFLOPs": round(fl / g_our / 1e6, 1), "speedup": round(g_ref / g_our, 2)},
        }), flush=True)
        del layers
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
