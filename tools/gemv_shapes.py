"""Per-shape decode GEMV timing (CUDA events, weights rotated through > L2 bytes so they stream from HBM).

    python tools/gemv_shapes.py [--m 1] [--iters 200]
Prints one JSON line per (shape, mode): achieved algorithmic GB/s against MEASURED_PEAKS.json.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qeft_b200 import _lib, qeft_cuda  # noqa: E402
from qeft_b200.synth import synth_tensors  # noqa: E402


def algo_bytes(N, K, m, r=128, G=128):
    return N * (K - r) // 2 + 4 * N * ((K - r) // G) + 2 * N * r + 2 * K * m + 2 * N * m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=1)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--shapes", default="4096x4096,11008x4096,4096x11008,12288x4096,22016x4096,8192x8192,28672x8192,8192x28672")
    args = ap.parse_args()
    peak = 6553.0
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    m = args.m
    for shp in args.shapes.split(","):
        N, K = map(int, shp.split("x"))
        per = algo_bytes(N, K, m)
        copies = max(2, int(400e6 // per) + 1)          # > 3x L2
        layers = [synth_tensors(N, K, seed=i) for i in range(copies)]
        x = torch.randn(m, K, device="cuda").half()
        ys = [torch.empty(m, N, device="cuda", dtype=torch.float16) for _ in range(copies)]

        def launch(i, pdl):
            t = layers[i % copies]
            qeft_cuda.gemv_w4(x, t["qweight"], t["scales"], t["scaled_zeros"], t["oweight_interleaved"], m, N, K, 128,
                              ow_layout=_lib.OW_INTERLEAVED, out=ys[i % copies], pdl=pdl)

        for mode in ("stream", "stream+pdl", "graph", "graph+pdl"):
            pdl = mode.endswith("pdl")
            iters = (args.iters // copies + 1) * copies
            if mode.startswith("graph"):
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    for i in range(copies):
                        launch(i, pdl)
                torch.cuda.current_stream().wait_stream(s)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for i in range(copies):
                        launch(i, pdl)
                run = lambda: [g.replay() for _ in range(iters // copies)]  # noqa: E731
            else:
                run = lambda: [launch(i, pdl) for i in range(iters)]  # noqa: E731
            run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / iters
            gbs = per / us / 1e3
            print(json.dumps({"shape": shp, "m": m, "mode": mode, "us": round(us, 3), "GBps": round(gbs, 1),
                              "frac_measured_peak": round(gbs / peak, 3), "copies": copies}), flush=True)
        del layers, ys
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
