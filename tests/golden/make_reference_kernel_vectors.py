"""Golden vectors from the reference's OWN CUDA kernels, run on a B200 (needs oracle/_ref/qeft_cuda_ref.so and a GPU).

    gpurun -- 'python tests/golden/make_reference_kernel_vectors.py gpurun_out/reference_kernel_vectors.npz'
    cp gpurun_out/reference_kernel_vectors.npz tests/golden/

For small seeded layers (oracle.synth_layer) it stores the packed tensors, the activations and what the reference
kernels returned for them: gemv_4bit_qeft (g128 and per-channel), gemv_4bit, gemm_4bit (+ the outlier F.linear the
reference adds, qlinear.py:264-266).  tests/test_oracle_golden.py checks the numpy oracle's forward against these on
the CPU, so the oracle's forward arithmetic stays pinned to reference-run outputs where no GPU / no reference module
exists.  The reference rounds every dequantised weight to fp16 and accumulates partly in fp16: agreement is to ~2e-3 of
the output scale, not bit-exact.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import build_ref  # noqa: E402

CASES = [
    # kind, N, K, r, G, m (tokens)
    ("gemv_qeft", 64, 256, 32, 128, 1),
    ("gemv_qeft", 128, 512, 128, 128, 2),
    ("gemv_qeft", 64, 1024, 64, 128, 7),
    ("gemv_qeft", 64, 512, 128, 512, 3),      # per-channel scales: gemv_kernel_qeft_perchannel
    ("gemv", 64, 256, 0, 128, 1),
    ("gemv", 128, 512, 0, 128, 5),
    ("gemm", 128, 256, 64, 128, 24),          # gemm_4bit tile config for <= 32 tokens (split-K 2)
    ("gemm", 128, 512, 128, 128, 100),        # <= 128 tokens
    ("gemm", 128, 256, 64, 128, 300),         # > 192 tokens (gemm_w4a16_T2)
]


def main(out):
    ref = build_ref.load()
    assert ref is not None and torch.cuda.is_available()
    d = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()  # noqa: E731
    store = {"cases": np.array([f"{k}:{N}:{K}:{r}:{G}:{m}" for k, N, K, r, G, m in CASES])}
    for i, (kind, N, K, r, G, m) in enumerate(CASES):
        L = oracle.synth_layer(N, K, r=r, G=G, seed=900 + i)
        x = np.random.default_rng(950 + i).standard_normal((m, K)).astype(np.float16)
        a = (d(x), d(L["qweight"]), d(L["scales"]), d(L["scaled_zeros"]))
        if kind == "gemv_qeft":
            y = ref.gemv_4bit_qeft(*a, d(L["oweight_interleaved"]), m, N, K, G)
        elif kind == "gemv":
            y = ref.gemv_4bit(*a, m, N, K, G)
        else:
            y = ref.gemm_4bit(*a) + torch.nn.functional.linear(a[0][..., -r:], d(L["oweight"]))
        torch.cuda.synchronize()
        for k in ("qweight", "scales", "scaled_zeros", "oweight", "oweight_interleaved"):
            if k in L:
                store[f"c{i}_{k}"] = L[k]
        store[f"c{i}_x"] = x
        store[f"c{i}_y_ref"] = y.cpu().numpy()
        want = oracle.forward(x, L["qweight"], L["scales"], L["scaled_zeros"], L.get("oweight"), None, group_size=G)
        err = np.max(np.abs(store[f"c{i}_y_ref"].astype(np.float64) - want.astype(np.float64))) / np.max(np.abs(want.astype(np.float64)))
        print(f"case {i} {kind} N={N} K={K} r={r} G={G} m={m}: reference kernel vs oracle {err:.2e}")
    store["device"] = np.array(torch.cuda.get_device_name(0))
    np.savez_compressed(out, **store)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "reference_kernel_vectors.npz"))
