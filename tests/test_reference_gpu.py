"""Parity against the reference's OWN kernels, run on the same B200.

`oracle/_ref/qeft_cuda_ref.so` holds the unmodified reference kernels of this path (gemv_4bit, gemv_4bit_qeft,
gemm_4bit), compiled for sm_100a from the sources under /root/reference by `oracle/build_ref.py` in the build container
and shipped prebuilt.  Three-way checks on the same seeded inputs:

  * reference kernel vs the numpy oracle  -> pins the oracle's forward arithmetic (layout, nibble order, scale /
    scaled-zero convention, interleaved outlier rows, dead columns) to reference-run outputs;
  * our kernel vs the reference kernel    -> the drop-in claim itself, through the same three signatures;
  * our error against the fp64 oracle is not larger than the reference's own.

The reference dequantises in fp16 (half2 fma) and accumulates partly in fp16, so it is the noisier of the two; the
tolerances are the north-star's (1e-3 of the output scale for fp16) with head-room for the reference's own rounding.
"""
import numpy as np
import pytest
import torch

import oracle
from oracle import build_ref

pytestmark = pytest.mark.gpu

ref = build_ref.load()
needs_ref = pytest.mark.skipif(ref is None, reason="oracle/_ref/qeft_cuda_ref.so not built (python oracle/build_ref.py)")

G = 128


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-30))


def layer(N, K, r, seed):
    L = oracle.synth_layer(N, K, r=r, G=G, seed=seed)
    return L, {k: dev(L[k]) for k in ("qweight", "scales", "scaled_zeros") + (("oweight", "oweight_interleaved") if r else ())}


@needs_ref
@pytest.mark.parametrize("m", [1, 2, 3, 4, 7])
@pytest.mark.parametrize("N,K,r", [(256, 512, 128), (512, 1024, 64), (1024, 4096, 128)])
def test_gemv_qeft_three_way(N, K, r, m):
    from qeft_b200 import qeft_cuda
    L, D = layer(N, K, r, seed=N + K + m)
    x = np.random.default_rng(m).standard_normal((m, K)).astype(np.float16)
    xd = dev(x)
    y_ref = ref.gemv_4bit_qeft(xd, D["qweight"], D["scales"], D["scaled_zeros"], D["oweight_interleaved"], m, N, K, G)
    y_our = qeft_cuda.gemv_4bit_qeft(xd, D["qweight"], D["scales"], D["scaled_zeros"], D["oweight_interleaved"], m, N, K, G)
    torch.cuda.synchronize()
    want = oracle.forward(x, L["qweight"], L["scales"], L["scaled_zeros"], L["oweight"], None)
    e_ref, e_our, e_x = rel(y_ref.cpu().numpy(), want), rel(y_our.cpu().numpy(), want), rel(y_our.cpu().numpy(), y_ref.cpu().numpy())
    print(f"gemv_qeft {N}x{K} r={r} m={m}: ref-oracle {e_ref:.2e}  ours-oracle {e_our:.2e}  ours-ref {e_x:.2e}")
    assert y_ref.shape == y_our.shape and y_ref.dtype == y_our.dtype
    assert e_ref <= 4e-3          # the oracle restates what the reference kernel computes
    assert e_our <= 1e-3          # north-star tolerance (fp16)
    assert e_x <= 4e-3
    assert e_our <= e_ref + 1e-4


@needs_ref
@pytest.mark.parametrize("m", [1, 2, 5])
@pytest.mark.parametrize("N,K,r", [(256, 512, 128), (512, 2048, 64)])
def test_gemv_qeft_per_channel_three_way(N, K, r, m):
    """group_size != 128 sends the reference to gemv_kernel_qeft_perchannel (gemv_cuda_qeft.cu:461-): scales [1, N]."""
    from qeft_b200 import qeft_cuda
    L = oracle.synth_layer(N, K, r=r, G=K, seed=7 * N + m)
    D = {k: dev(L[k]) for k in ("qweight", "scales", "scaled_zeros", "oweight_interleaved")}
    assert L["scales"].shape == (1, N)
    x = np.random.default_rng(30 + m).standard_normal((m, K)).astype(np.float16)
    xd = dev(x)
    y_ref = ref.gemv_4bit_qeft(xd, D["qweight"], D["scales"], D["scaled_zeros"], D["oweight_interleaved"], m, N, K, K)
    y_our = qeft_cuda.gemv_4bit_qeft(xd, D["qweight"], D["scales"], D["scaled_zeros"], D["oweight_interleaved"], m, N, K, K)
    torch.cuda.synchronize()
    want = oracle.forward(x, L["qweight"], L["scales"], L["scaled_zeros"], L["oweight"], None, group_size=K)
    e_ref, e_our = rel(y_ref.cpu().numpy(), want), rel(y_our.cpu().numpy(), want)
    print(f"gemv_qeft per-channel {N}x{K} m={m}: ref-oracle {e_ref:.2e}  ours-oracle {e_our:.2e}")
    assert e_ref <= 4e-3 and e_our <= 1e-3 and e_our <= e_ref + 1e-4


@needs_ref
@pytest.mark.parametrize("m", [1, 4, 7])
@pytest.mark.parametrize("N,K", [(256, 512), (1024, 2048)])
def test_gemv_plain_three_way(N, K, m):
    from qeft_b200 import qeft_cuda
    L, D = layer(N, K, 0, seed=3 * N + K + m)
    x = np.random.default_rng(10 + m).standard_normal((m, K)).astype(np.float16)
    xd = dev(x)
    y_ref = ref.gemv_4bit(xd, D["qweight"], D["scales"], D["scaled_zeros"], m, N, K, G)
    y_our = qeft_cuda.gemv_4bit(xd, D["qweight"], D["scales"], D["scaled_zeros"], m, N, K, G)
    torch.cuda.synchronize()
    want = oracle.forward(x, L["qweight"], L["scales"], L["scaled_zeros"], None, None)
    e_ref, e_our = rel(y_ref.cpu().numpy(), want), rel(y_our.cpu().numpy(), want)
    print(f"gemv {N}x{K} m={m}: ref-oracle {e_ref:.2e}  ours-oracle {e_our:.2e}")
    assert e_ref <= 4e-3 and e_our <= 1e-3 and e_our <= e_ref + 1e-4


@needs_ref
@pytest.mark.parametrize("M", [8, 24, 48, 100, 160, 256, 1000])      # one per tile configuration of gemm_cuda.cu:944-1029
@pytest.mark.parametrize("N,K,r", [(256, 512, 128), (1024, 2048, 128)])
def test_gemm_three_way(N, K, r, M):
    """Reference forward_gemm_qeft (qlinear.py:262-268): gemm_4bit over all K columns + F.linear on the outlier columns."""
    from qeft_b200 import qeft_cuda
    L, D = layer(N, K, r, seed=5 * N + K + M)
    x = np.random.default_rng(20 + M).standard_normal((M, K)).astype(np.float16)
    xd = dev(x)
    y_ref = ref.gemm_4bit(xd, D["qweight"], D["scales"], D["scaled_zeros"])
    y_ref = y_ref + torch.nn.functional.linear(xd[..., -r:], D["oweight"])
    y_api = qeft_cuda.gemm_4bit(xd, D["qweight"], D["scales"], D["scaled_zeros"])        # same signature, all K columns
    y_api = y_api + torch.nn.functional.linear(xd[..., -r:], D["oweight"])
    y_fused = qeft_cuda.gemm_w4(xd, D["qweight"], D["scales"], D["scaled_zeros"], D["oweight"], None, group_size=G)
    torch.cuda.synchronize()
    want = oracle.forward(x, L["qweight"], L["scales"], L["scaled_zeros"], L["oweight"], None)
    e_ref, e_api, e_fused = (rel(t.cpu().numpy(), want) for t in (y_ref, y_api, y_fused))
    print(f"gemm {N}x{K} M={M}: ref-oracle {e_ref:.2e}  ours(api) {e_api:.2e}  ours(fused) {e_fused:.2e}")
    assert e_ref <= 4e-3 and e_api <= 1.5e-3 and e_fused <= 1e-3
    assert rel(y_fused.cpu().numpy(), y_ref.cpu().numpy()) <= 4e-3


@needs_ref
@pytest.mark.parametrize("N,K", [(4096, 4096), (11008, 4096), (4096, 11008)])
def test_full_size_layers_against_reference_kernels(N, K):
    """BASELINE configs[1] shapes (Llama-2-7B), device-generated packed layers: ours vs the reference kernels."""
    from qeft_b200 import qeft_cuda
    from qeft_b200.synth import synth_tensors
    t = synth_tensors(N, K, r=128, G=G, seed=N ^ K)
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    x1 = torch.randn((1, K), device="cuda", generator=g).half()
    a = ref.gemv_4bit_qeft(x1, t["qweight"], t["scales"], t["scaled_zeros"], t["oweight_interleaved"], 1, N, K, G)
    b = qeft_cuda.gemv_4bit_qeft(x1, t["qweight"], t["scales"], t["scaled_zeros"], t["oweight_interleaved"], 1, N, K, G)
    xm = torch.randn((512, K), device="cuda", generator=g).half()
    c = ref.gemm_4bit(xm, t["qweight"], t["scales"], t["scaled_zeros"]) + torch.nn.functional.linear(xm[..., -128:], t["oweight"])
    d = qeft_cuda.gemm_w4(xm, t["qweight"], t["scales"], t["scaled_zeros"], t["oweight"], None, group_size=G)
    torch.cuda.synchronize()
    e_v, e_m = rel(b.float().cpu().numpy(), a.float().cpu().numpy()), rel(d.float().cpu().numpy(), c.float().cpu().numpy())
    print(f"{N}x{K}: gemv ours-ref {e_v:.2e}; gemm(M=512) ours-ref {e_m:.2e}")
    assert e_v <= 4e-3 and e_m <= 4e-3
