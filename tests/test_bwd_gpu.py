"""Parity of the backward kernels (dX, dOW) and of the autograd function against the CPU oracle.  Needs a B200."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

REL_TOL = 1e-3


def rel_err(got, want):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    return float(np.max(np.abs(got - want)) / max(np.max(np.abs(want)), 1e-6))


def assert_close(got, want, what="", atol_rms=2e-3):
    """Elementwise form of the tolerance (beside the global-normalised rel_err): |err| <= 1e-3 |ref| + 2e-3 rms(ref).
    rtol covers a one-ulp flip of the fp16 result (2^-10); the rms term covers what does not scale with |y| (fp16
    rounding of the weights' fma, summation order), so small-magnitude outputs are checked too."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    rms = float(np.sqrt(np.mean(want ** 2))) + 1e-12
    bad = np.abs(got - want) > 1e-3 * np.abs(want) + atol_rms * rms
    assert not bad.any(), (what, int(bad.sum()), float(np.max(np.abs(got - want)) / rms))


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("M,N,K,r,G", [
    (64, 128, 256, 0, 128),         # no outlier columns (QuantMatMul)
    (128, 256, 512, 128, 128),
    (200, 384, 320, 64, 64),        # ragged tokens, K % 256 != 0 (feature tile is clipped), G = 64
    (300, 128, 1024, 192, 256),     # three token blocks, r = 192
    (512, 1024, 768, 128, 768),     # per-channel scales
    (256, 2048, 512, 128, 128),     # few tiles, long contraction: split over N (2 splits), fp32 partials + last-arriver sum
    (300, 3072, 320, 64, 64),       # 4 splits, ragged tokens and a clipped feature tile in the partial round trip
    (640, 4096, 1024, 128, 128),    # 3 token tiles x 4 feature tiles x 4 splits
])
def test_dx_matches_oracle(M, N, K, r, G):
    from qeft_b200 import qeft_cuda
    L = oracle.synth_layer(N, K, r=r, G=G, seed=M + N + K)
    rng = np.random.default_rng(M)
    dy = rng.standard_normal((M, N)).astype(np.float16)
    x = rng.standard_normal((M, K)).astype(np.float16)
    want_dx, _ = oracle.backward(dy, x, L["qweight"], L["scales"], L["scaled_zeros"], L.get("oweight"), group_size=G)
    got = qeft_cuda.gemm_w4_dx(dev(dy), dev(L["qweight"]), dev(L["scales"]), dev(L["scaled_zeros"]),
                               dev(L["oweight"]) if r > 0 else None, K, group_size=G, pdl=False)
    torch.cuda.synchronize()
    assert got.shape == (M, K) and got.dtype == torch.float16
    assert rel_err(got.cpu().numpy(), want_dx) <= REL_TOL, rel_err(got.cpu().numpy(), want_dx)
    assert_close(got.cpu().numpy(), want_dx)
    # a second launch gives the same bits (split launches: the arrival counters reset themselves, the partials are
    # added in split order)
    again = qeft_cuda.gemm_w4_dx(dev(dy), dev(L["qweight"]), dev(L["scales"]), dev(L["scaled_zeros"]),
                                 dev(L["oweight"]) if r > 0 else None, K, group_size=G, pdl=False)
    torch.cuda.synchronize()
    assert torch.equal(got, again)


@pytest.mark.parametrize("M,N,K,r", [(64, 128, 256, 64), (300, 256, 512, 128), (1000, 384, 384, 256)])
def test_dow_matches_oracle(M, N, K, r):
    from qeft_b200 import qeft_cuda
    L = oracle.synth_layer(N, K, r=r, seed=M + r)
    rng = np.random.default_rng(M)
    dy = rng.standard_normal((M, N)).astype(np.float16)
    x = rng.standard_normal((M, K)).astype(np.float16)
    _, want = oracle.backward(dy, x, L["qweight"], L["scales"], L["scaled_zeros"], L["oweight"])
    # full activation (pitched view of the last r columns) and the compact copy give the same bits
    got_full = qeft_cuda.dow(dev(dy), dev(x), r)
    got_compact = qeft_cuda.dow(dev(dy), dev(x[:, K - r:]), r)
    torch.cuda.synchronize()
    assert got_full.shape == (N, r) and got_full.dtype == torch.float32
    assert rel_err(got_full.cpu().numpy(), want) <= 1e-4
    assert torch.equal(got_full, got_compact)
    # accumulate adds into the fp32 master gradient
    acc = got_full.clone()
    qeft_cuda.dow(dev(dy), dev(x), r, out=acc, accumulate=True)
    torch.cuda.synchronize()
    assert rel_err(acc.cpu().numpy(), 2 * want) <= 1e-4


def test_autograd_function_matches_dense_autograd():
    """QuantMatMulQEFT: gradients equal torch.autograd through the dense dequantised weight."""
    from qeft_b200 import qeft_cuda
    from qeft_b200.synth import synth_quantlinear
    N, K, r, M = 256, 512, 128, 96
    layer = synth_quantlinear(N, K, r=r, seed=5, bias=True, training=True)
    layer.set_for_wct()
    layer.train()
    g = torch.Generator(device="cuda")
    g.manual_seed(3)
    x = torch.randn((2, M // 2, K), device="cuda", generator=g).half().requires_grad_(True)
    y = layer(x)
    gy = torch.randn(y.shape, device="cuda", generator=g).half()
    y.backward(gy)
    # dense reference on the GPU, fp32
    W = qeft_cuda.dequant_w4(layer.qweight, layer.scales, layer.scaled_zeros, None, 128).float()
    Wd = torch.cat([W[:, :K - r], layer.oweight.detach().float()], dim=1).requires_grad_(True)
    x2 = x.detach().float().requires_grad_(True)
    y2 = torch.nn.functional.linear(x2, Wd, layer.bias.float())
    y2.backward(gy.float())
    assert rel_err(y.detach().float().cpu().numpy(), y2.detach().cpu().numpy()) <= REL_TOL
    assert rel_err(x.grad.float().cpu().numpy(), x2.grad.cpu().numpy()) <= REL_TOL
    assert layer.oweight.grad is not None and layer.oweight.grad.dtype == torch.float32
    assert rel_err(layer.oweight.grad.cpu().numpy(), Wd.grad[:, K - r:].cpu().numpy()) <= REL_TOL
    assert layer.qweight.grad is None


@pytest.mark.parametrize("N,K", [(13824, 5120), (4096, 11008)])
def test_dx_full_size_tail_wave_split(N, K):
    """BASELINE.json config 4 shape (13B gate_proj) and the 7B down_proj at M = 2048: 160 / 344 tiles on 148 SMs, so the
    tiles of the last wave are computed by groups of CTAs that split the contraction (fp32 partials, last arriver adds
    in split order).  Sampled rows of every token tile against the oracle (whole rows: every feature tile) + determinism."""
    from qeft_b200 import qeft_cuda
    from qeft_b200.synth import synth_tensors, to_numpy_layer
    M = 2048
    t = synth_tensors(N, K, seed=4)
    g = torch.Generator(device="cuda")
    g.manual_seed(6)
    dy = torch.randn((M, N), device="cuda", generator=g).half()
    dx = qeft_cuda.gemm_w4_dx(dy, t["qweight"], t["scales"], t["scaled_zeros"], t["oweight"], K, pdl=False)
    dx2 = qeft_cuda.gemm_w4_dx(dy, t["qweight"], t["scales"], t["scaled_zeros"], t["oweight"], K, pdl=False)
    torch.cuda.synchronize()
    assert torch.equal(dx, dx2)
    L = to_numpy_layer(t, N, K, 128, 128)
    rows = np.array([0, 127, 128, 255, 256, 600, 900, 1100, 1300, 1600, 1800, 2047])
    W = oracle.dense_weight(L["qweight"], L["scales"], L["scaled_zeros"], L["oweight"]).astype(np.float64)
    want = dy[rows].cpu().numpy().astype(np.float64) @ W
    got = dx[rows].cpu().numpy()
    assert rel_err(got, want) <= REL_TOL
    # per feature tile of 256 columns (a wrong tile would hide behind a global maximum)
    for f0 in range(0, K, 256):
        assert rel_err(got[:, f0:f0 + 256], want[:, f0:f0 + 256]) <= 2 * REL_TOL, f0
