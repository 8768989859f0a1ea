"""Parity of the tcgen05 prefill / fine-tune GEMM (through the C ABI) against the CPU oracle.  Needs a B200."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

REL_TOL = 1e-3   # north_star: max relative error 1e-3 for the fp16 path (vs the reference dequant + matmul)


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


def run_gemm(L, x, bias=None, with_outliers=True):
    from qeft_b200 import qeft_cuda
    ow = dev(L["oweight"]) if (with_outliers and L["r"] > 0) else None
    y = qeft_cuda.gemm_w4(dev(x), dev(L["qweight"]), dev(L["scales"]), dev(L["scaled_zeros"]), ow,
                          None if bias is None else dev(bias), group_size=L["G"], pdl=False)
    torch.cuda.synchronize()
    return y.cpu().numpy()


@pytest.mark.parametrize("M,N,K,r,G", [
    (8, 128, 128, 0, 128),          # smallest legal: one k-block pair, one 128-feature block
    (128, 256, 256, 64, 128),       # both feature blocks, one outlier k-block
    (77, 128, 512, 128, 128),       # ragged token count (TMA zero-fills, stores are masked)
    (300, 384, 1024, 128, 128),     # N % 256 == 128: last tile has one feature block; three token tiles
    (256, 256, 512, 0, 512),        # per-channel scales (G == K)
    (130, 512, 768, 192, 256),      # G = 256, r = 192
])
def test_gemm_matches_oracle(M, N, K, r, G):
    L = oracle.synth_layer(N, K, r=r, G=G, seed=M + N + K + r, bias=True)
    x = np.random.default_rng(M).standard_normal((M, K)).astype(np.float16)
    want = oracle.forward(x, L["qweight"], L["scales"], L["scaled_zeros"], L.get("oweight"), L["bias"], group_size=G)
    got = run_gemm(L, x, bias=L["bias"])
    assert got.shape == (M, N) and got.dtype == np.float16
    assert rel_err(got, want) <= REL_TOL, rel_err(got, want)
    assert_close(got, want)


@pytest.mark.parametrize("M,N,K,r,G", [
    (8, 256, 2048, 128, 128),       # 64-token tile, 2 feature blocks, 17 ring stages: K split over several CTAs per tile
    (24, 384, 4096, 128, 128),
    (48, 128, 1024, 64, 128),       # odd number of k-blocks (15 + 1): the last split ends on a half-filled stage
    (64, 512, 1536, 0, 512),        # per-channel scales, no outlier columns
    (100, 256, 2048, 128, 256),     # 128-token tile
    (128, 128, 8192, 128, 128),
    (300, 256, 4096, 128, 128),     # 256-token tiles, few of them: split as well
])
def test_small_m_split_k_matches_oracle(M, N, K, r, G):
    """8 <= M <= 128 (the reference's split-K tile configurations, gemm_cuda.cu:952-1004): narrow token tile, K split over
    gridDim.z CTAs per tile, fp32 partials added in split order by the last CTA to arrive.  Deterministic: two launches
    are bit-equal; the arrival counters are left at zero, so a second launch (and a CUDA-graph replay) works."""
    L = oracle.synth_layer(N, K, r=r, G=G, seed=M + N + K + r, bias=True)
    x = np.random.default_rng(M).standard_normal((M, K)).astype(np.float16)
    want = oracle.forward(x, L["qweight"], L["scales"], L["scaled_zeros"], L.get("oweight"), L["bias"], group_size=G)
    got = run_gemm(L, x, bias=L["bias"])
    assert got.shape == (M, N) and got.dtype == np.float16
    assert rel_err(got, want) <= REL_TOL, rel_err(got, want)
    assert_close(got, want)
    again = run_gemm(L, x, bias=L["bias"])
    assert np.array_equal(got.view(np.int16), again.view(np.int16))


def test_small_m_split_k_under_graph_replay():
    from qeft_b200 import qeft_cuda
    M, N, K, r = 16, 512, 4096, 128
    L = oracle.synth_layer(N, K, r=r, seed=5)
    x = dev(np.random.default_rng(5).standard_normal((M, K)).astype(np.float16))
    t = {k: dev(L[k]) for k in ("qweight", "scales", "scaled_zeros", "oweight")}
    y = torch.empty((M, N), dtype=torch.float16, device="cuda")
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):          # warm-up on the capture stream: the split-K workspace is allocated here
        qeft_cuda.gemm_w4(x, t["qweight"], t["scales"], t["scaled_zeros"], t["oweight"], None, out=y)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    first = y.clone()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        qeft_cuda.gemm_w4(x, t["qweight"], t["scales"], t["scaled_zeros"], t["oweight"], None, out=y)
    for _ in range(3):
        y.zero_()
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(y.view(torch.int16), first.view(torch.int16))
    want = oracle.forward(x.cpu().numpy(), L["qweight"], L["scales"], L["scaled_zeros"], L["oweight"], None)
    assert rel_err(y.cpu().numpy(), want) <= REL_TOL


def test_gemm_reference_signature_uses_all_int4_columns():
    """`gemm_4bit` of the reference has no outlier term: every int4 column counts (gemm_cuda.cu:929-1033)."""
    from qeft_b200 import qeft_cuda
    L = oracle.synth_layer(256, 512, r=128, seed=9)
    x = np.random.default_rng(9).standard_normal((64, 512)).astype(np.float16)
    W = oracle.dequant_weight(L["qweight"], L["scales"], L["scaled_zeros"]).astype(np.float64)
    want = (x.astype(np.float64) @ W.T).astype(np.float16)
    got = qeft_cuda.gemm_4bit(dev(x), dev(L["qweight"]), dev(L["scales"]), dev(L["scaled_zeros"]))
    torch.cuda.synchronize()
    assert rel_err(got.cpu().numpy(), want) <= REL_TOL


def test_gemm_equals_gemv_on_the_same_rows():
    """Prefill and decode agree: the GEMM's rows match the GEMV of the same activations within tolerance."""
    from qeft_b200 import _lib, qeft_cuda
    N, K, r = 512, 1024, 128
    L = oracle.synth_layer(N, K, r=r, seed=21, bias=True)
    x = np.random.default_rng(21).standard_normal((16, K)).astype(np.float16)
    yg = run_gemm(L, x, bias=L["bias"])
    xd = dev(x[:4])
    yv = qeft_cuda.gemv_w4(xd, dev(L["qweight"]), dev(L["scales"]), dev(L["scaled_zeros"]), dev(L["oweight_interleaved"]),
                           4, N, K, 128, ow_layout=_lib.OW_INTERLEAVED, bias=dev(L["bias"]))
    torch.cuda.synchronize()
    assert rel_err(yg[:4], yv.cpu().numpy()) <= REL_TOL


@pytest.mark.parametrize("shape", [(4096, 4096), (11008, 4096), (4096, 11008)])
def test_gemm_llama7b_shapes_m2048_properties(shape):
    """Full size (BASELINE.json config 3): checked by linearity against sampled oracle rows/columns."""
    from qeft_b200 import qeft_cuda
    from qeft_b200.synth import synth_tensors, to_numpy_layer
    N, K = shape
    M = 2048
    t = synth_tensors(N, K, seed=3)
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    x = torch.randn((M, K), device="cuda", generator=g).half()
    y = qeft_cuda.gemm_w4(x, t["qweight"], t["scales"], t["scaled_zeros"], t["oweight"], None, pdl=False)
    torch.cuda.synchronize()
    L = to_numpy_layer(t, N, K, 128, 128)
    rows = np.array([0, 1, 127, 128, 1000, 2047])
    want = oracle.forward(x[rows].cpu().numpy(), L["qweight"], L["scales"], L["scaled_zeros"], L["oweight"])
    assert rel_err(y[rows].cpu().numpy(), want) <= REL_TOL
    # linearity in x: gemm(2 x) == 2 gemm(x) exactly in fp16 (power-of-two scaling commutes with every rounding
    # of a NORMAL fp16 result; subnormal results round on an absolute grid and are excluded)
    y2 = qeft_cuda.gemm_w4((x * 2).half(), t["qweight"], t["scales"], t["scaled_zeros"], t["oweight"], None, pdl=False)
    torch.cuda.synchronize()
    normal = y.float().abs() >= 2.0 ** -13
    assert torch.equal(y2[normal], (y.float() * 2).half()[normal])
    # and the kernel is deterministic
    y3 = qeft_cuda.gemm_w4(x, t["qweight"], t["scales"], t["scaled_zeros"], t["oweight"], None, pdl=False)
    torch.cuda.synchronize()
    assert torch.equal(y, y3)


def _bf16_round(a):
    """fp32 -> bf16 (round to nearest even) -> fp32, in numpy."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


@pytest.mark.parametrize("M,N,K,r", [(96, 256, 512, 128), (300, 384, 1024, 64)])
def test_gemm_and_backward_bf16(M, N, K, r):
    """bf16 activations / outputs (north_star tolerance 1e-2): int4 columns dequantised in fp32 with one rounding to
    bf16, outlier columns and activations bf16, fp32 accumulation."""
    from qeft_b200 import qeft_cuda
    L = oracle.synth_layer(N, K, r=r, seed=M + r, bias=True)
    rng = np.random.default_rng(M)
    x = _bf16_round(rng.standard_normal((M, K)).astype(np.float32))
    dy = _bf16_round(rng.standard_normal((M, N)).astype(np.float32))
    q = oracle.unpack_intweight(L["qweight"]).astype(np.float64)
    s = np.repeat(L["scales"].astype(np.float64).T, 128, axis=1)
    z = np.repeat(L["scaled_zeros"].astype(np.float64).T, 128, axis=1)
    W = _bf16_round((q * s + z).astype(np.float32)).astype(np.float64)       # fma in (at least) fp32, one rounding
    ow = _bf16_round(L["oweight"].astype(np.float32))
    W[:, K - r:] = ow
    want_y = x.astype(np.float64) @ W.T + L["bias"].astype(np.float64)
    want_dx = dy.astype(np.float64) @ W
    want_dw = dy.astype(np.float64).T @ x[:, K - r:].astype(np.float64)
    tb = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda().to(torch.bfloat16)  # noqa: E731  (exact: already bf16 values)
    y = qeft_cuda.gemm_w4(tb(x), dev(L["qweight"]), dev(L["scales"]), dev(L["scaled_zeros"]), tb(ow), dev(L["bias"]), pdl=False)
    dx = qeft_cuda.gemm_w4_dx(tb(dy), dev(L["qweight"]), dev(L["scales"]), dev(L["scaled_zeros"]), tb(ow), K, pdl=False)
    dw = qeft_cuda.dow(tb(dy), tb(x), r)
    torch.cuda.synchronize()
    assert y.dtype == torch.bfloat16 and dx.dtype == torch.bfloat16 and dw.dtype == torch.float32
    assert rel_err(y.float().cpu().numpy(), want_y) <= 1e-2
    assert rel_err(dx.float().cpu().numpy(), want_dx) <= 1e-2
    assert rel_err(dw.cpu().numpy(), want_dw) <= 1e-4
