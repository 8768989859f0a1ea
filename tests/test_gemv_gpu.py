"""Parity of the CUDA decode path (through the C ABI) against the CPU oracle.  Needs a B200."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

REL_TOL = 1e-3   # north_star: "within max relative error 1e-3 (fp16 accumulate)"; ours accumulates in fp32


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


def run_gemv(L, x, layout, bias=None, gather=None, pdl=False):
    from qeft_b200 import _lib, qeft_cuda
    N, K, r, G = L["N"], L["K"], L["r"], L["G"]
    ow = None
    if r > 0:
        ow = dev(L["oweight_interleaved"] if layout == _lib.OW_INTERLEAVED else L["oweight"])
    y = qeft_cuda.gemv_w4(dev(x), dev(L["qweight"]), dev(L["scales"]), dev(L["scaled_zeros"]), ow, x.shape[0], N, K, G,
                          ow_layout=layout if r > 0 else _lib.OW_NONE, bias=None if bias is None else dev(bias),
                          x_gather=None if gather is None else dev(gather.astype(np.int32)), pdl=pdl)
    torch.cuda.synchronize()
    return y.cpu().numpy()


@pytest.mark.parametrize("N,K,r,G", [
    (16, 128, 0, 128), (8, 64, 0, 64), (24, 256, 32, 128), (32, 256, 128, 128), (40, 384, 64, 128),
    (64, 512, 96, 512), (128, 1024, 128, 128), (4096, 4096, 128, 128), (1024, 8192, 128, 128),
])
@pytest.mark.parametrize("m", [1, 2, 3, 7, 8])
def test_gemv_matches_oracle(N, K, r, G, m):
    from qeft_b200 import _lib
    if N >= 1024 and m not in (1, 7):
        pytest.skip("full-size shapes: m=1 and m=7 only")
    L = oracle.synth_layer(N, K, r=r, G=G, seed=N + K + r, bias=True)
    rng = np.random.default_rng(m)
    x = rng.standard_normal((m, K)).astype(np.float16)
    want = oracle.forward(x, L["qweight"], L["scales"], L["scaled_zeros"], L.get("oweight"), L["bias"], group_size=G)
    for layout in ([_lib.OW_INTERLEAVED, _lib.OW_PLAIN] if r > 0 else [_lib.OW_NONE]):
        got = run_gemv(L, x, layout, bias=L["bias"])
        assert got.shape == (m, N) and got.dtype == np.float16
        assert rel_err(got, want) <= REL_TOL, (layout, rel_err(got, want))
        assert_close(got, want)


@pytest.mark.parametrize("shape", [(11008, 4096), (4096, 11008)])
def test_gemv_llama7b_ffn_shapes(shape):
    from qeft_b200 import _lib
    N, K = shape
    L = oracle.synth_layer(N, K, seed=5)
    x = np.random.default_rng(1).standard_normal((1, K)).astype(np.float16)
    want = oracle.forward(x, L["qweight"], L["scales"], L["scaled_zeros"], L["oweight"], acc=np.float32)
    got = run_gemv(L, x, _lib.OW_INTERLEAVED, pdl=True)
    assert rel_err(got, want) <= REL_TOL


@pytest.mark.parametrize("N,K,r,m", [
    (12288, 256, 64, 8),     # more rows than one launch's partial-sum slices hold at m = 8: split into row windows
    (64, 8192, 128, 8),      # x too large to stage in shared memory: activations read through L1/L2
    (72, 1024, 288, 2),      # r > 256 (three outlier units per tile), N % 16 == 8, int8 path
    (256, 640, 96, 5),       # r % 64 != 0, K - r not a multiple of 128 (dead chunks in the last step)
])
def test_gemv_uncommon_shapes(N, K, r, m):
    from qeft_b200 import _lib
    L = oracle.synth_layer(N, K, r=r, G=128, seed=N + m, bias=True)
    x = np.random.default_rng(m).standard_normal((m, K)).astype(np.float16)
    want = oracle.forward(x, L["qweight"], L["scales"], L["scaled_zeros"], L["oweight"], L["bias"])
    for layout in (_lib.OW_INTERLEAVED, _lib.OW_PLAIN):
        got = run_gemv(L, x, layout, bias=L["bias"])
        assert rel_err(got, want) <= REL_TOL, (layout, rel_err(got, want))
        assert_close(got, want)


def test_gemv_o_proj_gather_fused():
    from qeft_b200 import _lib
    N, K, r = 256, 512, 128
    L = oracle.synth_layer(N, K, r=r, seed=11, o_proj=True)
    ids = oracle.sparse_to_dense_ids(L["outlieridx"], K)
    x = np.random.default_rng(2).standard_normal((3, K)).astype(np.float16)
    want = oracle.forward(x, L["qweight"], L["scales"], L["scaled_zeros"], L["oweight"], reorder_ids=ids)
    got = run_gemv(L, x, _lib.OW_INTERLEAVED, gather=ids)
    assert rel_err(got, want) <= REL_TOL
    # and it is exactly the un-fused path on pre-gathered input
    got2 = run_gemv(L, np.take(x, ids, axis=-1), _lib.OW_INTERLEAVED)
    assert np.array_equal(got.view(np.uint16), got2.view(np.uint16))


def test_outlier_columns_replace_int4_columns():
    """GEMV semantics: the int4 image of the last r columns must not leak into y (gemv_cuda_qeft.cu:168-176)."""
    from qeft_b200 import _lib
    L = oracle.synth_layer(64, 512, r=128, seed=3)
    x = np.random.default_rng(3).standard_normal((1, 512)).astype(np.float16)
    a = run_gemv(L, x, _lib.OW_INTERLEAVED)
    q = L["intweight"].copy()
    q[:, -128:] = 15 - q[:, -128:]          # scramble the dead columns
    L2 = dict(L, qweight=oracle.pack_intweight(q))
    b = run_gemv(L2, x, _lib.OW_INTERLEAVED)
    assert np.array_equal(a.view(np.uint16), b.view(np.uint16))


def test_gemv_multi_equals_single_launches_bitwise():
    from qeft_b200 import _lib, qeft_cuda
    K, r, G, m = 1024, 128, 128, 2
    Ls = [oracle.synth_layer(N, K, r=r, seed=20 + i) for i, N in enumerate((256, 64, 72))]
    x = np.random.default_rng(4).standard_normal((m, K)).astype(np.float16)
    xd = dev(x)
    parts, singles = [], []
    for L in Ls:
        t = {k: dev(L[k]) for k in ("qweight", "scales", "scaled_zeros", "oweight_interleaved")}
        parts.append({"qweight": t["qweight"], "scales": t["scales"], "scaled_zeros": t["scaled_zeros"],
                      "oweight": t["oweight_interleaved"], "N": L["N"]})
        singles.append(qeft_cuda.gemv_4bit_qeft(xd, t["qweight"], t["scales"], t["scaled_zeros"],
                                               t["oweight_interleaved"], m, L["N"], K, G))
    outs = qeft_cuda.gemv_w4_multi(xd, parts, m, K, r, G, ow_layout=_lib.OW_INTERLEAVED)
    torch.cuda.synchronize()
    for a, b, L in zip(outs, singles, Ls):
        assert torch.equal(a, b)
        want = oracle.forward(x, L["qweight"], L["scales"], L["scaled_zeros"], L["oweight"])
        assert rel_err(a.cpu().numpy(), want) <= REL_TOL


def test_reference_signature_errors():
    from qeft_b200 import qeft_cuda
    x = torch.zeros(9, 128, dtype=torch.float16, device="cuda")
    qw = torch.zeros(16, 128, dtype=torch.int16, device="cuda")
    s = torch.zeros(1, 64, dtype=torch.float16, device="cuda")
    with pytest.raises(RuntimeError, match="Unsupported batch size"):
        qeft_cuda.gemv_4bit(x, qw, s, s, 9, 64, 128, 128)
    with pytest.raises(RuntimeError, match="Half"):
        qeft_cuda.gemv_4bit(x[:1].float(), qw, s, s, 1, 64, 128, 128)


def test_device_packers_bit_exact():
    from qeft_b200 import qeft_cuda
    rng = np.random.default_rng(9)
    for N, K in [(8, 64), (64, 256), (4096, 4096)]:
        q = rng.integers(0, 16, size=(N, K), dtype=np.int32)
        packed = qeft_cuda.pack_w4(dev(q))
        assert np.array_equal(packed.cpu().numpy(), oracle.pack_intweight(q))
        assert np.array_equal(qeft_cuda.unpack_w4(packed).cpu().numpy(), q)
    ow = rng.standard_normal((64, 128)).astype(np.float16)
    assert np.array_equal(qeft_cuda.interleave_oweight(dev(ow)).cpu().numpy().view(np.uint16),
                          oracle.pack_oweight(ow).view(np.uint16))
    assert np.array_equal(qeft_cuda.interleave_oweight(dev(ow.astype(np.float32))).cpu().numpy().view(np.uint16),
                          oracle.pack_oweight(ow).view(np.uint16))
    L = oracle.synth_layer(64, 512, r=128, seed=1)
    W = qeft_cuda.dequant_w4(dev(L["qweight"]), dev(L["scales"]), dev(L["scaled_zeros"]), dev(L["oweight"])).cpu().numpy()
    want = oracle.dense_weight(L["qweight"], L["scales"], L["scaled_zeros"], L["oweight"]).astype(np.float16)
    assert np.array_equal(W.view(np.uint16), want.view(np.uint16))   # dequant is bit-exact (single-rounding fma)


def test_quantlinear_module_decode_and_graph_capture():
    from qeft_b200.synth import synth_quantlinear, to_numpy_layer
    layer = synth_quantlinear(512, 1024, name="model.layers.0.self_attn.o_proj", bias=True, seed=4)
    t = {k: getattr(layer, k) for k in ("qweight", "scales", "scaled_zeros", "oweight", "bias")}
    L = to_numpy_layer(t, 512, 1024, 128, 128)
    x = torch.randn(1, 1, 1024, device="cuda").half()
    y = layer(x)
    want = oracle.forward(x.cpu().numpy(), L["qweight"], L["scales"], L["scaled_zeros"], L["oweight"], L["bias"],
                          reorder_ids=layer.reorder_ids.cpu().numpy())
    assert y.shape == (1, 1, 512)
    assert rel_err(y.cpu().numpy(), want) <= REL_TOL
    # the launch is capturable (current stream, no allocation inside the C ABI)
    g = torch.cuda.CUDAGraph()
    static_x = x.clone()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        layer(static_x)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        static_y = layer(static_x)
    static_x.copy_(torch.randn_like(static_x))
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(static_y, layer(static_x))


def test_decode_stack_graph_equals_eager_and_oracle():
    from qeft_b200.decode import PackedDecoderStack
    from qeft_b200.synth import to_numpy_layer
    st = PackedDecoderStack("7b", layers=2, seed=1)
    eager = {n: v.clone() for n, v in st.out[1].items()}
    st.step_eager()
    torch.cuda.synchronize()
    eager = {n: v.clone() for n, v in st.out[1].items()}
    st.capture()
    for o in st.out:
        for v in o.values():
            v.zero_()
    st.step()
    torch.cuda.synchronize()
    for n, v in st.out[1].items():
        assert torch.equal(v, eager[n]), n
    # one projection of each input width against the oracle
    for n, x in (("k", st.x_h), ("down", st.x_f), ("o", st.x_h)):
        t = st.blocks[1][n]
        N, K = t["N"], t["qweight"].shape[1]
        L = to_numpy_layer({k: t[k] for k in ("qweight", "scales", "scaled_zeros", "oweight")}, N, K, 128, 128)
        ids = t["reorder_ids32"].cpu().numpy() if n == "o" else None
        want = oracle.forward(x.cpu().numpy(), L["qweight"], L["scales"], L["scaled_zeros"], L["oweight"],
                              reorder_ids=ids, acc=np.float32)
        assert rel_err(st.out[1][n].cpu().numpy(), want) <= REL_TOL, n
    assert st.algorithmic_bytes_per_step() == 2 * 115_657_216
