"""The oracle against vectors produced by the reference's own Python (tests/golden/make_golden.py)."""
import numpy as np
import pytest

import oracle


def _cases(golden, prefix):
    idx = 0
    while f"{prefix}{idx}_" + ("q" if prefix == "piw" else "ow" if prefix == "pow" else "ids") in golden:
        yield idx
        idx += 1


def test_tile_index_map_matches_reference_probe(golden):
    e, i = oracle.tile_index_map()
    for j, kk, ge, gi in golden["piw_probe"]:
        assert (e[j, kk], i[j, kk]) == (ge, gi)


def test_pack_intweight_bit_exact(golden):
    n = 0
    for idx in _cases(golden, "piw"):
        q = golden[f"piw{idx}_q"].astype(np.int32)
        want = golden[f"piw{idx}_packed"]
        got = oracle.pack_intweight(q)
        assert got.dtype == np.int16 and got.shape == want.shape
        assert np.array_equal(got, want)
        assert np.array_equal(oracle.unpack_intweight(want), q)
        n += 1
    assert n == 7


def test_pack_oweight_bit_exact(golden):
    n = 0
    for idx in _cases(golden, "pow"):
        ow = golden[f"pow{idx}_ow"]
        want = golden[f"pow{idx}_packed"]
        got = oracle.pack_oweight(ow)
        assert np.array_equal(got.view(np.uint16), want.view(np.uint16))
        assert np.array_equal(oracle.unpack_oweight(want).view(np.uint16), ow.view(np.uint16))
        n += 1
    assert n == 4


def test_sparse_to_dense_ids(golden):
    for idx in range(4):
        got = oracle.sparse_to_dense_ids(golden[f"s2d{idx}_ids"], int(golden[f"s2d{idx}_K"]))
        assert np.array_equal(got, golden[f"s2d{idx}_dense"])


@pytest.mark.parametrize("idx", range(5))
def test_quantlinear_pack_matches_reference(golden, golden_schema, idx):
    c = golden_schema[f"qlp{idx}"]["case"]
    p = f"qlp{idx}_"
    res = oracle.quantize_for_pack(golden[p + "weight"], golden[p + "scales_in"], golden[p + "zeros_in"],
                                   c["r"], c["G"], sym=c["sym"])
    assert np.array_equal(oracle.pack_intweight(res["intweight"]), golden[p + "qweight"])
    assert np.array_equal(res["scales"].view(np.uint16), golden[p + "scales"].view(np.uint16))
    assert np.array_equal(res["scaled_zeros"].view(np.uint16), golden[p + "scaled_zeros"].view(np.uint16))
    if c["r"] > 0:
        assert np.array_equal(res["oweight"].view(np.uint16), golden[p + "oweight"].view(np.uint16))
        assert np.array_equal(oracle.pack_oweight(res["oweight"]).view(np.uint16),
                              golden[p + "oweight_interleaved"].view(np.uint16))
        # the int4 image of the outlier columns carries the zero point of their group (qlinear.py:200-202)
        q = oracle.unpack_intweight(golden[p + "qweight"])
        z = golden[p + "zeros_after"].astype(np.int32)
        K, G = c["K"], c["G"]
        cols = np.arange(K - c["r"], K)
        assert np.array_equal(q[:, K - c["r"]:], z[:, cols // G])
    if c["sym"]:
        assert np.array_equal(golden[p + "zeros_after"], golden[p + "zeros_in"] + 8)


def test_dequant_is_single_rounding_fma(golden, golden_schema):
    # dequant(pack(W)) reproduces W to within half a quantisation step on the int4 columns
    c = golden_schema["qlp0"]["case"]
    p = "qlp0_"
    W = oracle.dequant_weight(golden[p + "qweight"], golden[p + "scales"], golden[p + "scaled_zeros"], c["G"])
    w = golden[p + "weight"].astype(np.float32)
    K, r, G = c["K"], c["r"], c["G"]
    step = golden[p + "scales_in"].astype(np.float32)[:, np.arange(K - r) // G]
    assert np.all(np.abs(W[:, :K - r].astype(np.float32) - w[:, :K - r]) <= 0.5 * step + 2e-3 * np.abs(w[:, :K - r]) + 1e-4)
    # exhaustive check of the fma against python floats for one (s, sz) pair
    s = np.float16(0.01173); sz = np.float16(-0.0822)
    for qv in range(16):
        exact = float(qv) * float(s) + float(sz)
        qw = oracle.pack_intweight(np.full((4, 64), qv, dtype=np.int32))
        Wd = oracle.dequant_weight(qw, np.full((1, 4), s), np.full((1, 4), sz), 64)
        assert Wd[0, 0] == np.float16(exact)


def test_forward_semantics_and_backward_against_autograd():
    import torch
    L = oracle.synth_layer(32, 256, r=128, G=128, seed=3, bias=True)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((3, 256)).astype(np.float16)
    y_gemv = oracle.forward(x, L["qweight"], L["scales"], L["scaled_zeros"], L["oweight"], L["bias"])
    y_gemm = oracle.forward(x, L["qweight"], L["scales"], L["scaled_zeros"], L["oweight"], L["bias"], semantics="gemm")
    # the two reference code paths differ only by the tiny int4 residual of the outlier columns
    assert np.max(np.abs(y_gemv.astype(np.float32) - y_gemm.astype(np.float32))) < 5e-3
    W = torch.tensor(oracle.dense_weight(L["qweight"], L["scales"], L["scaled_zeros"], L["oweight"])).double()
    ow = W[:, -128:].clone().requires_grad_(True)
    xt = torch.tensor(x).double().requires_grad_(True)
    yt = torch.nn.functional.linear(xt, torch.cat([W[:, :-128], ow], dim=1))
    dy = rng.standard_normal((3, 32)).astype(np.float16)
    yt.backward(torch.tensor(dy).double())
    dx, dow = oracle.backward(dy, x, L["qweight"], L["scales"], L["scaled_zeros"], L["oweight"])
    assert np.allclose(dx.astype(np.float64), xt.grad.numpy(), rtol=2e-3, atol=2e-3)
    assert np.allclose(dow, ow.grad.numpy(), rtol=1e-5, atol=1e-5)


def test_algorithmic_bytes_match_baseline_md():
    assert oracle.gemv_algorithmic_bytes(4096, 4096) == 9_699_328
    assert oracle.gemv_algorithmic_bytes(11008, 4096) == 26_053_120
    assert oracle.gemv_algorithmic_bytes(4096, 11008) == 24_753_664
    per_layer = 4 * 9_699_328 + 2 * 26_053_120 + 24_753_664
    assert per_layer == 115_657_216


# ---- forward arithmetic: outputs of the reference's own CUDA kernels, run on a B200 ---------------------------------
# (tests/golden/make_reference_kernel_vectors.py; the kernels come from oracle/build_ref.py)
def _kernel_vectors():
    import os
    p = os.path.join(os.path.dirname(__file__), "golden", "reference_kernel_vectors.npz")
    if not os.path.exists(p):
        pytest.skip("tests/golden/reference_kernel_vectors.npz not generated yet")
    return np.load(p)


def test_oracle_forward_matches_reference_kernel_outputs():
    g = _kernel_vectors()
    cases = [str(c) for c in g["cases"]]
    assert len(cases) >= 9
    kinds = set()
    for i, c in enumerate(cases):
        kind, N, K, r, G, m = c.split(":")
        N, K, r, G, m = int(N), int(K), int(r), int(G), int(m)
        kinds.add((kind, G == K))
        ow = g[f"c{i}_oweight"] if r > 0 else None
        got = oracle.forward(g[f"c{i}_x"], g[f"c{i}_qweight"], g[f"c{i}_scales"], g[f"c{i}_scaled_zeros"], ow, None,
                             group_size=G)
        want = g[f"c{i}_y_ref"]
        assert got.shape == want.shape == (m, N) and want.dtype == np.float16
        err = np.max(np.abs(got.astype(np.float64) - want.astype(np.float64))) / np.max(np.abs(want.astype(np.float64)))
        # the reference rounds each dequantised weight to fp16 and keeps fp16 partial sums; the oracle accumulates wide
        assert err <= 4e-3, (c, err)
        if r > 0:
            # the interleaved outlier layout the GEMV kernel consumed is what the oracle's packer makes of `oweight`
            assert np.array_equal(oracle.pack_oweight(ow).view(np.uint16), g[f"c{i}_oweight_interleaved"].view(np.uint16))
    assert {("gemv_qeft", False), ("gemv_qeft", True), ("gemv", False), ("gemm", False)} <= kinds


def test_reference_kernel_vectors_detect_a_wrong_layout():
    """The pin has teeth: swapping two nibble positions or dropping the outlier override moves the oracle far outside
    the tolerance the real layout meets."""
    g = _kernel_vectors()
    i = 1                                  # gemv_qeft 128x512 r=128 m=2
    x, qw, s, z, ow = (g[f"c{i}_{k}"] for k in ("x", "qweight", "scales", "scaled_zeros", "oweight"))
    want = g[f"c{i}_y_ref"].astype(np.float64)
    rel = lambda y: np.max(np.abs(y.astype(np.float64) - want)) / np.max(np.abs(want))  # noqa: E731
    assert rel(oracle.forward(x, qw, s, z, ow)) <= 4e-3
    assert rel(oracle.forward(x, qw, s, z, None)) > 5e-2                       # int4 columns instead of outlier columns
    swapped = ((qw.view(np.uint16) >> 4) & 0x000f | (qw.view(np.uint16) << 4) & 0x00f0 | qw.view(np.uint16) & 0xff00).view(np.int16)
    assert rel(oracle.forward(x, swapped, s, z, ow)) > 5e-2                    # two nibbles of every word exchanged
