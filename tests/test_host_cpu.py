"""Host-side logic and the C-ABI surface, no GPU needed."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import qeft_b200
from qeft_b200 import _lib
from qeft_b200.qlinear import QuantLinear, pack_intweight, pack_oweight, unpack_intweight
from qeft_b200.quant import make_quant

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "qeft_b200.h")).read()
    declared = set(re.findall(r"QEFT_API\s+[\w\s\*]+?\b(qeft_\w+)\s*\(", hdr))
    assert declared, "header parse failed"
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().qeft_abi_version() == 1
    assert b"sm_100a" in _lib.load().qeft_build_info()


def test_argument_validation_without_a_gpu():
    lib = _lib.load()
    dummy = ctypes.c_void_p(16)  # aligned non-null; never dereferenced because validation fails first
    # NULL pointers
    assert lib.qeft_gemv_w4(None, dummy, dummy, dummy, None, 0, None, None, dummy, 1, 64, 128, 0, 128, 0, None) == -1
    # batch outside 1..8 -> the reference's "Unsupported batch size" error
    st = lib.qeft_gemv_w4(dummy, dummy, dummy, dummy, None, 0, None, None, dummy, 9, 64, 128, 0, 128, 0, None)
    assert st == -3 and b"Unsupported batch size" in lib.qeft_status_string(st)
    # shape rules: N % 8, K % 64, r % 32
    assert lib.qeft_gemv_w4(dummy, dummy, dummy, dummy, None, 0, None, None, dummy, 1, 60, 128, 0, 128, 0, None) == -2
    assert lib.qeft_gemv_w4(dummy, dummy, dummy, dummy, None, 0, None, None, dummy, 1, 64, 100, 0, 100, 0, None) == -2
    assert lib.qeft_gemv_w4(dummy, dummy, dummy, dummy, dummy, 2, None, None, dummy, 1, 64, 128, 48, 128, 0, None) == -2
    assert lib.qeft_pack_w4(None, dummy, 8, 64, None) == -1
    assert lib.qeft_pack_w4(dummy, dummy, 6, 64, None) == -2
    misaligned = ctypes.c_void_p(8)
    assert lib.qeft_gemv_w4(misaligned, dummy, dummy, dummy, None, 0, None, None, dummy, 1, 64, 128, 0, 128, 0, None) == -5


def test_gather_entries_validate_before_touching_the_device():
    """qeft_gemm_w4_gather / qeft_gemv_w4_multi_gather: descriptor rules are checked before any CUDA call."""
    lib = _lib.load()
    dummy = ctypes.c_void_p(16)
    gemm = lambda g, N=256, y_ld=512: lib.qeft_gemm_w4_gather(  # noqa: E731
        dummy, dummy, dummy, dummy, None, None, 64, N, 128, 0, 128, _lib.DT_F16, 0, g, None)
    assert gemm(None) == -1                                   # no descriptor
    g = _lib.Gather()
    g.nranks, g.y_ld = 0, 512
    g.epoch = 16
    assert gemm(ctypes.byref(g)) == -2                        # ranks outside 1..QEFT_MAX_RANKS
    g.nranks = _lib.MAX_RANKS + 1
    assert gemm(ctypes.byref(g)) == -2
    g.nranks, g.y_ld = 2, 128
    assert gemm(ctypes.byref(g)) == -2                        # gathered row pitch smaller than this rank's N
    g.y_ld = 512
    g.y_peer[0][0] = 16
    g.done_peer[0] = 16
    assert gemm(ctypes.byref(g)) == -1                        # rank 1 has no buffer / counter
    g.y_peer[1][0] = 24
    g.done_peer[1] = 16
    assert gemm(ctypes.byref(g)) == -5                        # misaligned peer buffer
    g.y_peer[1][0] = 32
    g.epoch = None
    assert gemm(ctypes.byref(g)) == -2                        # the step counter is mandatory
    assert gemm(ctypes.byref(g), N=200) == -2                 # N % 128 (checked before the descriptor)
    part = (_lib.GemvPart * 1)()
    part[0] = _lib.GemvPart(16, 16, 16, None, None, None, 64)
    g.epoch = 16
    gemv = lambda gp, xg=None: lib.qeft_gemv_w4_multi_gather(dummy, part, 1, 0, xg, 1, 128, 0, 128, 0, gp, None)  # noqa: E731
    assert gemv(None) == -1
    assert gemv(ctypes.byref(g), ctypes.c_void_p(8)) == -5    # x_gather must be 16-byte aligned
    g.nranks = 9
    assert gemv(ctypes.byref(g)) == -2


def test_no_cpu_fallback():
    x = torch.zeros(1, 128, dtype=torch.float16)
    qw = torch.zeros(16, 128, dtype=torch.int16)
    s = torch.zeros(1, 64, dtype=torch.float16)
    with pytest.raises(RuntimeError, match="no CPU path"):
        qeft_b200.qeft_cuda.gemv_4bit(x, qw, s, s, 1, 64, 128, 128)
    with pytest.raises(RuntimeError, match="no CPU path"):
        qeft_b200.qeft_cuda.gemm_4bit(x, qw, s, s)


def test_product_packers_match_reference_vectors(golden):
    for idx in range(7):
        q = torch.tensor(golden[f"piw{idx}_q"].astype(np.int32))
        want = golden[f"piw{idx}_packed"]
        got = pack_intweight(q, interleave=4, kstride=64)
        assert got.dtype == torch.int16
        assert np.array_equal(got.numpy(), want)
        assert np.array_equal(unpack_intweight(torch.tensor(want)).numpy(), q.numpy())
    for idx in range(4):
        ow = torch.tensor(golden[f"pow{idx}_ow"])
        got = pack_oweight(ow, interleave=4)
        assert np.array_equal(got.numpy().view(np.uint16), golden[f"pow{idx}_packed"].view(np.uint16))


def test_sparse_to_dense_ids_matches_reference(golden):
    for idx in range(4):
        got = qeft_b200.sparse_to_dense_ids(torch.tensor(golden[f"s2d{idx}_ids"]), int(golden[f"s2d{idx}_K"]))
        assert got.dtype == torch.int64
        assert np.array_equal(got.numpy(), golden[f"s2d{idx}_dense"])


@pytest.mark.parametrize("idx", range(5))
def test_quantlinear_schema_and_pack_match_reference(golden, golden_schema, idx):
    c = golden_schema[f"qlp{idx}"]["case"]
    layer = QuantLinear(4, c["K"], c["N"], c["bias"], torch.float16, c["r"], c["G"], True, c["name"])
    got_schema = {k: [list(v.shape), str(v.dtype)] for k, v in layer.state_dict().items()}
    assert got_schema == golden_schema[f"qlp{idx}"]["state_dict"]
    p = f"qlp{idx}_"
    lin = torch.nn.Linear(c["K"], c["N"], bias=c["bias"])
    lin.weight.data = torch.tensor(golden[p + "weight"])
    if c["bias"]:
        lin.bias.data = torch.tensor(golden[p + "bias"])
    zeros = torch.tensor(golden[p + "zeros_in"]).clone()
    layer.pack(lin, torch.tensor(golden[p + "scales_in"]), zeros, torch.tensor(golden[p + "outlieridx"]), sym=c["sym"])
    assert np.array_equal(zeros.numpy(), golden[p + "zeros_after"])          # in-place +8 for sym, like the reference
    assert np.array_equal(layer.qweight.numpy(), golden[p + "qweight"])
    for k in ("scales", "scaled_zeros") + (("oweight", "oweight_interleaved") if c["r"] else ()):
        assert np.array_equal(getattr(layer, k).numpy().view(np.uint16), golden[p + k].view(np.uint16)), k
    if c["bias"]:
        assert np.array_equal(layer.bias.numpy().view(np.uint16), golden[p + "bias"].view(np.uint16))


def test_make_quant_schema_matches_reference(golden_schema):
    import types

    class Blk(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.q_proj = torch.nn.Linear(256, 32, bias=False).half()
            self.o_proj = torch.nn.Linear(256, 32, bias=False).half()
            self.other = torch.nn.Linear(8, 8)

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.layers = torch.nn.ModuleList([Blk(), Blk()])
            self.lm_head = torch.nn.Linear(32, 16)

    tiny = Tiny()
    infos = {f"layers.{i}.{n}": types.SimpleNamespace(bits=4, n_out=128, group_size=128, reorder=True)
             for i in range(2) for n in ("q_proj", "o_proj")}
    make_quant(tiny, infos)
    got = {k: [list(v.shape), str(v.dtype)] for k, v in tiny.state_dict().items()}
    assert got == golden_schema["make_quant"]
    assert tiny.layers[1].o_proj.name == "layers.1.o_proj"


def test_checkpoint_roundtrip_schema(tmp_path):
    from argparse import Namespace
    from qeft_b200 import modelutils

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.q_proj = torch.nn.Linear(256, 32, bias=False).half()
            self.dtype = torch.float16

    m = M()
    qinfo = {"q_proj": Namespace(bits=4, n_out=128, group_size=128, reorder=True, sym=False,
                                 scale_group=torch.rand(32, 2).half() * 0.01 + 0.002,
                                 zero_group=torch.randint(0, 16, (32, 2)).half(),
                                 out_ids=torch.arange(128, 256, dtype=torch.int32))}
    path = str(tmp_path / "ckpt" / "packed.pth")
    modelutils.save_model(m, qinfo, path, packing=True, fake=False)
    ck = modelutils.load_checkpoint(path)
    assert set(ck) == {"model_state_dict", "quantinfos", "packing", "dtype", "bits", "group_size"}
    assert ck["packing"] is True and ck["bits"] == 4 and ck["group_size"] == 128
    assert vars(ck["quantinfos"]["q_proj"]) == dict(bits=4, sym=False, group_size=128, n_out=128, reorder=True)
    m2 = M()
    make_quant(m2, ck["quantinfos"])
    missing, unexpected = m2.load_state_dict(ck["model_state_dict"], strict=False)
    assert not missing and not unexpected
    assert torch.equal(m2.q_proj.qweight, m.q_proj.qweight)


def test_dx_launch_plan_host_logic():
    """The dX launcher's contraction-split plan (host arithmetic, no GPU): which tiles are computed by one CTA and which
    are shared by 2-4 CTAs.  The expected picks are the ones measured fastest on B200 (profiles/r02_dx_tail_split.log,
    r02_dx_small_m_splits.log)."""
    from qeft_b200 import qeft_cuda
    plan = qeft_cuda.gemm_w4_dx_plan
    # Llama-2-7B at M = 2048: one wave of 128 tiles -> no split; down_proj (K = 11008): 344 tiles = 2 waves + 48 tiles x 3
    assert plan(2048, 4096, 4096) == (1, 128, 128)
    assert plan(2048, 11008, 4096) == (1, 128, 128)
    assert plan(2048, 4096, 11008) == (3, 296, 296 + 48 * 3)
    # Llama-2-13B (BASELINE.json configs[3]): K = 5120 is 160 tiles = 148 whole + 12 x 4; K = 13824 (432 tiles) is not split
    assert plan(2048, 5120, 5120) == (4, 148, 148 + 12 * 4)
    assert plan(2048, 13824, 5120) == (4, 148, 148 + 12 * 4)
    assert plan(2048, 5120, 13824) == (1, 432, 432)
    # launches of few tiles: every tile split
    assert plan(512, 4096, 4096) == (4, 0, 32 * 4)
    assert plan(128, 11008, 4096) == (4, 0, 16 * 4)
    assert plan(512, 4096, 11008) == (1, 86, 86)
    # the fp32 partials of a launch stay under 1 GB
    assert plan(16384, 13824, 5120)[0] * 16384 * 5120 * 4 <= 1 << 30
    # a split keeps at least 12 k-blocks (N >= 1536 for two splits)
    assert plan(256, 1024, 512)[0] == 1
    # invariants over a sweep: the grid covers every tile, a shared tile has 2-4 CTAs, whole tiles fill whole waves
    for M in (8, 100, 256, 640, 2048, 4096, 8192):
        for N in (128, 1536, 4096, 5120, 11008, 13824):
            for K in (256, 4096, 5120, 11008, 13824):
                for sms in (148, 132, 64):
                    sp, whole, ctas = plan(M, N, K, sms)
                    tiles = -(-M // 256) * -(-K // 256)
                    assert 1 <= sp <= 4 and 0 <= whole <= tiles and ctas == whole + (tiles - whole) * sp
                    if sp == 1:
                        assert whole == tiles
                    else:
                        assert (N // 64) // sp >= 12 and (whole == 0 or (whole % sms == 0 and whole < tiles))
    lib = _lib.load()
    import ctypes
    z = ctypes.c_int(0)
    assert lib.qeft_gemm_w4_dx_plan(2048, 4096, 4096, 148, None, ctypes.byref(z), ctypes.byref(z)) == -1
    assert lib.qeft_gemm_w4_dx_plan(2048, 4000, 4096, 148, ctypes.byref(z), ctypes.byref(z), ctypes.byref(z)) == -2


def test_nvtx_ranges_are_env_gated(monkeypatch):
    """QEFT_NVTX=1 wraps the dispatch target chosen by set_kernel in an NVTX range; by default forward is
    the bound method itself.  The wrapper keeps the name, passes errors through and closes its range."""
    import torch
    from qeft_b200.qlinear import QuantLinear

    def make():
        layer = QuantLinear(4, 256, 128, False, torch.float16, 64, 128, True, "model.layers.0.self_attn.o_proj")
        layer.outlieridx = torch.arange(64)
        layer.set_kernel(False)
        return layer

    monkeypatch.delenv("QEFT_NVTX", raising=False)
    plain = make()
    assert plain.forward == plain.forward_outlier_out_proj
    monkeypatch.setenv("QEFT_NVTX", "1")
    traced = make()
    assert traced.forward != traced.forward_outlier_out_proj and traced.forward.__name__ == "forward_outlier_out_proj"
    pushed, popped = [], []
    monkeypatch.setattr(torch.cuda.nvtx, "range_push", lambda s: pushed.append(s))
    monkeypatch.setattr(torch.cuda.nvtx, "range_pop", lambda: popped.append(1))
    try:
        traced(torch.zeros(1, 256, dtype=torch.float16))      # no CPU path: raises inside the range
        raise AssertionError("expected the CUDA-only error")
    except RuntimeError as e:
        assert "no CPU path" in str(e)
    assert pushed == ["QuantLinear.forward_outlier_out_proj:model.layers.0.self_attn.o_proj"] and popped == [1]


def test_packers_property_random_shapes():
    """Property test over random legal shapes (hypothesis): the product's closed-form host packers agree bit for bit with
    the oracle's restatement of qlinear.py:70-121 (itself pinned to reference outputs above), unpack inverts pack, and a
    one-nibble change moves exactly one nibble of the packed image (the layout is a permutation of nibbles)."""
    from hypothesis import given, settings, strategies as st
    import oracle

    @settings(max_examples=25, deadline=None)
    @given(st.integers(1, 4), st.integers(1, 5), st.integers(0, 2 ** 31 - 1))
    def check(n8, k64, seed):
        N, K = 8 * n8, 64 * k64                      # (pack_oweight interleaves blocks of 8 rows)
        rng = np.random.default_rng(seed)
        q = rng.integers(0, 16, size=(N, K), dtype=np.int32)
        packed = pack_intweight(torch.tensor(q), interleave=4, kstride=64)
        assert packed.shape == (N // 4, K) and packed.dtype == torch.int16
        assert np.array_equal(packed.numpy(), oracle.pack_intweight(q))
        assert np.array_equal(unpack_intweight(packed).numpy(), q)
        assert np.array_equal(oracle.unpack_intweight(packed.numpy()), q)
        # flip one weight: exactly one nibble of one int16 changes
        i, j = int(rng.integers(0, N)), int(rng.integers(0, K))
        q2 = q.copy()
        q2[i, j] ^= 0xF
        diff = packed.numpy().view(np.uint16) ^ pack_intweight(torch.tensor(q2), 4, 64).numpy().view(np.uint16)
        nz = diff[diff != 0]
        assert nz.size == 1 and int(nz[0]) in (0xF, 0xF0, 0xF00, 0xF000)
        # outlier columns: interleave is a permutation of fp16 values, inverted by the oracle's unpacker
        r = 32 * int(rng.integers(1, 5))
        ow = rng.standard_normal((N, r)).astype(np.float16)
        owi = pack_oweight(torch.tensor(ow), interleave=4)
        assert np.array_equal(owi.numpy().view(np.uint16), oracle.pack_oweight(ow).view(np.uint16))
        assert np.array_equal(oracle.unpack_oweight(owi.numpy()).view(np.uint16), ow.view(np.uint16))

    check()


def test_wrapper_validation_of_packed_operands():
    """What the C ABI cannot see (pointers only) is checked by the Python wrappers before a launch: qweight is the packed
    int16 [N / 4, K] image of THIS K, the scale tables have the shape the kernel indexes with group_size, dtypes are the
    checkpoint's, a caller-provided result tensor is exactly what the kernel writes."""
    import pytest
    import oracle
    from qeft_b200 import qeft_cuda
    N, K, r, G = 128, 256, 64, 128
    L = oracle.synth_layer(N, K, r=r, G=G, seed=1, bias=True)
    d = lambda a: torch.as_tensor(np.ascontiguousarray(a))  # noqa: E731
    qw, sc, sz, ow, b = d(L["qweight"]), d(L["scales"]), d(L["scaled_zeros"]), d(L["oweight"]), d(L["bias"])
    chk = qeft_cuda._check_packed
    chk("t", torch.float16, qw, sc, sz, ow, b, K, G)                       # the shapes every call site passes
    chk("t", torch.float16, qw[: N // 8], sc[:, : N // 2].contiguous(), sz[:, : N // 2].contiguous(), ow[: N // 2], None, K, G)   # a row shard
    with pytest.raises(RuntimeError, match="qweight"):
        chk("t", torch.float16, qw, sc, sz, ow, b, K // 2, G)              # activations of another width
    with pytest.raises(RuntimeError, match="qweight"):
        chk("t", torch.float16, qw.to(torch.int32), sc, sz, ow, b, K, G)
    with pytest.raises(RuntimeError, match="qweight"):
        chk("t", torch.float16, qw[:, ::2][:, : K // 2], sc, sz, ow, b, K // 2, G)   # non-contiguous view
    with pytest.raises(RuntimeError, match="scales"):
        chk("t", torch.float16, qw, sc, sz, ow, b, K, 64)                  # packed with another group size
    with pytest.raises(RuntimeError, match="Half"):
        chk("t", torch.float16, qw, sc.float(), sz, ow, b, K, G)
    with pytest.raises(RuntimeError, match="oweight"):
        chk("t", torch.bfloat16, qw, sc, sz, ow, b, K, G)                  # fp16 outlier block under bf16 activations
    out = torch.empty((8, N), dtype=torch.float16)
    qeft_cuda._check_out("t", out, 8 * N, torch.float16, out.device)
    for bad in (out.float(), out[:, : N // 2], torch.empty((4, N), dtype=torch.float16)):
        with pytest.raises(RuntimeError, match="out must be"):
            qeft_cuda._check_out("t", bad, 8 * N, torch.float16, out.device)
