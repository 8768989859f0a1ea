"""The literal drop-in claim: the reference's OWN qeft/qlinear.py, UNMODIFIED (staged under baseline/_ref by
oracle/build_ref.stage_reference_python(); /root/reference does not exist on the GPU box), runs with
``sys.modules["qeft_cuda"] = qeft_b200.qeft_cuda`` -- its `import qeft_cuda` (qlinear.py:8-11) then binds this repo's
shim -- and its QuantLinear forwards (qlinear.py:244-330: decode GEMV, prefill GEMM + outlier F.linear, the o_proj
variant with index_select) give what the oracle says and what this repo's QuantLinear gives.  Needs a B200."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGE = os.path.join(ROOT, "baseline", "_ref")


def rel_err(got, want):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    return float(np.max(np.abs(got - want)) / max(np.max(np.abs(want)), 1e-6))


@pytest.fixture(scope="module")
def ref_qlinear():
    if not os.path.exists(os.path.join(STAGE, "qeft", "qlinear.py")):
        pytest.skip("baseline/_ref/qeft/qlinear.py not staged (run __graft_entry__.build() where /root/reference exists)")
    from qeft_b200 import qeft_cuda as shim
    saved = {k: sys.modules.get(k) for k in ("qeft_cuda", "qeft", "qeft.qlinear", "qeft.reorder")}
    sys.modules["qeft_cuda"] = shim
    sys.path.insert(0, STAGE)
    for k in ("qeft", "qeft.qlinear", "qeft.reorder"):
        sys.modules.pop(k, None)
    mod = importlib.import_module("qeft.qlinear")
    assert mod.qeft_cuda is shim
    assert os.path.abspath(mod.__file__).startswith(os.path.abspath(STAGE))
    yield mod
    sys.path.remove(STAGE)
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


def fill(layer, L, device="cuda"):
    for k in ("qweight", "scales", "scaled_zeros", "oweight", "oweight_interleaved", "outlieridx", "bias"):
        if k in L:
            setattr(layer, k, torch.as_tensor(np.ascontiguousarray(L[k])))
    return layer.to(device)


CASES = [
    dict(N=256, K=512, r=128, bias=True, name="model.layers.0.self_attn.q_proj"),
    dict(N=128, K=1024, r=128, bias=False, name="model.layers.0.self_attn.o_proj"),
    dict(N=384, K=256, r=0, bias=False, name="model.layers.0.mlp.gate_proj"),
    dict(N=4096, K=4096, r=128, bias=False, name="model.layers.3.self_attn.o_proj"),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c['N']}x{c['K']}r{c['r']}{'_o' if 'o_proj' in c['name'] else ''}")
@pytest.mark.parametrize("tokens", [1, 5, 8, 200])
def test_reference_quantlinear_on_our_shim(ref_qlinear, case, tokens):
    from qeft_b200.qlinear import QuantLinear as Ours
    N, K, r = case["N"], case["K"], case["r"]
    if N >= 4096 and tokens not in (1, 200):
        pytest.skip("full size: one decode and one prefill case")
    o_proj = "o_proj" in case["name"]
    L = oracle.synth_layer(N, K, r=r, seed=N + K + r, bias=case["bias"], o_proj=o_proj)
    x = np.random.default_rng(tokens).standard_normal((1, tokens, K)).astype(np.float16)
    ref_layer = fill(ref_qlinear.QuantLinear(4, K, N, case["bias"], torch.float16, r, 128, True, case["name"]), L)
    ref_layer.set_kernel(False)
    our_layer = fill(Ours(4, K, N, case["bias"], torch.float16, r, 128, True, case["name"]), L)
    our_layer.set_kernel(False)
    xd = torch.as_tensor(x).cuda()
    y_ref = ref_layer(xd)
    y_ours = our_layer(xd)
    torch.cuda.synchronize()
    assert y_ref.shape == (1, tokens, N) and y_ref.dtype == torch.float16
    ids = oracle.sparse_to_dense_ids(L["outlieridx"], K) if o_proj and r > 0 else None
    # the reference's GEMM path keeps the int4 image of the outlier columns and ADDS the dense columns (qlinear.py:264-266)
    sem = "gemv" if tokens < 8 else "gemm"
    want = oracle.forward(x, L["qweight"], L["scales"], L["scaled_zeros"], L.get("oweight"), L.get("bias"),
                          semantics=sem, reorder_ids=ids)
    # reference module: fp16 `y += F.linear(...)` and `y + bias` round two more times than the oracle's single rounding
    assert rel_err(y_ref.cpu().numpy(), want) <= 2e-3, rel_err(y_ref.cpu().numpy(), want)
    want_ours = oracle.forward(x, L["qweight"], L["scales"], L["scaled_zeros"], L.get("oweight"), L.get("bias"),
                               reorder_ids=ids)
    assert rel_err(y_ours.cpu().numpy(), want_ours) <= 1e-3
    assert rel_err(y_ref.cpu().numpy(), y_ours.cpu().numpy().astype(np.float64)) <= 2e-3


def test_reference_module_error_behaviour(ref_qlinear):
    """Same failures as the reference's extension: non-Half input raises, batch outside 1..7 on the GEMV raises."""
    from qeft_b200 import qeft_cuda as shim
    L = oracle.synth_layer(64, 256, r=128, seed=2)
    t = {k: torch.as_tensor(np.ascontiguousarray(L[k])).cuda() for k in ("qweight", "scales", "scaled_zeros", "oweight_interleaved")}
    x = torch.randn(1, 256, device="cuda")
    with pytest.raises(RuntimeError, match="Half"):
        shim.gemv_4bit_qeft(x, t["qweight"], t["scales"], t["scaled_zeros"], t["oweight_interleaved"], 1, 64, 256, 128)
    x9 = torch.randn(9, 256, device="cuda").half()
    with pytest.raises(RuntimeError, match="Unsupported batch size"):
        shim.gemv_4bit_qeft(x9, t["qweight"], t["scales"], t["scaled_zeros"], t["oweight_interleaved"], 9, 64, 256, 128)
