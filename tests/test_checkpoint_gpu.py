"""The WCT update path on the device: a packed checkpoint written by the reference's save_model is loaded onto the GPU,
the fine-tuned outlier columns of a reference-written WCT checkpoint are installed (replace_oweight), and the decode
GEMV -- which reads the interleaved copy -- uses the NEW columns (the reference's replace_oweight,
qeft/utils/modelutils.py:185-198, leaves that copy stale).  Needs a B200."""
import os
import sys

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)
import tiny_model  # noqa: E402


def rel_err(got, want):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    return float(np.max(np.abs(got - want)) / max(np.max(np.abs(want)), 1e-6))


def layer_np(L):
    return {k: getattr(L, k).detach().cpu().numpy() for k in ("qweight", "scales", "scaled_zeros")}


@pytest.mark.parametrize("tokens", [1, 16])
def test_wct_checkpoint_reaches_the_kernels(tokens):
    from qeft_b200 import modelutils
    from qeft_b200.qlinear import QuantLinear
    ckpt = modelutils.load_checkpoint(os.path.join(GOLD, "reference_packed_ckpt.pth"))
    model = modelutils.hfmodel_to_owqmodel(tiny_model.build(1), ckpt, device="cuda:0")
    expect = np.load(os.path.join(GOLD, "reference_ckpt_expect.npz"))
    mods = dict(model.named_modules())
    rng = np.random.default_rng(tokens)
    xs, before = {}, {}
    for n in tiny_model.quant_layer_names():
        L = mods[n]
        assert isinstance(L, QuantLinear) and L.qweight.is_cuda
        x = rng.standard_normal((tokens, L.infeatures)).astype(np.float16)
        xs[n] = x
        y = L(torch.as_tensor(x).cuda())
        torch.cuda.synchronize()
        ids = L.reorder_ids.cpu().numpy() if "o_proj" in n else None
        bias = L.bias.cpu().numpy() if L.bias is not None else None
        want = oracle.forward(x, **layer_np(L), oweight=L.oweight.cpu().numpy(), bias=bias, reorder_ids=ids)
        assert rel_err(y.cpu().numpy(), want) <= 1e-3, n
        before[n] = y.cpu().numpy()
    modelutils.replace_oweight(model, modelutils.load_checkpoint(os.path.join(GOLD, "reference_wct")))
    for n in tiny_model.quant_layer_names():
        L = mods[n]
        new_ow = expect[n + ".oweight_finetuned"]
        assert np.array_equal(L.oweight.cpu().numpy().view(np.uint16), new_ow.view(np.uint16))
        y = L(torch.as_tensor(xs[n]).cuda())
        torch.cuda.synchronize()
        ids = L.reorder_ids.cpu().numpy() if "o_proj" in n else None
        bias = L.bias.cpu().numpy() if L.bias is not None else None
        want = oracle.forward(xs[n], **layer_np(L), oweight=new_ow, bias=bias, reorder_ids=ids)
        assert rel_err(y.cpu().numpy(), want) <= 1e-3, n                      # the kernels see the fine-tuned columns
        assert rel_err(y.cpu().numpy(), before[n].astype(np.float64)) > 1e-3   # and the result moved


def test_finetune_step_then_decode_sees_updated_columns():
    """prepare_for_finetune -> one SGD step on the outlier columns through autograd (QuantMatMulQEFT) ->
    refresh_oweight_interleaved -> decode GEMV uses the stepped columns; save_wctmodel round trip on the device."""
    from qeft_b200 import modelutils
    from qeft_b200.qlinear import QuantLinear
    ckpt = modelutils.load_checkpoint(os.path.join(GOLD, "reference_packed_ckpt.pth"))
    model = modelutils.hfmodel_to_owqmodel(tiny_model.build(1), ckpt, training=True, device="cuda:0")
    model = modelutils.prepare_for_finetune(model)
    L = dict(model.named_modules())["layers.0.self_attn.q_proj"]
    assert isinstance(L, QuantLinear) and L.oweight.requires_grad and L.oweight.dtype == torch.float32
    x = torch.randn(64, L.infeatures, device="cuda").half()
    ow0 = L.oweight.detach().clone()
    y = L(x)
    loss = (y.float() ** 2).mean()
    loss.backward()
    assert L.oweight.grad is not None and L.oweight.grad.shape == L.oweight.shape
    # gradient against autograd through the dense twin
    W = torch.as_tensor(oracle.dense_weight(**layer_np(L), oweight=ow0.half().cpu().numpy())).cuda()
    Wd = W.clone().requires_grad_(True)
    yd = torch.nn.functional.linear(x.float(), Wd)
    ((yd.half().float() ** 2).mean()).backward()
    g_ref = Wd.grad[:, -L.outlierfeatures:]
    assert rel_err(L.oweight.grad.cpu().numpy(), g_ref.cpu().numpy()) <= 2e-3
    with torch.no_grad():
        L.oweight -= 50.0 * L.oweight.grad
    L.refresh_oweight_interleaved()
    L.training = False
    x1 = torch.randn(1, L.infeatures, device="cuda").half()
    y1 = L(x1)
    torch.cuda.synchronize()
    want = oracle.forward(x1.cpu().numpy(), **layer_np(L), oweight=L.oweight.detach().half().cpu().numpy())
    assert rel_err(y1.cpu().numpy(), want) <= 1e-3
    stale = oracle.forward(x1.cpu().numpy(), **layer_np(L), oweight=ow0.half().cpu().numpy())
    assert rel_err(y1.cpu().numpy(), stale) > 1e-3
