"""Checkpoint I/O against files WRITTEN BY THE REFERENCE (SURVEY.md 8c pin (3), 8f2): tests/golden/reference_packed_ckpt.pth
comes from the reference's save_model -> lm_pack -> QuantLinear.pack (qeft/utils/modelutils.py:219-268), and
tests/golden/reference_wct/model.pth from its save_wctmodel (:270-284); tests/golden/make_reference_checkpoint.py made both
by importing /root/reference in the build container.  CPU-only: loading, schema, the dense weights the packed layers stand
for, and the WCT update path (replace_oweight + refresh of the interleaved GEMV copy, which the reference forgets)."""
import os
import sys

import numpy as np
import pytest
import torch

import oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)
import tiny_model  # noqa: E402


@pytest.fixture(scope="module")
def expect():
    return np.load(os.path.join(GOLD, "reference_ckpt_expect.npz"))


def load_packed(device="cpu", training=False, seed=1):
    from qeft_b200 import modelutils
    ckpt = modelutils.load_checkpoint(os.path.join(GOLD, "reference_packed_ckpt.pth"))
    model = tiny_model.build(seed)              # other seed: every packed tensor must come from the file
    model = modelutils.hfmodel_to_owqmodel(model, ckpt, training=training, device=device)
    return model, ckpt


def test_reference_written_packed_checkpoint_schema_and_load(expect):
    from qeft_b200.qlinear import QuantLinear
    model, ckpt = load_packed()
    assert ckpt["packing"] is True and ckpt["bits"] == 4 and ckpt["group_size"] == tiny_model.GROUP
    assert ckpt["dtype"] == torch.float16
    names = tiny_model.quant_layer_names()
    assert sorted(ckpt["quantinfos"]) == sorted(names)
    info = ckpt["quantinfos"][names[0]]
    assert (info.bits, info.sym, info.group_size, info.n_out, info.reorder) == (4, False, 128, 128, True)
    mods = dict(model.named_modules())
    sd = ckpt["model_state_dict"]
    for n in names:
        layer = mods[n]
        assert isinstance(layer, QuantLinear)
        for key in ("qweight", "scales", "scaled_zeros", "oweight", "oweight_interleaved", "outlieridx"):
            got, want = getattr(layer, key), sd[f"{n}.{key}"]
            assert got.dtype == want.dtype and got.shape == want.shape, (n, key)
            assert torch.equal(got, want), (n, key)                       # byte-for-byte what the reference wrote
        if f"{n}.bias" in sd:
            assert torch.equal(layer.bias, sd[f"{n}.bias"])
        if "o_proj" in n:
            ids = oracle.sparse_to_dense_ids(expect[n + ".outlieridx"], layer.infeatures)
            assert np.array_equal(layer.reorder_ids.numpy(), ids)
    # everything that is not a packed layer is loaded as well (strict=False must not hide a schema mismatch)
    missing = [k for k in model.state_dict() if k not in sd and "reorder_ids" not in k]
    assert missing == [], missing


def test_reference_written_layers_stand_for_the_fake_quantised_weights(expect):
    """README "Result is Equal to Reconstruction": the packed layer dequantises to the fake-quantised weight the reference
    packed (up to the fp16 rounding of scaled_zeros = -(zero * scale), qlinear.py:207-208)."""
    model, _ = load_packed()
    mods = dict(model.named_modules())
    for n in tiny_model.quant_layer_names():
        L = mods[n]
        W = oracle.dense_weight(L.qweight.numpy(), L.scales.numpy(), L.scaled_zeros.numpy(), L.oweight.numpy(),
                                group_size=L.group_size)
        want = expect[n + ".fake_weight"].astype(np.float32)
        r = L.outlierfeatures
        assert np.array_equal(W[:, -r:], want[:, -r:])                    # outlier columns are stored exactly
        assert np.max(np.abs(W[:, :-r] - want[:, :-r])) <= 1.5e-4, n
        # the unpacked integers are valid 4-bit values and the dead columns hold the zero point
        q = oracle.unpack_intweight(L.qweight.numpy())
        assert q.min() >= 0 and q.max() <= 15


def test_reference_written_wct_checkpoint_replaces_and_refreshes(expect):
    from qeft_b200 import modelutils
    model, _ = load_packed()
    mods = dict(model.named_modules())
    before = {n: mods[n].oweight_interleaved.clone() for n in tiny_model.quant_layer_names()}
    wct = modelutils.load_checkpoint(os.path.join(GOLD, "reference_wct"))
    assert set(wct) == {"oweight_state_dict", "base_path"}
    assert sorted(wct["oweight_state_dict"]) == sorted(tiny_model.quant_layer_names())
    modelutils.replace_oweight(model, wct)
    for n in tiny_model.quant_layer_names():
        L = mods[n]
        want = expect[n + ".oweight_finetuned"]
        assert L.oweight.dtype == torch.float16
        assert np.array_equal(L.oweight.numpy().view(np.uint16), want.view(np.uint16)), n
        # the GEMV copy follows (the reference's replace_oweight, modelutils.py:185-198, leaves it stale)
        assert np.array_equal(L.oweight_interleaved.numpy().view(np.uint16), oracle.pack_oweight(want).view(np.uint16)), n
        assert not torch.equal(L.oweight_interleaved, before[n])


def test_our_wct_checkpoint_round_trip(tmp_path, expect):
    """save_wctmodel (ours) writes what the reference's writes; loading it back gives the same layers."""
    from qeft_b200 import modelutils
    from qeft_b200.qlinear import QuantLinear
    model, _ = load_packed()
    for m in model.modules():
        if isinstance(m, QuantLinear):
            m.set_for_wct()
            assert m.oweight.dtype == torch.float32 and m.oweight.requires_grad and not m.qweight.requires_grad
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, QuantLinear):
                m.oweight.add_(torch.randn(m.oweight.shape, generator=g) * 0.01)
    path = modelutils.save_wctmodel(model, os.path.join(GOLD, "reference_packed_ckpt.pth"), str(tmp_path / "wct"))
    ours = modelutils.load_checkpoint(path)
    ref = modelutils.load_checkpoint(os.path.join(GOLD, "reference_wct"))
    assert list(ours["oweight_state_dict"]) == list(ref["oweight_state_dict"])          # same keys, same order
    for k, v in ours["oweight_state_dict"].items():
        assert v.dtype == ref["oweight_state_dict"][k].dtype and v.shape == ref["oweight_state_dict"][k].shape
    fresh, _ = load_packed(seed=5)
    modelutils.replace_oweight(fresh, ours)
    a, b = dict(model.named_modules()), dict(fresh.named_modules())
    for n in tiny_model.quant_layer_names():
        assert torch.equal(a[n].oweight.detach().half(), b[n].oweight)
        assert torch.equal(b[n].oweight_interleaved, torch.as_tensor(oracle.pack_oweight(b[n].oweight.numpy())))


def test_prepare_for_finetune_hook():
    """The reference's get_training_model recipe (qeft/finetune.py:372-379, 452-470) as one call."""
    from qeft_b200 import modelutils
    from qeft_b200.qlinear import QuantLinear
    model, _ = load_packed(training=True)
    model = modelutils.prepare_for_finetune(model)
    trainable = [n for n, p in model.named_parameters() if p.requires_grad]
    assert trainable and all(n.endswith(".oweight") for n in trainable)
    assert len(trainable) == len(tiny_model.quant_layer_names())
    for n, p in model.named_parameters():
        if "oweight" in n:
            assert p.dtype == torch.float32
        if "norm" in n:
            assert p.dtype == torch.float32 and not p.requires_grad
    for m in model.modules():
        if isinstance(m, QuantLinear):
            assert m.training and m.matmul is not None
