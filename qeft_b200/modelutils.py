"""Checkpoint schema of the packed model and of the fine-tuned outlier columns ("WCT").

Byte-compatible with the reference (qeft/utils/modelutils.py:120-145, 185-198, 219-284):

* packed checkpoint: ``{'model_state_dict', 'quantinfos': {name: Namespace(bits, sym, group_size, n_out,
  reorder)}, 'packing': True, 'dtype', 'bits', 'group_size'}``
* WCT checkpoint: ``{'oweight_state_dict': {layer_name: tensor}, 'base_path'}``

Loading needs ``weights_only=False`` on torch >= 2.6 because of the Namespace objects.
"""
from __future__ import annotations

import os
from argparse import Namespace
from collections import OrderedDict

import torch

from .qlinear import QuantLinear
from .quant import find_layers, lm_pack, make_quant


def load_checkpoint(path):
    if os.path.isdir(path):
        path = os.path.join(path, "model.pth")
    return torch.load(path, map_location="cpu", weights_only=False)


def hfmodel_to_owqmodel(model, ckpt, training=False, device="cuda:0"):
    """Turn a freshly constructed (HF-style) model into the packed model described by ``ckpt``."""
    if ckpt["packing"]:
        make_quant(model, ckpt["quantinfos"])
        model.load_state_dict(ckpt["model_state_dict"], strict=False)
        if device not in ("cpu", None):
            model = model.to(device)
        for layer in find_layers(model, [QuantLinear]).values():
            layer.set_kernel(training)
    else:
        model.load_state_dict(ckpt["model_state_dict"], strict=False)
        if device not in ("cpu", None):
            model = model.to(device)
    return model


def packed_checkpoint(model, quantizers, dtype=None):
    """The dict the reference's ``save_model(packing=True)`` writes, for an already packed ``model``."""
    first = next(iter(quantizers.values()))
    infos = {n: Namespace(bits=q.bits, sym=getattr(q, "sym", False), group_size=getattr(q, "group_size", -1),
                          n_out=getattr(q, "n_out", 0), reorder=getattr(q, "reorder", False))
             for n, q in quantizers.items()}
    return {"model_state_dict": model.state_dict(), "quantinfos": infos, "packing": True,
            "dtype": dtype if dtype is not None else getattr(model, "dtype", torch.float16),
            "bits": first.bits, "group_size": getattr(first, "group_size", -1)}


def save_model(model, quantizers, save_path, packing=True, fake=False):
    if fake:
        raise NotImplementedError("fake-quant checkpoints are an offline-quantiser artefact (out of scope)")
    os.makedirs(os.path.dirname(os.path.abspath(save_path)), exist_ok=True)
    lm_pack(model, quantizers)
    torch.save(packed_checkpoint(model, quantizers), save_path)


def save_wctmodel(model, base_path, output_dir):
    """Fine-tuned outlier columns only (reference: modelutils.py:270-284)."""
    sd = OrderedDict()
    dtype = getattr(model, "dtype", torch.float16)
    for name, param in model.named_parameters():
        if "oweight" in name:
            sd[name.replace(".oweight", "")] = param.data.to(dtype)
    os.makedirs(output_dir, exist_ok=True)
    path = os.path.join(output_dir, "model.pth")
    torch.save({"oweight_state_dict": sd, "base_path": os.path.abspath(base_path)}, path)
    return path


def replace_oweight(model, ckpt_wct):
    """Install fine-tuned outlier columns AND refresh the interleaved GEMV copy (the reference only does
    the first, modelutils.py:185-198, so its decode path keeps using the pre-fine-tuning columns)."""
    sd = ckpt_wct["oweight_state_dict"]
    for name, module in model.named_modules():
        if isinstance(module, QuantLinear) and name in sd:
            module.oweight.data = sd[name].data.to(device=module.oweight.device, dtype=module.oweight.dtype)
            module.refresh_oweight_interleaved()
    return model


def prepare_for_finetune(model, gradient_checkpointing=False):
    """Put a packed model into the state the reference fine-tunes it in -- what ``get_training_model`` does around
    ``QuantLinear`` (qeft/finetune.py:372-379, 452-470, with peft's ``prepare_model_for_kbit_training`` :292-356):

    * every ``QuantLinear``: ``set_kernel(training=True)`` (autograd path, ``QuantMatMulQEFT``) and ``set_for_wct()``
      (packed weight frozen, outlier columns an fp32 trainable parameter);
    * every other parameter frozen; parameters of modules whose name contains ``norm`` cast to fp32;
    * only parameters with ``oweight`` in their name train;
    * inputs of the embedding require grad when the model offers ``enable_input_require_grads`` (so that gradient
      checkpointing has a differentiable path), and gradient checkpointing is switched on on request.
    Returns the model; ``save_wctmodel`` then writes only the fine-tuned columns."""
    for module in model.modules():
        if isinstance(module, QuantLinear):
            if not module.training or module.matmul is None:
                module.set_kernel(training=True)
            if not isinstance(module.oweight if module.outlierfeatures > 0 else None, torch.nn.Parameter):
                module.set_for_wct()
    for name, param in model.named_parameters():
        param.requires_grad = False
    for name, module in model.named_modules():
        if "norm" in name:
            module.to(torch.float32)
    for name, param in model.named_parameters():
        if "oweight" in name:
            param.requires_grad = True
    if hasattr(model, "enable_input_require_grads"):
        model.enable_input_require_grads()
    if gradient_checkpointing and hasattr(model, "gradient_checkpointing_enable"):
        model.gradient_checkpointing_enable()
    return model


# --------------------------------------------------------------------------------------------------
# column (output-feature) sharding of one packed layer -- SURVEY.md 8(e)
# --------------------------------------------------------------------------------------------------
def shard_rows(N, rank, world, multiple=128):
    """Rows ``[lo, hi)`` of an N-row layer owned by ``rank``: contiguous, equal, a multiple of ``multiple`` each
    (128 for the tensor-core tiles; the GEMV alone needs 8)."""
    if N % (world * multiple) != 0:
        raise ValueError(f"N={N} does not split into {world} shards of a multiple of {multiple} rows")
    per = N // world
    return rank * per, (rank + 1) * per


def shard_layer_tensors(t, rank, world, multiple=128):
    """The rank's slice of a packed layer's tensors (dict with the checkpoint's buffer names).

    No repacking: the interleave groups are 4 rows (qweight) / 8 rows (oweight_interleaved), so a row range that is
    a multiple of 8 is a contiguous slab of every buffer.  x is replicated; rank p computes ``y[:, lo:hi]``;
    the slices are concatenated along the feature dimension by one all-gather per launch group."""
    N = t["qweight"].shape[0] * 4
    lo, hi = shard_rows(N, rank, world, multiple)
    out = {"qweight": t["qweight"][lo // 4:hi // 4].contiguous(),
           "scales": t["scales"][:, lo:hi].contiguous(),
           "scaled_zeros": t["scaled_zeros"][:, lo:hi].contiguous()}
    if t.get("bias") is not None:
        out["bias"] = t["bias"][lo:hi].contiguous()
    if t.get("oweight") is not None:
        out["oweight"] = t["oweight"][lo:hi].contiguous()
    if t.get("oweight_interleaved") is not None:
        out["oweight_interleaved"] = t["oweight_interleaved"][lo // 2:hi // 2].contiguous()
    if t.get("outlieridx") is not None:
        out["outlieridx"] = t["outlieridx"]
    return out
