"""Module surgery that creates / fills ``QuantLinear`` layers (reference: qeft/quant.py:194-234)."""
from __future__ import annotations

import torch.nn as nn

from .qlinear import QuantLinear


def find_layers(module, layers=(nn.Conv2d, nn.Linear), name=""):
    """Dotted-name -> module for every leaf of one of the given types (reference: utils/misc.py:8-16)."""
    if type(module) in tuple(layers):
        return {name: module}
    found = {}
    for child_name, child in module.named_children():
        found.update(find_layers(child, layers, f"{name}.{child_name}" if name else child_name))
    return found


def make_quant(module, quantinfos, name=""):
    """Replace every linear whose dotted name is a key of ``quantinfos`` by an empty ``QuantLinear``.

    ``quantinfos[name]`` carries ``bits`` and optionally ``n_out`` (outlier columns), ``group_size`` and
    ``reorder`` -- the Namespace objects stored in a packed checkpoint (utils/modelutils.py:253-259).
    """
    if isinstance(module, QuantLinear):
        return
    for attr, child in list(module.named_children()):
        full = f"{name}.{attr}" if name else attr
        if full in quantinfos:
            info = quantinfos[full]
            qlayer = QuantLinear(info.bits, child.in_features, child.out_features, child.bias is not None,
                                 child.weight.dtype, getattr(info, "n_out", 0), getattr(info, "group_size", -1),
                                 getattr(info, "reorder", False), full)
            setattr(module, attr, qlayer.to(child.weight.device))
        else:
            make_quant(child, quantinfos, full)


def lm_pack(model, quantinfos, linears=(nn.Linear,)):
    """Swap in ``QuantLinear`` layers and pack each from its fake-quantised linear + quantizer state."""
    layers = find_layers(model, linears)
    layers = {n: layers[n] for n in quantinfos}
    make_quant(model, quantinfos)
    qlayers = find_layers(model, [QuantLinear])
    for name, qlayer in qlayers.items():
        info = quantinfos[name] = quantinfos[name].cpu() if hasattr(quantinfos[name], "cpu") else quantinfos[name]
        qlayer.pack(layers[name],
                    scales=getattr(info, "scale_group", getattr(info, "scale", None)),
                    zeros=getattr(info, "zero_group", getattr(info, "zero", None)),
                    outlieridx=getattr(info, "out_ids", None),
                    sym=getattr(info, "sym", False))
    return model
