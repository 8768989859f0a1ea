"""Outlier-guided reordering (OGR), host side: the index helper ``QuantLinear.set_kernel`` uses for o_proj and the
producer that moves the global outlier channels of a decoder stack to the END of every hidden-size dimension, so that
a packed layer's dense fp16 outlier columns are its last ``r`` input columns.

Reference: qeft/reorder.py (``sparse_to_dense_ids`` :6-12, ``reorder_embeds`` :14-25, ``reorder_qkv_ffn1_ln`` :27-38,
``reorder_out`` :40-67, ``reorder_in_mlp`` :105-145, ``make_reorder`` :148-176) and the selection of the global
indices at the end of qeft/extract_outidx.py (:159-179).  Same function names, argument meaning and in-place effect on the
modules / quantizers; the tensor work is one helper (``_take``) instead of the reference's repeated blocks.  This is
offline model surgery on the host: no kernel is involved.
"""
from typing import Iterable, List, Sequence

import torch


def sparse_to_dense_ids(sparse_ids: torch.Tensor, length: int) -> torch.Tensor:
    """Permutation that moves the columns listed in ``sparse_ids`` to the end, keeping the rest in order."""
    if not len(sparse_ids) < length:
        raise AssertionError("outlier index list must be shorter than the feature dimension")
    ids = sparse_ids.to(torch.long)
    keep = torch.ones(length, dtype=torch.bool, device=ids.device)
    keep[ids] = False
    rest = torch.nonzero(keep, as_tuple=False).flatten()
    return torch.cat([rest, ids])


def select_global_outlier_ids(h_diags: Iterable[torch.Tensor], target_rank: int) -> List[int]:
    """The global OGR indices: every q/k/v and gate/up layer contributes its Hessian diagonal normalised by its mean;
    the ``target_rank`` input channels with the largest summed sensitivity, ascending (extract_outidx.py:159-179)."""
    total = None
    for h in h_diags:
        s = h / h.mean()
        total = s.clone() if total is None else total + s
    if total is None:
        raise ValueError("no sensitivities given")
    return sorted(torch.topk(total, target_rank).indices.cpu().tolist())


def _take(t: torch.Tensor, dim: int, ids: torch.Tensor) -> torch.Tensor:
    return torch.index_select(t, dim, ids.to(t.device))


def _reorder_linear_rows(layers: Sequence, quantizers: Sequence, ids: torch.Tensor) -> None:
    """Permute the OUTPUT channels of linear layers: weight rows, bias, and the per-row quantisation parameters."""
    for layer in layers:
        layer.weight.data = _take(layer.weight.data, 0, ids)
        if getattr(layer, "bias", None) is not None:
            layer.bias.data = _take(layer.bias.data, 0, ids)
    for q in quantizers:
        for grouped, plain in (("scale_group", "scale"), ("zero_group", "zero")):
            name = grouped if hasattr(q, grouped) else plain
            setattr(q, name, _take(getattr(q, name), 0, ids))


def reorder_embeds(l_pres: Sequence, l_posts: Sequence, out_ids: torch.Tensor) -> None:
    """Embeddings (hidden size is their LAST dimension) and the final norm / lm_head (reference :14-25)."""
    dst = sparse_to_dense_ids(out_ids, l_pres[0].weight.data.shape[1])
    for layer in l_pres:
        layer.weight.data = _take(layer.weight.data, 1, dst)
    for layer in l_posts:
        layer.weight.data = _take(layer.weight.data, -1, dst)
        if getattr(layer, "bias", None) is not None:
            layer.bias.data = _take(layer.bias.data, -1, dst)


def reorder_qkv_ffn1_ln(l_qkv_ffn1: Sequence, l_ln: Sequence, out_ids: torch.Tensor) -> None:
    """Input channels of q/k/v and gate/up, and the norms that feed them (reference :27-38)."""
    dst = sparse_to_dense_ids(out_ids, l_qkv_ffn1[0].weight.shape[-1])
    for layer in l_qkv_ffn1:
        layer.weight.data = _take(layer.weight.data, 1, dst)
    for norm in l_ln:
        norm.weight.data = _take(norm.weight.data, 0, dst)
        if getattr(norm, "bias", None) is not None:
            norm.bias.data = _take(norm.bias.data, 0, dst)


def reorder_out(l_out: Sequence, l_out_quantizers: Sequence, out_ids: torch.Tensor) -> None:
    """o_proj: its INPUT channels by the layer's own outlier ids (kept on the module as ``reorder_ids``: the gather
    ``QuantLinear`` applies to x at run time, qlinear.py:273-275), its output channels by the global ids (reference :40-67)."""
    out_ch, in_ch = l_out[0].weight.shape
    if l_out_quantizers[0].out_ids.numel() > 0:
        dst = sparse_to_dense_ids(l_out_quantizers[0].out_ids, in_ch)
        for layer in l_out:
            layer.weight.data = _take(layer.weight.data, 1, dst)
            layer.reorder_ids = dst
    _reorder_linear_rows(l_out, l_out_quantizers, sparse_to_dense_ids(out_ids, out_ch))


def reorder_in_mlp(l_ffn1: Sequence, l_ffn2: Sequence, l_ffn1_quantizers: Sequence, l_ffn2_quantizers: Sequence) -> None:
    """Inside the MLP: gate/up output channels and down_proj input channels by down_proj's own outlier ids, then
    down_proj's output channels by gate/up's (= the global) ids (reference :105-145)."""
    dst = sparse_to_dense_ids(l_ffn2_quantizers[0].out_ids, l_ffn2[0].weight.shape[-1])
    _reorder_linear_rows(l_ffn1, l_ffn1_quantizers, dst)
    for layer in l_ffn2:
        layer.weight.data = _take(layer.weight.data, 1, dst)
    dst = sparse_to_dense_ids(l_ffn1_quantizers[0].out_ids, l_ffn2[0].weight.shape[0])
    _reorder_linear_rows(l_ffn2, l_ffn2_quantizers, dst)


def make_reorder(blocks: Sequence[dict], quantizers: Sequence[dict], pre_layers: Sequence, post_layers: Sequence,
                 global_ids: torch.Tensor) -> None:
    """The whole pass (reference :148-176) over an explicit description of the model instead of HF module parsing:
    ``blocks[i]`` = {"ln": [attn_norm, mlp_norm], "qkv": [...], "out": [...], "ffn1": [...], "ffn2": [...]} (modules),
    ``quantizers[i]`` = the same keys without "ln" (objects with ``out_ids`` and scale / zero tensors)."""
    reorder_embeds(pre_layers, post_layers, global_ids)
    for blk, qz in zip(blocks, quantizers):
        reorder_qkv_ffn1_ln(l_qkv_ffn1=list(blk["qkv"]) + list(blk["ffn1"]), l_ln=blk["ln"], out_ids=global_ids)
        reorder_out(l_out=blk["out"], l_out_quantizers=qz["out"], out_ids=global_ids)
        reorder_in_mlp(l_ffn1=blk["ffn1"], l_ffn2=blk["ffn2"], l_ffn1_quantizers=qz["ffn1"], l_ffn2_quantizers=qz["ffn2"])
