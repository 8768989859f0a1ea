"""A seeded toy decoder stack (plain nn modules + quantizer stand-ins) for the OGR reorder tests: hidden 32, ffn 48,
2 blocks; every layer carries distinct random weights so that any wrong permutation shows."""
import types

import numpy as np
import torch
import torch.nn as nn

H, F, V, NBLK, R = 32, 48, 16, 2, 4


def _lin(g, n_out, n_in, bias):
    layer = nn.Linear(n_in, n_out, bias=bias)
    layer.weight.data = torch.randn(n_out, n_in, generator=g)
    if bias:
        layer.bias.data = torch.randn(n_out, generator=g)
    return layer


def _quantizer(g, n_out, out_ids, grouped):
    q = types.SimpleNamespace(out_ids=out_ids)
    if grouped:
        q.scale_group = torch.rand(n_out, 3, generator=g)
        q.zero_group = torch.rand(n_out, 3, generator=g)
    else:
        q.scale = torch.rand(n_out, 1, generator=g)
        q.zero = torch.rand(n_out, 1, generator=g)
    return q


def build(seed):
    g = torch.Generator().manual_seed(seed)
    global_ids = torch.sort(torch.randperm(H, generator=g)[:R]).values
    pre = [nn.Embedding(V, H)]
    pre[0].weight.data = torch.randn(V, H, generator=g)
    norm_f = nn.LayerNorm(H)
    norm_f.weight.data = torch.randn(H, generator=g)
    norm_f.bias.data = torch.randn(H, generator=g)
    post = [norm_f, _lin(g, V, H, False)]
    blocks, quantizers = [], []
    for _ in range(NBLK):
        ln = []
        for _ in range(2):
            n = nn.LayerNorm(H)
            n.weight.data = torch.randn(H, generator=g)
            n.bias.data = torch.randn(H, generator=g)
            ln.append(n)
        blk = {"ln": ln, "qkv": [_lin(g, H, H, True) for _ in range(3)], "out": [_lin(g, H, H, True)],
               "ffn1": [_lin(g, F, H, False) for _ in range(2)], "ffn2": [_lin(g, H, F, True)]}
        out_ids_o = torch.sort(torch.randperm(H, generator=g)[:R]).values       # o_proj's own outlier input channels
        out_ids_d = torch.sort(torch.randperm(F, generator=g)[:R]).values       # down_proj's
        qz = {"qkv": [_quantizer(g, H, global_ids, True) for _ in range(3)],
              "out": [_quantizer(g, H, out_ids_o, True)],
              "ffn1": [_quantizer(g, F, global_ids, bool(seed % 2)) for _ in range(2)],
              "ffn2": [_quantizer(g, H, out_ids_d, False)]}
        blocks.append(blk)
        quantizers.append(qz)
    return {"global_ids": global_ids, "pre": pre, "post": post, "blocks": blocks, "quantizers": quantizers}


def snapshot(m):
    out = {}
    for i, layer in enumerate(m["pre"] + m["post"]):
        out[f"io{i}/w"] = layer.weight.data.numpy().copy()
        if getattr(layer, "bias", None) is not None:
            out[f"io{i}/b"] = layer.bias.data.numpy().copy()
    for bi, (blk, qz) in enumerate(zip(m["blocks"], m["quantizers"])):
        for key, layers in blk.items():
            for li, layer in enumerate(layers):
                out[f"b{bi}/{key}{li}/w"] = layer.weight.data.numpy().copy()
                if getattr(layer, "bias", None) is not None:
                    out[f"b{bi}/{key}{li}/b"] = layer.bias.data.numpy().copy()
                if hasattr(layer, "reorder_ids"):
                    out[f"b{bi}/{key}{li}/reorder_ids"] = layer.reorder_ids.numpy().astype(np.int64)
        for key, qs in qz.items():
            for qi, q in enumerate(qs):
                for name in ("scale_group", "zero_group", "scale", "zero"):
                    if hasattr(q, name):
                        out[f"b{bi}/q_{key}{qi}/{name}"] = getattr(q, name).numpy().copy()
    return out
