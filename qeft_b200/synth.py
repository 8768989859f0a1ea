"""Synthetic packed layers with the statistics SURVEY.md 8(d) prescribes, generated on the device.

There is no network for real checkpoints; bench.py, smoke() and the full-size GPU tests use these.
"""
from __future__ import annotations

import torch

from . import qeft_cuda
from .qlinear import QuantLinear
from .reorder import sparse_to_dense_ids

LLAMA_SHAPES = {
    # name: (hidden, ffn, layers, kv_out)
    "7b": (4096, 11008, 32, 4096),
    "13b": (5120, 13824, 40, 5120),
    "70b": (8192, 28672, 80, 1024),
}


def decoder_linears(model: str):
    """[(name, N, K)] of one decoder block, reference naming (q/k/v/o/gate/up/down_proj)."""
    h, f, _, kv = LLAMA_SHAPES[model]
    return [("self_attn.q_proj", h, h), ("self_attn.k_proj", kv, h), ("self_attn.v_proj", kv, h),
            ("self_attn.o_proj", h, h), ("mlp.gate_proj", f, h), ("mlp.up_proj", f, h), ("mlp.down_proj", h, f)]


@torch.no_grad()
def synth_tensors(N, K, r=128, G=128, seed=0, device="cuda", bias=False, o_proj=False, fast=False):
    """Packed tensors of one layer: q ~ U{0..15}, scale ~ U(0.002, 0.012), zero ~ U{0..15}, oweight ~ N(0, 0.02^2).

    ``fast``: draw the packed int16 words directly (uniform random nibbles, which is what packing uniform q gives; the
    dead outlier columns then hold random nibbles instead of the zero point, and no kernel reads them).  Used by the
    benchmark stacks, where generating and packing 35 GB of int32 weights would dominate the run."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    ng = K // G
    zero = torch.randint(0, 16, (ng, N), dtype=torch.int32, device=device, generator=gen)
    scale = (torch.rand((ng, N), device=device, generator=gen) * 0.010 + 0.002).half()
    if fast:
        qweight = torch.randint(-32768, 32768, (N // 4, K), dtype=torch.int16, device=device, generator=gen)
        q = None
    else:
        q = torch.randint(0, 16, (N, K), dtype=torch.int32, device=device, generator=gen)
        if r > 0:
            cols = torch.arange(K - r, K, device=device)
            q[:, K - r:] = zero[cols // G, :].t()
        qweight = qeft_cuda.pack_w4(q)
    out = {
        "qweight": qweight,
        "scales": scale,
        "scaled_zeros": (-(zero.float() * scale.float())).half(),
    }
    del q
    if r > 0:
        ow = (torch.randn((N, r), device=device, generator=gen) * 0.02).half()
        out["oweight"] = ow
        out["oweight_interleaved"] = qeft_cuda.interleave_oweight(ow)
        if o_proj:
            idx = torch.randperm(K, device=device, generator=gen)[:r].sort().values.to(torch.int32)
        else:
            idx = torch.arange(K - r, K, device=device, dtype=torch.int32)
        out["outlieridx"] = idx
    if bias:
        out["bias"] = (torch.randn((N,), device=device, generator=gen) * 0.02).half()
    return out


@torch.no_grad()
def synth_quantlinear(N, K, r=128, G=128, seed=0, device="cuda", bias=False, name="model.layers.0.self_attn.q_proj",
                      training=False) -> QuantLinear:
    t = synth_tensors(N, K, r, G, seed, device, bias, o_proj=("o_proj" in name))
    layer = QuantLinear(4, K, N, bias, torch.float16, r, G, True, name).to(device)
    for k, v in t.items():
        setattr(layer, k, v)
    layer.set_kernel(training)
    return layer


def to_numpy_layer(t, N, K, r, G):
    """Device tensors of `synth_tensors` -> the dict layout oracle.* functions take."""
    d = {"N": N, "K": K, "r": r, "G": G}
    for k, v in t.items():
        d[k] = v.detach().cpu().numpy()
    return d
