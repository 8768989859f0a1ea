"""OGR producer (qeft_b200/reorder.py) against the reference's own qeft/reorder.py: the reference functions were run on a
seeded toy decoder stack in the build container (tests/golden/make_reorder_golden.py); here the same stack is rebuilt from
the seed, reordered by OUR functions and compared bit for bit -- weights, biases, norm parameters, per-row quantisation
parameters and the o_proj ``reorder_ids``.  Plus the property the reordering exists for: the network computes the same
function, with its hidden channels permuted so that the global outlier channels are last."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from tiny_reorder_model import H, R, build, snapshot  # noqa: E402

from qeft_b200 import reorder  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "reference_reorder.npz"))


def _describe(m):
    return m["blocks"], m["quantizers"], m["pre"], m["post"], m["global_ids"]


def test_make_reorder_matches_the_reference_bit_for_bit():
    for seed in (1, 2):
        m = build(seed)
        reorder.make_reorder(*_describe(m))
        snap = snapshot(m)
        keys = [k[len(f"s{seed}/"):] for k in GOLD.files if k.startswith(f"s{seed}/") and not k.endswith("outidx")]
        assert sorted(keys) == sorted(snap.keys())
        for k in keys:
            assert np.array_equal(snap[k], GOLD[f"s{seed}/{k}"]), (seed, k)


def test_select_global_outlier_ids_matches_the_reference_selection():
    for seed in (1, 2):
        g = torch.Generator().manual_seed(100 + seed)
        hs = [torch.rand(48, generator=g) + 0.01 for _ in range(5)]
        got = reorder.select_global_outlier_ids(hs, 6)
        assert got == GOLD[f"s{seed}/outidx"].tolist()
        assert got == sorted(got) and len(set(got)) == 6


def _forward(m, tokens):
    """Toy decoder forward through the modules as they are (no attention mixing: per-token MLP stack is enough to
    exercise every reordered dimension): embed -> blocks -> final norm -> head."""
    x = m["pre"][0](tokens)
    for blk in m["blocks"]:
        h = blk["ln"][0](x)
        q, k, v = (layer(h) for layer in blk["qkv"])
        a = torch.tanh(q) * torch.sigmoid(k) + v                       # elementwise stand-in for attention
        o = blk["out"][0]
        a = a if not hasattr(o, "reorder_ids") else torch.index_select(a, -1, o.reorder_ids)
        x = x + o(a)
        h = blk["ln"][1](x)
        gate, up = (layer(h) for layer in blk["ffn1"])
        x = x + blk["ffn2"][0](torch.nn.functional.silu(gate) * up)
    return m["post"][1](m["post"][0](x))


def test_reordered_network_computes_the_same_function():
    torch.manual_seed(0)
    tokens = torch.randint(0, 16, (5, 7))
    for seed in (1, 2):
        m = build(seed)
        with torch.no_grad():
            want = _forward(m, tokens)
        gids = m["global_ids"].clone()
        reorder.make_reorder(*_describe(m))
        # LayerNorm statistics are permutation invariant; the normalised shape stays H
        with torch.no_grad():
            got = _forward(m, tokens)
        assert torch.allclose(got, want, rtol=1e-4, atol=1e-4), float((got - want).abs().max())
        # the hidden dimension now ends with the global outlier channels: the embedding's last R columns are the
        # original columns gids
        m0 = build(seed)
        assert torch.equal(m["pre"][0].weight.data[:, H - R:], m0["pre"][0].weight.data[:, gids])


def test_sparse_to_dense_ids_edge_cases():
    ids = torch.tensor([5, 1, 3])
    got = reorder.sparse_to_dense_ids(ids, 6)
    assert got.tolist() == [0, 2, 4, 5, 1, 3]                         # (unsorted ids keep their order at the end)
    assert reorder.sparse_to_dense_ids(torch.tensor([], dtype=torch.long), 4).tolist() == [0, 1, 2, 3]
