"""Generate tests/golden/*.npz|json by IMPORTING the reference's own Python.

Run in the build container only (needs /root/reference, which does not exist on the
GPU box):   python tests/golden/make_golden.py
The reference has no test vectors of its own; these files pin the
oracle and the product packers to what the reference code itself computes:
  qeft/qlinear.py:70-79   pack_oweight
  qeft/qlinear.py:81-121  pack_intweight
  qeft/qlinear.py:125-215 QuantLinear.__init__/pack  (buffers, state-dict schema)
  qeft/reorder.py:6-12    sparse_to_dense_ids
  qeft/quant.py:194-214   make_quant
"""
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def import_reference():
    stub = tempfile.mkdtemp()
    with open(os.path.join(stub, "qeft_cuda.py"), "w") as f:
        f.write("# empty stand-in so that qeft.qlinear imports without the CUDA extension\n")
    sys.path.insert(0, stub)
    sys.path.insert(0, REF)
    import qeft.qlinear as ql
    import qeft.quant as qq
    import qeft.reorder as qr
    return ql, qq, qr


def main():
    ql, qq, qr = import_reference()
    rng = np.random.default_rng(20261018)
    out = {}

    # --- pack_intweight -------------------------------------------------------------
    shapes = [(4, 64), (8, 64), (8, 128), (16, 192), (64, 256), (128, 512), (24, 320)]
    for idx, (N, K) in enumerate(shapes):
        q = rng.integers(0, 16, size=(N, K), dtype=np.int32)
        packed = ql.pack_intweight(torch.tensor(q), interleave=4, kstride=64).numpy()
        out[f"piw{idx}_q"] = q.astype(np.uint8)
        out[f"piw{idx}_packed"] = packed
    # structured probe: every (row-in-tile, k-in-tile) position lights one nibble
    q = np.zeros((4, 64), dtype=np.int32)
    probes = []
    for j in range(4):
        for kk in range(64):
            q[:] = 0
            q[j, kk] = 0xF
            p = ql.pack_intweight(torch.tensor(q), interleave=4, kstride=64).numpy().view(np.uint16)[0]
            e = int(np.nonzero(p)[0][0])
            i = int(np.log2(int(p[e]) // 0xF)) // 4
            probes.append((j, kk, e, i))
    out["piw_probe"] = np.array(probes, dtype=np.int32)

    # --- pack_oweight ---------------------------------------------------------------
    for idx, (N, r) in enumerate([(8, 32), (16, 64), (64, 128), (24, 96)]):
        ow = rng.standard_normal((N, r)).astype(np.float16)
        packed = ql.pack_oweight(torch.tensor(ow), interleave=4).numpy()
        out[f"pow{idx}_ow"] = ow
        out[f"pow{idx}_packed"] = packed

    # --- sparse_to_dense_ids --------------------------------------------------------
    for idx, (K, r) in enumerate([(64, 8), (256, 32), (512, 128)]):
        ids = np.sort(rng.choice(K, size=r, replace=False)).astype(np.int32)
        dense = qr.sparse_to_dense_ids(torch.tensor(ids), K).numpy()
        out[f"s2d{idx}_ids"] = ids
        out[f"s2d{idx}_dense"] = dense
        out[f"s2d{idx}_K"] = np.array(K)
    # unsorted ids are kept in the given order
    ids = np.array([5, 2, 60, 33], dtype=np.int32)
    out["s2d3_ids"] = ids
    out["s2d3_dense"] = qr.sparse_to_dense_ids(torch.tensor(ids), 64).numpy()
    out["s2d3_K"] = np.array(64)

    # --- QuantLinear.pack -----------------------------------------------------------
    cases = [
        dict(N=32, K=256, G=128, r=128, sym=False, bias=False, name="model.layers.0.self_attn.q_proj"),
        dict(N=16, K=384, G=128, r=64, sym=False, bias=True, name="model.layers.0.mlp.down_proj"),
        dict(N=32, K=256, G=128, r=128, sym=True, bias=False, name="model.layers.0.self_attn.o_proj"),
        dict(N=16, K=256, G=256, r=32, sym=False, bias=False, name="model.layers.0.mlp.up_proj"),  # per-channel
        dict(N=16, K=256, G=128, r=0, sym=False, bias=False, name="model.layers.0.mlp.gate_proj"),
    ]
    schema = {}
    for idx, c in enumerate(cases):
        N, K, G, r = c["N"], c["K"], c["G"], c["r"]
        lin = torch.nn.Linear(K, N, bias=c["bias"])
        w = (rng.standard_normal((N, K)) * 0.02).astype(np.float32)
        lin.weight.data = torch.tensor(w).half()
        if c["bias"]:
            lin.bias.data = torch.tensor(rng.standard_normal(N).astype(np.float32) * 0.01).half()
        ng = K // G
        wg = lin.weight.data.float().reshape(N, ng, G)
        wmax, wmin = wg.amax(-1), wg.amin(-1)
        if c["sym"]:
            amax = torch.maximum(wmax.abs(), wmin.abs())
            scales = (amax / 7.5).half()
            zeros = torch.zeros_like(scales)           # pack() adds 8 in place when sym
        else:
            scales = ((wmax - wmin) / 15).half()
            zeros = torch.round(-wmin / scales.float()).half()
        ids = (np.sort(rng.choice(K, size=r, replace=False)).astype(np.int32) if "o_proj" in c["name"]
               else np.arange(K - r, K, dtype=np.int32))
        layer = ql.QuantLinear(4, K, N, c["bias"], torch.float16, r, G, True, c["name"])
        zeros_in = zeros.clone()
        layer.pack(lin, scales.clone(), zeros_in, torch.tensor(ids), sym=c["sym"])
        p = f"qlp{idx}_"
        out[p + "weight"] = lin.weight.data.numpy()
        out[p + "scales_in"] = scales.numpy()
        out[p + "zeros_in"] = zeros.numpy()
        out[p + "zeros_after"] = zeros_in.numpy()       # documents the in-place +8 for sym
        out[p + "outlieridx"] = ids
        out[p + "qweight"] = layer.qweight.numpy()
        out[p + "scales"] = layer.scales.numpy()
        out[p + "scaled_zeros"] = layer.scaled_zeros.numpy()
        if r > 0:
            out[p + "oweight"] = layer.oweight.numpy()
            out[p + "oweight_interleaved"] = layer.oweight_interleaved.numpy()
        if c["bias"]:
            out[p + "bias"] = layer.bias.detach().numpy()
        fresh = ql.QuantLinear(4, K, N, c["bias"], torch.float16, r, G, True, c["name"])
        schema[f"qlp{idx}"] = {
            "case": c,
            "state_dict": {k: [list(v.shape), str(v.dtype)] for k, v in fresh.state_dict().items()},
        }

    # --- make_quant on a tiny module tree --------------------------------------------
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
    qq.make_quant(tiny, infos)
    schema["make_quant"] = {k: [list(v.shape), str(v.dtype)] for k, v in tiny.state_dict().items()}

    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    with open(os.path.join(HERE, "reference_schema.json"), "w") as f:
        json.dump(schema, f, indent=1, sort_keys=True)
    print("wrote", len(out), "arrays;", os.path.getsize(os.path.join(HERE, "reference_vectors.npz")), "bytes")


if __name__ == "__main__":
    main()
