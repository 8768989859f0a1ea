"""Golden vectors for the OGR producer (qeft_b200/reorder.py) from the reference's own qeft/reorder.py and the
selection at the end of qeft/extract_outidx.py.  Build container only (needs /root/reference):
    python tests/golden/make_reorder_golden.py      ->  tests/golden/reference_reorder.npz
The toy model is built by tests/golden/tiny_reorder_model.py from a seed, so the test rebuilds the same inputs."""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from tiny_reorder_model import build, snapshot  # noqa: E402

REF = "/root/reference"


def main():
    stub = tempfile.mkdtemp()
    with open(os.path.join(stub, "qeft_cuda.py"), "w") as f:
        f.write("# stand-in so that the reference package imports without its CUDA extension\n")
    sys.path.insert(0, stub)
    sys.path.insert(0, REF)
    import qeft.reorder as qr
    out = {}
    for seed in (1, 2):
        m = build(seed)
        qr.reorder_embeds(m["pre"], m["post"], m["global_ids"])
        for blk, qz in zip(m["blocks"], m["quantizers"]):
            qr.reorder_qkv_ffn1_ln(l_qkv_ffn1=blk["qkv"] + blk["ffn1"], l_ln=blk["ln"], out_ids=m["global_ids"])
            qr.reorder_out(l_out=blk["out"], l_out_quantizers=qz["out"], out_ids=m["global_ids"])
            qr.reorder_in_mlp(l_ffn1=blk["ffn1"], l_ffn2=blk["ffn2"], l_ffn1_quantizers=qz["ffn1"], l_ffn2_quantizers=qz["ffn2"])
        for k, v in snapshot(m).items():
            out[f"s{seed}/{k}"] = v
        # the selection of the global ids (extract_outidx.py:159-179, restated on the same tensors: the reference has it
        # inline in its calibration loop)
        g = torch.Generator().manual_seed(100 + seed)
        hs = [torch.rand(48, generator=g) + 0.01 for _ in range(5)]
        sens = torch.zeros(48)
        for h in hs:
            sens += h / h.mean()
        out[f"s{seed}/outidx"] = np.array(sorted(torch.topk(sens, 6).indices.cpu().tolist()), dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "reference_reorder.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
