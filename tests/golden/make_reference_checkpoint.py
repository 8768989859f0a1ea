"""Write tests/golden/reference_packed_ckpt.pth and reference_wct/model.pth WITH THE REFERENCE'S OWN CODE:
qeft/utils/modelutils.py:219-268 (save_model -> lm_pack -> QuantLinear.pack) and :270-284 (save_wctmodel), imported
from /root/reference in the build container (python tests/golden/make_reference_checkpoint.py).  The tests load these
files with this repo's loader (SURVEY.md 8c pin (3)); reference_ckpt_expect.npz holds what the loaded layers must
stand for (the fake-quantised dense weights the reference packed, the fine-tuned outlier columns)."""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tiny_model  # noqa: E402

REF = "/root/reference"


def main():
    stub = tempfile.mkdtemp()
    with open(os.path.join(stub, "qeft_cuda.py"), "w") as f:
        f.write("# empty stand-in so that qeft.qlinear imports without the CUDA extension\n")
    sys.path.insert(0, stub)
    sys.path.insert(0, REF)
    import qeft.qlinear as ql
    import qeft.utils.modelutils as mu

    rng = np.random.default_rng(20261018)
    model = tiny_model.build(0)
    names = tiny_model.quant_layer_names()
    mods = dict(model.named_modules())
    G, r = tiny_model.GROUP, tiny_model.NOUT
    quantizers, expect = {}, {}
    for i, n in enumerate(names):
        lin = mods[n]
        N, K = lin.weight.shape
        w = (rng.standard_normal((N, K)) * 0.02).astype(np.float32)
        ng = K // G
        wg = torch.tensor(w).reshape(N, ng, G)
        wmax, wmin = wg.amax(-1), wg.amin(-1)
        scales = ((wmax - wmin) / 15).half()
        zeros = torch.round(-wmin / scales.float()).half()
        # the fake-quantised weight the quantiser leaves in the model (what pack() re-derives the integers from);
        # the last r columns are the fp16 outlier columns (OGR-reordered layout)
        s_full = torch.repeat_interleave(scales.float(), G, dim=1)
        z_full = torch.repeat_interleave(zeros.float(), G, dim=1)
        q = torch.clamp(torch.round(torch.tensor(w) / s_full) + z_full, 0, 15)
        fake = (s_full * (q - z_full))
        fake[:, K - r:] = torch.tensor(w)[:, K - r:]
        lin.weight.data = fake.half()
        ids = (np.sort(rng.choice(K, size=r, replace=False)).astype(np.int32) if "o_proj" in n
               else np.arange(K - r, K, dtype=np.int32))
        qz = types.SimpleNamespace(bits=4, sym=False, group_size=G, n_out=r, reorder=True, scale=scales, zero=zeros,
                                   out_ids=torch.tensor(ids))
        qz.cpu = (lambda self=qz: self)
        quantizers[n] = qz
        expect[n + ".fake_weight"] = lin.weight.data.numpy().copy()
        expect[n + ".outlieridx"] = ids
        if lin.bias is not None:
            expect[n + ".bias"] = lin.bias.data.numpy().copy()
    path = os.path.join(HERE, "reference_packed_ckpt.pth")
    mu.save_model(model, quantizers, path, packing=True, fake=False)

    # fine-tune stand-in: set_for_wct (qlinear.py:239-242), perturb the outlier columns, save with the reference
    for n, mod in model.named_modules():
        if isinstance(mod, ql.QuantLinear):
            mod.set_for_wct()
            mod.oweight.data += torch.tensor(rng.standard_normal(tuple(mod.oweight.shape)).astype(np.float32) * 0.01)
            expect[n + ".oweight_finetuned"] = mod.oweight.data.to(torch.float16).numpy().copy()
    wct_dir = os.path.join(HERE, "reference_wct")
    mu.save_wctmodel(model, path, wct_dir)
    np.savez_compressed(os.path.join(HERE, "reference_ckpt_expect.npz"), **expect)
    print("wrote", path, os.path.getsize(path), "bytes;", os.path.getsize(os.path.join(wct_dir, "model.pth")), "bytes")


if __name__ == "__main__":
    main()
