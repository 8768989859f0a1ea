"""In-kernel timeline of the persistent decode program (QEFT_DECODE_STAMPS=1): per stage, for 4 CTAs, the times of
[barrier passed, x staged, units consumed (CTA-wide), rows stored + arrival signalled].
    QEFT_DECODE_STAMPS=1 python tools/decode_timeline.py [model] [layers]"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

os.environ.setdefault("QEFT_DECODE_STAMPS", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qeft_b200 import _lib  # noqa: E402
from qeft_b200.decode import PackedDecoderStack  # noqa: E402


def main():
    model = sys.argv[1] if len(sys.argv) > 1 else "7b"
    layers = int(sys.argv[2]) if len(sys.argv) > 2 else None
    st = PackedDecoderStack(model, layers=layers, fast_synth=True)
    prog = st.enable_program()
    for _ in range(5):
        st.step_eager()
    torch.cuda.synchronize()
    n = prog.nstages
    buf = np.zeros((n, 4, 4), dtype=np.uint64)
    lib = _lib.load()
    lib.qeft_decode_debug_stamps.restype = C.c_int
    lib.qeft_decode_debug_stamps.argtypes = [C.c_void_p, C.c_void_p]
    rc = lib.qeft_decode_debug_stamps(prog._h, buf.ctypes.data)
    assert rc == 0, rc
    t = buf.astype(np.int64)
    t0 = t[0, :, 0].min()
    rel = (t - t0) / 1e3     # us
    names = ["qkv", "o", "gateup", "down"]
    agg = {k: {"wait": [], "stage_x": [], "consume": [], "reduce_store": []} for k in names}
    for s in range(1, n):
        k = names[s % 4]
        for c in range(4):
            agg[k]["wait"].append(rel[s, c, 0] - rel[s - 1, c, 3])
            agg[k]["stage_x"].append(rel[s, c, 1] - rel[s, c, 0])
            agg[k]["consume"].append(rel[s, c, 2] - rel[s, c, 1])
            agg[k]["reduce_store"].append(rel[s, c, 3] - rel[s, c, 2])
    out = {k: {kk: round(float(np.median(vv)), 2) for kk, vv in v.items()} for k, v in agg.items()}
    out["total_us"] = round(float(rel[n - 1, :, 3].max()), 1)
    out["first_stages_cta0_us"] = [[round(float(x), 2) for x in rel[s, 0]] for s in range(min(n, 8))]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
