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
    buf = np.zeros((n * 32 + 24,), dtype=np.uint64)
    lib = _lib.load()
    lib.qeft_decode_debug_stamps.restype = C.c_int
    lib.qeft_decode_debug_stamps.argtypes = [C.c_void_p, C.c_void_p]
    rc = lib.qeft_decode_debug_stamps(prog._h, buf.ctypes.data)
    assert rc == 0, rc
    extra = buf[n * 32:].astype(np.int64)
    t = buf[:n * 32].reshape(n, 4, 8).astype(np.int64)
    t0 = t[0, :, 0].min()
    rel = (t - t0) / 1e3     # us
    names = ["qkv", "o", "gateup", "down"]
    agg = {k: {"wait": [], "stage_x": [], "consume": [], "reduce_store": [], "x_loads_max": [], "x_digits": [], "x_tail": []} for k in names}
    for s in range(1, n):
        k = names[s % 4]
        for c in range(4):
            agg[k]["wait"].append(rel[s, c, 0] - rel[s - 1, c, 3])
            agg[k]["stage_x"].append(rel[s, c, 1] - rel[s, c, 0])
            agg[k]["consume"].append(rel[s, c, 2] - rel[s, c, 1])
            agg[k]["reduce_store"].append(rel[s, c, 3] - rel[s, c, 2])
            agg[k]["x_loads_max"].append(rel[s, c, 4] - rel[s, c, 0])
            agg[k]["x_digits"].append(rel[s, c, 6] - rel[s, c, 4])
            agg[k]["x_tail"].append(rel[s, c, 1] - rel[s, c, 6])
    out = {k: {kk: round(float(np.median(vv)), 2) for kk, vv in v.items()} for k, v in agg.items()}
    out["total_us"] = round(float(rel[n - 1, :, 3].max()), 1)
    out["first_stages_cta0_us"] = [[round(float(x), 2) for x in rel[s, 0, :4]] for s in range(min(n, 8))]
    out["cta0_producer"] = {"issue_cycles": int(extra[16]), "issues": int(extra[17]),
                            "cycles_per_block": round(float(extra[16]) / max(1, int(extra[17])), 1)}
    for nm, e in (("warp0", extra[:8]), ("warp15", extra[8:16])):
        out["cta0_" + nm] = {"wait_cycles": int(e[0]), "math_cycles": int(e[1]), "issue_cycles": int(e[2]),
                             "block_period_cycles_sum": int(e[3]), "issues": int(e[4]), "blocks_waited": int(e[5]),
                             "blocks": int(e[6])}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
