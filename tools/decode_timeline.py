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
    layers = int(sys.argv[2]) if len(sys.argv) > 2 and int(sys.argv[2]) > 0 else None
    st = PackedDecoderStack(model, layers=layers, fast_synth=True)
    mode = sys.argv[3] if len(sys.argv) > 3 else "fixed"      # fixed | chain_barrier | chain_dataflow
    prog = st.enable_program() if mode == "fixed" else st.enable_chain_program(dataflow=(mode == "chain_dataflow"))
    for _ in range(5):
        st.step_eager()
    torch.cuda.synchronize()
    print(json.dumps(collect(prog)))


def collect(prog):
    """Summary of the stamps of the program's last run (QEFT_DECODE_STAMPS=1 must have been set before the first run)."""
    n = prog.nstages
    buf = np.zeros((n * (32 + 640) + 24,), dtype=np.uint64)
    lib = _lib.load()
    lib.qeft_decode_debug_stamps.restype = C.c_int
    lib.qeft_decode_debug_stamps.argtypes = [C.c_void_p, C.c_void_p]
    rc = lib.qeft_decode_debug_stamps(prog._h, buf.ctypes.data)
    assert rc == 0, rc
    extra = buf[n * 32:n * 32 + 24].astype(np.int64)
    allc = buf[n * 32 + 24:].reshape(n, 160, 4)[:, :148].astype(np.int64)
    t = buf[:n * 32].reshape(n, 4, 8).astype(np.int64)
    t0 = t[0, :, 0].min()
    rel = (t - t0) / 1e3     # us
    names = ["qkv", "o", "gateup", "down"]
    agg = {k: {"wait": [], "stage_x": [], "consume": [], "reduce_store": [], "x_loads_max": [], "x_digits": [], "x_tail": [], "reduce": [], "rs_release": []} for k in names}
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
            agg[k]["reduce"].append(rel[s, c, 5] - rel[s, c, 2])
            agg[k]["rs_release"].append(rel[s, c, 3] - rel[s, c, 5])
    out = {k: {kk: round(float(np.median(vv)), 2) for kk, vv in v.items()} for k, v in agg.items()}
    out["total_us"] = round(float(rel[n - 1, :, 3].max()), 1)
    out["first_stages_cta0_us"] = [[round(float(x), 2) for x in rel[s, 0, :4]] for s in range(min(n, 8))]
    out["cta0_producer"] = {"issue_cycles": int(extra[16]), "issues": int(extra[17]),
                            "cycles_per_block": round(float(extra[16]) / max(1, int(extra[17])), 1)}
    # all CTAs: spread of the moments the CTAs finish consuming / storing a stage, and of their consume durations
    sk = {k: {"consume_end_spread": [], "store_spread": [], "consume_min": [], "consume_max": [], "enter_spread": [], "staged_spread": []} for k in names}
    for s in range(1, n):
        k = names[s % 4]
        a = allc[s] / 1e3
        sk[k]["enter_spread"].append(a[:, 0].max() - a[:, 0].min())
        sk[k]["staged_spread"].append(a[:, 1].max() - a[:, 1].min())
        sk[k]["consume_end_spread"].append(a[:, 2].max() - a[:, 2].min())
        sk[k]["store_spread"].append(a[:, 3].max() - a[:, 3].min())
        d = a[:, 2] - a[:, 1]
        sk[k]["consume_min"].append(d.min())
        sk[k]["consume_max"].append(d.max())
    out["all_ctas"] = {k: {kk: round(float(np.median(vv)), 2) for kk, vv in v.items()} for k, v in sk.items()}
    s_mid = 4 * (n // 8) + 2      # one gate+up stage in the middle: per-CTA consume time and end time (relative to the first end)
    a = allc[s_mid] / 1e3
    out["gateup_mid_consume_us"] = [round(float(x), 2) for x in (a[:, 2] - a[:, 1])]
    out["gateup_mid_end_rel_us"] = [round(float(x), 2) for x in (a[:, 2] - a[:, 2].min())]
    for nm, e in (("warp0", extra[:8]), ("warp15", extra[8:16])):
        out["cta0_" + nm] = {"wait_cycles": int(e[0]), "math_cycles": int(e[1]), "issue_cycles": int(e[2]),
                             "block_period_cycles_sum": int(e[3]), "issues": int(e[4]), "blocks_waited": int(e[5]),
                             "blocks": int(e[6])}
    return out


if __name__ == "__main__":
    main()
