"""Time one decode token through the persistent decode program against the round-1 launch chain (same stack, same
buffers) and compare their outputs.  python tools/decode_program_time.py [model] [layers] [batch]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qeft_b200.decode import PackedDecoderStack  # noqa: E402


def timed(fn, steps=20, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    model = sys.argv[1] if len(sys.argv) > 1 else "7b"
    layers = int(sys.argv[2]) if len(sys.argv) > 2 else None
    batch = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    st = PackedDecoderStack(model, layers=layers, fast_synth=True, batch=batch)
    nbytes = st.algorithmic_bytes_per_step()
    st.capture()
    ms_old = timed(st.step)
    old = [{k: v.clone() for k, v in o.items()} for o in st.out]
    for o in st.out:
        for v in o.values():
            v.fill_(float("nan"))
    st.graph = None
    st.enable_program()
    st.step_eager()
    torch.cuda.synchronize()
    worst = 0.0
    for o, n in zip(old, st.out):
        for k in o:
            a, b = o[k].float(), n[k].float()
            if not torch.isfinite(b).all():
                worst = float("inf")
            worst = max(worst, float((a - b).abs().max() / a.abs().max()))
    ms_eager = timed(st.step_eager)
    st.capture()
    ms_graph = timed(st.step)
    print(json.dumps({"model": model, "layers": st.nlayers, "batch": batch, "bytes": nbytes,
                      "old_chain_ms": round(ms_old, 4), "old_GBps": round(nbytes / ms_old / 1e6, 1),
                      "program_eager_ms": round(ms_eager, 4), "program_graph_ms": round(ms_graph, 4),
                      "program_GBps": round(nbytes / min(ms_eager, ms_graph) / 1e6, 1),
                      "max_diff_vs_old_over_max": worst}))


if __name__ == "__main__":
    main()
