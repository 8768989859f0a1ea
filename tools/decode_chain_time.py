"""Time one decode token of the fused decoder chain (stages feeding each other, glue fused in) ordered by data-flow words
vs by barriers, beside the fixed-buffer program.  python tools/decode_chain_time.py [model] [layers] [batch]"""
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
    out = {"model": model, "layers": st.nlayers, "batch": batch, "bytes": nbytes}
    for name, fn in (("chain_dataflow", lambda: st.enable_chain_program(True)), ("chain_barrier", lambda: st.enable_chain_program(False)),
                     ("fixed_buffers", st.enable_program)):
        prog = fn()
        ms = timed(prog.run)
        out[name + "_ms"] = round(ms, 4)
        out[name + "_GBps"] = round(nbytes / ms / 1e6, 1)
        if name.startswith("chain"):
            y = st.chain[-1]["out"].float()
            out[name + "_finite"] = bool(torch.isfinite(y).all())
            out[name + "_absmax"] = float(y.abs().max())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
