"""Time one decode token of the column-sharded chain program (torchrun, one rank per GPU) beside the round-1 fused-gather
launch chain.   torchrun --nproc-per-node N tools/sharded_program_time.py [model] [layers]"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from qeft_b200.decode import PackedDecoderStack  # noqa: E402


def timed(fn, steps=20, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    model = sys.argv[1] if len(sys.argv) > 1 else "70b"
    layers = int(sys.argv[2]) if len(sys.argv) > 2 and int(sys.argv[2]) > 0 else None
    st = PackedDecoderStack(model, layers=layers, shard=(rank, world), device=f"cuda:{local}", fast_synth=True)
    nbytes = st.algorithmic_bytes_per_step() * world
    out = {"model": model, "layers": st.nlayers, "ranks": world, "bytes_all_ranks": nbytes}
    prog = st.enable_sharded_chain_program(dist.group.WORLD)
    ms = timed(prog.run)
    out["sharded_program_ms"] = round(ms, 4)
    out["sharded_program_GBps"] = round(nbytes / ms / 1e6, 1)
    y = st.result().float()
    out["finite"] = bool(torch.isfinite(y).all())
    if os.environ.get("QEFT_DECODE_STAMPS"):
        from decode_timeline import collect
        tl = collect(prog)
        if rank == 0:
            print(json.dumps({k: tl[k] for k in ("qkv", "o", "gateup", "down", "total_us", "all_ctas")}), flush=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        prog.run()
    ms = timed(g.replay)
    out["sharded_program_graph_ms"] = round(ms, 4)
    del st, prog, g
    torch.cuda.empty_cache()
    if "--old" in sys.argv:
        st = PackedDecoderStack(model, layers=layers, shard=(rank, world), device=f"cuda:{local}", fast_synth=True)
        st.enable_fused_gather(dist.group.WORLD)
        st.capture()
        ms = timed(st.step)
        out["fused_gather_chain_ms"] = round(ms, 4)
        out["fused_gather_chain_GBps"] = round(nbytes / ms / 1e6, 1)
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
