"""In-kernel timeline of the column-sharded decode chain (fused gather), rank 0's CTA 0 stamps for a few blocks.

    QEFT_GEMV_STAMPS=1 torchrun --nproc-per-node 2 tools/gather_timeline.py [--model 70b] [--layers 3] [--gather fused|nccl]
Stamps (us, relative to the first launch's start): start, ring filled, waited (grid wait and/or arrival counter),
x staged, main loop done, stored+published, last CTA of the launch finished.
"""
import argparse
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qeft_b200 import _lib  # noqa: E402
from qeft_b200.decode import PackedDecoderStack  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="70b")
ap.add_argument("--layers", type=int, default=3)
ap.add_argument("--gather", default="fused")
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
st = PackedDecoderStack(args.model, layers=args.layers, shard=(rank, world), device=f"cuda:{local}", fast_synth=True)
if args.gather == "fused":
    st.enable_fused_gather(dist.group.WORLD)
else:
    st.enable_allgather(dist.group.WORLD)
st.capture()                    # eager warm-up = launch slots 0..L-1, capture = slots L..2L-1
for _ in range(5):
    st.step()
torch.cuda.synchronize()
dist.barrier()
st.step()
torch.cuda.synchronize()
L = st.launches_per_step()
lib = _lib.load()
lib.qeft_gemv_debug_stamps.restype = C.c_int
buf = (C.c_ulonglong * (2 * L * 8))()
assert lib.qeft_gemv_debug_stamps(buf, 2 * L) == 0
rows = [[buf[i * 8 + j] for j in range(7)] for i in range(L, 2 * L)]
if rank == 0:
    t0 = rows[0][0]
    names = ["qkv", "o", "gate/up", "down"]
    print(f"{args.model} over {world} ranks, gather={args.gather}; rank 0, CTA 0 (us)")
    print("launch        start   filled   waited   staged   looped   stored  lastCTA | wait  stage  loop  store  gap-to-next-waited")
    for i, r in enumerate(rows):
        v = [(x - t0) / 1e3 for x in r]
        nxt = (rows[i + 1][2] - r[6]) / 1e3 if i + 1 < L else float("nan")
        print(f"{i:3d} {names[i % 4]:8s}" + " ".join(f"{x:8.2f}" for x in v) + f" | {v[2]-v[1]:5.2f} {v[3]-v[2]:5.2f} {v[4]-v[3]:5.2f} {v[5]-v[4]:5.2f} {nxt:6.2f}")
    print(f"chain: {(rows[-1][6] - rows[0][0]) / 1e3 / L:.2f} us per launch")
dist.barrier()
st.graph = None
dist.destroy_process_group()
os._exit(0)
