import os, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import torch.distributed._symmetric_memory as symm_mem
try:
    t = symm_mem.empty((1024,), dtype=torch.float16, device=f"cuda:{local}")
    hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name if hasattr(dist.group.WORLD, "group_name") else dist.group.WORLD)
    print(rank, "rendezvous ok", [hex(p) for p in hdl.buffer_ptrs], "signal pads", [hex(p) for p in hdl.signal_pad_ptrs], flush=True)
    t.fill_(rank + 1)
    hdl.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (1024,), torch.float16)
    print(rank, "peer value", peer[0].item(), flush=True)
    hdl.barrier()
except Exception as e:
    print(rank, "FAILED", type(e).__name__, e, flush=True)
dist.destroy_process_group()
