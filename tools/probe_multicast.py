"""Does torch symmetric memory give a multicast (NVLS) mapping on this box?  torchrun --nproc-per-node N tools/probe_multicast.py"""
import os
import torch
import torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import torch.distributed._symmetric_memory as symm_mem
t = symm_mem.empty((1 << 20,), dtype=torch.uint8, device=f"cuda:{local}")
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
mc = getattr(hdl, "multicast_ptr", None)
print(rank, "multicast_ptr", hex(mc) if mc else mc, "has_multicast_support",
      getattr(symm_mem, "has_multicast_support", lambda *a: "n/a")(torch.device("cuda").type, local), flush=True)
hdl.barrier()
dist.destroy_process_group()
