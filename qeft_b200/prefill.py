"""One prefill (or fine-tune forward) pass of M tokens through the packed linears of a Llama-shaped stack.

BASELINE.json configs[2] (one GPU) and the prefill half of configs[4] (70B shapes column-sharded over 2/4/8 GPUs,
SURVEY.md 8e).  Per decoder block the seven ``QuantLinear`` GEMMs (reference: ``forward_outlier`` -> ``gemm_4bit`` +
outlier ``F.linear``, qeft/qlinear.py:262-268; here ONE fused tcgen05 launch each).  Sharded execution: rank p owns a
row slab of every layer (``modelutils.shard_layer_tensors``, no repacking), x is replicated, rank p computes
``y[:, slab]``; the slabs meet in a ``[M, N]`` buffer on every rank either

* ``enable_fused_gather``: by the GEMM's own epilogue -- each output tile is stored into every rank's buffer through
  peer-mapped pointers (torch symmetric memory over NVLink) while other tiles still compute; per-launch arrival
  counters order a launch after the gathered input it reads.  No collective call, no extra pass over y; or
* ``enable_allgather``: local ``[M, N/P]`` output, NCCL all-gather to ``[P, M, N/P]``, one permuting copy to
  ``[M, N]`` (the baseline).

As in decode.py, attention / norms / activation functions are outside this path: every launch reads a fixed
activation buffer of the right width, so a step's cost is exactly the packed-linear work (+ the exchange).
"""
from __future__ import annotations

from typing import List

import torch

from . import _lib, qeft_cuda
from .modelutils import shard_layer_tensors
from .synth import LLAMA_SHAPES, synth_tensors

NAMES = ("q", "k", "v", "o", "gate", "up", "down")


class PackedPrefillStack:
    def __init__(self, model="7b", M=2048, layers=None, r=128, G=128, device="cuda", seed=0, shard=(0, 1), pdl=True,
                 fast_synth=True, shard_from_full=False):
        h, f, nl, kv = LLAMA_SHAPES[model] if isinstance(model, str) else model
        self.model, self.M, self.r, self.G, self.device = model, M, r, G, torch.device(device)
        self.h, self.f, self.kv = h, f, kv
        self.nlayers = nl if layers is None else layers
        self.rank, self.world = shard
        self.pdl = pdl
        rank, world = shard
        self.full = {"q": h, "k": kv, "v": kv, "o": h, "gate": f, "up": f, "down": h}      # N of every linear
        self.kin = {"q": h, "k": h, "v": h, "o": h, "gate": h, "up": h, "down": f}         # K
        for n, N in self.full.items():
            assert N % (world * 128) == 0, f"{n}: N={N} does not split into {world} slabs of a multiple of 128 rows"
        self.blocks: List[dict] = []
        for li in range(self.nlayers):
            blk = {}
            for pi, name in enumerate(NAMES):
                N, K = self.full[name], self.kin[name]
                if shard_from_full:      # tests: every rank derives its slab from the same full layer
                    t = shard_layer_tensors(synth_tensors(N, K, r, G, seed=seed * 100003 + li * 16 + pi,
                                                          device=self.device, fast=fast_synth), rank, world)
                else:                    # bench: the slab is generated directly (values do not matter for timing)
                    t = synth_tensors(N // world, K, r, G, seed=seed * 100003 + li * 16 + pi + rank * 7919,
                                      device=self.device, fast=fast_synth)
                t["N"] = N // world
                blk[name] = t
            self.blocks.append(blk)
        g = torch.Generator(device=self.device)
        g.manual_seed(seed + 29)
        self.x_h = torch.randn((M, h), device=self.device, generator=g).half()
        self.x_f = torch.randn((M, f), device=self.device, generator=g).half()
        # outputs: two sets of seven buffers, alternating by block parity (a launch never overwrites a buffer that a
        # launch less than seven launches back may still be reading)
        self.mode = "local"
        self.y_local = [{n: torch.empty((M, self.full[n] // world), dtype=torch.float16, device=self.device) for n in NAMES}
                        for _ in range(2)]
        self.y_full = None
        self.pg = None

    # ---- accounting -----------------------------------------------------------------------------
    def flops_per_step(self) -> float:
        """2 M N K over this rank's slabs (outlier columns counted once, as part of K)."""
        return sum(2.0 * self.M * (self.full[n] // self.world) * self.kin[n] for n in NAMES) * self.nlayers

    def gathered_bytes_per_step(self) -> int:
        """Bytes this rank sends to its peers per step (its slab of every output, to each of the other ranks)."""
        return sum(2 * self.M * (self.full[n] // self.world) for n in NAMES) * self.nlayers * (self.world - 1)

    def launches_per_step(self) -> int:
        return 7 * self.nlayers

    # ---- exchange -------------------------------------------------------------------------------
    def enable_allgather(self, process_group):
        self.pg, self.mode = process_group, "nccl"
        P, M = self.world, self.M
        self.stage = {n: torch.empty((P, M, self.full[n] // P), dtype=torch.float16, device=self.device) for n in NAMES}
        self.y_full = [{n: torch.empty((M, self.full[n]), dtype=torch.float16, device=self.device) for n in NAMES}
                       for _ in range(2)]

    def enable_fused_gather(self, process_group, multicast=None):
        """``multicast``: store each tile once to the NVLS multicast mapping of the gathered buffer (the NVSwitch
        replicates it to all ranks) when the symmetric allocation has one; else one store per rank.  Default: only
        from 8 ranks up (measured on B200: 10 % faster than per-rank stores at 8 ranks, 6-8 % slower at 2 and 4,
        where the copy that comes back to the sender costs more than the saved egress)."""
        if multicast is None:
            multicast = self.world >= 8
        import torch.distributed._symmetric_memory as symm_mem
        P, M = self.world, self.M
        nlaunch = 7 * self.nlayers
        flag_bytes = ((nlaunch * 4 + 255) // 256) * 256
        widths = [self.full[n] for n in NAMES]
        total = flag_bytes + 2 * sum(2 * M * w for w in widths)
        buf = symm_mem.empty((total,), dtype=torch.uint8, device=self.device)
        buf.zero_()
        hdl = symm_mem.rendezvous(buf, process_group)
        self._symm = (buf, hdl)
        mc = int(getattr(hdl, "multicast_ptr", 0) or 0) if multicast else 0
        self.multicast = bool(mc)
        self.epoch = torch.zeros((1,), dtype=torch.int32, device=self.device)
        offs, off = [], flag_bytes
        self.y_full = []
        for s in range(2):
            o, views = {}, {}
            for n, w in zip(NAMES, widths):
                o[n] = off
                views[n] = buf[off:off + 2 * M * w].view(torch.float16).view(M, w)
                off += 2 * M * w
            offs.append(o)
            self.y_full.append(views)
        self.gathers, prev_flag, li_flat = [], None, 0
        for li in range(self.nlayers):
            row = {}
            for n in NAMES:
                w = self.full[n]
                g = _lib.Gather()
                g.nranks, g.y_ld = P, w
                for pr in range(P):
                    g.y_peer[pr][0] = hdl.buffer_ptrs[pr] + offs[li % 2][n] + 2 * self.rank * (w // P)
                    g.done_peer[pr] = hdl.buffer_ptrs[pr] + 4 * li_flat
                if mc:
                    g.y_mc[0] = mc + offs[li % 2][n] + 2 * self.rank * (w // P)
                g.wait_flag = prev_flag
                g.epoch = self.epoch.data_ptr()
                prev_flag = buf.data_ptr() + 4 * li_flat
                row[n] = g
                li_flat += 1
            self.gathers.append(row)
        hdl.barrier()
        self.pg, self.mode = None, "fused_mc" if mc else "fused"

    # ---- one pass -------------------------------------------------------------------------------
    def step(self):
        G = self.G
        fused = self.mode in ("fused", "fused_mc")
        if fused:
            self.epoch.add_(1)
        for li, blk in enumerate(self.blocks):
            s = li % 2
            for n in NAMES:
                t = blk[n]
                x = self.x_f if n == "down" else self.x_h
                if fused:
                    qeft_cuda.gemm_w4_gather(x, t["qweight"], t["scales"], t["scaled_zeros"], t["oweight"], t.get("bias"),
                                             self.gathers[li][n], group_size=G, pdl=self.pdl)
                    continue
                y = self.y_local[s][n]
                qeft_cuda.gemm_w4(x, t["qweight"], t["scales"], t["scaled_zeros"], t["oweight"], t.get("bias"),
                                  group_size=G, out=y, pdl=self.pdl)
                if self.mode == "nccl":
                    import torch.distributed as dist
                    dist.all_gather_into_tensor(self.stage[n].view(-1), y.view(-1), group=self.pg)
                    self.y_full[s][n].view(self.M, self.world, -1).copy_(self.stage[n].permute(1, 0, 2))
        return (self.y_full if self.mode != "local" else self.y_local)[(self.nlayers - 1) % 2]["down"]
