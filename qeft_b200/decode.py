"""One decode token through the packed linears of a Llama-shaped stack, as a replayable CUDA graph.

This is the data-parallel hot path BASELINE.json's metric is quoted on: per decoder block the seven
``QuantLinear`` GEMVs (reference call stack: qeft/main.py:356-366 -> HF LlamaDecoderLayer ->
``QuantLinear.forward_outlier`` -> ``qeft_cuda.gemv_4bit_qeft``).  B200-first differences:

* q/k/v and gate/up share their input, so each group is ONE launch (4 launches per block, not 7+);
* every launch uses programmatic dependent launch, so the weight stream of launch i+1 starts
  while launch i drains;
* the whole token (4 x layers launches) is captured once in a CUDA graph and replayed.

Attention, norms and the activation function are not part of this path (out of scope); the
stack feeds each group from fixed activation buffers, so a step's cost is exactly the packed-linear work.
"""
from __future__ import annotations

from typing import List

import torch

from . import _lib, qeft_cuda
from .reorder import sparse_to_dense_ids
from .synth import LLAMA_SHAPES, synth_tensors


class PackedDecoderStack:
    def __init__(self, model="7b", layers=None, r=128, G=128, device="cuda", seed=0, fused=True, pdl=True,
                 shard=(0, 1), batch=1, fast_synth=False):
        h, f, nl, kv = LLAMA_SHAPES[model]
        self.model, self.r, self.G, self.device = model, r, G, torch.device(device)
        self.h, self.f, self.kv = h, f, kv
        self.nlayers = nl if layers is None else layers
        self.fused, self.pdl, self.batch = fused, pdl, batch
        rank, world = shard
        self.rank, self.world = rank, world

        def sl(n):   # column (output-feature) shard owned by this rank
            assert n % (world * 16) == 0
            return n // world

        self.nq, self.nkv, self.nf, self.no = sl(h), sl(kv), sl(f), sl(h)
        self.blocks: List[dict] = []
        for li in range(self.nlayers):
            blk = {}
            for pi, (name, N, K) in enumerate((("q", self.nq, h), ("k", self.nkv, h), ("v", self.nkv, h),
                                               ("o", self.no, h), ("gate", self.nf, h), ("up", self.nf, h),
                                               ("down", self.no, f))):
                blk[name] = synth_tensors(N, K, r, G, seed=seed * 100003 + li * 16 + pi + rank * 7919,
                                          device=self.device, o_proj=(name == "o"), fast=fast_synth)
                blk[name]["N"] = N
            if r > 0:   # o_proj gathers its outlier channels to the back (qlinear.py:273-275): fused into the GEMV
                # the OGR outlier indices are global: the same on every rank of a column-sharded layer
                gi = torch.Generator(device=self.device)
                gi.manual_seed(seed * 100003 + li * 16 + 999)
                blk["o"]["outlieridx"] = torch.randperm(h, device=self.device, generator=gi)[:r].sort().values.to(torch.int32)
                blk["o"]["reorder_ids32"] = sparse_to_dense_ids(blk["o"]["outlieridx"], h).to(torch.int32)
            self.blocks.append(blk)
        g = torch.Generator(device=self.device)
        g.manual_seed(seed + 17)
        m = batch
        self.x_h = torch.randn((m, h), device=self.device, generator=g).half()     # hidden-size input (q/k/v/o/gate/up)
        self.x_f = torch.randn((m, f), device=self.device, generator=g).half()     # ffn-size input (down)
        # outputs: one contiguous buffer per launch group so that a sharded run needs ONE all-gather per launch
        self.groups = (("q", "k", "v"), ("o",), ("gate", "up"), ("down",))
        self.out, self.grp_local, self.grp_full = [], [], []
        for blk in self.blocks:
            o, gl = {}, []
            for names in self.groups:
                width = sum(blk[n]["N"] for n in names)
                buf = torch.empty((m, width), dtype=torch.float16, device=self.device)
                gl.append(buf)
                off = 0
                for n in names:
                    # batch 1: a column slice of the group buffer is itself a contiguous [1, N] output
                    o[n] = buf[:, off:off + blk[n]["N"]] if m == 1 else torch.empty((m, blk[n]["N"]), dtype=torch.float16,
                                                                                     device=self.device)
                    off += blk[n]["N"]
            self.out.append(o)
            self.grp_local.append(gl)
        self.pg = None
        self.graph = None
        self.program = None

    def enable_program(self):
        """The whole token as ONE launch of the persistent decode kernel (``qeft_cuda.DecodeProgram``): 4 stages per
        block, stage boundaries are gpu-scope barriers inside the kernel.  Same inputs, outputs and arithmetic
        bytes as :meth:`step_eager`; unsharded stacks only."""
        assert self.world == 1 and self.fused
        stages = []
        for blk, out in zip(self.blocks, self.out):
            for names in self.groups:
                x = self.x_f if names[0] == "down" else self.x_h
                st = {"x": x, "K": x.shape[-1], "r": self.r, "G": self.G,
                      "parts": [{"qweight": blk[n]["qweight"], "scales": blk[n]["scales"],
                                 "scaled_zeros": blk[n]["scaled_zeros"], "oweight": blk[n].get("oweight"),
                                 "bias": blk[n].get("bias"), "N": blk[n]["N"], "y": out[n]} for n in names]}
                if names[0] == "o" and self.r > 0:
                    st["x_gather"] = blk["o"]["reorder_ids32"]
                stages.append(st)
        self.program = qeft_cuda.DecodeProgram(stages, m=self.batch)
        return self.program

    def enable_chain_program(self, dataflow=False, eps=1e-5, ln=None, hidden=None):
        """The token as ONE launch in which the stages really feed each other, with a Llama block's elementwise glue
        fused in: per block [RMSNorm -> q|k|v], [o_proj (reorder gather) + residual], [RMSNorm -> gate|up ->
        SiLU*mul], [down_proj + residual]; the next block's q|k|v reads this block's output.  Attention is not part of
        the path: o_proj consumes the q projection in its place.  ``dataflow``: stages are ordered by the data-flow words
        of their inputs (no barrier); False (default, measured faster: 1.47 vs 2.30 ms per 7B token, profiles/r02_decode_chain_*):
        by gpu-scope barriers (``QEFT_DECODE_LL=0``).  Same weights and weight bytes
        as :meth:`enable_program`."""
        import os
        assert self.world == 1 and self.fused
        m, dev = self.batch, self.device
        g = torch.Generator(device=dev)
        g.manual_seed(4242)
        self.ln = ln if ln is not None else [(1 + 0.1 * torch.randn((2, self.h), device=dev, generator=g)).half() for _ in self.blocks]
        self.chain = []
        stages = []
        hidden = self.x_h if hidden is None else hidden

        def parts(blk, names, ys):
            return [{"qweight": blk[n]["qweight"], "scales": blk[n]["scales"], "scaled_zeros": blk[n]["scaled_zeros"],
                     "oweight": blk[n].get("oweight"), "bias": blk[n].get("bias"), "N": blk[n]["N"], "y": y}
                    for n, y in zip(names, ys)]

        for li, blk in enumerate(self.blocks):
            e = lambda n: torch.empty((m, n), dtype=torch.float16, device=dev)  # noqa: E731
            buf = {"q": e(self.h), "k": e(self.kv), "v": e(self.kv), "h2": e(self.h), "act": e(self.f), "up": e(self.f),
                   "out": e(self.h)}
            kw = {"K": self.h, "r": self.r, "G": self.G}
            stages.append({"x": hidden, **kw, "norm_weight": self.ln[li][0], "norm_eps": eps,
                           "parts": parts(blk, ("q", "k", "v"), (buf["q"], buf["k"], buf["v"]))})
            st = {"x": buf["q"], **kw, "epilogue": "residual", "residual": hidden, "parts": parts(blk, ("o",), (buf["h2"],))}
            if self.r > 0:
                st["x_gather"] = blk["o"]["reorder_ids32"]
            stages.append(st)
            stages.append({"x": buf["h2"], **kw, "norm_weight": self.ln[li][1], "norm_eps": eps, "epilogue": "swiglu",
                           "parts": parts(blk, ("gate", "up"), (buf["act"], buf["up"]))})
            stages.append({"x": buf["act"], "K": self.f, "r": self.r, "G": self.G, "epilogue": "residual",
                           "residual": buf["h2"], "parts": parts(blk, ("down",), (buf["out"],))})
            self.chain.append(buf)
            hidden = buf["out"]
        old = os.environ.get("QEFT_DECODE_LL")
        os.environ["QEFT_DECODE_LL"] = "1" if dataflow else "0"
        try:
            self.program = qeft_cuda.DecodeProgram(stages, m=self.batch)
        finally:
            if old is None:
                del os.environ["QEFT_DECODE_LL"]
            else:
                os.environ["QEFT_DECODE_LL"] = old
        self.chain_mode = "dataflow" if dataflow else "barrier"
        return self.program

    def enable_sharded_chain_program(self, process_group, eps=1e-5):
        """The chained token of :meth:`enable_chain_program` on a column-sharded stack, as ONE launch per rank with the
        all-gathers inside the kernel: every projection computes this rank's row slab and stores its slice of the
        gathered activation row straight into every rank's copy (symmetric memory, NVLink stores); the next stage starts
        on the elements as they arrive (data-flow by sentinel: no collective call, no fence, no counter per stage; two
        rank barriers per token).  SURVEY.md 8e; replaces the per-launch exchange of :meth:`enable_fused_gather`."""
        import torch.distributed._symmetric_memory as symm_mem
        assert self.batch == 1 and self.fused and self.kv % self.world == 0
        P, rank, dev = self.world, self.rank, self.device
        g = torch.Generator(device=dev)
        g.manual_seed(4242)
        self.ln = [(1 + 0.1 * torch.randn((2, self.h), device=dev, generator=g)).half() for _ in self.blocks]
        # one symmetric allocation: [barrier word | per block: q, k, v, h2, act, out gathered rows]
        widths = {"q": self.h, "k": self.kv, "v": self.kv, "h2": self.h, "act": self.f, "out": self.h}
        per_block = sum(widths.values()) * 2
        total = 256 + per_block * len(self.blocks)
        buf = symm_mem.empty((total,), dtype=torch.uint8, device=dev)
        buf.zero_()
        hdl = symm_mem.rendezvous(buf, process_group)
        self._symm = (buf, hdl)
        self.chain, stages, offs = [], [], []
        hidden = self.x_h
        off = 256

        def parts(blk, names, ys):
            return [{"qweight": blk[n]["qweight"], "scales": blk[n]["scales"], "scaled_zeros": blk[n]["scaled_zeros"],
                     "oweight": blk[n].get("oweight"), "bias": blk[n].get("bias"), "N": blk[n]["N"], "y": y}
                    for n, y in zip(names, ys)]

        for li, blk in enumerate(self.blocks):
            full, o = {}, {}
            for name, w in widths.items():
                full[name] = buf[off:off + 2 * w].view(torch.float16).view(1, w)
                o[name] = off
                off += 2 * w
            up = torch.empty((1, self.nf), dtype=torch.float16, device=dev)

            def mine(name, n_local):       # this rank's slice of the local copy
                return full[name][:, rank * n_local:(rank + 1) * n_local]

            kw = {"K": self.h, "r": self.r, "G": self.G}
            stages.append({"x": hidden, **kw, "norm_weight": self.ln[li][0], "norm_eps": eps,
                           "parts": parts(blk, ("q", "k", "v"), (mine("q", self.nq), mine("k", self.nkv), mine("v", self.nkv)))})
            st = {"x": full["q"], **kw, "epilogue": "residual", "residual": hidden[:, rank * self.no:(rank + 1) * self.no],
                  "parts": parts(blk, ("o",), (mine("h2", self.no),))}
            if self.r > 0:
                st["x_gather"] = blk["o"]["reorder_ids32"]
            stages.append(st)
            stages.append({"x": full["h2"], **kw, "norm_weight": self.ln[li][1], "norm_eps": eps, "epilogue": "swiglu",
                           "parts": parts(blk, ("gate", "up"), (mine("act", self.nf), up))})
            stages.append({"x": full["act"], "K": self.f, "r": self.r, "G": self.G, "epilogue": "residual",
                           "residual": full["h2"][:, rank * self.no:(rank + 1) * self.no],
                           "parts": parts(blk, ("down",), (mine("out", self.no),))})
            self.chain.append(full)
            offs.append(o)
            hidden = full["out"]
        prog = qeft_cuda.DecodeProgram(stages, m=1)
        prog.set_ranks(P, rank, [hdl.buffer_ptrs[p] for p in range(P)])
        for li, o in enumerate(offs):
            for si, names in enumerate((("q", "k", "v"), ("h2",), ("act",), ("out",))):
                for pi, name in enumerate(names):
                    prog.shard(4 * li + si, pi, [hdl.buffer_ptrs[p] + o[name] for p in range(P)])
        hdl.barrier()
        self.program = prog
        self.chain_mode = "sharded_dataflow"
        self.fused_gather = None
        self.pg = None
        return prog

    def gathered_twin(self, layers=None):
        """The UNSHARDED stack with the same weights (every rank's slabs all-gathered; for parity checks of the sharded
        programs): a single-GPU :class:`PackedDecoderStack` on this rank's device.  ``layers``: block indices to keep."""
        import torch.distributed as dist
        assert self.world > 1

        def cat(t, dim):
            raw = t.contiguous().view(torch.uint8)               # (NCCL has no int16: gather the bytes)
            parts = [torch.empty_like(raw) for _ in range(self.world)]
            dist.all_gather(parts, raw)
            return torch.cat([q.view(t.dtype).view(t.shape) for q in parts], dim=dim).contiguous()

        idx = list(range(self.nlayers)) if layers is None else list(layers)
        twin = PackedDecoderStack.__new__(PackedDecoderStack)
        twin.model, twin.r, twin.G, twin.device = self.model, self.r, self.G, self.device
        twin.h, twin.f, twin.kv = self.h, self.f, self.kv
        twin.nlayers, twin.fused, twin.pdl, twin.batch = len(idx), True, self.pdl, 1
        twin.rank, twin.world = 0, 1
        twin.nq, twin.nkv, twin.nf, twin.no = self.h, self.kv, self.f, self.h
        twin.blocks = []
        for li in idx:
            blk, nb = self.blocks[li], {}
            for name in ("q", "k", "v", "o", "gate", "up", "down"):
                t = blk[name]
                nb[name] = {"qweight": cat(t["qweight"], 0), "scales": cat(t["scales"], 1), "scaled_zeros": cat(t["scaled_zeros"], 1),
                            "N": t["N"] * self.world}
                if "oweight" in t:
                    nb[name]["oweight"] = cat(t["oweight"], 0)
                if t.get("bias") is not None:
                    nb[name]["bias"] = cat(t["bias"], 0)
            if self.r > 0:
                nb["o"]["outlieridx"] = blk["o"]["outlieridx"]
                nb["o"]["reorder_ids32"] = blk["o"]["reorder_ids32"]
            twin.blocks.append(nb)
        twin.x_h, twin.x_f = self.x_h, self.x_f
        twin.groups = self.groups
        twin.out, twin.grp_local, twin.grp_full = [], [], []
        twin.pg = twin.graph = twin.program = None
        return twin

    def enable_allgather(self, process_group):
        """Column-sharded execution: after each launch group, all-gather the ranks' output slices (NCCL over
        NVLink).  The gathered buffer is [world, group_width]: rank-major, then q|k|v (or gate|up) inside."""
        assert self.batch == 1, "the sharded stack is exercised at batch 1"
        self.pg = process_group
        self.grp_full = [[torch.empty((self.world, b.shape[1]), dtype=torch.float16, device=self.device) for b in gl]
                         for gl in self.grp_local]

    def enable_fused_gather(self, process_group):
        """Column-sharded execution WITHOUT a collective call: the GEMV epilogue stores each rank's slice into every
        rank's gathered buffer (symmetric memory, peer pointers over NVLink) and signals per-launch arrival counters;
        a launch waits for the counter of the launch before it (the chain's data dependency).  One symmetric
        allocation holds the counters and all gathered outputs."""
        import torch.distributed._symmetric_memory as symm_mem
        assert self.batch == 1 and self.fused
        m, P = self.batch, self.world
        widths = [[b.shape[1] for b in gl] for gl in self.grp_local]          # local width of every launch group
        nlaunch = sum(len(w) for w in widths)
        flag_bytes = ((nlaunch * 4 + 255) // 256) * 256
        total = flag_bytes + sum(2 * m * P * w for ws in widths for w in ws)
        buf = symm_mem.empty((total,), dtype=torch.uint8, device=self.device)
        buf.zero_()
        hdl = symm_mem.rendezvous(buf, process_group)
        self._symm = (buf, hdl)
        self.epoch = torch.zeros((1,), dtype=torch.int32, device=self.device)
        self.fused_gather = []
        off, li_flat, prev_flag = flag_bytes, 0, None
        for li, ws in enumerate(widths):
            row = []
            for gi, w in enumerate(ws):
                g = _lib.Gather()
                g.nranks, g.y_ld = P, P * w
                names = self.groups[gi]
                for pr in range(P):
                    base = hdl.buffer_ptrs[pr] + off + 2 * self.rank * w      # my column slot in rank pr's buffer
                    col = 0
                    for i, n in enumerate(names):
                        g.y_peer[pr][i] = base + 2 * col
                        col += self.blocks[li][n]["N"]
                    g.done_peer[pr] = hdl.buffer_ptrs[pr] + 4 * li_flat
                g.wait_flag = prev_flag
                g.epoch = self.epoch.data_ptr()
                prev_flag = buf.data_ptr() + 4 * li_flat
                row.append((g, buf[off:off + 2 * m * P * w].view(torch.float16).view(m, P * w)))
                off += 2 * m * P * w
                li_flat += 1
            self.fused_gather.append(row)
        self._last_flag = prev_flag
        hdl.barrier()
        self.pg = None

    def _gather(self, li, gi):
        if self.pg is not None:
            import torch.distributed as dist
            dist.all_gather_into_tensor(self.grp_full[li][gi].view(-1), self.grp_local[li][gi].view(-1), group=self.pg)

    # ---- accounting -----------------------------------------------------------------------------
    def algorithmic_bytes_per_step(self) -> int:
        m, r, G = self.batch, self.r, self.G
        tot = 0
        for blk in self.blocks:
            for n, t in blk.items():
                N, K = t["N"], t["qweight"].shape[1]
                tot += N * (K - r) // 2 + 4 * N * ((K - r) // G) + 2 * N * r + 2 * K * m + 2 * N * m
        return tot

    def launches_per_step(self) -> int:
        if self.program is not None:
            return 1
        return (4 if self.fused else 7) * self.nlayers

    # ---- one token ------------------------------------------------------------------------------
    def _part(self, t, y):
        return {"qweight": t["qweight"], "scales": t["scales"], "scaled_zeros": t["scaled_zeros"],
                "oweight": t["oweight_interleaved"], "bias": t.get("bias"), "N": t["N"], "y": y}

    def step_eager(self):
        if self.program is not None:
            self.program.run()
            return self.result()
        m, r, G, h, f = self.batch, self.r, self.G, self.h, self.f
        lay = _lib.OW_INTERLEAVED
        fg = getattr(self, "fused_gather", None)
        if fg is not None:
            self.epoch.add_(1)                     # one step: every arrival counter advances by `world`
        for li, (blk, out) in enumerate(zip(self.blocks, self.out)):
            for gi, names in enumerate(self.groups):
                x = self.x_f if names[0] == "down" else self.x_h
                gather = blk["o"].get("reorder_ids32") if names[0] == "o" else None
                if fg is not None:
                    qeft_cuda.gemv_w4_multi_gather(x, [self._part(blk[n], None) for n in names], m, x.shape[-1], r, G,
                                                   fg[li][gi][0], ow_layout=lay, x_gather=gather, pdl=self.pdl)
                    if li == len(self.blocks) - 1 and gi == len(self.groups) - 1:
                        # the step's result is the gathered buffer of the last launch: order whatever follows in the stream
                        # (the copy to the host) after the arrival of every rank's slice
                        qeft_cuda.gather_wait(self._last_flag, self.epoch, self.world)
                    continue
                if self.fused:
                    qeft_cuda.gemv_w4_multi(x, [self._part(blk[n], out[n]) for n in names], m, x.shape[-1], r, G,
                                            ow_layout=lay, x_gather=gather, pdl=self.pdl)
                else:
                    for n in names:
                        t = blk[n]
                        qeft_cuda.gemv_w4(x, t["qweight"], t["scales"], t["scaled_zeros"], t["oweight_interleaved"], m,
                                          t["N"], x.shape[-1], G, ow_layout=lay, out=out[n], pdl=self.pdl,
                                          x_gather=gather)
                self._gather(li, gi)
        return self.result()

    def result(self):
        """The tensor a step produces: the last projection's output (the GATHERED [1, N] row when sharded)."""
        if getattr(self, "chain", None) and self.program is not None:
            return self.chain[-1]["out"]
        if getattr(self, "fused_gather", None) is not None:
            return self.fused_gather[-1][-1][1]
        if self.pg is not None:
            return self.grp_full[-1][-1].view(1, -1)
        return self.out[-1]["down"]

    def capture(self):
        """Capture one token into a CUDA graph (after a warm-up run on a side stream)."""
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            self.step_eager()
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.step_eager()
        self.graph = g
        return g

    def step(self):
        if self.graph is None:
            return self.step_eager()
        self.graph.replay()
        return self.result()

    def step_from_host(self, x_host_h: torch.Tensor, x_host_f: torch.Tensor, y_host: torch.Tensor):
        """The end-to-end call: pinned host activations in, last projection's output back on the host."""
        self.x_h.copy_(x_host_h, non_blocking=True)
        self.x_f.copy_(x_host_f, non_blocking=True)
        y = self.step()
        y_host.copy_(y, non_blocking=True)
        return y_host
