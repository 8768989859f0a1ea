#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on its config, one JSON line on stdout (rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N = 1 workload (BASELINE.json configs[1]): Llama-2-7B-shape, w4 g128 r128, batch 1 decode: one "step" is
one token through the 224 packed QuantLinear GEMVs of the 32 decoder blocks (3.701 GB of algorithmic bytes,
far larger than the 126 MB L2, so every step streams its weights from HBM).
N > 1 workload (configs[4]): Llama-2-70B shapes, every linear column-sharded over the N ranks, the all-gather of
each launch's output fused into the GEMV epilogue over NVLink (`--gather nccl`: one NCCL all-gather per launch);
strong scaling (total work fixed).  The line also carries `prefill_sharded`: the same shapes at M = 2048 tokens.

`--impl reference`: the reference has no CPU implementation of this path (every forward calls its CUDA
extension).  When `oracle/_ref/qeft_cuda_ref.so` (the reference's own kernels, recompiled for sm_100a by
oracle/build_ref.py) and a GPU are present, this arm runs the reference's decode path as the reference runs it: per
decoder block seven `gemv_4bit_qeft` calls plus o_proj's `index_select` (qeft/qlinear.py:244-304), eager launches on
the legacy default stream, same shapes and byte accounting as our arm.  Otherwise it times the oracle's restatement of
the reference arithmetic ("torch dequant+matmul on CPU", BASELINE.json configs[0]) with all host threads on a bounded
sample: one decoder block per step.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "QuantLinear GEMV HBM GB/s & decode tok/s, GEMM TFLOP/s (Llama-2-7B w4g128r128)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default=None, help="7b | 13b | 70b (default: 7b at N=1, 70b at N>1)")
    ap.add_argument("--layers", type=int, default=None, help="debug: fewer decoder blocks (INVALID as a bench number)")
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-fused", action="store_true")
    ap.add_argument("--no-pdl", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="program", choices=["program", "fused", "nccl"],
                    help="N > 1: the all-gather fused into the GEMV epilogue (peer stores over NVLink + arrival counters; "
                         "default, measured 11-22 %% faster than NCCL at N = 2 and 4) or one NCCL all-gather per launch group")
    ap.add_argument("--cpu-port", action="store_true", help="--impl reference: time the oracle's CPU port even when "
                                                             "the reference's own kernels (oracle/_ref) are available")
    ap.add_argument("--no-gemm", action="store_true", help="skip the prefill GEMM / backward extras (M=2048 TFLOP/s)")
    ap.add_argument("--no-program", action="store_true",
                    help="N = 1: the round-1 chain of 4 x layers PDL launches instead of the persistent decode program")
    ap.add_argument("--no-single", action="store_true", help="N > 1: skip the single-GPU run of the same workload on rank 0")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the gathered-output parity check before timing")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu_index = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()   # exact PID we started
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [c.strip() for c in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops", 1590.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(avg_alg_bytes_per_launch, kernel="gemv"):
    """DRAM bytes per launch from the committed `ncu --set full` capture of the dominant kernel
    (profiles/<kernel>_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum and the algorithmic bytes of the
    profiled launch): the measured traffic / algorithmic ratio applied to this run's algorithmic bytes per launch.
    It is a figure FROM THE PROFILE, not of this run (bench.py never runs under ncu)."""
    p = os.path.join(ROOT, "profiles", f"{kernel}_traffic.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        d = json.load(f)
    return d["dram_bytes"] / d["algorithmic_bytes"] * avg_alg_bytes_per_launch


# --------------------------------------------------------------------------------------------------
def cpu_block_baseline(model="7b", reps=1, warm=0, threads=None):
    """Oracle port of the reference arithmetic on the host: one decoder block, batch 1 (bounded sample)."""
    import numpy as np
    import torch
    import oracle
    from qeft_b200.synth import decoder_linears

    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    layers = []
    for i, (name, N, K) in enumerate(decoder_linears(model)):
        layers.append(oracle.synth_layer(N, K, seed=1000 + i))
    rng = np.random.default_rng(0)
    xs = [rng.standard_normal((1, L["K"])).astype(np.float16) for L in layers]
    nbytes = sum(oracle.gemv_algorithmic_bytes(L["N"], L["K"]) for L in layers)
    times = []
    for it in range(warm + reps):
        t0 = time.perf_counter()
        for L, x in zip(layers, xs):
            oracle.cpu_dequant_matmul(x, L)
        dt = time.perf_counter() - t0
        if it >= warm:
            times.append(dt)
    return nbytes, times, threads


def reference_kernels_arm(args, ref):
    """The reference's own CUDA kernels (oracle/_ref) on this GPU: one decode token = 7 gemv_4bit_qeft per block."""
    import torch

    from qeft_b200.synth import LLAMA_SHAPES       # shape table only; no kernel of ours runs in this arm

    torch.cuda.set_device(0)
    model = args.model or ("7b" if args.gpus == 1 else "70b")
    h, f, nl, kv = LLAMA_SHAPES[model]
    nl = args.layers or nl
    r, G, m = 128, 128, 1
    gen = torch.Generator(device="cuda")
    gen.manual_seed(0)
    names = (("q", h, h), ("k", kv, h), ("v", kv, h), ("o", h, h), ("gate", f, h), ("up", f, h), ("down", h, f))
    blocks, nbytes = [], 0
    for _ in range(nl):
        blk = []
        for name, N, K in names:
            zero = torch.randint(0, 16, (K // G, N), device="cuda", generator=gen).float()
            scale = (torch.rand((K // G, N), device="cuda", generator=gen) * 0.010 + 0.002).half()
            t = {"qweight": torch.randint(-32768, 32768, (N // 4, K), dtype=torch.int16, device="cuda", generator=gen),
                 "scales": scale, "scaled_zeros": (-(zero * scale.float())).half(),
                 "oweight_interleaved": (torch.randn((N // 2, 2 * r), device="cuda", generator=gen) * 0.02).half(),
                 "N": N, "K": K}
            if name == "o":      # qlinear.py:275: index_select(x, -1, reorder_ids) ahead of the GEMV
                t["reorder_ids"] = torch.randperm(K, device="cuda", generator=gen)
            blk.append(t)
            nbytes += N * (K - r) // 2 + 4 * N * ((K - r) // G) + 2 * N * r + 2 * K * m + 2 * N * m
        blocks.append(blk)
    x_h = torch.randn((m, h), device="cuda", generator=gen).half()
    x_f = torch.randn((m, f), device="cuda", generator=gen).half()

    def step():
        y = None
        for blk in blocks:
            for t in blk:
                x = x_f if t["K"] == f else x_h
                if "reorder_ids" in t:
                    x = torch.index_select(x, -1, t["reorder_ids"])
                y = ref.gemv_4bit_qeft(x, t["qweight"], t["scales"], t["scaled_zeros"], t["oweight_interleaved"],
                                       m, t["N"], t["K"], G)
        return y

    def timed(fn, steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    for _ in range(max(3, args.warmup)):
        step()
    sampler = ClockSampler(0)
    sampler.start()
    ms = timed(step, args.steps)
    clocks = sampler.stop()
    xh, xf = torch.randn(x_h.shape).half().pin_memory(), torch.randn(x_f.shape).half().pin_memory()
    yh = torch.empty((m, h), dtype=torch.float16).pin_memory()

    def step_from_host():
        x_h.copy_(xh, non_blocking=True)
        x_f.copy_(xf, non_blocking=True)
        yh.copy_(step(), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(3):
        step_from_host()
    ms_e2e = timed(step_from_host, args.steps)
    gbs, gbs_e2e = nbytes / ms / 1e6, nbytes / ms_e2e / 1e6
    what = (f"the reference's own kernels (oracle/_ref/qeft_cuda_ref.so: gemv_cuda_qeft.cu recompiled for sm_100a, "
            f"launch configuration fixed, see oracle/build_ref.py) on this GPU, {7 * nl} eager gemv_4bit_qeft launches per "
            f"token on the legacy default stream; the reference has no CPU implementation of this path")
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": {"workload": f"llama2-{model} decode b1: {nl} decoder blocks x 7 packed QuantLinear (w4 g128 r128)"
                               + (" on ONE GPU (the reference has no column-sharded mode)" if args.gpus > 1 else ""),
                   "algorithmic_bytes_per_step": int(nbytes), "device": torch.cuda.get_device_name(0),
                   "l2_policy": "inputs larger than L2"},
        "decode_tok_s": 1e3 / ms, "clocks": clocks,
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": 1, "kind": "reference", "sample": what},
        "e2e": {"value": gbs_e2e, "unit": "GB/s", "tok_s": 1e3 / ms_e2e,
                "h2d_bytes_per_step": int(2 * (xh.numel() + xf.numel())), "d2h_bytes_per_step": int(2 * yh.numel())},
        "gpu_launches": 0, "reference_gpu_launches": int(7 * nl * args.steps),
    }
    if args.layers:
        line["config"]["INVALID"] = "reduced layer count (debug run)"
    print(json.dumps(line), flush=True)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not args.cpu_port:
        try:
            import torch
            from oracle import build_ref
            ref = build_ref.load() if torch.cuda.is_available() else None
        except Exception as e:  # noqa: BLE001  (a module built against another libtorch: fall back to the CPU port)
            print(f"bench.py: oracle/_ref not usable ({e}); timing the CPU port", file=sys.stderr)
            ref = None
        if ref is not None:
            return reference_kernels_arm(args, ref)
    model = args.model or "7b"
    from qeft_b200.synth import LLAMA_SHAPES
    nl = LLAMA_SHAPES[model][2]
    nbytes, times, threads = cpu_block_baseline(model, reps=max(1, args.steps), warm=args.warmup)
    t = statistics.median(times)
    gbs = nbytes / t / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak",
        "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": {"workload": (f"llama2-{model} decode b1: {nl} decoder blocks x 7 packed QuantLinear (w4 g128 r128)"),
                   "sample": f"1 of {nl} decoder blocks per step (7 dequant+matmul calls on the host CPU); tok/s extrapolated x{nl}",
                   "algorithmic_bytes_per_step": int(nbytes)},
        "decode_tok_s": 1.0 / (t * nl),
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": threads, "kind": "port",
                         "sample": f"one llama2-{model} decoder block (7 linears, {nbytes} algorithmic bytes) per step, "
                                   f"median of {len(times)}"},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
def program_parity_check(stack, world, rank):
    """N > 1, sharded decode program, before the timed loop: one run, then EVERY rank compares the six gathered activation
    rows of the first and the last decoder block with the UNSHARDED chain program run on that block's all-gathered weights
    from the same block input (the sharded run's own gathered row): bit-equal (a row's arithmetic does not depend on how
    the rows are partitioned).  Returns the dict for the JSON line; the caller exits non-zero on a mismatch."""
    import torch
    import torch.distributed as dist
    stack.program.run()
    torch.cuda.synchronize()
    dist.barrier()
    nl = stack.nlayers
    bad, checked, worst = 0, 0, 0.0
    for li in sorted({0, nl - 1}):
        twin = stack.gathered_twin(layers=[li])
        hidden = stack.x_h if li == 0 else stack.chain[li - 1]["out"].clone()
        twin.enable_chain_program(dataflow=True, ln=[stack.ln[li]], hidden=hidden)
        twin.program.run()
        torch.cuda.synchronize()
        for name in ("q", "k", "v", "h2", "act", "out"):
            a, b = stack.chain[li][name], twin.chain[0][name]
            checked += 1
            if not torch.equal(a.view(torch.int16), b.view(torch.int16)):
                bad += 1
                worst = max(worst, float((a.float() - b.float()).abs().max()))
        del twin
        torch.cuda.empty_cache()
    res = torch.tensor([bad], dtype=torch.int64, device="cuda")
    dist.all_reduce(res, op=dist.ReduceOp.SUM)
    return {"groups": checked, "ok": bool(res.item() == 0), "mode": "program", "mismatches_all_ranks": int(res.item()),
            "max_abs_diff_this_rank": worst,
            "what": "q, k, v, o+residual, SiLU(gate)*up, down+residual rows of the first and last block, every rank: bit-equal "
                    "to the unsharded chain program on the all-gathered weights, from the same block input"}


def sharded_parity_check(stack, world, rank, gather_mode):
    """N > 1, before the timed loop: one eager step in the active exchange mode, then EVERY rank checks the gathered
    buffer of all four launch groups of the first and of the last decoder block against
      (a) this rank's local kernel outputs all-gathered by NCCL (same kernel, same row partition): bit-equal, and
      (b) the UNSHARDED kernel (qeft_gemv_w4_multi on the weights of all ranks, gathered here for the check, same x):
          equal to one fp16 ulp + 1e-5 rms (the row partition changes the fp32 summation order of a row's K-slices).
    Returns {"groups": n, "ok": bool, ...}; the caller exits non-zero when not ok."""
    import torch
    import torch.distributed as dist

    from qeft_b200 import _lib, qeft_cuda

    stack.step_eager()
    torch.cuda.synchronize()
    dist.barrier()
    m, r, G = stack.batch, stack.r, stack.G
    checked, bad_a, bad_b, worst_ulp = 0, 0, 0, 0.0

    def gather_cat(t, dim):
        t = t.contiguous()
        raw = t.view(torch.uint8)                        # (NCCL has no int16: gather the bytes)
        parts = [torch.empty_like(raw) for _ in range(world)]
        dist.all_gather(parts, raw)
        return torch.cat([p.view(t.dtype).view(t.shape) for p in parts], dim=dim).contiguous()

    for li in sorted({0, stack.nlayers - 1}):
        blk = stack.blocks[li]
        for gi, names in enumerate(stack.groups):
            x = stack.x_f if names[0] == "down" else stack.x_h
            gather = blk["o"].get("reorder_ids32") if names[0] == "o" else None
            if gather_mode == "fused":
                got = stack.fused_gather[li][gi][1].view(world, -1)
            else:
                got = stack.grp_full[li][gi]
            # (a) local kernel + NCCL
            local = qeft_cuda.gemv_w4_multi(x, [stack._part(blk[n], None) for n in names], m, x.shape[-1], r, G,
                                            ow_layout=_lib.OW_INTERLEAVED, x_gather=gather, pdl=False)
            loc = torch.cat([y.reshape(1, -1) for y in local], dim=1).contiguous()
            want_a = torch.empty((world, loc.shape[1]), dtype=loc.dtype, device=loc.device)
            dist.all_gather_into_tensor(want_a.view(-1), loc.view(-1))
            bad_a += int(not torch.equal(want_a, got))
            # (b) the unsharded kernel on everybody's weights
            full_parts = []
            for n in names:
                t = blk[n]
                full_parts.append({"qweight": gather_cat(t["qweight"], 0), "scales": gather_cat(t["scales"], 1),
                                   "scaled_zeros": gather_cat(t["scaled_zeros"], 1),
                                   "oweight": gather_cat(t["oweight_interleaved"], 0), "N": t["N"] * world})
            full = qeft_cuda.gemv_w4_multi(x, full_parts, m, x.shape[-1], r, G, ow_layout=_lib.OW_INTERLEAVED,
                                           x_gather=gather, pdl=False)
            torch.cuda.synchronize()
            want_b = torch.cat([torch.cat([y.reshape(-1)[p * blk[n]["N"]:(p + 1) * blk[n]["N"]] for y, n in zip(full, names)])
                                for p in range(world)]).view(world, -1)
            d = (want_b.float() - got.float()).abs()
            # one fp16 ulp at |want| (a rounding flip) + the fp32 summation-order noise of a row (~1e-6 of the rows' rms)
            tol = torch.clamp(want_b.float().abs(), min=2.0 ** -14) * 2.0 ** -10 + 1e-5 * want_b.float().pow(2).mean().sqrt()
            worst = float((d / tol).max())
            worst_ulp = max(worst_ulp, worst)
            bad_b += int(worst > 1.0)
            del full_parts, full
            checked += 1
    torch.cuda.empty_cache()
    res = torch.tensor([bad_a, bad_b], dtype=torch.int64, device="cuda")
    dist.all_reduce(res, op=dist.ReduceOp.SUM)
    wu = torch.tensor([worst_ulp], dtype=torch.float64, device="cuda")
    dist.all_reduce(wu, op=dist.ReduceOp.MAX)
    return {"groups": checked, "ok": bool(res.sum().item() == 0), "mode": gather_mode,
            "mismatch_vs_local_plus_nccl": int(res[0].item()), "groups_off_by_more_than_1ulp_vs_unsharded": int(res[1].item()),
            "worst_ulp_vs_unsharded": wu.item(),
            "what": "all 4 launch groups of the first and last block, every rank; bit-equal to local kernel + NCCL all-gather, "
                    "<= 1 fp16 ulp from the unsharded kernel on the gathered weights"}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from qeft_b200 import qeft_cuda
    from qeft_b200.decode import PackedDecoderStack

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; qeft_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    model = args.model or ("7b" if world == 1 else "70b")
    stack = PackedDecoderStack(model, layers=args.layers, fused=not args.no_fused, pdl=not args.no_pdl,
                               shard=(rank, world), batch=args.batch, device=f"cuda:{local}", fast_synth=True)
    gather_mode = args.gather
    if world > 1:
        if gather_mode == "program" and (args.no_program or args.batch != 1):
            gather_mode = "fused"
        if gather_mode == "program":
            try:
                # one persistent launch per rank and token, the all-gathers inside the kernel (csrc/decode_w4.cu)
                stack.enable_sharded_chain_program(dist.group.WORLD)
                stack.step_eager()
                torch.cuda.synchronize()
            except Exception as e:  # noqa: BLE001
                print(f"bench.py: sharded decode program unavailable ({type(e).__name__}: {e}); using the fused-gather chain",
                      file=sys.stderr)
                stack.program = None
                stack.chain = None
                gather_mode = "fused"
        if gather_mode == "fused":
            try:
                stack.enable_fused_gather(dist.group.WORLD)
            except Exception as e:  # noqa: BLE001  (no peer-mapped symmetric memory on this box: NCCL still works)
                print(f"bench.py: fused gather unavailable ({type(e).__name__}: {e}); using NCCL all-gather", file=sys.stderr)
                gather_mode = "nccl"
                stack.fused_gather = None
        if gather_mode == "nccl":
            stack.enable_allgather(dist.group.WORLD)
    parity = None
    if world > 1 and not args.no_parity:
        parity = program_parity_check(stack, world, rank) if gather_mode == "program" else \
            sharded_parity_check(stack, world, rank, gather_mode)
        if not parity["ok"]:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "n_gpus": world, "parity_checked": parity,
                                  "error": "gathered outputs differ from the unsharded kernel"}), flush=True)
            os._exit(3)
    use_program = world == 1 and not args.no_program and not args.no_fused and args.batch <= 2
    sharded_program = world > 1 and gather_mode == "program"
    if use_program:
        try:
            stack.enable_program()   # one persistent cooperative launch per token (csrc/decode_w4.cu)
            stack.step_eager()
            torch.cuda.synchronize()
        except RuntimeError as e:    # (a box that cannot co-schedule 148 CTAs: the launch chain still measures the path)
            print(f"bench.py: decode program unavailable ({e}); timing the launch chain", file=sys.stderr)
            stack.program = None
            use_program = False
    if not args.no_graph:
        stack.capture()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        stack.step()
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = qeft_cuda.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        stack.step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = qeft_cuda.launch_count() - l0
    if stack.graph is not None:
        launches = stack.launches_per_step() * args.steps      # graph replays do not pass through the C ABI
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end: pinned host activations in, result back on the host, every step ------------------
    xh = torch.randn(stack.x_h.shape).half().pin_memory()
    xf = torch.randn(stack.x_f.shape).half().pin_memory()
    yh = torch.empty(stack.result().shape, dtype=torch.float16).pin_memory()
    for _ in range(3):
        stack.step_from_host(xh, xf, yh)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        stack.step_from_host(xh, xf, yh)
        torch.cuda.current_stream().synchronize()       # the caller reads y before the next token
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    nbytes = stack.algorithmic_bytes_per_step()
    launches_per_step = stack.launches_per_step()
    workload_layers, workload_fused, workload_pdl, workload_graph = stack.nlayers, stack.fused, stack.pdl, stack.graph is not None
    tot = torch.tensor([nbytes], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    nbytes_all = tot.item()

    extra = {}
    if use_program and rank == 0:
        # beside the headline: the same token (a) as one launch per decoder block (4 stages each: where attention would sit
        # between launches in a full decoder) and (b) as the round-1 chain of 4 x layers PDL launches in a CUDA graph
        def time_steps(fn, n):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / n

        prog, nl = stack.program, stack.nlayers

        def per_block():
            for li in range(nl):
                prog.run(4 * li, 4 * li + 4)

        ms_blk = time_steps(per_block, args.steps)
        stack.program, stack.graph = None, None
        stack.capture()
        ms_chain = time_steps(stack.step, args.steps)
        stack.graph, stack.program = None, prog
        nb1 = stack.algorithmic_bytes_per_step()
        extra["decode_variants"] = {
            "one_launch_per_block": {"launches_per_token": nl, "ms_per_token": ms_blk, "GBps": nb1 / ms_blk / 1e6},
            "round1_launch_chain": {"launches_per_token": 4 * nl, "ms_per_token": ms_chain, "GBps": nb1 / ms_chain / 1e6,
                                    "what": "gemv_w4_kernel, 4 PDL launches per block, CUDA graph (round-1 headline path)"},
        }
    if not args.no_gemm and rank == 0 and world == 1:
        del stack
        torch.cuda.empty_cache()
        extra["prefill_finetune"] = bench_gemm(model)
        if model == "7b":
            # configs[3]: the 13B shapes of the fine-tuning step (fwd + dX + dOW per linear, M = 2048)
            extra["finetune_13b"] = bench_gemm("13b")
        stack = None
    if world > 1 and rank == 0 and not args.no_single:
        # strong-scaling reference: the SAME workload (all layers unsharded) on this rank's GPU alone, through our kernels
        try:
            single = PackedDecoderStack(model, layers=args.layers, fused=True, pdl=True, shard=(0, 1), batch=args.batch,
                                        device=f"cuda:{local}", fast_synth=True)

            def t_steps(fn, n=5):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(n):
                    fn()
                b.record()
                torch.cuda.synchronize()
                return a.elapsed_time(b) / n

            nb = single.algorithmic_bytes_per_step()
            single.capture()
            ms_chain1 = t_steps(single.step)
            res = {"ms_per_step_launch_chain": ms_chain1, "GBps_launch_chain": nb / ms_chain1 / 1e6}
            try:
                single.graph = None
                single.enable_program()
                ms_prog1 = t_steps(single.step_eager)
                res.update({"ms_per_step_program": ms_prog1, "GBps_program": nb / ms_prog1 / 1e6})
            except RuntimeError as e:
                res["program"] = f"unavailable: {e}"
            best = min(v for k, v in res.items() if k.startswith("ms_per_step"))
            res.update({"ms_per_step": best, "speedup_of_this_run": best / (ms / args.steps),
                        "strong_scaling_efficiency": best / (ms / args.steps) / world,
                        "what": f"llama2-{model}, all layers unsharded on one B200 (rank 0), same kernels"})
            extra["single_gpu_same_workload"] = res
            del single
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            extra["single_gpu_same_workload"] = {"unavailable": f"{type(e).__name__}: {e}"}
    if not args.no_gemm and world > 1:
        # configs[4], prefill half: the same shapes, M = 2048 tokens, every linear column-sharded; all three exchanges
        stack.graph = None
        del stack
        stack = None
        torch.cuda.empty_cache()
        pre = bench_prefill_sharded(model, rank, world, local, barrier, layers=args.layers)
        if rank == 0:
            extra["prefill_sharded"] = pre

    if rank == 0:
        ms_step = ms / args.steps
        gbs = nbytes_all / (ms_step * 1e-3) / 1e9
        gbs_e2e = nbytes_all / (ms_e2e / args.steps * 1e-3) / 1e9
        hbm_peak, _, peak_src = measured_peaks()
        per_gpu = gbs / world
        line = {
            "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak" if world == 1 else "strong",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {
                "workload": (f"llama2-{model} decode b{args.batch}: {workload_layers} decoder blocks x 7 packed QuantLinear "
                             f"(w4 g128 r128) = "
                             + (f"{4 * workload_layers} dependent GEMV stages/token in ONE persistent cooperative launch "
                                f"(gpu-scope barrier between stages)" if use_program else
                                (f"{4 * workload_layers} chained stages/token (RMSNorm, SiLU*mul, residual fused; every stage reads "
                                 f"the previous one's output) in ONE persistent cooperative launch per rank" if sharded_program
                                 else f"{launches_per_step} GEMV launches/token"))
                             + (f", column-sharded over {world} ranks, all-gather "
                                + ("inside the kernel: the epilogue stores each slice into every rank's gathered row over NVLink, "
                                   "the next stage polls the elements as they arrive (no collective, fence or counter per stage)"
                                   if sharded_program else
                                   ("fused into the GEMV epilogue (peer stores over NVLink)" if gather_mode == "fused" else "by NCCL"))
                                if world > 1 else "")),
                "algorithmic_bytes_per_step": int(nbytes_all), "l2_policy": "inputs larger than L2 (3.7 GB of weights streamed per step)",
                "cuda_graph": workload_graph, "fused_qkv_gateup": workload_fused, "pdl": workload_pdl,
                "layers": workload_layers, "persistent_program": bool(use_program or sharded_program),
            },
            "decode_tok_s": args.batch * 1e3 / ms_step,
            "decode_tok_s_note": (f"packed linears only ({7 * workload_layers} QuantLinear of the {workload_layers} blocks"
                                  + (", chained: each stage reads the previous one's gathered output, o_proj takes the q projection "
                                     "in place of the attention output" if sharded_program else " fed from fixed activation buffers")
                                  + "); attention, embeddings and lm_head are not part of this path (SURVEY.md 8)"),
            "clocks": clocks,
            "e2e": {"value": gbs_e2e, "unit": "GB/s", "tok_s": args.batch * 1e3 / (ms_e2e / args.steps),
                    "h2d_bytes_per_step": int(xh.numel() * 2 + xf.numel() * 2), "d2h_bytes_per_step": int(yh.numel() * 2)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": per_gpu, "peak": hbm_peak, "unit": "GB/s", "frac": per_gpu / hbm_peak,
                         "frac_of_nominal_8TBs": per_gpu / 8000.0, "peak_source": peak_src,
                         "kernel": "decode_w4_kernel" if (use_program or sharded_program) else "gemv_w4_kernel",
                         "traffic": ncu_traffic(nbytes_all / world / launches_per_step, "decode" if (use_program or sharded_program) else "gemv"),
                         "traffic_unit": "bytes per launch = traffic_ratio_from_profile x algorithmic bytes of this run's launch "
                                         "(ratio: dram__bytes_read.sum + dram__bytes_write.sum of the committed ncu --set full "
                                         "capture of this kernel / its algorithmic bytes; profiles/*_traffic.json)",
                         "traffic_ratio_from_profile": ncu_traffic(1.0, "decode" if (use_program or sharded_program) else "gemv"),
                         "algorithmic_bytes_per_launch": nbytes_all / world / launches_per_step},
        }
        if parity is not None:
            line["parity_checked"] = parity
        if args.layers is not None:
            line["config"]["INVALID"] = "reduced layer count (debug run)"
        if world == 1 and not args.no_cpu_baseline:
            nb, times, threads = cpu_block_baseline(model if model != "70b" else "7b", reps=2, warm=1)
            tmed = statistics.median(times)
            line["cpu_baseline"] = {"value": nb / tmed / 1e9, "unit": "GB/s", "cores": threads, "kind": "port",
                                    "sample": f"one llama2-{model} decoder block (7 dequant+matmul calls, {nb} algorithmic bytes), "
                                              f"median of {len(times)} after 1 warm-up; {tmed:.3f} s per block",
                                    "tok_s_extrapolated": 1.0 / (tmed * workload_layers)}
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        # a captured graph keeps NCCL work alive: drop it before the communicator, and never let teardown hang the box
        import threading
        threading.Timer(30.0, lambda: os._exit(0)).start()
        sys.stdout.flush()
        if stack is not None:
            stack.graph = None
        del stack
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        os._exit(0)


def bench_prefill_sharded(model, rank, world, local, barrier, M=2048, steps=3, layers=None):
    """70B-shape prefill, M tokens, N-sharded over `world` ranks: TFLOP/s of the whole job (max-over-ranks time) with
    no exchange (each rank keeps its slab), with the all-gather fused into the GEMM epilogue, and with NCCL."""
    import torch
    import torch.distributed as dist

    from qeft_b200.prefill import PackedPrefillStack

    _, tpeak, src = measured_peaks()
    st = PackedPrefillStack(model, M=M, layers=layers, shard=(rank, world), device=f"cuda:{local}", fast_synth=True)
    flops_all = st.flops_per_step() * world
    out = {"M": M, "layers": st.nlayers, "launches_per_step": st.launches_per_step(), "flops_per_step": flops_all,
           "tensor_peak_TFLOPs_per_gpu": tpeak, "peak_source": src,
           "sent_bytes_per_rank_per_step": st.gathered_bytes_per_step(), "modes": {}}

    def timed():
        st.step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            st.step()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    for mode in ("local", "fused", "fused_multicast", "nccl"):
        if mode == "fused":
            st.enable_fused_gather(dist.group.WORLD, multicast=False)
        elif mode == "fused_multicast":
            st.enable_fused_gather(dist.group.WORLD, multicast=True)
            if not st.multicast:
                out["modes"][mode] = {"unavailable": "the symmetric allocation has no multicast mapping on this box"}
                continue
        elif mode == "nccl":
            st.enable_allgather(dist.group.WORLD)
        ms = timed()
        tf = flops_all / ms / 1e9
        out["modes"][mode] = {"ms_per_step": ms, "TFLOPs": tf, "TFLOPs_per_gpu": tf / world,
                              "frac_of_tensor_peak": tf / world / tpeak}
    out["note"] = ("local = no exchange (each rank keeps its [M, N/P] slab); fused = every output tile stored into every "
                   "rank's [M, N] buffer by the GEMM epilogue (peer stores over NVLink, arrival counters); fused_multicast = the "
                   "same with ONE store per tile piece to the NVLS multicast mapping (the NVSwitch replicates it); nccl = "
                   "all_gather_into_tensor + one permuting copy per launch")
    del st
    torch.cuda.empty_cache()
    return out


def bench_gemm(model, M=2048, iters=10):
    """Configs 3 and 4 of BASELINE.json beside the headline: per-shape prefill GEMM (fused outliers + bias path) and
    backward (dX, dOW) TFLOP/s at M = 2048 tokens, against the measured bf16 tensor peak."""
    import torch

    from qeft_b200 import qeft_cuda
    from qeft_b200.synth import LLAMA_SHAPES, synth_tensors

    _, tpeak, src = measured_peaks()
    h, f, _, kv = LLAMA_SHAPES[model]
    shapes = [("qkv/o_proj", h, h), ("gate/up_proj", f, h), ("down_proj", h, f)]
    out = {"M": M, "tensor_peak_TFLOPs": tpeak, "peak_source": src, "nominal_dense_f16_TFLOPs": 2250.0, "dtype": "f16", "shapes": []}

    def timeit(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / iters

    for name, N, K in shapes:
        t = synth_tensors(N, K, seed=7)
        x = torch.randn(M, K, device="cuda").half()
        dy = torch.randn(M, N, device="cuda").half()
        y = torch.empty(M, N, device="cuda", dtype=torch.float16)
        dx = torch.empty(M, K, device="cuda", dtype=torch.float16)
        dow = torch.empty(N, 128, device="cuda", dtype=torch.float32)
        xo = x[:, K - 128:].contiguous()
        flops = 2.0 * M * N * K
        tf = timeit(lambda: qeft_cuda.gemm_w4(x, t["qweight"], t["scales"], t["scaled_zeros"], t["oweight"], None, out=y, pdl=False))
        tb = timeit(lambda: qeft_cuda.gemm_w4_dx(dy, t["qweight"], t["scales"], t["scaled_zeros"], t["oweight"], K, out=dx, pdl=False))
        tw = timeit(lambda: qeft_cuda.dow(dy, xo, 128, out=dow))
        xb, owb = x.bfloat16(), t["oweight"].bfloat16()
        yb = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        tfb = timeit(lambda: qeft_cuda.gemm_w4(xb, t["qweight"], t["scales"], t["scaled_zeros"], owb, None, out=yb, pdl=False))
        del xb, owb, yb
        out["shapes"].append({"name": name, "N": N, "K": K, "fwd_bf16_TFLOPs": flops / tfb / 1e12,
                              "fwd_TFLOPs": flops / tf / 1e12, "fwd_frac": flops / tf / 1e12 / tpeak,
                              "fwd_frac_of_nominal_2250": flops / tf / 1e12 / 2250.0,
                              "dx_TFLOPs": flops / tb / 1e12, "dx_frac": flops / tb / 1e12 / tpeak,
                              "dx_frac_of_nominal_2250": flops / tb / 1e12 / 2250.0,
                              "dow_us": tw * 1e6,
                              "finetune_step_TFLOPs": (2 * flops + 2.0 * M * N * 128) / (tf + tb + tw) / 1e12})
        del t, x, dy, y, dx
        torch.cuda.empty_cache()
    return out


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
