"""Column-sharded decode over 2 GPUs: the all-gather fused into the GEMV epilogue (peer stores + arrival counters)
gives bit-identical gathered outputs to local GEMVs + NCCL all-gather, eagerly and under CUDA-graph replay.
Needs 2 B200s on one node (skipped otherwise)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from qeft_b200.decode import PackedDecoderStack
        kw = dict(layers=2, seed=3, shard=(rank, world), device=f"cuda:{rank}")
        ref = PackedDecoderStack("7b", **kw)
        ref.enable_allgather(dist.group.WORLD)
        ref.step_eager()
        torch.cuda.synchronize()
        fused = PackedDecoderStack("7b", **kw)
        fused.enable_fused_gather(dist.group.WORLD)
        ok = True
        for mode in ("eager", "graph", "graph"):
            if mode == "graph" and fused.graph is None:
                fused.capture()
            for row in fused.fused_gather:
                for _, buf in row:
                    buf.zero_()
            torch.cuda.synchronize()
            dist.barrier()
            fused.step()
            torch.cuda.synchronize()
            dist.barrier()          # every rank's peer stores are done once every rank has synchronised
            for li in range(fused.nlayers):
                for gi in range(4):
                    want = ref.grp_full[li][gi]                       # [world, width]
                    got = fused.fused_gather[li][gi][1].view(world, -1)
                    ok = ok and torch.equal(want, got)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_fused_gather_equals_nccl_allgather():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(rank, world, port, q)) for rank in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=500) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


# ---- prefill: the all-gather fused into the tcgen05 GEMM's epilogue --------------------------------------------------
def _prefill_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    # the test is about the exchange being exact: every launch here must add its k-blocks in the same order, so the
    # split-K configurations of the plain GEMM (few tiles at M = 300; the gather launches never split) are switched off
    os.environ["QEFT_GEMM_SMALLM"] = "0"
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from qeft_b200 import qeft_cuda
        from qeft_b200.prefill import NAMES, PackedPrefillStack
        from qeft_b200.synth import synth_tensors
        shape = (512, 1024, 2, 256)                  # hidden, ffn, blocks, kv width: every N splits into 2 x k x 128 rows
        M = 300                                      # ragged: the last 256-token tile is partial
        kw = dict(M=M, seed=5, shard=(rank, world), device=f"cuda:{rank}", fast_synth=False, shard_from_full=True)
        nccl = PackedPrefillStack(shape, **kw)
        nccl.enable_allgather(dist.group.WORLD)
        nccl.step()
        torch.cuda.synchronize()
        fused = PackedPrefillStack(shape, **kw)
        ok = True
        for it in range(4):
            if it == 0:
                fused.enable_fused_gather(dist.group.WORLD, multicast=False)      # one store per rank
            elif it == 2:
                fused.enable_fused_gather(dist.group.WORLD, multicast=True)       # NVLS multicast stores, if available
            for s in range(2):
                for n in NAMES:
                    fused.y_full[s][n].zero_()
            torch.cuda.synchronize()
            dist.barrier()
            fused.step()
            torch.cuda.synchronize()
            dist.barrier()
            for li in range(2):
                for pi, n in enumerate(NAMES):
                    got = fused.y_full[li % 2][n]
                    ok = ok and torch.equal(got, nccl.y_full[li % 2][n])
                    # and against the unsharded layer on this rank
                    t = synth_tensors(fused.full[n], fused.kin[n], 128, 128, seed=5 * 100003 + li * 16 + pi,
                                      device=f"cuda:{rank}")
                    x = fused.x_f if n == "down" else fused.x_h
                    want = qeft_cuda.gemm_w4(x, t["qweight"], t["scales"], t["scaled_zeros"], t["oweight"], None)
                    ok = ok and torch.equal(got, want)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_prefill_fused_gather_equals_nccl_and_unsharded():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_prefill_worker, args=(rank, world, port, q)) for rank in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=500) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_launch_on_a_device_that_is_not_current():
    """The reference launches on the legacy default stream of the current device; this wrapper launches on the
    tensors' device whatever the current one is (the device guard is skipped only when they coincide)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import numpy as np

    import oracle
    from qeft_b200 import _lib, qeft_cuda
    N, K, r, m = 256, 512, 128, 1
    L = oracle.synth_layer(N, K, r=r, G=128, seed=3)
    x = np.random.default_rng(0).standard_normal((m, K)).astype(np.float16)
    want = oracle.forward(x, L["qweight"], L["scales"], L["scaled_zeros"], L["oweight"], None)
    torch.cuda.set_device(0)
    d1 = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to("cuda:1")  # noqa: E731
    y = qeft_cuda.gemv_w4(d1(x), d1(L["qweight"]), d1(L["scales"]), d1(L["scaled_zeros"]), d1(L["oweight_interleaved"]),
                          m, N, K, 128, ow_layout=_lib.OW_INTERLEAVED)
    xm = np.random.default_rng(1).standard_normal((40, K)).astype(np.float16)
    ym = qeft_cuda.gemm_w4(d1(xm), d1(L["qweight"]), d1(L["scales"]), d1(L["scaled_zeros"]), d1(L["oweight"]), None)
    torch.cuda.synchronize(1)
    assert torch.cuda.current_device() == 0 and y.device.index == 1 and ym.device.index == 1
    rel = lambda a, b: float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64))) / np.max(np.abs(b.astype(np.float64))))  # noqa: E731
    assert rel(y.cpu().numpy(), want) <= 1e-3
    assert rel(ym.cpu().numpy(), oracle.forward(xm, L["qweight"], L["scales"], L["scaled_zeros"], L["oweight"], None)) <= 1e-3


# ---- the acquire-then-consume path: launch i+1 READS the gathered buffer of launch i ---------------------------------
def _chain_worker(rank, world, port, q):
    """Three column-sharded GEMVs chained through their gathered buffers (x of launch i+1 = gathered y of launch i, ordered
    only by the arrival counter under programmatic dependent launch), eagerly and replayed from a CUDA graph, against
    the same chain with local kernels + NCCL all-gathers.  Exercises the coherent x loads of launches with a wait flag
    (a gathered x must not be read through the read-only path)."""
    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm_mem
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from qeft_b200 import _lib, qeft_cuda
        from qeft_b200.modelutils import shard_layer_tensors
        from qeft_b200.synth import synth_tensors
        dev = f"cuda:{rank}"
        H, r, G, nl = 1024, 128, 128, 3
        # small weights so that three chained layers stay in fp16 range
        fulls = []
        for i in range(nl):
            t = synth_tensors(H, H, r, G, seed=900 + i, device=dev)
            t["scales"] = (t["scales"].float() * 0.05).half()
            t["scaled_zeros"] = (t["scaled_zeros"].float() * 0.05).half()
            t["oweight"] = (t["oweight"].float() * 0.05).half()
            t["oweight_interleaved"] = qeft_cuda.interleave_oweight(t["oweight"])
            fulls.append(t)
        shards = [shard_layer_tensors(t, rank, world, multiple=16) for t in fulls]
        w = H // world
        gen = torch.Generator(device=dev)
        gen.manual_seed(4)
        x0 = torch.randn((1, H), device=dev, generator=gen).half()

        def part(t):
            return {"qweight": t["qweight"], "scales": t["scales"], "scaled_zeros": t["scaled_zeros"],
                    "oweight": t["oweight_interleaved"], "N": w}

        # reference chain: local kernel + NCCL all-gather
        ref = []
        x = x0
        for t in shards:
            y = qeft_cuda.gemv_w4_multi(x, [part(t)], 1, H, r, G, ow_layout=_lib.OW_INTERLEAVED, pdl=False)[0]
            full = torch.empty((1, H), dtype=torch.float16, device=dev)
            dist.all_gather_into_tensor(full.view(-1), y.reshape(-1).contiguous())
            ref.append(full)
            x = full
        torch.cuda.synchronize()

        # fused chain through symmetric memory: flags first, then the gathered buffers
        flag_bytes = 256
        buf = symm_mem.empty((flag_bytes + nl * 2 * H,), dtype=torch.uint8, device=dev)
        buf.zero_()
        hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
        epoch = torch.zeros((1,), dtype=torch.int32, device=dev)
        gathers, outs, prev = [], [], None
        for i in range(nl):
            g = _lib.Gather()
            g.nranks, g.y_ld = world, H
            off = flag_bytes + i * 2 * H
            for pr in range(world):
                g.y_peer[pr][0] = hdl.buffer_ptrs[pr] + off + 2 * rank * w
                g.done_peer[pr] = hdl.buffer_ptrs[pr] + 4 * i
            g.wait_flag = prev
            g.epoch = epoch.data_ptr()
            prev = buf.data_ptr() + 4 * i
            gathers.append(g)
            outs.append(buf[off:off + 2 * H].view(torch.float16).view(1, H))
        hdl.barrier()

        def step():
            epoch.add_(1)
            x = x0
            for t, g, o in zip(shards, gathers, outs):
                qeft_cuda.gemv_w4_multi_gather(x, [part(t)], 1, H, r, G, g, ow_layout=_lib.OW_INTERLEAVED, pdl=True)
                x = o
            qeft_cuda.gather_wait(prev, epoch, world)

        ok = True
        graph = None
        for mode in ("eager", "eager", "graph", "graph", "graph"):
            if mode == "graph" and graph is None:
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    step()
                torch.cuda.current_stream().wait_stream(s)
                torch.cuda.synchronize()
                dist.barrier()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    step()
            if mode == "eager":
                step()
            else:
                graph.replay()
            torch.cuda.synchronize()
            dist.barrier()
            for o, want in zip(outs, ref):
                ok = ok and torch.equal(o, want)
            dist.barrier()
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_chained_gathered_input_equals_nccl_chain():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_chain_worker, args=(rank, world, port, q)) for rank in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=500) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


# ---- decode: the column-sharded chain as ONE persistent launch per rank, all-gathers by data-flow inside the kernel ----
def _program_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from qeft_b200.decode import PackedDecoderStack
        st = PackedDecoderStack("7b", layers=3, seed=5, shard=(rank, world), device=f"cuda:{rank}", fast_synth=True)
        prog = st.enable_sharded_chain_program(dist.group.WORLD)
        # the unsharded chain on everybody's weights: same arithmetic per output row, so the results are bit-equal
        twin = st.gathered_twin()
        twin.enable_chain_program(dataflow=True, ln=st.ln)
        twin.program.run()
        torch.cuda.synchronize()
        ok = True
        why = ""

        def check(tag):
            nonlocal ok, why
            for li in range(st.nlayers):
                for name in ("q", "k", "v", "h2", "act", "out"):
                    a, b = st.chain[li][name], twin.chain[li][name]
                    if not torch.equal(a.view(torch.int16), b.view(torch.int16)):
                        ok = False
                        why = why or f"{tag}: block {li} {name}: {(a.float() - b.float()).abs().max().item()}"

        prog.run()
        torch.cuda.synchronize()
        dist.barrier()
        check("eager")
        for b in st.chain:
            for v in b.values():
                v.fill_(float("nan"))
        torch.cuda.synchronize()
        dist.barrier()
        prog.run()
        prog.run(0, 6)
        prog.run(6, 12)                  # sub-ranges: the second launch reads gathered rows of the first as plain inputs
        torch.cuda.synchronize()
        dist.barrier()
        check("again + sub-ranges")
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            prog.run()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        dist.barrier()
        with torch.cuda.graph(g):
            prog.run()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        dist.barrier()
        check("graph replay")
        q.put((rank, bool(ok), why))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_sharded_decode_program_equals_unsharded_chain():
    """2 ranks: the column-sharded chain program (peer stores from the epilogue, next stage polls its gathered input)
    gives on every rank exactly the gathered rows of the unsharded chain program on the all-gathered weights: eagerly,
    run after run, on sub-ranges and under CUDA-graph replay."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_program_worker, args=(rank, world, port, q)) for rank in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=500) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True, ""), (1, True, "")], res
