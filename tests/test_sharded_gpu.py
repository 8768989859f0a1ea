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
