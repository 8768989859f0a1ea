"""Column sharding of a packed layer over ranks: host logic, world_size 2 over gloo on the CPU.

The product has no CPU compute path, so each rank's slice is evaluated with the oracle (checker only); what is under
test is the slicing (`shard_layer_tensors`), the all-gather layout and that shards reproduce the unsharded result.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from qeft_b200.modelutils import shard_layer_tensors, shard_rows


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, N, K, r, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = oracle.synth_layer(N, K, r=r, seed=3, bias=True)
        t = {k: torch.from_numpy(np.ascontiguousarray(L[k])) for k in
             ("qweight", "scales", "scaled_zeros", "oweight", "oweight_interleaved", "bias")}
        sh = shard_layer_tensors(t, rank, world, multiple=8)
        x = np.random.default_rng(0).standard_normal((3, K)).astype(np.float16)
        # the interleaved slice is exactly pack_oweight of the plain slice (no repacking needed)
        assert np.array_equal(oracle.pack_oweight(sh["oweight"].numpy()), sh["oweight_interleaved"].numpy())
        y_local = oracle.forward(x, sh["qweight"].numpy(), sh["scales"].numpy(), sh["scaled_zeros"].numpy(),
                                 sh["oweight"].numpy(), sh["bias"].numpy())
        shape = y_local.shape
        y_local = torch.from_numpy(np.ascontiguousarray(y_local).view(np.uint8).copy())      # gloo moves bytes
        # one all-gather per launch group: gathered buffer is [world, m * width], rank-major
        gathered = torch.empty((world, y_local.numel()), dtype=torch.uint8)
        dist.all_gather_into_tensor(gathered.view(-1), y_local.view(-1))
        y = np.concatenate([gathered[p].numpy().view(np.float16).reshape(shape) for p in range(world)], axis=1)
        want = oracle.forward(x, L["qweight"], L["scales"], L["scaled_zeros"], L["oweight"], L["bias"])
        ok = np.array_equal(y.view(np.uint16), want.view(np.uint16))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def _prefill_worker(rank, world, port, N, K, r, M, q):
    """The prefill exchange's layout (prefill.py): local [M, N/P] slab -> all-gather to [P, M, N/P] -> permuting copy to
    [M, N]; and the fused variant's addressing (row pitch N, column offset rank * N/P) gives the same matrix."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = oracle.synth_layer(N, K, r=r, seed=5)
        t = {k: torch.from_numpy(np.ascontiguousarray(L[k])) for k in
             ("qweight", "scales", "scaled_zeros", "oweight", "oweight_interleaved")}
        sh = shard_layer_tensors(t, rank, world)                 # 128-row multiples, as the tensor-core tiles need
        x = np.random.default_rng(1).standard_normal((M, K)).astype(np.float16)
        y_local = oracle.forward(x, sh["qweight"].numpy(), sh["scales"].numpy(), sh["scaled_zeros"].numpy(),
                                 sh["oweight"].numpy(), None)
        assert y_local.shape == (M, N // world)
        stage = torch.empty((world, M, (N // world) * 2), dtype=torch.uint8)
        dist.all_gather_into_tensor(stage.view(-1), torch.from_numpy(np.ascontiguousarray(y_local).view(np.uint8).copy()).view(-1))
        y_full = torch.empty((M, N * 2), dtype=torch.uint8)
        y_full.view(M, world, -1).copy_(stage.permute(1, 0, 2))          # the copy enable_allgather's step() makes
        want = oracle.forward(x, L["qweight"], L["scales"], L["scaled_zeros"], L["oweight"], None)
        ok = np.array_equal(y_full.numpy().view(np.uint16), want.view(np.uint16))
        # fused addressing: element (tok, f) of the slab lands at tok * y_ld + rank * N/P + f of every rank's buffer
        flat = np.zeros(M * N, dtype=np.uint16)
        lo = rank * (N // world)
        idx = (np.arange(M)[:, None] * N + lo + np.arange(N // world)[None, :]).reshape(-1)
        flat[idx] = y_local.view(np.uint16).reshape(-1)
        ok = ok and np.array_equal(flat.reshape(M, N)[:, lo:lo + N // world], want.view(np.uint16)[:, lo:lo + N // world])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_prefill_exchange_layout():
    world, N, K, r, M = 2, 256, 256, 64, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_prefill_worker, args=(rank, world, port, N, K, r, M, q)) for rank in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


@pytest.mark.timeout(120)
def test_two_rank_shards_reproduce_the_unsharded_layer():
    world, N, K, r = 2, 64, 256, 64
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(rank, world, port, N, K, r, q)) for rank in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def test_shard_rows_rules():
    assert shard_rows(8192, 3, 8) == (3072, 4096)
    assert shard_rows(1024, 7, 8) == (896, 1024)          # 70B kv projections over 8 ranks: 128 rows each
    assert shard_rows(28672, 1, 2) == (14336, 28672)
    with pytest.raises(ValueError):
        shard_rows(1024, 0, 16)                           # 64 rows per rank: not a multiple of 128
