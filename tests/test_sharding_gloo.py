"""N > 1 path on CPU: two gloo ranks shard the work exactly as bench.py does on GPUs (activations by batch rows,
weights by layer, one big tensor by channel blocks), each rank computes its share with the CPU oracle standing in
for the kernel, results are all-gathered and must equal the unsharded computation.  No collective is needed to
COMPUTE anything; the gather is verification only."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from mct_quantizers_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fq(x, scale, zp, C, inner, offset=0):
    """oracle fake-quant of a flat slice that starts at logical element `offset` of a [.., C, inner] tensor."""
    n = x.size
    ch = ((offset + np.arange(n, dtype=np.int64)) // inner) % C
    out = np.empty(n, dtype=np.float32)
    for c in np.unique(ch):
        m = ch == c
        out[m] = oracle.fq_affine(x[m], oracle.F32, scale[c:c + 1], zp[c:c + 1], 1, 1, -128, 127)
    return out


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(42)                               # same data on every rank
        # (1) activations: batch-sharded, per-tensor parameters replicated
        batch, feat = 13, 37
        x = rng.standard_normal((batch, feat)).astype(np.float32)
        s, z = np.array([0.031], np.float32), np.zeros(1, np.int32)
        b0, b1 = sharding.shard_batch(batch, world, rank)
        mine = torch.from_numpy(_fq(x[b0:b1].reshape(-1), s, z, 1, 1))
        sizes = [(sharding.shard_batch(batch, world, r)[1] - sharding.shard_batch(batch, world, r)[0]) * feat for r in range(world)]
        pad = max(sizes)                                               # gloo all_gather wants equal sizes: pad, then trim
        parts = [torch.empty(pad, dtype=torch.float32) for _ in sizes]
        dist.all_gather(parts, torch.cat([mine, torch.zeros(pad - mine.numel())]))
        ok_act = np.array_equal(torch.cat([p_[:n_] for p_, n_ in zip(parts, sizes)]).numpy(), _fq(x.reshape(-1), s, z, 1, 1))

        # (2) one big weight tensor: channel-block sharded along axis 1 of [outer, C, inner]; parameters sharded identically
        shape, axis = (3, 10, 7), 1
        w = rng.standard_normal(shape).astype(np.float32)
        sc = (np.abs(rng.standard_normal(10)) * 0.02 + 0.01).astype(np.float32)
        zp = np.zeros(10, np.int32)
        slices, (c0, c1) = sharding.shard_channel_blocks(shape, axis, world, rank)
        sub = np.ascontiguousarray(w[slices])
        y_sub = oracle.fq_affine(sub, oracle.F32, sc[c0:c1], zp[c0:c1], c1 - c0, 7, -128, 127)
        full = oracle.fq_affine(w, oracle.F32, sc, zp, 10, 7, -128, 127)
        ok_w = np.array_equal(y_sub, full[slices])

        # (3) flat range sharding with elem_offset (what the host staging path and the size sweep use)
        n = w.size
        a0, a1 = sharding.shard_range(n, world, rank, align=8)
        ok_flat = np.array_equal(_fq(w.reshape(-1)[a0:a1], sc, zp, 10, 7, offset=a0), full.reshape(-1)[a0:a1])

        # (4) layers: every layer owned by exactly one rank, loads balanced; checksums gathered like bench.py does
        layer_sizes = [864, 288, 512, 1536, 864, 2304, 3456, 1296, 3456, 20736, 1296, 4608, 6144, 1728]
        bins = sharding.shard_layers(layer_sizes, world)
        chk = torch.tensor([sum(sharding.checksum64(torch.full((layer_sizes[i],), float(i))) for i in bins[rank])], dtype=torch.int64)
        allchk = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allchk, chk)
        want = sum(sharding.checksum64(torch.full((n_,), float(i))) for i, n_ in enumerate(layer_sizes))
        ok_layers = sum(int(c.item()) for c in allchk) == want and sorted(sum(bins, [])) == list(range(len(layer_sizes)))
        loads = [sum(layer_sizes[i] for i in b) for b in bins]
        ok_balance = max(loads) - min(loads) <= max(layer_sizes)
        if rank == 0:
            q.put((ok_act, ok_w, ok_flat, ok_layers, ok_balance))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_sharding_matches_unsharded():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    assert res == (True, True, True, True, True), res


def test_shard_helpers_edge_cases():
    assert sharding.shard_range(10, 4, 0) == (0, 3) and sharding.shard_range(10, 4, 3) == (8, 10)
    assert [sharding.shard_range(3, 8, r) for r in range(8)][3:] == [(3, 3)] * 5          # more ranks than items
    assert sharding.shard_range(100, 3, 2, align=16) == (80, 100)
    assert sum(b - a for a, b in (sharding.shard_range(1001, 8, r, align=8) for r in range(8))) == 1001
    assert sharding.shard_layers([], 2) == [[], []]
    assert sharding.shard_layers([5, 5, 5, 5], 2) == [[0, 2], [1, 3]]
    sl, (c0, c1) = sharding.shard_channel_blocks((4, 6), -1, 2, 1)
    assert sl == (slice(None), slice(3, 6)) and (c0, c1) == (3, 6)
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)
    t = torch.tensor([1.0, -0.0, 0.0])
    assert sharding.checksum64(t) != sharding.checksum64(torch.tensor([1.0, 0.0, 0.0]))      # sign of zero is visible
