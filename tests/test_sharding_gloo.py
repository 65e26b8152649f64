"""The N>1 host logic on CPU: two gloo ranks partition pairs and z-slabs exactly like the NCCL
run does (same code in meshode_b200/sharding.py), and the collectives reassemble bit-identical
results.  No GPU."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _field(N):
    z, y, x = np.meshgrid(np.arange(N), np.arange(N), np.arange(N), indexing="ij")
    return (z * 10000.0 + y * 100.0 + x).astype(np.float64)


def _worker(rank, world, port, N, n_pairs, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from meshode_b200 import sharding as S
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # z-slabs: each rank contributes its slices of a known field, in three dtypes
        z0, z1 = S.slab_range(N, rank, world)
        full = torch.from_numpy(_field(N))
        for dt in (torch.float64, torch.float32, torch.int32):
            got = S.allgather_slabs(full[z0:z1].to(dt), N)
            assert got.shape == (N, N, N) and torch.equal(got, full.to(dt)), (rank, dt)
        # pairs: ragged per-pair results come back in pair order on rank 0
        lo, hi = S.shard_range(n_pairs, rank, world)
        local = [torch.full((5 + (i % 4), 3), float(i), dtype=torch.float32) + torch.arange(3) for i in range(lo, hi)]
        allv = S.gather_pair_vertices(local, n_pairs)
        if rank == 0:
            assert len(allv) == n_pairs
            for i, v in enumerate(allv):
                assert v.shape == (5 + (i % 4), 3) and torch.equal(v, torch.full((5 + (i % 4), 3), float(i)) + torch.arange(3))
        else:
            assert allv is None
        # cyclic z-tile layers: rank r owns layers r, r+world, ...; one in-place all-gather per group of `world` layers
        # assembles the field (the same indexing build_template_sharded runs over NCCL on the template's own buffers)
        L = 4
        M = 4 * L * world
        groups = S.layer_groups(M, world, L)
        assert groups == 4 and S.layer_groups(M + L, world, L) == 0
        fullM = torch.from_numpy(_field(M))
        mine = torch.full((M, M, M), -1.0, dtype=torch.float64)
        for layer in range(rank, M // L, world):
            mine[layer * L:(layer + 1) * L] = fullM[layer * L:(layer + 1) * L]
        span = L * world
        for j in range(groups):
            chunk = mine[j * span:(j + 1) * span]
            parts = [torch.empty_like(chunk[:L]) for _ in range(world)]
            dist.all_gather(parts, chunk[rank * L:(rank + 1) * L].contiguous())
            chunk.copy_(torch.cat(parts, dim=0))
        assert torch.equal(mine, fullM)
        # a problem seen by ONE rank raises on every rank instead of leaving the others inside a collective
        try:
            S._check_collectively(rank != 1, "rank 1 is unhappy")
            raised = False
        except ValueError:
            raised = True
        assert raised
        # a rank that owns no pair still takes part in the final gather
        lo1, hi1 = S.shard_range(1, rank, world)
        one = S.gather_pair_vertices([torch.ones((4, 3))] * (hi1 - lo1), 1)
        if rank == 0:
            assert len(one) == 1 and torch.equal(one[0], torch.ones((4, 3)))
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == 10.0 + world - 1
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("N,n_pairs", [(8, 7), (9, 4)])
def test_two_rank_gloo(tmp_path, N, n_pairs):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), N, n_pairs, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))


def test_partitions_cover_everything():
    from meshode_b200.sharding import shard_range, slab_range
    for n in (0, 1, 7, 64, 3625):
        for world in (1, 2, 4, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    assert [slab_range(256, k, 8) for k in range(8)] == [(32 * k, 32 * k + 32) for k in range(8)]
