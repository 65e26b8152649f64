"""Multi-GPU paths on real devices (skipped with fewer than 2 GPUs): z-slab sharded grid build with
one NCCL all-gather == single-GPU build bit for bit; pair-sharded deformation + final gather ==
single-GPU results bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from meshode_b200 import engine, sharding
    from meshode_b200 import pyDeform as pd
    from meshode_b200.synth import synth_mesh, synth_pair
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        # ---- z-slab sharded build of one grid: equal slabs (in-place all-gather on the template's own fields)
        #      ragged slabs (N = 51: padded all-gather) and cyclic layers (N = 48) ----------------------------
        V, F = synth_mesh(3000, 11)
        tV, tF = torch.from_numpy(V).to(dev), torch.from_numpy(F).to(dev)
        for N in (51, 48, 50):   # ragged slabs (padded all-gather), cyclic z-tile layers (48 = 4 * world * 6), equal slabs
            pid = sharding.build_template_sharded(tV, tF, N)
            g64, g32, idx = pd.GetGrid(pid)
            ref = pd.InitializeDeformTemplate(tV, tF, 0, N)
            r64, r32, ridx = pd.GetGrid(ref)
            assert torch.equal(g64, r64) and torch.equal(g32, r32) and torch.equal(idx, ridx), N
            if N != 50:
                pd.DestroyTemplate(pid); pd.DestroyTemplate(ref)
        # the assembled template serves lookups everywhere (a vertex needs slices z and z+1)
        P = torch.rand((4000, 3), device=dev)
        assert torch.equal(pd.DistanceFieldLoss_backward(P, pid), pd.DistanceFieldLoss_backward(P, ref))
        # ---- pair-sharded deformation, final gather -------------------------------------------------
        n_pairs, iters = 5, 40
        lo, hi = sharding.shard_range(n_pairs, rank, world)
        pairs = [tuple(torch.from_numpy(a) for a in synth_pair(i, 600 + 40 * i, 500)) for i in range(lo, hi)]
        b = engine.PairBatch(pairs, grid_resolution=32, device=dev)
        b.deform(iters=iters, exact=True)
        outs = sharding.gather_pair_vertices(b.finalize(), n_pairs)
        if rank == 0:
            for i in range(n_pairs):
                b1 = engine.PairBatch([tuple(torch.from_numpy(a) for a in synth_pair(i, 600 + 40 * i, 500))], 32, device=dev)
                b1.deform(iters=iters, exact=True)
                assert torch.equal(b1.finalize()[0], outs[i]), i
                b1.release()
        b.release()
        torch.cuda.synchronize()
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_gpu_slabs_and_pairs(tmp_path, pd):
    world = min(torch.cuda.device_count(), 2)
    if world < 2:
        pytest.skip("needs 2 GPUs")
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))
