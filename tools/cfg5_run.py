"""cfg5 of BASELINE.json: one 256^3 distance field on a 500 000-triangle synthetic target, z-slab sharded over the
ranks (one NCCL all-gather per field), then the graph loss of cad_neural_deform2.py on the assembled field.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/cfg5_run.py [N=256] [target_verts=250002]
Prints, on rank 0, the slab build time (max over ranks), the all-gather time and the single-GPU build beside it."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from meshode_b200 import capi, sharding  # noqa: E402
from meshode_b200 import pyDeform as pd  # noqa: E402
from meshode_b200.synth import synth_mesh, synth_params, unique_edges  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nv = int(sys.argv[2]) if len(sys.argv) > 2 else 250002
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
V, F = synth_mesh(nv, 1)
tV, tF = torch.from_numpy(V).to(dev), torch.from_numpy(F).to(dev)
ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731


def sync():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def maxr(x):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


z0, z1 = sharding.slab_range(N, rank, world)
s = torch.cuda.current_stream().cuda_stream
res = {}
for rep in range(3):
    sync()
    e0, e1, e2 = ev(), ev(), ev()
    e0.record()
    pid = sharding.build_template_sharded(tV, tF, N)
    e2.record()
    sync()
    res = {"sharded_build_ms": maxr(e0.elapsed_time(e2))}
    # the slab build alone, for the split
    e0, e1 = ev(), ev()
    e0.record()
    p2 = capi.template_create_slab(tV.data_ptr(), tV.shape[0], tF.data_ptr(), tF.shape[0], N, z0, z1, s)
    e1.record()
    sync()
    res["slab_build_ms"] = maxr(e0.elapsed_time(e1))
    pd.DestroyTemplate(p2)
    if rep < 2:
        pd.DestroyTemplate(pid)
# graph loss on the assembled field (20 000-node graph, a19 / cad_neural_deform2.py)
gV, gF = synth_mesh(20000, 0, axis_scale=synth_params(1)[3])
gE = unique_edges(gF)
GV = torch.from_numpy(gV).to(dev); GE = torch.from_numpy(gE).to(dev)
pd.NormalizeByTemplate(GV, pid)
pd.StoreGraphInformation(GV, GE, pid)
moved = (GV + 1e-3 * torch.sin(37.0 * GV)).contiguous()
for _ in range(3):
    loss, grad = pd.LossForwardBackward(moved, pid, pid, 1.0, 0.5 * 0.03 ** 2)
sync()
e0, e1 = ev(), ev()
e0.record()
for _ in range(100):
    loss, grad = pd.LossForwardBackward(moved, pid, pid, 1.0, 0.5 * 0.03 ** 2)
e1.record()
sync()
res["graph_loss_us"] = maxr(e0.elapsed_time(e1) * 10.0)
res["loss"] = float(loss.item())
if rank == 0:
    full = []
    for rep in range(3):
        e0, e1 = ev(), ev()
        e0.record()
        p1 = pd.InitializeDeformTemplate(tV, tF, 0, N)
        e1.record()
        torch.cuda.synchronize()
        full.append(e0.elapsed_time(e1))
        if rep == 2:
            a, b = pd.GetGrid(p1)[0], pd.GetGrid(pid)[0]
            res["bit_identical_to_single_gpu"] = bool(torch.equal(a, b))
        pd.DestroyTemplate(p1)
    res["single_gpu_build_ms"] = min(full)
    print("cfg5 N=%d triangles=%d ranks=%d: %s" % (N, F.shape[0], world, res))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
