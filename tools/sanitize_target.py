"""A small pass over the newer kernels for compute-sanitizer (memcheck):
  compute-sanitizer --tool memcheck python tools/sanitize_target.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from meshode_b200 import capi, engine  # noqa: E402
from meshode_b200 import pyDeform as pd  # noqa: E402
from meshode_b200.synth import synth_pair  # noqa: E402

dev = "cuda:0"
pairs = [tuple(torch.from_numpy(a) for a in synth_pair(i, 700 + 300 * i, 600)) for i in range(3)]
b = engine.PairBatch(pairs, 32, device=dev)
b.deform(iters=25, schedule="cta")      # fused exact loop, one CTA per pair (tensor-memory records, repeat flags)
b.deform(iters=25, schedule="cluster")  # the same loop on thread-block clusters (double-buffered replicas over DSMEM)
b.deform(iters=5, exact=False)          # fast loop
V = b.finalize()
b.release()
srcV, srcF, tarV, tarF = synth_pair(9, 500, 500)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
pid = pd.InitializeDeformTemplate(t(tarV), t(tarF), 0, 32)
sV = t(srcV)
pd.NormalizeByTemplate(sV, pid)
pd.StoreRigidityInformation(sV, t(srcF), pid)
pd.RigidEdgeLoss_backward(sV, t(srcF), pid)          # per-incidence records
pd.LossForwardBackward(sV, pid, pid)
V0 = sV.double()
a = srcF.reshape(-1); bb = np.roll(srcF, -1, axis=1).reshape(-1)
I = t(np.stack([a, bb], 1).astype(np.int32)); rest = V0[I[:, 0].long()] - V0[I[:, 1].long()]
for kind in (capi.CERES_EDGE, capi.CERES_ROT_EDGE):
    Vv = V0.clone(); R = torch.zeros_like(Vv)
    print(pd.CeresSolve(pid, kind, Vv, R, I, rest, 1.0, max_iterations=4)["final_cost"])
pd.DestroyTemplate(pid)
torch.cuda.synchronize()
print("sanitize target done")
