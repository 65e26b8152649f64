"""Small, fixed workloads for ncu captures (never a source of bench numbers).
  python tools/prof_target.py sdf128     one 128^3 build on the 50k-triangle synthetic target (+1 warm-up)
  python tools/prof_target.py sdf64      four 64^3 builds on 10k-triangle targets
  python tools/prof_target.py deform     148 cfg4 pairs x 300 Adam iterations (+ set-up)
  python tools/prof_target.py loss       per-call kernels on the cfg4 sizes
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from meshode_b200 import engine  # noqa: E402
from meshode_b200 import pyDeform as pd  # noqa: E402
from meshode_b200.synth import synth_mesh, synth_pair  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "sdf128"
dev = "cuda:0"
if mode == "sdf128":
    V, F = synth_mesh(25002, 1)
    tV, tF = torch.from_numpy(V).to(dev), torch.from_numpy(F).to(dev)
    for _ in range(2):
        pid = pd.InitializeDeformTemplate(tV, tF, 0, 128)
        torch.cuda.synchronize()
        pd.DestroyTemplate(pid)
elif mode == "sdf64":
    for i in range(4):
        V, F = synth_mesh(5000, 2 * i + 1)
        pid = pd.InitializeDeformTemplate(torch.from_numpy(V).to(dev), torch.from_numpy(F).to(dev), 0, 64)
        torch.cuda.synchronize()
        pd.DestroyTemplate(pid)
elif mode == "deform":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 148
    it = int(sys.argv[3]) if len(sys.argv) > 3 else 300
    pairs = [tuple(torch.from_numpy(a).to(dev) for a in synth_pair(i)) for i in range(n)]
    b = engine.PairBatch(pairs, 64)
    b.deform(iters=it)
    b.finalize()
    torch.cuda.synchronize()
    b.release()
elif mode == "loss":
    sV, sF, tV, tF = [torch.from_numpy(a).to(dev) for a in synth_pair(0)]
    pid = pd.InitializeDeformTemplate(tV, tF, 0, 64)
    pd.NormalizeByTemplate(sV, pid)
    pd.StoreRigidityInformation(sV, sF, pid)
    for _ in range(3):
        pd.DistanceFieldLoss_forward(sV, pid); pd.DistanceFieldLoss_backward(sV, pid)
        pd.RigidEdgeLoss_forward(sV, sF, pid); pd.RigidEdgeLoss_backward(sV, sF, pid)
        pd.LossForwardBackward(sV, pid, pid)
    torch.cuda.synchronize()
torch.cuda.synchronize()
print("done", mode)
