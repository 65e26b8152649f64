"""Achieved bandwidth of the per-call loss kernels on a mesh large enough to be memory bound:
K copies of a 50 000-vertex synthetic source (disjoint union) against one 128^3 field.
  python tools/loss_bench.py [copies=40] [reps=20]
Algorithmic bytes: distance fwd 16 B/vertex, bwd 24, fused 28; edge fwd 20 + 36 B/edge
(the API's [E,3] output), bwd 20 B/edge + 12 B/vertex gradient + 12 B/vertex position; fused loss
28 B/vertex + 20 B/edge."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from meshode_b200 import pyDeform as pd  # noqa: E402
from meshode_b200.synth import synth_mesh, synth_params  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 40
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = "cuda:0"
tV, tF = synth_mesh(25002, 1)
sV, sF = synth_mesh(50000, 0, axis_scale=synth_params(1)[3])
n1 = sV.shape[0]
rng = np.random.default_rng(0)
V = np.concatenate([sV + rng.normal(0, 1e-3, sV.shape).astype(np.float32) for _ in range(K)])
F = np.concatenate([sF + k * n1 for k in range(K)]).astype(np.int32)
nV, nE = V.shape[0], 3 * F.shape[0]
dT, dTF = torch.from_numpy(tV).to(dev), torch.from_numpy(tF).to(dev)
dV, dF = torch.from_numpy(V).to(dev), torch.from_numpy(F).to(dev)
pid = pd.InitializeDeformTemplate(dT, dTF, 0, 128)
pd.NormalizeByTemplate(dV, pid)
pd.StoreRigidityInformation(dV, dF, pid)
moved = (dV + 1e-3 * torch.sin(37.0 * dV)).contiguous()
peaks = {}
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peaks = json.load(open(p))
hbm = float(peaks.get("hbm_gbs", 6650.0))


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


cases = [
    ("DistanceFieldLoss_forward", lambda: pd.DistanceFieldLoss_forward(moved, pid), 16 * nV),
    ("DistanceFieldLoss_backward", lambda: pd.DistanceFieldLoss_backward(moved, pid), 24 * nV),
    ("DistanceFieldLoss fused fwd+bwd", lambda: pd.DistanceFieldLoss_forward_backward(moved, pid), 28 * nV),
    ("RigidEdgeLoss_forward", lambda: pd.RigidEdgeLoss_forward(moved, dF, pid), 56 * nE),
    ("RigidEdgeLoss_backward", lambda: pd.RigidEdgeLoss_backward(moved, dF, pid), 20 * nE + 24 * nV),
    ("LossForwardBackward (one launch)", lambda: pd.LossForwardBackward(moved, pid, pid), 28 * nV + 20 * nE),
]
print("vertices %d, directed edges %d, grid 128^3 (8 MB, L2 resident); HBM peak %.0f GB/s; inputs %.0f MB" %
      (nV, nE, hbm, (12 * nV + 20 * nE) / 1e6))
out = {}
for name, fn, nbytes in cases:
    ms = timed(fn)
    gbs = nbytes / (ms * 1e-3) / 1e9
    out[name] = {"ms": ms, "algorithmic_GBps": gbs, "frac_of_hbm": gbs / hbm}
    print("%-36s %8.3f ms  %8.1f GB/s algorithmic  = %5.1f %% of HBM peak" % (name, ms, gbs, 100 * gbs / hbm))
print(json.dumps(out))
