"""Times the distance-field build alone (CUDA events): python tools/sdf_bench.py [N] [target_verts] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from meshode_b200 import capi  # noqa: E402
from meshode_b200 import pyDeform as pd  # noqa: E402
from meshode_b200.synth import synth_mesh  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
nv = int(sys.argv[2]) if len(sys.argv) > 2 else 25002
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
V, F = synth_mesh(nv, 1)
tV, tF = torch.from_numpy(V).cuda(), torch.from_numpy(F).cuda()
ts = []
for k in range(reps + 2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pid = pd.InitializeDeformTemplate(tV, tF, 0, N)
    e1.record()
    torch.cuda.synchronize()
    if k >= 2:
        ts.append(e0.elapsed_time(e1))
    pd.DestroyTemplate(pid)
# the counts come from one instrumented build (the timed ones run without the counters)
capi.lib().mo_build_stats_enable(1)
pid = pd.InitializeDeformTemplate(tV, tF, 0, N)
st = capi.template_build_stats(pid)
pd.DestroyTemplate(pid)
capi.lib().mo_build_stats_enable(0)
nvox = N ** 3
print("N=%d tris=%d: %.3f ms (min %.3f) | per voxel: dense %.1f  disc %.1f  cluster %.1f  fp64 %.2f" %
      (N, F.shape[0], np.mean(ts), np.min(ts), st["fp32_tests"] / nvox, st["disc_tests"] / nvox, st["cull_tests"] / nvox,
       st["fp64_tests"] / nvox))
