"""How sensitive is the 10 000-iteration Adam loop to rounding?  Runs the EXACT loop twice, the second
time with every normalised source coordinate moved by one float32 ulp, and prints the deviation of the
results.  (python tools/chaos_probe.py [pairs] [iters])"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from scipy.spatial import cKDTree  # noqa: E402

from meshode_b200 import engine  # noqa: E402
from meshode_b200.synth import synth_pair  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
pairs = [tuple(torch.from_numpy(a).cuda() for a in synth_pair(i)) for i in range(n)]
a = engine.PairBatch(pairs, 64)
b = engine.PairBatch(pairs, 64)
for v in b.V:
    v.copy_(torch.nextafter(v, torch.full_like(v, 2.0)))   # +1 ulp on the normalised coordinates
a.deform(iters=iters, exact=True)
b.deform(iters=iters, exact=True)
torch.cuda.synchronize()
for i in range(n):
    A = a.V[i].cpu().numpy(); B = b.V[i].cpu().numpy()
    d = np.abs(A - B).max(1)
    ch = cKDTree(B).query(A)[0].mean() + cKDTree(A).query(B)[0].mean()
    print("pair %d (+1 ulp start, %d its): max|dV| %.3g  p99 %.3g  median %.3g  chamfer %.3g" %
          (i, iters, d.max(), np.quantile(d, 0.99), np.median(d), ch))
