"""Deviation of the fast deformation loop from the exact (bit-identical-to-CPU) loop over full-length
runs: python tools/fast_vs_exact.py [pairs] [iters].  Prints per-pair max |dV|, quantiles and Chamfer."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from scipy.spatial import cKDTree  # noqa: E402

from meshode_b200 import engine  # noqa: E402
from meshode_b200.synth import synth_pair  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
pairs = [tuple(torch.from_numpy(a).cuda() for a in synth_pair(i)) for i in range(n)]
a = engine.PairBatch(pairs, 64); a.deform(iters=iters, exact=True)
b = engine.PairBatch(pairs, 64); b.deform(iters=iters, exact=False)
torch.cuda.synchronize()
for i in range(n):
    A = a.V[i].cpu().numpy(); B = b.V[i].cpu().numpy()
    d = np.abs(A - B).max(1)
    ch = cKDTree(B).query(A)[0].mean() + cKDTree(A).query(B)[0].mean()
    moved = np.abs(A - a_src[i]).max() if False else 0
    print("pair %d: max|dV| %.3g  p99.9 %.3g  p99 %.3g  median %.3g  frac<=1e-6 %.4f  chamfer %.3g" %
          (i, d.max(), np.quantile(d, 0.999), np.quantile(d, 0.99), np.median(d), (d <= 1e-6).mean(), ch))
