"""Host enqueue time vs GPU completion time of the per-pair set-up (template build, normalise, store):
  python tools/setup_bench.py [pairs] [streams]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from meshode_b200 import engine  # noqa: E402
from meshode_b200.synth import synth_pair  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 296
streams = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = "cuda:0"
pairs = [tuple(torch.from_numpy(a).to(dev) for a in synth_pair(i, 5000, 5000)) for i in range(n)]
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    b = engine.PairBatch(pairs, 64, n_streams=streams)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    b.deform(iters=1)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    b.finalize()
    b.release()
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    print("pairs=%d streams=%d: enqueue %.1f us/pair, until idle %.1f us/pair, adjacency+1 iteration %.1f us/pair, "
          "finalize+release %.1f us/pair" % (n, streams, (t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6, (t3 - t2) / n * 1e6,
                                             (t4 - t3) / n * 1e6))
