"""Times the any-size Adam loop (mo_deform_adam_large: one fused loss launch + one Adam launch per iteration) on the
cfg1 pair (21 542-vertex source, 29 532-triangle target, grid 64): python tools/large_bench.py [iters]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from meshode_b200 import engine  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "meshes.npz"))
pair = tuple(torch.from_numpy(np.ascontiguousarray(d[k])).cuda() for k in ("srcV", "srcF", "tarV", "tarF"))
for rep in range(3):
    b = engine.PairBatch([pair], 64)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    b.deform(iters=iters)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    print("cfg1 source %d vertices, %d iterations: %.1f ms on the GPU (%.2f us per iteration), %.1f ms wall" %
          (pair[0].shape[0], iters, e0.elapsed_time(e1), 1e3 * e0.elapsed_time(e1) / iters, 1e3 * wall))
    b.release()
