"""Times the persistent deformation kernel alone (CUDA events), for kernel tuning:
  [MESHODE_DEFORM_THREADS=768] python tools/deform_bench.py [pairs] [iters] [verts]
Prints microseconds per pair-iteration per SM.  Not a bench.py number."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from meshode_b200 import engine  # noqa: E402
from meshode_b200.synth import synth_pair  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 148
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 500
verts = int(sys.argv[3]) if len(sys.argv) > 3 else 5000
exact = os.environ.get("MESHODE_EXACT", "0") == "1"
schedule = os.environ.get("MESHODE_SCHEDULE", "auto")   # auto | cta | cluster
dev = "cuda:0"
pairs = [tuple(torch.from_numpy(a).to(dev) for a in synth_pair(i, verts, verts)) for i in range(n)]
res = []
for rep in range(3):
    b = engine.PairBatch(pairs, 64)
    b.deform(iters=1, exact=exact, schedule=schedule)   # builds the adjacency and the cell records (not part of the kernel timing)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    b.deform(iters=iters, exact=exact, schedule=schedule)
    e1.record()
    torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1))
    b.release()
ms = min(res)
waves = -(-n // 148)
print("exact=%s schedule=%s pairs=%d iters=%d verts=%d: %.2f ms -> %.2f us per iteration of a full wave (%.2f us per pair-iteration)" %
      (exact, schedule, n, iters, verts, ms, ms * 1e3 / (waves * iters), ms * 1e3 / (n * iters)))
