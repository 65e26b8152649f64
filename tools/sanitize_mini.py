"""One schedule of the exact deformation loop on two small pairs, for compute-sanitizer:
  compute-sanitizer --tool racecheck python tools/sanitize_mini.py cta|cluster"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from meshode_b200 import engine  # noqa: E402
from meshode_b200.synth import synth_pair  # noqa: E402

schedule = sys.argv[1] if len(sys.argv) > 1 else "cta"
pairs = [tuple(torch.from_numpy(a) for a in synth_pair(i, 1100 + 300 * i, 600)) for i in range(2)]
b = engine.PairBatch(pairs, 32, device="cuda:0")
b.deform(iters=12, schedule=schedule)
b.finalize()
b.release()
torch.cuda.synchronize()
print("sanitize mini done:", schedule)
