#!/usr/bin/env python
"""batch_deform.py -- a list of shape pairs deformed on all GPUs of one box (the cfg4 workload of BASELINE.json
on real files).  Every line of --filelist is `source.obj target.obj output.obj`; ranks take contiguous
blocks of lines (no collective on the data path), each rank runs its pairs through the fused Adam loop in
chunks of --chunk pairs.

  python scripts/batch_deform.py --filelist pairs.txt                                  (one GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
      scripts/batch_deform.py --filelist pairs.txt                                      (eight)
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import pyDeform  # noqa: E402
from meshode_b200 import engine  # noqa: E402
from meshode_b200.sharding import shard_range  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--filelist', required=True)
    ap.add_argument('--niter', type=int, default=10000)
    ap.add_argument('--grid', type=int, default=64)
    ap.add_argument('--chunk', type=int, default=592, help='pairs resident on the GPU at a time')
    a = ap.parse_args()
    rank = int(os.environ.get('RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    lines = [ln.split() for ln in open(a.filelist) if ln.strip() and not ln.startswith('#')]
    lo, hi = shard_range(len(lines), rank, world)
    t0 = time.perf_counter()
    done = 0
    for c0 in range(lo, hi, a.chunk):
        block = lines[c0:min(hi, c0 + a.chunk)]
        pairs, faces = [], []
        for src, tar, _ in block:
            sV, sF = pyDeform.LoadMesh(src)
            tV, tF = pyDeform.LoadMesh(tar)
            pairs.append((sV, sF, tV, tF)); faces.append(sF)
        batch = engine.PairBatch(pairs, grid_resolution=a.grid, device=dev)
        batch.deform(iters=a.niter, lr=1e-3)
        for (_, _, out), V, F in zip(block, batch.finalize(), faces):
            pyDeform.SaveMesh(out, V, F)
        batch.release()
        done += len(block)
    torch.cuda.synchronize()
    print('rank %d/%d: %d pairs in %.2f s' % (rank, world, done, time.perf_counter() - t0))


if __name__ == '__main__':
    main()
