#!/usr/bin/env python
"""rigid_deform.py -- the reference's src/python/rigid_deform.py on the B200 path: same arguments
(--source --target --output), same loss (RigidLossLayer), same optimiser (Adam, lr 1e-3, 10 000
iterations), same progress lines.

  --engine fused   (default) the whole loop in one persistent kernel (meshode_b200.engine, bit-identical to
                   the float32 CPU loop); progress lines are not available inside the kernel
  --engine layers  the reference's per-iteration loop through the loss layer and torch.optim.Adam
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402  (torch first, as with the reference's pyDeform)
from torch import nn  # noqa: E402
import torch.optim as optim  # noqa: E402

import pyDeform  # noqa: E402
from meshode_b200 import engine  # noqa: E402
from meshode_b200.layers.rigid_loss_layer import Finalize, RigidLossLayer  # noqa: E402

parser = argparse.ArgumentParser(description='Rigid Deformation.')
parser.add_argument('--source', default='../data/source.obj')
parser.add_argument('--target', default='../data/target.obj')
parser.add_argument('--output', default='./output.obj')
parser.add_argument('--engine', default='fused', choices=['fused', 'layers'])
parser.add_argument('--niter', type=int, default=10000)
parser.add_argument('--grid', type=int, default=64)
args = parser.parse_args()

src_V, src_F = pyDeform.LoadMesh(args.source)
tar_V, tar_F = pyDeform.LoadMesh(args.target)
dev = torch.device('cuda', torch.cuda.current_device())

if args.engine == 'fused':
    batch = engine.PairBatch([(src_V, src_F, tar_V, tar_F)], grid_resolution=args.grid, device=dev)
    batch.deform(iters=args.niter, lr=1e-3)
    out_V = batch.finalize()[0]
    pyDeform.SaveMesh(args.output, out_V, src_F)
    batch.release()
else:
    src_V, src_F, tar_V, tar_F = (t.to(dev) for t in (src_V, src_F, tar_V, tar_F))
    rigid_deform = RigidLossLayer(src_V, src_F, tar_V, tar_F, grid_resolution=args.grid)
    param_id = rigid_deform.param_id
    src_V = nn.Parameter(src_V)
    optimizer = optim.Adam([src_V], lr=1e-3)
    for it in range(0, args.niter):
        optimizer.zero_grad()
        loss = rigid_deform(src_V, src_F)
        loss.backward()
        optimizer.step()
        if it % 100 == 0:
            print('iter=%d loss=%.6f' % (it, loss.item()))
    Finalize(src_V, param_id)
    pyDeform.SaveMesh(args.output, src_V, src_F)
