#!/usr/bin/env python
"""cad_deform2.py -- the reference's src/python/cad_deform2.py on the B200 path (cfg2 of BASELINE.json): same
arguments (--source --target --output --rigidity), same pipeline: LoadCadMesh (subdivision + deformation graph,
host), GraphLossLayer + ReverseLossLayer on the graph nodes (GPU, no host round trip), Adam lr 1e-3 with the
reference's stopping rule every 100 iterations, then the sparse post-solve (SolveLinear, host) and SaveMesh."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402
from torch import nn  # noqa: E402
import torch.optim as optim  # noqa: E402

import pyDeform  # noqa: E402
from meshode_b200.layers.graph_loss_layer import Finalize, GraphLossLayer  # noqa: E402
from meshode_b200.layers.reverse_loss_layer import ReverseLossLayer  # noqa: E402

parser = argparse.ArgumentParser(description='Rigid Deformation.')
parser.add_argument('--source', default='../data/cad-source.obj')
parser.add_argument('--target', default='../data/cad-target.obj')
parser.add_argument('--output', default='./cad-output.obj')
parser.add_argument('--rigidity', default='1')
parser.add_argument('--niter', type=int, default=10000)
args = parser.parse_args()

rigidity = float(args.rigidity)
dev = torch.device('cuda', torch.cuda.current_device())
src_V, src_F, src_E, src_to_graph, graph_V, graph_E = pyDeform.LoadCadMesh(args.source)
tar_V, tar_F, tar_E, tar_to_graph, graph_V_tar, graph_E_tar = pyDeform.LoadCadMesh(args.target)

graph_V, graph_E, graph_V_tar = graph_V.to(dev), graph_E.to(dev), graph_V_tar.to(dev)
graph_deform = GraphLossLayer(graph_V, graph_E, tar_V, tar_F, rigidity, dev)   # normalises graph_V in place
param_id = graph_deform.param_id
reverse_deform = ReverseLossLayer()

graph_V = nn.Parameter(graph_V)
optimizer = optim.Adam([graph_V], lr=1e-3)

pyDeform.NormalizeByTemplate(graph_V_tar, param_id.tolist())
prev_loss_src, prev_loss_tar = 1e30, 1e30
for it in range(0, args.niter):
    optimizer.zero_grad()
    loss_src2tar = graph_deform(graph_V, graph_E)
    loss_tar2src = reverse_deform(graph_V, graph_V_tar, dev)
    loss = loss_src2tar / graph_V.shape[0] + loss_tar2src / graph_V_tar.shape[0]
    loss.backward()
    optimizer.step()
    if it % 100 == 0:
        current_loss_src = np.sqrt(loss_src2tar.item() / graph_V.shape[0])
        current_loss_tar = np.sqrt(loss_tar2src.item() / graph_V_tar.shape[0])
        print('iter=%d, loss_src2tar=%.6f loss_tar2src=%.6f' % (it, current_loss_src, current_loss_tar))
        if prev_loss_src - current_loss_src < 1e-6 and prev_loss_tar - current_loss_tar < 1e-6:
            break
        prev_loss_src, prev_loss_tar = current_loss_src, current_loss_tar

Finalize(src_V, src_F, src_E, src_to_graph, graph_V.detach().cpu(), 1, param_id)
pyDeform.SaveMesh(args.output, src_V, src_F)
