#!/usr/bin/env python
"""cad_neural_deform2.py -- the reference's src/python/cad_neural_deform2.py on the B200 path (the loss of cfg5 of
BASELINE.json): same arguments; a NeuralODE flow (4-50-50-50-3 MLP, RK4) deforms the two deformation graphs
towards each other under GraphLoss2Layer + ReverseLossLayer, everything on the GPU; the source mesh is then
pushed through the flow and cleaned up by the rotation-aware sparse solve (SolveLinear(..., 1, 1))."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.optim as optim  # noqa: E402

import pyDeform  # noqa: E402
from meshode_b200.layers.graph_loss2_layer import GraphLoss2Layer  # noqa: E402
from meshode_b200.layers.neuralode import NeuralODE, load_checkpoint, save_checkpoint  # noqa: E402
from meshode_b200.layers.reverse_loss_layer import ReverseLossLayer  # noqa: E402

parser = argparse.ArgumentParser(description='Rigid Deformation.')
parser.add_argument('--source', default='../data/cad-source.obj')
parser.add_argument('--target', default='../data/cad-target.obj')
parser.add_argument('--output', default='./cad-output.obj')
parser.add_argument('--rigidity', default='0.1')
parser.add_argument('--device', default='cuda')
parser.add_argument('--save_path', default='./cad-output.ckpt')
parser.add_argument('--niter', type=int, default=1000)
parser.add_argument('--resume_path', default='')
args = parser.parse_args()

rigidity = float(args.rigidity)
if os.environ.get('MESHODE_SEED'):   # (additive: a fixed initialisation of the flow for tests; the reference draws it unseeded)
    torch.manual_seed(int(os.environ['MESHODE_SEED']))
device = torch.device(args.device)
if device.type != 'cuda':
    raise SystemExit('meshode_b200 runs on a CUDA device (no CPU fallback)')
if device.index is None:
    device = torch.device('cuda', torch.cuda.current_device())

V1, F1, E1, V2G1, GV1, GE1 = pyDeform.LoadCadMesh(args.source)
V2, F2, E2, V2G2, GV2, GE2 = pyDeform.LoadCadMesh(args.target)

graph_loss = GraphLoss2Layer(V1, F1, GV1, GE1, V2, F2, GV2, GE2, rigidity, device)   # normalises GV1 / GV2 in place
param_id1, param_id2 = graph_loss.param_id1, graph_loss.param_id2
reverse_loss = ReverseLossLayer()
if args.resume_path != '' and os.path.exists(args.resume_path):
    func, optimizer = load_checkpoint(args.resume_path, device)   # the reference's layout or round 1's state_dicts
else:
    func = NeuralODE(device)
    optimizer = optim.Adam(func.parameters(), lr=1e-3)
GV1_device, GV2_device = GV1.to(device), GV2.to(device)
GV1_origin, GV2_origin = GV1_device.clone(), GV2_device.clone()

for it in range(0, args.niter):
    optimizer.zero_grad()
    GV1_deformed = func.forward(GV1_device)
    GV2_deformed = func.inverse(GV2_device)
    loss1_forward = graph_loss(GV1_deformed, GE1, GV2_device, GE2, 0)
    loss1_backward = reverse_loss(GV1_deformed, GV2_origin, device)
    loss2_forward = graph_loss(GV1_device, GE1, GV2_deformed, GE2, 1)
    loss2_backward = reverse_loss(GV2_deformed, GV1_origin, device)
    loss = loss1_forward + loss1_backward + loss2_forward + loss2_backward
    loss.backward()
    optimizer.step()
    print('iter=%d, loss1_forward=%.6f loss1_backward=%.6f loss2_forward=%.6f loss2_backward=%.6f'
          % (it, np.sqrt(loss1_forward.item() / GV1.shape[0]), np.sqrt(loss1_backward.item() / GV2.shape[0]),
             np.sqrt(loss2_forward.item() / GV2.shape[0]), np.sqrt(loss2_backward.item() / GV1.shape[0])))

if args.save_path != '':
    save_checkpoint(args.save_path, func, optimizer)   # {'func': func, 'optim': optimizer}, as cad_neural_deform2.py:108

V1_copy = V1.clone()
pyDeform.NormalizeByTemplate(V1_copy, param_id1.tolist())
V1_origin = V1_copy.clone()
with torch.no_grad():
    V1_copy = func.forward(V1_copy.to(device)).cpu()
src_to_src = torch.from_numpy(np.arange(V1_origin.shape[0], dtype=np.int32))
pyDeform.SolveLinear(V1_origin, F1, E1, src_to_src, V1_copy, 1, 1)
pyDeform.DenormalizeByTemplate(V1_origin, param_id2.tolist())
pyDeform.SaveMesh(args.output, V1_origin, F1)
