#!/usr/bin/env python
"""cad_deform2.py -- the reference's src/python/cad_deform2.py on the B200 path (cfg2 of BASELINE.json): same
arguments (--source --target --output --rigidity), same pipeline: LoadCadMesh (subdivision + deformation graph,
host), GraphLossLayer + ReverseLossLayer on the graph nodes (GPU, no host round trip), Adam lr 1e-3 with the
reference's stopping rule every 100 iterations, then the sparse post-solve (SolveLinear, host) and SaveMesh."""
import argparse
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402  (torch before pyDeform, as with the reference's extension)

import pyDeform  # noqa: E402
from meshode_b200.layers.graph_loss_layer import Finalize, GraphLossLayer  # noqa: E402
from meshode_b200.layers.reverse_loss_layer import ReverseLossLayer  # noqa: E402


def parse_args():
    ap = argparse.ArgumentParser(description='Rigid Deformation.')
    ap.add_argument('--source', default='../data/cad-source.obj')
    ap.add_argument('--target', default='../data/cad-target.obj')
    ap.add_argument('--output', default='./cad-output.obj')
    ap.add_argument('--rigidity', default='1')
    ap.add_argument('--niter', type=int, default=10000, help='iteration budget (reference: 10000)')
    ap.add_argument('--check-every', type=int, default=100, help='progress line and stopping test period (reference: 100)')
    return ap.parse_args()


def optimise_graph(nodes, edges, target_nodes, src2tar, tar2src, niter, period, dev):
    """Adam on the graph nodes; stops when neither RMS loss improved by 1e-6 over one period (cad_deform2.py:62-69)."""
    nodes = torch.nn.Parameter(nodes)
    opt = torch.optim.Adam([nodes], lr=1e-3)
    n_src, n_tar = nodes.shape[0], target_nodes.shape[0]
    best = (1e30, 1e30)
    for it in range(niter):
        opt.zero_grad()
        l_fwd = src2tar(nodes, edges)
        l_bwd = tar2src(nodes, target_nodes, dev)
        (l_fwd / n_src + l_bwd / n_tar).backward()
        opt.step()
        if it % period:
            continue
        rms = (math.sqrt(l_fwd.item() / n_src), math.sqrt(l_bwd.item() / n_tar))
        print('iter=%d, loss_src2tar=%.6f loss_tar2src=%.6f' % (it, rms[0], rms[1]))
        if best[0] - rms[0] < 1e-6 and best[1] - rms[1] < 1e-6:
            break
        best = rms
    return nodes.detach()


def main():
    a = parse_args()
    dev = torch.device('cuda', torch.cuda.current_device())
    src_V, src_F, src_E, src_to_graph, graph_V, graph_E = pyDeform.LoadCadMesh(a.source)
    tar_V, tar_F, _, _, graph_V_tar, _ = pyDeform.LoadCadMesh(a.target)
    graph_V, graph_E, graph_V_tar = graph_V.to(dev), graph_E.to(dev), graph_V_tar.to(dev)
    layer = GraphLossLayer(graph_V, graph_E, tar_V, tar_F, float(a.rigidity), dev)   # normalises graph_V in place
    pyDeform.NormalizeByTemplate(graph_V_tar, layer.param_id.tolist())
    moved = optimise_graph(graph_V, graph_E, graph_V_tar, layer, ReverseLossLayer(), a.niter, a.check_every, dev)
    Finalize(src_V, src_F, src_E, src_to_graph, moved.cpu(), 1, layer.param_id)     # SolveLinear between the normalisations
    pyDeform.SaveMesh(a.output, src_V, src_F)


if __name__ == '__main__':
    main()
