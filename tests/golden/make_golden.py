"""Generates the committed fixtures under tests/golden/ (run once, in the build container).

Inputs are the reference's shipped meshes (/root/reference/data/*.obj, inputs only -- the
reference ships no expected outputs).  Outputs are produced by the CPU oracle
(oracle/meshode_oracle.cc); the FP64 brute-force search is the ground truth for the grid.
The GPU box has no /root/reference, so the meshes themselves are stored as arrays too.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from meshode_b200.objio import read_obj  # noqa: E402
from oracle import oracle as O  # noqa: E402

REF = "/root/reference/data"
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    srcV, srcF = read_obj(os.path.join(REF, "source.obj"))
    tarV, tarF = read_obj(os.path.join(REF, "target.obj"))
    csV, csF = read_obj(os.path.join(REF, "cad-source.obj"))
    ctV, ctF = read_obj(os.path.join(REF, "cad-target.obj"))
    np.savez_compressed(os.path.join(OUT, "meshes.npz"), srcV=srcV, srcF=srcF, tarV=tarV, tarF=tarF, cadSrcV=csV,
                        cadSrcF=csF, cadTarV=ctV, cadTarF=ctF)

    Vn, scale, trans = O.normalize_target(tarV)
    N = 32
    grid, idx = O.build_grid(Vn, tarF, N, fast=False)          # FP64 brute force
    grid_fast, idx_fast = O.build_grid(Vn, tarF, N, fast=True)
    assert np.array_equal(grid, grid_fast) and np.array_equal(idx, idx_fast)
    src_n = O.normalize_by_template(srcV, scale, trans)
    sel = np.arange(0, src_n.shape[0], 7)
    P = src_n[sel]
    fwd = O.distfield_forward(grid, P)
    bwd = O.distfield_backward(grid, P)
    val64, grad64 = O.distance_double_jet(grid, P.astype(np.float64))
    rest = O.store_rigid(src_n, srcF)
    # perturb deterministically so the edge residuals are non-zero
    rng = np.random.default_rng(7)
    moved = (src_n + rng.normal(0, 2e-3, src_n.shape)).astype(np.float32)
    rf = O.rigid_forward(moved, srcF, rest)
    rb = O.rigid_backward(moved, srcF, rest)
    E = np.ascontiguousarray(np.stack([srcF[:, 0], srcF[:, 1]], axis=1)[:5000], dtype=np.int32)
    grest = O.store_graph(src_n, E)
    gf = O.graph_forward(moved, E, grest)
    gb = O.graph_backward(moved, E, grest)
    crest, clam = O.store_cad(src_n, srcF[:3000], E)
    cf = O.cad_forward(moved, srcF[:3000], E, crest, clam)
    cb = O.cad_backward(moved, srcF[:3000], E, crest, clam)
    adamV, log = O.rigid_adam(grid, src_n[:0].copy() if False else src_n, srcF, rest, 20, 1e-3, log_every=5)
    np.savez_compressed(
        os.path.join(OUT, "golden_cfg1.npz"), N=N, scale=scale, trans=trans, Vn_head=Vn[:64], grid=grid, nearest=idx,
        sel=sel, dist_fwd=fwd, dist_bwd=bwd, val64=val64, grad64=grad64, moved_seed=7, rigid_fwd_rows=rf[::97],
        rigid_fwd_sum=np.float64(rf.astype(np.float64).sum()), rigid_bwd=rb[::11], graph_fwd=gf[::13], graph_bwd=gb[::11],
        cad_lambda=clam[::17], cad_fwd=cf[::13], cad_bwd=cb[::11], adam20_rows=adamV[::101], adam_log=log)
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
