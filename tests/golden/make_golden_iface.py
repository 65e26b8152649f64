"""Generates tests/golden/golden_iface.npz from the REFERENCE's own interface loops (oracle/_ref/libmeshode_ref.so =
/root/reference/src/interface/{distance,rigid,graph,cad}_layer.cc + normalize.cc + src/lib/uniformgrid.cc compiled
where they lie, `make -C oracle ref`): SURVEY.md s8 rows a11-a16.  Run in the build container, where /root/reference
exists; the GPU box and the CPU test run only read the committed .npz."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import ref as R  # noqa: E402
from ifacecases import SCALE, TRANS, iface_case  # noqa: E402


def main():
    R.build(force=True)
    grid = np.load(os.path.join(HERE, "golden_cfg1.npz"))["grid"]
    V, F, E, moved, raw = iface_case()
    P = R.Params(grid, SCALE, TRANS)
    out = {"normalize": P.normalize(raw), "denormalize": P.normalize(V, inverse=True),
           "dist_fwd": P.dist_forward(moved), "dist_bwd": P.dist_backward(moved)}
    P.rigid_store(V, F)
    out["rigid_fwd"] = P.rigid_forward(moved, F); out["rigid_bwd"] = P.rigid_backward(moved, F)
    P.graph_store(V, E)
    out["graph_fwd"] = P.graph_forward(moved, E); out["graph_bwd"] = P.graph_backward(moved, E)
    P.cad_store(V, F, E)
    out["cad_lambda"] = P.cad_lambda(E.shape[0] + 3 * F.shape[0])
    out["cad_fwd"] = P.cad_forward(moved, F, E); out["cad_bwd"] = P.cad_backward(moved, F, E)
    np.savez_compressed(os.path.join(HERE, "golden_iface.npz"), **out)
    print("wrote golden_iface.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
