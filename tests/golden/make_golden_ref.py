"""Generates tests/golden/golden_ref.npz from the REFERENCE's own compiled code
(oracle/_ref/libmeshode_ref.so = /root/reference/src/lib/uniformgrid.cc + distanceloss.h +
edgeloss.h, built by `make -C oracle ref`).  Run in the build container, where /root/reference
exists; the GPU box and the CPU test run only read the committed .npz."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import ref as R  # noqa: E402
from refcases import functor_cases, sampler_cases  # noqa: E402


def main():
    R.build(force=True)
    golden = np.load(os.path.join(HERE, "golden_cfg1.npz"))
    out = {}
    for name, (grid, P) in sampler_cases(golden["grid"]).items():
        G = R.Grid(grid)
        out[name + "/d64"] = G.distance_double(P)
        v, g = G.distance_double_jet(P)
        out[name + "/j64v"] = v; out[name + "/j64g"] = g
        P32 = P.astype(np.float32)
        out[name + "/d32"] = G.distance_float(P32)
        v, g = G.distance_float_jet(P32)
        out[name + "/j32v"] = v; out[name + "/j32g"] = g
        r = [G.distance_loss(p) for p in P[:50]]
        out[name + "/dl_r"] = np.stack([x[0] for x in r]); out[name + "/dl_J"] = np.stack([x[1] for x in r])
    p1, p2, rot1, rot2, v, lam = functor_cases()
    e = [R.edge_loss(p1[i], p2[i], v[i], lam[i], False) for i in range(len(lam))]
    a = [R.edge_loss(p1[i], p2[i], v[i], lam[i], True) for i in range(len(lam))]
    rr = [R.edge_rot(p1[i], p2[i], rot1[i], rot2[i], v[i], lam[i]) for i in range(len(lam))]
    out["edge/r"] = np.stack([x[0] for x in e]); out["edge/lam"] = np.array([x[1] for x in e])
    out["aedge/r"] = np.stack([x[0] for x in a]); out["aedge/lam"] = np.array([x[1] for x in a])
    out["rot/r"] = np.stack([x[0] for x in rr]); out["rot/J"] = np.stack([x[1] for x in rr])
    np.savez_compressed(os.path.join(HERE, "golden_ref.npz"), **out)
    print("wrote golden_ref.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
