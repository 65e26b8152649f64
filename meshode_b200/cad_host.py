"""Host-side stages of the C++ ``cad_deform`` driver, run out of process (apps/cad_deform.cc): the CGAL-free
restatements of Subdivision::Subdivide / ComputeGeometryNeighbors / ComputeRepresentativeGraph
(reference src/lib/subdivision.cc:28-399, here meshode_b200/cadmesh.py) and of Subdivision::LinearSolve ->
LinearEstimation (src/lib/subdivision.cc:462-470, src/lib/linear.cc:10-117, here meshode_b200/linear.py).
Neither is on the GPU hot path; the driver's distance field and Ceres problem are.

  python cad_host.py prepare cad.obj out.bin          # clean-up + subdivision + neighbour pairs + deformation graph
  python cad_host.py linear  in.bin  out.bin [rigidity=2.0]

File layout (little endian): int32 magic 0x4d4f4344, nV, nF, nE, nG, nGE, then V f64[nV,3], F i32[nF,3],
E i32[nE,2] (geometric neighbour pairs), REF i32[nV] (graph node of every vertex), GV f64[nG,3], GE i32[nGE,2].
``linear`` writes V f64[nV,3] only.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

MAGIC = 0x4D4F4344


def write_bundle(path, V, F, E, ref, GV, GE):
    with open(path, "wb") as fh:
        np.array([MAGIC, V.shape[0], F.shape[0], E.shape[0], GV.shape[0], GE.shape[0]], dtype="<i4").tofile(fh)
        np.ascontiguousarray(V, dtype="<f8").tofile(fh)
        np.ascontiguousarray(F, dtype="<i4").tofile(fh)
        np.ascontiguousarray(E, dtype="<i4").tofile(fh)
        np.ascontiguousarray(ref, dtype="<i4").reshape(-1).tofile(fh)
        np.ascontiguousarray(GV, dtype="<f8").tofile(fh)
        np.ascontiguousarray(GE, dtype="<i4").tofile(fh)


def read_bundle(path):
    with open(path, "rb") as fh:
        h = np.fromfile(fh, dtype="<i4", count=6)
        if h[0] != MAGIC:
            raise ValueError("%s is not a cad_host bundle" % path)
        nV, nF, nE, nG, nGE = [int(x) for x in h[1:]]
        V = np.fromfile(fh, dtype="<f8", count=3 * nV).reshape(nV, 3)
        F = np.fromfile(fh, dtype="<i4", count=3 * nF).reshape(nF, 3)
        E = np.fromfile(fh, dtype="<i4", count=2 * nE).reshape(nE, 2)
        ref = np.fromfile(fh, dtype="<i4", count=nV)
        GV = np.fromfile(fh, dtype="<f8", count=3 * nG).reshape(nG, 3)
        GE = np.fromfile(fh, dtype="<i4", count=2 * nGE).reshape(nGE, 2)
    return V, F, E, ref, GV, GE


def prepare(obj_path, out_path):
    from meshode_b200 import cadmesh
    from meshode_b200.objio import read_obj
    V, F = read_obj(obj_path, vertex_dtype=np.float64)
    V = np.asarray(V, dtype=np.float64)
    F = np.asarray(F, dtype=np.int64).reshape(-1, 3)
    F = cadmesh.remove_degenerated(V, F)                # cad_deform.cc:43
    V, F = cadmesh.merge_duplex(V, F)                   # :44
    V, F = cadmesh.subdivide(V, F, 2e-2)                # :48
    E = cadmesh.geometry_neighbors(V, F, 1.5e-2)        # :49
    ref, GV, GE = cadmesh.representative_graph(V, F, E, 1e-2)   # :50
    write_bundle(out_path, V, F, E, ref, GV, GE)


def linear(in_path, out_path, rigidity=2.0):
    from meshode_b200 import linear as L
    V, F, E, ref, GV, _ = read_bundle(in_path)
    out = L.linear_estimation(V, F, E, ref, GV, rigidity)   # linear.h: rigidity = 2.0 by default
    with open(out_path, "wb") as fh:
        np.ascontiguousarray(out, dtype="<f8").tofile(fh)


def main(argv):
    if len(argv) >= 4 and argv[1] == "prepare":
        prepare(argv[2], argv[3])
    elif len(argv) >= 4 and argv[1] == "linear":
        linear(argv[2], argv[3], float(argv[4]) if len(argv) > 4 else 2.0)
    else:
        print(__doc__)
        return 2
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
