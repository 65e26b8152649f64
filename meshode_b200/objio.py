"""OBJ reader/writer with the reference's semantics (host-side, not on the hot path).

ReadOBJ follows src/lib/mesh.cc:14-45: only ``v`` and ``f`` records, the leading
integer of each ``a/b/c`` face token, 1-based indices, first three tokens of a face.
WriteOBJ follows src/lib/mesh.cc:47-64.
"""
import numpy as np


def read_obj(path, vertex_dtype=np.float32):
    """``vertex_dtype=np.float64`` keeps the doubles Mesh::ReadOBJ parses (the CAD preprocessing works on them)."""
    V, F = [], []
    with open(path, "r", errors="replace") as fh:
        for line in fh:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "v" and len(tok) >= 4:
                V.append((float(tok[1]), float(tok[2]), float(tok[3])))
            elif tok[0] == "f" and len(tok) >= 4:
                f = []
                for t in tok[1:4]:
                    lead = t.split("/")[0]
                    f.append(int(lead) - 1 if lead else -1)
                F.append(f)
    V = np.asarray(V, dtype=np.float64).reshape(-1, 3)
    F = np.asarray(F, dtype=np.int32).reshape(-1, 3)
    # LoadMesh exports float32 vertices / int32 faces (src/interface/mesh_tensor.cc:87-100)
    return np.ascontiguousarray(V, dtype=vertex_dtype), np.ascontiguousarray(F, dtype=np.int32)


def write_obj(path, V, F):
    V = np.asarray(V, dtype=np.float64)
    F = np.asarray(F)
    with open(path, "w") as fh:
        for v in V:
            fh.write("v %.6g %.6g %.6g\n" % (v[0], v[1], v[2]))
        for f in F:
            fh.write("f %d %d %d\n" % (f[0] + 1, f[1] + 1, f[2] + 1))
