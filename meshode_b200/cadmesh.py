"""LoadCadMesh on the host: the CAD preprocessing of the reference (src/interface/mesh_tensor.cc:102-178 ->
src/lib/mesh.cc RemoveDegenerated / MergeDuplex, src/lib/subdivision.cc Subdivide / DelaunaySubdivision /
ComputeGeometryNeighbors / ComputeRepresentativeGraph) without CGAL: the 2-D and 3-D Delaunay triangulations
come from scipy.spatial (Qhull).

It is host code in the reference and host code here; it runs once per shape and is not on the GPU path
(SURVEY.md s8f rank 4).  A Delaunay triangulation is unique only up to co-circular / co-spherical ties, so
the output is functionally equivalent to the reference's, not index-for-index identical (parity unpinned):
same vertex sets (edge splits and lattice points are deterministic), the same edge-length bounds, the same
graph construction rules.
"""
from collections import deque

import numpy as np


def _trunc_key(v, step):
    """make_key of subdivision.cc:258-262 / :346-349: int(v / step), C truncation toward zero, per axis."""
    return np.trunc(np.asarray(v, dtype=np.float64) / step).astype(np.int64)


def remove_degenerated(V, F):
    """Mesh::RemoveDegenerated (mesh.cc:320-334): drops faces whose normal has zero length."""
    a, b, c = V[F[:, 0]], V[F[:, 1]], V[F[:, 2]]
    n = np.cross(b - a, c - a)
    return F[np.sqrt((n * n).sum(1)) > 0]


def merge_duplex(V, F):
    """Mesh::MergeDuplex (mesh.cc:181-233): vertices equal after int(v * 1e6) are merged (first one wins),
    faces with a repeated vertex or a repeated vertex set are dropped (first one wins)."""
    key = np.trunc(V * 1e6).astype(np.int64)
    _, first, inv = np.unique(key, axis=0, return_index=True, return_inverse=True)
    order = np.argsort(first)                     # keep vertices in order of first appearance
    rank = np.empty_like(order); rank[order] = np.arange(order.size)
    shrink = rank[inv.reshape(-1)]
    Vn = V[first[order]]
    Fn = shrink[F]
    ok = (Fn[:, 0] != Fn[:, 1]) & (Fn[:, 1] != Fn[:, 2]) & (Fn[:, 2] != Fn[:, 0])
    Fn = Fn[ok]
    _, firstf = np.unique(np.sort(Fn, axis=1), axis=0, return_index=True)
    return Vn, Fn[np.sort(firstf)]


def _delaunay_2d(P):
    """Counter-clockwise triangles of the 2-D Delaunay triangulation of P [n,2] (Delaunay2D, delaunay.cc:50-78)."""
    from scipy.spatial import Delaunay, QhullError
    try:
        T = Delaunay(P).simplices
    except (QhullError, ValueError):
        try:
            T = Delaunay(P, qhull_options="QJ Pp").simplices
        except (QhullError, ValueError):
            return np.zeros((0, 3), dtype=np.int64)
    a, b, c = P[T[:, 0]], P[T[:, 1]], P[T[:, 2]]
    area2 = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0])
    T = T.copy()
    flip = area2 < 0
    T[flip] = T[flip][:, [0, 2, 1]]
    return T[area2 != 0]


def _delaunay_edges_3d(P):
    """Edges (local index pairs) of the Delaunay triangulation of P [n,3] (Delaunay3D, delaunay.cc:80-102).
    CGAL triangulates degenerate inputs in their own dimension (coplanar points -> a 2-D triangulation,
    collinear points -> a chain); Qhull does not, so the dimension is detected first."""
    from scipy.spatial import Delaunay, QhullError
    n = P.shape[0]
    if n < 2:
        return np.zeros((0, 2), dtype=np.int64)
    c = P - P.mean(0)
    U, S, Vt = np.linalg.svd(c, full_matrices=False)
    tol = max(S[0], 1e-300) * 1e-9
    dim = int((S > tol).sum())
    if dim == 0:
        return np.zeros((0, 2), dtype=np.int64)
    if dim == 1:
        o = np.argsort(c @ Vt[0])
        return np.stack([o[:-1], o[1:]], 1)
    try:
        if dim == 2:
            T = _delaunay_2d(c @ Vt[:2].T)
            E = np.concatenate([T[:, [0, 1]], T[:, [1, 2]], T[:, [2, 0]]])
        else:
            T = Delaunay(P).simplices
            E = np.concatenate([T[:, [i, j]] for i in range(4) for j in range(i + 1, 4)])
    except (QhullError, ValueError):
        i, j = np.triu_indices(n, 1)          # tiny, nearly degenerate cell: connect everything
        return np.stack([i, j], 1)
    return np.unique(np.sort(E, axis=1), axis=0)


def _face_order(F):
    """Faces sorted by (connected component over shared edges, face index) -- subdivision.cc:38-84."""
    nF = F.shape[0]
    edge_faces = {}
    for i in range(nF):
        for j in range(3):
            a, b = int(F[i, j]), int(F[i, (j + 1) % 3])
            edge_faces.setdefault((a, b) if a < b else (b, a), []).append(i)
    group = np.full(nF, -1, dtype=np.int64)
    g = 0
    for i in range(nF):
        if group[i] >= 0:
            continue
        group[i] = g
        q = deque([i])
        while q:
            f = q.popleft()
            for j in range(3):
                a, b = int(F[f, j]), int(F[f, (j + 1) % 3])
                for nf in edge_faces[(a, b) if a < b else (b, a)]:
                    if group[nf] < 0:
                        group[nf] = g
                        q.append(nf)
        g += 1
    return np.lexsort((np.arange(nF), group))


def subdivide(V, F, len_thres):
    """Subdivision::Subdivide + DelaunaySubdivision (subdivision.cc:28-250): every edge is split into
    int(len / len_thres) + 1 pieces, every face receives the lattice points v0 + tx*px*len_thres + ty*py*len_thres
    strictly inside it, and is re-triangulated (2-D Delaunay of its boundary points pushed onto three huge
    circles plus the lattice points; triangles with an edge longer than 3*len_thres are dropped)."""
    V = [np.asarray(v, dtype=np.float64) for v in V]
    faces = F[_face_order(F)]
    edge_pts = {}
    outF = []
    for face in faces:
        bidx = []
        for j in range(3):
            v0, v1 = int(face[j]), int(face[(j + 1) % 3])
            h = (v0, v1) if v0 < v1 else (v1, v0)
            if h not in edge_pts:                                      # :94-110 (direction of first use)
                diff = V[v1] - V[v0]
                ns = int(np.linalg.norm(diff) / len_thres + 1)
                diff = diff / float(ns)
                ids = [v0]
                for k in range(1, ns):
                    ids.append(len(V)); V.append(V[v0] + diff * k)
                ids.append(v1)
                edge_pts[h] = ids
            bidx.append(edge_pts[h])
        p0, p1, p2 = V[int(face[0])], V[int(face[1])], V[int(face[2])]
        n = np.cross(p1 - p0, p2 - p0); n = n / np.linalg.norm(n)
        ids, pts = [], []
        seen = set()
        for i in range(3):                                             # :146-160 boundary points on huge circles
            x, y = V[int(face[i])], V[int(face[(i + 1) % 3])]
            d = np.cross(n, y - x); d = d / np.linalg.norm(d)
            c = (y + x) * 0.5 + 1e3 * d
            ln = np.linalg.norm(c - x)
            for p in bidx[i]:
                w = V[p] - c
                cp = w / np.linalg.norm(w) * ln + c
                if p in seen:
                    pts[ids.index(p)] = cp                              # curved_point[p] is overwritten (:158)
                else:
                    seen.add(p); ids.append(p); pts.append(cp)
        min_axis = 0                                                   # :169-174 (sic: "min_axis = 1")
        for i in range(1, 3):
            if abs(n[i]) < abs(n[min_axis]):
                min_axis = 1
        tx = np.zeros(3); tx[min_axis] = 1.0
        tx = np.cross(tx, n); tx = tx / np.linalg.norm(tx)
        ty = np.cross(n, tx)
        v1x, v1y = (p1 - p0) @ tx / len_thres, (p1 - p0) @ ty / len_thres
        v2x, v2y = (p2 - p0) @ tx / len_thres, (p2 - p0) @ ty / len_thres
        minX, minY = int(min(0.0, v1x, v2x)), int(min(0.0, v1y, v2y))
        maxX, maxY = int(max(0.0, v1x, v2x) + 0.999999), int(max(0.0, v1y, v2y) + 0.999999)
        inv = 1.0 / (v1x * v2y - v2x * v1y) if (v1x * v2y - v2x * v1y) != 0 else np.inf
        for py in range(minY, maxY + 1):                               # :199-216
            for px in range(minX, maxX + 1):
                beta = (px * v2y - v2x * py) * inv                     # signed-area ratios of :442-460
                gamma = (v1x * py - px * v1y) * inv
                alpha = 1.0 - beta - gamma
                if 0.0 < alpha < 1.0 and 0.0 < beta < 1.0 and 0.0 < gamma < 1.0:
                    rp = p0 + tx * (px * len_thres) + ty * (py * len_thres)
                    ids.append(len(V)); V.append(rp); pts.append(rp)
        P = np.asarray(pts) - p0
        T = _delaunay_2d(np.stack([P @ tx, P @ ty], 1))
        ids = np.asarray(ids)
        for t in T:                                                    # :231-249
            v = ids[t]
            a, b, c = V[v[0]], V[v[1]], V[v[2]]
            if max(np.linalg.norm(a - b), np.linalg.norm(a - c), np.linalg.norm(b - c)) > 3 * len_thres:
                continue
            outF.append(v)
    return np.asarray(V, dtype=np.float64), np.asarray(outF, dtype=np.int64).reshape(-1, 3)


def geometry_neighbors(V, F, thres):
    """Subdivision::ComputeGeometryNeighbors (subdivision.cc:252-341): vertices hashed into the eight cells of
    size `thres` around them; the Delaunay edges of every cell's vertex set, plus the face edges.
    Returns the sorted pair set as [e,2] (v1 < v2)."""
    step = thres
    offs = np.array([[0, 0, 0], [0, 0, step], [0, step, 0], [0, step, step], [step, 0, 0], [step, 0, step], [step, step, 0],
                     [step, step, step]], dtype=np.float64)
    n = V.shape[0]
    keys = _trunc_key((V[:, None, :] + offs[None, :, :]).reshape(-1, 3), step)
    vid = np.repeat(np.arange(n), 8)
    # a vertex may reach the same cell through two offsets (truncation toward zero around the origin)
    kv = np.unique(np.concatenate([keys, vid[:, None]], axis=1), axis=0)
    cell_id = np.unique(kv[:, :3], axis=0, return_inverse=True)[1].reshape(-1)
    order = np.argsort(cell_id, kind="stable")
    cell_id, members = cell_id[order], kv[order, 3]
    bounds = np.flatnonzero(np.diff(cell_id)) + 1
    pairs = []
    for grp in np.split(members, bounds):
        if grp.size < 2:
            continue
        if grp.size == 2:
            pairs.append(np.sort(grp)[None, :])
            continue
        e = _delaunay_edges_3d(V[grp])
        if e.size:
            pairs.append(np.sort(grp[e], axis=1))
    fe = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]])
    pairs.append(np.sort(fe, axis=1))
    E = np.unique(np.concatenate(pairs), axis=0)
    return E[E[:, 0] != E[:, 1]]


def representative_graph(V, F, E, thres):
    """Subdivision::ComputeRepresentativeGraph (subdivision.cc:343-399): one graph node per occupied cell of size
    `thres` (cells in lexicographic key order, node = mean of its vertices); graph edges connect the nodes of
    every neighbour pair and every face edge.  Returns (reference [n], graphV [g,3], graphE [ge,2])."""
    key = _trunc_key(V, thres)
    _, ref = np.unique(key, axis=0, return_inverse=True)            # np.unique sorts rows lexicographically, like std::map
    ref = ref.reshape(-1)
    g = int(ref.max()) + 1 if ref.size else 0
    cnt = np.bincount(ref, minlength=g).astype(np.float64)
    GV = np.stack([np.bincount(ref, weights=V[:, j], minlength=g) / cnt for j in range(3)], 1)
    fe = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]])
    ge = np.sort(ref[np.concatenate([E, fe])], axis=1)
    ge = ge[ge[:, 0] != ge[:, 1]]
    return ref, GV, np.unique(ge, axis=0)


def load_cad_mesh(V, F, len_thres=2e-2, neighbor_thres=1.5e-2, graph_thres=1e-2):
    """LoadCadMesh (mesh_tensor.cc:102-178) on arrays: returns (V f32 [n,3], F i32 [m,3], E i32 [e,2],
    V2G i32 [n,1], GV f32 [g,3], GE i32 [ge,2])."""
    V = np.asarray(V, dtype=np.float64); F = np.asarray(F, dtype=np.int64).reshape(-1, 3)
    F = remove_degenerated(V, F)
    V, F = merge_duplex(V, F)
    V, F = subdivide(V, F, len_thres)
    E = geometry_neighbors(V, F, neighbor_thres)
    ref, GV, GE = representative_graph(V, F, E, graph_thres)
    return (np.ascontiguousarray(V, dtype=np.float32), np.ascontiguousarray(F, dtype=np.int32),
            np.ascontiguousarray(E, dtype=np.int32), np.ascontiguousarray(ref.reshape(-1, 1), dtype=np.int32),
            np.ascontiguousarray(GV, dtype=np.float32), np.ascontiguousarray(GE, dtype=np.int32))
