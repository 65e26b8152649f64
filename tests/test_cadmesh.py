"""CPU: LoadCadMesh (host-side CAD preprocessing, src/lib/subdivision.cc without CGAL) on the shipped CAD target
(data/cad-target.obj, 70 vertices / 272 faces, kept in tests/golden/meshes.npz).  The Delaunay triangulation is
unique only up to co-circular ties, so the checks are the construction rules, not an index-for-index listing."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")


def _area(V, F):
    a, b, c = (V[F[:, k]].astype(np.float64) for k in range(3))
    return 0.5 * np.sqrt((np.cross(b - a, c - a) ** 2).sum(1)).sum()


def test_load_cad_mesh_rules(tmp_path, meshes):
    import pyDeform
    from meshode_b200 import cadmesh
    V0, F0 = meshes["cadTarV"].astype(np.float64), meshes["cadTarF"]
    path = str(tmp_path / "cad.obj")
    with open(path, "w") as fh:
        for v in V0:
            fh.write("v %.17g %.17g %.17g\n" % tuple(v))
        for f in F0:
            fh.write("f %d %d %d\n" % tuple(int(x) + 1 for x in f))
    V, F, E, V2G, GV, GE = [t.numpy() for t in pyDeform.LoadCadMesh(path)]
    assert V.dtype == np.float32 and F.dtype == np.int32 and E.dtype == np.int32 and V2G.shape == (V.shape[0], 1)
    Fc = cadmesh.remove_degenerated(V0, F0.astype(np.int64)); Vc, Fc = cadmesh.merge_duplex(V0, Fc)
    # the cleaned input vertices come first, unchanged; the re-triangulation tiles every face exactly
    assert np.array_equal(V[:Vc.shape[0]], Vc.astype(np.float32))
    assert abs(_area(V, F) - _area(Vc, Fc)) <= 1e-5 * _area(Vc, Fc)
    assert F.min() >= 0 and F.max() < V.shape[0] and len(np.unique(F)) == V.shape[0]      # every vertex is used
    # subdivision: no triangle edge above 3 * 2e-2 (subdivision.cc:238-241), typical edges at the 2e-2 lattice
    el = np.sqrt(((V[F].astype(np.float64) - V[np.roll(F, -1, axis=1)]) ** 2).sum(2))
    assert el.max() <= 3 * 2e-2 + 1e-6 and np.median(el) <= 2.9e-2
    # consistent orientation with the input: summed normals agree
    def nsum(V, F):
        a, b, c = (V[F[:, k]].astype(np.float64) for k in range(3))
        return np.cross(b - a, c - a).sum(0)
    assert np.allclose(nsum(V, F), nsum(Vc, Fc), atol=1e-5)
    # neighbour pairs: sorted, unique, contain every face edge, stay local (cells of 1.5e-2 around a vertex)
    assert (E[:, 0] < E[:, 1]).all() and len(np.unique(E, axis=0)) == len(E)
    fe = np.sort(np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]]), axis=1)
    Eset = set(map(tuple, E))
    assert all(tuple(e) in Eset for e in fe[::7])
    fset = set(map(tuple, fe))
    non_face = np.array([e for e in E if tuple(e) not in fset])
    if len(non_face):
        assert np.sqrt(((V[non_face[:, 0]] - V[non_face[:, 1]]).astype(np.float64) ** 2).sum(1)).max() <= 2 * 1.5e-2 * np.sqrt(3) + 1e-6
    # deformation graph: node = mean of its vertices, one node per occupied 1e-2 cell, edges between distinct nodes
    ref = V2G.reshape(-1)
    assert ref.min() == 0 and ref.max() == GV.shape[0] - 1
    for g in (0, GV.shape[0] // 2, GV.shape[0] - 1):
        assert np.allclose(GV[g], V[ref == g].astype(np.float64).mean(0), atol=1e-6)
    keys = np.trunc(V.astype(np.float64) / 1e-2).astype(np.int64)
    assert len(np.unique(keys, axis=0)) in (GV.shape[0], GV.shape[0] + 1, GV.shape[0] - 1)   # float32 export may move a vertex across a cell wall
    assert (GE[:, 0] < GE[:, 1]).all() and GE.max() < GV.shape[0]
    gfe = np.sort(ref[fe], axis=1)
    assert set(map(tuple, gfe[gfe[:, 0] != gfe[:, 1]][::5])) <= set(map(tuple, GE))


def test_delaunay_edges_in_every_dimension():
    from meshode_b200 import cadmesh
    rng = np.random.default_rng(0)
    line = np.outer(np.array([0.0, 1.0, 3.0, 2.0]), np.array([1.0, 2.0, -1.0]))            # collinear: a chain
    assert sorted(map(tuple, np.sort(cadmesh._delaunay_edges_3d(line), axis=1).tolist())) == [(0, 1), (1, 3), (2, 3)]
    plane = np.concatenate([rng.uniform(0, 1, (12, 2)), np.zeros((12, 1))], axis=1) @ np.linalg.qr(rng.normal(size=(3, 3)))[0]
    e2 = cadmesh._delaunay_edges_3d(plane)                                                  # coplanar: planar triangulation
    assert len(e2) <= 3 * 12 - 6 and len(e2) >= 12 - 1
    vol = rng.uniform(0, 1, (15, 3))
    e3 = cadmesh._delaunay_edges_3d(vol)
    assert len(e3) >= 15 - 1 and (e3[:, 0] < e3[:, 1]).all()


def test_mesh_cleanup_rules():
    """Mesh::RemoveDegenerated (mesh.cc:320-334) and Mesh::MergeDuplex (mesh.cc:181-233)."""
    from meshode_b200 import cadmesh
    V = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 0, 0.0000001], [2, 0, 0], [0, 0, 1]], dtype=np.float64)
    F = np.array([[0, 1, 2], [0, 1, 4], [0, 3, 2], [2, 1, 0], [0, 2, 5], [1, 3, 5]], dtype=np.int64)
    F1 = cadmesh.remove_degenerated(V, F)
    assert F1.tolist() == [[0, 1, 2], [0, 3, 2], [2, 1, 0], [0, 2, 5], [1, 3, 5]]       # the collinear face goes
    V2, F2 = cadmesh.merge_duplex(V, F1)
    # vertex 3 equals vertex 1 after int(v * 1e6); first appearance wins; later indices shift down
    assert V2.shape == (5, 3) and np.array_equal(V2[3], V[4]) and np.array_equal(V2[4], V[5])
    # [0,3,2] and [2,1,0] repeat the vertex set of [0,1,2]; [1,3,5] collapses to a repeated vertex
    assert F2.tolist() == [[0, 1, 2], [0, 2, 4]]
