"""Known-answer tests of the CPU oracle (no GPU): closed-form point-triangle regions, exact
brute-force vs accelerated grid build, golden fixtures from the reference's shipped meshes,
finite-difference gradients, and the Adam restatement against the real torch.optim.Adam."""
import numpy as np
import pytest


# ---- point-triangle distance: the 7 Voronoi regions in closed form -----------------------------
A = np.array([0.0, 0.0, 0.0]); B = np.array([1.0, 0.0, 0.0]); C = np.array([0.0, 1.0, 0.0])


@pytest.mark.parametrize("p,closest", [
    ((0.25, 0.25, 0.7), (0.25, 0.25, 0.0)),      # face
    ((-0.5, -0.5, 0.3), (0.0, 0.0, 0.0)),        # vertex a
    ((2.0, -0.5, 0.1), (1.0, 0.0, 0.0)),         # vertex b
    ((-0.5, 2.0, -0.2), (0.0, 1.0, 0.0)),        # vertex c
    ((0.5, -1.0, 0.4), (0.5, 0.0, 0.0)),         # edge ab
    ((-1.0, 0.5, 0.4), (0.0, 0.5, 0.0)),         # edge ac
    ((1.0, 1.0, 0.5), (0.5, 0.5, 0.0)),          # edge bc
    ((0.0, 0.0, 0.0), (0.0, 0.0, 0.0)),          # on a vertex
    ((0.5, 0.5, 0.0), (0.5, 0.5, 0.0)),          # on the hypotenuse
])
def test_point_triangle_regions(oracle, p, closest):
    d, q = oracle.point_triangle_sqr(np.array(p), A, B, C)
    assert np.allclose(q, closest, atol=1e-15)
    assert d == pytest.approx(float(np.sum((np.array(p) - np.array(closest)) ** 2)), abs=1e-15)


def test_point_triangle_degenerate(oracle):
    # zero-area triangles (segment, point) must not produce NaN (documented deviation: 0/0 guarded)
    d, _ = oracle.point_triangle_sqr(np.array([0.5, 1.0, 0.0]), A, B, B)
    assert d == pytest.approx(1.0)
    d, _ = oracle.point_triangle_sqr(np.array([3.0, 4.0, 0.0]), A, A, A)
    assert d == pytest.approx(25.0)
    d, _ = oracle.point_triangle_sqr(np.array([0.5, 0.5, 0.5]), A, 0.5 * B, B)   # collinear
    assert np.isfinite(d) and d == pytest.approx(0.5)


def test_point_triangle_random_vs_sampling(oracle):
    rng = np.random.default_rng(5)
    u = rng.uniform(0, 1, size=(20000, 2))
    u[u.sum(1) > 1] = 1 - u[u.sum(1) > 1]
    for _ in range(20):
        a, b, c, p = rng.normal(0, 1, (4, 3))
        pts = a + u[:, :1] * (b - a) + u[:, 1:] * (c - a)
        dense = ((pts - p) ** 2).sum(1).min()
        d, q = oracle.point_triangle_sqr(p, a, b, c)
        assert d <= dense + 1e-12 and d >= dense - 2e-2 * max(dense, 1e-3)
        assert d == pytest.approx(((p - q) ** 2).sum(), rel=1e-13)


def test_point_triangle_vs_independent_formulation(oracle):
    """An independent algorithm for the same quantity (not Ericson's region walk): the minimum over the
    orthogonal projection onto the plane (if it falls inside, by barycentric coordinates from a 2x2 solve),
    the three clamped edge projections and the three vertices.  libigl is absent, so this is what the
    nearest-triangle restatement can be held against besides brute-force sampling."""
    rng = np.random.default_rng(11)

    def seg(p, a, b):
        ab = b - a
        t = np.clip(((p - a) @ ab) / max(ab @ ab, 1e-300), 0.0, 1.0)
        return (((a + t * ab) - p) ** 2).sum()

    worst = 0.0
    for k in range(4000):
        a, b, c, p = rng.normal(0, 1, (4, 3))
        if k % 4 == 1:            # slivers
            c = a + (b - a) * rng.uniform(-0.5, 1.5) + rng.normal(0, 1e-4, 3)
        if k % 4 == 2:            # points close to the plane
            u, v = rng.uniform(-0.3, 1.3, 2)
            n = np.cross(b - a, c - a)
            p = a + u * (b - a) + v * (c - a) + 1e-3 * rng.normal() * n / np.linalg.norm(n)
        cand = [seg(p, a, b), seg(p, b, c), seg(p, c, a)]
        ab, ac, ap = b - a, c - a, p - a
        G = np.array([[ab @ ab, ab @ ac], [ab @ ac, ac @ ac]])
        if abs(np.linalg.det(G)) > 1e-14 * G[0, 0] * G[1, 1]:
            s, t = np.linalg.solve(G, np.array([ab @ ap, ac @ ap]))
            if s >= 0 and t >= 0 and s + t <= 1:
                cand.append((((a + s * ab + t * ac) - p) ** 2).sum())
        want = min(cand)
        d, _ = oracle.point_triangle_sqr(p, a, b, c)
        worst = max(worst, abs(d - want) / max(want, 1e-30))
        assert abs(d - want) <= 1e-9 * max(want, 1e-12), (k, d, want)
    print("largest relative difference", worst)


# ---- grid build --------------------------------------------------------------------------------
def test_normalize_target(oracle, meshes):
    V = meshes["tarV"]
    Vn, scale, pos = oracle.normalize_target(V)
    V64 = V.astype(np.float64)
    ext = V64.max(0) - V64.min(0)
    assert scale == ext.max() * 1.1                              # mesh.cc:80-81
    assert np.array_equal(pos, V64.min(0) - 0.05 * scale)        # mesh.cc:82-83
    assert np.array_equal(Vn, (V64 - pos) / scale)
    j = int(np.argmax(ext))
    assert Vn[:, j].min() == pytest.approx(0.05) and Vn[:, j].max() == pytest.approx(0.05 + 1 / 1.1)


def test_grid_layout_single_triangle(oracle):
    """Voxel (i,j,k) holds the distance of point (k/N, j/N, i/N) (mesh.cc:112-120, :143-147)."""
    N = 8
    Vn = np.array([[0.25, 0.25, 0.5], [0.75, 0.25, 0.5], [0.25, 0.75, 0.5]])
    F = np.array([[0, 1, 2]], dtype=np.int32)
    grid, idx = oracle.build_grid(Vn, F, N, fast=False)
    assert (idx == 0).all()
    for (i, j, k) in [(0, 0, 0), (4, 2, 2), (7, 1, 6), (2, 7, 3)]:
        d, _ = oracle.point_triangle_sqr(np.array([k / N, j / N, i / N]), Vn[0], Vn[1], Vn[2])
        assert grid[i, j, k] == np.sqrt(d)
    assert grid[4, 2, 2] == 0.0 and grid[0, 2, 2] == 0.5


def test_grid_fast_equals_brute(oracle, meshes):
    """Both accelerated CPU builders -- the bounding-box tree (libigl's AABB query restated; the CPU baseline) and the
    uniform cell index with a ring search -- return the brute-force field and index bit for bit, ties included."""
    rng = np.random.default_rng(1)
    Vn, _, _ = oracle.normalize_target(meshes["cadTarV"])
    F = meshes["cadTarF"]
    for N in (7, 16):
        gb, ib = oracle.build_grid(Vn, F, N, fast=False)
        for mode in ("bvh", "cells"):
            gf, if_ = oracle.build_grid(Vn, F, N, fast=mode)
            assert np.array_equal(gb, gf) and np.array_equal(ib, if_), (N, mode)
    # random soup with triangles sticking out of the unit cube (plus exact duplicates: ties), slab build
    V = rng.uniform(-0.2, 1.2, size=(300, 3)); F = rng.integers(0, 300, size=(500, 3)).astype(np.int32)
    F = np.concatenate([F, F[:40]])
    gb, ib = oracle.build_grid(V, F, 12, fast=False)
    for mode in (True, "cells"):
        gf, if_ = oracle.build_grid(V, F, 12, fast=mode)
        assert np.array_equal(gb, gf) and np.array_equal(ib, if_), mode
        gs, _ = oracle.build_grid(V, F, 12, z0=3, z1=7, fast=mode)
        assert np.array_equal(gs[3:7], gb[3:7]) and (gs[:3] == 1e30).all() and (gs[7:] == 1e30).all()
    # the two accelerated builders against each other at a size brute force would take minutes for
    from meshode_b200.synth import synth_mesh
    Vs, Fs = synth_mesh(2000, 5)
    Vsn, _, _ = oracle.normalize_target(Vs)
    g1, i1 = oracle.build_grid(Vsn, Fs, 40, fast="bvh")
    g2, i2 = oracle.build_grid(Vsn, Fs, 40, fast="cells")
    assert np.array_equal(g1, g2) and np.array_equal(i1, i2)


def test_grid_matches_golden(oracle, meshes, golden):
    """The committed cfg1 fixture (data/target.obj, N = 32, FP64 brute force) is reproducible."""
    Vn, scale, trans = oracle.normalize_target(meshes["tarV"])
    assert scale == golden["scale"] and np.array_equal(trans, golden["trans"])
    assert np.array_equal(Vn[:64], golden["Vn_head"])
    grid, idx = oracle.build_grid(Vn, meshes["tarF"], int(golden["N"]), fast=True)
    assert np.array_equal(grid, golden["grid"]) and np.array_equal(idx, golden["nearest"])


# ---- losses --------------------------------------------------------------------------------------
def test_distance_layers_match_golden(oracle, meshes, golden):
    src_n = oracle.normalize_by_template(meshes["srcV"], float(golden["scale"]), golden["trans"])
    P = src_n[golden["sel"]]
    assert np.array_equal(oracle.distfield_forward(golden["grid"], P), golden["dist_fwd"])
    assert np.array_equal(oracle.distfield_backward(golden["grid"], P), golden["dist_bwd"])
    # forward = d^2 and backward = d * grad d of the float sampler
    v, g = oracle.distance_float_jet(golden["grid"], P)
    assert np.array_equal(golden["dist_fwd"], v * v)
    assert np.allclose(golden["dist_bwd"], v[:, None] * g, rtol=2e-6, atol=1e-12)


def test_gradient_finite_differences(oracle, golden):
    grid = golden["grid"]
    N = grid.shape[0]
    rng = np.random.default_rng(3)
    P = (rng.integers(1, N - 2, size=(400, 3)) + rng.uniform(0.2, 0.8, size=(400, 3))) / N   # cell interiors
    v, g = oracle.distance_double_jet(grid, P)
    h = 1e-7
    for k in range(3):
        d = np.zeros(3); d[k] = h
        fd = (oracle.distance_double(grid, P + d) - oracle.distance_double(grid, P - d)) / (2 * h)
        keep = v > 0     # the cut-off branch is a constant
        assert np.allclose(fd[keep], g[keep, k], rtol=1e-5, atol=1e-7)


def test_edge_layers_vs_numpy(oracle, meshes, golden):
    rng = np.random.default_rng(7)
    V0 = oracle.normalize_by_template(meshes["srcV"], float(golden["scale"]), golden["trans"])
    F = meshes["srcF"]
    V = (V0 + rng.normal(0, 2e-3, V0.shape)).astype(np.float32)
    rest = oracle.store_rigid(V0, F)
    v0 = F.reshape(-1); v1 = np.roll(F, -1, axis=1).reshape(-1)            # rigid_layer.cc:34-35
    assert np.array_equal(rest, V0[v1] - V0[v0])
    r = (V[v1] - V[v0]) - rest
    assert np.array_equal(oracle.rigid_forward(V, F, rest), r * r)
    gb = oracle.rigid_backward(V, F, rest)
    ref = np.zeros(V.shape, np.float64)
    np.add.at(ref, v0, -r.astype(np.float64)); np.add.at(ref, v1, r.astype(np.float64))
    assert np.allclose(gb, ref, rtol=0, atol=2e-6)
    assert np.array_equal(gb[::11], golden["rigid_bwd"])
    # CAD weights
    E = np.ascontiguousarray(np.stack([F[:, 0], F[:, 1]], axis=1)[:5000], dtype=np.int32)
    crest, lam = oracle.store_cad(V0, F[:3000], E)
    n = np.sqrt((crest.astype(np.float32) ** 2).sum(1, dtype=np.float32))
    assert np.allclose(lam, 2e-2 / (n.astype(np.float64) + 1e-8), rtol=1e-6)
    assert np.array_equal(lam[::17], golden["cad_lambda"])


def test_adam_restatement_vs_torch(oracle, meshes, golden):
    """orc_rigid_adam against the real torch.optim.Adam (float32, CPU) driving the oracle's loss
    gradient.  The oracle uses torch's operation order and FMAs (lerp, addcmul, addcdiv), so the
    moments agree bit for bit; torch's CPU sqrt (MKL VML) is not correctly rounded, which leaves
    rare 1-ulp differences in the parameters -- hence a tolerance of a few float32 ulps, far
    inside the 1e-4 Chamfer gate of the north star."""
    torch = pytest.importorskip("torch")
    grid = golden["grid"]
    src_all = oracle.normalize_by_template(meshes["srcV"], float(golden["scale"]), golden["trans"])
    src_n = src_all[:3000]
    F = meshes["srcF"]
    F = np.ascontiguousarray(F[(F < 3000).all(1)])
    rest = oracle.store_rigid(src_n, F)
    iters = 25
    got, _ = oracle.rigid_adam(grid, src_n, F, rest, iters, 1e-3)
    p = torch.nn.Parameter(torch.from_numpy(src_n.copy()))
    opt = torch.optim.Adam([p], lr=1e-3)
    for it in range(iters):
        opt.zero_grad()
        Vc = p.detach().numpy()
        g = oracle.distfield_backward(grid, Vc) + oracle.rigid_backward(Vc, F, rest)   # rigid_loss_layer.py:24-27
        p.grad = torch.from_numpy(g)
        opt.step()
        if it == 0:   # one step from identical state: moments identical, parameters within 1 ulp
            one, _ = oracle.rigid_adam(grid, src_n, F, rest, 1, 1e-3)
            assert np.abs(one - p.detach().numpy()).max() <= 6e-8
            w1 = np.float32(1.0 - 0.9)
            assert np.array_equal(opt.state[p]["exp_avg"].numpy(), (w1 * g.astype(np.float64)).astype(np.float32))
    want = p.detach().numpy()
    assert np.abs(want - src_n).max() > 1e-3
    assert np.abs(got - want).max() <= 5e-7, "max |d| = %g" % np.abs(got - want).max()
    assert (got == want).mean() > 0.95
    rest_all = oracle.store_rigid(src_all, meshes["srcF"])
    assert np.array_equal(oracle.rigid_adam(grid, src_all, meshes["srcF"], rest_all, 20, 1e-3)[0][::101], golden["adam20_rows"])
