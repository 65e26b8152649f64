"""GPU parity tests: the CUDA path, called through the C-ABI (ctypes -> libmeshode_b200.so),
against the CPU oracle on the same inputs.  Integer/index work and every FP64-refined or
contraction-free quantity is compared bit for bit; the tolerances that remain are written
next to the assertion."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _t(a, dev="cuda"):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _tie_aware_index_check(Vn, F, grid, idx_gpu, idx_ref, O):
    """Indices must agree wherever the runner-up is more than 1e-6 (relative) away."""
    diff = np.argwhere(idx_gpu != idx_ref)
    N = grid.shape[0]
    for (z, y, x) in diff[:200]:
        p = np.array([x / N, y / N, z / N])
        a, b = idx_gpu[z, y, x], idx_ref[z, y, x]
        da = O.point_triangle_sqr(p, *[Vn[F[a, k]] for k in range(3)])[0]
        db = O.point_triangle_sqr(p, *[Vn[F[b, k]] for k in range(3)])[0]
        assert abs(np.sqrt(da) - np.sqrt(db)) <= 1e-6 * max(np.sqrt(db), 1e-300), (z, y, x, a, b, da, db)
    return len(diff)


@pytest.fixture(scope="module")
def cfg1(meshes, oracle, pd):
    """Template of data/target.obj at N=32 on the GPU and its brute-force oracle twin."""
    tarV, tarF = meshes["tarV"], meshes["tarF"]
    pid = pd.InitializeDeformTemplate(_t(tarV), _t(tarF), 0, 32)
    Vn, scale, trans = oracle.normalize_target(tarV)
    return dict(pid=pid, Vn=Vn, scale=scale, trans=trans, tarV=tarV, tarF=tarF)


def test_normalize_exact(cfg1, pd):
    info = pd.GetTemplateInfo(cfg1["pid"])
    assert info["N"] == 32 and info["nV"] == cfg1["tarV"].shape[0] and info["nF"] == cfg1["tarF"].shape[0]
    assert info["scale"] == cfg1["scale"]                       # FP64, bit exact (mesh.cc:80-83)
    assert np.array_equal(np.array(info["trans"]), cfg1["trans"])


def test_grid_matches_golden_bruteforce(cfg1, golden, pd, oracle):
    g64, g32, idx = [t.cpu().numpy() for t in pd.GetGrid(cfg1["pid"])]
    ref = golden["grid"]
    # distances: north_star tolerance is 1e-5 relative; the FP64 refinement is in fact bit exact
    np.testing.assert_allclose(g64, ref, rtol=1e-5, atol=0)
    assert np.array_equal(g64, ref), "FP64 grid not bit-identical: max rel %g" % np.max(np.abs(g64 - ref) / ref)
    assert np.array_equal(g32, ref.astype(np.float32))
    nd = _tie_aware_index_check(cfg1["Vn"], cfg1["tarF"], ref, idx, golden["nearest"], oracle)
    assert nd == 0, "%d nearest indices differ (all within tie tolerance)" % nd


@pytest.mark.parametrize("N", [64, 40])
def test_grid_cfg1_full_resolution(meshes, oracle, pd, N):
    tarV, tarF = meshes["tarV"], meshes["tarF"]
    pid = pd.InitializeDeformTemplate(_t(tarV), _t(tarF), 0, N)
    g64, g32, idx = [t.cpu().numpy() for t in pd.GetGrid(pid)]
    Vn, _, _ = oracle.normalize_target(tarV)
    ref, ridx = oracle.build_grid(Vn, tarF, N, fast=True)
    np.testing.assert_allclose(g64, ref, rtol=1e-5, atol=0)
    assert np.array_equal(g64, ref)
    assert np.array_equal(idx, ridx)
    capi = __import__("meshode_b200.capi", fromlist=["x"])
    stats = capi.template_build_stats(pid)
    assert stats["fp32_tests"] == 0 and stats["fp64_tests"] == 0   # the plain build does not count
    pd.DestroyTemplate(pid)
    # the instrumented instantiation of the search kernel: the same field, plus its test counts
    assert capi.lib().mo_build_stats_enable(1) == 0
    try:
        pid = pd.InitializeDeformTemplate(_t(tarV), _t(tarF), 0, N)
    finally:
        assert capi.lib().mo_build_stats_enable(0) == 1
    h64, _, hidx = [t.cpu().numpy() for t in pd.GetGrid(pid)]
    assert np.array_equal(h64, ref) and np.array_equal(hidx, ridx)
    stats = capi.template_build_stats(pid)
    assert stats["fp32_tests"] > 0 and stats["fp64_tests"] >= N ** 3
    pd.DestroyTemplate(pid)


def test_grid_synthetic_and_slabs(oracle, pd):
    from meshode_b200.synth import synth_mesh
    V, F = synth_mesh(2000, 3)
    N = 48
    pid = pd.InitializeDeformTemplate(_t(V), _t(F), 0, N)
    g64, g32, idx = [t.cpu().numpy() for t in pd.GetGrid(pid)]
    Vn, _, _ = oracle.normalize_target(V)
    ref, ridx = oracle.build_grid(Vn, F, N, fast=True)
    assert np.array_equal(g64, ref) and np.array_equal(idx, ridx)
    # z-slab build: assembled slabs == single build, bit for bit
    from meshode_b200 import capi
    parts = []
    dV, dF = _t(V), _t(F)   # keep the device buffers alive across the raw-pointer calls
    for z0, z1 in ((0, 13), (13, 32), (32, 48)):
        p = capi.template_create_slab(dV.data_ptr(), V.shape[0], dF.data_ptr(), F.shape[0], N, z0, z1,
                                      torch.cuda.current_stream().cuda_stream)
        s64, s32, sidx = pd.GetGrid(p)
        full = s64.cpu().numpy()
        assert np.all(full[z0:z1] == ref[z0:z1])
        outside = np.delete(full, np.s_[z0:z1], axis=0)
        assert np.all(outside == 1e30)                      # uniformgrid.cc:9-17 initial value
        parts.append(full[z0:z1])
        pd.DestroyTemplate(p)
    assert np.array_equal(np.concatenate(parts, 0), ref)
    pd.DestroyTemplate(pid)


def test_grid_degenerate_and_coarse_meshes(meshes, oracle, pd):
    # CAD target: 272 huge triangles; plus slivers / zero-area triangles appended
    V, F = meshes["cadTarV"].copy(), meshes["cadTarF"].copy()
    extra_v = np.array([[0.0, 0.0, 0.0], [0.1, 0.0, 0.0], [0.2, 0.0, 0.0],           # collinear
                        [0.05, 0.05, 0.05], [0.05, 0.05, 0.05], [0.05, 0.05, 0.05],  # point triangle
                        [0.0, 0.1, 0.0], [0.2, 0.1, 0.0], [0.1, 0.1000001, 0.0]],    # sliver
                       dtype=np.float32)
    n0 = V.shape[0]
    V = np.concatenate([V, extra_v])
    F = np.concatenate([F, np.arange(9, dtype=np.int32).reshape(3, 3) + n0])
    N = 24
    pid = pd.InitializeDeformTemplate(_t(V), _t(F), 0, N)
    g64, _, idx = [t.cpu().numpy() for t in pd.GetGrid(pid)]
    Vn, _, _ = oracle.normalize_target(V)
    ref, ridx = oracle.build_grid(Vn, F, N, fast=False)
    np.testing.assert_allclose(g64, ref, rtol=1e-5, atol=1e-300)
    assert np.array_equal(g64, ref)
    _tie_aware_index_check(Vn, F, ref, idx, ridx, oracle)
    pd.DestroyTemplate(pid)


def test_distance_forward_backward_bit_exact(cfg1, meshes, golden, oracle, pd):
    src = oracle.normalize_by_template(meshes["srcV"], cfg1["scale"], cfg1["trans"])
    rng = np.random.default_rng(0)
    # real vertices + random points incl. out-of-cube, negative fractions, exact cell boundaries
    extra = rng.uniform(-0.2, 1.2, size=(4096, 3)).astype(np.float32)
    edge = np.array([[0.0, 0.0, 0.0], [1.0, 1.0, 1.0], [31 / 32, 0.5, 0.5], [0.5, 30.999 / 32, 0.5], [-1e-3, 0.3, 0.3],
                     [-1 / 32, 0.3, 0.3], [-1.5 / 32, 0.3, 0.3], [0.25, 0.25, 0.25], [1.03125, 0.5, 0.5], [0.5, 0.5, 2.0]],
                    dtype=np.float32)
    P = np.concatenate([src, extra, edge]).astype(np.float32)
    grid = golden["grid"]
    fwd = pd.DistanceFieldLoss_forward(_t(P), cfg1["pid"]).cpu().numpy()
    bwd = pd.DistanceFieldLoss_backward(_t(P), cfg1["pid"]).cpu().numpy()
    f2, b2 = [t.cpu().numpy() for t in pd.DistanceFieldLoss_forward_backward(_t(P), cfg1["pid"])]
    rf, rb = oracle.distfield_forward(grid, P), oracle.distfield_backward(grid, P)
    # tolerance demanded: 1e-5 relative; achieved: identical bits (same operation sequence, no contraction)
    np.testing.assert_allclose(fwd, rf, rtol=1e-5, atol=1e-12)
    np.testing.assert_allclose(bwd, rb, rtol=1e-5, atol=1e-5 * np.abs(rb).max())
    assert np.array_equal(fwd, rf) and np.array_equal(bwd, rb)
    assert np.array_equal(f2, rf) and np.array_equal(b2, rb)
    # golden rows
    sel = golden["sel"]
    assert np.array_equal(fwd[: src.shape[0]][sel], golden["dist_fwd"])
    assert np.array_equal(bwd[: src.shape[0]][sel], golden["dist_bwd"])
    # CPU tensors are staged through the GPU and come back on the CPU
    cpu_out = pd.DistanceFieldLoss_forward(torch.from_numpy(P), cfg1["pid"])
    assert cpu_out.device.type == "cpu" and np.array_equal(cpu_out.numpy(), rf)


def test_distance_f64_bit_exact(cfg1, golden, oracle, pd):
    from meshode_b200 import capi
    rng = np.random.default_rng(1)
    P = rng.uniform(-0.1, 1.1, size=(5000, 3))
    tP = _t(P)
    val = torch.empty(P.shape[0], dtype=torch.float64, device="cuda")
    grad = torch.empty((P.shape[0], 3), dtype=torch.float64, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    capi.check(capi.lib().mo_distance_f64(tP.data_ptr(), P.shape[0], cfg1["pid"], val.data_ptr(), grad.data_ptr(), s))
    rv, rg = oracle.distance_double_jet(golden["grid"], P)
    assert np.array_equal(val.cpu().numpy(), rv) and np.array_equal(grad.cpu().numpy(), rg)
    capi.check(capi.lib().mo_distance_f64(tP.data_ptr(), P.shape[0], cfg1["pid"], val.data_ptr(), 0, s))
    assert np.array_equal(val.cpu().numpy(), oracle.distance_double(golden["grid"], P))


def test_normalize_by_template(cfg1, meshes, oracle, pd):
    V = meshes["srcV"].copy()
    t = _t(V)
    pd.NormalizeByTemplate(t, cfg1["pid"])
    ref = oracle.normalize_by_template(V, cfg1["scale"], cfg1["trans"])
    assert np.array_equal(t.cpu().numpy(), ref)
    pd.DenormalizeByTemplate(t, cfg1["pid"])
    assert np.array_equal(t.cpu().numpy(), oracle.denormalize_by_template(ref, cfg1["scale"], cfg1["trans"]))
    c = torch.from_numpy(V.copy())            # CPU tensor: in place through staging
    pd.NormalizeByTemplate(c, cfg1["pid"])
    assert np.array_equal(c.numpy(), ref)


def _moved(src_n, seed=7):
    rng = np.random.default_rng(seed)
    return (src_n + rng.normal(0, 2e-3, src_n.shape)).astype(np.float32)


def test_rigid_edges(cfg1, meshes, golden, oracle, pd):
    src_n = oracle.normalize_by_template(meshes["srcV"], cfg1["scale"], cfg1["trans"])
    F = meshes["srcF"]
    tF = _t(F)
    pd.StoreRigidityInformation(_t(src_n), tF, cfg1["pid"])
    moved = _moved(src_n)
    rest = oracle.store_rigid(src_n, F)
    fwd = pd.RigidEdgeLoss_forward(_t(moved), tF, cfg1["pid"]).cpu().numpy()
    bwd = pd.RigidEdgeLoss_backward(_t(moved), tF, cfg1["pid"]).cpu().numpy()
    rf, rb = oracle.rigid_forward(moved, F, rest), oracle.rigid_backward(moved, F, rest)
    assert fwd.shape == (3 * F.shape[0], 3) and bwd.shape == (src_n.shape[0], 3)
    assert np.array_equal(fwd, rf)
    assert np.array_equal(bwd, rb), "CSR gather must reproduce the serial scatter order bit for bit"
    assert np.array_equal(fwd[::97], golden["rigid_fwd_rows"]) and np.array_equal(bwd[::11], golden["rigid_bwd"])
    # edge-parallel variant with warp-aggregated atomics: float32 summation order differs.
    # tolerance: 1e-5 relative with atol = 1e-5 * max|g| (gradient components cancel)
    atom = pd.EdgeLoss_backward_atomic(0, _t(moved), tF, None, cfg1["pid"]).cpu().numpy()
    np.testing.assert_allclose(atom, rb, rtol=1e-5, atol=1e-5 * np.abs(rb).max())


def test_graph_and_cad_edges(cfg1, meshes, golden, oracle, pd):
    src_n = oracle.normalize_by_template(meshes["srcV"], cfg1["scale"], cfg1["trans"])
    F = meshes["srcF"]
    E = np.ascontiguousarray(np.stack([F[:, 0], F[:, 1]], axis=1)[:5000], dtype=np.int32)
    moved = _moved(src_n)
    tE = _t(E)
    pd.StoreGraphInformation(_t(src_n), tE, cfg1["pid"])
    rest = oracle.store_graph(src_n, E)
    gf = pd.GraphEdgeLoss_forward(_t(moved), tE, cfg1["pid"]).cpu().numpy()
    gb = pd.GraphEdgeLoss_backward(_t(moved), tE, cfg1["pid"]).cpu().numpy()
    assert np.array_equal(gf, oracle.graph_forward(moved, E, rest))
    assert np.array_equal(gb, oracle.graph_backward(moved, E, rest))
    assert np.array_equal(gf[::13], golden["graph_fwd"]) and np.array_equal(gb[::11], golden["graph_bwd"])
    # storing another kind overwrites the edge set (rigid_layer.cc:29 / cad_layer.cc:33)
    Fc = F[:3000]
    tFc = _t(Fc)
    pd.StoreCadInformation(_t(src_n), tFc, tE, cfg1["pid"])
    crest, clam = oracle.store_cad(src_n, Fc, E)
    cf = pd.CadEdgeLoss_forward(_t(moved), tFc, tE, cfg1["pid"]).cpu().numpy()
    cb = pd.CadEdgeLoss_backward(_t(moved), tFc, tE, cfg1["pid"]).cpu().numpy()
    assert cf.shape == (E.shape[0] + 3 * Fc.shape[0], 3)
    assert np.array_equal(cf, oracle.cad_forward(moved, Fc, E, crest, clam))
    assert np.array_equal(cb, oracle.cad_backward(moved, Fc, E, crest, clam))
    assert np.array_equal(cf[::13], golden["cad_fwd"]) and np.array_equal(cb[::11], golden["cad_bwd"])
    atom = pd.EdgeLoss_backward_atomic(2, _t(moved), tFc, tE, cfg1["pid"]).cpu().numpy()
    rb = oracle.cad_backward(moved, Fc, E, crest, clam)
    np.testing.assert_allclose(atom, rb, rtol=1e-5, atol=1e-5 * np.abs(rb).max())
    with pytest.raises(Exception):      # graph edges are gone now
        pd.GraphEdgeLoss_forward(_t(moved), tE, cfg1["pid"])


def test_edge_cases_and_errors(cfg1, pd):
    empty = torch.empty((0, 3), dtype=torch.float32, device="cuda")
    assert pd.DistanceFieldLoss_forward(empty, cfg1["pid"]).shape == (0,)
    assert pd.DistanceFieldLoss_backward(empty, cfg1["pid"]).shape == (0, 3)
    with pytest.raises(Exception):
        pd.DistanceFieldLoss_forward(empty, 12345)                       # bad handle
    with pytest.raises(TypeError):
        pd.DistanceFieldLoss_forward(empty.double(), cfg1["pid"])        # dtype
    with pytest.raises(ValueError):
        pd.DistanceFieldLoss_forward(torch.zeros((4, 4), device="cuda"), cfg1["pid"])
    with pytest.raises(Exception):
        pd.InitializeDeformTemplate(empty, torch.empty((0, 3), dtype=torch.int32, device="cuda"), 0, 16)


def test_fused_loss_matches_layer_composition(cfg1, meshes, golden, oracle, pd):
    src_n = oracle.normalize_by_template(meshes["srcV"], cfg1["scale"], cfg1["trans"])
    F = meshes["srcF"]
    pd.StoreRigidityInformation(_t(src_n), _t(F), cfg1["pid"])
    moved = _moved(src_n)
    rest = oracle.store_rigid(src_n, F)
    loss, grad = pd.LossForwardBackward(_t(moved), cfg1["pid"], cfg1["pid"], 1.0, 0.0)
    gD, gR = oracle.distfield_backward(golden["grid"], moved), oracle.rigid_backward(moved, F, rest)
    assert np.array_equal(grad.cpu().numpy(), gD + gR)                   # rigid_loss_layer.py:27
    ref_loss = 0.5 * oracle.distfield_forward(golden["grid"], moved).astype(np.float64).sum() + \
        0.5 * oracle.rigid_forward(moved, F, rest).astype(np.float64).sum()
    assert abs(loss.item() - ref_loss) <= 1e-9 * abs(ref_loss)           # FP64 accumulation, order differs
    # graph-layer form: gradient mask at 0.5*0.03^2 and rigidity^2 weight (graph_loss_layer.py:18,40-42)
    w = 2.5 ** 2
    loss2, grad2 = pd.LossForwardBackward(_t(moved), cfg1["pid"], cfg1["pid"], w, 0.5 * 0.03 * 0.03)
    lossD = oracle.distfield_forward(golden["grid"], moved) * np.float32(0.5)
    mask = (lossD < np.float32(0.5 * 0.03 * 0.03)).astype(np.float32)[:, None]
    assert np.array_equal(grad2.cpu().numpy(), gD * mask + gR * np.float32(w))


def test_interface_entries_match_the_references_own_loops(golden, pd):
    """The CUDA path against outputs of the REFERENCE's own src/interface/{distance,rigid,graph,cad}_layer.cc and
    normalize.cc (compiled in place into oracle/_ref, tests/golden/make_golden_iface.py -> golden_iface.npz), with no
    restatement in between: SURVEY.md s8 rows a11-a16, bit for bit."""
    import os
    import sys
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, gdir)
    from ifacecases import SCALE, TRANS, iface_case
    from meshode_b200 import capi
    ref = dict(np.load(os.path.join(gdir, "golden_iface.npz")))
    V, F, E, moved, raw = iface_case()
    grid = golden["grid"]
    N = grid.shape[0]
    # a template that holds the fixture's field, scale and translation (what InitializeDeformTemplate would leave)
    tri = _t(np.array([[0.3, 0.3, 0.3], [0.6, 0.3, 0.3], [0.3, 0.6, 0.3]], dtype=np.float64))
    pid = capi.template_create_normalized(tri.data_ptr(), 3, _t(np.array([[0, 1, 2]], dtype=np.int32)).data_ptr(), 1, N, SCALE, TRANS,
                                          torch.cuda.current_stream().cuda_stream)
    pd.SetGrid(pid, _t(grid), _t(grid.astype(np.float32)), _t(np.zeros(grid.shape, dtype=np.int32)))
    same = lambda t, k: np.array_equal(t.cpu().numpy(), ref[k], equal_nan=True)  # noqa: E731
    v = _t(raw); pd.NormalizeByTemplate(v, pid); assert same(v, "normalize")
    v = _t(V); pd.DenormalizeByTemplate(v, pid); assert same(v, "denormalize")
    dV, dF, dE, dM = _t(V), _t(F), _t(E), _t(moved)
    assert same(pd.DistanceFieldLoss_forward(dM, pid), "dist_fwd") and same(pd.DistanceFieldLoss_backward(dM, pid), "dist_bwd")
    pd.StoreRigidityInformation(dV, dF, pid)
    assert same(pd.RigidEdgeLoss_forward(dM, dF, pid), "rigid_fwd") and same(pd.RigidEdgeLoss_backward(dM, dF, pid), "rigid_bwd")
    pd.StoreGraphInformation(dV, dE, pid)
    assert same(pd.GraphEdgeLoss_forward(dM, dE, pid), "graph_fwd") and same(pd.GraphEdgeLoss_backward(dM, dE, pid), "graph_bwd")
    pd.StoreCadInformation(dV, dF, dE, pid)
    assert same(pd.CadEdgeLoss_forward(dM, dF, dE, pid), "cad_fwd") and same(pd.CadEdgeLoss_backward(dM, dF, dE, pid), "cad_bwd")
    # the fused per-iteration entry composes the same two gradients (rigid_loss_layer.py:27)
    pd.StoreRigidityInformation(dV, dF, pid)
    _, g = pd.LossForwardBackward(dM, pid, pid, 1.0, 0.0)
    assert np.array_equal(g.cpu().numpy(), ref["dist_bwd"] + ref["rigid_bwd"])
    pd.DestroyTemplate(pid)
