"""GPU parity of the device-native loss layers (the mirror of src/python/layers), the nearest-vertex
search behind ReverseLossLayer and the FP64 Ceres loss terms, against the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def pair():
    from meshode_b200.synth import synth_pair, unique_edges
    srcV, srcF, tarV, tarF = synth_pair(5, 1500, 1300)
    return srcV, srcF, tarV, tarF, unique_edges(srcF)


def _moved(V, seed=3):
    rng = np.random.default_rng(seed)
    return (V + rng.normal(0, 3e-3, V.shape)).astype(np.float32)


def _close(a, b, rtol=1e-5):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    scale = max(np.abs(b).max(), 1e-30)
    assert np.abs(a - b).max() <= rtol * scale, "max |d| = %g vs scale %g" % (np.abs(a - b).max(), scale)


def test_rigid_layer_matches_reference_composition(oracle, pd, pair):
    from meshode_b200.layers import Finalize, RigidLossLayer
    srcV, srcF, tarV, tarF, _ = pair
    N = 32
    sV = torch.from_numpy(srcV.copy())     # CPU tensor as in rigid_deform.py: normalised in place by the layer
    layer = RigidLossLayer(sV, torch.from_numpy(srcF), torch.from_numpy(tarV), torch.from_numpy(tarF), grid_resolution=N)
    tm = oracle.Template(tarV, tarF, N)
    src_n = oracle.normalize_by_template(srcV, tm.scale, tm.trans)
    assert np.array_equal(sV.numpy(), src_n)
    rest = oracle.store_rigid(src_n, srcF)
    mv = _moved(src_n)
    p = torch.nn.Parameter(_t(mv))
    loss = layer(p, _t(srcF))
    (2.5 * loss).backward()
    want = 0.5 * oracle.distfield_forward(tm.grid, mv).astype(np.float64).sum() + \
        0.5 * oracle.rigid_forward(mv, srcF, rest).astype(np.float64).sum()        # rigid_loss_layer.py:11-17
    assert abs(loss.item() - want) <= 1e-5 * want                                  # tolerance: float32 scalar
    g = oracle.distfield_backward(tm.grid, mv) + oracle.rigid_backward(mv, srcF, rest)   # :24-27
    assert np.array_equal(p.grad.cpu().numpy(), np.float32(2.5) * g)
    out = p.detach().clone()
    Finalize(out, layer.param_id)
    assert np.array_equal(out.cpu().numpy(), oracle.denormalize_by_template(mv, tm.scale, tm.trans))


def test_graph_layers_mask_and_cross_templates(oracle, pd, pair):
    from meshode_b200.layers import GraphLoss2Layer, GraphLossLayer
    srcV, srcF, tarV, tarF, E = pair
    N, rig = 32, 1.7
    layer = GraphLossLayer(_t(srcV.copy()), _t(E), _t(tarV), _t(tarF), rig, grid_resolution=N)
    tm = oracle.Template(tarV, tarF, N)
    src_n = oracle.normalize_by_template(srcV, tm.scale, tm.trans)
    rest = oracle.store_graph(src_n, E)
    mv = _moved(src_n, 9)
    mv[::7] += np.float32(0.04)            # push some vertices beyond the 0.03 mask
    p = torch.nn.Parameter(_t(mv))
    loss = layer(p, _t(E))
    loss.backward()
    lossD = oracle.distfield_forward(tm.grid, mv) * np.float32(0.5)
    mask = lossD < np.float32(0.5 * 0.03 * 0.03)                                   # graph_loss_layer.py:18
    assert mask.any() and (~mask).any()
    want = lossD.astype(np.float64).sum() + 0.5 * oracle.graph_forward(mv, E, rest).astype(np.float64).sum() * rig * rig
    assert abs(loss.item() - want) <= 1e-5 * want
    r2 = np.float32(rig * rig)
    g = oracle.distfield_backward(tm.grid, mv) * mask[:, None] + oracle.graph_backward(mv, E, rest) * r2   # :40-42
    assert np.array_equal(p.grad.cpu().numpy(), g.astype(np.float32))

    # GraphLoss2: distance field of the OTHER mesh, rest edges of its own (graph_loss2_layer.py:18-19)
    from meshode_b200.synth import unique_edges
    E2 = unique_edges(tarF)
    l2 = GraphLoss2Layer(_t(srcV), _t(srcF), _t(srcV.copy()), _t(E), _t(tarV), _t(tarF), _t(tarV.copy()), _t(E2), rig,
                         grid_resolution=N)
    tm1 = oracle.Template(srcV, srcF, N)
    g1 = oracle.normalize_by_template(srcV, tm1.scale, tm1.trans)
    g2 = oracle.normalize_by_template(tarV, tm.scale, tm.trans)
    mv1 = _moved(g1, 4)
    p1 = torch.nn.Parameter(_t(mv1))
    l2(p1, _t(E), _t(g2), _t(E2), 0).backward()
    want1 = oracle.distfield_backward(tm.grid, mv1) + oracle.graph_backward(mv1, E, oracle.store_graph(g1, E)) * r2
    assert np.array_equal(p1.grad.cpu().numpy(), want1.astype(np.float32))
    mv2 = _moved(g2, 6)
    p2 = torch.nn.Parameter(_t(mv2))
    l2(_t(g1), _t(E), p2, _t(E2), 1).backward()
    want2 = oracle.distfield_backward(tm1.grid, mv2) + oracle.graph_backward(mv2, E2, oracle.store_graph(g2, E2)) * r2
    assert np.array_equal(p2.grad.cpu().numpy(), want2.astype(np.float32))


def test_cad_layer(oracle, pd, pair):
    from meshode_b200.layers import CadLossLayer
    srcV, srcF, tarV, tarF, E = pair
    N = 24
    layer = CadLossLayer(_t(srcV.copy()), _t(srcF), _t(E[:700]), _t(tarV), _t(tarF), grid_resolution=N)
    tm = oracle.Template(tarV, tarF, N)
    src_n = oracle.normalize_by_template(srcV, tm.scale, tm.trans)
    rest, lam = oracle.store_cad(src_n, srcF, E[:700])
    mv = _moved(src_n, 12)
    p = torch.nn.Parameter(_t(mv))
    loss = layer(p)
    loss.backward()
    want = 0.5 * oracle.distfield_forward(tm.grid, mv).astype(np.float64).sum() + \
        0.5 * oracle.cad_forward(mv, srcF, E[:700], rest, lam).astype(np.float64).sum()
    assert abs(loss.item() - want) <= 1e-5 * want
    g = oracle.distfield_backward(tm.grid, mv) + oracle.cad_backward(mv, srcF, E[:700], rest, lam)
    assert np.array_equal(p.grad.cpu().numpy(), g)


def test_nearest_vertex_and_reverse_layer(pd, pair):
    from scipy.spatial import cKDTree
    from meshode_b200.layers import ReverseLossLayer
    srcV, _, tarV, _, _ = pair
    rng = np.random.default_rng(0)
    for P, Q in ((srcV, tarV), (rng.normal(size=(4097, 3)).astype(np.float32), rng.normal(size=(1031, 3)).astype(np.float32)),
                 (srcV[:1], tarV[:5])):
        dd, ii = cKDTree(P.astype(np.float64)).query(Q.astype(np.float64), k=1)      # reverse_loss_layer.py:18-19
        idx, d2 = pd.NearestVertex(_t(Q), _t(P), return_dist2=True)
        idx = idx.cpu().numpy(); d2 = d2.cpu().numpy()
        exact = ((P[idx].astype(np.float64) - Q.astype(np.float64)) ** 2).sum(1)
        assert np.allclose(d2, exact, rtol=1e-15, atol=0)
        assert np.allclose(np.sqrt(d2), dd, rtol=1e-14, atol=0)                      # same minimum distance
        same = idx == ii
        assert same.mean() > 0.999 and np.allclose(exact[~same], dd[~same] ** 2, rtol=1e-14)   # differences are exact ties
    # duplicate points: the lowest index wins
    P = np.repeat(srcV[:50], 2, axis=0)
    idx = pd.NearestVertex(_t(srcV[:50]), _t(P)).cpu().numpy()
    assert np.array_equal(idx, 2 * np.arange(50))
    layer = ReverseLossLayer()
    s = torch.nn.Parameter(_t(srcV))
    loss = layer(s, _t(tarV))
    loss.backward()
    dd, ii = cKDTree(srcV).query(tarV, k=1)
    want = 0.5 * ((srcV[ii] - tarV).astype(np.float64) ** 2).sum()
    assert abs(loss.item() - want) <= 1e-5 * want
    g = np.zeros_like(srcV, dtype=np.float64)
    np.add.at(g, ii, (srcV[ii] - tarV).astype(np.float64))
    _close(s.grad.cpu().numpy(), g, 1e-5)


def test_ceres_edge_blocks(oracle, pd):
    from meshode_b200 import capi
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from refcases import functor_cases
    p1, p2, rot1, rot2, v, lam = functor_cases()
    n = len(lam)
    V = np.concatenate([p1, p2]); R = np.concatenate([rot1, rot2])
    I = np.stack([np.arange(n), n + np.arange(n)], 1).astype(np.int32)
    for i in (0, 9, 17, 40):      # one lambda per launch: compare block i
        res, jac = pd.CeresEdges(capi.CERES_ROT_EDGE, _t(V), _t(R), _t(I), _t(v), lam[i], True)
        r, J = oracle.edge_rot(p1[i], p2[i], rot1[i], rot2[i], v[i], lam[i])
        # tolerance 1e-12 relative: CUDA's sin/cos differ from glibc's in the last bit
        _close(res[i].cpu().numpy(), r, 1e-12)
        _close(jac[i].cpu().numpy(), J, 1e-12)
        for kind, adaptive in ((capi.CERES_EDGE, False), (capi.CERES_ADAPTIVE_EDGE, True)):
            res, jac = pd.CeresEdges(kind, _t(V), None, _t(I), _t(v), lam[i], True)
            r, le = oracle.edge_loss(p1[i], p2[i], v[i], lam[i], adaptive)
            assert np.array_equal(res[i].cpu().numpy(), r)
            assert np.array_equal(jac[i].cpu().numpy(), np.concatenate([np.eye(3) * le, -np.eye(3) * le], 1))
    # small-angle rows (rot1 = 0): exactly the first-order branch
    res, jac = pd.CeresEdges(capi.CERES_ROT_EDGE, _t(V), _t(R), _t(I), _t(v), 1.0, True)
    r, J = oracle.edge_rot(p1[0], p2[0], rot1[0], rot2[0], v[0], 1.0)
    assert np.array_equal(res[0].cpu().numpy(), r) and np.array_equal(jac[0].cpu().numpy(), J)


def test_ceres_problems_cfg3_style(oracle, pd, pair):
    """Cost and gradient of Deformer::Deform / DeformWithRot (src/lib/deformer.cc:32-53, :118-131)."""
    from meshode_b200 import capi
    srcV, srcF, tarV, tarF, _ = pair
    N = 32
    tm = oracle.Template(tarV, tarF, N)
    pid = pd.InitializeDeformTemplate(_t(tarV), _t(tarF), 0, N)
    V0 = ((srcV.astype(np.float64) - tm.trans) / tm.scale)          # Mesh::ApplyTransform (mesh.cc:98-105)
    a = srcF.reshape(-1); b = np.roll(srcF, -1, axis=1).reshape(-1)
    rest = V0[a] - V0[b]                                            # deformer.cc:44 / :121
    I = np.stack([a, b], 1).astype(np.int32)
    rng = np.random.default_rng(2)
    V = V0 + rng.normal(0, 2e-3, V0.shape)
    R = rng.normal(0, 0.05, V0.shape); R[::3] = 0.0
    lam = 1.3
    cost, gV, gR = pd.CeresProblem(pid, capi.CERES_ROT_EDGE, _t(V), _t(R), _t(I), _t(rest), lam)
    cd, ce, oV, oR = oracle.rot_problem_cost_grad(tm.grid, V, R, srcF, rest, lam)
    assert abs(cost[0].item() - cd) <= 1e-12 * cd and abs(cost[1].item() - ce) <= 1e-11 * ce
    _close(gV.cpu().numpy(), oV, 1e-11)
    _close(gR.cpu().numpy(), oR, 1e-11)
    for kind, adaptive in ((capi.CERES_EDGE, False), (capi.CERES_ADAPTIVE_EDGE, True)):
        cost, gV, _ = pd.CeresProblem(pid, kind, _t(V), None, _t(I), _t(rest), lam)
        cd, ce, oV = oracle.deform_problem_cost_grad(tm.grid, V, srcF, rest, lam, adaptive)
        assert abs(cost[0].item() - cd) <= 1e-12 * cd and abs(cost[1].item() - ce) <= 1e-11 * ce
        _close(gV.cpu().numpy(), oV, 1e-11)
    pd.DestroyTemplate(pid)


def test_neuralode_flow_with_graph_loss(pd, pair):
    """cad_neural_deform2.py:57-76 in miniature: the MLP flow on cuda, the loss without leaving the GPU."""
    from meshode_b200.layers import GraphLoss2Layer, NeuralODE, ReverseLossLayer
    from meshode_b200.synth import unique_edges
    srcV, srcF, tarV, tarF, E = pair
    E2 = unique_edges(tarF)
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    GV1, GV2 = _t(srcV.copy()), _t(tarV.copy())
    layer = GraphLoss2Layer(_t(srcV), _t(srcF), GV1, _t(E), _t(tarV), _t(tarF), GV2, _t(E2), 1.0, dev, grid_resolution=32)
    func = NeuralODE(dev)
    opt = torch.optim.Adam(func.parameters(), lr=1e-3)
    rev = ReverseLossLayer()
    GV1o, GV2o = GV1.clone(), GV2.clone()
    losses = []
    for it in range(6):
        opt.zero_grad()
        d1 = func.forward(GV1); d2 = func.inverse(GV2)
        loss = layer(d1, None, GV2, None, 0) + rev(d1, GV2o, dev) + layer(GV1, None, d2, None, 1) + rev(d2, GV1o, dev)
        loss.backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in func.parameters())
        opt.step()
        losses.append(loss.item())
    assert np.isfinite(losses).all() and losses[-1] < losses[0]
    # RK4 (3/8 rule) is exact for constant fields and fourth order in general
    from meshode_b200.layers.neuralode import odeint_rk4
    y = odeint_rk4(lambda t, y: -y, torch.ones(1, dtype=torch.float64), torch.tensor([0.0, 0.5, 1.0], dtype=torch.float64))
    assert abs(y[-1].item() - np.exp(-1.0)) < 1e-3


def test_exported_autograd_functions_are_differentiable(oracle, pd, pair):
    """The reference exports RigidLossFunction / GraphLossFunction / GraphLoss2Function / CadLossFunction as autograd
    Functions with a working backward (rigid_loss_layer.py:7-27, graph_loss_layer.py:9-43, graph_loss2_layer.py:9-41,
    cad_loss_layer.py:7-27): ``X.apply(...).backward()`` must give the gradient the layers give."""
    from meshode_b200.layers import (CadLossFunction, GraphLoss2Function, GraphLossFunction, RigidLossFunction)
    import pyDeform as top
    srcV, srcF, tarV, tarF, E = pair
    N = 32
    pid = pd.InitializeDeformTemplate(_t(tarV), _t(tarF), 0, N)
    pid2 = pd.InitializeDeformTemplate(_t(srcV), _t(srcF), 0, N)
    tm = oracle.Template(tarV, tarF, N)
    src_n = oracle.normalize_by_template(srcV, tm.scale, tm.trans)
    mv = _moved(src_n)
    F_, E_ = _t(srcF), _t(E)
    pidt = torch.tensor(pid)

    # rigid
    pd.StoreRigidityInformation(_t(src_n), F_, pid)
    p = torch.nn.Parameter(_t(mv))
    loss = RigidLossFunction.apply(p, F_, pidt)
    (3.0 * loss).backward()
    rest = oracle.store_rigid(src_n, srcF)
    g = oracle.distfield_backward(tm.grid, mv) + oracle.rigid_backward(mv, srcF, rest)
    assert np.array_equal(p.grad.cpu().numpy(), np.float32(3.0) * g)

    # graph (mask 0.5*0.03^2 on the distance gradient, rigidity^2 on the edge term)
    pd.StoreGraphInformation(_t(src_n), E_, pid)
    p = torch.nn.Parameter(_t(mv))
    GraphLossFunction.apply(p, E_, torch.tensor(1.7 * 1.7), pidt).backward()
    grest = oracle.store_graph(src_n, E)
    d = oracle.distfield_backward(tm.grid, mv)
    d[~(0.5 * oracle.distfield_forward(tm.grid, mv) < np.float32(0.5 * 0.03 * 0.03))] = 0
    want = d + oracle.graph_backward(mv, E, grest) * np.float32(1.7 * 1.7)
    _close(p.grad.cpu().numpy(), want, 1e-6)

    # graph2: distance field of the OTHER template (pid2), edges stored in its own (pid)
    p = torch.nn.Parameter(_t(mv))
    GraphLoss2Function.apply(p, E_, torch.tensor(1.0), pidt, torch.tensor(pid2)).backward()
    g64_2, _, _ = pd.GetGrid(pid2)
    want = oracle.distfield_backward(g64_2.cpu().numpy(), mv) + oracle.graph_backward(mv, E, grest)
    assert np.array_equal(p.grad.cpu().numpy(), want)

    # cad
    pd.StoreCadInformation(_t(src_n), F_, E_, pid)
    p = torch.nn.Parameter(_t(mv))
    CadLossFunction.apply(p, F_, E_, pidt).backward()
    crest, lam = oracle.store_cad(src_n, srcF, E)
    want = oracle.distfield_backward(tm.grid, mv) + oracle.cad_backward(mv, srcF, E, crest, lam)
    assert np.array_equal(p.grad.cpu().numpy(), want)
    assert top.DistanceFieldLoss_forward is pd.DistanceFieldLoss_forward   # the top-level shim is the same module surface
    pd.DestroyTemplate(pid); pd.DestroyTemplate(pid2)
