"""GPU Levenberg-Marquardt (mo_ceres_solve, the ceres::Solve of Deformer::Deform / DeformWithRot /
DeformSubdivision, src/lib/deformer.cc) against the CPU restatement with a sparse direct solve
(oracle/lm.py).  Ceres itself is absent: the SOLVER's parity is unpinned; residuals and Jacobians are the
pinned functors."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def problem():
    from meshode_b200.synth import synth_pair
    srcV, srcF, tarV, tarF = synth_pair(7, 260, 400)
    return srcV, srcF, tarV, tarF


def _setup(oracle, pd, problem, N=32):
    srcV, srcF, tarV, tarF = problem
    tm = oracle.Template(tarV, tarF, N)
    pid = pd.InitializeDeformTemplate(_t(tarV), _t(tarF), 0, N)
    V0 = (srcV.astype(np.float64) - tm.trans) / tm.scale          # Mesh::ApplyTransform (mesh.cc:98-105)
    a = srcF.reshape(-1); b = np.roll(srcF, -1, axis=1).reshape(-1)
    rest = V0[a] - V0[b]                                          # deformer.cc:44 / :121
    I = np.stack([a, b], 1).astype(np.int32)
    return tm, pid, V0, I, rest


@pytest.mark.parametrize("kind_name", ["EDGE", "ADAPTIVE_EDGE", "ROT_EDGE"])
def test_lm_matches_direct_solve(oracle, pd, problem, kind_name):
    from meshode_b200 import capi
    from oracle import lm
    tm, pid, V0, I, rest = _setup(oracle, pd, problem)
    kind = getattr(capi, "CERES_" + kind_name)
    lam = 1.0
    iters = 12 if kind_name == "ROT_EDGE" else 25
    R0 = np.zeros_like(V0)                                        # deformer.cc:118
    V = _t(V0.copy()); R = _t(R0.copy())
    s = pd.CeresSolve(pid, kind, V, R if kind_name == "ROT_EDGE" else None, _t(I), _t(rest), lam, max_iterations=iters,
                      cg_tolerance=1e-12)
    oV, oR, os_ = lm.solve(tm.grid, getattr(lm, kind_name), V0, R0, I, rest, lam, max_iterations=iters)
    print(kind_name, s, os_)
    assert s["final_cost"] < s["initial_cost"]
    assert abs(s["initial_cost"] - os_["initial_cost"]) <= 1e-12 * os_["initial_cost"]
    assert s["iterations"] == os_["iterations"] and s["accepted"] == os_["accepted"]
    assert s["termination"] == os_["termination"]
    assert abs(s["final_cost"] - os_["final_cost"]) <= 1e-7 * os_["final_cost"]
    assert np.abs(V.cpu().numpy() - oV).max() <= 1e-6
    if kind_name == "ROT_EDGE":
        assert np.abs(R.cpu().numpy() - oR).max() <= 1e-5
    # the summary's split is what Problem::Evaluate gives at the solution (deformer.cc:76-91)
    cost, _, _ = pd.CeresProblem(pid, kind, V, R if kind_name == "ROT_EDGE" else None, _t(I), _t(rest), lam)
    assert abs(cost[0].item() - s["vertices_cost"]) <= 1e-12 * max(s["vertices_cost"], 1e-30)
    assert abs(cost[1].item() - s["rigidity_cost"]) <= 1e-12 * max(s["rigidity_cost"], 1e-30)
    pd.DestroyTemplate(pid)


def test_lm_argument_checks(pd):
    from meshode_b200 import capi
    V = torch.zeros((4, 3), dtype=torch.float64, device="cuda")
    I = torch.zeros((2, 2), dtype=torch.int32, device="cuda")
    rest = torch.zeros((2, 3), dtype=torch.float64, device="cuda")
    with pytest.raises(Exception):
        pd.CeresSolve(10 ** 6, capi.CERES_EDGE, V, None, I, rest, 1.0)       # unknown template
    with pytest.raises(Exception):
        pd.CeresSolve(None, capi.CERES_ROT_EDGE, V, None, I, rest, 1.0)      # ROT needs R
    s = pd.CeresSolve(None, capi.CERES_EDGE, V, None, I, rest, 1.0)          # no distance term: already optimal
    assert s["final_cost"] == 0.0 and s["termination"] == "gradient tolerance"
