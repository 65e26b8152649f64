"""CPU: SolveLinear (host-side sparse post-process, src/lib/linear.cc) against its literal dense restatement."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")


def _case(n=140, g=9, seed=0):
    from meshode_b200.synth import synth_mesh
    V, F = synth_mesh(n, seed)
    rng = np.random.default_rng(seed)
    ref = rng.integers(0, g, size=n).astype(np.int32); ref[:g] = np.arange(g)      # every node owns a vertex
    GV = np.stack([V[ref == i].mean(0) for i in range(g)]) + rng.normal(0, 0.05, (g, 3))
    E = rng.integers(0, n, size=(25, 2)).astype(np.int32); E = E[E[:, 0] != E[:, 1]]
    return V, F, E, ref, GV.astype(np.float32)


def test_linear_estimation_matches_dense_restatement():
    from meshode_b200 import linear
    from oracle import linear as ref_linear
    V, F, E, ref, GV = _case()
    for rigidity in (1.0, 0.1):
        got = linear.linear_estimation(V, F, E, ref, GV, rigidity)
        want = ref_linear.linear_estimation(V, F, E, ref, GV, rigidity)
        assert np.abs(got - want).max() <= 1e-7 * np.abs(want).max()
    TV = V + np.random.default_rng(1).normal(0, 0.03, V.shape)
    got = linear.linear_estimation_with_rot(V, F, TV, 1.0)
    want = ref_linear.linear_estimation_with_rot(V, F, TV, 1.0)
    assert np.abs(got - want).max() <= 1e-9 * np.abs(want).max()
    # a rigid motion of the targets is reproduced exactly by the rotation-aware solve (scale 1, R = the rotation)
    c, s = np.cos(0.4), np.sin(0.4)
    Rz = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
    Vd = V.astype(np.float64)
    moved = linear.linear_estimation_with_rot(Vd, F, Vd @ Rz.T + 0.3, 5.0)
    assert np.abs(moved - (Vd @ Rz.T + 0.3)).max() <= 1e-7      # the 1e-8 guard in the scale (linear.cc:155)


def test_pydeform_solve_linear_in_place():
    import pyDeform
    from oracle import linear as ref_linear
    V, F, E, ref, GV = _case(seed=2)
    tV = torch.from_numpy(V.copy())
    pyDeform.SolveLinear(tV, torch.from_numpy(F), torch.from_numpy(E), torch.from_numpy(ref.reshape(-1, 1)), torch.from_numpy(GV), 1.0, 0)
    want = ref_linear.linear_estimation(V, F, E, ref, GV, 1.0)
    assert np.abs(tV.numpy() - want).max() <= 1e-5 * np.abs(want).max()       # float32 tensor in, float32 out
    with pytest.raises(ValueError):
        pyDeform.SolveLinear(tV, torch.from_numpy(F), torch.from_numpy(E), torch.from_numpy(ref.reshape(-1, 1)), torch.from_numpy(GV), 1.0, 1)
