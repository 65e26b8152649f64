"""CPU: the Levenberg-Marquardt restatement (oracle/lm.py) on a small Deformer problem -- monotone cost,
consistent bookkeeping, and a stationary point when run to convergence."""
import numpy as np


def test_lm_restatement_converges(oracle):
    from meshode_b200.synth import synth_pair
    from oracle import lm
    srcV, srcF, tarV, tarF = synth_pair(7, 120, 200)
    tm = oracle.Template(tarV, tarF, 16)
    V0 = (srcV.astype(np.float64) - tm.trans) / tm.scale
    a = srcF.reshape(-1); b = np.roll(srcF, -1, axis=1).reshape(-1)
    rest = V0[a] - V0[b]
    I = np.stack([a, b], 1).astype(np.int32)
    for kind in (lm.EDGE, lm.ADAPTIVE_EDGE):
        log = []
        V, _, s = lm.solve(tm.grid, kind, V0, None, I, rest, 1.0, max_iterations=100, log=log)
        assert s["final_cost"] < s["initial_cost"]
        assert s["termination"] in ("function tolerance", "gradient tolerance", "parameter tolerance")
        accepted_costs = [s["initial_cost"]] + [c for (_, _, c, rho, _) in log if rho > 1e-3]
        assert all(x >= y for x, y in zip(accepted_costs, accepted_costs[1:]))
        cd, ce, g = oracle.deform_problem_cost_grad(tm.grid, V, srcF, rest, 1.0, kind == lm.ADAPTIVE_EDGE)
        assert abs((cd + ce) - s["final_cost"]) <= 1e-12 * s["final_cost"]
    # EdgeLossWithRot: a few iterations, cost and gradient agree with the pinned problem evaluation
    V, R, s = lm.solve(tm.grid, lm.ROT_EDGE, V0, np.zeros_like(V0), I, rest, 1.0, max_iterations=3)
    cd, ce, gV, gR = oracle.rot_problem_cost_grad(tm.grid, V, R, srcF, rest, 1.0)
    assert abs((cd + ce) - s["final_cost"]) <= 1e-12 * s["final_cost"] and s["final_cost"] < s["initial_cost"]
