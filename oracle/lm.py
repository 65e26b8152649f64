"""TEST INFRASTRUCTURE (not product code): CPU restatement of ceres::Solve for the Deformer problems
(reference src/lib/deformer.cc:18-92 Deform, :94-171 DeformWithRot, :173-257 DeformSubdivision).

Ceres itself (ceres-solver @ d93fac4b, un-vendored) is absent, so its trust-region Levenberg-Marquardt loop is
restated from its published algorithm (trust_region_minimizer.cc, levenberg_marquardt_strategy.cc, default
Solver::Options with max_num_iterations = 100): parity of the SOLVER is therefore unpinned; the residual
blocks and their Jacobians are the pinned oracle functors (oracle.edge_loss / edge_rot /
distance_double_jet).  The normal equations are assembled with scipy.sparse and solved with a sparse
direct factorisation, like Ceres' SPARSE_NORMAL_CHOLESKY.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import oracle as O

EDGE, ADAPTIVE_EDGE, ROT_EDGE = 0, 1, 2


def _linearize(grid, kind, V, R, I, rest, lam):
    """Residual vector r and sparse Jacobian J over x = [V.ravel(), R.ravel()] (R only for ROT)."""
    nV, nE = V.shape[0], I.shape[0]
    rot = kind == ROT_EDGE
    n = (6 if rot else 3) * nV
    d, gd = O.distance_double_jet(grid, V) if grid is not None else (np.zeros(nV), np.zeros((nV, 3)))
    rows, cols, vals = [], [], []
    rows.append(np.repeat(np.arange(nV), 3)); cols.append(np.arange(3 * nV)); vals.append(gd.ravel())
    res = [d]
    a, b = I[:, 0].astype(np.int64), I[:, 1].astype(np.int64)
    if not rot:
        if kind == ADAPTIVE_EDGE:
            le = lam * (2e-2 / (np.sqrt((rest * rest).sum(1)) + 1e-8))
        else:
            le = np.full(nE, float(lam))
        r = ((V[a] - V[b]) - rest) * le[:, None]
        res.append(r.ravel())
        base = nV + 3 * np.arange(nE)
        for c in range(3):
            rows += [base + c, base + c]; cols += [3 * a + c, 3 * b + c]; vals += [le, -le]
    else:
        r = np.empty((nE, 6)); J = np.empty((nE, 6, 12))
        for e in range(nE):
            r[e], J[e] = O.edge_rot(V[a[e]], V[b[e]], R[a[e]], R[b[e]], rest[e], lam)
        res.append(r.ravel())
        base = nV + 6 * np.arange(nE)
        colblocks = [3 * a, 3 * b, 3 * nV + 3 * a, 3 * nV + 3 * b]
        for m in range(6):
            for blk in range(4):
                for c in range(3):
                    rows.append(base + m); cols.append(colblocks[blk] + c); vals.append(J[:, m, 3 * blk + c])
    r = np.concatenate(res)
    J = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(r.size, n))
    return r, J, 0.5 * float(d @ d), 0.5 * float(res[1] @ res[1])


def solve(grid, kind, V, R, I, rest, lam, max_iterations=100, log=None):
    """Returns (V, R, summary).  V, R float64 [n,3] (R ignored / None unless ROT_EDGE)."""
    V = np.array(V, dtype=np.float64); nV = V.shape[0]
    rot = kind == ROT_EDGE
    R = np.array(R, dtype=np.float64) if rot else None
    unpack = lambda x: (x[:3 * nV].reshape(nV, 3), x[3 * nV:].reshape(nV, 3) if rot else None)  # noqa: E731
    x = np.concatenate([V.ravel(), R.ravel()]) if rot else V.ravel().copy()
    r, J, cd, ce = _linearize(grid, kind, *unpack(x), I, rest, lam)
    cost = cd + ce
    g = J.T @ r
    diag = np.asarray(J.multiply(J).sum(0)).ravel()
    scale = 1.0 / (1.0 + np.sqrt(diag))                 # Jacobi scaling, taken once
    radius, decrease = 1e4, 2.0
    summary = {"initial_cost": cost, "iterations": 0, "accepted": 0, "termination": "iteration limit"}
    invalid = 0
    if np.abs(g).max() <= 1e-10:
        summary["termination"] = "gradient tolerance"
    else:
        while summary["iterations"] < max_iterations:
            summary["iterations"] += 1
            s2 = scale * scale
            damp = np.clip(diag * s2, 1e-6, 1e32) / radius / s2
            H = (J.T @ J + sp.diags(damp)).tocsc()
            delta = spla.spsolve(H, -g)
            Jd = J @ delta
            model_change = -(g @ delta + 0.5 * (Jd @ Jd))
            if not model_change > 0.0:
                invalid += 1
                if invalid >= 5:
                    summary["termination"] = "invalid steps"; break
                radius /= decrease; decrease *= 2.0
                continue
            invalid = 0
            if np.linalg.norm(delta) <= 1e-8 * (np.linalg.norm(x) + 1e-8):
                summary["termination"] = "parameter tolerance"; break
            xn = x + delta
            rn, Jn, cdn, cen = _linearize(grid, kind, *unpack(xn), I, rest, lam)
            new_cost = cdn + cen
            rho = (cost - new_cost) / model_change
            if log is not None:
                log.append((summary["iterations"], cost, new_cost, rho, radius))
            # TrustRegionMinimizer::Minimize order: ParameterToleranceReached (above), FunctionToleranceReached on the
            # CANDIDATE (accepted or not; the step is not applied), then IsStepSuccessful
            if abs(cost - new_cost) <= 1e-6 * cost:
                summary["termination"] = "function tolerance"; break
            if rho > 1e-3:
                t = 2.0 * rho - 1.0
                radius = min(1e16, radius / max(1.0 / 3.0, 1.0 - t * t * t)); decrease = 2.0
                x, r, J, cd, ce, cost = xn, rn, Jn, cdn, cen, new_cost
                g = J.T @ r
                diag = np.asarray(J.multiply(J).sum(0)).ravel()
                summary["accepted"] += 1
                if np.abs(g).max() <= 1e-10:
                    summary["termination"] = "gradient tolerance"; break
            else:
                radius /= decrease; decrease *= 2.0
                if radius < 1e-32:
                    summary["termination"] = "radius underflow"; break
    Vn, Rn = unpack(x)
    summary.update(final_cost=cost, vertices_cost=cd, rigidity_cost=ce)
    return Vn.copy(), (Rn.copy() if rot else None), summary
