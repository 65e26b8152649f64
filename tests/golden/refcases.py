"""Seeded inputs shared by make_golden_ref.py (which evaluates them with the reference's own
compiled sampler / functors, oracle/_ref) and tests/test_oracle_ref.py (which evaluates them
with the oracle restatement and, when oracle/_ref is present, with the reference again)."""
import numpy as np


def sampler_cases(golden_grid):
    """Returns {name: (grid [N,N,N] f64, P [n,3] f64)} covering SURVEY s8c's edge cases."""
    rng = np.random.default_rng(20261017)
    cases = {}
    # 1. the cfg1 distance field (data/target.obj, N = 32): interior, near-surface, cut-off and OOB points
    N = golden_grid.shape[0]
    P = rng.uniform(-0.15, 1.15, size=(4000, 3))
    cases["cfg1_uniform"] = (golden_grid, P)
    # 2. random field with values straddling the 0.2 cut-off
    N2 = 9
    g2 = rng.uniform(0.0, 0.4, size=(N2, N2, N2))
    P2 = rng.uniform(-0.2, 1.2, size=(3000, 3))
    cases["random_cutoff"] = (g2, P2)
    # 3. points exactly on cell boundaries (p*N integral), including 0, (N-2)/N, (N-1)/N and 1
    k = np.arange(-2, N2 + 3, dtype=np.float64) / N2
    X, Y, Z = np.meshgrid(k, k[::3], k[::4], indexing="ij")
    P3 = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    cases["lattice"] = (g2 * 0.45, P3)
    # 4. negative fractions in (-1/N, 0): C-cast truncation gives index 0 and a negative weight
    P4 = rng.uniform(0.05, 0.8, size=(500, 3))
    ax = rng.integers(0, 3, size=500)
    P4[np.arange(500), ax] = -rng.uniform(0.0, 1.0 / N2, size=500)
    cases["negative_fraction"] = (g2 * 0.3, P4)
    # 5. dead band [ (N-1)/N, 1 ) and the >= N penalty, one or several axes
    P5 = rng.uniform(0.05, 0.8, size=(600, 3))
    ax = rng.integers(0, 3, size=600)
    P5[np.arange(600), ax] = rng.uniform((N2 - 1.0) / N2, 1.3, size=600)
    P5[:100, (ax[:100] + 1) % 3] = rng.uniform(1.0, 1.2, size=100)
    P5[100:200, (ax[100:200] + 2) % 3] = -rng.uniform(0.0, 0.3, size=100)
    cases["deadband_penalty"] = (g2 * 0.3, P5)
    # 6. an affine field: trilinear interpolation must reproduce it exactly up to rounding
    zz, yy, xx = np.meshgrid(np.arange(N2), np.arange(N2), np.arange(N2), indexing="ij")
    g6 = (0.01 + 0.004 * xx + 0.007 * yy + 0.002 * zz).astype(np.float64)
    cases["affine"] = (g6, rng.uniform(0.0, (N2 - 1.0) / N2, size=(800, 3)))
    return cases


def functor_cases():
    """Inputs of the Ceres functors: (p1, p2, rot1, rot2, v, lambda) rows; rot covers the
    small-angle branch (0 and ~1e-9), the threshold region and ordinary rotations."""
    rng = np.random.default_rng(77)
    n = 64
    p1 = rng.normal(0.5, 0.2, size=(n, 3)); p2 = rng.normal(0.5, 0.2, size=(n, 3))
    v = rng.normal(0.0, 0.02, size=(n, 3))
    rot1 = rng.normal(0.0, 0.5, size=(n, 3)); rot2 = rng.normal(0.0, 0.5, size=(n, 3))
    rot1[:8] = 0.0
    rot1[8:16] = rng.normal(0.0, 1e-9, size=(8, 3))
    rot1[16:24] = rng.normal(0.0, 1.2e-8, size=(8, 3))   # theta^2 around DBL_EPSILON
    lam = rng.uniform(0.1, 5.0, size=n)
    return p1, p2, rot1, rot2, v, lam
