"""Seeded inputs of the interface-loop fixtures (shared by make_golden_iface.py and the tests): a 32^3 field (the cfg1
fixture's), a small closed mesh placed inside the unit cube with some vertices pushed outside it (out-of-bounds
branch, negative fractions, the N-1 dead band), its unique edges, and a moved copy."""
import numpy as np

SCALE, TRANS = 1.37, np.array([0.1, -0.2, 0.3])


def iface_case():
    from meshode_b200.synth import synth_mesh, unique_edges
    rng = np.random.default_rng(20261018)
    V, F = synth_mesh(900, 3)
    V = (V * 0.35 + 0.5).astype(np.float32)
    V[:60] += rng.normal(0, 0.4, (60, 3)).astype(np.float32)
    V[60:70, 0] = np.float32(31.0 / 32.0) + rng.uniform(0, 1.0 / 32.0, 10).astype(np.float32)   # index N-1: dead band
    V[70:80, 1] = -rng.uniform(0, 1.0 / 32.0, 10).astype(np.float32)                           # negative fraction, index 0
    E = unique_edges(F).astype(np.int32)
    moved = (V + np.float32(2e-3) * np.sin(np.float32(37.0) * V)).astype(np.float32)
    raw = (V * np.float32(SCALE) + TRANS.astype(np.float32)).astype(np.float32)                  # input of NormalizeByTemplate
    return V, F.astype(np.int32), E, moved, raw
