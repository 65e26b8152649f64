"""Pins the CPU oracle (oracle/meshode_oracle.cc) against the REFERENCE's own compiled sampler and
loss functors: always through tests/golden/golden_ref.npz (generated from oracle/_ref by
tests/golden/make_golden_ref.py), and live against oracle/_ref when it can be built here
(/root/reference present).  Bit-exact: the oracle restates the same operation sequence."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from refcases import functor_cases, sampler_cases  # noqa: E402


@pytest.fixture(scope="module")
def gref():
    return dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_ref.npz")))


@pytest.fixture(scope="module")
def cases(golden):
    return sampler_cases(golden["grid"])


def _same(a, b):
    return np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("name", ["cfg1_uniform", "random_cutoff", "lattice", "negative_fraction", "deadband_penalty", "affine"])
def test_sampler_matches_reference_golden(oracle, gref, cases, name):
    grid, P = cases[name]
    assert _same(oracle.distance_double(grid, P), gref[name + "/d64"])
    v, g = oracle.distance_double_jet(grid, P)
    assert _same(v, gref[name + "/j64v"]) and _same(g, gref[name + "/j64g"])
    P32 = P.astype(np.float32)
    assert _same(oracle.distance_float(grid, P32), gref[name + "/d32"])
    v, g = oracle.distance_float_jet(grid, P32)
    assert _same(v, gref[name + "/j32v"]) and _same(g, gref[name + "/j32g"])
    # DistanceLoss: residual = [distance, 0, 0], Jacobian row 0 = the Jet partials
    v, g = oracle.distance_double_jet(grid, P[:50])
    assert _same(gref[name + "/dl_r"][:, 0], v) and not gref[name + "/dl_r"][:, 1:].any()
    assert _same(gref[name + "/dl_J"][:, 0, :], g) and not gref[name + "/dl_J"][:, 1:, :].any()


def test_case_coverage(gref, cases):
    """The fixtures really exercise every branch of the sampler."""
    grid, P = cases["deadband_penalty"]
    N = grid.shape[0]
    idx = (P * N).astype(np.int64)
    assert (idx == N - 1).any() and (idx >= N).any() and (P < 0).any()
    assert (gref["random_cutoff/d64"] == 0).any() and (gref["random_cutoff/d64"] > 0.19).any()
    grid, P = cases["negative_fraction"]
    assert ((P > -1.0 / grid.shape[0]) & (P < 0)).any()
    # affine field is reproduced by trilinear interpolation
    grid, P = cases["affine"]
    N = grid.shape[0]
    expect = 0.01 + 0.004 * P[:, 0] * N + 0.007 * P[:, 1] * N + 0.002 * P[:, 2] * N
    assert np.allclose(gref["affine/d64"], expect, rtol=0, atol=1e-14)
    assert np.allclose(gref["affine/j64g"], np.array([0.004, 0.007, 0.002]) * N, rtol=0, atol=1e-12)


def test_functors_match_reference_golden(oracle, gref):
    p1, p2, rot1, rot2, v, lam = functor_cases()
    for i in range(len(lam)):
        r, le = oracle.edge_loss(p1[i], p2[i], v[i], lam[i], False)
        assert _same(r, gref["edge/r"][i]) and le == gref["edge/lam"][i]
        r, le = oracle.edge_loss(p1[i], p2[i], v[i], lam[i], True)
        assert _same(r, gref["aedge/r"][i]) and le == gref["aedge/lam"][i]
        r, J = oracle.edge_rot(p1[i], p2[i], rot1[i], rot2[i], v[i], lam[i])
        assert _same(r, gref["rot/r"][i]) and _same(J, gref["rot/J"][i])


def test_live_reference_if_present(oracle, cases):
    """Where /root/reference exists, compile it (oracle/_ref) and compare on fresh random inputs."""
    from oracle import ref as R
    if not R.available():
        pytest.skip("no /root/reference and no prebuilt oracle/_ref here")
    rng = np.random.default_rng(123)
    for N in (4, 17, 32):
        grid = rng.uniform(0.0, 0.3, size=(N, N, N))
        P = rng.uniform(-0.3, 1.3, size=(5000, 3))
        G = R.Grid(grid)
        assert _same(G.distance_double(P), oracle.distance_double(grid, P))
        rv, rg = G.distance_double_jet(P)
        ov, og = oracle.distance_double_jet(grid, P)
        assert _same(rv, ov) and _same(rg, og)
        P32 = P.astype(np.float32)
        assert _same(G.distance_float(P32), oracle.distance_float(grid, P32))
        rv, rg = G.distance_float_jet(P32)
        ov, og = oracle.distance_float_jet(grid, P32)
        assert _same(rv, ov) and _same(rg, og)
    for _ in range(200):
        a, b, r1, r2, v = (rng.normal(0, 0.7, 3) for _ in range(5))
        lam = float(rng.uniform(0.1, 3))
        assert _same(R.edge_rot(a, b, r1, r2, v, lam)[1], oracle.edge_rot(a, b, r1, r2, v, lam)[1])
        assert R.edge_loss(a, b, v, lam, True)[1] == oracle.edge_loss(a, b, v, lam, True)[1]


# ---- the interface loops of the pyDeform module ---------------------------------------
@pytest.fixture(scope="module")
def giface():
    return dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_iface.npz")))


def _iface_oracle(oracle, grid):
    """What the oracle computes for the fixture's inputs, keyed like golden_iface.npz."""
    from ifacecases import SCALE, TRANS, iface_case
    V, F, E, moved, raw = iface_case()
    out = {"normalize": oracle.normalize_by_template(raw, SCALE, TRANS),
           "denormalize": oracle.denormalize_by_template(V, SCALE, TRANS),
           "dist_fwd": oracle.distfield_forward(grid, moved), "dist_bwd": oracle.distfield_backward(grid, moved)}
    rest = oracle.store_rigid(V, F)
    out["rigid_fwd"] = oracle.rigid_forward(moved, F, rest); out["rigid_bwd"] = oracle.rigid_backward(moved, F, rest)
    rest = oracle.store_graph(V, E)
    out["graph_fwd"] = oracle.graph_forward(moved, E, rest); out["graph_bwd"] = oracle.graph_backward(moved, E, rest)
    rest, lam = oracle.store_cad(V, F, E)
    out["cad_lambda"] = lam
    out["cad_fwd"] = oracle.cad_forward(moved, F, E, rest, lam); out["cad_bwd"] = oracle.cad_backward(moved, F, E, rest, lam)
    return out


def test_interface_loops_match_reference_golden(oracle, golden, giface):
    """DistanceFieldLoss_*, {Rigid,Graph,Cad}EdgeLoss_*, Store*Information, Normalize/DenormalizeByTemplate: the oracle's
    restatement against the outputs of the reference's own src/interface/*.cc (compiled in place), bit for bit."""
    got = _iface_oracle(oracle, golden["grid"])
    assert set(got) == set(giface)
    for k in giface:
        assert got[k].dtype == giface[k].dtype and _same(got[k], giface[k]), k
    # the fixture reaches the branches that matter: out-of-bounds penalty, dead band, cut-off, non-trivial lambda
    from ifacecases import iface_case
    moved = iface_case()[3]
    N = golden["grid"].shape[0]
    idx = np.trunc(moved * N)
    assert (idx >= N).any() and (idx == N - 1).any() and (moved < 0).any()
    assert (giface["dist_fwd"] == 0).any() and (giface["dist_fwd"] > 0).any()
    assert np.unique(giface["cad_lambda"]).size > 100 and np.abs(giface["cad_bwd"]).max() > 0


def test_interface_loops_match_reference_live(oracle, golden, giface):
    from oracle import ref as R
    if not os.path.exists(os.path.join(R.REFERENCE, "src", "interface", "rigid_layer.cc")):
        pytest.skip("/root/reference is not present here: covered by golden_iface.npz")
    from ifacecases import SCALE, TRANS, iface_case
    R.build()
    V, F, E, moved, raw = iface_case()
    P = R.Params(golden["grid"], SCALE, TRANS)
    assert _same(P.dist_forward(moved), giface["dist_fwd"]) and _same(P.dist_backward(moved), giface["dist_bwd"])
    # a second, differently seeded case straight against the compiled reference (no fixture in between)
    rng = np.random.default_rng(99)
    V2 = (V + rng.normal(0, 0.02, V.shape)).astype(np.float32)
    mv2 = (V2 + rng.normal(0, 3e-3, V.shape)).astype(np.float32)
    grid2 = np.abs(rng.normal(0.1, 0.08, golden["grid"].shape))
    P2 = R.Params(grid2, 0.731, (-0.4, 0.25, 0.05))
    assert _same(P2.dist_forward(mv2), oracle.distfield_forward(grid2, mv2))
    assert _same(P2.dist_backward(mv2), oracle.distfield_backward(grid2, mv2))
    assert _same(P2.normalize(mv2), oracle.normalize_by_template(mv2, 0.731, np.array([-0.4, 0.25, 0.05])))
    assert _same(P2.normalize(mv2, inverse=True), oracle.denormalize_by_template(mv2, 0.731, np.array([-0.4, 0.25, 0.05])))
    P2.rigid_store(V2, F); rest = oracle.store_rigid(V2, F)
    assert _same(P2.rigid_forward(mv2, F), oracle.rigid_forward(mv2, F, rest))
    assert _same(P2.rigid_backward(mv2, F), oracle.rigid_backward(mv2, F, rest))
    P2.graph_store(V2, E); rest = oracle.store_graph(V2, E)
    assert _same(P2.graph_forward(mv2, E), oracle.graph_forward(mv2, E, rest))
    assert _same(P2.graph_backward(mv2, E), oracle.graph_backward(mv2, E, rest))
    P2.cad_store(V2, F, E); rest, lam = oracle.store_cad(V2, F, E)
    assert _same(P2.cad_lambda(lam.shape[0]), lam)
    assert _same(P2.cad_forward(mv2, F, E), oracle.cad_forward(mv2, F, E, rest, lam))
    assert _same(P2.cad_backward(mv2, F, E), oracle.cad_backward(mv2, F, E, rest, lam))
