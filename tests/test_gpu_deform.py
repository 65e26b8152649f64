"""GPU parity of the fused deformation loops against the oracle's restatement of
src/python/rigid_deform.py:32-41 (loss layers + torch.optim.Adam in float32)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _chamfer(A, B):
    from scipy.spatial import cKDTree
    da = cKDTree(B).query(A)[0]
    db = cKDTree(A).query(B)[0]
    return float(da.mean() + db.mean())


@pytest.mark.parametrize("schedule", ["cta", "cluster"])
@pytest.mark.parametrize("n_src,n_tar,N,iters", [(1500, 1200, 32, 300), (5000, 5000, 64, 150)])
def test_persistent_engine_bit_exact(oracle, pd, n_src, n_tar, N, iters, schedule):
    """Both schedules of the exact loop (one CTA per pair; one thread-block cluster per pair with the positions
    exchanged through distributed shared memory) against the CPU loop."""
    from meshode_b200 import engine
    from meshode_b200.synth import synth_pair
    srcV, srcF, tarV, tarF = synth_pair(3, n_src, n_tar)
    batch = engine.PairBatch([(torch.from_numpy(srcV), torch.from_numpy(srcF), torch.from_numpy(tarV), torch.from_numpy(tarF))],
                             grid_resolution=N)
    tmpl = oracle.Template(tarV, tarF, N)
    src_n = oracle.normalize_by_template(srcV, tmpl.scale, tmpl.trans)
    assert np.array_equal(batch.V[0].cpu().numpy(), src_n)
    rest = oracle.store_rigid(src_n, srcF)
    ref, _ = oracle.rigid_adam(tmpl.grid, src_n, srcF, rest, iters, 1e-3)
    batch.deform(iters=iters, lr=1e-3, exact=True, schedule=schedule)
    got = batch.V[0].cpu().numpy()
    # north_star gate: Chamfer <= 1e-4 (normalised units); achieved: identical bits
    assert _chamfer(got, ref) <= 1e-4
    assert np.array_equal(got, ref), "max |dV| = %g" % np.abs(got - ref).max()
    assert np.abs(got - src_n).max() > 1e-3, "the optimisation must actually move the mesh"
    out = batch.finalize()[0].cpu().numpy()
    assert np.array_equal(out, oracle.denormalize_by_template(ref, tmpl.scale, tmpl.trans))
    batch.release()


def test_batch_of_pairs_matches_single(oracle, pd):
    """Pairs in one launch are independent: batch result == per-pair result, bit for bit."""
    from meshode_b200 import engine
    from meshode_b200.synth import synth_pair
    pairs = [tuple(torch.from_numpy(a) for a in synth_pair(i, 700 + 50 * i, 800)) for i in range(5)]
    iters = 120
    b_all = engine.PairBatch(pairs, grid_resolution=32)
    b_all.deform(iters=iters, exact=True)
    for i, p in enumerate(pairs):
        b1 = engine.PairBatch([p], grid_resolution=32)
        b1.deform(iters=iters, exact=True)
        assert torch.equal(b1.V[0], b_all.V[i])
        b1.release()
    # and against the oracle for one of them
    srcV, srcF, tarV, tarF = [a.numpy() for a in pairs[2]]
    tmpl = oracle.Template(tarV, tarF, 32)
    src_n = oracle.normalize_by_template(srcV, tmpl.scale, tmpl.trans)
    ref, _ = oracle.rigid_adam(tmpl.grid, src_n, srcF, oracle.store_rigid(src_n, srcF), iters, 1e-3)
    assert np.array_equal(b_all.V[2].cpu().numpy(), ref)
    b_all.release()


def test_full_length_cfg4_pair_bit_exact(oracle, pd):
    """One pair of BASELINE.json's cfg4 (5 000-vertex source, 9 996-triangle target, grid 64^3) through the FULL
    10 000 Adam iterations of src/python/rigid_deform.py:32-41.  A 1-ulp difference anywhere grows to ~2e-4
    Chamfer over this length (tools/chaos_probe.py), above the 1e-4 gate of north_star, so the gate is only
    meaningful at full length -- and only bit-identical arithmetic passes it."""
    from meshode_b200 import engine
    from meshode_b200.synth import synth_pair
    iters = 10000
    srcV, srcF, tarV, tarF = synth_pair(17, 5000, 5000)
    P = tuple(torch.from_numpy(a) for a in (srcV, srcF, tarV, tarF))
    tmpl = oracle.Template(tarV, tarF, 64)
    src_n = oracle.normalize_by_template(srcV, tmpl.scale, tmpl.trans)
    ref, _ = oracle.rigid_adam(tmpl.grid, src_n, srcF, oracle.store_rigid(src_n, srcF), iters, 1e-3)
    for schedule in ("cta", "cluster"):
        batch = engine.PairBatch([P], grid_resolution=64)
        batch.deform(iters=iters, lr=1e-3, exact=True, schedule=schedule)
        got = batch.V[0].cpu().numpy()
        assert _chamfer(got, ref) <= 1e-4, schedule
        assert np.array_equal(got, ref), "%s: max |dV| = %g" % (schedule, np.abs(got - ref).max())
        batch.release()
    assert np.abs(ref - src_n).max() > 1e-2


def test_partial_wave_goes_to_clusters_same_bits(pd):
    """A batch that is not a multiple of the SM count: the automatic schedule runs the full waves one CTA per
    pair and the remainder on clusters; every pair's result equals the CTA-only schedule bit for bit."""
    from meshode_b200 import engine
    from meshode_b200.synth import synth_pair
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    base = [tuple(torch.from_numpy(a).cuda() for a in synth_pair(i, 1100 + 37 * i, 600)) for i in range(6)]
    pairs = [base[i % 6] for i in range(sms + 5)]
    a = engine.PairBatch(pairs, grid_resolution=24)
    b = engine.PairBatch(pairs, grid_resolution=24)
    a.deform(iters=60, exact=True, schedule="auto")
    b.deform(iters=60, exact=True, schedule="cta")
    for x, y in zip(a.V, b.V):
        assert torch.equal(x, y)
    # repeated inputs give repeated outputs whichever kernel took them
    for i in range(6, len(pairs)):
        assert torch.equal(a.V[i], a.V[i % 6])
    a.release(); b.release()


def test_large_mesh_loop_cfg1_long(meshes, oracle, pd):
    """cfg1's 21 542-vertex source through 2 000 iterations of the cooperative loop (mo_deform_adam_large)."""
    from meshode_b200 import engine
    N, iters = 32, 2000
    p = [torch.from_numpy(meshes[k]) for k in ("srcV", "srcF", "tarV", "tarF")]
    batch = engine.PairBatch([tuple(p)], grid_resolution=N)
    batch.deform(iters=iters, exact=True)
    tmpl = oracle.Template(meshes["tarV"], meshes["tarF"], N)
    src_n = oracle.normalize_by_template(meshes["srcV"], tmpl.scale, tmpl.trans)
    ref, _ = oracle.rigid_adam(tmpl.grid, src_n, meshes["srcF"], oracle.store_rigid(src_n, meshes["srcF"]), iters, 1e-3)
    got = batch.V[0].cpu().numpy()
    assert _chamfer(got, ref) <= 1e-4
    assert np.array_equal(got, ref), "max |dV| = %g" % np.abs(got - ref).max()
    batch.release()


def test_large_mesh_loop_cfg1(meshes, oracle, pd):
    """data/source.obj -> data/target.obj (21 542 vertices: HBM-resident loop, two launches per iteration)."""
    from meshode_b200 import engine
    N, iters = 32, 25
    p = [torch.from_numpy(meshes[k]) for k in ("srcV", "srcF", "tarV", "tarF")]
    batch = engine.PairBatch([tuple(p)], grid_resolution=N)
    batch.deform(iters=iters, exact=True)
    tmpl = oracle.Template(meshes["tarV"], meshes["tarF"], N)
    src_n = oracle.normalize_by_template(meshes["srcV"], tmpl.scale, tmpl.trans)
    ref, _ = oracle.rigid_adam(tmpl.grid, src_n, meshes["srcF"], oracle.store_rigid(src_n, meshes["srcF"]), iters, 1e-3)
    got = batch.V[0].cpu().numpy()
    assert _chamfer(got, ref) <= 1e-4
    assert np.array_equal(got, ref), "max |dV| = %g" % np.abs(got - ref).max()
    batch.release()


@pytest.mark.parametrize("n_src,n_tar,N,iters", [(1500, 1200, 32, 300), (5000, 5000, 64, 400)])
def test_fast_engine_is_at_the_rounding_sensitivity_floor(oracle, pd, n_src, n_tar, N, iters):
    """The opt-in fast loop (exact=False) sums the edge term over distinct neighbours on displacements:
    mathematically the reference's sum, not its float32 order.  Adam's normalised steps make the
    trajectory chaotic at the 1e-4 level -- moving every start coordinate by ONE ulp changes the exact
    loop's result by the same amount (tools/chaos_probe.py: Chamfer ~1.9e-4 after 10 000 iterations) --
    so the fast loop cannot meet the 1e-4 Chamfer gate against the CPU and is NOT the default; this
    test pins it to that sensitivity floor: no further from the oracle than the oracle is from itself
    under a 1-ulp perturbation (x3 margin)."""
    from meshode_b200 import engine
    from meshode_b200.synth import synth_pair
    srcV, srcF, tarV, tarF = synth_pair(3, n_src, n_tar)
    P = (torch.from_numpy(srcV), torch.from_numpy(srcF), torch.from_numpy(tarV), torch.from_numpy(tarF))
    tmpl = oracle.Template(tarV, tarF, N)
    src_n = oracle.normalize_by_template(srcV, tmpl.scale, tmpl.trans)
    rest = oracle.store_rigid(src_n, srcF)
    ref, _ = oracle.rigid_adam(tmpl.grid, src_n, srcF, rest, iters, 1e-3)
    bumped, _ = oracle.rigid_adam(tmpl.grid, np.nextafter(src_n, np.float32(2.0)), srcF, rest, iters, 1e-3)
    floor = _chamfer(bumped, ref)
    fast = engine.PairBatch([P], grid_resolution=N)
    fast.deform(iters=iters, lr=1e-3, exact=False)
    got = fast.V[0].cpu().numpy()
    assert np.abs(ref - src_n).max() > 1e-3
    cham = _chamfer(got, ref)
    print("fast vs oracle after %d iterations: chamfer %.3g (1-ulp sensitivity floor %.3g), max |dV| %.3g" %
          (iters, cham, floor, np.abs(got - ref).max()))
    assert cham <= max(3.0 * floor, 2e-5)
    assert cham <= 5e-4
    fast.release()


def test_fast_and_exact_agree_over_a_few_iterations(oracle, pd):
    """Before rounding differences are amplified the two loops coincide to float32 resolution."""
    from meshode_b200 import engine
    from meshode_b200.synth import synth_pair
    srcV, srcF, tarV, tarF = synth_pair(9, 3000, 2500)
    P = (torch.from_numpy(srcV), torch.from_numpy(srcF), torch.from_numpy(tarV), torch.from_numpy(tarF))
    a = engine.PairBatch([P], grid_resolution=48)
    b = engine.PairBatch([P], grid_resolution=48)
    a.deform(iters=3, exact=True)
    b.deform(iters=3, exact=False)
    d = (a.V[0] - b.V[0]).abs()
    # Adam's first steps are +-lr per coordinate; a gradient at the sign threshold may flip one of them
    assert (d <= 1e-6).float().mean().item() > 0.995 and d.max().item() <= 6.1e-3
    a.release(); b.release()


@pytest.mark.parametrize("schedule", ["cta", "cluster"])
@pytest.mark.parametrize("fan", [1, 9, 26])
def test_hub_vertices_take_the_long_adjacency_rows(oracle, pd, fan, schedule):
    """A vertex with many incident edges: 2*fan + its own ~12 incidences exceed the eight adjacency words that the loop
    reads with vector loads (fan = 1: exactly eight words; 9: eight more rows; 26: many), so the remaining rows come from
    the row-major table.  Same bits as the CPU loop on both schedules, in a batch next to an ordinary pair."""
    from meshode_b200 import engine
    from meshode_b200.synth import synth_pair
    srcV, srcF, tarV, tarF = synth_pair(31, 1100, 900)
    rng = np.random.default_rng(fan)
    hub, ring = 17, rng.choice(np.arange(100, 1100), size=max(fan, 2), replace=False)
    extra = np.stack([np.full(ring.size, hub), ring, np.roll(ring, 1)], 1).astype(np.int32)[:fan]   # a fan of triangles around the hub
    srcF2 = np.ascontiguousarray(np.concatenate([srcF[:900], extra, srcF[900:]]))        # in the middle of the edge order
    deg = np.bincount(srcF2.reshape(-1), minlength=srcV.shape[0])
    assert 2 * deg[hub] == {1: 16, 9: 32, 26: 66}[fan]   # exactly eight words; one table row beyond them; many
    other = synth_pair(32, 1000, 900)
    pairs = [(srcV, srcF2, tarV, tarF), other]
    batch = engine.PairBatch([tuple(torch.from_numpy(a) for a in p) for p in pairs], grid_resolution=32)
    iters = 80
    batch.deform(iters=iters, lr=1e-3, exact=True, schedule=schedule)
    for k, (sV, sF, tV, tF) in enumerate(pairs):
        tmpl = oracle.Template(tV, tF, 32)
        src_n = oracle.normalize_by_template(sV, tmpl.scale, tmpl.trans)
        ref, _ = oracle.rigid_adam(tmpl.grid, src_n, sF, oracle.store_rigid(src_n, sF), iters, 1e-3)
        got = batch.V[k].cpu().numpy()
        assert np.array_equal(got, ref), "pair %d: max |dV| = %g" % (k, np.abs(got - ref).max())
    batch.release()


def _grid_mesh(m, n, closed):
    """m x n vertices on a torus (closed: every vertex has 12 incident edge slots) or on a bent open patch (boundary
    vertices have as few as 4, so the adjacency rows are padded from the third word on)."""
    u, v = np.meshgrid(np.arange(m), np.arange(n), indexing="ij")
    if closed:
        a, b = 2 * np.pi * u / m, 2 * np.pi * v / n
        V = np.stack([(0.62 + 0.2 * np.cos(b)) * np.cos(a), (0.62 + 0.2 * np.cos(b)) * np.sin(a), 0.2 * np.sin(b)], -1)
    else:
        x, y = u / (m - 1) - 0.5, v / (n - 1) - 0.5
        V = np.stack([1.2 * x, 1.2 * y, 0.35 * np.cos(2.2 * x) * np.cos(1.7 * y) - 0.2], -1)
    idx = lambda i, j: (i % m) * n + (j % n)  # noqa: E731
    F = []
    for i in range(m if closed else m - 1):
        for j in range(n if closed else n - 1):
            F.append((idx(i, j), idx(i + 1, j), idx(i + 1, j + 1)))
            F.append((idx(i, j), idx(i + 1, j + 1), idx(i, j + 1)))
    return np.ascontiguousarray(V.reshape(-1, 3), dtype=np.float32), np.asarray(F, dtype=np.int32)


@pytest.mark.parametrize("schedule", ["cta", "cluster"])
@pytest.mark.parametrize("closed", [True, False])
def test_regular_and_open_meshes_six_word_rows(oracle, pd, closed, schedule):
    """Sources whose vertices have at most 12 incident edges take the six-word instantiation of the loops; an open patch
    adds rows that are padded early (the vertex itself: an exact zero term).  Same bits as the CPU loop."""
    from meshode_b200 import engine
    from meshode_b200.synth import synth_mesh
    srcV, srcF = _grid_mesh(40, 31, closed)
    assert np.bincount(srcF.reshape(-1)).max() == 6 and (np.bincount(srcF.reshape(-1)).min() == 6) == closed
    tarV, tarF = synth_mesh(900, 5)
    batch = engine.PairBatch([tuple(torch.from_numpy(a) for a in (srcV, srcF, tarV, tarF))], grid_resolution=32)
    iters = 90
    batch.deform(iters=iters, lr=1e-3, exact=True, schedule=schedule)
    tmpl = oracle.Template(tarV, tarF, 32)
    src_n = oracle.normalize_by_template(srcV, tmpl.scale, tmpl.trans)
    ref, _ = oracle.rigid_adam(tmpl.grid, src_n, srcF, oracle.store_rigid(src_n, srcF), iters, 1e-3)
    got = batch.V[0].cpu().numpy()
    assert np.array_equal(got, ref), "max |dV| = %g" % np.abs(got - ref).max()
    batch.release()
