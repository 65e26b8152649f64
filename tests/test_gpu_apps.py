"""The re-hosted C++ drivers (apps/, reference src/app/rigid_deform.cc and rigid_rot_deform.cc) end to end on
the GPU: OBJ in, OBJ out, same console lines as the reference binaries."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_obj(path, V, F):
    with open(path, "w") as fh:
        for v in V:
            fh.write("v %.9g %.9g %.9g\n" % tuple(float(x) for x in v))
        for f in F:
            fh.write("f %d/%d/%d %d %d//%d\n" % (f[0] + 1, f[0] + 1, f[0] + 1, f[1] + 1, f[2] + 1, f[2] + 1))


def _read_obj(path):
    V, F = [], []
    for ln in open(path):
        t = ln.split()
        if t and t[0] == "v":
            V.append([float(x) for x in t[1:4]])
        elif t and t[0] == "f":
            F.append([int(x.split("/")[0]) - 1 for x in t[1:4]])
    return np.array(V), np.array(F)


def _exe(name):
    from meshode_b200 import build
    exes = dict(zip(build.APPS, build.build_apps()))
    return exes[name]


def _run(name, args):
    p = subprocess.run([_exe(name)] + [str(a) for a in args], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    return p.stdout


def _costs(out):
    m = {k: float(re.search(k + r": ([-+0-9.eE]+)", out).group(1)) for k in ("Vertices cost", "Rigidity cost", "Final cost")}
    first = float(re.search(r"^\s*0\s+([-+0-9.eE]+)", out, re.M).group(1))
    return m, first


@pytest.mark.parametrize("name,kind_name", [("rigid_deform", "EDGE"), ("rigid_rot_deform", "ROT_EDGE")])
def test_driver_matches_the_library_path(tmp_path, oracle, pd, name, kind_name):
    from meshode_b200 import capi
    from meshode_b200.synth import synth_pair
    srcV, srcF, tarV, tarF = synth_pair(11, 300, 500)
    s_obj, t_obj, o_obj = (str(tmp_path / x) for x in ("s.obj", "t.obj", "o.obj"))
    _write_obj(s_obj, srcV, srcF); _write_obj(t_obj, tarV, tarF)
    lam = 0.7
    out = _run(name, [s_obj, t_obj, o_obj, 32, 5000, lam])
    assert "Source:\t\tNum vertices: 300\tNum faces: 596" in out and "Deformed" in out
    costs, first = _costs(out)
    assert costs["Final cost"] < first
    oV, oF = _read_obj(o_obj)
    assert np.array_equal(oF, srcF)
    # the same problem through the Python binding, built the way the driver builds it (FP64 from the OBJ text)
    tv = tarV.astype(np.float64)
    mn, mx = tv.min(0), tv.max(0)
    scale = (mx - mn).max() * 1.1; pos = mn - 0.05 * scale
    tm = oracle.Template(tarV, tarF, 32)                      # float32 -> FP64 normalisation: the same numbers
    assert abs(tm.scale - scale) <= 1e-15 * scale
    V0 = (srcV.astype(np.float64) - pos) / scale
    a = srcF.reshape(-1); b = np.roll(srcF, -1, axis=1).reshape(-1)
    I = np.stack([a, b], 1).astype(np.int32); rest = V0[a] - V0[b]
    pid = pd.InitializeDeformTemplate(torch.from_numpy(tarV).cuda(), torch.from_numpy(tarF).cuda(), 0, 32)
    V = torch.from_numpy(V0.copy()).cuda(); R = torch.zeros_like(V)
    kind = getattr(capi, "CERES_" + kind_name)
    s = pd.CeresSolve(pid, kind, V, R if kind_name == "ROT_EDGE" else None, torch.from_numpy(I).cuda(),
                      torch.from_numpy(rest).cuda(), lam)
    assert abs(s["final_cost"] - costs["Final cost"]) <= 2e-5 * s["final_cost"]      # 6 significant digits on the console
    want = V.cpu().numpy() * scale + pos                                              # Mesh::WriteOBJ denormalises
    assert np.abs(oV - want).max() <= 2e-5 * np.abs(want).max()
    pd.DestroyTemplate(pid)


def test_rigid_deform_cfg1(tmp_path, meshes):
    """cfg1 of BASELINE.json: rigid_deform data/source.obj -> data/target.obj, GRID_RESOLUTION=64, lambda=1."""
    s_obj, t_obj, o_obj = (str(tmp_path / x) for x in ("source.obj", "target.obj", "out.obj"))
    _write_obj(s_obj, meshes["srcV"], meshes["srcF"]); _write_obj(t_obj, meshes["tarV"], meshes["tarF"])
    out = _run("rigid_deform", [s_obj, t_obj, o_obj, 64, 5000, 1])
    print(out[-1500:])
    costs, first = _costs(out)
    assert costs["Final cost"] < 0.5 * first
    oV, oF = _read_obj(o_obj)
    assert oV.shape == meshes["srcV"].shape and np.array_equal(oF, meshes["srcF"]) and np.isfinite(oV).all()


def test_usage_without_arguments():
    p = subprocess.run([_exe("rigid_rot_deform")], capture_output=True, text=True, timeout=60)
    assert p.returncode == 0 and "rigid_rot_deform source.obj reference.obj output.obj" in p.stdout


def test_python_scripts_mirror_the_reference_cli(tmp_path):
    """scripts/rigid_deform.py (reference src/python/rigid_deform.py: --source --target --output) with both engines,
    and scripts/batch_deform.py on a two-line file list."""
    import sys
    from meshode_b200.synth import synth_pair
    outs = {}
    srcV, srcF, tarV, tarF = synth_pair(13, 400, 500)
    s_obj, t_obj = str(tmp_path / "s.obj"), str(tmp_path / "t.obj")
    _write_obj(s_obj, srcV, srcF); _write_obj(t_obj, tarV, tarF)
    for eng in ("fused", "layers"):
        o = str(tmp_path / ("o_%s.obj" % eng))
        p = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "rigid_deform.py"), "--source", s_obj, "--target", t_obj,
                            "--output", o, "--engine", eng, "--niter", "60", "--grid", "32"], capture_output=True, text=True,
                           timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        if eng == "layers":
            assert "iter=0 loss=" in p.stdout
        outs[eng] = _read_obj(o)
    assert np.array_equal(outs["fused"][1], srcF) and np.array_equal(outs["layers"][1], srcF)
    # same loss, same optimiser: the two engines agree to the OBJ's six significant digits after 60 steps
    assert np.abs(outs["fused"][0] - outs["layers"][0]).max() <= 2e-5 * np.abs(outs["fused"][0]).max()
    assert np.abs(outs["fused"][0] - srcV).max() > 1e-3          # and the mesh did move
    lst = tmp_path / "pairs.txt"
    lst.write_text("%s %s %s\n%s %s %s\n" % (s_obj, t_obj, tmp_path / "b0.obj", t_obj, s_obj, tmp_path / "b1.obj"))
    p = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "batch_deform.py"), "--filelist", str(lst), "--niter", "60",
                        "--grid", "32"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "2 pairs" in p.stdout, p.stderr[-2000:]
    b0 = _read_obj(str(tmp_path / "b0.obj"))
    assert np.array_equal(b0[0], outs["fused"][0])               # a pair deforms the same alone or in a batch
    assert _read_obj(str(tmp_path / "b1.obj"))[0].shape == tarV.shape


def test_cad_deform2_cfg2(tmp_path, meshes):
    """cfg2 of BASELINE.json: cad_deform2.py data/cad-source.obj -> data/cad-target.obj, rigidity 1, grid 64
    (a few hundred of the 10 000 iterations): LoadCadMesh, graph + reverse losses on the GPU, SolveLinear, SaveMesh."""
    import sys
    s_obj, t_obj, o_obj = (str(tmp_path / x) for x in ("cad-source.obj", "cad-target.obj", "cad-output.obj"))
    _write_obj(s_obj, meshes["cadSrcV"], meshes["cadSrcF"]); _write_obj(t_obj, meshes["cadTarV"], meshes["cadTarF"])
    p = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "cad_deform2.py"), "--source", s_obj, "--target", t_obj,
                        "--output", o_obj, "--rigidity", "1", "--niter", "301"], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-3000:]
    print(p.stdout[-600:])
    vals = [(float(a), float(b)) for a, b in re.findall(r"loss_src2tar=([0-9.eE+-]+) loss_tar2src=([0-9.eE+-]+)", p.stdout)]
    assert len(vals) >= 3 and vals[-1][0] < vals[0][0] and vals[-1][1] < vals[0][1]      # both directions improve
    oV, oF = _read_obj(o_obj)
    assert oV.shape[0] > meshes["cadSrcV"].shape[0] and np.isfinite(oV).all() and oF.max() < oV.shape[0]


def test_cad_neural_deform2(tmp_path, meshes):
    """The NeuralODE script of cfg5 (cad_neural_deform2.py) on the shipped CAD pair, a few iterations."""
    import sys
    s_obj, t_obj, o_obj = (str(tmp_path / x) for x in ("cad-source.obj", "cad-target.obj", "cad-output.obj"))
    _write_obj(s_obj, meshes["cadSrcV"], meshes["cadSrcF"]); _write_obj(t_obj, meshes["cadTarV"], meshes["cadTarF"])
    p = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "cad_neural_deform2.py"), "--source", s_obj, "--target", t_obj,
                        "--output", o_obj, "--niter", "12", "--save_path", str(tmp_path / "flow.ckpt")], capture_output=True,
                       text=True, timeout=900, env=dict(os.environ, MESHODE_SEED="7"))
    assert p.returncode == 0, p.stderr[-3000:]
    tot = [sum(float(x) for x in m) for m in re.findall(
        r"loss1_forward=([0-9.eE+-]+) loss1_backward=([0-9.eE+-]+) loss2_forward=([0-9.eE+-]+) loss2_backward=([0-9.eE+-]+)", p.stdout)]
    assert len(tot) == 12 and np.isfinite(tot).all() and tot[-1] < tot[0]
    oV, oF = _read_obj(o_obj)
    assert np.isfinite(oV).all() and oF.max() < oV.shape[0] and os.path.exists(str(tmp_path / "flow.ckpt"))
    # the checkpoint has the reference's layout ({'func': NeuralODE object, 'optim': optimizer}, cad_neural_deform2.py:108)
    # and resumes: the first loss line of a resumed run continues from the trained flow, not from a fresh one
    sys.path.insert(0, ROOT)
    ck = torch.load(str(tmp_path / "flow.ckpt"), map_location="cpu", weights_only=False)
    assert type(ck["func"]).__name__ == "NeuralODE" and isinstance(ck["optim"], torch.optim.Adam)
    p2 = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "cad_neural_deform2.py"), "--source", s_obj, "--target", t_obj,
                         "--output", o_obj, "--niter", "2", "--save_path", "", "--resume_path", str(tmp_path / "flow.ckpt")],
                        capture_output=True, text=True, timeout=900)
    assert p2.returncode == 0, p2.stderr[-3000:]
    tot2 = [sum(float(x) for x in m) for m in re.findall(
        r"loss1_forward=([0-9.eE+-]+) loss1_backward=([0-9.eE+-]+) loss2_forward=([0-9.eE+-]+) loss2_backward=([0-9.eE+-]+)", p2.stdout)]
    # (Adam's first dozen steps from a random flow are not monotone: the resumed loss is compared with the trained
    #  level loosely and with the untrained one strictly)
    assert len(tot2) == 2 and tot2[0] < tot[0] and tot2[0] <= 1.25 * tot[-1]


def test_cad_deform_driver_cfg2_pair(tmp_path, meshes):
    """The re-hosted cad_deform (reference src/app/cad_deform.cc:21-120) on the shipped CAD pair: host-side clean-up +
    subdivision + deformation graph, the reference's distance field and Deformer::DeformGraph on the GPU through the
    C-ABI, host-side LinearSolve; same command line and console lines as the reference binary."""
    import sys
    s_obj, t_obj, o_obj = (str(tmp_path / x) for x in ("cad-source.obj", "cad-target.obj", "cad-output.obj"))
    _write_obj(s_obj, meshes["cadSrcV"], meshes["cadSrcF"]); _write_obj(t_obj, meshes["cadTarV"], meshes["cadTarF"])
    env = dict(os.environ, MESHODE_PYTHON=sys.executable)
    p = subprocess.run([_exe("cad_deform"), s_obj, t_obj, o_obj, "64", "5000", "1"], capture_output=True, text=True, timeout=900,
                       env=env)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    out = p.stdout
    assert re.search(r"Source:\t\tNum vertices: \d+\tNum faces: \d+", out) and "Reference:\t" in out and out.rstrip().endswith("Deformed")
    m, first = _costs(out)
    assert abs(m["Final cost"] - (m["Vertices cost"] + m["Rigidity cost"])) <= 1e-9 * max(m["Final cost"], 1e-30)
    assert m["Final cost"] < first            # DeformGraph lowered the cost of its problem
    oV, oF = _read_obj(o_obj)
    nsub = int(re.search(r"Source:\t\tNum vertices: (\d+)", out).group(1))
    assert oV.shape[0] == nsub > meshes["cadSrcV"].shape[0] and np.isfinite(oV).all() and oF.max() < oV.shape[0]
    # the deformed CAD model moved towards the target: mean distance-field value of its vertices decreased
    from meshode_b200 import pyDeform as pd
    from meshode_b200 import cadmesh
    tV, tF = torch.from_numpy(meshes["cadTarV"]).cuda(), torch.from_numpy(meshes["cadTarF"]).cuda()
    pid = pd.InitializeDeformTemplate(tV, tF, 0, 64)
    def mean_dist(V):
        v = torch.from_numpy(np.ascontiguousarray(V, dtype=np.float32)).cuda()
        pd.NormalizeByTemplate(v, pid)
        return float(pd.DistanceFieldLoss_forward(v, pid).sqrt().mean())
    from meshode_b200.objio import read_obj
    V0, F0 = read_obj(s_obj, vertex_dtype=np.float64)      # what the driver's host helper read (cad_host.prepare)
    V0, F0 = np.asarray(V0, dtype=np.float64), np.asarray(F0, dtype=np.int64).reshape(-1, 3)
    F0 = cadmesh.remove_degenerated(V0, F0)
    V0, F0 = cadmesh.merge_duplex(V0, F0)
    V0, F0 = cadmesh.subdivide(V0, F0, 2e-2)
    assert V0.shape[0] == nsub
    assert mean_dist(oV) < 0.8 * mean_dist(V0)
    pd.DestroyTemplate(pid)
    # usage line when called without arguments, exit code 0 (cad_deform.cc:22-27)
    u = subprocess.run([_exe("cad_deform")], capture_output=True, text=True, timeout=60)
    assert u.returncode == 0 and u.stdout.startswith("./cad_deform cad.obj reference.obj output.obj")
