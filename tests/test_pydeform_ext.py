"""The compiled ``pyDeform`` module (meshode_b200/csrc/pydeform_ext.cpp -> meshode_b200/ext/pyDeform.<abi>.so): the
torch C++ extension a maintainer of the reference would build instead of src/interface/pydeform.cc.

CPU part: it builds, needs ``import torch`` first like the reference's module (README.md:46-50), exports the
reference's 18 names, validates its arguments and refuses to compute without a GPU.
GPU part: same bits as the ctypes module on CUDA and on CPU tensors; and -- where the reference's own Python tree is
available (``MESHODE_REFERENCE_PY`` or /root/reference/src/python; it is not on the driver's GPU box, see
tools/run_reference_scripts.sh for how the run is made there) -- the reference's UNMODIFIED src/python/rigid_deform.py
with the reference's own layers/ on top of this module, checked against the oracle's gradient driven by the very
same torch.optim.Adam."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE_NAMES = ["LoadMesh", "LoadCadMesh", "SaveMesh", "InitializeDeformTemplate", "NormalizeByTemplate",
                   "DenormalizeByTemplate", "SolveLinear", "DistanceFieldLoss_forward", "DistanceFieldLoss_backward",
                   "RigidEdgeLoss_forward", "RigidEdgeLoss_backward", "StoreRigidityInformation", "CadEdgeLoss_forward",
                   "CadEdgeLoss_backward", "StoreCadInformation", "GraphEdgeLoss_forward", "GraphEdgeLoss_backward",
                   "StoreGraphInformation"]   # src/interface/pydeform.cc:15-38


@pytest.fixture(scope="module")
def ext_dir():
    from meshode_b200 import build
    path = build.build_ext()
    assert os.path.exists(path)
    return os.path.dirname(path)


def _py(ext_dir, code, timeout=300):
    env = dict(os.environ, PYTHONPATH=ext_dir + os.pathsep + ROOT)
    return subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=timeout, env=env, cwd="/tmp")


def test_extension_surface_and_import_order(ext_dir, tmp_path, meshes):
    p = _py(ext_dir, "import torch, pyDeform; print(pyDeform.__file__); print(' '.join(sorted(n for n in dir(pyDeform) "
                     "if not n.startswith('_'))))")
    assert p.returncode == 0, p.stderr[-2000:]
    path, names = p.stdout.strip().split("\n")
    assert path.startswith(ext_dir) and path.endswith(".so")           # the compiled module, not the ctypes shim
    assert set(REFERENCE_NAMES) <= set(names.split())
    # the reference's rule: torch first (the module links libtorch and finds it only once torch is loaded)
    q = _py(ext_dir, "import pyDeform")
    assert q.returncode != 0 and "libtorch" in q.stderr or "libc10" in q.stderr
    # it is bound to the C-ABI library next to it
    ldd = subprocess.run(["ldd", path], capture_output=True, text=True).stdout
    assert re.search(r"libmeshode_b200\.so => .*meshode_b200/", ldd)
    # host side: OBJ semantics of the reference reader / writer
    obj = tmp_path / "m.obj"
    V, F = meshes["srcV"][:50], meshes["srcF"][meshes["srcF"].max(1) < 50][:40]
    obj.write_text("# c\n" + "".join("v %.9g %.9g %.9g\n" % tuple(v) for v in V) + "vn 0 0 1\n" +
                   "".join("f %d/1/1 %d//2 %d\n" % tuple(f + 1) for f in F))
    code = ("import torch, pyDeform, numpy as np\n"
            "V, F = pyDeform.LoadMesh(%r)\n"
            "assert V.dtype == torch.float32 and F.dtype == torch.int32 and V.shape == (50, 3)\n"
            "pyDeform.SaveMesh(%r, V, F)\n"
            "V2, F2 = pyDeform.LoadMesh(%r)\n"
            "assert torch.equal(F, F2) and (V - V2).abs().max() <= 1e-5 * V.abs().max()\n"
            "np.save(%r, V.numpy()); np.save(%r, F.numpy())\n") % (str(obj), str(tmp_path / "o.obj"), str(tmp_path / "o.obj"),
                                                                 str(tmp_path / "V.npy"), str(tmp_path / "F.npy"))
    r = _py(ext_dir, code)
    assert r.returncode == 0, r.stderr[-2000:]
    assert np.array_equal(np.load(tmp_path / "V.npy"), V.astype(np.float32)) and np.array_equal(np.load(tmp_path / "F.npy"), F)


def test_extension_validates_and_has_no_cpu_path(ext_dir):
    code = """
import torch, pyDeform
V = torch.rand(10, 3); F = torch.zeros(4, 3, dtype=torch.int32)
def raises(fn, *a):
    try:
        fn(*a)
    except (RuntimeError, TypeError) as e:
        return str(e)
    raise SystemExit("no exception from %s" % fn.__name__)
assert "dtype" in raises(pyDeform.DistanceFieldLoss_forward, V.double(), 0)
assert "shape" in raises(pyDeform.DistanceFieldLoss_forward, torch.rand(10, 4), 0)
assert "contiguous" in raises(pyDeform.NormalizeByTemplate, torch.rand(3, 10).t(), 0)
assert "dtype" in raises(pyDeform.StoreRigidityInformation, V, F.long(), 0)
if not torch.cuda.is_available():
    for fn, a in [(pyDeform.InitializeDeformTemplate, (V, F, 0, 16)), (pyDeform.DistanceFieldLoss_backward, (V, 0)),
                  (pyDeform.RigidEdgeLoss_forward, (V, F, 0)), (pyDeform.NormalizeByTemplate, (V, 0))]:
        assert "no CUDA device" in raises(fn, *a), fn.__name__
print("ok")
"""
    p = _py(ext_dir, code)
    assert p.returncode == 0 and p.stdout.strip() == "ok", p.stdout + p.stderr[-2000:]


@pytest.mark.gpu
def test_extension_matches_the_ctypes_module_bit_for_bit(ext_dir, tmp_path):
    """All 14 hot functions, CUDA tensors and CPU tensors, against meshode_b200.pyDeform (which the other GPU tests
    pin to the oracle)."""
    code = """
import sys, torch, pyDeform as X
assert X.__file__.endswith(".so")
from meshode_b200 import pyDeform as P
from meshode_b200.synth import synth_pair
from meshode_b200.synth import unique_edges
import numpy as np
sV, sF, tV, tF = [torch.from_numpy(a) for a in synth_pair(5, 900, 700)]
E = torch.from_numpy(unique_edges(sF.numpy()).astype(np.int32))
for dev in ("cuda", "cpu"):
    a = [t.to(dev) for t in (sV, sF, tV, tF, E)]
    b = [t.clone() for t in a]
    px, pp = X.InitializeDeformTemplate(a[2], a[3], 0, 32), P.InitializeDeformTemplate(b[2], b[3], 0, 32)
    X.NormalizeByTemplate(a[0], px); P.NormalizeByTemplate(b[0], pp)
    assert torch.equal(a[0], b[0]) and a[0].device.type == dev
    moved = (a[0] + 2e-3 * torch.sin(41.0 * a[0])).contiguous()
    assert torch.equal(X.DistanceFieldLoss_forward(moved, px), P.DistanceFieldLoss_forward(moved, pp))
    assert torch.equal(X.DistanceFieldLoss_backward(moved, px), P.DistanceFieldLoss_backward(moved, pp))
    X.StoreRigidityInformation(a[0], a[1], px); P.StoreRigidityInformation(b[0], b[1], pp)
    fx = X.RigidEdgeLoss_forward(moved, a[1], px)
    assert fx.shape == (3 * sF.shape[0], 3) and fx.device.type == dev
    assert torch.equal(fx, P.RigidEdgeLoss_forward(moved, b[1], pp))
    assert torch.equal(X.RigidEdgeLoss_backward(moved, a[1], px), P.RigidEdgeLoss_backward(moved, b[1], pp))
    lx, gx = X.LossForwardBackward(moved, px, px, 1.0, 0.0)
    lp, gp = P.LossForwardBackward(moved, pp, pp, 1.0, 0.0)
    assert torch.equal(lx, lp) and torch.equal(gx, gp)
    X.StoreGraphInformation(a[0], a[4], px); P.StoreGraphInformation(b[0], b[4], pp)
    assert torch.equal(X.GraphEdgeLoss_forward(moved, a[4], px), P.GraphEdgeLoss_forward(moved, b[4], pp))
    assert torch.equal(X.GraphEdgeLoss_backward(moved, a[4], px), P.GraphEdgeLoss_backward(moved, b[4], pp))
    X.StoreCadInformation(a[0], a[1], a[4], px); P.StoreCadInformation(b[0], b[1], b[4], pp)
    assert torch.equal(X.CadEdgeLoss_forward(moved, a[1], a[4], px), P.CadEdgeLoss_forward(moved, b[1], b[4], pp))
    assert torch.equal(X.CadEdgeLoss_backward(moved, a[1], a[4], px), P.CadEdgeLoss_backward(moved, b[1], b[4], pp))
    X.DenormalizeByTemplate(a[0], px); P.DenormalizeByTemplate(b[0], pp)
    assert torch.equal(a[0], b[0])
    if dev == "cuda":
        X.NormalizeByTemplate(a[0], px); P.NormalizeByTemplate(b[0], pp)
        X.StoreRigidityInformation(a[0], a[1], px); P.StoreRigidityInformation(b[0], b[1], pp)
        X.DeformBatchAdam([a[0]], [px], 50, 1e-3)
        from meshode_b200 import engine
        engine.deform_batch_adam([b[0]], [pp], [pp], 50, 1e-3)
        assert torch.equal(a[0], b[0])
    try:
        X.DistanceFieldLoss_forward(moved, 12345)
        raise SystemExit("bad param_id accepted")
    except RuntimeError:
        pass
    X.DestroyTemplate(px); P.DestroyTemplate(pp)
torch.cuda.synchronize()
print("ok")
"""
    p = _py(ext_dir, code, timeout=600)
    assert p.returncode == 0 and p.stdout.strip().endswith("ok"), p.stdout[-2000:] + p.stderr[-3000:]


def _reference_python():
    for d in (os.environ.get("MESHODE_REFERENCE_PY"), "/root/reference/src/python"):
        if d and os.path.isfile(os.path.join(d, "rigid_deform.py")) and os.path.isfile(os.path.join(d, "layers", "rigid_loss_layer.py")):
            return d
    return None


@pytest.mark.gpu
@pytest.mark.parametrize("module", ["compiled", "ctypes"])
def test_reference_rigid_deform_script_unmodified(ext_dir, tmp_path, oracle, module):
    """src/python/rigid_deform.py:25-44 exactly as the reference ships it -- its own argparse, its own
    layers/rigid_loss_layer.py, 10 000 Adam iterations on CPU tensors -- with ``import pyDeform`` resolving to this
    repository's module.  Expected result: the same loop with the ORACLE's gradient (distfield_backward +
    rigid_backward) under the same torch.optim.Adam, which must agree to the last bit on every iteration, so the
    saved OBJ and every printed loss line are compared as text."""
    ref_py = _reference_python()
    if ref_py is None:
        pytest.skip("the reference's src/python is not available here (set MESHODE_REFERENCE_PY)")
    from meshode_b200.objio import read_obj, write_obj
    from meshode_b200.synth import synth_pair
    srcV, srcF, tarV, tarF = synth_pair(21, 1500, 1200)
    s_obj, t_obj, o_obj = (str(tmp_path / x) for x in ("source.obj", "target.obj", "output.obj"))
    write_obj(s_obj, srcV, srcF); write_obj(t_obj, tarV, tarF)
    path = (ext_dir if module == "compiled" else ROOT)
    env = dict(os.environ, PYTHONPATH=path)
    p = subprocess.run([sys.executable, os.path.join(ref_py, "rigid_deform.py"), "--source", s_obj, "--target", t_obj,
                        "--output", o_obj], capture_output=True, text=True, timeout=1500, env=env, cwd=str(tmp_path))
    assert p.returncode == 0, p.stdout[-1000:] + p.stderr[-3000:]
    lines = re.findall(r"^iter=(\d+) loss=([0-9.]+)$", p.stdout, re.M)
    assert len(lines) == 100 and lines[0][0] == "0" and lines[-1][0] == "9900"

    # the same loop on the oracle (what the reference's C++ computes) with the same optimiser
    sV, sF = read_obj(s_obj); tV, tF = read_obj(t_obj)        # the float32 values the script loaded
    T = oracle.Template(tV, tF, 64)
    V0 = oracle.normalize_by_template(sV, T.scale, T.trans)
    rest = oracle.store_rigid(V0, sF)
    V = torch.nn.Parameter(torch.from_numpy(V0.copy()))
    opt = torch.optim.Adam([V], lr=1e-3)
    expect = []
    for it in range(10000):
        opt.zero_grad()
        v = V.detach().numpy()
        if it % 100 == 0:
            lossD = torch.from_numpy(oracle.distfield_forward(T.grid, v)) * 0.5
            lossR = torch.from_numpy(oracle.rigid_forward(v, sF, rest)) * 0.5
            expect.append((str(it), "%.6f" % (lossD.sum() + lossR.sum()).item()))
        V.grad = torch.from_numpy(oracle.distfield_backward(T.grid, v)) + torch.from_numpy(oracle.rigid_backward(v, sF, rest))
        opt.step()
    assert lines == expect
    out = oracle.denormalize_by_template(V.detach().numpy(), T.scale, T.trans)
    write_obj(str(tmp_path / "expected.obj"), out, sF)
    assert open(o_obj).read() == open(str(tmp_path / "expected.obj")).read()
