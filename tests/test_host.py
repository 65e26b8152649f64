"""Host-side logic without a GPU: synthetic shape generator, OBJ I/O semantics, argument
validation of the pyDeform mirror, and its loud failure when no CUDA device exists."""
import os

import numpy as np
import pytest


def test_synth_mesh_is_deterministic_closed_manifold():
    from meshode_b200.synth import synth_mesh, synth_pair, unique_edges
    V, F = synth_mesh(500, 3)
    V2, F2 = synth_mesh(500, 3)
    assert np.array_equal(V, V2) and np.array_equal(F, F2)
    assert V.dtype == np.float32 and F.dtype == np.int32 and F.shape == (2 * 500 - 4, 3)
    E = unique_edges(F)
    assert E.shape[0] == 3 * 500 - 6                      # Euler: closed genus-0 triangulation
    # every undirected edge is used by exactly two triangles, once per direction (oriented manifold)
    d = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]])
    key = d[:, 0].astype(np.int64) * 500 + d[:, 1]
    assert np.unique(key).size == key.size
    assert np.isin(d[:, 1].astype(np.int64) * 500 + d[:, 0], key).all()
    # outward orientation: positive signed volume
    a, b, c = (V[F[:, k]].astype(np.float64) for k in range(3))
    assert np.einsum("ij,ij->i", a, np.cross(b, c)).sum() > 0
    sV, sF, tV, tF = synth_pair(2, 300, 400)
    assert sV.shape == (300, 3) and tV.shape == (400, 3) and sF.shape == (596, 3) and tF.shape == (796, 3)


def test_obj_io_reference_semantics(tmp_path, meshes):
    from meshode_b200.objio import read_obj, write_obj
    p = tmp_path / "m.obj"
    p.write_text("# comment\nvn 0 0 1\nv 0 0 0\nv 1 0 0\nv 0 1 0\nv 0 0 1\nvt 0 0\n"
                 "f 1//1 2//1 3//1\nf 1/5/2 3/4/2 4/1/2 2/2/2\n\ng grp\nf 2 3 4\n")
    V, F = read_obj(str(p))
    assert V.dtype == np.float32 and F.dtype == np.int32
    assert V.shape == (4, 3)
    # leading index of a/b/c tokens, 1-based, first three tokens of a face (mesh.cc:14-45)
    assert F.tolist() == [[0, 1, 2], [0, 2, 3], [1, 2, 3]]
    q = tmp_path / "o.obj"
    write_obj(str(q), meshes["cadTarV"], meshes["cadTarF"])
    V2, F2 = read_obj(str(q))
    assert np.array_equal(F2, meshes["cadTarF"])
    assert np.allclose(V2, meshes["cadTarV"], rtol=1e-5, atol=0)   # "%.6g"-style text precision


def test_pydeform_surface_matches_reference_names():
    import pyDeform
    names = ["LoadMesh", "LoadCadMesh", "SaveMesh", "InitializeDeformTemplate", "NormalizeByTemplate",
             "DenormalizeByTemplate", "SolveLinear", "DistanceFieldLoss_forward", "DistanceFieldLoss_backward",
             "RigidEdgeLoss_forward", "RigidEdgeLoss_backward", "StoreRigidityInformation", "CadEdgeLoss_forward",
             "CadEdgeLoss_backward", "StoreCadInformation", "GraphEdgeLoss_forward", "GraphEdgeLoss_backward",
             "StoreGraphInformation"]          # src/interface/pydeform.cc:15-37
    for n in names:
        assert callable(getattr(pyDeform, n)), n


def test_pydeform_validates_and_fails_loudly_without_gpu():
    torch = pytest.importorskip("torch")
    from meshode_b200 import capi
    from meshode_b200 import pyDeform as pd
    V = torch.zeros((4, 3), dtype=torch.float32); F = torch.zeros((2, 3), dtype=torch.int32)
    with pytest.raises(TypeError):
        pd.InitializeDeformTemplate(V.double(), F, 0, 8)
    with pytest.raises(TypeError):
        pd.InitializeDeformTemplate(V, F.long(), 0, 8)
    with pytest.raises(ValueError):
        pd.DistanceFieldLoss_forward(torch.zeros((4, 2), dtype=torch.float32), 0)
    with pytest.raises(ValueError):
        pd.DistanceFieldLoss_forward(torch.zeros((3, 8), dtype=torch.float32).t()[:, :3], 0)
    if capi.device_count() == 0:
        with pytest.raises(capi.MeshodeError):          # no CPU fallback
            pd.InitializeDeformTemplate(V, F, 0, 8)
        with pytest.raises(capi.MeshodeError):
            pd.DistanceFieldLoss_forward(V, 0)


def test_golden_fixtures_are_self_consistent(meshes, golden):
    assert meshes["tarV"].shape == (14762, 3) and meshes["tarF"].shape == (29532, 3)      # data/target.obj
    assert meshes["srcV"].shape == (21542, 3) and meshes["srcF"].shape == (43100, 3)      # data/source.obj
    assert golden["grid"].shape == (32, 32, 32) and golden["nearest"].min() >= 0
    assert golden["nearest"].max() < meshes["tarF"].shape[0]


def test_scripts_expose_the_reference_command_lines():
    """scripts/*.py keep the reference scripts' arguments (src/python/rigid_deform.py:15-18, cad_deform2.py:18-22,
    cad_neural_deform2.py:22-28); --help needs no GPU."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    want = {"rigid_deform.py": ["--source", "--target", "--output"],
            "cad_deform2.py": ["--source", "--target", "--output", "--rigidity"],
            "cad_neural_deform2.py": ["--source", "--target", "--output", "--rigidity", "--device", "--save_path"],
            "batch_deform.py": ["--filelist", "--niter", "--grid"]}
    for name, flags in want.items():
        p = subprocess.run([sys.executable, os.path.join(root, "scripts", name), "--help"], capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stderr[-1000:]
        for f in flags:
            assert f in p.stdout, (name, f)


def test_reference_import_paths_and_function_surface():
    """`import pyDeform` and `from layers.X import ...` resolve with the repository root on the path, as the reference's
    scripts import them (src/python/rigid_deform.py:10-11, cad_neural_deform2.py:9-13); the exported autograd Functions
    define their own backward (the reference's are differentiable)."""
    import torch  # noqa: F401  (torch first, README.md:46-50)
    import pyDeform
    from torch.autograd import Function
    from layers.cad_loss_layer import CadLossFunction, CadLossLayer, Finalize as cad_finalize  # noqa: F401
    from layers.graph_loss2_layer import Finalize as g2_finalize, GraphLoss2Function, GraphLoss2Layer  # noqa: F401
    from layers.graph_loss_layer import Finalize as g_finalize, GraphLossFunction, GraphLossLayer  # noqa: F401
    from layers.neuralode_fast import NeuralODE
    from layers.reverse_loss_layer import ReverseLossLayer  # noqa: F401
    from layers.rigid_loss_layer import Finalize, RigidLossFunction, RigidLossLayer  # noqa: F401
    for fn in (RigidLossFunction, GraphLossFunction, GraphLoss2Function, CadLossFunction):
        assert fn.backward is not Function.backward and fn.forward is not Function.forward
    for name in ("LoadMesh", "LoadCadMesh", "SaveMesh", "InitializeDeformTemplate", "NormalizeByTemplate",
                 "DenormalizeByTemplate", "SolveLinear", "DistanceFieldLoss_forward", "DistanceFieldLoss_backward",
                 "RigidEdgeLoss_forward", "RigidEdgeLoss_backward", "StoreRigidityInformation", "CadEdgeLoss_forward",
                 "CadEdgeLoss_backward", "StoreCadInformation", "GraphEdgeLoss_forward", "GraphEdgeLoss_backward",
                 "StoreGraphInformation"):   # src/interface/pydeform.cc:14-39
        assert callable(getattr(pyDeform, name))
    assert NeuralODE.__module__ == "layers.neuralode_fast"


def test_checkpoint_layout_is_the_references(tmp_path):
    """cad_neural_deform2.py:108 pickles {'func': func, 'optim': optimizer}; cad_neural_animate.py:52-57 reads the
    objects back.  The file written here has that layout under the reference's module path, and the loader also
    accepts the state_dict layout of this repository's first round."""
    import zipfile
    import torch
    from meshode_b200.layers.neuralode import NeuralODE, load_checkpoint, save_checkpoint
    f = NeuralODE(torch.device("cpu"))
    o = torch.optim.Adam(f.parameters(), lr=1e-3)
    f.forward(torch.rand(10, 3)).sum().backward()
    o.step()
    p = str(tmp_path / "a.ckpt")
    save_checkpoint(p, f, o)
    z = zipfile.ZipFile(p)
    data = z.read([n for n in z.namelist() if n.endswith("data.pkl")][0])
    assert b"layers.neuralode_fast" in data and b"meshode_b200" not in data
    ck = torch.load(p, map_location="cpu", weights_only=False)
    assert isinstance(ck["func"], NeuralODE) and isinstance(ck["optim"], torch.optim.Adam)   # objects, as the reference reads them
    x = torch.rand(7, 3)
    f2, o2 = load_checkpoint(p, torch.device("cpu"))
    assert torch.equal(f.forward(x), f2.forward(x)) and torch.equal(f.integrate(x, 0, 0.4, "cpu"), f2.integrate(x, 0, 0.4, "cpu"))
    assert o2.state_dict()["state"][0]["step"] == o.state_dict()["state"][0]["step"]
    torch.save({"func": f.func.state_dict(), "optim": o.state_dict()}, p)
    f3, o3 = load_checkpoint(p, torch.device("cpu"))
    assert torch.equal(f.forward(x), f3.forward(x))
