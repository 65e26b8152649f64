"""Drop-in ``pyDeform`` module: the 18 functions of the reference's pybind11 module
(src/interface/pydeform.cc:14-39), same names and positional signatures, executed by
libmeshode_b200.so on the GPU.

Tensors may live on a CUDA device (zero-copy: the kernels read ``data_ptr()`` on torch's
current stream and results are allocated on the same device) or on the CPU as in the
reference's scripts (they are staged through the GPU and results come back on the CPU;
in-place functions write back into the caller's storage).  Unlike the reference, dtype,
contiguity, shape and ``param_id`` are checked and violations raise.

``import torch`` must precede ``import pyDeform`` as in the reference (README.md:46-50) --
here simply because this module imports torch itself.
"""
import numpy as np
import torch

from . import capi
from .capi import EDGES_CAD, EDGES_GRAPH, EDGES_RIGID, MeshodeError
from .objio import read_obj, write_obj

__all__ = [
    "LoadMesh", "LoadCadMesh", "SaveMesh", "InitializeDeformTemplate", "NormalizeByTemplate",
    "DenormalizeByTemplate", "SolveLinear", "DistanceFieldLoss_forward", "DistanceFieldLoss_backward",
    "RigidEdgeLoss_forward", "RigidEdgeLoss_backward", "StoreRigidityInformation", "CadEdgeLoss_forward",
    "CadEdgeLoss_backward", "StoreCadInformation", "GraphEdgeLoss_forward", "GraphEdgeLoss_backward",
    "StoreGraphInformation",
]


def _device():
    capi.require_device()
    return torch.device("cuda", torch.cuda.current_device())


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _check(t, dtype, cols, name):
    """Raises unless ``t`` is a contiguous [n, cols] tensor of ``dtype`` (the reference checks nothing)."""
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if t.dtype != dtype:
        raise TypeError("%s must have dtype %s, got %s" % (name, dtype, t.dtype))
    if t.dim() != 2 or t.shape[1] != cols:
        raise ValueError("%s must have shape [n, %d], got %s" % (name, cols, tuple(t.shape)))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)


def _dev(t, dtype, cols, name):
    """Validated device view of ``t`` ([n, cols], ``dtype``, contiguous)."""
    _check(t, dtype, cols, name)
    if t.is_cuda:
        return t
    return t.to(_device(), non_blocking=False)


def _back(res, like):
    return res if like.is_cuda else res.cpu()


def _pid(param_id):
    if isinstance(param_id, torch.Tensor):
        param_id = param_id.item()
    return int(param_id)


def _ptr(t):
    return t.data_ptr() if t is not None and t.numel() > 0 else 0


# ---- mesh I/O (host side, src/interface/mesh_tensor.cc:87-100, :180-186) ---------------
def LoadMesh(filename):
    V, F = read_obj(filename)
    return [torch.from_numpy(V), torch.from_numpy(F)]


def SaveMesh(filename, tensorV, tensorF):
    write_obj(filename, tensorV.detach().cpu().numpy(), tensorF.detach().cpu().numpy())


def LoadCadMesh(filename):
    """src/interface/mesh_tensor.cc:102-178: [V f32 [n,3], F i32 [m,3], E i32 [e,2], V2G i32 [n,1], GV f32 [g,3],
    GE i32 [ge,2]] -- the CAD mesh cleaned, subdivided to 2e-2 edges, its geometric neighbour pairs (1.5e-2 cells)
    and its deformation graph (1e-2 cells).  Host code as in the reference, with scipy's Delaunay instead of CGAL's."""
    from . import cadmesh
    V, F = read_obj(filename, vertex_dtype=np.float64)
    return [torch.from_numpy(a) for a in cadmesh.load_cad_mesh(V, F)]


def SolveLinear(tensorV, tensorF, tensorE, tensorRef, tensorGraphV, rigidity, with_rot):
    """src/interface/linear_layer.cc:5-84: ``tensorV`` (float32 [n,3]) is overwritten with the least-squares
    positions that follow the deformed graph ``tensorGraphV``.  Host code, as in the reference (the sparse
    solve runs once per shape, after the optimisation).  With ``with_rot`` the reference reinterprets the
    graph tensor as the per-vertex target positions (it must then have n rows)."""
    from . import linear
    V = tensorV.detach().cpu().numpy().astype(np.float64)
    F = tensorF.detach().cpu().numpy()
    GV = tensorGraphV.detach().cpu().numpy().astype(np.float64)
    if not with_rot:
        out = linear.linear_estimation(V, F, tensorE.detach().cpu().numpy(), tensorRef.detach().cpu().numpy().reshape(-1), GV,
                                       float(rigidity))
    else:
        if GV.shape[0] != V.shape[0]:
            raise ValueError("with_rot: tensorGraphV must hold one target position per vertex")
        out = linear.linear_estimation_with_rot(V, F, GV, float(rigidity))
    with torch.no_grad():
        tensorV.copy_(torch.from_numpy(out.astype(np.float32)).to(tensorV.device))


# ---- template ----------------------------------------------------------------------------
def InitializeDeformTemplate(tensorV, tensorF, symmetry, grid_resolution):
    """deform_params.cc:16-40 -> param_id (int)."""
    _check(tensorV, torch.float32, 3, "tensorV")
    _check(tensorF, torch.int32, 3, "tensorF")
    V = _dev(tensorV, torch.float32, 3, "tensorV")
    F = _dev(tensorF, torch.int32, 3, "tensorF").to(V.device)
    with torch.cuda.device(V.device):
        return capi.template_create(_ptr(V), V.shape[0], _ptr(F), F.shape[0], int(symmetry), int(grid_resolution),
                                    _stream(V))


def DestroyTemplate(param_id, stream_ordered=True):
    """Additive: frees the device buffers of a template (the reference never frees g_params).
    ``stream_ordered`` (default): the buffers return to the pool in the order of torch's CURRENT stream, so every
    use of the template must have been enqueued on it or be ordered before it (``wait_stream``) -- the same rule
    torch's caching allocator applies to tensors.  ``stream_ordered=False`` waits for the device instead."""
    if stream_ordered and torch.cuda.is_available():
        capi.template_destroy(_pid(param_id), torch.cuda.current_stream().cuda_stream)
    else:
        capi.template_destroy(_pid(param_id))


def _normalize(tensorV, param_id, inverse):
    V = _dev(tensorV, torch.float32, 3, "tensorV")
    with torch.cuda.device(V.device):
        capi.check(capi.lib().mo_normalize_by_template(_ptr(V), V.shape[0], _pid(param_id), inverse, _stream(V)))
    if V is not tensorV:
        tensorV.copy_(V)


def NormalizeByTemplate(tensorV, param_id):
    _normalize(tensorV, param_id, 0)


def DenormalizeByTemplate(tensorV, param_id):
    _normalize(tensorV, param_id, 1)


# ---- distance field loss -----------------------------------------------------------------
def DistanceFieldLoss_forward(tensorV, param_id):
    V = _dev(tensorV, torch.float32, 3, "tensorV")
    out = torch.empty(V.shape[0], dtype=torch.float32, device=V.device)
    with torch.cuda.device(V.device):
        capi.check(capi.lib().mo_distance_forward(_ptr(V), V.shape[0], _pid(param_id), _ptr(out), _stream(V)))
    return _back(out, tensorV)


def DistanceFieldLoss_backward(tensorV, param_id):
    V = _dev(tensorV, torch.float32, 3, "tensorV")
    out = torch.empty((V.shape[0], 3), dtype=torch.float32, device=V.device)
    with torch.cuda.device(V.device):
        capi.check(capi.lib().mo_distance_backward(_ptr(V), V.shape[0], _pid(param_id), _ptr(out), _stream(V)))
    return _back(out, tensorV)


def DistanceFieldLoss_forward_backward(tensorV, param_id):
    """Additive: (forward, backward) from one pass over V."""
    V = _dev(tensorV, torch.float32, 3, "tensorV")
    out = torch.empty(V.shape[0], dtype=torch.float32, device=V.device)
    grad = torch.empty((V.shape[0], 3), dtype=torch.float32, device=V.device)
    with torch.cuda.device(V.device):
        capi.check(capi.lib().mo_distance_forward_backward(_ptr(V), V.shape[0], _pid(param_id), _ptr(out), _ptr(grad),
                                                           _stream(V)))
    return _back(out, tensorV), _back(grad, tensorV)


# ---- edge losses -------------------------------------------------------------------------
def _edges(fn, kind, tensorV, tensorF, tensorE, param_id, out_rows):
    _check(tensorV, torch.float32, 3, "tensorV")
    if tensorF is not None:
        _check(tensorF, torch.int32, 3, "tensorF")
    if tensorE is not None:
        _check(tensorE, torch.int32, 2, "tensorE")
    V = _dev(tensorV, torch.float32, 3, "tensorV")
    F = _dev(tensorF, torch.int32, 3, "tensorF").to(V.device) if tensorF is not None else None
    E = _dev(tensorE, torch.int32, 2, "tensorE").to(V.device) if tensorE is not None else None
    nF = F.shape[0] if F is not None else 0
    nE = E.shape[0] if E is not None else 0
    out = None
    args = [_pid(param_id), kind, _ptr(V), V.shape[0], _ptr(F), nF, _ptr(E), nE]
    if out_rows is not None:
        rows = {"edges": {EDGES_RIGID: 3 * nF, EDGES_GRAPH: nE, EDGES_CAD: nE + 3 * nF}[kind], "verts": V.shape[0]}[out_rows]
        out = torch.empty((rows, 3), dtype=torch.float32, device=V.device)
        args.append(_ptr(out))
    with torch.cuda.device(V.device):
        capi.check(fn(*args, _stream(V)))
    return _back(out, tensorV) if out is not None else None


def StoreRigidityInformation(tensorV, tensorF, param_id):
    _edges(capi.lib().mo_edges_store, EDGES_RIGID, tensorV, tensorF, None, param_id, None)


def RigidEdgeLoss_forward(tensorV, tensorF, param_id):
    return _edges(capi.lib().mo_edges_forward, EDGES_RIGID, tensorV, tensorF, None, param_id, "edges")


def RigidEdgeLoss_backward(tensorV, tensorF, param_id):
    return _edges(capi.lib().mo_edges_backward, EDGES_RIGID, tensorV, tensorF, None, param_id, "verts")


def StoreGraphInformation(tensorV, tensorE, param_id):
    _edges(capi.lib().mo_edges_store, EDGES_GRAPH, tensorV, None, tensorE, param_id, None)


def GraphEdgeLoss_forward(tensorV, tensorE, param_id):
    return _edges(capi.lib().mo_edges_forward, EDGES_GRAPH, tensorV, None, tensorE, param_id, "edges")


def GraphEdgeLoss_backward(tensorV, tensorE, param_id):
    return _edges(capi.lib().mo_edges_backward, EDGES_GRAPH, tensorV, None, tensorE, param_id, "verts")


def StoreCadInformation(tensorV, tensorF, tensorE, param_id):
    _edges(capi.lib().mo_edges_store, EDGES_CAD, tensorV, tensorF, tensorE, param_id, None)


def CadEdgeLoss_forward(tensorV, tensorF, tensorE, param_id):
    return _edges(capi.lib().mo_edges_forward, EDGES_CAD, tensorV, tensorF, tensorE, param_id, "edges")


def CadEdgeLoss_backward(tensorV, tensorF, tensorE, param_id):
    return _edges(capi.lib().mo_edges_backward, EDGES_CAD, tensorV, tensorF, tensorE, param_id, "verts")


# ---- additive helpers used by the device-native layers and the tests ------------------------
def EdgeLoss_backward_atomic(kind, tensorV, tensorF, tensorE, param_id):
    return _edges(capi.lib().mo_edges_backward_atomic, kind, tensorV, tensorF, tensorE, param_id, "verts")


def LossForwardBackward(tensorV, dist_param_id, edge_param_id, w_edge=1.0, mask_threshold=0.0, want_loss=True,
                        want_grad=True):
    """One launch: loss (0-d float64 tensor) and gradient [n,3] of the layers' combined loss."""
    V = _dev(tensorV, torch.float32, 3, "tensorV")
    loss = torch.empty((), dtype=torch.float64, device=V.device) if want_loss else None
    grad = torch.empty((V.shape[0], 3), dtype=torch.float32, device=V.device) if want_grad else None
    with torch.cuda.device(V.device):
        capi.check(capi.lib().mo_loss_forward_backward(_pid(dist_param_id), _pid(edge_param_id), _ptr(V), V.shape[0],
                                                       float(w_edge), float(mask_threshold),
                                                       loss.data_ptr() if want_loss else 0, _ptr(grad), _stream(V)))
    return (_back(loss, tensorV) if want_loss else None), (_back(grad, tensorV) if want_grad else None)


def GetTemplateInfo(param_id):
    return capi.template_info(_pid(param_id), torch.cuda.current_stream().cuda_stream)


def GetGrid(param_id, z0=0, z1=None):
    """Additive (tests / z-slab all-gather): (grid_f64 [N,N,N], grid_f32, nearest) torch tensors
    holding slices [z0,z1) of the template's fields (other slices zero)."""
    pid = _pid(param_id)
    N = capi.template_info(pid, torch.cuda.current_stream().cuda_stream)["N"]
    z1 = N if z1 is None else z1
    dev = _device()
    g64 = torch.zeros((N, N, N), dtype=torch.float64, device=dev)
    g32 = torch.zeros((N, N, N), dtype=torch.float32, device=dev)
    idx = torch.zeros((N, N, N), dtype=torch.int32, device=dev)
    capi.check(capi.lib().mo_template_copy_grid(pid, 0, z0, z1, g64.data_ptr(), g32.data_ptr(), idx.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream))
    return g64, g32, idx


def SetGrid(param_id, g64, g32, idx, z0=0, z1=None):
    """Additive: writes slices [z0,z1) of full-size device tensors into the template (after an all-gather)."""
    pid = _pid(param_id)
    N = g32.shape[0]
    z1 = N if z1 is None else z1
    capi.check(capi.lib().mo_template_copy_grid(pid, 1, z0, z1, _ptr(g64), _ptr(g32), _ptr(idx),
                                                torch.cuda.current_stream().cuda_stream))


class _DeviceArray:
    """A raw device buffer exposed through __cuda_array_interface__ so that torch can alias it."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3,
                                         "strides": None}


def GridViews(param_id):
    """Additive: torch tensors ALIASING the template's own device fields (grid_f64, grid_f32, nearest), each
    [N,N,N], for in-place collectives on a z-slab sharded build.  They are views: do not use them after
    DestroyTemplate, and destroy the derived corner tables by calling SetGrid if the field is changed."""
    pid = _pid(param_id)
    N = capi.template_info(pid, torch.cuda.current_stream().cuda_stream)["N"]
    p64, p32, pidx = capi.template_grid(pid)
    dev = _device()
    return (torch.as_tensor(_DeviceArray(p64, (N, N, N), "<f8"), device=dev),
            torch.as_tensor(_DeviceArray(p32, (N, N, N), "<f4"), device=dev),
            torch.as_tensor(_DeviceArray(pidx, (N, N, N), "<i4"), device=dev))


def NearestVertex(tensorQ, tensorP, return_dist2=False):
    """Additive (ReverseLossLayer): index [nQ] int32 of the nearest row of ``tensorP`` for every row of
    ``tensorQ`` -- scipy's cKDTree(P).query(Q, k=1) on the GPU, exact, FP64 distances."""
    _check(tensorQ, torch.float32, 3, "tensorQ")
    _check(tensorP, torch.float32, 3, "tensorP")
    Q = _dev(tensorQ, torch.float32, 3, "tensorQ")
    P = _dev(tensorP, torch.float32, 3, "tensorP").to(Q.device)
    idx = torch.empty(Q.shape[0], dtype=torch.int32, device=Q.device)
    d2 = torch.empty(Q.shape[0], dtype=torch.float64, device=Q.device) if return_dist2 else None
    with torch.cuda.device(Q.device):
        capi.check(capi.lib().mo_nearest_vertex(_ptr(Q), Q.shape[0], _ptr(P), P.shape[0], _ptr(idx),
                                                d2.data_ptr() if return_dist2 else 0, _stream(Q)))
    return (_back(idx, tensorQ), _back(d2, tensorQ)) if return_dist2 else _back(idx, tensorQ)


# ---- the Ceres loss terms of the C++ drivers (additive: the reference has no Python binding for them) ----
def _f64(t, cols, name, device=None):
    _check(t, torch.float64, cols, name)
    return t if t.is_cuda else t.to(device if device is not None else _device())


def CeresEdges(kind, V, R, I, rest, lam, want_jacobian=False):
    """Residual blocks EdgeLoss / AdaptiveEdgeLoss / EdgeLossWithRot (src/lib/edgeloss.h) for the edge list
    ``I`` [e,2]: returns residuals [e,3] ([e,6] for capi.CERES_ROT_EDGE) and, optionally, the autodiff
    Jacobians [e,3,6] / [e,6,12].  float64 tensors."""
    V = _f64(V, 3, "V")
    dev = V.device
    R = _f64(R, 3, "R", dev) if R is not None else None
    _check(I, torch.int32, 2, "I")
    I = I.to(dev)
    rest = _f64(rest, 3, "rest", dev)
    rot = kind == capi.CERES_ROT_EDGE
    res = torch.empty((I.shape[0], 6 if rot else 3), dtype=torch.float64, device=dev)
    jac = torch.empty((I.shape[0], 6, 12) if rot else (I.shape[0], 3, 6), dtype=torch.float64, device=dev) if want_jacobian else None
    with torch.cuda.device(dev):
        capi.check(capi.lib().mo_ceres_edges(int(kind), _ptr(V), _ptr(R), V.shape[0], _ptr(I), _ptr(rest), I.shape[0],
                                             float(lam), _ptr(res), _ptr(jac), _stream(V)))
    return (res, jac) if want_jacobian else res


def CeresProblem(dist_param_id, kind, V, R, I, rest, lam):
    """Cost and gradient of the Deformer problems (src/lib/deformer.cc): returns
    (cost_distance, cost_edges) as a float64 [2] tensor, gV [n,3] and gR [n,3] (None unless ROT)."""
    V = _f64(V, 3, "V")
    dev = V.device
    rot = kind == capi.CERES_ROT_EDGE
    R = _f64(R, 3, "R", dev) if R is not None else None
    _check(I, torch.int32, 2, "I")
    I = I.to(dev)
    rest = _f64(rest, 3, "rest", dev)
    cost = torch.empty(2, dtype=torch.float64, device=dev)
    gV = torch.empty_like(V)
    gR = torch.empty_like(V) if rot else None
    with torch.cuda.device(dev):
        capi.check(capi.lib().mo_ceres_problem(_pid(dist_param_id) if dist_param_id is not None else -1, int(kind), _ptr(V),
                                               _ptr(R), V.shape[0], _ptr(I), _ptr(rest), I.shape[0], float(lam),
                                               _ptr(cost), _ptr(gV), _ptr(gR), _stream(V)))
    return cost, gV, gR


SOLVE_TERMINATION = ("function tolerance", "gradient tolerance", "parameter tolerance", "iteration limit",
                     "invalid steps", "radius underflow")


def CeresSolve(dist_param_id, kind, V, R, I, rest, lam, max_iterations=100, max_cg_iterations=4000, cg_tolerance=1e-10,
               verbose=False):
    """ceres::Solve for the Deformer problems (src/lib/deformer.cc:55-74, :135-153): Levenberg-Marquardt with
    Ceres' default options, the LM step solved matrix-free with preconditioned CG on the GPU.  ``V`` [n,3]
    (and ``R`` [n,3] for capi.CERES_ROT_EDGE) float64 CUDA tensors are optimised IN PLACE; returns a summary
    dict (initial_cost, final_cost, vertices_cost, rigidity_cost, iterations, accepted, cg_iterations,
    termination)."""
    import ctypes as C
    _check(V, torch.float64, 3, "V")
    if not V.is_cuda:
        raise ValueError("CeresSolve optimises in place: V must be a CUDA tensor")
    dev = V.device
    rot = kind == capi.CERES_ROT_EDGE
    if rot:
        _check(R, torch.float64, 3, "R")
        if not R.is_cuda:
            raise ValueError("CeresSolve optimises in place: R must be a CUDA tensor")
    _check(I, torch.int32, 2, "I")
    I = I.to(dev)
    rest = _f64(rest, 3, "rest", dev)
    summ = (C.c_double * 10)()
    with torch.cuda.device(dev):
        capi.check(capi.lib().mo_ceres_solve(_pid(dist_param_id) if dist_param_id is not None else -1, int(kind), _ptr(V),
                                             _ptr(R) if rot else 0, V.shape[0], _ptr(I), _ptr(rest), I.shape[0], float(lam),
                                             int(max_iterations), int(max_cg_iterations), float(cg_tolerance),
                                             int(bool(verbose)), C.addressof(summ), _stream(V)))
    return {"initial_cost": summ[0], "final_cost": summ[1], "vertices_cost": summ[2], "rigidity_cost": summ[3],
            "iterations": int(summ[4]), "accepted": int(summ[5]), "cg_iterations": int(summ[6]),
            "termination": SOLVE_TERMINATION[int(summ[7])], "radius": summ[8], "gradient_max": summ[9]}
