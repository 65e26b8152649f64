"""ctypes front-end of the CPU oracle (oracle/meshode_oracle.cc).

TEST INFRASTRUCTURE ONLY -- sampler, loss functors and Adam are pinned, the nearest-triangle search is
parity unpinned (see the header of meshode_oracle.cc).
Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs; never by the product package ``meshode_b200``.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmeshode_oracle.so")
_lib = None

_f32 = np.float32
_f64 = np.float64
_i32 = np.int32


def build(force=False):
    """Compile the oracle with the recipe in oracle/Makefile."""
    src = os.path.join(_HERE, "meshode_oracle.cc")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libmeshode_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_point_triangle_sqr.restype = C.c_double
        _lib.orc_rot_problem_cost_grad.restype = C.c_double
        _lib.orc_deform_problem_cost_grad.restype = C.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def ncores():
    return os.cpu_count() or 1


# ---- template -------------------------------------------------------------
def normalize_target(V):
    V = _c(V, _f32)
    Vn = np.empty((V.shape[0], 3), _f64)
    scale = C.c_double()
    pos = np.empty(3, _f64)
    lib().orc_normalize_target(_p(V), C.c_int(V.shape[0]), _p(Vn), C.byref(scale), _p(pos))
    return Vn, scale.value, pos


def point_triangle_sqr(p, a, b, c):
    p, a, b, c = (_c(x, _f64) for x in (p, a, b, c))
    q = np.empty(3, _f64)
    d = lib().orc_point_triangle_sqr(_p(p), _p(a), _p(b), _p(c), _p(q))
    return d, q


def build_grid(Vn, F, N, z0=0, z1=None, fast=True, threads=None, want_idx=True):
    """FP64 distance grid [z][y][x] and nearest-triangle index (mesh.cc:106-152).  ``fast``: True / "bvh" = bounding-box
    tree (what libigl's query does; the CPU baseline), "cells" = uniform cell index with a ring search, False = brute
    force over all triangles (the ground truth the other two are checked against).  Same result from all three."""
    Vn = _c(Vn, _f64)
    F = _c(F, _i32)
    z1 = N if z1 is None else z1
    threads = ncores() if threads is None else threads
    grid = np.full((N, N, N), 1e30, _f64)  # uniformgrid.cc:9-17 initial value
    idx = np.full((N, N, N), -1, _i32) if want_idx else None
    ip = _p(idx) if want_idx else None
    if fast is True or fast == "bvh":
        lib().orc_build_grid_bvh(_p(Vn), C.c_int(Vn.shape[0]), _p(F), C.c_int(F.shape[0]), C.c_int(N), C.c_int(z0),
                                 C.c_int(z1), _p(grid), ip, C.c_int(threads))
    elif fast == "cells":
        lib().orc_build_grid_fast(_p(Vn), C.c_int(Vn.shape[0]), _p(F), C.c_int(F.shape[0]), C.c_int(N), C.c_int(z0),
                                  C.c_int(z1), _p(grid), ip, C.c_int(threads))
    else:
        lib().orc_build_grid_brute(_p(Vn), _p(F), C.c_int(F.shape[0]), C.c_int(N), C.c_int(z0), C.c_int(z1), _p(grid),
                                   ip, C.c_int(threads))
    return grid, idx


class Template:
    """What InitializeDeformTemplate leaves in g_params (deform_params.cc:16-40)."""

    def __init__(self, tarV, tarF, N, fast=True, threads=None):
        self.N = int(N)
        self.F = _c(tarF, _i32)
        self.Vn, self.scale, self.trans = normalize_target(tarV)
        self.grid, self.idx = build_grid(self.Vn, self.F, self.N, fast=fast, threads=threads)


# ---- samplers -------------------------------------------------------------
def distance_float(grid, P):
    grid = _c(grid, _f64); P = _c(P, _f32)
    out = np.empty(P.shape[0], _f32)
    lib().orc_distance_float(_p(grid), C.c_int(grid.shape[0]), _p(P), C.c_int(P.shape[0]), _p(out))
    return out


def distance_double(grid, P):
    grid = _c(grid, _f64); P = _c(P, _f64)
    out = np.empty(P.shape[0], _f64)
    lib().orc_distance_double(_p(grid), C.c_int(grid.shape[0]), _p(P), C.c_int(P.shape[0]), _p(out))
    return out


def distance_double_jet(grid, P):
    grid = _c(grid, _f64); P = _c(P, _f64)
    val = np.empty(P.shape[0], _f64); grad = np.empty((P.shape[0], 3), _f64)
    lib().orc_distance_double_jet(_p(grid), C.c_int(grid.shape[0]), _p(P), C.c_int(P.shape[0]), _p(val), _p(grad))
    return val, grad


def distance_float_jet(grid, P):
    grid = _c(grid, _f64); P = _c(P, _f32)
    val = np.empty(P.shape[0], _f32); grad = np.empty((P.shape[0], 3), _f32)
    lib().orc_distance_float_jet(_p(grid), C.c_int(grid.shape[0]), _p(P), C.c_int(P.shape[0]), _p(val), _p(grad))
    return val, grad


def distfield_forward(grid, V):
    grid = _c(grid, _f64); V = _c(V, _f32)
    out = np.empty(V.shape[0], _f32)
    lib().orc_distfield_forward(_p(grid), C.c_int(grid.shape[0]), _p(V), C.c_int(V.shape[0]), _p(out))
    return out


def distfield_backward(grid, V):
    grid = _c(grid, _f64); V = _c(V, _f32)
    out = np.empty((V.shape[0], 3), _f32)
    lib().orc_distfield_backward(_p(grid), C.c_int(grid.shape[0]), _p(V), C.c_int(V.shape[0]), _p(out))
    return out


# ---- edges ----------------------------------------------------------------
def store_rigid(V, F):
    V = _c(V, _f32); F = _c(F, _i32)
    rest = np.empty((3 * F.shape[0], 3), _f32)
    lib().orc_store_rigid(_p(V), _p(F), C.c_int(F.shape[0]), _p(rest))
    return rest


def rigid_forward(V, F, rest):
    V = _c(V, _f32); F = _c(F, _i32); rest = _c(rest, _f32)
    out = np.empty((3 * F.shape[0], 3), _f32)
    lib().orc_rigid_forward(_p(V), _p(F), C.c_int(F.shape[0]), _p(rest), _p(out))
    return out


def rigid_backward(V, F, rest):
    V = _c(V, _f32); F = _c(F, _i32); rest = _c(rest, _f32)
    out = np.empty((V.shape[0], 3), _f32)
    lib().orc_rigid_backward(_p(V), C.c_int(V.shape[0]), _p(F), C.c_int(F.shape[0]), _p(rest), _p(out))
    return out


def store_graph(V, E):
    V = _c(V, _f32); E = _c(E, _i32)
    rest = np.empty((E.shape[0], 3), _f32)
    lib().orc_store_graph(_p(V), _p(E), C.c_int(E.shape[0]), _p(rest))
    return rest


def graph_forward(V, E, rest):
    V = _c(V, _f32); E = _c(E, _i32); rest = _c(rest, _f32)
    out = np.empty((E.shape[0], 3), _f32)
    lib().orc_graph_forward(_p(V), _p(E), C.c_int(E.shape[0]), _p(rest), _p(out))
    return out


def graph_backward(V, E, rest):
    V = _c(V, _f32); E = _c(E, _i32); rest = _c(rest, _f32)
    out = np.empty((V.shape[0], 3), _f32)
    lib().orc_graph_backward(_p(V), C.c_int(V.shape[0]), _p(E), C.c_int(E.shape[0]), _p(rest), _p(out))
    return out


def store_cad(V, F, E):
    V = _c(V, _f32); F = _c(F, _i32); E = _c(E, _i32).reshape(-1, 2)
    n = E.shape[0] + 3 * F.shape[0]
    rest = np.empty((n, 3), _f32); lam = np.empty(n, _f32)
    lib().orc_store_cad(_p(V), _p(F), C.c_int(F.shape[0]), _p(E), C.c_int(E.shape[0]), _p(rest), _p(lam))
    return rest, lam


def cad_forward(V, F, E, rest, lam):
    V = _c(V, _f32); F = _c(F, _i32); E = _c(E, _i32).reshape(-1, 2); rest = _c(rest, _f32); lam = _c(lam, _f32)
    out = np.empty((E.shape[0] + 3 * F.shape[0], 3), _f32)
    lib().orc_cad_forward(_p(V), _p(F), C.c_int(F.shape[0]), _p(E), C.c_int(E.shape[0]), _p(rest), _p(lam), _p(out))
    return out


def cad_backward(V, F, E, rest, lam):
    V = _c(V, _f32); F = _c(F, _i32); E = _c(E, _i32).reshape(-1, 2); rest = _c(rest, _f32); lam = _c(lam, _f32)
    out = np.empty((V.shape[0], 3), _f32)
    lib().orc_cad_backward(_p(V), C.c_int(V.shape[0]), _p(F), C.c_int(F.shape[0]), _p(E), C.c_int(E.shape[0]),
                           _p(rest), _p(lam), _p(out))
    return out


def normalize_by_template(V, scale, trans):
    V = np.array(V, dtype=_f32, order="C", copy=True)
    trans = _c(trans, _f64)
    lib().orc_normalize_by_template(_p(V), C.c_int(V.shape[0]), C.c_double(scale), _p(trans))
    return V


def denormalize_by_template(V, scale, trans):
    V = np.array(V, dtype=_f32, order="C", copy=True)
    trans = _c(trans, _f64)
    lib().orc_denormalize_by_template(_p(V), C.c_int(V.shape[0]), C.c_double(scale), _p(trans))
    return V


# ---- Ceres functors -------------------------------------------------------
def edge_loss(p1, p2, v, lam, adaptive=False):
    p1, p2, v = (_c(x, _f64) for x in (p1, p2, v))
    r = np.empty(3, _f64); le = C.c_double()
    lib().orc_edge_loss(_p(p1), _p(p2), _p(v), C.c_double(lam), C.c_int(int(adaptive)), _p(r), C.byref(le))
    return r, le.value


def edge_rot(p1, p2, rot1, rot2, v, lam):
    p1, p2, rot1, rot2, v = (_c(x, _f64) for x in (p1, p2, rot1, rot2, v))
    r = np.empty(6, _f64); J = np.empty((6, 12), _f64)
    lib().orc_edge_rot(_p(p1), _p(p2), _p(rot1), _p(rot2), _p(v), C.c_double(lam), _p(r), _p(J))
    return r, J


def rot_problem_cost_grad(grid, V, R, F, rest, lam):
    grid = _c(grid, _f64); V = _c(V, _f64); R = _c(R, _f64); F = _c(F, _i32); rest = _c(rest, _f64)
    gV = np.empty_like(V); gR = np.empty_like(R)
    cd = C.c_double(); ce = C.c_double()
    lib().orc_rot_problem_cost_grad(_p(grid), C.c_int(grid.shape[0]), _p(V), _p(R), C.c_int(V.shape[0]), _p(F),
                                    C.c_int(F.shape[0]), _p(rest), C.c_double(lam), _p(gV), _p(gR), C.byref(cd),
                                    C.byref(ce))
    return cd.value, ce.value, gV, gR


def deform_problem_cost_grad(grid, V, F, rest, lam, adaptive=False):
    grid = _c(grid, _f64); V = _c(V, _f64); F = _c(F, _i32); rest = _c(rest, _f64)
    gV = np.empty_like(V)
    cd = C.c_double(); ce = C.c_double()
    lib().orc_deform_problem_cost_grad(_p(grid), C.c_int(grid.shape[0]), _p(V), C.c_int(V.shape[0]), _p(F),
                                       C.c_int(F.shape[0]), _p(rest), C.c_double(lam), C.c_int(int(adaptive)), _p(gV),
                                       C.byref(cd), C.byref(ce))
    return cd.value, ce.value, gV


# ---- whole loops ----------------------------------------------------------
def rigid_adam(grid, V, F, rest, iters, lr=1e-3, log_every=0):
    """src/python/rigid_deform.py:32-41 on an already normalised source. Returns (V, loss_log)."""
    grid = _c(grid, _f64); F = _c(F, _i32); rest = _c(rest, _f32)
    V = np.array(V, dtype=_f32, order="C", copy=True)
    log = np.zeros((iters + log_every - 1) // log_every if log_every else 0, _f64)
    lib().orc_rigid_adam(_p(grid), C.c_int(grid.shape[0]), _p(V), C.c_int(V.shape[0]), _p(F), C.c_int(F.shape[0]),
                         _p(rest), C.c_int(iters), C.c_double(lr), _p(log) if log_every else None, C.c_int(log_every))
    return V, log


def deform_pair(tarV, tarF, srcV, srcF, N=64, iters=10000, lr=1e-3, build_threads=1):
    """InitializeDeformTemplate + NormalizeByTemplate + StoreRigidityInformation + Adam + Denormalize."""
    tarV = _c(tarV, _f32); tarF = _c(tarF, _i32); srcF = _c(srcF, _i32)
    V = np.array(srcV, dtype=_f32, order="C", copy=True)
    lib().orc_deform_pair(_p(tarV), C.c_int(tarV.shape[0]), _p(tarF), C.c_int(tarF.shape[0]), _p(V),
                          C.c_int(V.shape[0]), _p(srcF), C.c_int(srcF.shape[0]), C.c_int(N), C.c_int(iters),
                          C.c_double(lr), C.c_int(build_threads))
    return V
