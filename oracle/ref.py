"""ctypes front-end of oracle/_ref/libmeshode_ref.so: the REFERENCE's own sampler and loss
functors (src/lib/uniformgrid.cc, distanceloss.h, edgeloss.h) compiled from /root/reference
against the stand-in third-party headers of oracle/stubs/ (recipe: ``make -C oracle ref``).

TEST INFRASTRUCTURE ONLY.  Exists only where /root/reference does (the build container);
tests that need it skip elsewhere and fall back to tests/golden/golden_ref.npz, which was
generated from it by tests/golden/make_golden_ref.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libmeshode_ref.so")
REFERENCE = os.environ.get("MESHODE_REFERENCE", "/root/reference")
_lib = None


def available():
    return os.path.exists(SO) or os.path.exists(os.path.join(REFERENCE, "src", "lib", "uniformgrid.cc"))


def build(force=False):
    """Compiles the reference's sources in place (no-op when /root/reference is absent)."""
    if not os.path.exists(os.path.join(REFERENCE, "src", "lib", "uniformgrid.cc")):
        return SO if os.path.exists(SO) else None
    stale = os.path.exists(SO) and any(os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(SO)
                                       for f in ("ref_shim.cc", "ref_iface_shim.cc", "Makefile"))
    if force or stale or not os.path.exists(SO):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B" if (force or stale) else "-s", "ref", "REF=" + REFERENCE])
    return SO


def lib():
    global _lib
    if _lib is None:
        if build() is None:
            raise RuntimeError("oracle/_ref is not built and /root/reference is absent")
        _lib = C.CDLL(SO)
        _lib.ref_grid_create.restype = C.c_void_p
        _lib.ref_grid_get.restype = C.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Grid:
    """A reference UniformGrid filled through SetDistance(i=z, j=y, k=x)."""

    def __init__(self, grid):
        grid = np.ascontiguousarray(grid, dtype=np.float64)
        self.N = grid.shape[0]
        self.h = C.c_void_p(lib().ref_grid_create(C.c_int(self.N), _p(grid)))

    def __del__(self):
        try:
            lib().ref_grid_destroy(self.h)
        except Exception:
            pass

    def distance_double(self, P):
        P = np.ascontiguousarray(P, dtype=np.float64)
        out = np.empty(P.shape[0], np.float64)
        lib().ref_distance_double(self.h, _p(P), C.c_int(P.shape[0]), _p(out))
        return out

    def distance_float(self, P):
        P = np.ascontiguousarray(P, dtype=np.float32)
        out = np.empty(P.shape[0], np.float32)
        lib().ref_distance_float(self.h, _p(P), C.c_int(P.shape[0]), _p(out))
        return out

    def distance_double_jet(self, P):
        P = np.ascontiguousarray(P, dtype=np.float64)
        val = np.empty(P.shape[0], np.float64); grad = np.empty((P.shape[0], 3), np.float64)
        lib().ref_distance_double_jet(self.h, _p(P), C.c_int(P.shape[0]), _p(val), _p(grad))
        return val, grad

    def distance_float_jet(self, P):
        P = np.ascontiguousarray(P, dtype=np.float32)
        val = np.empty(P.shape[0], np.float32); grad = np.empty((P.shape[0], 3), np.float32)
        lib().ref_distance_float_jet(self.h, _p(P), C.c_int(P.shape[0]), _p(val), _p(grad))
        return val, grad

    def distance_loss(self, p):
        p = np.ascontiguousarray(p, dtype=np.float64)
        r = np.empty(3, np.float64); J = np.empty((3, 3), np.float64)
        lib().ref_distance_loss(self.h, _p(p), _p(r), _p(J))
        return r, J


def edge_loss(p1, p2, v, lam, adaptive=False):
    p1, p2, v = (np.ascontiguousarray(x, dtype=np.float64) for x in (p1, p2, v))
    r = np.empty(3, np.float64); le = C.c_double()
    lib().ref_edge_loss(_p(p1), _p(p2), _p(v), C.c_double(lam), C.c_int(int(adaptive)), _p(r), C.byref(le))
    return r, le.value


def edge_rot(p1, p2, rot1, rot2, v, lam):
    p1, p2, rot1, rot2, v = (np.ascontiguousarray(x, dtype=np.float64) for x in (p1, p2, rot1, rot2, v))
    r = np.empty(6, np.float64); J = np.empty((6, 12), np.float64)
    lib().ref_edge_rot(_p(p1), _p(p2), _p(rot1), _p(rot2), _p(v), C.c_double(lam), _p(r), _p(J))
    return r, J


class Params:
    """A reference DeformParams (src/interface/deform_params.h:7-25) as InitializeDeformTemplate leaves it -- grid,
    scale, translation -- driven through the reference's own interface loops (src/interface/*_layer.cc, normalize.cc,
    compiled where they lie; entry points in oracle/ref_iface_shim.cc)."""

    def __init__(self, grid, scale=1.0, trans=(0.0, 0.0, 0.0)):
        grid = np.ascontiguousarray(grid, dtype=np.float64)
        t = np.ascontiguousarray(trans, dtype=np.float64)
        self.id = int(lib().ref_params_create(C.c_int(grid.shape[0]), _p(grid), C.c_double(float(scale)), _p(t)))

    @staticmethod
    def _v(V):
        V = np.ascontiguousarray(V, dtype=np.float32)
        assert V.ndim == 2 and V.shape[1] == 3
        return V

    @staticmethod
    def _i(I, cols):
        I = np.ascontiguousarray(I, dtype=np.int32)
        assert I.ndim == 2 and I.shape[1] == cols
        return I

    def normalize(self, V, inverse=False):
        V = self._v(V).copy()
        lib().ref_iface_normalize(_p(V), C.c_int(V.shape[0]), C.c_int(self.id), C.c_int(int(inverse)))
        return V

    def dist_forward(self, V):
        V = self._v(V); out = np.empty(V.shape[0], dtype=np.float32)
        lib().ref_iface_dist_forward(_p(V), C.c_int(V.shape[0]), C.c_int(self.id), _p(out))
        return out

    def dist_backward(self, V):
        V = self._v(V); out = np.empty((V.shape[0], 3), dtype=np.float32)
        lib().ref_iface_dist_backward(_p(V), C.c_int(V.shape[0]), C.c_int(self.id), _p(out))
        return out

    def _edges(self, name, V, F, E, rows):
        V = self._v(V)
        args = [_p(V), C.c_int(V.shape[0])]
        if F is not None:
            F = self._i(F, 3); args += [_p(F), C.c_int(F.shape[0])]
        if E is not None:
            E = self._i(E, 2); args += [_p(E), C.c_int(E.shape[0])]
        args.append(C.c_int(self.id))
        out = None
        if rows is not None:
            out = np.empty((rows, 3), dtype=np.float32); args.append(_p(out))
        getattr(lib(), name)(*args)
        return out

    def rigid_store(self, V, F): self._edges("ref_iface_rigid_store", V, F, None, None)
    def rigid_forward(self, V, F): return self._edges("ref_iface_rigid_forward", V, F, None, 3 * len(F))
    def rigid_backward(self, V, F): return self._edges("ref_iface_rigid_backward", V, F, None, len(V))
    def graph_store(self, V, E): self._edges("ref_iface_graph_store", V, None, E, None)
    def graph_forward(self, V, E): return self._edges("ref_iface_graph_forward", V, None, E, len(E))
    def graph_backward(self, V, E): return self._edges("ref_iface_graph_backward", V, None, E, len(V))
    def cad_store(self, V, F, E): self._edges("ref_iface_cad_store", V, F, E, None)
    def cad_forward(self, V, F, E): return self._edges("ref_iface_cad_forward", V, F, E, len(E) + 3 * len(F))
    def cad_backward(self, V, F, E): return self._edges("ref_iface_cad_backward", V, F, E, len(V))

    def cad_lambda(self, n):
        f = lib().ref_iface_cad_lambda
        f.restype = C.c_float
        return np.array([f(C.c_int(self.id), C.c_int(i)) for i in range(n)], dtype=np.float32)
