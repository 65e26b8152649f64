"""ctypes binding of libmeshode_b200.so (include/meshode_b200.h).

Every function takes raw device addresses (ints, e.g. ``tensor.data_ptr()``) and a CUDA
stream handle.  There is no CPU fallback: if the library has not been built, or no CUDA
device is visible, the calls raise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MESHODE_B200_LIB: another build of the same library (A/B timing of kernel variants inside one GPU session)
LIB_PATH = os.environ.get("MESHODE_B200_LIB") or os.path.join(_HERE, "libmeshode_b200.so")

MO_OK = 0
EDGES_RIGID, EDGES_GRAPH, EDGES_CAD = 0, 1, 2
CERES_EDGE, CERES_ADAPTIVE_EDGE, CERES_ROT_EDGE = 0, 1, 2
DEFORM_EXACT, DEFORM_CTA_ONLY, DEFORM_CLUSTER_ONLY = 1, 2, 4
LAYER_SLICES = 4   # MO_LAYER_SLICES: voxel slices per z-tile layer of the cyclic sharded build

_vp, _i, _d, _f = C.c_void_p, C.c_int, C.c_double, C.c_float
_ip, _dp, _ullp = C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_ulonglong)

# name -> argtypes; every function returns int unless listed in _RESTYPES
SIGNATURES = {
    "mo_version": [],
    "mo_last_error": [],
    "mo_device_count": [],
    "mo_launch_count": [],
    "mo_build_stats_enable": [_i],
    "mo_microbench_fp32": [_i, _i, _i, _vp, _vp],
    "mo_microbench_fp32x2": [_i, _i, _i, _vp, _vp],
    "mo_template_create": [_vp, _i, _vp, _i, _i, _i, _vp, _ip],
    "mo_template_create_slab": [_vp, _i, _vp, _i, _i, _i, _i, _vp, _ip],
    "mo_template_create_layers": [_vp, _i, _vp, _i, _i, _i, _i, _vp, _ip],
    "mo_template_create_normalized": [_vp, _i, _vp, _i, _i, _d, _dp, _vp, _ip],
    "mo_template_destroy": [_i],
    "mo_template_destroy_async": [_i, _vp],
    "mo_template_info": [_i, _vp, _ip, _ip, _ip, _dp, _dp],
    "mo_template_grid": [_i, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)],
    "mo_template_copy_grid": [_i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "mo_template_vertices": [_i, C.POINTER(_vp)],
    "mo_template_build_stats": [_i, _vp, _ullp, _ullp, _ullp, _ullp],
    "mo_normalize_by_template": [_vp, _i, _i, _i, _vp],
    "mo_distance_forward": [_vp, _i, _i, _vp, _vp],
    "mo_distance_backward": [_vp, _i, _i, _vp, _vp],
    "mo_distance_forward_backward": [_vp, _i, _i, _vp, _vp, _vp],
    "mo_distance_f64": [_vp, _i, _i, _vp, _vp, _vp],
    "mo_edges_store": [_i, _i, _vp, _i, _vp, _i, _vp, _i, _vp],
    "mo_edges_forward": [_i, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp],
    "mo_edges_backward": [_i, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp],
    "mo_edges_backward_atomic": [_i, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp],
    "mo_loss_forward_backward": [_i, _i, _vp, _i, _f, _f, _vp, _vp, _vp],
    "mo_deform_batch_adam": [_ip, _ip, C.POINTER(_vp), _i, _i, _d, _d, _d, _d, _i, _vp],
    "mo_deform_adam_large": [_i, _i, _vp, _i, _f, _f, _i, _d, _d, _d, _d, _vp],
    "mo_nearest_vertex": [_vp, _i, _vp, _i, _vp, _vp, _vp],
    "mo_ceres_edges": [_i, _vp, _vp, _i, _vp, _vp, _i, _d, _vp, _vp, _vp],
    "mo_ceres_problem": [_i, _i, _vp, _vp, _i, _vp, _vp, _i, _d, _vp, _vp, _vp, _vp],
    "mo_ceres_solve": [_i, _i, _vp, _vp, _i, _vp, _vp, _i, _d, _i, _i, _d, _i, _vp, _vp],
}
_RESTYPES = {"mo_last_error": C.c_char_p, "mo_launch_count": C.c_ulonglong}

_lib = None


class MeshodeError(RuntimeError):
    pass


def lib():
    """The loaded shared library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MeshodeError(
                "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(meshode_b200 has no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(L, name, None)
            if fn is None:
                continue
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, C.c_int)
        _lib = L
    return _lib


def check(rc):
    if rc != MO_OK:
        msg = lib().mo_last_error()
        raise MeshodeError("libmeshode_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))


def device_count():
    return lib().mo_device_count()


def require_device():
    if device_count() <= 0:
        raise MeshodeError("no CUDA device visible: meshode_b200 runs on the GPU only (no CPU fallback)")


# ---- thin typed wrappers (addresses are python ints; 0/None = NULL) ---------------------
def template_create(dV, nV, dF, nF, symmetry, N, stream=0):
    pid = C.c_int(-1)
    check(lib().mo_template_create(dV, nV, dF, nF, int(symmetry), int(N), stream, C.byref(pid)))
    return pid.value


def template_create_slab(dV, nV, dF, nF, N, z0, z1, stream=0):
    pid = C.c_int(-1)
    check(lib().mo_template_create_slab(dV, nV, dF, nF, int(N), int(z0), int(z1), stream, C.byref(pid)))
    return pid.value


def template_create_layers(dV, nV, dF, nF, N, first_layer, layer_stride, stream=0):
    pid = C.c_int(-1)
    check(lib().mo_template_create_layers(dV, nV, dF, nF, int(N), int(first_layer), int(layer_stride), stream, C.byref(pid)))
    return pid.value


def template_create_normalized(dVn, nV, dF, nF, N, scale, trans, stream=0):
    pid = C.c_int(-1)
    t = (C.c_double * 3)(*[float(x) for x in trans])
    check(lib().mo_template_create_normalized(dVn, nV, dF, nF, int(N), float(scale), t, stream, C.byref(pid)))
    return pid.value


def template_destroy(pid, stream=None):
    """stream=None: waits for the device, then frees; otherwise frees in the order of ``stream``."""
    if stream is None:
        check(lib().mo_template_destroy(int(pid)))
    else:
        check(lib().mo_template_destroy_async(int(pid), stream))


def template_info(pid, stream=0):
    N, nV, nF = C.c_int(), C.c_int(), C.c_int()
    scale = C.c_double()
    trans = (C.c_double * 3)()
    check(lib().mo_template_info(int(pid), stream, C.byref(N), C.byref(nV), C.byref(nF), C.byref(scale), trans))
    return {"N": N.value, "nV": nV.value, "nF": nF.value, "scale": scale.value, "trans": [trans[0], trans[1], trans[2]]}


def template_grid(pid):
    g64, g32, idx = _vp(), _vp(), _vp()
    check(lib().mo_template_grid(int(pid), C.byref(g64), C.byref(g32), C.byref(idx)))
    return g64.value, g32.value, idx.value


def template_vertices(pid):
    p = _vp()
    check(lib().mo_template_vertices(int(pid), C.byref(p)))
    return p.value


def template_build_stats(pid, stream=0):
    a, b, c, d = C.c_ulonglong(), C.c_ulonglong(), C.c_ulonglong(), C.c_ulonglong()
    check(lib().mo_template_build_stats(int(pid), stream, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
    return {"fp32_tests": a.value, "fp64_tests": b.value, "cull_tests": c.value, "disc_tests": d.value}
