"""In-tree build of libmeshode_b200.so (nvcc, sm_100a only) and of the pyDeform extension."""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmeshode_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
SOURCES = ["sdf_build.cu", "sampler.cu", "edges.cu", "capi.cu", "deform.cu", "microbench.cu", "nearest.cu", "ceres_path.cu"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force=False, verbose=False):
    """Compile meshode_b200/csrc/*.cu into meshode_b200/libmeshode_b200.so. Returns the path."""
    srcs = [os.path.join(CSRC, f) for f in SOURCES if os.path.exists(os.path.join(CSRC, f))]
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "meshode_b200.h"))
    objdir = os.path.join(HERE, "..", "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if force or _newer(obj, [src] + hdrs):
            cmd = [nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            subprocess.check_call(cmd)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    if force or _newer(LIB, objs):
        subprocess.check_call([nvcc] + NVCC_FLAGS[:2] + ["-shared", "-o", LIB] + objs)
    return LIB


APPS = ["rigid_deform", "rigid_rot_deform", "cad_deform"]
BIN = os.path.join(HERE, "bin")


def build_apps(force=False):
    """The re-hosted C++ drivers (apps/*.cc, reference src/app/*.cc) linked against the C-ABI library."""
    build_lib()
    os.makedirs(BIN, exist_ok=True)
    appdir = os.path.join(HERE, "..", "apps")
    deps = [os.path.join(appdir, f) for f in os.listdir(appdir) if f.endswith(".h")] + [LIB]
    out = []
    for a in APPS:
        exe = os.path.join(BIN, a)
        src = os.path.join(appdir, a + ".cc")
        if force or _newer(exe, [src] + deps):
            subprocess.check_call([_nvcc(), "-O2", "-std=c++17", "-x", "cu", *NVCC_FLAGS[:2], src, "-o", exe, "-L" + HERE,
                                   "-lmeshode_b200", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/.."])
        out.append(exe)
    return out


EXT_DIR = os.path.join(HERE, "ext")


def ext_path():
    import sysconfig
    return os.path.join(EXT_DIR, "pyDeform" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_ext(force=False, verbose=False):
    """The compiled ``pyDeform`` module (csrc/pydeform_ext.cpp: pybind11 + libtorch over the C-ABI), in-tree at
    meshode_b200/ext/pyDeform.<abi>.so -- put that directory on PYTHONPATH to get it as ``import pyDeform``, the
    way the reference's scripts find their build/ directory.  Compiled with g++ directly (torch's headers, the
    pybind11 ABI tags of the running torch); libtorch is resolved from the already imported torch at load time,
    which is the reference's "import torch first" rule (README.md:46-50)."""
    build_lib()
    import sysconfig

    import torch
    from torch.utils import cpp_extension as ce
    out = ext_path()
    src = os.path.join(CSRC, "pydeform_ext.cpp")
    deps = [src, os.path.join(HERE, "..", "include", "meshode_b200.h"), os.path.join(HERE, "..", "apps", "mesh_host.h")]
    if not (force or _newer(out, deps)):
        return out
    os.makedirs(EXT_DIR, exist_ok=True)
    inc = ce.include_paths() + [sysconfig.get_paths()["include"], "/usr/local/cuda/include"]
    abi = []
    for name in ("COMPILER_TYPE", "STDLIB", "BUILD_ABI"):
        val = getattr(torch._C, "_PYBIND11_" + name, None)
        if val is not None:
            abi.append('-DPYBIND11_%s="%s"' % (name, val))
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = (["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-DTORCH_EXTENSION_NAME=pyDeform",
            "-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)] + abi +
           ["-I" + i for i in inc] + [src, "-o", out, "-L" + tlib, "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lc10",
                                      "-lc10_cuda", "-L" + HERE, "-lmeshode_b200", "-Wl,-rpath,$ORIGIN/.."])
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose=True))
    print(build_apps(force="--force" in sys.argv))
    print(build_ext(force="--force" in sys.argv, verbose=True))
