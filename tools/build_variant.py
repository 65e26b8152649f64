"""Builds a variant of libmeshode_b200.so with extra nvcc flags for A/B timing inside one GPU session:
  python tools/build_variant.py NAME -DMO_SDF_CTAS=2 ...   ->  build/variants/libmeshode_NAME.so
Use it with MESHODE_B200_LIB=build/variants/libmeshode_NAME.so.  Development tool, not part of the product."""
import concurrent.futures
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from meshode_b200 import build as B  # noqa: E402

name, extra = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(ROOT, "build", "variants")
obj_dir = os.path.join(out_dir, "obj_" + name)
os.makedirs(obj_dir, exist_ok=True)


def one(f):
    obj = os.path.join(obj_dir, f[:-3] + ".o")
    subprocess.check_call([B._nvcc()] + B.NVCC_FLAGS + extra + ["-c", os.path.join(B.CSRC, f), "-o", obj])
    return obj


with concurrent.futures.ThreadPoolExecutor(8) as ex:
    objs = list(ex.map(one, B.SOURCES))
lib = os.path.join(out_dir, "libmeshode_%s.so" % name)
subprocess.check_call([B._nvcc()] + B.NVCC_FLAGS[:2] + ["-shared", "-o", lib] + objs)
print(lib)
