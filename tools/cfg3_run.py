"""cfg3 of BASELINE.json end to end: rigid_rot_deform on a synthetic 50 000-vertex source and a
50 000-triangle target at GRID_RESOLUTION=128 (python tools/cfg3_run.py [n_src] [n_tar_verts] [grid])."""
import os
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshode_b200 import build  # noqa: E402
from meshode_b200.synth import synth_mesh, synth_params  # noqa: E402

n_src = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
n_tar = int(sys.argv[2]) if len(sys.argv) > 2 else 25002
grid = int(sys.argv[3]) if len(sys.argv) > 3 else 128
name = sys.argv[4] if len(sys.argv) > 4 else "rigid_rot_deform"


def write_obj(path, V, F):
    with open(path, "w") as fh:
        fh.write("".join("v %.9g %.9g %.9g\n" % tuple(v) for v in V.tolist()))
        fh.write("".join("f %d %d %d\n" % (f[0] + 1, f[1] + 1, f[2] + 1) for f in F.tolist()))


exe = dict(zip(build.APPS, build.build_apps()))[name]
tV, tF = synth_mesh(n_tar, 1)
sV, sF = synth_mesh(n_src, 0, axis_scale=synth_params(1)[3])
with tempfile.TemporaryDirectory() as d:
    s, t, o = (os.path.join(d, x) for x in ("s.obj", "t.obj", "o.obj"))
    write_obj(s, sV, sF); write_obj(t, tV, tF)
    t0 = time.perf_counter()
    p = subprocess.run([exe, s, t, o, str(grid), "5000", "1"], capture_output=True, text=True)
    dt = time.perf_counter() - t0
    lines = p.stdout.splitlines()
    print("\n".join(lines[:6] + ["..."] + lines[-8:]))
    print(p.stderr[-500:])
    print("%s: %d source vertices, %d target triangles, grid %d: %.2f s wall (process start, OBJ parse, build, solve, OBJ write)"
          % (name, sV.shape[0], tF.shape[0], grid, dt))
