"""Build-time guard for the persistent deformation kernels: they run with ~205 KB of shared memory carved out, so L1
keeps ~28 KB and a register spill is an L2 round trip (a 1024 x 64 build that spilled three registers ran 13 % slower).
ptxas must fit every instantiation without spills."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="nvcc not available")
def test_deformation_kernels_compile_without_spills(tmp_path):
    from meshode_b200 import build as B
    out = subprocess.run([B._nvcc()] + B.NVCC_FLAGS + ["-Xptxas", "-v", "-cubin", "-o", str(tmp_path / "deform.cubin"),
                          os.path.join(B.CSRC, "deform.cu")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-2000:]
    blocks = re.findall(r"Function properties for (\S+)\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads",
                        out.stderr)
    seen = 0
    for name, _stack, st, ld in blocks:
        if "k_deform_adam_fused2" in name or "k_deform_adam_cluster" in name:
            seen += 1
            assert int(st) == 0 and int(ld) == 0, "%s spills (%s B stores, %s B loads)" % (name, st, ld)
    assert seen >= 5, "expected three adjacency widths of the fused loop and two of the cluster loop, found %d" % seen
