#!/bin/bash
# Runs the reference's UNMODIFIED src/python/rigid_deform.py (with the reference's own layers/) on a B200 on top of this
# repository's pyDeform (tests/test_pydeform_ext.py::test_reference_rigid_deform_script_unmodified).
# /root/reference does not exist on the GPU box and reference sources are never copied into this repository, so the
# few Python files travel INSIDE the gpurun command line (a base64 tarball unpacked under /tmp on the box).
# usage: tools/run_reference_scripts.sh [LOGFILE]        (from the repository root, in the build container)
set -eu
log=${1:-gpurun_out/reference_scripts_call.log}
ref=/root/reference/src/python
blob=$(tar -C "$ref" -czf - rigid_deform.py layers/rigid_loss_layer.py | base64 -w0)
sums=$(cd "$ref" && sha256sum rigid_deform.py layers/rigid_loss_layer.py | tr '\n' ';')
cmd="mkdir -p /tmp/refpy gpurun_out && echo $blob | base64 -d | tar -C /tmp/refpy -xzf - && (cd /tmp/refpy && sha256sum rigid_deform.py layers/rigid_loss_layer.py) > gpurun_out/reference_scripts.log && echo 'expected: $sums' >> gpurun_out/reference_scripts.log && MESHODE_REFERENCE_PY=/tmp/refpy timeout 1500 python -m pytest tests/test_pydeform_ext.py -q -m gpu -rA >> gpurun_out/reference_scripts.log 2>&1; tail -15 gpurun_out/reference_scripts.log"
exec "$(dirname "$0")/gpurun_retry.sh" "$log" --timeout 1800 -- "$cmd"
