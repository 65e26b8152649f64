# scratch session script for gpurun (edited per experiment)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_deform.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
