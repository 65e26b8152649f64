set -u
mkdir -p gpurun_out
export MESHODE_EXACT=1
MESHODE_SCHEDULE=cluster timeout 120 python tools/deform_bench.py 9 400 5000
MESHODE_SCHEDULE=cluster timeout 120 python tools/deform_bench.py 29 400 5000
MESHODE_SCHEDULE=cluster timeout 120 python tools/deform_bench.py 1 400 5000
MESHODE_SCHEDULE=auto timeout 120 python tools/deform_bench.py 157 400 5000
MESHODE_SCHEDULE=auto timeout 120 python tools/deform_bench.py 453 400 5000
MESHODE_SCHEDULE=cta timeout 120 python tools/deform_bench.py 453 400 5000
unset MESHODE_EXACT
timeout 600 python -m pytest tests/test_gpu_deform.py -x -q 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
