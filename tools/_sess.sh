set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
export MESHODE_EXACT=1 MESHODE_SCHEDULE=cta
timeout 120 python tools/deform_bench.py 148 400 5000
timeout 120 python tools/deform_bench.py 148 400 5000
unset MESHODE_EXACT MESHODE_SCHEDULE
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_deform_adam -c 2 -f \
    -o gpurun_out/prof_deform_v10 python tools/prof_target.py deform 157 300 > gpurun_out/ncu_deform_v10.log 2>&1
timeout 400 compute-sanitizer --tool memcheck python tools/sanitize_target.py > gpurun_out/sanitize_memcheck_v10.log 2>&1; tail -4 gpurun_out/sanitize_memcheck_v10.log
