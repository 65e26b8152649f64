set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_deform.py -x -q -m gpu > gpurun_out/pytest_deform.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_deform.log; tail -15 gpurun_out/pytest_deform.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
for cfg in "cta 148" "cluster 148" "cluster 29" "cluster 9" "cluster 1" "cta 1" "auto 157" "cta 157" "auto 453" "cta 453"; do
  set -- $cfg
  MESHODE_EXACT=1 MESHODE_SCHEDULE=$1 timeout 300 python tools/deform_bench.py $2 400 5000 >> gpurun_out/deform_bench.log 2>&1
done
cat gpurun_out/deform_bench.log
