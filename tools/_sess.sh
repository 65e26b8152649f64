set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
timeout 200 python tools/sdf_bench.py 128 25002 8 > gpurun_out/sdf_bench.log 2>&1; timeout 100 python tools/sdf_bench.py 64 5000 8 >> gpurun_out/sdf_bench.log 2>&1; timeout 200 python tools/sdf_bench.py 256 250002 4 >> gpurun_out/sdf_bench.log 2>&1; cat gpurun_out/sdf_bench.log
