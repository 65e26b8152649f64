set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_deform.py -q -m gpu -rA 2>&1 | tail -30 | cut -c1-200
