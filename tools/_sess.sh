set -u
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_target.py > gpurun_out/sanitize_racecheck.log 2>&1; tail -4 gpurun_out/sanitize_racecheck.log
timeout 900 compute-sanitizer --tool synccheck python tools/sanitize_target.py > gpurun_out/sanitize_synccheck.log 2>&1; tail -3 gpurun_out/sanitize_synccheck.log
timeout 600 python -m pytest tests/test_gpu_deform.py -q -m gpu 2>&1 | tail -2
