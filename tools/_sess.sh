set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_deform_adam -c 2 -f -o gpurun_out/prof_deform_r2b python tools/prof_target.py deform 157 300 > gpurun_out/ncu_deform.log 2>&1; tail -2 gpurun_out/ncu_deform.log
timeout 900 python -m pytest tests/test_gpu_deform.py tests/test_gpu_apps.py -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py --no-cpu > gpurun_out/bench_n1_nocpu.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench_n1_nocpu.json
