set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -10 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-200 gpurun_out/bench_ref.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_deform_adam -c 2 -f -o gpurun_out/prof_deform_r2c python tools/prof_target.py deform 157 300 > gpurun_out/ncu_deform.log 2>&1; tail -2 gpurun_out/ncu_deform.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sdf_tiles -s 1 -c 1 -f -o gpurun_out/prof_sdf128_r2c python tools/prof_target.py sdf128 > gpurun_out/ncu_sdf128.log 2>&1; tail -2 gpurun_out/ncu_sdf128.log
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_target.py > gpurun_out/sanitize_memcheck.log 2>&1; tail -4 gpurun_out/sanitize_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_target.py > gpurun_out/sanitize_racecheck.log 2>&1; tail -4 gpurun_out/sanitize_racecheck.log
