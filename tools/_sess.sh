set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_bench_v10.csv python bench.py --steps 1 --warmup 1 --pairs 148 --iters 1000 --no-cpu --no-sdf128 --no-percall > gpurun_out/launches_bench_v10.log 2>&1; echo "exit $?"
tail -2 gpurun_out/launches_bench_v10.log | cut -c1-300
