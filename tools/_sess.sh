set -u
mkdir -p gpurun_out
timeout 600 python bench.py --pairs 148 --iters 500 --steps 1 --warmup 3 --no-cpu > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "exit $?"; python - <<'P'
import json
d=json.load(open('gpurun_out/bench_small.json'))
print(d['value'], d['sdf_build_128']['value'], d['sdf_build_128']['roofline'])
P
tail -3 gpurun_out/bench_small.err
