set -u
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['share_of_step'], d['cpu_baseline']['value'], d['clocks'])
print(d['sdf_build_128']['value'], d['per_call_path']['us_per_iteration'], d['large_mesh_path']['us_per_iteration'])
P
tail -3 gpurun_out/bench_n1.err
timeout 600 python -m pytest tests/test_gpu_apps.py -x -q 2>&1 | tail -3
