set -u
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['share_of_step'], d['cpu_baseline']['value'], d['clocks'])
print(d['sdf_build_128']['value'], d['sdf_build_128']['roofline']['frac'], d['per_call_path']['us_per_iteration'], d['large_mesh_path']['us_per_iteration'])
P
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_deform_adam -c 2 -f \
    -o gpurun_out/prof_deform_v11 python tools/prof_target.py deform 157 300 > gpurun_out/ncu_deform_v11.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sdf_tiles -s 1 -c 1 -f \
    -o gpurun_out/prof_sdf128_v8 python tools/prof_target.py sdf128 > gpurun_out/ncu_sdf128_v8.log 2>&1
