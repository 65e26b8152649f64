# scratch session script for gpurun (edited per experiment)
set -u
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sdf_tiles -s 1 -c 2 -f \
    -o gpurun_out/prof_sdf64_v8 python tools/prof_target.py sdf64 > gpurun_out/ncu_sdf64_v8.log 2>&1
tail -1 gpurun_out/ncu_sdf64_v8.log
timeout 120 python tools/setup_bench.py 592 4 | tail -2
