set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_deform_adam -c 2 -f \
    -o gpurun_out/prof_deform_v12 python tools/prof_target.py deform 157 300 > gpurun_out/ncu_deform_v12.log 2>&1
tail -2 gpurun_out/ncu_deform_v12.log
