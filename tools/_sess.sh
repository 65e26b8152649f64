set -u
mkdir -p gpurun_out
export MESHODE_EXACT=1 MESHODE_SCHEDULE=cta
echo "== v1"; MESHODE_FUSED_V1=1 python tools/deform_bench.py 148 400 5000
echo "== v2 896"; python tools/deform_bench.py 148 400 5000
echo "== v2 1024"; MESHODE_B200_LIB=build/variants/libmeshode_nt1024.so python tools/deform_bench.py 148 400 5000
echo "== v2 768"; MESHODE_B200_LIB=build/variants/libmeshode_nt768.so python tools/deform_bench.py 148 400 5000
echo "== v1"; MESHODE_FUSED_V1=1 python tools/deform_bench.py 148 400 5000
echo "== v2 896"; python tools/deform_bench.py 148 400 5000
unset MESHODE_EXACT MESHODE_SCHEDULE
timeout 600 python -m pytest tests/test_gpu_deform.py -x -q 2>&1 | tail -5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_deform_adam_fused2 -c 1 -f \
    -o gpurun_out/prof_deform_v2c python tools/prof_target.py deform 148 300 > gpurun_out/ncu_deform_v2c.log 2>&1
