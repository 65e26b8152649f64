set -u
mkdir -p gpurun_out
export MESHODE_EXACT=1 MESHODE_SCHEDULE=cta
for rep in 1 2; do
echo "== before"; MESHODE_B200_LIB=build/variants/libmeshode_noreuse3.so timeout 120 python tools/deform_bench.py 148 400 5000
echo "== reuse3"; timeout 120 python tools/deform_bench.py 148 400 5000
done
unset MESHODE_EXACT MESHODE_SCHEDULE
timeout 600 python -m pytest tests/test_gpu_deform.py -x -q 2>&1 | tail -5
