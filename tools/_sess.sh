set -u
mkdir -p gpurun_out
for rep in 1 2; do
echo "== old"; for c in "128 25002" "64 5000" "256 250002"; do MESHODE_B200_LIB=build/variants/libmeshode_sdfold.so timeout 120 python tools/sdf_bench.py $c 6; done
echo "== new"; for c in "128 25002" "64 5000" "256 250002"; do timeout 120 python tools/sdf_bench.py $c 6; done
done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3
