set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/pytest_sdf.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_sdf.log; tail -5 gpurun_out/pytest_sdf.log
rm -f gpurun_out/sdf_bench.log
for v in default ctas2 ctas4; do
  if [ $v = default ]; then unset MESHODE_B200_LIB; else export MESHODE_B200_LIB=$PWD/build/variants/libmeshode_$v.so; fi
  echo "== $v" >> gpurun_out/sdf_bench.log
  timeout 200 python tools/sdf_bench.py 128 25002 8 >> gpurun_out/sdf_bench.log 2>&1; timeout 100 python tools/sdf_bench.py 64 5000 8 >> gpurun_out/sdf_bench.log 2>&1; timeout 200 python tools/sdf_bench.py 256 250002 4 >> gpurun_out/sdf_bench.log 2>&1
done
cat gpurun_out/sdf_bench.log
unset MESHODE_B200_LIB
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sdf_tiles -s 1 -c 1 -f -o gpurun_out/prof_sdf128_r2 python tools/prof_target.py sdf128 > gpurun_out/ncu_sdf128.log 2>&1
tail -3 gpurun_out/ncu_sdf128.log
