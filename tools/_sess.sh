set -u
mkdir -p gpurun_out
rm -f gpurun_out/ab.log
for v in default d13; do
  if [ $v = default ]; then unset MESHODE_B200_LIB; else export MESHODE_B200_LIB=$PWD/build/variants/libmeshode_$v.so; fi
  echo "== $v" >> gpurun_out/ab.log
  MESHODE_EXACT=1 MESHODE_SCHEDULE=cta timeout 300 python tools/deform_bench.py 148 400 5000 >> gpurun_out/ab.log 2>&1
  MESHODE_EXACT=1 MESHODE_SCHEDULE=cta timeout 300 python tools/deform_bench.py 148 400 5000 >> gpurun_out/ab.log 2>&1
  if [ $v != default ]; then timeout 900 python -m pytest tests/test_gpu_deform.py -q -m gpu 2>&1 | tail -3 >> gpurun_out/ab.log; fi
done
for v in default directed; do
  if [ $v = default ]; then unset MESHODE_B200_LIB; else export MESHODE_B200_LIB=$PWD/build/variants/libmeshode_$v.so; fi
  echo "== $v" >> gpurun_out/ab.log
  timeout 200 python tools/sdf_bench.py 128 25002 8 >> gpurun_out/ab.log 2>&1; timeout 100 python tools/sdf_bench.py 64 5000 8 >> gpurun_out/ab.log 2>&1; timeout 200 python tools/sdf_bench.py 256 250002 4 >> gpurun_out/ab.log 2>&1
  if [ $v != default ]; then timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -m gpu 2>&1 | tail -3 >> gpurun_out/ab.log; fi
done
unset MESHODE_B200_LIB
cat gpurun_out/ab.log
