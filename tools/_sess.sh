set -u
mkdir -p gpurun_out
rm -f gpurun_out/deform_bench.log
for v in default zpack zdb; do
  if [ $v = default ]; then unset MESHODE_B200_LIB; else export MESHODE_B200_LIB=$PWD/build/variants/libmeshode_$v.so; fi
  echo "== $v" >> gpurun_out/deform_bench.log
  MESHODE_EXACT=1 MESHODE_SCHEDULE=cta timeout 300 python tools/deform_bench.py 148 400 5000 >> gpurun_out/deform_bench.log 2>&1
  MESHODE_EXACT=1 MESHODE_SCHEDULE=cta timeout 300 python tools/deform_bench.py 148 400 3000 >> gpurun_out/deform_bench.log 2>&1
  MESHODE_EXACT=1 MESHODE_SCHEDULE=auto timeout 300 python tools/deform_bench.py 453 400 5000 >> gpurun_out/deform_bench.log 2>&1
  if [ $v = zdb ]; then timeout 900 python -m pytest tests/test_gpu_deform.py -q -m gpu 2>&1 | tail -3 >> gpurun_out/deform_bench.log; fi
done
unset MESHODE_B200_LIB
cat gpurun_out/deform_bench.log
