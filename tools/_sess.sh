set -u
mkdir -p gpurun_out
rm -f gpurun_out/ab.log
for v in default ell8 default ell8; do
  if [ $v = default ]; then unset MESHODE_B200_LIB; else export MESHODE_B200_LIB=$PWD/build/variants/libmeshode_$v.so; fi
  echo "== $v" >> gpurun_out/ab.log
  MESHODE_EXACT=1 MESHODE_SCHEDULE=cta timeout 300 python tools/deform_bench.py 148 400 5000 >> gpurun_out/ab.log 2>&1
done
export MESHODE_B200_LIB=$PWD/build/variants/libmeshode_ell8.so
timeout 900 python -m pytest tests/test_gpu_deform.py -q -m gpu 2>&1 | tail -3 >> gpurun_out/ab.log
unset MESHODE_B200_LIB
cat gpurun_out/ab.log
