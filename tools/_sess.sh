set -u
mkdir -p gpurun_out
# 1. SDF variants: timing, then parity of each variant
rm -f gpurun_out/sdf_bench.log
for v in default compact compact16 premask premask_compact pc_ctas2; do
  if [ $v = default ]; then unset MESHODE_B200_LIB; else export MESHODE_B200_LIB=$PWD/build/variants/libmeshode_$v.so; fi
  echo "== $v" >> gpurun_out/sdf_bench.log
  timeout 200 python tools/sdf_bench.py 128 25002 8 >> gpurun_out/sdf_bench.log 2>&1; timeout 100 python tools/sdf_bench.py 64 5000 8 >> gpurun_out/sdf_bench.log 2>&1; timeout 200 python tools/sdf_bench.py 256 250002 4 >> gpurun_out/sdf_bench.log 2>&1
  if [ $v != default ]; then timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -m gpu 2>&1 | tail -3 >> gpurun_out/sdf_bench.log; fi
done
unset MESHODE_B200_LIB
cat gpurun_out/sdf_bench.log
# 2. the whole GPU suite (with the reference's own script where the call shipped it under /tmp/refpy)
if [ -f /tmp/refpy/rigid_deform.py ]; then export MESHODE_REFERENCE_PY=/tmp/refpy; (cd /tmp/refpy && sha256sum rigid_deform.py layers/rigid_loss_layer.py) > gpurun_out/reference_scripts.log; fi
timeout 1800 python -m pytest tests -q -m gpu --durations=8 -rA > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -12 gpurun_out/pytest_gpu.log
grep -E "PASSED|FAILED|SKIPPED" gpurun_out/pytest_gpu.log | grep -i "pydeform_ext" >> gpurun_out/reference_scripts.log
# 3. profiles
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sdf_tiles -s 1 -c 1 -f -o gpurun_out/prof_sdf128_r2b python tools/prof_target.py sdf128 > gpurun_out/ncu_sdf128.log 2>&1; tail -2 gpurun_out/ncu_sdf128.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_bench_r2.csv python bench.py --steps 1 --warmup 1 --pairs 148 --iters 1000 --no-cpu --no-sdf128 --no-percall > gpurun_out/launches_bench_r2.log 2>&1; tail -2 gpurun_out/launches_bench_r2.log | cut -c1-300
