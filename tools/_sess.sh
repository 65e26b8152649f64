set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi.txt 2>&1
rm -f gpurun_out/sdf_bench.log
for v in default old c8s0 c4s1; do
  if [ $v = default ]; then unset MESHODE_B200_LIB; else export MESHODE_B200_LIB=$PWD/build/variants/libmeshode_$v.so; fi
  echo "== $v" >> gpurun_out/sdf_bench.log
  timeout 200 python tools/sdf_bench.py 128 25002 8 >> gpurun_out/sdf_bench.log 2>&1; timeout 100 python tools/sdf_bench.py 64 5000 8 >> gpurun_out/sdf_bench.log 2>&1; timeout 200 python tools/sdf_bench.py 256 250002 4 >> gpurun_out/sdf_bench.log 2>&1
done
unset MESHODE_B200_LIB
cat gpurun_out/sdf_bench.log
timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; cat gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
