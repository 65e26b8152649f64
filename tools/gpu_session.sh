#!/bin/bash
# One gpurun call: GPU parity tests, the default bench, launch lists and ncu captures.
#   gpurun --timeout 1500 -- 'bash tools/gpu_session.sh [tests] [bench] [lists] [ncu]'
# (no argument = everything).  Outputs land in gpurun_out/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
what="${*:-tests bench lists ncu}"
has() { [[ " $what " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -c "import os; print('host cores', os.cpu_count())" >> gpurun_out/smi.txt

if has tests; then
  timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
  tail -5 gpurun_out/pytest_gpu.log
fi
if has smoke; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
  echo "smoke exit $?" >> gpurun_out/smoke.log
  tail -3 gpurun_out/smoke.log
fi
if has bench; then
  timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench exit $?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
fi
if has benchref; then
  timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
  cat gpurun_out/bench_ref.json
fi
if has lists; then
  # launch list of the bench command itself at a reduced pair / iteration count (shares, not absolutes)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
    --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --pairs 148 --iters 1000 --no-cpu \
    > gpurun_out/launches_bench.log 2>&1
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_sdf128.csv python tools/prof_target.py sdf128 > gpurun_out/launches_sdf128.log 2>&1
fi
if has ncu; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sdf_tiles -s 1 -c 1 -f \
    -o gpurun_out/prof_sdf128 python tools/prof_target.py sdf128 > gpurun_out/ncu_sdf128.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_deform_adam -c 1 -f \
    -o gpurun_out/prof_deform python tools/prof_target.py deform 148 300 > gpurun_out/ncu_deform.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_distance_f32|k_edges|k_loss_fused' -c 12 -f \
    -o gpurun_out/prof_loss python tools/prof_target.py loss > gpurun_out/ncu_loss.log 2>&1
fi
ls -la gpurun_out
