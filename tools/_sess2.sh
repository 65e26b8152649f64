set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi2.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/pytest_multi.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_multi.log; tail -15 gpurun_out/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 1 --pairs 310 --iters 1000 > gpurun_out/bench_n2_small.json 2> gpurun_out/bench_n2_small.err; echo "bench exit $?"; cat gpurun_out/bench_n2_small.json; tail -20 gpurun_out/bench_n2_small.err
