set -u
mkdir -p gpurun_out
n=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 2 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; echo "bench n=$n exit $?"; grep '^{' gpurun_out/bench_n$n.json | cut -c1-200; tail -2 gpurun_out/bench_n$n.err
