#!/bin/bash
# lean multi-GPU record: bench config B (sharded, parity-checked) and config E at N ranks.  usage: tools/gpu_round_n8.sh <tag> <N>
tag=${1:-r02n}; N=${2:-8}
mkdir -p gpurun_out
nvidia-smi -L | head -$N
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${tag}_n$N.json 2> gpurun_out/bench_${tag}_n$N.err
echo "bench exit $?"; tail -c 1800 gpurun_out/bench_${tag}_n$N.json; tail -3 gpurun_out/bench_${tag}_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --config E --steps 20 --warmup 5 > gpurun_out/bench_${tag}_E_n$N.json 2> gpurun_out/bench_${tag}_E_n$N.err
echo "bench E exit $?"; tail -c 600 gpurun_out/bench_${tag}_E_n$N.json
