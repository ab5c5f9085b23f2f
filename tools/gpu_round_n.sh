#!/bin/bash
# multi-GPU: sharded parity tests + bench at N ranks (both transports) + the reference arm.  usage: tools/gpu_round_n.sh <tag> <N>
tag=${1:-r02n}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -$N
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q > gpurun_out/pytest_sharded_$tag.log 2>&1; echo "exit $?" >> gpurun_out/pytest_sharded_$tag.log; tail -5 gpurun_out/pytest_sharded_$tag.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 60 --warmup 10 > gpurun_out/bench_${tag}_n$N.json 2> gpurun_out/bench_${tag}_n$N.err
echo "bench exit $?"; tail -c 2500 gpurun_out/bench_${tag}_n$N.json; tail -5 gpurun_out/bench_${tag}_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --config E --steps 30 --warmup 5 > gpurun_out/bench_${tag}_E_n$N.json 2> gpurun_out/bench_${tag}_E_n$N.err
tail -c 700 gpurun_out/bench_${tag}_E_n$N.json
