#!/bin/bash
# tests + bench lines of configs B, E, D on one GPU.  usage: tools/gpu_round2.sh <tag>
tag=${1:-r02d}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$tag.log
tail -5 gpurun_out/pytest_$tag.log
timeout 300 python tools/stage_times.py > gpurun_out/stage_$tag.txt 2>&1; cat gpurun_out/stage_$tag.txt
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err; tail -c 1500 gpurun_out/bench_${tag}_n1.json
timeout 600 python bench.py --config E --steps 30 --warmup 5 > gpurun_out/bench_${tag}_E_n1.json 2> gpurun_out/bench_${tag}_E_n1.err; tail -c 2500 gpurun_out/bench_${tag}_E_n1.json; tail -3 gpurun_out/bench_${tag}_E_n1.err
timeout 600 python bench.py --config E --no-graph --steps 30 --warmup 5 > gpurun_out/bench_${tag}_E_nograph_n1.json 2> /dev/null; tail -c 600 gpurun_out/bench_${tag}_E_nograph_n1.json
timeout 900 python bench.py --config D --steps 5 --warmup 3 > gpurun_out/bench_${tag}_D_n1.json 2> gpurun_out/bench_${tag}_D_n1.err; tail -c 2500 gpurun_out/bench_${tag}_D_n1.json; tail -3 gpurun_out/bench_${tag}_D_n1.err
