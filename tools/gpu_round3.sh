#!/bin/bash
# tests + bench B + bench D + stage times.  usage: tools/gpu_round3.sh <tag>
tag=${1:-r02e}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$tag.log
tail -15 gpurun_out/pytest_$tag.log
timeout 300 python tools/stage_times.py > gpurun_out/stage_$tag.txt 2>&1; cat gpurun_out/stage_$tag.txt
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err; tail -c 800 gpurun_out/bench_${tag}_n1.json
timeout 900 python bench.py --config D --steps 5 --warmup 3 > gpurun_out/bench_${tag}_D_n1.json 2> gpurun_out/bench_${tag}_D_n1.err; tail -c 1200 gpurun_out/bench_${tag}_D_n1.json; tail -3 gpurun_out/bench_${tag}_D_n1.err
cat gpurun_out/inpaint_config_d.json
