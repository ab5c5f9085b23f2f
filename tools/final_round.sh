#!/bin/bash
# Round-end record on one GPU: tests, bench lines of configs B / D / E, launch list, full ncu capture, sanitizers.
# usage: tools/final_round.sh <tag>
tag=${1:-r02z}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$tag.log
tail -4 gpurun_out/pytest_$tag.log
timeout 400 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err; tail -c 600 gpurun_out/bench_${tag}_n1.json
timeout 300 python bench.py --config E --steps 20 --warmup 5 > gpurun_out/bench_${tag}_E_n1.json 2> gpurun_out/bench_${tag}_E_n1.err; tail -c 400 gpurun_out/bench_${tag}_E_n1.json
timeout 400 python bench.py --config D --steps 4 --warmup 3 > gpurun_out/bench_${tag}_D_n1.json 2> gpurun_out/bench_${tag}_D_n1.err; tail -c 900 gpurun_out/bench_${tag}_D_n1.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_${tag}_reference.json 2> /dev/null; tail -c 300 gpurun_out/bench_${tag}_reference.json
bash tools/evidence.sh $tag > gpurun_out/evidence_$tag.log 2>&1; tail -12 gpurun_out/evidence_$tag.log
