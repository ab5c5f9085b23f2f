#!/bin/bash
# One gpurun call: GPU tests, stage times, bench line, ncu launch list and full captures of the frame kernels.
# usage: tools/gpu_round.sh <tag> [skip-tests]
tag=${1:-r02a}
mkdir -p gpurun_out
if [ "$2" != "skip-tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$tag.log
  tail -5 gpurun_out/pytest_$tag.log
fi
timeout 300 python tools/stage_times.py > gpurun_out/stage_$tag.txt 2>&1; cat gpurun_out/stage_$tag.txt
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err; tail -c 3000 gpurun_out/bench_${tag}_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file gpurun_out/launches_$tag.csv python tools/stage_times.py > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_raycast|k_post|k_finalize|k_record|k_fuse|k_index' -s 60 -c 12 -o gpurun_out/prof_$tag -f python tools/stage_times.py > gpurun_out/ncu_$tag.log 2>&1
ls -la gpurun_out | tail -8
