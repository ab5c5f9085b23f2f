#!/bin/bash
# A/B of build variants (EMAP_LIB) on one GPU: smoke (bounded), tests with the product library, then stage times of each
# variant on configs B and D, then the bench line.  usage: tools/ab.sh <tag> <variant>...
tag=$1; shift
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1 || { echo "SMOKE FAILED"; tail -5 gpurun_out/smoke_$tag.log; exit 1; }
tail -1 gpurun_out/smoke_$tag.log
timeout 200 python tools/stage_times.py 1024 2>&1 | tail -1 | tee -a gpurun_out/ab_$tag.txt || { echo "STAGE FAILED"; exit 1; }
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$tag.log
tail -6 gpurun_out/pytest_$tag.log
P=elevation_mapping_cupy_b200
for v in "" "$@"; do
  lib=$P/libemap.so; [ -n "$v" ] && lib=$P/libemap_$v.so
  for cfg in 1024 D; do
    EMAP_LIB=$PWD/$lib timeout 300 python tools/stage_times.py $cfg 2>&1 | tail -1 | tee -a gpurun_out/ab_$tag.txt
  done
done
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err; tail -c 1200 gpurun_out/bench_${tag}_n1.json
