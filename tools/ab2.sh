#!/bin/bash
# stage times of build variants (EMAP_LIB) on configs B and D, no tests.  usage: tools/ab2.sh <tag> <variant>...
tag=$1; shift
mkdir -p gpurun_out
P=elevation_mapping_cupy_b200
for v in "" "$@"; do
  lib=$P/libemap.so; [ -n "$v" ] && lib=$P/libemap_$v.so
  for cfg in 1024 D; do
    EMAP_LIB=$PWD/$lib timeout 300 python tools/stage_times.py $cfg 2>&1 | tail -1 | tee -a gpurun_out/ab_$tag.txt
  done
done
