#!/bin/bash
# Evidence run (one GPU): compute-sanitizer memcheck / racecheck / initcheck of the smoke frames, atomic / reduction counters and
# full captures of the frame kernels, SASS excerpts.  usage: tools/evidence.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
cat > /tmp/smoke_small.py <<'PY'
import sys; sys.path.insert(0, '.')
import __graft_entry__ as g
g.smoke()
# plugin stencils + inpainting + semantic fusion + batched export on a small map (sanitizer coverage)
import numpy as np
from elevation_mapping_cupy_b200.parameter import core_parameter
from elevation_mapping_cupy_b200.elevation_mapping import ElevationMap
from elevation_mapping_cupy_b200 import workloads as wl
p = core_parameter(130)
em = ElevationMap(p)
pts, R, t = wl.uniform_cloud(0, 1, n=4000, half_extent=2.2)
cloud = np.concatenate([pts, np.random.default_rng(0).random((len(pts), 2), dtype=np.float32)], 1)
em.input_pointcloud(cloud, ["x", "y", "z", "a", "rgb"], R, t, 0.02, 0.02)
import os
layers = ["elevation", "traversability", "min_filter", "smooth", "inpaint", "erosion"]
if os.environ.get("SKIP_INPAINT"):       # initcheck slows the cooperative march down by > 1000x: covered by memcheck / racecheck only
    layers.remove("inpaint")
out = np.zeros((len(layers), 128, 128), np.float32)
em.get_maps_with_names_ref(layers, out)
em.move(np.array([0.1, 0.0, 0.0])); em.update_variance(); em.update_time()
print("sanitizer workload ok")
PY
for tool in memcheck racecheck initcheck; do
  skip=""; [ "$tool" = "initcheck" ] && skip=1
  SKIP_INPAINT=$skip timeout 240 compute-sanitizer --tool $tool --log-file gpurun_out/sanitizer_${tool}_$tag.log python /tmp/smoke_small.py > gpurun_out/sanitizer_${tool}_$tag.out 2>&1
  echo "$tool exit $?"; tail -3 gpurun_out/sanitizer_${tool}_$tag.log
done
# atomic / reduction counters of the frame kernels (north_star: "atomic contention counters")
timeout 900 ncu --clock-control none -k regex:'k_raycast|k_post|k_finalize|k_record|k_fuse|k_index|k_drift' -s 56 -c 7 \
  --metrics gpu__time_duration.sum,lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,lts__t_requests_op_red.sum,lts__t_requests_op_atom.sum,l1tex__t_set_conflicts_pipe_lsu_mem_global_op_red.sum,l1tex__t_set_conflicts_pipe_lsu_mem_global_op_atom.sum,l1tex__t_requests_pipe_lsu_mem_global_op_red.sum,l1tex__t_requests_pipe_lsu_mem_global_op_atom.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum,smsp__inst_executed_op_shared_ld.sum \
  --csv --log-file gpurun_out/atomics_$tag.csv python tools/stage_times.py > /dev/null 2>&1
tail -9 gpurun_out/atomics_$tag.csv | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file gpurun_out/launches_$tag.csv python tools/stage_times.py > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_raycast|k_post|k_finalize|k_record|k_fuse|k_index' -s 60 -c 6 -o gpurun_out/prof_$tag -f python tools/stage_times.py > gpurun_out/ncu_$tag.log 2>&1
ls -la gpurun_out | tail -12
