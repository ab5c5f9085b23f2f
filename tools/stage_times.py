"""Print the per-stage device times of config-B frames (library CUDA events); quick A/B helper."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elevation_mapping_cupy_b200.parameter import core_parameter
from elevation_mapping_cupy_b200.elevation_mapping import ElevationMap
from elevation_mapping_cupy_b200 import workloads as wl
# usage: stage_times.py [cell_n | D]   (D = config D: 2048^2 grid, 1M-point depth-camera frames)
arg = sys.argv[1] if len(sys.argv) > 1 else "1024"
cell_n = 2048 if arg == "D" else int(arg)
p = core_parameter(cell_n)
em = ElevationMap(p)
frames = [wl.depth_camera_cloud(3, f) for f in range(4)] if arg == "D" else [wl.lidar_cloud(1, f) for f in range(4)]
dev = [torch.from_numpy(x[0]).cuda() for x in frames]
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
em.enable_stage_timing(True)
acc = np.zeros(8); n = 0
for it in range(24):
    f = it % 4
    em.move_to(frames[f][2], frames[f][1]); em.update_variance(); em.update_time(); flush.zero_(); torch.cuda.synchronize()
    em.input_pointcloud(dev[f], ["x", "y", "z"], frames[f][1], frames[f][2], 0.02, 0.02); em.synchronize()
    if it >= 8:
        acc += em.stage_ms(); n += 1
names = ["index", "drift", "fusion", "record", "raycast", "finalize", "post", "total"]
print(os.path.basename(os.environ.get("EMAP_LIB", "libemap.so")), arg, " ".join(f"{k}={1e3*v/n:.1f}us" for k, v in zip(names, acc)))
