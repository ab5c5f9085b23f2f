"""Frame times of the other BASELINE.json configs (A, D, E) on one GPU -- parity cases in tests/, timed here for
DESIGN.md.  Device-resident clouds, CUDA events around each frame, L2 flushed between frames except for E
(E measures throughput of independent maps issued back to back on their own streams)."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elevation_mapping_cupy_b200.parameter import core_parameter
from elevation_mapping_cupy_b200.elevation_mapping import ElevationMap
from elevation_mapping_cupy_b200 import workloads as wl

flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
out = {}


def timed(em, frames, n_iter, warm):
    s = torch.cuda.Stream()
    torch.cuda.set_stream(s)
    em._check(em._L.emap_set_stream(em._h, s.cuda_stream))
    dev = [torch.from_numpy(f[0]).cuda() for f in frames]
    ms = []
    for it in range(warm + n_iter):
        f = it % len(frames)
        em.move_to(frames[f][2], frames[f][1]); em.update_variance(); em.update_time(); flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(s); em.input_pointcloud(dev[f], ["x", "y", "z"], frames[f][1], frames[f][2], 0.02, 0.02); e1.record(s)
        torch.cuda.synchronize()
        if it >= warm:
            ms.append(e0.elapsed_time(e1))
    return float(np.mean(ms))


# A: 256^2, 10k-point uniform cloud
em = ElevationMap(core_parameter(256))
fr = [wl.uniform_cloud(0, f) for f in range(4)]
m = timed(em, fr, 40, 8); out["A 256^2 / 10k pts"] = {"ms_per_frame": m, "mpoints_s": 10000 / m / 1e3}
# D: 2048^2, 1M-point depth camera (5 % NaN)
em = ElevationMap(core_parameter(2048))
fr = [wl.depth_camera_cloud(3, f) for f in range(2)]
m = timed(em, fr, 12, 4); out["D 2048^2 / 1M pts"] = {"ms_per_frame": m, "mpoints_s": 1e6 / m / 1e3}
data = np.zeros((2046, 2046), np.float32)
torch.cuda.synchronize(); t0 = time.perf_counter()
for name in ("min_filter", "smooth", "erosion"):
    em.get_map_with_name_ref(name, data)
out["D plugin chain (min_filter 30 it + smooth + erosion, incl. D2H of 3 layers)"] = {"ms": 1e3 * (time.perf_counter() - t0)}
del em
# E: 8 independent 512^2 maps x 100k points on one GPU, one stream each, frames issued round robin
ems = [ElevationMap(core_parameter(512)) for _ in range(8)]
clouds = [[wl.lidar_cloud(4, 10 * k + f, n_rings=32, n_az=3125, max_range=12.0) for f in range(2)] for k in range(8)]
dev = [[torch.from_numpy(c[0]).cuda() for c in cs] for cs in clouds]
def rnd(f):
    for k, e in enumerate(ems):
        pts, R, t = clouds[k][f]
        e.move_to(t, R); e.input_pointcloud(dev[k][f], ["x", "y", "z"], R, t, 0.02, 0.02); e.update_time()
for it in range(4):
    rnd(it % 2)
for e in ems: e.synchronize()
t0 = time.perf_counter(); n = 20
for it in range(n):
    rnd(it % 2)
for e in ems: e.synchronize()
dt = time.perf_counter() - t0
out["E 8 maps x 512^2 / 100k pts, 1 GPU (wall clock, warm L2)"] = {"map_frames_per_s": 8 * n / dt, "mpoints_s": 8 * n * 1e5 / dt / 1e6}
print(json.dumps(out, indent=1))
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "config_times.json"), "w"), indent=1)
