"""2-GPU sharded frame (NCCL): each rank fuses one sensor; every replica must equal the single-process
multi-sensor oracle bit for bit.  Skipped unless >= 2 CUDA devices are visible."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
rk = int(os.environ["RANK"]); ws = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rk)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{rk}"))
from elevation_mapping_cupy_b200.parameter import core_parameter
from elevation_mapping_cupy_b200.elevation_mapping import ElevationMap
from elevation_mapping_cupy_b200.sharded import ShardedElevationMap
from elevation_mapping_cupy_b200 import workloads as wl
from oracle import oracle as O
p = core_parameter(256)
em = ElevationMap(p, device=rk)
sh = ShardedElevationMap(em, mode=os.environ.get('EMAP_SHARD_MODE', 'auto'))
print('mode', sh.mode, flush=True)
om = O.OracleElevationMap(p, nthreads=0)
for f in range(3):
    clouds, Rs, ts = [], [], []
    for s in range(ws):
        pts, R, t = wl.lidar_cloud(2, f, n_rings=16, n_az=500 + 37 * s, max_range=6.0, sensor=s, n_sensors=ws)
        clouds.append(pts); Rs.append(R); ts.append(t)
    R0, t0 = Rs[0], ts[0]
    em.move_to(t0, R0); om.move_to(t0, R0)
    sh.input_sensors([clouds[rk]], [Rs[rk]], [ts[rk]], 0.02, 0.02, overlap_z=float(t0[2]))
    om.input_sensors(clouds, Rs, ts, 0.02, 0.02)
    state, normal = em.get_state()
    for li in (0, 1, 2, 4, 5, 6):
        assert np.array_equal(state[li], om.elevation_map[li]), (rk, f, li)
    assert np.abs(state[3] - om.elevation_map[3]).max() < 2e-6
    assert np.array_equal(normal, om.normal_map)
    em.update_time(); om.update_time()
dist.barrier()
if rk == 0:
    print("SHARDED_OK")
dist.destroy_process_group()
"""


@pytest.mark.parametrize("mode", ["nccl", "multicast"])
def test_two_rank_sharded_frame_matches_oracle(tmp_path, mode):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "w.py"
    script.write_text(_WORKER % {"root": ROOT})
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29544", str(script)],
                       capture_output=True, text=True, timeout=600, env=dict(os.environ, EMAP_SHARD_MODE=mode))
    if mode == "multicast" and "multicast is not available" in (r.stdout + r.stderr):
        pytest.skip("fabric has no NVLink multicast")
    assert "SHARDED_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
