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


_WORKER_C = r"""
import os, sys, time, ctypes as C
sys.path.insert(0, %(root)r)
import numpy as np
rk = int(os.environ["RANK"]); ws = int(os.environ["WORLD_SIZE"])
os.environ["CUDA_VISIBLE_DEVICES"] = os.environ.get("CUDA_VISIBLE_DEVICES", "")
from elevation_mapping_cupy_b200.parameter import core_parameter
from elevation_mapping_cupy_b200.elevation_mapping import ElevationMap
from elevation_mapping_cupy_b200 import workloads as wl, _lib
from oracle import oracle as O
# no torch.distributed anywhere: the communicator is created through the C ABI (emap_comm_unique_id / emap_comm_init)
L = _lib.load()
idfile = %(idfile)r
buf = (C.c_char * 128)()
if rk == 0:
    assert L.emap_comm_unique_id(buf) == 0
    open(idfile + ".tmp", "wb").write(bytes(buf)); os.rename(idfile + ".tmp", idfile)
else:
    for _ in range(600):
        if os.path.exists(idfile): break
        time.sleep(0.1)
    buf = (C.c_char * 128).from_buffer_copy(open(idfile, "rb").read())
p = core_parameter(256)
em = ElevationMap(p, device=rk)
em._check(L.emap_comm_init(em._h, buf, rk, ws))
om = O.OracleElevationMap(p, nthreads=0)
for f in range(3):
    clouds, Rs, ts = [], [], []
    for s in range(ws):
        pts, R, t = wl.lidar_cloud(2, f, n_rings=16, n_az=500 + 37 * s, max_range=6.0, sensor=s, n_sensors=ws)
        clouds.append(pts); Rs.append(R); ts.append(t)
    em.move_to(ts[0], Rs[0]); om.move_to(ts[0], Rs[0])
    mine = np.ascontiguousarray(clouds[rk])
    ptrs = (C.c_void_p * 1)(mine.ctypes.data); cnt = (C.c_int64 * 1)(len(mine))
    off = sum(len(c) for c in clouds[:rk])
    Rm = np.ascontiguousarray(Rs[rk], np.float32).reshape(9); tm = np.ascontiguousarray(ts[rk], np.float32)
    em._check(L.emap_input_sensors_sharded(em._h, 1, ptrs, cnt, 3, 0, 0, Rm.ctypes.data, tm.ctypes.data, off, float(ts[0][2]), 0.02, 0.02))
    om.input_sensors(clouds, Rs, ts, 0.02, 0.02)
    state, normal = em.get_state()
    for li in (0, 1, 2, 4, 5, 6):
        assert np.array_equal(state[li], om.elevation_map[li]), (rk, f, li)
    assert np.abs(state[3] - om.elevation_map[3]).max() < 2e-6
    assert np.array_equal(normal, om.normal_map)
    em.update_time(); om.update_time()
print("C_SHARDED_OK", rk, flush=True)
em.close()
"""


def test_two_rank_sharded_frame_through_the_c_abi_only(tmp_path):
    """SURVEY 8(b) emap_comm_init: ranks create the NCCL communicator through libemap.so (no torch.distributed) and every
    frame is one emap_input_sensors_sharded call; each replica must equal the multi-sensor oracle bit for bit."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "wc.py"
    script.write_text(_WORKER_C % {"root": ROOT, "idfile": str(tmp_path / "nccl_id.bin")})
    procs = []
    for rk in range(2):
        env = dict(os.environ, RANK=str(rk), WORLD_SIZE="2")
        procs.append(subprocess.Popen([sys.executable, str(script)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all("C_SHARDED_OK" in o for o in outs), "\n-----\n".join(o[-3000:] for o in outs)
