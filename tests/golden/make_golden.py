"""Generates tests/golden/*.npz.  Run HERE (needs /root/reference): the "ref_*" arrays come from the
REFERENCE'S OWN kernel source (custom_kernels.py strings compiled for the host by oracle/build_ref.py and
driven in the order of elevation_mapping.py:316-391); the "oracle_*" arrays from oracle/emap_oracle.c.

    python tests/golden/make_golden.py

Inputs are regenerated from seeds by elevation_mapping_cupy_b200/workloads.py, so only outputs are stored.
The reference kernel is racy; `racy` marks the cells whose outcome differs between executing the
points in input order, in reverse order and in a fixed random order -- parity is asserted outside that mask.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from elevation_mapping_cupy_b200 import workloads as wl          # noqa: E402
from elevation_mapping_cupy_b200.parameter import Parameter, core_parameter   # noqa: E402
from oracle import oracle as O                                   # noqa: E402


def index_case(seed):
    """reference test shape (test_elevation_mapping.py:51-61), default parameters, 202^2 map"""
    p = Parameter(); p.update()
    pts, R, t = wl.reference_test_cloud(seed, n=20000)
    if seed % 2:
        pts = (pts * np.float32(9.0) - np.float32(4.5)).astype(np.float32)
    rm = O.RefKernelMap(p, "default202")
    rm.input_pointcloud(pts.copy(), ["x", "y", "z"], R, t, 0, 0)
    idx, valid, inside = rm.last_point_record
    return dict(idx=idx, valid=valid, inside=inside)


def frames_case(cell_n=130, n_frames=4, tag="core130"):
    p = core_parameter(cell_n)
    om = O.OracleElevationMap(p)
    out = {}
    perm_rng = np.random.default_rng(99)
    for f in range(n_frames):
        pts, R, t = wl.lidar_cloud(0, f, n_rings=24, n_az=500, max_range=4.0)
        refs = []
        for order in ("fwd", "rev", "perm"):
            rm = O.RefKernelMap(p, tag)
            rm.elevation_map = om.elevation_map.copy(); rm.normal_map = om.normal_map.copy()
            rm.center = om.center.copy(); rm.additive_mean_error = om.additive_mean_error
            pp = pts if order == "fwd" else (pts[::-1].copy() if order == "rev" else pts[perm_rng.permutation(len(pts))])
            rm.move_to(t, R)
            rm.input_pointcloud(pp, ["x", "y", "z"], R, t, 0.02, 0.02)
            refs.append(rm)
        om.move_to(t, R)
        om.input_pointcloud(pts, ["x", "y", "z"], R, t, 0.02, 0.02)
        racy = np.zeros((cell_n, cell_n), bool)
        for li in (0, 1, 2, 4, 5, 6):
            for other in refs[1:]:
                racy |= np.abs(refs[0].elevation_map[li] - other.elevation_map[li]) > 1e-6
        out[f"ref_state_{f}"] = refs[0].elevation_map.copy()
        out[f"ref_normal_{f}"] = refs[0].normal_map.copy()
        out[f"racy_{f}"] = racy
        out[f"oracle_state_{f}"] = om.elevation_map.copy()
        out[f"oracle_normal_{f}"] = om.normal_map.copy()
        out[f"ref_point_idx_{f}"] = refs[0].last_point_record[0]
        om.update_variance(); om.update_time()
    return out


if __name__ == "__main__":
    for seed in (0, 1):
        np.savez_compressed(os.path.join(HERE, f"index_default202_seed{seed}.npz"), **index_case(seed))
    np.savez_compressed(os.path.join(HERE, "frames_core130.npz"), **frames_case())
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
