"""The engine against the REFERENCE'S OWN kernels running on the same B200 (tests/ref_gpu.py), plus their
timing side by side (written to gpurun_out/ref_gpu_timing.json when that directory exists)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mk(param):
    from elevation_mapping_cupy_b200.elevation_mapping import ElevationMap
    return ElevationMap(param)


def _ref(param, tag):
    from ref_gpu import RefGpuMap
    try:
        return RefGpuMap(param, tag)
    except FileNotFoundError:
        pytest.skip("oracle/_ref/libref_gpu_*.so not prebuilt")


def test_indices_bit_identical_to_reference_gpu_kernels():
    import torch
    from elevation_mapping_cupy_b200.parameter import Parameter, core_parameter
    from elevation_mapping_cupy_b200 import workloads as wl
    for tag, p, cloud in (("default202", Parameter(), "rand"), ("core1024", core_parameter(1024), "lidar")):
        p.update()
        em = _mk(p); rg = _ref(p, tag)
        if cloud == "rand":
            pts, R, t = wl.reference_test_cloud(3)
            pts = (pts * np.float32(9.0) - np.float32(4.5)).astype(np.float32)
        else:
            pts, R, t = wl.lidar_cloud(1, 2)
        em.input_pointcloud(pts, ["x", "y", "z"], R, t, 0, 0)
        wb = rg.input_pointcloud(pts, R, t, 0, 0).cpu().numpy()
        idx, valid, inside = em.get_point_record(len(pts))
        assert np.array_equal(idx, wb[:, 0].astype(np.int32)), tag
        assert np.array_equal(valid, wb[:, 1].astype(np.uint8)), tag
        assert np.array_equal(inside, wb[:, 2].astype(np.uint8)), tag


def test_state_matches_reference_gpu_kernels_on_order_independent_cells():
    """Same pre-frame state in both; the reference GPU kernel is run three times per frame (its races resolve
    differently from launch to launch and for permuted inputs): cells on which those runs agree are order-
    independent, and there the engine must match to 1e-4 (BASELINE) -- in practice ~1e-6."""
    import torch
    from elevation_mapping_cupy_b200.parameter import core_parameter
    from elevation_mapping_cupy_b200 import workloads as wl
    p = core_parameter(256)
    em = _mk(p); refs = [_ref(p, "core256") for _ in range(3)]
    rng = np.random.default_rng(4)
    worst = 0.0
    worst_trav = 0.0
    for f in range(5):
        pts, R, t = wl.lidar_cloud(0, f, n_rings=32, n_az=625, max_range=8.0)
        em.move_to(t, R)
        st, nm = em.get_state()
        outs = []
        for k, rg in enumerate(refs):
            rg.set_state(st, nm, em.center)
            pp = pts if k == 0 else (pts[::-1].copy() if k == 1 else pts[rng.permutation(len(pts))])
            rg.input_pointcloud(pp, R, t, 0.02, 0.02)
            outs.append(rg.elevation_map.cpu().numpy())
        em.input_pointcloud(pts, ["x", "y", "z"], R, t, 0.02, 0.02)
        state, _ = em.get_state()
        # height / variance / validity / time: compare where the reference agrees with itself.  Three runs of a racy
        # kernel can still agree on a cell whose outcome is order-dependent (e.g. outlier-inlier-outlier in input order
        # looks the same forward and reversed), so a handful of cells per frame may legitimately differ: the bar is
        # <= 1e-4 on all but at most 0.1 % of the compared cells, and the typical difference is reported.
        racy = np.zeros((256, 256), bool)
        for li in (0, 1, 2, 4):
            for o in outs[1:]:
                racy |= np.abs(outs[0][li] - o[li]) > 1e-6
        assert racy.mean() < 0.06, racy.mean()
        n_cmp = int((~racy).sum())
        # traversability (layer 3): the reference computes it with torch / cuDNN convolutions (traversability_filter.py:
        # 15-42, TF32 off); it depends on the dilated upper bound within a 13 x 13 window, so compare where no racy or
        # upper-bound-unstable cell lies in that window
        ub_ok = ~racy
        for o in outs[1:]:
            ub_ok &= (np.abs(outs[0][5] - o[5]) <= 1e-6) & (outs[0][6] == o[6])
        ub_ok &= (np.abs(state[5] - outs[0][5]) <= 1e-6) & (state[6] == outs[0][6])
        from scipy import ndimage
        clean = ndimage.minimum_filter(ub_ok.astype(np.uint8), size=13, mode="constant", cval=0).astype(bool)
        clean[:3, :] = clean[-3:, :] = False; clean[:, :3] = clean[:, -3:] = False
        if clean.sum() > 100:
            dt = np.abs(state[3] - outs[0][3])[clean]
            assert float(dt.max()) <= 1e-4, (f, "traversability vs cuDNN", float(dt.max()), int(clean.sum()))
            worst_trav = max(worst_trav, float(dt.max()))
        for li in (0, 1, 2, 4):
            d = np.abs(state[li] - outs[0][li])[~racy]
            bad = int((d > 1e-4).sum())
            assert bad <= max(5, n_cmp // 1000), (f, li, bad, float(d.max()))
            worst = max(worst, float(np.sort(d)[-(bad + 1)]) if bad < len(d) else 0.0)
        # upper_bound: the reference's check-then-store (CK.py:230-233,253-256) loses updates, so a carved cell
        # holds SOME ray's height; the engine holds the true minimum: never above any reference outcome, and
        # is_upper_bound identical wherever the three reference runs agree with each other
        ub_stable = ~racy
        for o in outs[1:]:
            ub_stable &= (np.abs(outs[0][5] - o[5]) <= 1e-6) & (outs[0][6] == o[6])
        assert int((state[6][ub_stable] != outs[0][6][ub_stable]).sum()) <= max(5, n_cmp // 1000)
        carved = ub_stable & (state[6] > 0.5)
        for o in outs:
            both = carved & (o[6] > 0.5)
            assert int((state[5][both] > o[5][both] + 1e-6).sum()) <= max(5, n_cmp // 1000)
        # (upper_bound of a cell fused by several points is the new_h of an ARBITRARY one of them in the reference,
        # CK.py:191; the engine takes the last in input order -- not comparable cell by cell)
        em.update_variance(); em.update_time()
    print("largest abs difference vs reference GPU kernels over the accepted order-independent cells:", worst,
          "; traversability vs the cuDNN path:", worst_trav)


def test_reference_gpu_kernels_timing_config_b():
    """REF-GPU: the reference's kernels (nvcc build of its own source) + torch/cuDNN traversability on the same
    B200, same frames as bench.py; device time by CUDA events.  Not an assertion of speed, a record."""
    import torch
    from elevation_mapping_cupy_b200.parameter import core_parameter
    from elevation_mapping_cupy_b200 import workloads as wl
    p = core_parameter(1024)
    em = _mk(p); rg = _ref(p, "core1024")
    frames = [wl.lidar_cloud(1, f) for f in range(4)]
    dev = [torch.from_numpy(x[0]).cuda() for x in frames]
    t_ref, t_new = [], []
    for it in range(12):
        f = it % 4
        pts, R, t = frames[f]
        em.move_to(t, R); em.update_variance(); em.update_time()
        st, nm = em.get_state()
        rg.set_state(st, nm, em.center)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); rg.input_pointcloud(pts, R, t, 0.02, 0.02, points_dev=dev[f]); e1.record(); torch.cuda.synchronize()
        if it >= 4:
            t_ref.append(e0.elapsed_time(e1))
        em.enable_stage_timing(True)
        em.input_pointcloud(dev[f], ["x", "y", "z"], R, t, 0.02, 0.02); em.synchronize()
        if it >= 4:
            t_new.append(float(em.stage_ms()[7]))
    rec = {"workload": "config B: 1024^2, 200k-pt LiDAR frame, warm L2", "ref_gpu_ms_per_frame": float(np.mean(t_ref)),
           "engine_ms_per_frame": float(np.mean(t_new)), "speedup": float(np.mean(t_ref) / np.mean(t_new)),
           "note": "reference = its own CUDA-C kernel strings compiled by nvcc for sm_100a + torch/cuDNN convs, incl. its 2 host syncs"}
    print(json.dumps(rec))
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        json.dump(rec, open(os.path.join(out, "ref_gpu_timing.json"), "w"))
    assert rec["engine_ms_per_frame"] > 0
