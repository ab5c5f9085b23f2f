"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes front-end of oracle/emap_oracle.c plus the numpy restatement of the array-level
orchestration in elevation_mapping.py (EM.py = elevation_mapping_cupy/script/
elevation_mapping_cupy/elevation_mapping.py of the reference):

  OracleElevationMap   -- the canonical-serialisation oracle; the thing
                          the CUDA engine is compared with.
  RefKernelMap         -- the same orchestration, but every kernel is the REFERENCE'S OWN
                          source compiled for the host by oracle/build_ref.py
                          (oracle/_ref/libref_cpu_<tag>.so), executed one element at a
                          time in input order.  Used to pin the oracle to the reference.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

LAYER_NAMES = ["elevation", "variance", "is_valid", "traversability", "time", "upper_bound", "is_upper_bound"]


class OracleParams(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "cell_n", "dilation_size", "enable_edge_sharpen", "enable_drift_compensation",
        "enable_visibility_cleanup", "enable_overlap_clearance", "cell_min", "cell_max")] + [
        (n, C.c_double) for n in (
            "resolution", "sensor_noise_factor", "mahalanobis_thresh", "outlier_variance",
            "drift_compensation_variance_inlier", "traversability_inlier", "wall_num_thresh",
            "min_height_drift_cnt", "max_ray_length", "cleanup_step", "cleanup_cos_thresh",
            "min_valid_distance", "max_height_range", "ramped_height_range_a", "ramped_height_range_b",
            "ramped_height_range_c", "max_variance", "initial_variance", "max_drift",
            "drift_compensation_alpha", "position_noise_thresh", "orientation_noise_thresh",
            "overlap_clear_range_z")]


class FrameStats(C.Structure):
    _fields_ = [("mean_error", C.c_float), ("additive_mean_error", C.c_float), ("shift_applied", C.c_float),
                ("error_sum", C.c_float), ("error_cnt", C.c_int64), ("drift_applied", C.c_int32), ("drift_evaluated", C.c_int32),
                ("ray_visits", C.c_int64), ("ray_steps", C.c_int64)]


def build_oracle(force=False):
    """gcc-compile the C restatement into oracle/_build/liboracle.so (git-ignored)."""
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "liboracle.so")
    src = os.path.join(HERE, "emap_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["gcc", "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC",
                        "-Wall", "-o", so, src, "-lm"], check=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build_oracle())
        _LIB.oracle_max_threads.restype = C.c_int
        _LIB.oracle_min_filter.restype = C.c_int
    return _LIB


def _p(a, t=C.c_float):
    return a.ctypes.data_as(C.POINTER(t))


def overlap_window(param):
    """EM.py:87-91"""
    cell_range = int(param.overlap_clear_range_xy / param.resolution)
    cell_range = int(np.clip(cell_range, 0, param.cell_n))
    return param.cell_n // 2 - cell_range // 2, param.cell_n // 2 + cell_range // 2


def make_params(param):
    op = OracleParams()
    op.cell_n = int(param.cell_n)
    op.dilation_size = int(param.dilation_size)
    for k in ("enable_edge_sharpen", "enable_drift_compensation", "enable_visibility_cleanup",
              "enable_overlap_clearance"):
        setattr(op, k, int(bool(getattr(param, k))))
    op.cell_min, op.cell_max = overlap_window(param)
    for n, _ in OracleParams._fields_[8:]:
        setattr(op, n, float(getattr(param, n)))
    return op


def point_index(param, points, R, t_rel):
    """(idx, valid, inside, xyzv) per point: the write-back of CK.py:260-262."""
    pts = np.ascontiguousarray(points, dtype=np.float32)
    n = pts.shape[0]
    idx = np.zeros(n, np.int32); valid = np.zeros(n, np.uint8); inside = np.zeros(n, np.uint8)
    xyzv = np.zeros((n, 4), np.float32)
    op = make_params(param)
    R = np.ascontiguousarray(R, np.float32).reshape(9); t = np.ascontiguousarray(t_rel, np.float32).reshape(3)
    lib().oracle_point_index(C.byref(op), _p(pts), C.c_int64(n), C.c_int64(pts.shape[1]), _p(R), _p(t),
                             _p(idx, C.c_int32), _p(valid, C.c_uint8), _p(inside, C.c_uint8), _p(xyzv))
    return idx, valid, inside, xyzv


def point_index_numpy(param, points, R, t_rel):
    """Vectorised NumPy restatement of CK.py:22-81,160-167 (independent of the C code):
    explicit np.float16 roundings, double where a literal is involved."""
    f16 = lambda a: np.asarray(a, np.float32).astype(np.float16).astype(np.float32)
    P = f16(points[:, :3]); R16 = f16(np.asarray(R, np.float32).reshape(3, 3)); t16 = f16(np.asarray(t_rel, np.float32))
    W = param.cell_n
    xyz = []
    for r in range(3):
        a = (R16[r, 0] * P[:, 0]).astype(np.float32) + (R16[r, 1] * P[:, 1]).astype(np.float32)
        a = a.astype(np.float32) + (R16[r, 2] * P[:, 2]).astype(np.float32)
        xyz.append((a.astype(np.float32) + t16[r]).astype(np.float32))
    x, y, z = xyz
    X, Y, Z = f16(x), f16(y), f16(z)

    def axis(c16):
        i = np.trunc(c16.astype(np.float64) / param.resolution + 0.5 * W)
        i = np.clip(i, -2147483648.0, 2147483647.0).astype(np.int64)
        c = f16(i.astype(np.float32))
        return np.maximum(np.minimum(c, f16(np.float32(W - 1))), np.float32(0)).astype(np.int64)
    ix, iy = axis(X), axis(Y)
    idx = (W * ix + iy).astype(np.int32)
    inside = (ix != 0) & (ix != W - 1) & (iy != 0) & (iy != W - 1)
    dx, dy, dz = X - t16[0], Y - t16[1], Z - t16[2]
    # FMAD contraction of CK.py:64 as nvcc emits it: fma(dz,dz, fma(dx,dx, dy*dy)); emulate in float64
    # (exact for the fp32 inputs' products) with one rounding per fma.
    d = (dy * dy).astype(np.float32)
    d = (dx.astype(np.float64) * dx.astype(np.float64) + d.astype(np.float64)).astype(np.float32)
    d = (dz.astype(np.float64) * dz.astype(np.float64) + d.astype(np.float64)).astype(np.float32)
    sq = np.sqrt((X * X + Y * Y).astype(np.float32)).astype(np.float32)
    dxy = np.maximum(sq.astype(np.float64) - param.ramped_height_range_b, 0.0).astype(np.float32)
    zr = (Z - t16[2]).astype(np.float32).astype(np.float64)
    # double fma(dxy, a, c): float64 product of an fp32 by a double is not exact; np has no fma, so use
    # extended precision for the single rounding.
    lim = (dxy.astype(np.longdouble) * np.longdouble(param.ramped_height_range_a)
           + np.longdouble(param.ramped_height_range_c)).astype(np.float64)
    valid = ~(d.astype(np.float64) < param.min_valid_distance * param.min_valid_distance)
    valid &= ~((zr > lim) | (zr > param.max_height_range))
    return idx, valid.astype(np.uint8), inside.astype(np.uint8)


def load_weights(param):
    if not np.any(param.w1):
        param.load_weights(param.weight_file)
    return [np.ascontiguousarray(w, np.float32).reshape(-1) for w in (param.w1, param.w2, param.w3, param.w_out)]


class _MapBase:
    """numpy restatement of the array-level part of ElevationMap (EM.py:52-226,393-432,579-775)."""

    def __init__(self, param):
        self.param = param
        self.resolution = param.resolution
        self.cell_n = W = param.cell_n
        self.center = np.zeros(3, np.float32)
        self.base_rotation = np.eye(3, dtype=np.float32)
        self.elevation_map = np.zeros((7, W, W), np.float32)
        self.layer_names = list(LAYER_NAMES)
        self.normal_map = np.zeros((3, W, W), np.float32)
        self.traversability_input = np.zeros((W, W), np.float32)
        self.traversability_buffer = np.full((W, W), np.nan)
        self.initial_variance = param.initial_variance
        self.elevation_map[1] += np.float32(self.initial_variance)
        self.elevation_map[3] += 1.0
        self.cell_min, self.cell_max = overlap_window(param)
        self.mean_error = 0.0
        self.additive_mean_error = 0.0
        self.weights = load_weights(param)
        self.last_point_record = None

    # EM.py:119-128
    def clear(self):
        self.elevation_map *= 0.0
        self.elevation_map[1] += np.float32(self.initial_variance)
        self.mean_error = 0.0
        self.additive_mean_error = 0.0

    def get_position(self, position):
        position[0][:] = self.center

    # EM.py:139-152
    def move(self, delta_position):
        delta_position = np.asarray(delta_position, dtype=np.float64)
        delta_pixel = np.round(delta_position[:2] / self.resolution)
        delta_position_xy = delta_pixel * self.resolution
        self.center[:2] += delta_position_xy
        self.center[2] += delta_position[2]
        self.shift_map_xy(delta_pixel)
        self.shift_map_z(-delta_position[2])

    # EM.py:154-170
    def move_to(self, position, R):
        self.base_rotation = np.asarray(R, dtype=np.float32)
        position = np.asarray(position, dtype=np.float64)
        delta = position - self.center.astype(np.float64)
        delta_pixel = np.around(delta[:2] / self.resolution)
        delta_xy = delta_pixel * self.resolution
        self.center[:2] += delta_xy
        self.center[2] += delta[2]
        self.shift_map_xy(-delta_pixel)
        self.shift_map_z(-delta[2])

    # EM.py:172-214
    def shift_map_xy(self, delta_pixel):
        sv = np.asarray(delta_pixel).astype(np.int32)
        if np.abs(sv).sum() == 0:
            return
        m = np.roll(self.elevation_map, sv, axis=(1, 2))
        for idx, value in ((None, 0.0), (1, self.initial_variance)):
            sel = slice(None) if idx is None else idx
            if sv[0] > 0:
                m[sel, : sv[0], :] = value
            elif sv[0] < 0:
                m[sel, sv[0]:, :] = value
            if sv[1] > 0:
                m[sel, :, : sv[1]] = value
            elif sv[1] < 0:
                m[sel, :, sv[1]:] = value
        self.elevation_map = m

    # EM.py:216-226: float32 plane += float64 0-d array -> computed in float64, stored fp32
    def shift_map_z(self, delta_z):
        dz = np.float64(delta_z)
        self.elevation_map[0] = (self.elevation_map[0].astype(np.float64) + dz).astype(np.float32)
        self.elevation_map[5] = (self.elevation_map[5].astype(np.float64) + dz).astype(np.float32)

    # EM.py:420-426
    def update_variance(self):
        self.elevation_map[1] += np.float32(self.param.time_variance) * self.elevation_map[2]

    def update_time(self):
        self.elevation_map[4] += np.float32(self.param.time_interval)

    def exists_layer(self, name):
        return name in self.layer_names

    def get_layer(self, name):
        return self.elevation_map[self.layer_names.index(name)]

    # EM.py:579-670
    def _publish(self, m, fill_nan=False, add_z=False):
        m = m.copy()
        if fill_nan:
            m = np.where(self.elevation_map[2] > 0.5, m, np.nan)
        if add_z:
            m = m + self.center[2]
        return m[1:-1, 1:-1]

    def _upper_valid(self):
        em = self.elevation_map
        if self.param.use_only_above_for_upper_bound:
            return np.logical_or(np.logical_and(em[5] > 0.0, em[6] > 0.5), em[2] > 0.5)
        return np.logical_or(em[2] > 0.5, em[6] > 0.5)

    def export_layer(self, name):
        """EM.py:720-775 without the D2H: the (W-2,W-2) array get_map_with_name_ref writes."""
        em = self.elevation_map
        if name == "elevation":
            m = self._publish(em[0], fill_nan=True, add_z=True)
        elif name == "variance":
            m = self._publish(em[1])
        elif name == "traversability":
            trav = np.where((em[2] + em[6]) > 0.5, em[3].copy(), np.nan)
            self.traversability_buffer[3:-3, 3:-3] = trav[3:-3, 3:-3]
            m = self.traversability_buffer[1:-1, 1:-1]
        elif name == "time":
            m = self._publish(em[4])
        elif name == "upper_bound":
            m = np.where(self._upper_valid(), em[5].copy(), np.nan)[1:-1, 1:-1] + self.center[2]
        elif name == "is_upper_bound":
            m = np.where(self._upper_valid(), em[6].copy(), np.nan)[1:-1, 1:-1]
        elif name in ("normal_x", "normal_y", "normal_z"):
            m = self.normal_map[("normal_x", "normal_y", "normal_z").index(name), 1:-1, 1:-1]
        else:
            raise KeyError(name)
        m = np.flip(np.flip(m, 0), 1)
        return m.astype(np.float32)

    def get_map_with_name_ref(self, name, data):
        data[...] = self.export_layer(name)

    # EM.py:434-466
    def input_pointcloud(self, raw_points, channels, R, t, position_noise, orientation_noise):
        pts = np.asarray(raw_points, dtype=np.float32)
        t = np.array(t, dtype=np.float32).reshape(3)
        R = np.asarray(R, dtype=np.float32).reshape(3, 3)
        self.update_map_with_kernel(pts, channels[3:], R, t, position_noise, orientation_noise)

    input = input_pointcloud


class OracleElevationMap(_MapBase):
    """Canonical-serialisation oracle of one ElevationMap (multi-sensor aware)."""

    def __init__(self, param, nthreads=1):
        super().__init__(param)
        self.nthreads = nthreads
        self.stats = None
        self.counts = np.zeros((3, self.cell_n, self.cell_n), np.float32)   # new_map[3], new_map[4], new_map[2]

    @property
    def counts_fused(self):
        """new_map[2] of the last frame (CK.py:185): points fused per cell"""
        return self.counts[2]

    def update_map_with_kernel(self, points_all, channels, R, t, position_noise, orientation_noise):
        self.input_sensors([points_all], [R], [t], position_noise, orientation_noise)

    def input_sensors(self, clouds, Rs, ts, position_noise, orientation_noise):
        """One frame fusing several sensors' clouds against the same snapshot;
        with one sensor this is EM.py:316-391."""
        pts = np.ascontiguousarray(np.concatenate([np.asarray(c, np.float32)[:, :3] for c in clouds], 0))
        offs = np.zeros(len(clouds) + 1, np.int64)
        offs[1:] = np.cumsum([len(c) for c in clouds])
        Rm = np.ascontiguousarray(np.stack([np.asarray(r, np.float32).reshape(9) for r in Rs]))
        tm = np.ascontiguousarray(np.stack([(np.asarray(t, np.float32).reshape(3) - self.center).astype(np.float32)
                                            for t in ts]))          # EM.py:314,333
        n = pts.shape[0]
        idx = np.zeros(n, np.int32); valid = np.zeros(n, np.uint8); inside = np.zeros(n, np.uint8)
        st = FrameStats(); op = make_params(self.param)
        w1, w2, w3, wo = self.weights
        em = self.elevation_map
        assert em.flags.c_contiguous and em.dtype == np.float32
        lib().oracle_frame(C.byref(op), _p(em), _p(self.normal_map), _p(self.traversability_input),
                           _p(w1), _p(w2), _p(w3), _p(wo), _p(pts), C.c_int64(n), C.c_int64(3),
                           _p(Rm), _p(tm), _p(offs, C.c_int64), C.c_int32(len(clouds)),
                           C.c_float(position_noise), C.c_float(orientation_noise),
                           C.c_float(self.additive_mean_error),
                           _p(idx, C.c_int32), _p(valid, C.c_uint8), _p(inside, C.c_uint8),
                           _p(self.counts), C.byref(st), C.c_int(self.nthreads))
        if st.drift_evaluated:
            self.mean_error = st.mean_error
        self.additive_mean_error = st.additive_mean_error
        self.stats = st
        self.last_point_record = (idx, valid, inside)

    def update_normal(self, dilated_map):
        self.normal_map[...] = 0
        lib().oracle_normal(C.c_int(self.cell_n), C.c_double(self.resolution),
                            _p(np.ascontiguousarray(dilated_map, np.float32)),
                            _p(np.ascontiguousarray(self.elevation_map[2])), _p(self.normal_map))


# ---- standalone kernels -----------------------------------------------------------------

def dilation(W, k, h, mask):
    out = np.zeros((W, W), np.float32); om = np.zeros((W, W), np.float32)
    lib().oracle_dilation(C.c_int(W), C.c_int(k), _p(np.ascontiguousarray(h, np.float32)),
                          _p(np.ascontiguousarray(mask, np.float32)), _p(out), _p(om))
    return out, om


def traversability(W, x, weights):
    out = np.zeros((W - 6, W - 6), np.float32)
    w1, w2, w3, wo = weights
    lib().oracle_traversability(C.c_int(W), _p(np.ascontiguousarray(x, np.float32)), _p(w1), _p(w2), _p(w3), _p(wo), _p(out))
    return out


def normal(W, resolution, h, mask):
    out = np.zeros((3, W, W), np.float32)
    lib().oracle_normal(C.c_int(W), C.c_double(resolution), _p(np.ascontiguousarray(h, np.float32)),
                        _p(np.ascontiguousarray(mask, np.float32)), _p(out))
    return out


def min_filter(W, k, iteration_n, h, mask):
    out = np.zeros((W, W), np.float32)
    it = lib().oracle_min_filter(C.c_int(W), C.c_int(k), C.c_int(iteration_n), _p(np.ascontiguousarray(h, np.float32)),
                                 _p(np.ascontiguousarray(mask, np.float32)), _p(out))
    return out, it


def max_filter(W, k, iteration_n, h, mask):
    out = np.zeros((W, W), np.float32)
    lib().oracle_max_filter.restype = C.c_int
    it = lib().oracle_max_filter(C.c_int(W), C.c_int(k), C.c_int(iteration_n), _p(np.ascontiguousarray(h, np.float32)),
                                 _p(np.ascontiguousarray(mask, np.float32)), _p(out))
    return out, it


def robot_centric(W, resolution, threshold, use_threshold, elev, valid, R):
    out = np.zeros((W, W), np.float32)
    lib().oracle_robot_centric(C.c_int(W), C.c_double(resolution), C.c_double(threshold), C.c_int(int(use_threshold)),
                               _p(np.ascontiguousarray(elev, np.float32)), _p(np.ascontiguousarray(valid, np.float32)),
                               _p(np.ascontiguousarray(R, np.float32).reshape(9)), _p(out))
    return out


def erosion_cv2(layer, kernel_size=3, iterations=1, reverse=False):
    """plugins/erosion.py:96-113 verbatim (OpenCV on the host)."""
    import cv2 as cv
    layer_np = np.asarray(layer, np.float32)
    kernel = np.ones((kernel_size, kernel_size), np.uint8)
    if reverse:
        layer_np = 1 - layer_np
    layer_min = float(layer_np.min()); layer_max = float(layer_np.max())
    norm = ((layer_np - layer_min) * 255 / (layer_max - layer_min)).astype("uint8")
    er = cv.erode(norm, kernel, iterations=iterations)
    er = er.astype(np.float32) * (layer_max - layer_min) / 255 + layer_min
    if reverse:
        er = 1 - er
    return er


def smooth(W, h):
    out = np.zeros((W, W), np.float32)
    lib().oracle_smooth(C.c_int(W), _p(np.ascontiguousarray(h, np.float32)), _p(out))
    return out


def inpaint_cv2(elevation, is_valid, method="telea"):
    """plugins/inpainting.py:53-63 verbatim semantics (OpenCV on the host)."""
    import cv2 as cv
    mask = (is_valid < 0.5).astype("uint8")
    if (mask < 1).any():
        h = elevation
        h_max = float(h[mask < 1].max()); h_min = float(h[mask < 1].min())
        # CuPy's float32 -> uint8 conversion is CUDA's saturating cvt (NumPy's wraps): invalid cells outside the valid
        # range only matter through cv2's shifted border reads, but keep the reference's (GPU) semantics
        q = ((elevation.astype(np.float32) - np.float32(h_min)) * np.float32(255) / np.float32(h_max - h_min))
        h8 = np.clip(np.nan_to_num(np.trunc(q), nan=0.0), 0, 255).astype("uint8")
        dst = np.array(cv.inpaint(h8, mask, 1, cv.INPAINT_NS if method == "ns" else cv.INPAINT_TELEA))
        return (dst.astype(np.float32) * (h_max - h_min) / 255 + h_min).astype(np.float64)
    return elevation



# ---- semantic point-channel fusion ------------------------------------------------------------

def semantic_fuse(W, idx, valid, inside, feats, kinds, semantic_map, cnt_fused, alpha=0.5):
    """fusion/pointcloud_average.py:41-79, pointcloud_class_average.py, pointcloud_color.py:51-116 in NumPy, with the sums
    accumulated in float64 (the reference's float atomics are order-dependent; this is the order-free value they
    approximate).  idx/valid/inside: the write-back of CK.py:260-262; feats (N, k) float32; kinds: list of
    'average' | 'class_average' | 'color' per column; semantic_map (k, W, W) float32 is updated in place; cnt_fused (W*W,)
    = new_map[2] of the frame."""
    C = W * W
    sel = (valid > 0) & (inside > 0)
    cells = idx[sel].astype(np.int64)
    cnt = cnt_fused.reshape(-1).astype(np.float32)
    for k, kind in enumerate(kinds):
        layer = semantic_map[k].reshape(-1)
        f = feats[sel, k].astype(np.float32)
        if kind == "color":
            col = f.view(np.uint32)
            n = np.bincount(cells, minlength=C).astype(np.uint32)
            r = np.bincount(cells, weights=((col >> 16) & 0xFF).astype(np.float64), minlength=C).astype(np.uint32)
            g = np.bincount(cells, weights=((col >> 8) & 0xFF).astype(np.float64), minlength=C).astype(np.uint32)
            b = np.bincount(cells, weights=(col & 0xFF).astype(np.float64), minlength=C).astype(np.uint32)
            m = n > 0
            rgb = ((r[m] // n[m]) << 16) + ((g[m] // n[m]) << 8) + (b[m] // n[m])
            layer[m] = rgb.astype(np.uint32).view(np.float32)
            continue
        sums = np.bincount(cells, weights=f.astype(np.float64), minlength=C).astype(np.float32)
        m = cnt > 0
        if kind == "average":
            layer[m] = sums[m] / (np.float32(1) * cnt[m])
        else:
            prev = layer.copy()
            first = m & (prev == 0)
            layer[first] = sums[first] / (np.float32(1) * cnt[first])
            later = m & (prev != 0)
            layer[later] = (alpha * prev[later].astype(np.float64)
                            + (1 - alpha) * sums[later].astype(np.float64) / cnt[later].astype(np.float64)).astype(np.float32)
    return semantic_map


# ---- map initialiser: EM.py:899-922, map_initializer.py:25-62 ----------------------------------

def initialize_map_planes(param, state, points, center, method="cubic"):
    """NumPy restatement of initialize_map on a cleared map: griddata interpolation of `points` (x, y, z in the map frame),
    two Jacobi passes of dilation_filter_kernel (CK.py:392-449, dilation_size_initialize, mask = is_valid; the reference runs
    them in place) and update_upper_bound_with_valid_elevation (EM.py:428-432).  Returns the (7,W,W) state."""
    from scipy.interpolate import griddata
    W = param.cell_n
    st = state.copy()
    pts = np.array(points, np.float32, copy=True)
    idx = ((pts[:, :2] - np.asarray(center, np.float32)[:2].reshape(1, 2)) / np.float32(param.resolution) + W / 2).astype(np.int32)
    pts[:, :2] = idx.astype(np.float32); pts[:, 2] -= np.float32(center[2])
    vi = np.where(st[2] > 0.5)
    pidx = np.vstack([np.stack(vi).T.astype(np.float32), pts[:, :2]])
    vals = np.hstack([st[0][vi], pts[:, 2]])
    gx, gy = np.mgrid[0:W, 0:W]
    interp = griddata(pidx, vals, (gx, gy), method=method)
    ok = ~np.isnan(interp)
    st[0] = np.nan_to_num(interp).astype(np.float32)
    st[1] = np.where(ok, np.float32(param.initialized_variance), np.float32(param.initial_variance))
    st[2] = ok.astype(np.float32)
    k = int(param.dilation_size_initialize)
    for it in range(2 if k > 0 else 0):
        h, m = st[0].reshape(-1).copy(), st[2].reshape(-1).copy()
        nh, nm = h.copy(), m.copy()
        C = W * W
        for i in np.nonzero(m < 0.5)[0]:
            dist, near = 100.0, 0.0
            for dy in range(-k, k + 1):
                for dx in range(-k, k + 1):
                    j = i + W * dy + dx
                    if j < 0 or j >= C:
                        continue
                    ix, iy = divmod(j, W)
                    if ix <= 0 or ix >= W - 1 or iy <= 0 or iy >= W - 1:
                        continue
                    if m[j] > 0.5 and dx + dy < dist:
                        dist, near = dx + dy, h[j]
            if dist < 100:
                nh[i], nm[i] = near, 1.0
        st[0], st[2] = nh.reshape(W, W), nm.reshape(W, W)
    mask = st[2] > 0.5
    st[5] = np.where(mask, st[0], st[5]); st[6] = np.where(mask, 0.0, st[6])
    return st

# ---- the reference's own kernel source, compiled for the host ------------------------------

class RefKernelMap(_MapBase):
    """EM.py:316-391 driving oracle/_ref/libref_cpu_<tag>.so.  `parallel=False` executes the
    kernel bodies one element at a time in input order (a legal, deterministic interleaving of
    the reference kernel); `parallel=True` runs them on all host threads with real atomics
    (racy exactly like the GPU original) -- used only for timing."""

    def __init__(self, param, tag, parallel=False):
        super().__init__(param)
        from . import build_ref
        from .configs import ref_dict
        self.lib = C.CDLL(build_ref.build(ref_dict(param), tag=tag, gpu=False))      # tag None: named by parameter hash
        assert self.lib.ref_cell_n() == param.cell_n
        W = self.cell_n
        self.new_map = np.zeros((7, W, W), np.float32)
        self.mask_dummy = np.zeros((W, W), np.float32)
        self.parallel = int(parallel)
        self.zero = np.zeros(1, np.float32)

    def update_map_with_kernel(self, points_all, channels, R, t, position_noise, orientation_noise):
        pr = self.param
        W = self.cell_n
        L = self.lib
        self.new_map *= 0.0
        error = np.zeros(1, np.float32); error_cnt = np.zeros(1, np.float32)
        points_all = points_all[~np.isnan(points_all[:, :3]).any(axis=1)]            # EM.py:458
        points = np.ascontiguousarray(points_all[:, :3], np.float32)
        n = points.shape[0]
        R = np.ascontiguousarray(R, np.float32).reshape(9)
        t = (t - self.center).astype(np.float32)                                      # EM.py:333
        em = self.elevation_map
        L.ref_error_counting(C.c_longlong(n), _p(em), _p(points), _p(self.zero), _p(self.zero), _p(R), _p(t),
                             _p(self.new_map), _p(error), _p(error_cnt), C.c_int(self.parallel))
        if (pr.enable_drift_compensation and error_cnt[0] > pr.min_height_drift_cnt
                and (position_noise > pr.position_noise_thresh or orientation_noise > pr.orientation_noise_thresh)):
            self.mean_error = error / error_cnt
            self.additive_mean_error += self.mean_error
            if np.abs(self.mean_error) < pr.max_drift:
                em[0] += (self.mean_error * np.float32(pr.drift_compensation_alpha)).astype(np.float32)
        L.ref_add_points(C.c_longlong(n), _p(self.zero), _p(self.zero), _p(R), _p(t), _p(self.normal_map),
                         _p(points), _p(em), _p(self.new_map), C.c_int(self.parallel))
        L.ref_average_map(C.c_longlong(W * W), _p(self.new_map), _p(em), C.c_int(self.parallel))
        self.last_point_record = (points[:, 0].astype(np.int32), points[:, 1].astype(np.uint8),
                                  points[:, 2].astype(np.uint8))
        if pr.enable_overlap_clearance:
            self.clear_overlap_map(t)
        self.traversability_input *= 0.0
        mask = em[2] + em[6]
        L.ref_dilation_filter(C.c_longlong(W * W), _p(np.ascontiguousarray(em[5])), _p(mask),
                              _p(self.traversability_input), _p(self.mask_dummy), C.c_int(self.parallel))
        em[3][3:-3, 3:-3] = self.traversability_filter(self.traversability_input)
        self.update_normal(self.traversability_input)

    def traversability_filter(self, x):
        """TF.py:15-42 with torch on the CPU."""
        import torch
        import torch.nn.functional as F
        w1, w2, w3, wo = [torch.from_numpy(w) for w in self.weights]
        with torch.no_grad():
            e = torch.from_numpy(np.ascontiguousarray(x)).view(1, 1, *x.shape)
            o1 = F.conv2d(e, w1.view(4, 1, 3, 3), dilation=1)[:, :, 2:-2, 2:-2]
            o2 = F.conv2d(e, w2.view(4, 1, 3, 3), dilation=2)[:, :, 1:-1, 1:-1]
            o3 = F.conv2d(e, w3.view(4, 1, 3, 3), dilation=3)
            out = F.conv2d(torch.cat((o1, o2, o3), 1).abs(), wo.view(1, 12, 1, 1))
            return torch.exp(-out)[0, 0].numpy()

    # EM.py:393-410
    def clear_overlap_map(self, t):
        pr = self.param
        height_min = np.float32(t[2] - np.float32(pr.overlap_clear_range_z))
        height_max = np.float32(t[2] + np.float32(pr.overlap_clear_range_z))
        near = self.elevation_map[:, self.cell_min:self.cell_max, self.cell_min:self.cell_max]
        ok = ~np.logical_or(near[0] < height_min, near[0] > height_max)
        near[0] = np.where(ok, near[0], 0.0)
        near[1] = np.where(ok, near[1], np.float32(self.initial_variance))
        near[2] = np.where(ok, near[2], 0.0)
        ok = ~np.logical_or(near[5] < height_min, near[5] > height_max)
        near[5] = np.where(ok, near[5], 0.0)
        near[6] = np.where(ok, near[6], 0.0)

    def update_normal(self, dilated_map):
        W = self.cell_n
        self.normal_map *= 0.0
        self.lib.ref_normal_filter(C.c_longlong(W * W), _p(np.ascontiguousarray(dilated_map, np.float32)),
                                   _p(np.ascontiguousarray(self.elevation_map[2])), _p(self.normal_map),
                                   C.c_int(self.parallel))

    def min_filter(self, iteration_n):
        """plugins/min_filter.py:100-118 with the reference's in-place kernel."""
        W = self.cell_n
        h = np.ascontiguousarray(self.elevation_map[0]); m = np.ascontiguousarray(self.elevation_map[2])
        nh = h.copy(); nm = m.copy()
        for _ in range(iteration_n):
            self.lib.ref_min_filter(C.c_longlong(W * W), _p(h), _p(m), _p(nh), _p(nm), C.c_int(0))
            if (nm > 0.5).all():
                break
        return np.where(nm > 0.5, nh, np.nan)
