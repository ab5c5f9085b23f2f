"""Shared helpers of the parity tests."""
import numpy as np

LAYERS = ["elevation", "variance", "is_valid", "traversability", "time", "upper_bound", "is_upper_bound"]


def compare_state(eng_map, eng_normal, om, trav_tol=2e-6, exact_layers=(0, 1, 2, 4, 5, 6), label=""):
    """CUDA engine state vs oracle state.  By construction (fixed-point accumulation, pinned
    contractions) every layer except traversability is expected BIT-identical; traversability goes
    through expf (CUDA vs glibc) and is compared to `trav_tol`."""
    for li in exact_layers:
        a, b = eng_map[li], om.elevation_map[li]
        if not np.array_equal(a, b):
            bad = np.argwhere(a != b)
            r, c = bad[0]
            raise AssertionError(f"{label} layer {LAYERS[li]}: {len(bad)} cells differ, first at ({r},{c}): "
                                 f"engine {a[r, c]!r} oracle {b[r, c]!r}; max abs diff {np.nanmax(np.abs(a - b))}")
    d = np.abs(eng_map[3] - om.elevation_map[3])
    assert d.max() <= trav_tol, f"{label} traversability max diff {d.max()}"
    assert np.array_equal(eng_normal, om.normal_map), f"{label} normals differ: max {np.abs(eng_normal - om.normal_map).max()}"
