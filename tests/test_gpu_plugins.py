"""GPU parity of the plugin stencils and of the drop-in API surface (reference test_elevation_mapping.py /
test_plugins.py shapes: 202^2 map, randn layers)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _mk(param):
    from elevation_mapping_cupy_b200.elevation_mapping import ElevationMap
    return ElevationMap(param)


def _random_state(W, seed, valid_frac=0.5):
    rng = np.random.default_rng(seed)
    m = np.zeros((7, W, W), np.float32)
    m[0] = rng.standard_normal((W, W)).astype(np.float32)
    m[1] = 0.05
    m[2] = (rng.random((W, W)) < valid_frac).astype(np.float32)
    m[3] = 1.0
    m[5] = m[0]
    return m


@pytest.mark.parametrize("k,iters,valid_frac", [(1, 30, 0.5), (2, 3, 0.05), (1, 5, 0.001)])
def test_min_filter_matches_oracle(oracle_mod, k, iters, valid_frac):
    import torch
    from elevation_mapping_cupy_b200.parameter import core_parameter
    from elevation_mapping_cupy_b200.plugins.min_filter import MinFilter
    p = core_parameter(202)
    em = _mk(p)
    st = _random_state(202, 5, valid_frac)
    em.set_state(st)
    plug = MinFilter(cell_n=202, dilation_size=k, iteration_n=iters, engine=em)
    out = plug(em.elevation_map, em.layer_names, None, []).cpu().numpy()
    ref, _ = oracle_mod.min_filter(202, k, iters, st[0], st[2])
    assert np.array_equal(np.isnan(out), np.isnan(ref))
    assert np.array_equal(np.nan_to_num(out), np.nan_to_num(ref))


def test_smooth_filter_matches_oracle_and_scipy(oracle_mod):
    from scipy import ndimage
    from elevation_mapping_cupy_b200.parameter import core_parameter
    from elevation_mapping_cupy_b200.plugins.smooth_filter import SmoothFilter
    p = core_parameter(202)
    em = _mk(p)
    st = _random_state(202, 7)
    em.set_state(st)
    plug = SmoothFilter(cell_n=202, input_layer_name="elevation", engine=em)
    out = plug(em.elevation_map, em.layer_names, None, []).cpu().numpy()
    ref = oracle_mod.smooth(202, st[0])
    assert np.array_equal(out, ref)
    sp = ndimage.uniform_filter(ndimage.uniform_filter(st[0], size=3), size=3)   # smooth_filter.py:57-58 on the host
    assert np.abs(out - sp).max() < 1e-6


def test_reference_api_surface_smoke():
    """The call sequence of the reference's own test (test_elevation_mapping.py:46-118), minus the
    out-of-scope semantic / polygon / initialiser calls."""
    import torch
    from elevation_mapping_cupy_b200.parameter import Parameter
    from elevation_mapping_cupy_b200 import workloads as wl
    p = Parameter(); p.update()
    em = _mk(p)
    assert len(em.layer_names) == em.elevation_map.shape[0]
    pts, R, t = wl.reference_test_cloud(0, n_ch=5)
    em.input_pointcloud(torch.from_numpy(pts).cuda(), ["x", "y", "z", "a", "b"], R, t, 0, 0)
    em.input(pts.astype(np.float64), ["x", "y", "z", "a", "b"], R, t, 0, 0)
    em.update_normal(em.elevation_map[0].contiguous())
    for i in range(20):
        em.move_to(np.array([i * 0.01, i * 0.02, i * 0.01]), np.random.rand(3, 3))
    data = np.zeros((em.cell_n - 2, em.cell_n - 2), dtype=np.float32)
    for layer in ["elevation", "variance", "traversability", "min_filter", "smooth"]:
        em.get_map_with_name_ref(layer, data)
        assert em.exists_layer(layer)
    pos = np.random.rand(1, 3)
    em.get_position(pos)
    em.clear()
    em.move(np.random.rand(3))
    assert not em.exists_layer("nonexistent")
    with pytest.raises(KeyError):
        em._ptr("nonexistent")


def test_move_and_shift_match_oracle(oracle_mod):
    from elevation_mapping_cupy_b200.parameter import core_parameter
    p = core_parameter(202)
    em = _mk(p); om = oracle_mod.OracleElevationMap(p)
    st = _random_state(202, 11)
    em.set_state(st); om.elevation_map = st.copy()
    rng = np.random.default_rng(3)
    for i in range(12):
        pos = rng.uniform(-0.6, 0.6, 3) * (i + 1) / 4
        if i % 3 == 2:
            d = rng.uniform(-0.2, 0.2, 3)
            em.move(d); om.move(d)
        else:
            em.move_to(pos, np.eye(3)); om.move_to(pos, np.eye(3))
        state, _ = em.get_state()
        assert np.array_equal(state, om.elevation_map), f"step {i}"
        assert np.array_equal(em.center, om.center)
    em.clear(); om.clear()
    state, _ = em.get_state()
    assert np.array_equal(state, om.elevation_map)
