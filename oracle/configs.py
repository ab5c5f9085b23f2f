"""oracle/configs.py -- TEST INFRASTRUCTURE ONLY.

Parameter sets for which oracle/build_ref.py prebuilds the reference's kernel source
(constants are baked into the source, as the reference does with string.Template)."""
from elevation_mapping_cupy_b200.parameter import Parameter, core_parameter


def ref_dict(param: Parameter, min_filter_dilation_size=1):
    """The python-typed values ElevationMap.compile_kernels passes to the kernel
    factories (elevation_mapping.py:240-282)."""
    keys = ["resolution", "cell_n", "sensor_noise_factor", "mahalanobis_thresh", "outlier_variance",
            "wall_num_thresh", "max_ray_length", "cleanup_step", "min_valid_distance", "max_height_range",
            "cleanup_cos_thresh", "ramped_height_range_a", "ramped_height_range_b", "ramped_height_range_c",
            "enable_edge_sharpen", "enable_visibility_cleanup", "drift_compensation_variance_inlier",
            "traversability_inlier", "max_variance", "initial_variance", "dilation_size"]
    d = {k: getattr(param, k) for k in keys}
    d["min_filter_dilation_size"] = min_filter_dilation_size
    d["rce_threshold"] = 1.1          # tests/plugin_config.yaml of the reference
    return d


# parameters under which the drift path (elevation_mapping.py:346-357) triggers with a few thousand points
DRIFT_OVERRIDES = dict(traversability_inlier=0.0, drift_compensation_variance_inlier=10.0, min_height_drift_cnt=10)


def _default202():
    p = Parameter()
    p.update()
    return p


NAMED_PARAMS = {
    "default202": _default202,                      # dataclass defaults, 8 m / 0.04 m -> 202^2 (reference test shape)
    "core130": lambda: core_parameter(130),         # small map for the committed golden fixtures
    "drift130": lambda: core_parameter(130, **DRIFT_OVERRIDES),   # drift compensation fires on small test clouds
    "core202": lambda: core_parameter(202),         # deployed core_param.yaml values
    "core256": lambda: core_parameter(256),         # BASELINE config A
    "core512": lambda: core_parameter(512),         # BASELINE config E
    "core1024": lambda: core_parameter(1024),       # BASELINE config B / C
    "core2048": lambda: core_parameter(2048),       # BASELINE config D
}

REF_CONFIGS = {name: ref_dict(f()) for name, f in NAMED_PARAMS.items()}
