"""Point-channel semantic layers of the map, fused by libemap.so inside the frame.

Mirrors the point-cloud half of the reference class
(elevation_mapping_cupy/script/elevation_mapping_cupy/semantic_map.py:13-259: layer bookkeeping, `get_fusion`,
`get_matching_fusion`, `get_indices_fusion`, `update_layers_pointcloud`) for the three fusion algorithms that ride on the
frame's point record: `average` (fusion/pointcloud_average.py), `class_average` (fusion/pointcloud_class_average.py, the
reference's default) and `color` (fusion/pointcloud_color.py).  The kernels are csrc/emap_semantic.cuh; this class only
decides which column of the cloud goes to which layer and hands the layer buffer to the engine
(`emap_semantic_configure`).  Image fusion, bayesian / class-max fusion and the feature extractors are out of scope.
"""
import ctypes as C
import re
from typing import Dict, List

import numpy as np

KINDS = {"average": 0, "class_average": 1, "color": 2}


class SemanticMap:
    def __init__(self, param, engine=None):
        self.param = param
        self.engine = engine
        self.layer_specs_points: Dict[str, str] = {}
        self.layer_names: List[str] = []
        self.semantic_map = None           # torch (n_layers, W, W) float32 on the engine's device, grown by add_layer
        self._configured = None
        self.elements_to_shift = {}

    # ---- layer bookkeeping (semantic_map.py:47-99) ----------------------------------------------------------------
    def _alloc(self, n):
        import torch
        W = self.param.cell_n
        return torch.zeros((n, W, W), dtype=torch.float32, device=f"cuda:{self.engine.device}")

    def clear(self):
        if self.semantic_map is not None:
            self.semantic_map.zero_()

    def add_layer(self, name):
        import torch
        if name in self.layer_names:
            return
        self.layer_names.append(name)
        new = self._alloc(1)
        self.semantic_map = new if self.semantic_map is None else torch.cat([self.semantic_map, new], 0)
        self._configured = None            # the buffer moved: re-register with the engine

    def get_matching_fusion(self, channel: str, fusion_algs: Dict[str, str]):
        for pattern, alg in fusion_algs.items():          # semantic_map.py:186-191
            if re.match(f"^{pattern}$", channel):
                return alg
        return None

    def get_fusion(self, channels: List[str], channel_fusions: Dict[str, str], layer_specs: Dict[str, str]):
        """semantic_map.py:150-184: the fusion algorithm of every channel (regex match, then `default`)."""
        fusion_list, process_channels = [], []
        for channel in channels:
            if channel not in layer_specs:
                matched = self.get_matching_fusion(channel, channel_fusions)
                if matched is None:
                    if "default" in channel_fusions:
                        matched = channel_fusions["default"]
                        print(f"[WARNING] Layer {channel} not found in layer_specs. Using {matched} algorithm as default.")
                    else:
                        print(f"[WARNING] Layer {channel} not found in layer_specs ({layer_specs}) and no default fusion is configured. Skipping.")
                        continue
                layer_specs[channel] = matched
            fusion_list.append(layer_specs[channel])
            process_channels.append(channel)
        return process_channels, fusion_list

    # ---- the frame hook (semantic_map.py:223-259) --------------------------------------------------------------------
    def configure_for(self, channels: List[str]):
        """Called by ElevationMap.input_pointcloud with the cloud's channel names (x, y, z first): registers the feature
        columns with the engine so that the frame fuses them.  Idempotent for an unchanged channel list."""
        feats = list(channels[3:])
        key = tuple(channels)
        if self._configured == key:
            return
        eng = self.engine
        if not feats:
            eng._check(eng._L.emap_semantic_configure(eng._h, 0, None, None, None, None, 0, 0.0))
            self._configured = key
            return
        fusions = getattr(self.param, "pointcloud_channel_fusions", None) or {"rgb": "color", "default": "class_average"}
        process, fusion = self.get_fusion(feats, fusions, self.layer_specs_points)
        cols, kinds, layers = [], [], []
        for ch, alg in zip(process, fusion):
            if alg not in KINDS:
                raise NotImplementedError(f"point-cloud fusion '{alg}' (channel {ch}) is not implemented; available: {sorted(KINDS)}")
            if ch not in self.layer_names:
                print(f"Layer {ch} not found, adding it to the semantic map")
                self.add_layer(ch)
            cols.append(3 + feats.index(ch)); kinds.append(KINDS[alg]); layers.append(self.layer_names.index(ch))
        n = len(cols)
        if n == 0:                          # no channel has a fusion: nothing to fuse
            eng._check(eng._L.emap_semantic_configure(eng._h, 0, None, None, None, None, 0, 0.0))
            self._configured = key
            return
        arr = lambda v: (C.c_int32 * n)(*v)
        eng._check(eng._L.emap_semantic_configure(eng._h, n, arr(cols), arr(kinds), arr(layers),
                                                  C.c_void_p(self.semantic_map.data_ptr()), len(self.layer_names),
                                                  float(getattr(self.param, "average_weight", 0.5))))
        self._configured = tuple(channels)

    # ---- queries (semantic_map.py:318-385) ----------------------------------------------------------------------------
    def get_map_with_name(self, name):
        return self.semantic_map[self.layer_names.index(name)][1:-1, 1:-1].clone()

    def get_index(self, name):
        return self.layer_names.index(name) if name in self.layer_names else -1

    def shift_map_xy(self, shift_value):
        """semantic_map.py:131-148: roll + zero padding, applied with the elevation map's shift."""
        import torch
        if self.semantic_map is None or (shift_value[0] == 0 and shift_value[1] == 0):
            return
        sx, sy = int(shift_value[0]), int(shift_value[1])
        m = torch.roll(self.semantic_map, (sx, sy), dims=(1, 2))
        if sx > 0: m[:, :sx, :] = 0.0
        elif sx < 0: m[:, sx:, :] = 0.0
        if sy > 0: m[:, :, :sy] = 0.0
        elif sy < 0: m[:, :, sy:] = 0.0
        self.semantic_map.copy_(m)
