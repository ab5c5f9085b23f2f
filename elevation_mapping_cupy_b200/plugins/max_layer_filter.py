"""max_layer_filter plugin -- per-cell extremum over several prepared layers.

Behavioural mirror of the reference plugin (elevation_mapping_cupy/script/elevation_mapping_cupy/plugins/
max_layer_filter.py:12-108): every named layer is (1) patched where it is exactly zero (constant or another layer),
(2) optionally mirrored to 1-x, (3) optionally scaled, (4) optionally binarised against a threshold; the result is the
cell-wise max (or min) of what is left.  Only `float` entries of `scales` / `thresholds` are active, exactly like the
reference's isinstance tests.  Plain element-wise tensor algebra on the device; no library kernel involved.
"""
from typing import List, Optional

from .plugin_manager import PluginBase


class MaxLayerFilter(PluginBase):
    def __init__(self, cell_n: int = 100, layers: list = ["traversability"], reverse: list = [True],
                 min_or_max: str = "max", thresholds: list = [False], scales: list = [1.0],
                 default_value: float = 0.0, **kwargs):
        super().__init__()
        self.layers, self.reverse = layers, reverse
        self.thresholds, self.scales = thresholds, scales
        self.min_or_max = min_or_max
        self.default_value = default_value

    # one prepared operand, or None when the layer does not exist
    def _prepare(self, pos: int, name: str, lookup) -> Optional["object"]:
        import torch
        x = lookup(name)
        if x is None:
            return None
        fill = self.default_value
        if isinstance(fill, float):
            x = torch.where(x == 0.0, torch.full_like(x, fill), x)
        elif isinstance(fill, str):
            x = torch.where(x == 0, lookup(fill), x)
        if self.reverse[pos]:
            x = 1.0 - x
        if pos < len(self.scales) and isinstance(self.scales[pos], float):
            x = x * self.scales[pos]
        gate = self.thresholds[pos]
        if isinstance(gate, float):
            x = (x > gate).to(x.dtype)
        return x

    def __call__(self, elevation_map, layer_names: List[str], plugin_layers, plugin_layer_names: List[str],
                 semantic_map, semantic_layer_names: List[str], *args):
        import torch
        sem_names = semantic_layer_names or []

        def lookup(name):
            return self.get_layer_data(elevation_map, layer_names, plugin_layers, plugin_layer_names, semantic_map,
                                       sem_names, name)

        operands = [op for op in (self._prepare(k, n, lookup) for k, n in enumerate(self.layers)) if op is not None]
        if not operands:
            print("No layers are found, returning traversability!")
            if isinstance(self.default_value, float):
                return torch.full_like(elevation_map[0], float(self.default_value))
            return elevation_map[layer_names.index("traversability")]
        stacked = torch.stack(operands)
        reduce = torch.amin if self.min_or_max == "min" else torch.amax
        return reduce(stacked, dim=0)
