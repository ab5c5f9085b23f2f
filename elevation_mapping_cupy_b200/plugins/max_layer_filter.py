"""max_layer_filter plugin: per-cell max (or min) over a list of layers after optional default / reverse / scale /
threshold steps (reference: elevation_mapping_cupy/script/elevation_mapping_cupy/plugins/max_layer_filter.py:12-108).
Purely element-wise array algebra in the reference (cupy); the same expressions on torch CUDA tensors here."""
from typing import List

from .plugin_manager import PluginBase


class MaxLayerFilter(PluginBase):
    def __init__(self, cell_n: int = 100, layers: list = ["traversability"], reverse: list = [True],
                 min_or_max: str = "max", thresholds: list = [False], scales: list = [1.0],
                 default_value: float = 0.0, **kwargs):
        super().__init__()
        self.layers = layers
        self.reverse = reverse
        self.min_or_max = min_or_max
        self.thresholds = thresholds
        self.scales = scales
        self.default_value = default_value

    def __call__(self, elevation_map, layer_names: List[str], plugin_layers, plugin_layer_names: List[str],
                 semantic_map, semantic_layer_names: List[str], *args):
        import torch
        semantic_layer_names = semantic_layer_names or []
        layers = []
        for it, name in enumerate(self.layers):
            layer = self.get_layer_data(elevation_map, layer_names, plugin_layers, plugin_layer_names, semantic_map,
                                        semantic_layer_names, name)
            if layer is None:
                continue
            if isinstance(self.default_value, float):
                layer = torch.where(layer == 0.0, torch.full_like(layer, float(self.default_value)), layer)
            elif isinstance(self.default_value, str):
                default_layer = self.get_layer_data(elevation_map, layer_names, plugin_layers, plugin_layer_names,
                                                    semantic_map, semantic_layer_names, self.default_value)
                layer = torch.where(layer == 0, default_layer, layer)
            if self.reverse[it]:
                layer = 1.0 - layer
            if len(self.scales) > it and isinstance(self.scales[it], float):
                layer = layer * float(self.scales[it])
            if isinstance(self.thresholds[it], float):
                layer = torch.where(layer > float(self.thresholds[it]), torch.ones_like(layer), torch.zeros_like(layer))
            layers.append(layer)
        if len(layers) == 0:
            print("No layers are found, returning traversability!")
            if isinstance(self.default_value, float):
                return torch.ones_like(elevation_map[0]) * float(self.default_value)
            return elevation_map[layer_names.index("traversability")]
        result = torch.stack(layers, dim=0)
        return result.min(dim=0).values if self.min_or_max == "min" else result.max(dim=0).values
