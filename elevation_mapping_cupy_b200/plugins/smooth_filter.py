"""smooth_filter plugin: two passes of a 3x3 uniform filter with reflect borders
(reference: elevation_mapping_cupy/script/elevation_mapping_cupy/plugins/smooth_filter.py:12-59,
which calls cupyx.scipy.ndimage.uniform_filter twice).  Runs in libemap.so (`emap_smooth_filter`)."""
from typing import List

from .plugin_manager import PluginBase
from ._engine import require_engine, as_plane, sync_in


class SmoothFilter(PluginBase):
    def __init__(self, cell_n: int = 100, input_layer_name: str = "elevation", engine=None, **kwargs):
        super().__init__()
        self.input_layer_name = input_layer_name
        self.engine = engine

    def __call__(self, elevation_map, layer_names: List[str], plugin_layers, plugin_layer_names: List[str], *args):
        import torch
        eng = require_engine(self.engine, "SmoothFilter")
        if self.input_layer_name in layer_names:
            h = elevation_map[layer_names.index(self.input_layer_name)]
        elif self.input_layer_name in plugin_layer_names:
            h = plugin_layers[plugin_layer_names.index(self.input_layer_name)]
        else:
            print("layer name {} was not found. Using elevation layer.".format(self.input_layer_name))
            h = elevation_map[0]
        h = as_plane(h)
        out = torch.empty_like(h)
        sync_in()
        eng._check(eng._L.emap_smooth_filter(eng._h, h.data_ptr(), out.data_ptr()))
        eng.synchronize()
        return out
