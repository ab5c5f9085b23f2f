"""Plugin layer API -- mirror of the reference's plugin manager
(elevation_mapping_cupy/script/elevation_mapping_cupy/plugins/plugin_manager.py:15-235):
same `PluginParams`, `PluginBase.__call__` signature, YAML schema (`enable, fill_nan,
is_height_layer, layer_name, type?, extra_params`, :134-151), `cell_n` injection (:128) and
arity-based dispatch (:193-225).  Layers are torch CUDA tensors instead of CuPy arrays; PyYAML
replaces ruamel.  The built-in plugins run libemap.so kernels through the `engine` handle the
manager passes to their constructors (user plugins may ignore it: every reference plugin takes **kwargs).
"""
from abc import ABC
from dataclasses import dataclass
import importlib
import inspect
from inspect import signature
from typing import Dict, List, Optional

import yaml


@dataclass
class PluginParams:
    name: str
    layer_name: str
    fill_nan: bool = False        # fill nan to invalid region
    is_height_layer: bool = False  # if this is a height layer


class PluginBase(ABC):
    """Base class of plugins (plugin_manager.py:23-108)."""

    def __init__(self, *args, **kwargs):
        pass

    def __call__(self, elevation_map, layer_names: List[str], plugin_layers, plugin_layer_names: List[str],
                 semantic_map=None, semantic_layer_names: List[str] = None, *args, **kwargs):
        """elevation_map layers: 0 elevation, 1 variance, 2 is_valid, 3 traversability, 4 time,
        5 upper_bound, 6 is_upper_bound.  Return a (cell_n, cell_n) device array."""
        pass

    def get_layer_data(self, elevation_map, layer_names, plugin_layers, plugin_layer_names, semantic_map,
                       semantic_layer_names, name) -> Optional[object]:
        """plugin_manager.py:71-108: a copy of the named layer, or None."""
        if name in layer_names:
            return elevation_map[layer_names.index(name)].clone()
        if name in plugin_layer_names:
            return plugin_layers[plugin_layer_names.index(name)].clone()
        if semantic_layer_names and name in semantic_layer_names:
            return semantic_map[semantic_layer_names.index(name)].clone()
        print(f"Could not find layer {name}!")
        return None


class PluginManager(object):
    """Manages the plugins (plugin_manager.py:111-235)."""

    def __init__(self, cell_n: int, engine=None, package: str = "elevation_mapping_cupy_b200.plugins"):
        self.cell_n = cell_n
        self.engine = engine
        self.package = package
        self.plugin_params: List[PluginParams] = []
        self.plugins = []
        self.layers = None
        self.layer_names: List[str] = []
        self.plugin_names: List[str] = []

    def init(self, plugin_params: List[PluginParams], extra_params: List[Dict]):
        self.plugin_params = plugin_params
        self.plugins = []
        for param, extra_param in zip(plugin_params, extra_params):
            m = importlib.import_module("." + param.name, package=self.package)
            for name, obj in inspect.getmembers(m):
                if inspect.isclass(obj) and issubclass(obj, PluginBase) and name != "PluginBase":
                    extra_param = dict(extra_param or {})
                    extra_param["cell_n"] = self.cell_n          # plugin_manager.py:128
                    extra_param["engine"] = self.engine
                    self.plugins.append(obj(**extra_param))
        self.layers = None                                        # allocated lazily on the engine's device
        self.layer_names = self.get_layer_names()
        self.plugin_names = self.get_plugin_names()

    def _ensure_layers(self, like):
        if self.layers is None:
            import torch
            self.layers = torch.zeros((len(self.plugins), self.cell_n, self.cell_n), dtype=torch.float32,
                                      device=like.device)

    def load_plugin_settings(self, file_path: str):
        print("Start loading plugins...")
        with open(file_path, "r") as f:
            cfg = yaml.safe_load(f) or {}
        plugin_params, extra_params = [], []
        for k, v in cfg.items():
            if v["enable"]:
                plugin_params.append(PluginParams(name=k if "type" not in v else v["type"], layer_name=v["layer_name"],
                                                  fill_nan=v["fill_nan"], is_height_layer=v["is_height_layer"]))
                extra_params.append(v.get("extra_params", {}))
        self.init(plugin_params, extra_params)
        print("Loaded plugins are ", *self.plugin_names)

    def get_layer_names(self):
        return [obj.layer_name for obj in self.plugin_params]

    def get_plugin_names(self):
        return [obj.name for obj in self.plugin_params]

    def get_plugin_index_with_name(self, name: str) -> int:
        try:
            return self.plugin_names.index(name)
        except Exception as e:
            print("Error with plugin {}: {}".format(name, e))
            return None

    def get_layer_index_with_name(self, name: str) -> int:
        try:
            return self.layer_names.index(name)
        except Exception as e:
            print("Error with layer {}: {}".format(name, e))
            return None

    def update_with_name(self, name: str, elevation_map, layer_names: List[str], semantic_map=None,
                         semantic_params=None, rotation=None, elements_to_shift={}):
        """plugin_manager.py:181-225: dispatch by the number of parameters of the plugin's __call__."""
        idx = self.get_layer_index_with_name(name)
        if idx is not None and idx < len(self.plugins):
            self._ensure_layers(elevation_map)
            n_param = len(signature(self.plugins[idx]).parameters)
            if n_param == 5:
                out = self.plugins[idx](elevation_map, layer_names, self.layers, self.layer_names)
            elif n_param == 7:
                out = self.plugins[idx](elevation_map, layer_names, self.layers, self.layer_names, semantic_map,
                                        semantic_params)
            elif n_param == 8:
                out = self.plugins[idx](elevation_map, layer_names, self.layers, self.layer_names, semantic_map,
                                        semantic_params, rotation)
            else:
                out = self.plugins[idx](elevation_map, layer_names, self.layers, self.layer_names, semantic_map,
                                        semantic_params, rotation, elements_to_shift)
            self.layers[idx] = out

    def get_map_with_name(self, name: str):
        idx = self.get_layer_index_with_name(name)
        if idx is not None:
            return self.layers[idx]

    def get_param_with_name(self, name: str) -> PluginParams:
        idx = self.get_layer_index_with_name(name)
        if idx is not None:
            return self.plugin_params[idx]
