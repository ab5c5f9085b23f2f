"""CPU tests of the host side: C-ABI library exports, Parameter mirror, plugin manager, sharded-frame host
logic over gloo (world_size 2).  No compute call is made on the library (no GPU here)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol():
    from elevation_mapping_cupy_b200 import _lib
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "emap.h")).read()
    declared = set(re.findall(r"\b(emap_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/emap.h but not exported by libemap.so"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_abi_rejects_bad_config_without_gpu():
    from elevation_mapping_cupy_b200 import _lib
    from elevation_mapping_cupy_b200.elevation_mapping import config_from_parameter
    from elevation_mapping_cupy_b200.parameter import core_parameter
    lib = _lib.load()
    cfg = config_from_parameter(core_parameter(256))
    cfg.cell_n = 4096                       # beyond the fp16-exact clamp of CK.py:22-25
    h = C.c_void_p()
    assert lib.emap_create(C.byref(cfg), 0, C.byref(h)) == -1
    assert b"cell_n" in lib.emap_last_error(None)
    cfg.cell_n = 256; cfg.abi_version = 99
    assert lib.emap_create(C.byref(cfg), 0, C.byref(h)) == -1
    assert lib.emap_destroy(None) == 0
    assert lib.emap_cell_n(None) == -1


def test_no_cpu_fallback_when_extension_missing(tmp_path):
    """the product path must fail loudly without libemap.so"""
    code = ("import sys; sys.path.insert(0, %r); from elevation_mapping_cupy_b200 import _lib; "
            "_lib.LIB_PATH = %r; _lib._lib = None\n"
            "try:\n    _lib.load()\nexcept ImportError as e:\n    print('RAISED', 'no CPU fallback' in str(e))\n"
            % (ROOT, str(tmp_path / "missing.so")))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True).stdout
    assert "RAISED True" in out


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "elevation_mapping_cupy_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt.replace(
                    "oracle/emap_oracle.c", "").replace("oracle/emap_oracle.c fix32", ""), f


def test_parameter_mirror():
    from elevation_mapping_cupy_b200.parameter import Parameter, core_parameter
    p = Parameter()
    p.update()
    assert p.cell_n == 202 and p.true_cell_n == 200                 # parameter.py:282-289
    p.set_value("resolution", 0.1)
    assert p.get_value("resolution") == 0.1                         # test_parameter.py:5-17
    assert "resolution" in p.get_names() and len(p.get_names()) == len(p.get_types())
    b = core_parameter(1024)
    assert b.cell_n == 1024 and abs(b.map_length - 40.88) < 1e-9
    assert b.drift_compensation_variance_inlier == 0.1              # the YAML key is misspelt upstream
    p.load_weights(p.weight_file)
    assert p.w1.shape == (4, 1, 3, 3) and p.w_out.shape == (1, 12, 1, 1)


def test_plugin_manager_yaml_and_dispatch(tmp_path):
    torch = pytest.importorskip("torch")
    from elevation_mapping_cupy_b200.plugins.plugin_manager import PluginManager, PluginBase
    plug = tmp_path / "userplug"
    plug.mkdir()
    (plug / "__init__.py").write_text("")
    (plug / "five.py").write_text(
        "from elevation_mapping_cupy_b200.plugins.plugin_manager import PluginBase\n"
        "class Five(PluginBase):\n"
        "    def __init__(self, cell_n=1, add=0.0, **kw):\n        super().__init__(); self.add = add; self.cell_n = cell_n\n"
        "    def __call__(self, elevation_map, layer_names, plugin_layers, plugin_layer_names, *args):\n"
        "        return elevation_map[0] + self.add\n")
    (plug / "eight.py").write_text(
        "from elevation_mapping_cupy_b200.plugins.plugin_manager import PluginBase\n"
        "class Eight(PluginBase):\n"
        "    def __init__(self, cell_n=1, **kw):\n        super().__init__()\n"
        "    def __call__(self, elevation_map, layer_names, plugin_layers, plugin_layer_names, semantic_map, semantic_layer_names, rotation, *args):\n"
        "        return plugin_layers[plugin_layer_names.index('five_layer')] * 2 + float(rotation[0][0])\n")
    cfgf = tmp_path / "plugins.yaml"
    cfgf.write_text(
        "five:\n  enable: True\n  fill_nan: True\n  is_height_layer: True\n  layer_name: five_layer\n  extra_params:\n    add: 1.5\n"
        "second:\n  type: eight\n  enable: True\n  fill_nan: False\n  is_height_layer: False\n  layer_name: eight_layer\n  extra_params: {}\n"
        "off:\n  type: five\n  enable: False\n  fill_nan: False\n  is_height_layer: False\n  layer_name: off_layer\n  extra_params: {}\n")
    sys.path.insert(0, str(tmp_path))
    try:
        pm = PluginManager(cell_n=8, package="userplug")
        pm.load_plugin_settings(str(cfgf))
        assert pm.layer_names == ["five_layer", "eight_layer"] and pm.plugin_names == ["five", "eight"]
        assert pm.plugins[0].cell_n == 8                            # cell_n injected (plugin_manager.py:128)
        em = torch.arange(7 * 64, dtype=torch.float32).reshape(7, 8, 8)
        names = ["elevation", "variance", "is_valid", "traversability", "time", "upper_bound", "is_upper_bound"]
        pm.update_with_name("five_layer", em, names)
        pm.update_with_name("eight_layer", em, names, None, [], np.eye(3) * 3.0, {})
        assert torch.equal(pm.get_map_with_name("five_layer"), em[0] + 1.5)
        assert torch.equal(pm.get_map_with_name("eight_layer"), (em[0] + 1.5) * 2 + 3.0)
        assert pm.get_param_with_name("five_layer").fill_nan is True
        assert pm.get_layer_index_with_name("nope") is None
    finally:
        sys.path.remove(str(tmp_path))


def test_reduce_plan():
    from elevation_mapping_cupy_b200 import sharded
    plan = sharded.reduce_plan([(0x10, 8, 3), (0x20, 2, 0), (0x30, 4, 1), (0x40, 4, 2)])
    assert plan == [(0x10, 8, "int32", "SUM"), (0x20, 2, "int64", "SUM"), (0x30, 4, "int64", "MAX"),
                    (0x40, 4, "int32", "MIN")]


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from elevation_mapping_cupy_b200 import sharded
dist.init_process_group("gloo")
rk, ws = dist.get_rank(), dist.get_world_size()
# (1) global point offsets from per-rank counts
off, total = sharded.global_point_offsets(1000 * (rk + 1))
assert (off, total) == (sum(1000 * (r + 1) for r in range(rk)), sum(1000 * (r + 1) for r in range(ws))), (off, total)
# (2) the exchange reductions reproduce the single-process accumulators:
rng = np.random.default_rng(7)
C = 257
full = {"cnt": rng.integers(0, 5, (ws, C)), "sum": rng.integers(-2**40, 2**40, (ws, C)),
        "last": rng.integers(0, 2**62, (ws, C)), "ukey": rng.integers(-2**31, 2**31 - 1, (ws, C))}
mine = [(torch.from_numpy(full["cnt"][rk].astype(np.int32)), "SUM"), (torch.from_numpy(full["sum"][rk].astype(np.int64)), "SUM"),
        (torch.from_numpy(full["last"][rk].astype(np.int64)), "MAX"), (torch.from_numpy(full["ukey"][rk].astype(np.int32)), "MIN")]
sharded.all_reduce_buffers(mine)
assert np.array_equal(mine[0][0].numpy(), full["cnt"].sum(0).astype(np.int32))
assert np.array_equal(mine[1][0].numpy(), full["sum"].sum(0))
assert np.array_equal(mine[2][0].numpy(), full["last"].max(0))
assert np.array_equal(mine[3][0].numpy(), full["ukey"].min(0).astype(np.int32))
dist.barrier()
if rk == 0:
    print("GLOO_OK")
dist.destroy_process_group()
"""


def test_sharded_host_logic_gloo_world2(tmp_path):
    pytest.importorskip("torch")
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert "GLOO_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_c_client_compiles_against_the_header(tmp_path):
    """include/emap.h is plain C: a C99 client (examples/c_api_demo.c) compiles and links against libemap.so."""
    from elevation_mapping_cupy_b200 import _lib
    pkg = os.path.dirname(_lib.LIB_PATH)
    exe = tmp_path / "c_api_demo"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "c_api_demo.c"), "-L", pkg, "-lemap", f"-Wl,-rpath,{pkg}",
                        "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    import shutil
    if shutil.which("nvidia-smi") is None:          # no GPU here: the client must fail loudly through the error channel
        assert run.returncode != 0 and "emap_create" in run.stderr
