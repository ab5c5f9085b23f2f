"""Helpers shared by the built-in plugins: call a libemap.so stencil on torch CUDA tensors.
Ordering with the caller's torch stream is by stream waits (ElevationMap._after_framework / _before_framework), never a host sync."""
import ctypes as C


def require_engine(engine, who):
    if engine is None:
        raise RuntimeError(f"{who} runs libemap.so kernels and needs the ElevationMap engine "
                           "(PluginManager(cell_n, engine=elevation_map)); there is no CPU fallback")
    return engine


def as_plane(t):
    """contiguous float32 CUDA (W,W) tensor"""
    import torch
    if not t.is_cuda:
        raise TypeError("plugin layers must be CUDA tensors")
    return t.to(torch.float32).contiguous()
