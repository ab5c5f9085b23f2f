"""Inpainting (plugins/inpainting.py:53-63) without a GPU: the sequential restatement of cv2's Telea fill
(oracle/telea_seq.py) and the PARALLEL round / event-order algorithm the CUDA kernel runs (csrc/emap_inpaint.cuh, replayed on
the CPU by tests/native/inpaint_host.cpp, which calls the very same inline functions) are both pinned against the cv2
binary in this image -- bit for bit."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _host_lib():
    src = os.path.join(ROOT, "tests", "native", "inpaint_host.cpp")
    out_dir = os.path.join(ROOT, "tests", "native", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libinpaint_host.so")
    hdr = os.path.join(ROOT, "elevation_mapping_cupy_b200", "csrc", "emap_inpaint.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, src], check=True)
    return C.CDLL(so)


def _run_host(L, img, mask, cap=4096):
    H, W = img.shape
    out = np.zeros_like(img)
    r, mj, nc = C.c_int(), C.c_int(), C.c_int()
    L.inpaint_host(img.ctypes.data_as(C.c_void_p), mask.ctypes.data_as(C.c_void_p), H, W, out.ctypes.data_as(C.c_void_p),
                   cap, C.byref(r), C.byref(mj), C.byref(nc), None, None)
    return out, r.value, mj.value, nc.value


def _terrain_case(W, rng, dropout=0.08):
    from elevation_mapping_cupy_b200 import workloads as wl
    xs = (np.arange(W) - W / 2) * 0.04
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    g = wl.terrain(X, Y).astype(np.float32)
    half = W * 0.02
    valid = (np.abs(np.arctan2(Y, X)) < 0.76) & (np.hypot(X, Y) > 0.1 * half) & (np.hypot(X, Y) < min(12.0, 0.9 * half))
    valid &= rng.random((W, W)) > dropout
    hmax, hmin = float(g[valid].max()), float(g[valid].min())
    img = np.clip((np.where(valid, g, 0) - hmin) * 255 / (hmax - hmin), 0, 255).astype(np.uint8)
    return np.ascontiguousarray(img), np.ascontiguousarray((~valid).astype(np.uint8))


def test_sequential_restatement_equals_cv2():
    from oracle.telea_seq import telea_radius1
    rng = np.random.default_rng(0)
    for trial, frac in enumerate([0.3, 0.69, 0.9, 0.995]):
        H, W = 22 + trial, 29
        img = rng.integers(0, 256, (H, W), dtype=np.uint8)
        mask = (rng.random((H, W)) < frac).astype(np.uint8)
        assert np.array_equal(telea_radius1(img, mask), cv2.inpaint(img, mask, 1, cv2.INPAINT_TELEA)), trial
    img, mask = _terrain_case(48, rng)          # smooth data: exact .5 ties in the rounding
    assert np.array_equal(telea_radius1(img, mask), cv2.inpaint(img, mask, 1, cv2.INPAINT_TELEA))


def test_parallel_rounds_equal_cv2_bit_for_bit():
    L = _host_lib()
    rng = np.random.default_rng(5)
    # the reference's own plugin test shape (tests/test_plugins.py): randn(7, 200, 200) -> here 202^2
    em = rng.standard_normal((7, 202, 202)).astype(np.float32)
    mask = np.ascontiguousarray((em[2] < 0.5).astype(np.uint8))
    h = em[0]
    hmax, hmin = float(h[mask < 1].max()), float(h[mask < 1].min())
    img = np.ascontiguousarray(np.clip((h - hmin) * 255 / (hmax - hmin), 0, 255).astype(np.uint8))
    got, rounds, max_j, nc = _run_host(L, img, mask)
    assert nc == 0 and np.array_equal(got, cv2.inpaint(img, mask, 1, cv2.INPAINT_TELEA)), (rounds, max_j)
    # random images, sparse to almost fully masked, odd sizes
    for trial, frac in enumerate([0.2, 0.5, 0.8, 0.97, 0.999]):
        H, W = 37 + 3 * trial, 53 - 2 * trial
        img = np.ascontiguousarray(rng.integers(0, 256, (H, W), dtype=np.uint8))
        mask = np.ascontiguousarray((rng.random((H, W)) < frac).astype(np.uint8))
        if mask.all():
            mask[H // 2, W // 2] = 0
        got, rounds, max_j, nc = _run_host(L, img, mask)
        assert nc == 0 and np.array_equal(got, cv2.inpaint(img, mask, 1, cv2.INPAINT_TELEA)), (trial, rounds, max_j)
    # depth-camera-like terrain map: one big hole hundreds of cells deep (many rounds) + pixel dropouts
    img, mask = _terrain_case(384, rng)
    got, rounds, max_j, nc = _run_host(L, img, mask)
    assert rounds > 200 and nc == 0
    assert np.array_equal(got, cv2.inpaint(img, mask, 1, cv2.INPAINT_TELEA)), (rounds, max_j)


def test_parallel_rounds_expanding_front_equal_cv2():
    """Config-D-like map: a small known blob in an empty image -- the front expands as smooth rings (T ties at every level of
    the popper chains: the cached event-order masks and the level-synchronous chain walks decide everything)."""
    L = _host_lib()
    W = 300
    img = np.zeros((W, W), np.uint8); mask = np.ones((W, W), np.uint8)
    yy, xx = np.mgrid[0:W, 0:W]
    blob = ((yy - 150) ** 2 + (xx - 140) ** 2 < 25 ** 2)
    img[blob] = (128 + 60 * np.sin(xx[blob] / 7.0) + 40 * np.cos(yy[blob] / 5.0)).astype(np.uint8); mask[blob] = 0
    got, rounds, max_j, nc = _run_host(L, np.ascontiguousarray(img), np.ascontiguousarray(mask))
    assert rounds > 150 and nc == 0
    assert np.array_equal(got, cv2.inpaint(img, mask, 1, cv2.INPAINT_TELEA)), (rounds, max_j)
