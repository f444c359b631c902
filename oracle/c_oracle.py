"""TEST INFRASTRUCTURE -- ctypes wrapper of oracle/_build/liboracle.so (the plain-C restatement of the reference's
Numba engines).  Used as the faster checker at larger sizes and as the `cpu_baseline` / `--impl reference` arm of
bench.py.  Never imported by xdem_b200/."""

from __future__ import annotations

import ctypes
import os

import numpy as np

from . import build_oracle
from .terrain_oracle import SURFACE_ATTRS

_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        path = build_oracle.LIB if os.path.exists(build_oracle.LIB) else build_oracle.build()
        _lib = ctypes.CDLL(path)
        _lib.xo_num_threads.restype = ctypes.c_int
    return _lib


def num_threads() -> int:
    return int(lib().xo_num_threads())


def surface_attributes(dem: np.ndarray, resolution: float, attrs: list[str], surface_fit: str = "Florinsky",
                       curv_method: str = "geometric", degrees: bool = False, clip_hillshade: bool = False,
                       hillshade_azimuth: float = 315.0, hillshade_altitude: float = 45.0,
                       hillshade_z_factor: float = 1.0, nthreads: int = 0) -> np.ndarray:
    dem = np.ascontiguousarray(dem)
    assert dem.dtype in (np.float32, np.float64)
    H, W = dem.shape
    out = np.empty((len(attrs), H, W), dtype=dem.dtype)
    planes = (ctypes.c_void_p * 10)()
    mask = 0
    for i, a in enumerate(attrs):
        k = SURFACE_ATTRS.index(a)
        mask |= 1 << k
        planes[k] = out[i].ctypes.data
    fit = {"horn": 0, "zevenbergthorne": 1, "florinsky": 2}[surface_fit.lower()]
    rc = lib().xo_surface(ctypes.c_void_p(dem.ctypes.data), int(dem.dtype == np.float64), ctypes.c_int64(H),
                          ctypes.c_int64(W), ctypes.c_double(resolution), fit,
                          int(curv_method.lower() == "directional"), ctypes.c_uint32(mask), int(degrees),
                          int(clip_hillshade), ctypes.c_double(hillshade_azimuth), ctypes.c_double(hillshade_altitude),
                          ctypes.c_double(hillshade_z_factor), planes, int(nthreads))
    assert rc == 0
    return out


def windowed_indexes(dem: np.ndarray, window_size: int, attrs: list[str], tri_method: str = "Riley",
                     nthreads: int = 0) -> np.ndarray:
    dem = np.ascontiguousarray(dem, dtype=np.float32)
    H, W = dem.shape
    order = ["topographic_position_index", "terrain_ruggedness_index", "roughness"]
    out = np.empty((len(attrs), H, W), dtype=np.float32)
    planes = (ctypes.c_void_p * 3)()
    mask = 0
    for i, a in enumerate(attrs):
        k = order.index(a)
        mask |= 1 << k
        planes[k] = out[i].ctypes.data
    rc = lib().xo_windowed_f32(ctypes.c_void_p(dem.ctypes.data), ctypes.c_int64(H), ctypes.c_int64(W),
                               int(window_size), ctypes.c_uint32(mask), int(tri_method.lower() == "wilson"), planes,
                               int(nthreads))
    assert rc == 0
    return out


def rugosity(dem: np.ndarray, resolution: float, nthreads: int = 0) -> np.ndarray:
    """Rugosity as the reference's Numba engine computes it (window.py:505-595), float32 DEM."""
    dem = np.ascontiguousarray(dem, dtype=np.float32)
    H, W = dem.shape
    out = np.empty((H, W), dtype=np.float32)
    rc = lib().xo_rugosity_f32(ctypes.c_void_p(dem.ctypes.data), ctypes.c_int64(H), ctypes.c_int64(W),
                               ctypes.c_double(resolution), ctypes.c_void_p(out.ctypes.data), int(nthreads))
    assert rc == 0
    return out


def variogram_pairs(coords: np.ndarray, values: np.ndarray, edges: np.ndarray, nthreads: int = 0
                    ) -> tuple[np.ndarray, np.ndarray, float]:
    """All-pairs lag binning (C/OpenMP): returns (count[int64], sumsq[float64], dmax)."""
    x = np.ascontiguousarray(coords[:, 0], dtype=np.float64)
    y = np.ascontiguousarray(coords[:, 1], dtype=np.float64)
    v = np.ascontiguousarray(values, dtype=np.float64)
    e = np.ascontiguousarray(edges, dtype=np.float64)
    count = np.zeros(len(e), dtype=np.int64)
    sumsq = np.zeros(len(e), dtype=np.float64)
    dmax = ctypes.c_double(0.0)
    rc = lib().xo_variogram_pairs(ctypes.c_void_p(x.ctypes.data), ctypes.c_void_p(y.ctypes.data),
                                  ctypes.c_void_p(v.ctypes.data), ctypes.c_int64(len(v)),
                                  ctypes.c_void_p(e.ctypes.data), int(len(e)), ctypes.c_void_p(count.ctypes.data),
                                  ctypes.c_void_p(sumsq.ctypes.data), ctypes.byref(dmax), int(nthreads))
    assert rc == 0
    return count, sumsq, float(dmax.value)
