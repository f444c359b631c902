"""Thin typed layer over the C ABI: allocates output planes as torch.cuda tensors and launches the CUDA kernels."""

from __future__ import annotations

import ctypes
from typing import Sequence

import numpy as np
import torch

from . import _lib

SURFACE_ORDER = ["slope", "aspect", "hillshade", "curvature", "profile_curvature", "tangential_curvature",
                 "planform_curvature", "flowline_curvature", "max_curvature", "min_curvature"]  # surfit.py:407-418
WINDOW_ORDER = ["topographic_position_index", "terrain_ruggedness_index", "roughness", "rugosity"]  # window.py:752-758
FIT_IDS = {"horn": 0, "zevenbergthorne": 1, "florinsky": 2}  # surfit.py:1240
CURV_IDS = {"geometric": 0, "directional": 1}  # surfit.py:1244
N_PLANES = 14


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return 0
    if t.dtype == torch.float64:
        return 1
    raise TypeError(f"unsupported raster dtype {t.dtype}")


def terrain_fused(
    dem: torch.Tensor,
    resolution: float,
    surface_attributes: Sequence[str] = (),
    windowed_indexes: Sequence[str] = (),
    surface_fit: str = "Florinsky",
    curv_method: str = "geometric",
    tri_method: str = "Riley",
    window_size: int = 3,
    degrees: bool = False,
    clip_hillshade: bool = False,
    hillshade_azimuth: float = 315.0,
    hillshade_altitude: float = 45.0,
    hillshade_z_factor: float = 1.0,
    row_begin: int = 0,
    row_end: int | None = None,
    out: torch.Tensor | None = None,
) -> torch.Tensor:
    """One fused pass (xb_terrain_fused).  Returns a (n_attr, rows, cols) CUDA tensor whose planes follow
    ``list(surface_attributes) + list(windowed_indexes)``.  ``dem`` may carry halo rows: only rows
    [row_begin,row_end) are produced, rows outside the buffer count as NaN."""
    if not dem.is_cuda or dem.dim() != 2:
        raise ValueError("dem must be a 2-D CUDA tensor")
    if dem.stride(1) != 1:
        dem = dem.contiguous()
    L = _lib.lib()
    rows_buf, cols = dem.shape
    ld = dem.stride(0)
    if row_end is None:
        row_end = rows_buf
    names = list(surface_attributes) + list(windowed_indexes)
    n_rows = row_end - row_begin
    out_ld = cols
    if out is None:
        out = torch.empty((len(names), n_rows, cols), dtype=dem.dtype, device=dem.device)
    elif out.shape != (len(names), n_rows, cols) or out.dtype != dem.dtype or (n_rows > 0 and out.stride(2) != 1):
        raise ValueError("bad `out` tensor")  # planes may be row slices of a larger tensor (row stride = out_ld)
    out_ld = out.stride(1) if n_rows > 1 else cols
    planes = (ctypes.c_void_p * N_PLANES)()
    surf_mask = 0
    win_mask = 0
    for i, a in enumerate(names):
        if i < len(surface_attributes):
            slot = SURFACE_ORDER.index(a)
            if surf_mask >> slot & 1:
                raise ValueError(f"duplicate attribute {a}")
            surf_mask |= 1 << slot
        else:
            j = WINDOW_ORDER.index(a)
            slot = 10 + j
            if win_mask >> j & 1:
                raise ValueError(f"duplicate attribute {a}")
            win_mask |= 1 << j
        planes[slot] = out[i].data_ptr()
    stream = torch.cuda.current_stream(dem.device).cuda_stream
    with torch.cuda.device(dem.device):
        rc = L.xb_terrain_fused(
            dem.data_ptr(), _dtype_code(dem), rows_buf, cols, ld, row_begin, row_end, float(resolution),
            FIT_IDS[surface_fit.lower()], CURV_IDS[curv_method.lower()], surf_mask, win_mask, int(window_size),
            0 if tri_method.lower() == "riley" else 1, int(bool(degrees)), int(bool(clip_hillshade)),
            float(hillshade_azimuth), float(hillshade_altitude), float(hillshade_z_factor), planes, int(out_ld),
            ctypes.c_void_p(stream))
    _lib.check(rc)
    return out


def _masks_and_slots(surface_attributes: Sequence[str], windowed_indexes: Sequence[str]) -> tuple[int, int, list[int]]:
    surf_mask = 0
    win_mask = 0
    slots = []
    for a in surface_attributes:
        k = SURFACE_ORDER.index(a)
        if surf_mask >> k & 1:
            raise ValueError(f"duplicate attribute {a}")
        surf_mask |= 1 << k
        slots.append(k)
    for a in windowed_indexes:
        j = WINDOW_ORDER.index(a)
        if win_mask >> j & 1:
            raise ValueError(f"duplicate attribute {a}")
        win_mask |= 1 << j
        slots.append(10 + j)
    return surf_mask, win_mask, slots


def terrain_fused_host(
    dem: np.ndarray,
    resolution: float,
    surface_attributes: Sequence[str] = (),
    windowed_indexes: Sequence[str] = (),
    surface_fit: str = "Florinsky",
    curv_method: str = "geometric",
    tri_method: str = "Riley",
    window_size: int = 3,
    degrees: bool = False,
    clip_hillshade: bool = False,
    hillshade_azimuth: float = 315.0,
    hillshade_altitude: float = 45.0,
    hillshade_z_factor: float = 1.0,
    out: "np.ndarray | Sequence[np.ndarray] | None" = None,
    rows_per_block: int = 0,
    row_begin: int = 0,
    row_end: int | None = None,
) -> np.ndarray:
    """Host-buffer path (xb_terrain_fused_host_rows): a C-contiguous float32/float64 NumPy raster is streamed through
    the GPU in row blocks with halo rows (H2D / kernel / D2H overlapped); returns a (n_attr, rows, W) NumPy array.
    ``dem`` may be pageable (a reference caller's ndarray: staged block-wise through pinned scratch by a few host
    threads) or pinned.  When ``out`` is not given the planes are allocated in page-locked memory (torch's caching
    host allocator recycles the blocks once the caller drops the array), so the device-to-host copies -- 4/5 of the
    traffic of a 4-attribute request -- run at link rate; ``XDEM_B200_PINNED_OUTPUT=0`` allocates plain NumPy memory.
    ``out`` may also be a (n_attr, rows, W) array or a list of (rows, W) planes to fill.
    ``row_begin``/``row_end``: only these rows of ``dem`` are computed (the rest are halo rows of a row shard)."""
    if not torch.cuda.is_available():
        raise RuntimeError("xdem_b200 needs a CUDA device (B200, sm_100a): no CPU fallback exists.")
    if dem.ndim != 2 or dem.dtype not in (np.float32, np.float64):
        raise ValueError("dem must be a 2-D float32/float64 array")
    dem = np.ascontiguousarray(dem)
    L = _lib.lib()
    rows, cols = dem.shape
    if row_end is None:
        row_end = rows
    n_rows = row_end - row_begin
    names = list(surface_attributes) + list(windowed_indexes)
    if out is None:
        out = host_planes(len(names), n_rows, cols, dem.dtype)
    if len(out) != len(names) or any(o.shape != (n_rows, cols) or o.dtype != dem.dtype or not o.flags.c_contiguous
                                     for o in out):
        raise ValueError("bad `out`: one C-contiguous (rows, cols) plane of the raster's dtype per attribute")
    surf_mask, win_mask, slots = _masks_and_slots(surface_attributes, windowed_indexes)
    planes = (ctypes.c_void_p * N_PLANES)()
    for i, slot in enumerate(slots):
        planes[slot] = out[i].ctypes.data
    with torch.cuda.device(torch.cuda.current_device()):
        rc = L.xb_terrain_fused_host_rows(
            ctypes.c_void_p(dem.ctypes.data), 0 if dem.dtype == np.float32 else 1, rows, cols, int(row_begin),
            int(row_end), float(resolution), FIT_IDS[surface_fit.lower()], CURV_IDS[curv_method.lower()], surf_mask,
            win_mask, int(window_size), 0 if tri_method.lower() == "riley" else 1, int(bool(degrees)),
            int(bool(clip_hillshade)), float(hillshade_azimuth), float(hillshade_altitude), float(hillshade_z_factor),
            planes, int(rows_per_block))
    _lib.check(rc)
    return out


def host_planes(n: int, rows: int, cols: int, dtype: np.dtype) -> np.ndarray:
    """(n, rows, cols) output planes for the host-buffer path: a NumPy view of page-locked memory from torch's caching
    host allocator (kept alive by the array's base), or plain NumPy memory when XDEM_B200_PINNED_OUTPUT=0 or the
    pinned allocation fails."""
    import os

    if os.environ.get("XDEM_B200_PINNED_OUTPUT", "1") != "0" and n * rows * cols > 0:
        try:
            t = torch.empty((n, rows, cols), dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
            return t.numpy()
        except RuntimeError:
            pass
    return np.empty((n, rows, cols), dtype=dtype)


GENERIC_SLOTS = {"topographic_position_index": 0, "terrain_ruggedness_index": 1, "roughness": 2, "fractal_roughness": 4}


def windowed_generic(dem: torch.Tensor, window_size: int, windowed_indexes: Sequence[str], tri_method: str = "Riley",
                     row_begin: int = 0, row_end: int | None = None) -> torch.Tensor:
    """Any odd window (3..31): TPI / TRI / roughness / fractal roughness (xb_windowed_generic).  Returns
    (n_attr, rows, cols) in the order of ``windowed_indexes``."""
    if not dem.is_cuda or dem.dim() != 2:
        raise ValueError("dem must be a 2-D CUDA tensor")
    if dem.stride(1) != 1:
        dem = dem.contiguous()
    L = _lib.lib()
    rows_buf, cols = dem.shape
    if row_end is None:
        row_end = rows_buf
    out = torch.empty((len(windowed_indexes), row_end - row_begin, cols), dtype=dem.dtype, device=dem.device)
    planes = (ctypes.c_void_p * 5)()
    mask = 0
    for i, a in enumerate(windowed_indexes):
        slot = GENERIC_SLOTS[a]
        if mask >> slot & 1:
            raise ValueError(f"duplicate attribute {a}")
        mask |= 1 << slot
        planes[slot] = out[i].data_ptr()
    stream = torch.cuda.current_stream(dem.device).cuda_stream
    with torch.cuda.device(dem.device):
        rc = L.xb_windowed_generic(dem.data_ptr(), _dtype_code(dem), rows_buf, cols, dem.stride(0), row_begin, row_end,
                                   int(window_size), mask, 0 if tri_method.lower() == "riley" else 1, planes, cols,
                                   ctypes.c_void_p(stream))
    _lib.check(rc)
    return out
