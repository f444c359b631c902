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
    if out is None:
        out = torch.empty((len(names), n_rows, cols), dtype=dem.dtype, device=dem.device)
    elif out.shape != (len(names), n_rows, cols) or out.dtype != dem.dtype or not out.is_contiguous():
        raise ValueError("bad `out` tensor")
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
            float(hillshade_azimuth), float(hillshade_altitude), float(hillshade_z_factor), planes, cols,
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
    out: np.ndarray | None = None,
    rows_per_block: int = 0,
) -> np.ndarray:
    """Host-buffer path (xb_terrain_fused_host): a C-contiguous float32/float64 NumPy raster is streamed through the GPU
    in row blocks with halo rows (H2D / kernel / D2H overlapped); returns a (n_attr, H, W) NumPy array.  Pinned ``dem`` /
    ``out`` buffers (e.g. views of torch pinned tensors) reach PCIe line rate; pageable ones work but copy slower."""
    if not torch.cuda.is_available():
        raise RuntimeError("xdem_b200 needs a CUDA device (B200, sm_100a): no CPU fallback exists.")
    if dem.ndim != 2 or dem.dtype not in (np.float32, np.float64):
        raise ValueError("dem must be a 2-D float32/float64 array")
    dem = np.ascontiguousarray(dem)
    L = _lib.lib()
    rows, cols = dem.shape
    names = list(surface_attributes) + list(windowed_indexes)
    if out is None:
        out = np.empty((len(names), rows, cols), dtype=dem.dtype)
    elif out.shape != (len(names), rows, cols) or out.dtype != dem.dtype or not out.flags.c_contiguous:
        raise ValueError("bad `out` array")
    surf_mask, win_mask, slots = _masks_and_slots(surface_attributes, windowed_indexes)
    planes = (ctypes.c_void_p * N_PLANES)()
    for i, slot in enumerate(slots):
        planes[slot] = out[i].ctypes.data
    with torch.cuda.device(torch.cuda.current_device()):
        rc = L.xb_terrain_fused_host(
            ctypes.c_void_p(dem.ctypes.data), 0 if dem.dtype == np.float32 else 1, rows, cols, float(resolution),
            FIT_IDS[surface_fit.lower()], CURV_IDS[curv_method.lower()], surf_mask, win_mask, int(window_size),
            0 if tri_method.lower() == "riley" else 1, int(bool(degrees)), int(bool(clip_hillshade)),
            float(hillshade_azimuth), float(hillshade_altitude), float(hillshade_z_factor), planes,
            int(rows_per_block))
    _lib.check(rc)
    return out


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
