"""Surface-fit seam: drop-in for ``xdem.terrain.surfit._get_surface_attributes`` (surfit.py:1197-1305)."""

from __future__ import annotations

from typing import Any

import numpy as np
import torch

from . import _arrays, _engine


def _get_surface_attributes(
    dem: Any,
    resolution: float,
    surface_attributes: list[str],
    out_dtype: Any = np.float32,
    surface_fit: str = "Florinsky",
    curv_method: str = "geometric",
    engine: str = "b200",
    **kwargs: Any,
) -> Any:
    """Same array-level contract as the reference seam: returns ``(n_attr, H, W)`` in the order of
    ``surface_attributes``; slope/aspect in radians, hillshade unclipped; NaN where the w x w window (w=5 for Florinsky,
    else 3) holds a NaN or leaves the raster.  ``engine`` is accepted for signature compatibility; the CUDA path always
    runs.  kwargs: hillshade_azimuth, hillshade_altitude, hillshade_z_factor (surfit.py:483-485).

    A host raster (what ``xdem.terrain.terrain`` passes, terrain.py:574-583) is streamed through the GPU in row blocks
    (``xb_terrain_fused_host``: H2D / kernel / D2H overlapped, planes land in page-locked memory) -- the raster and its
    planes never have to fit in device memory; a ``torch.cuda`` tensor is processed in place and a tensor is returned."""
    hs = dict(hillshade_azimuth=kwargs.get("hillshade_azimuth", 315.0),
              hillshade_altitude=kwargs.get("hillshade_altitude", 45.0),
              hillshade_z_factor=kwargs.get("hillshade_z_factor", 1.0))
    if isinstance(dem, torch.Tensor) and dem.is_cuda:
        t, kind = _arrays.to_device(dem)
        out = _engine.terrain_fused(t, resolution, surface_attributes=surface_attributes, surface_fit=surface_fit,
                                    curv_method=curv_method, degrees=False, clip_hillshade=False, **hs)
        return _arrays.from_device(out, kind, out_dtype)
    arr = dem.numpy() if isinstance(dem, torch.Tensor) else _arrays.to_host_nan_array(dem)
    if arr.dtype not in (np.float32, np.float64):
        arr = arr.astype(np.float32)
    out = _engine.terrain_fused_host(arr, resolution, surface_attributes=surface_attributes, surface_fit=surface_fit,
                                     curv_method=curv_method, degrees=False, clip_hillshade=False, **hs)
    if out_dtype is not None and out.dtype != np.dtype(out_dtype):
        out = out.astype(out_dtype)
    return out
