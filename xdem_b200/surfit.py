"""Surface-fit seam: drop-in for ``xdem.terrain.surfit._get_surface_attributes`` (surfit.py:1197-1305)."""

from __future__ import annotations

from typing import Any

import numpy as np

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
    runs.  kwargs: hillshade_azimuth, hillshade_altitude, hillshade_z_factor (surfit.py:483-485)."""
    t, kind = _arrays.to_device(dem)
    out = _engine.terrain_fused(
        t, resolution, surface_attributes=surface_attributes, surface_fit=surface_fit, curv_method=curv_method,
        degrees=False, clip_hillshade=False,
        hillshade_azimuth=kwargs.get("hillshade_azimuth", 315.0),
        hillshade_altitude=kwargs.get("hillshade_altitude", 45.0),
        hillshade_z_factor=kwargs.get("hillshade_z_factor", 1.0))
    return _arrays.from_device(out, kind, out_dtype)
