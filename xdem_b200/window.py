"""Windowed-index seam: drop-in for ``xdem.terrain.window._get_windowed_indexes`` (window.py:926-1002)."""

from __future__ import annotations

from typing import Any

import numpy as np

from . import _arrays, _engine


def _get_windowed_indexes(
    dem: Any,
    window_size: int,
    windowed_indexes: list[str],
    resolution: float,
    out_dtype: Any = np.float32,
    tri_method: str = "Riley",
    engine: str = "b200",
    force_scipy_backend: Any = None,
) -> Any:
    """Returns ``(n_attr, H, W)`` in the order of ``windowed_indexes`` (TPI, TRI, roughness, rugosity); NaN where the
    window holds a NaN or leaves the raster (window.py:986, 111-112)."""
    if "fractal_roughness" in windowed_indexes:
        raise NotImplementedError("fractal_roughness is not part of the B200 hot path yet (SURVEY.md 8f rank 1)")
    if window_size not in (3, 5):
        raise NotImplementedError(f"the B200 engine supports window_size 3 or 5 (got {window_size})")
    t, kind = _arrays.to_device(dem)
    out = _engine.terrain_fused(t, resolution, windowed_indexes=windowed_indexes, tri_method=tri_method,
                                window_size=window_size)
    return _arrays.from_device(out, kind, out_dtype)
