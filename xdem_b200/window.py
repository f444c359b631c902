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
    """Returns ``(n_attr, H, W)`` in the order of ``windowed_indexes`` (TPI, TRI, roughness, rugosity, fractal
    roughness); NaN where the window holds a NaN or leaves the raster (window.py:986, 111-112).  3x3 / 5x5 windows use
    the fused kernel, other odd sizes (<= 31) and fractal roughness the generic odd-window kernel."""
    if window_size < 3 or window_size > 31 or window_size % 2 == 0:
        raise NotImplementedError(f"the B200 engine supports odd window sizes between 3 and 31 (got {window_size})")
    t, kind = _arrays.to_device(dem)
    fused = [a for a in windowed_indexes if a != "fractal_roughness" and window_size in (3, 5)
             and not (a == "rugosity" and window_size != 3)]
    planes = {}
    if fused:
        out = _engine.terrain_fused(t, resolution, windowed_indexes=fused, tri_method=tri_method,
                                    window_size=window_size)
        planes.update({a: out[i] for i, a in enumerate(fused)})
    if "rugosity" in windowed_indexes and "rugosity" not in fused:
        # the reference's SciPy engine always evaluates rugosity on 3x3 (window.py:909-914)
        planes["rugosity"] = _engine.terrain_fused(t, resolution, windowed_indexes=["rugosity"], window_size=3)[0]
    generic = [a for a in windowed_indexes if a not in planes]
    if generic:
        out = _engine.windowed_generic(t, window_size, generic, tri_method=tri_method)
        planes.update({a: out[i] for i, a in enumerate(generic)})
    import torch

    stacked = torch.stack([planes[a] for a in windowed_indexes])
    return _arrays.from_device(stacked, kind, out_dtype)
