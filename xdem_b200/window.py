"""Windowed-index seam: drop-in for ``xdem.terrain.window._get_windowed_indexes`` (window.py:926-1002)."""

from __future__ import annotations

from typing import Any

import numpy as np
import torch

from . import _arrays, _engine


def _split(window_size: int, windowed_indexes: list[str]) -> tuple[list[str], bool, list[str]]:
    """(indexes the fused 3x3 / 5x5 kernels take at this window size, rugosity needs its own 3x3 pass, the rest)."""
    fused = [a for a in windowed_indexes if a != "fractal_roughness" and window_size in (3, 5)
             and not (a == "rugosity" and window_size != 3)]
    # the reference's SciPy engine always evaluates rugosity on 3x3 (window.py:909-914)
    rug_separate = "rugosity" in windowed_indexes and "rugosity" not in fused
    generic = [a for a in windowed_indexes if a not in fused and a != "rugosity"]
    return fused, rug_separate, generic


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
    the fused kernels, other odd sizes (<= 31) and fractal roughness the generic odd-window kernel.

    A host raster (terrain.py:606-614) is streamed through the GPU in row blocks for the fused indexes, every plane
    written straight into the (page-locked) result array; only the generic odd-window kernel needs the raster resident
    on the device.  A ``torch.cuda`` tensor is processed in place and a tensor is returned."""
    if window_size < 3 or window_size > 31 or window_size % 2 == 0:
        raise NotImplementedError(f"the B200 engine supports odd window sizes between 3 and 31 (got {window_size})")
    fused, rug_separate, generic = _split(window_size, list(windowed_indexes))
    index = {a: i for i, a in enumerate(windowed_indexes)}

    if isinstance(dem, torch.Tensor) and dem.is_cuda:
        t, kind = _arrays.to_device(dem)
        out = torch.empty((len(windowed_indexes),) + tuple(t.shape), dtype=t.dtype, device=t.device)
        if fused:
            sub = _engine.terrain_fused(t, resolution, windowed_indexes=fused, tri_method=tri_method,
                                        window_size=window_size)
            for i, a in enumerate(fused):
                out[index[a]] = sub[i]
        if rug_separate:
            out[index["rugosity"]] = _engine.terrain_fused(t, resolution, windowed_indexes=["rugosity"],
                                                           window_size=3)[0]
        if generic:
            sub = _engine.windowed_generic(t, window_size, generic, tri_method=tri_method)
            for i, a in enumerate(generic):
                out[index[a]] = sub[i]
        return _arrays.from_device(out, kind, out_dtype)

    arr = dem.numpy() if isinstance(dem, torch.Tensor) else _arrays.to_host_nan_array(dem)
    if arr.dtype not in (np.float32, np.float64):
        arr = arr.astype(np.float32)
    arr = np.ascontiguousarray(arr)
    out = _engine.host_planes(len(windowed_indexes), arr.shape[0], arr.shape[1], arr.dtype)
    if fused:
        _engine.terrain_fused_host(arr, resolution, windowed_indexes=fused, tri_method=tri_method,
                                   window_size=window_size, out=[out[index[a]] for a in fused])
    if rug_separate:
        _engine.terrain_fused_host(arr, resolution, windowed_indexes=["rugosity"], window_size=3,
                                   out=[out[index["rugosity"]]])
    if generic:
        t = torch.from_numpy(arr).to(_arrays.require_cuda())
        sub = _engine.windowed_generic(t, window_size, generic, tri_method=tri_method)
        for i, a in enumerate(generic):
            torch.from_numpy(out[index[a]]).copy_(sub[i])
    if out_dtype is not None and out.dtype != np.dtype(out_dtype):
        out = out.astype(out_dtype)
    return out
