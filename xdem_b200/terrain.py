"""Public terrain API: drop-in for ``xdem.terrain`` (terrain.py:176-666 and the one-attribute wrappers :694-1763).

Same signatures, validation order, error / warning messages, dtype rules and output ordering as the reference; the
compute is ONE fused CUDA pass (surface-fit attributes and windowed indexes together, degree conversion and hillshade
clipping included) instead of the reference's engine calls plus whole-array NumPy post-processing.
"""

from __future__ import annotations

import warnings
from typing import Any, Sized

import numpy as np
import torch

from . import _arrays, _engine

# terrain.py:41-84
available_attributes = [
    "slope", "aspect", "hillshade", "profile_curvature", "tangential_curvature", "planform_curvature",
    "flowline_curvature", "max_curvature", "min_curvature", "topographic_position_index", "terrain_ruggedness_index",
    "roughness", "rugosity", "fractal_roughness", "texture_shading",
]
list_requiring_surface_fit = list(_engine.SURFACE_ORDER)
list_requiring_windowed_index = ["terrain_ruggedness_index", "topographic_position_index", "roughness", "rugosity"]
list_requiring_windowed_fractal_index = ["fractal_roughness"]
list_requiring_frequency_domain = ["texture_shading"]

_ENGINES = ("b200", "scipy", "numba")


def get_terrain_attribute(
    dem: Any,
    attribute: str | list[str],
    resolution: tuple[float, float] | float | None = None,
    degrees: bool = True,
    hillshade_altitude: float = 45.0,
    hillshade_azimuth: float = 315.0,
    hillshade_z_factor: float = 1.0,
    slope_method: str | None = None,
    surface_fit: str = "Florinsky",
    curv_method: str = "geometric",
    tri_method: str = "Riley",
    window_size: int = 3,
    window_size_fractal: int = 13,
    engine: str = "b200",
    texture_alpha: float = 0.8,
    out_dtype: Any = None,
    mp_config: Any = None,
) -> Any:
    """Derive one or multiple terrain attributes from a DEM on the GPU (see ``xdem.terrain.get_terrain_attribute``,
    terrain.py:176-485, for the attribute definitions).

    ``dem`` may be a NumPy array, a masked array, a Raster-like object (``.data/.transform/.crs/.res``) or a torch
    tensor; CUDA tensors are consumed in place and CUDA tensors are returned.  ``engine`` accepts the reference's
    values for signature compatibility but the CUDA engine always runs (there is no CPU fallback).

    Output order: a list request returns ``outs[i]`` = attribute ``attribute[i]``.  The reference intends the same but
    re-orders its category-grouped results with ``out[k] = grouped[attribute.index(grouped_names[k])]``
    (terrain.py:648-656), i.e. it applies the permutation instead of its inverse: the two agree whenever that
    permutation is an involution (any request inside one category, any two-category swap) and differ for 3-cycles across
    the surface / windowed / frequency categories such as ``["roughness", "texture_shading", "slope"]``, where the
    reference returns (texture_shading, slope, roughness).  This implementation deliberately keeps the requested order
    (pinned by ``tests/test_terrain_gpu.py::test_three_cycle_request_keeps_requested_order``).

    :examples:
        >>> dem = np.repeat(np.arange(3), 3)[::-1].reshape(3, 3)
        >>> slope, aspect = get_terrain_attribute(dem, ["slope", "aspect"], resolution=1,
        ...                                       surface_fit="ZevenbergThorne")  # doctest: +SKIP
        >>> slope[1, 1], aspect[1, 1]  # doctest: +SKIP
        (np.float32(45.0), np.float32(180.0))
    """
    # 0/ Deprecating slope method (terrain.py:284-291)
    if slope_method is not None:
        warnings.warn("'slope_method' is deprecated, use 'surface_fit' instead.", DeprecationWarning, stacklevel=2)
        surface_fit = slope_method
        slope_method = None

    # 1/ Input checks, in the reference's order (terrain.py:296-409)
    if surface_fit == "Horn":
        curvature_list = ["curvature", "profile_curvature", "tangential_curvature", "planform_curvature",
                          "flowline_curvature", "max_curvature", "min_curvature"]
        found = attribute in curvature_list if isinstance(attribute, str) else any(
            item in curvature_list for item in attribute)
        if found:
            raise ValueError(
                "'Horn' surface fit method cannot be used for to calculate curvatures. "
                "Use 'ZevenbergThorne' or 'Florinsky' instead."
            )

    is_raster = _arrays.is_raster_like(dem)
    if is_raster and resolution is None:
        resolution = dem.res

    if isinstance(attribute, str):
        attribute = [attribute]

    if out_dtype is None:
        in_dtype = _arrays.array_dtype(dem)
        if np.issubdtype(in_dtype, np.integer) or in_dtype == np.bool_:
            out_dtype = np.float32
        else:
            out_dtype = np.dtype(in_dtype)

    attributes_requiring_surface_fit = [attr for attr in attribute if attr in list_requiring_surface_fit]

    if "fractal_roughness" in attribute:
        if window_size_fractal < 5:
            warnings.warn(category=UserWarning, stacklevel=2,
                          message="Fractal roughness can only be computed on window sizes larger or equal to 5.")
        elif window_size_fractal < 13:
            warnings.warn(category=UserWarning, stacklevel=2,
                          message="Fractal roughness results with window size of less than 13 can be inaccurate.")

    attributes_requiring_resolution = attributes_requiring_surface_fit + (
        ["rugosity"] if "rugosity" in attribute else [])
    if len(attributes_requiring_resolution) > 0:
        if resolution is None:
            raise ValueError(
                f"'resolution' must be provided as an argument for attributes: {attributes_requiring_resolution}"
            )
        if not isinstance(resolution, Sized):
            resolution = (float(resolution), float(resolution))
        if resolution[0] != resolution[1]:
            raise ValueError(
                f"Surface fit and rugosity require the same X and Y resolution ({resolution} was given). "
                f"This was required by: {attributes_requiring_resolution}."
            )
    if resolution is None:
        resolution = 1
    elif isinstance(resolution, Sized):
        resolution = resolution[0]

    choices = (list_requiring_surface_fit + list_requiring_windowed_index + list_requiring_windowed_fractal_index
               + list_requiring_frequency_domain)
    for attr in attribute:
        if attr not in choices:
            raise ValueError(f"Attribute '{attr}' is not supported. Choices: {choices}")

    list_surface_fit = ["Horn", "ZevenbergThorne", "Florinsky"]
    if surface_fit.lower() not in [sm.lower() for sm in list_surface_fit]:
        raise ValueError(f"Surface fit '{surface_fit}' is not supported. Must be one of: {list_surface_fit}")
    list_curv_methods = ["geometric", "directional"]
    if curv_method.lower() not in [cm.lower() for cm in list_curv_methods]:
        raise ValueError(f"Curvature method '{curv_method}' is not supported. Must be one of: {list_curv_methods}")
    list_tri_methods = ["Riley", "Wilson"]
    if tri_method.lower() not in [tm.lower() for tm in list_tri_methods]:
        raise ValueError(f"TRI method '{tri_method}' is not supported. Must be one of: {list_tri_methods}")
    if (hillshade_azimuth < 0.0) or (hillshade_azimuth > 360.0):
        raise ValueError(f"Azimuth must be a value between 0 and 360 degrees (given value: {hillshade_azimuth})")
    if (hillshade_altitude < 0.0) or (hillshade_altitude > 90):
        raise ValueError("Altitude must be a value between 0 and 90 degrees (given value: {altitude})")
    if (hillshade_z_factor < 0.0) or not np.isfinite(hillshade_z_factor):
        raise ValueError(f"z_factor must be a non-negative finite value (given value: {hillshade_z_factor})")
    if engine not in _ENGINES:
        raise ValueError(f"Engine '{engine}' is not supported. Must be one of: {list(_ENGINES)}")

    if is_raster and not dem.crs.is_projected and len(attributes_requiring_surface_fit) > 0:
        warnings.warn(
            category=UserWarning,
            message=f"DEM is not in a projected CRS, the following surface fit attributes might be "
            f"wrong: {list_requiring_surface_fit}."
            f"Use DEM.reproject(crs=DEM.get_metric_crs()) to reproject in a projected CRS.",
        )

    if mp_config is not None:
        raise NotImplementedError(
            "mp_config (geoutils out-of-memory tiling, terrain.py:412-466) is replaced by row-sharding across GPUs "
            "(xdem_b200.distributed.sharded_terrain_attribute) and host streaming (xdem_b200.terrain_host)."
        )

    return _get_terrain_attribute(dem, attribute, resolution, degrees, hillshade_altitude, hillshade_azimuth,
                                  hillshade_z_factor, surface_fit, curv_method, tri_method, window_size,
                                  window_size_fractal, engine, texture_alpha, out_dtype)


def _get_terrain_attribute(
    dem: Any,
    attribute: list[str],
    resolution: float,
    degrees: bool = True,
    hillshade_altitude: float = 45.0,
    hillshade_azimuth: float = 315.0,
    hillshade_z_factor: float = 1.0,
    surface_fit: str = "Florinsky",
    curv_method: str = "geometric",
    tri_method: str = "Riley",
    window_size: int = 3,
    window_size_fractal: int = 13,
    engine: str = "b200",
    texture_alpha: float = 0.8,
    out_dtype: Any = None,
) -> Any:
    """terrain.py:528-666 with the surface-fit and windowed groups fused into one kernel launch.  Window sizes other
    than 3/5 and fractal roughness (terrain.py:620-633) go through the generic odd-window kernel."""
    surf = list(dict.fromkeys(a for a in attribute if a in list_requiring_surface_fit))
    win = list(dict.fromkeys(a for a in attribute if a in list_requiring_windowed_index))
    frac = [a for a in attribute if a in list_requiring_windowed_fractal_index]
    other = [a for a in attribute if a in list_requiring_frequency_domain]
    for ws in ([window_size] if win else []) + ([window_size_fractal] if frac else []):
        if ws < 3 or ws > 31 or ws % 2 == 0:
            raise NotImplementedError(f"the B200 engine supports odd window sizes between 3 and 31 (got {ws})")

    # which launches are needed: the fused kernel takes 3x3 / 5x5 windows; rugosity is always 3x3 in the reference's
    # default engine (window.py:909-914 ignores window_size)
    fused_win = [a for a in win if (window_size in (3, 5) and not (a == "rugosity" and window_size != 3))]
    rug_separate = "rugosity" in win and "rugosity" not in fused_win
    generic_win = [a for a in win if a not in fused_win and a != "rugosity"]
    needs_device = bool(generic_win or frac or rug_separate or other)

    is_raster = _arrays.is_raster_like(dem)
    kwargs = dict(surface_fit=surface_fit, curv_method=curv_method, tri_method=tri_method, degrees=degrees,
                  clip_hillshade=True, hillshade_azimuth=hillshade_azimuth, hillshade_altitude=hillshade_altitude,
                  hillshade_z_factor=hillshade_z_factor)
    on_device = isinstance(dem, torch.Tensor) and dem.is_cuda
    as_tensor = isinstance(dem, torch.Tensor)
    planes: dict[str, Any] = {}
    if on_device or needs_device:
        t, _ = _arrays.to_device(dem)
        if surf or fused_win:
            out = _engine.terrain_fused(t, float(resolution), surface_attributes=surf, windowed_indexes=fused_win,
                                        window_size=window_size if fused_win else 3, **kwargs)
            planes.update({a: out[i] for i, a in enumerate(surf + fused_win)})
        if rug_separate:
            planes["rugosity"] = _engine.terrain_fused(t, float(resolution), windowed_indexes=["rugosity"],
                                                       window_size=3)[0]
        if generic_win:
            out = _engine.windowed_generic(t, window_size, generic_win, tri_method=tri_method)
            planes.update({a: out[i] for i, a in enumerate(generic_win)})
        if frac:
            planes["fractal_roughness"] = _engine.windowed_generic(t, window_size_fractal, ["fractal_roughness"])[0]
        if other:
            # terrain.py:641-643: global frequency-domain attribute (cuFFT + the xb_texture_* kernels)
            from xdem_b200 import freq

            planes["texture_shading"] = freq.texture_shading_device(t, texture_alpha)
        kind = "torch" if on_device else "numpy"
        outs = [_arrays.from_device(planes[a], kind, out_dtype) for a in attribute]
        if as_tensor and not on_device:
            outs = [torch.from_numpy(o) for o in outs]
    else:
        # host raster (ndarray / masked array / Raster / CPU tensor): streamed through the GPU in row blocks
        arr = dem.numpy() if as_tensor else _arrays.to_host_nan_array(dem)
        if arr.dtype not in (np.float32, np.float64):
            arr = arr.astype(np.float32)
        planes_h = _engine.terrain_fused_host(arr, float(resolution), surface_attributes=surf,
                                              windowed_indexes=fused_win, window_size=window_size, **kwargs)
        index = {a: i for i, a in enumerate(surf + fused_win)}
        outs = []
        for a in attribute:
            o = planes_h[index[a]]
            if out_dtype is not None and o.dtype != np.dtype(out_dtype):
                o = o.astype(out_dtype)
            outs.append(torch.from_numpy(o) if as_tensor else o)

    if is_raster:
        try:
            import geoutils as gu  # optional

            outs = [gu.Raster.from_array(o, transform=dem.transform, crs=dem.crs, nodata=-99999) for o in outs]
        except ImportError:  # pragma: no cover
            pass
    return outs if len(outs) > 1 else outs[0]


def _deprecated_method(method: str | None, surface_fit: str) -> str:
    if method is not None:
        warnings.warn("'method' is deprecated, use 'surface_fit' instead.", DeprecationWarning, stacklevel=3)
        return method
    return surface_fit


# One-attribute wrappers (terrain.py:694-1763): same positional order and defaults as the reference.


def slope(dem: Any, method: str | None = None, surface_fit: str = "Florinsky", degrees: bool = True,
          resolution: float | tuple[float, float] | None = None, mp_config: Any = None, engine: str = "b200") -> Any:
    """Slope (terrain.py:694-747)."""
    surface_fit = _deprecated_method(method, surface_fit)
    return get_terrain_attribute(dem, attribute="slope", surface_fit=surface_fit, resolution=resolution,
                                 degrees=degrees, mp_config=mp_config, engine=engine)


def aspect(dem: Any, method: str | None = None, surface_fit: str = "Florinsky", degrees: bool = True,
           mp_config: Any = None, engine: str = "b200") -> Any:
    """Aspect; always uses resolution=1.0 like the reference (terrain.py:773-835)."""
    surface_fit = _deprecated_method(method, surface_fit)
    return get_terrain_attribute(dem, attribute="aspect", surface_fit=surface_fit, resolution=1.0, degrees=degrees,
                                 mp_config=mp_config, engine=engine)


def hillshade(dem: Any, method: str | None = None, surface_fit: str = "Florinsky", azimuth: float = 315.0,
              altitude: float = 45.0, z_factor: float = 1.0, resolution: float | tuple[float, float] | None = None,
              mp_config: Any = None, engine: str = "b200") -> Any:
    """Hillshade (terrain.py:867-920)."""
    surface_fit = _deprecated_method(method, surface_fit)
    return get_terrain_attribute(dem, attribute="hillshade", resolution=resolution, surface_fit=surface_fit,
                                 hillshade_azimuth=azimuth, hillshade_altitude=altitude, hillshade_z_factor=z_factor,
                                 mp_config=mp_config, engine=engine)


def curvature(dem: Any, resolution: float | tuple[float, float] | None = None, surface_fit: str = "Florinsky",
              mp_config: Any = None, engine: str = "b200") -> Any:
    """Deprecated total curvature (terrain.py:944-990)."""
    warnings.warn("The curvature attribute is deprecated, refer to docs for specific curvature functions.",
                  DeprecationWarning, stacklevel=2)
    return get_terrain_attribute(dem=dem, attribute="curvature", surface_fit=surface_fit, resolution=resolution,
                                 mp_config=mp_config, engine=engine)


def _curv_wrapper(name: str, dem: Any, resolution: Any, surface_fit: str, curv_method: str, mp_config: Any,
                  engine: str) -> Any:
    return get_terrain_attribute(dem=dem, attribute=name, surface_fit=surface_fit, curv_method=curv_method,
                                 resolution=resolution, mp_config=mp_config, engine=engine)


def profile_curvature(dem: Any, resolution: float | tuple[float, float] | None = None, surface_fit: str = "Florinsky",
                      curv_method: str = "geometric", mp_config: Any = None, engine: str = "b200") -> Any:
    """terrain.py:1016-1066."""
    return _curv_wrapper("profile_curvature", dem, resolution, surface_fit, curv_method, mp_config, engine)


def tangential_curvature(dem: Any, resolution: float | tuple[float, float] | None = None,
                         surface_fit: str = "Florinsky", curv_method: str = "geometric", mp_config: Any = None,
                         engine: str = "b200") -> Any:
    """terrain.py:1092-1143."""
    return _curv_wrapper("tangential_curvature", dem, resolution, surface_fit, curv_method, mp_config, engine)


def planform_curvature(dem: Any, resolution: float | tuple[float, float] | None = None, surface_fit: str = "Florinsky",
                       curv_method: str = "geometric", mp_config: Any = None, engine: str = "b200") -> Any:
    """terrain.py:1169-1218."""
    return _curv_wrapper("planform_curvature", dem, resolution, surface_fit, curv_method, mp_config, engine)


def flowline_curvature(dem: Any, resolution: float | tuple[float, float] | None = None, surface_fit: str = "Florinsky",
                       curv_method: str = "geometric", mp_config: Any = None, engine: str = "b200") -> Any:
    """terrain.py:1244-1294."""
    return _curv_wrapper("flowline_curvature", dem, resolution, surface_fit, curv_method, mp_config, engine)


def max_curvature(dem: Any, resolution: float | tuple[float, float] | None = None, surface_fit: str = "Florinsky",
                  curv_method: str = "geometric", mp_config: Any = None, engine: str = "b200") -> Any:
    """terrain.py:1320-1370."""
    return _curv_wrapper("max_curvature", dem, resolution, surface_fit, curv_method, mp_config, engine)


def min_curvature(dem: Any, resolution: float | tuple[float, float] | None = None, surface_fit: str = "Florinsky",
                  curv_method: str = "geometric", mp_config: Any = None, engine: str = "b200") -> Any:
    """terrain.py:1396-1446."""
    return _curv_wrapper("min_curvature", dem, resolution, surface_fit, curv_method, mp_config, engine)


def topographic_position_index(dem: Any, window_size: int = 3, mp_config: Any = None, engine: str = "b200") -> Any:
    """terrain.py:1468-1507."""
    return get_terrain_attribute(dem=dem, attribute="topographic_position_index", window_size=window_size,
                                 mp_config=mp_config, engine=engine)


def terrain_ruggedness_index(dem: Any, method: str = "Riley", window_size: int = 3, mp_config: Any = None,
                             engine: str = "b200") -> Any:
    """terrain.py:1531-1578."""
    return get_terrain_attribute(dem=dem, attribute="terrain_ruggedness_index", tri_method=method,
                                 window_size=window_size, mp_config=mp_config, engine=engine)


def roughness(dem: Any, window_size: int = 3, mp_config: Any = None, engine: str = "b200") -> Any:
    """terrain.py:1600-1639."""
    return get_terrain_attribute(dem=dem, attribute="roughness", window_size=window_size, mp_config=mp_config,
                                 engine=engine)


def rugosity(dem: Any, resolution: float | tuple[float, float] | None = None, mp_config: Any = None,
             engine: str = "b200") -> Any:
    """terrain.py:1661-1700."""
    return get_terrain_attribute(dem=dem, attribute="rugosity", resolution=resolution, mp_config=mp_config,
                                 engine=engine)


def fractal_roughness(dem: Any, window_size_fractal: int = 13, mp_config: Any = None, engine: str = "b200") -> Any:
    """terrain.py:1722-1763 (box-counting fractal dimension on a `window_size_fractal` window, default 13)."""
    return get_terrain_attribute(dem=dem, attribute="fractal_roughness", window_size_fractal=window_size_fractal,
                                 mp_config=mp_config, engine=engine)


def texture_shading(dem: Any, alpha: float = 0.8, mp_config: Any = None) -> Any:
    """terrain.py:1783-1838: texture shaded relief (fractional Laplacian, exponent `alpha` in [0, 2])."""
    return get_terrain_attribute(dem=dem, attribute="texture_shading", texture_alpha=alpha, mp_config=mp_config)


__all__ = [
    "available_attributes", "get_terrain_attribute", "slope", "aspect", "hillshade", "curvature",
    "profile_curvature", "tangential_curvature", "planform_curvature", "flowline_curvature", "max_curvature",
    "min_curvature", "topographic_position_index", "terrain_ruggedness_index", "roughness", "rugosity",
    "fractal_roughness", "texture_shading",
]
_ = torch  # torch is the device-memory provider
