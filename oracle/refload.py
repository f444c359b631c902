"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference (GlacioHack/xdem) in the build container.

The reference is a pure-Python package whose third-party geo stack (geoutils, geopandas, rasterio, pyproj, affine,
scikit-gstat) is not installed here, so ``import xdem`` fails (xdem/__init__.py:19-23).  This module registers small
stub modules for those packages and then loads the reference's own source files *where they lie* under
``/root/reference`` with ``importlib`` (nothing is copied).  The reference code that then runs unmodified is:

* ``xdem/terrain/{surfit,window,terrain}.py`` -- both engines ("scipy", "numba") of every terrain attribute,
* ``xdem/spatialstats.py``                     -- ``nd_binning``, ``convolution``, the variogram glue,
* ``xdem/coreg/{base,affine}.py``              -- ``_nuth_kaab_aux_vars``, ``_nuth_kaab_iteration_step``,
                                                  ``_nuth_kaab_bin_fit``, ``_bin_or_and_fit_nd``, ``_iterate_method``.

Third-party pieces that are *restated* (parity unpinned for them):
``geoutils.raster.get_array_and_mask`` (ndarray / masked array -> NaN array + mask), ``geoutils.stats.nmad``,
``geoutils._interp_points`` (bilinear, NaN-propagating) and ``subsample_array``.

This file is only used (a) by ``oracle/make_golden.py`` to generate the committed fixtures under ``tests/golden`` and
(b) by tests marked ``needs_reference`` that are skipped when ``/root/reference`` is absent (the GPU box).
Nothing under ``xdem_b200/`` may import it.
"""

from __future__ import annotations

import importlib.util
import os
import sys
import types
from typing import Any

import numpy as np

REFERENCE_ROOT = os.environ.get("XDEM_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "xdem", "terrain", "surfit.py"))


class _AutoMock(types.ModuleType):
    """Module whose unknown attributes resolve to further auto-mocks / dummy classes (for type annotations only)."""

    def __getattr__(self, name: str) -> Any:
        if name.startswith("__"):
            raise AttributeError(name)
        full = f"{self.__name__}.{name}"
        if name[:1].isupper():
            obj: Any = type(name, (), {})
        else:
            obj = _AutoMock(full)
            sys.modules.setdefault(full, obj)
        setattr(self, name, obj)
        return obj

    def __call__(self, *a: Any, **k: Any) -> Any:  # used as decorator / function placeholder
        raise RuntimeError(f"stubbed third-party function {self.__name__} was called")


def _mod(name: str) -> _AutoMock:
    if name in sys.modules and isinstance(sys.modules[name], _AutoMock):
        return sys.modules[name]  # type: ignore
    m = _AutoMock(name)
    m.__path__ = []  # type: ignore  # behave like a package
    sys.modules[name] = m
    return m


# ---------------------------------------------------------------------------------------------------------------
# Restated third-party helpers (geoutils==0.2.5 is not installed; behaviour restated from its documented semantics)
# ---------------------------------------------------------------------------------------------------------------


def get_array_and_mask(array: Any, check_shape: bool = True, copy: bool = True) -> tuple[np.ndarray, np.ndarray]:
    """ndarray / masked array -> (float array with NaN at invalid cells, boolean invalid mask)."""
    if isinstance(array, np.ma.MaskedArray):
        data = np.array(array.data, copy=True)
        if not np.issubdtype(data.dtype, np.floating):
            data = data.astype(np.float32)
        mask = np.ma.getmaskarray(array) | ~np.isfinite(data)
        data[mask] = np.nan
        return data, mask
    arr = np.asarray(array)
    if not np.issubdtype(arr.dtype, np.floating):
        # integer arrays have no invalid values; terrain.py:560 casts them itself
        return (arr.copy() if copy else arr), np.zeros(arr.shape, dtype=bool)
    arr = arr.copy() if copy else arr
    return arr, ~np.isfinite(arr)


def nmad(data: Any, nfact: float = 1.4826) -> Any:
    data = np.asarray(data)
    return nfact * np.nanmedian(np.abs(data - np.nanmedian(data)))


def subsample_array(array: Any, subsample: float | int, return_indices: bool = False, random_state: Any = None) -> Any:
    """Random subsample among valid (finite, unmasked) values; restated, RNG stream NOT identical to geoutils."""
    rng = np.random.default_rng(random_state)
    if isinstance(array, np.ma.MaskedArray):
        valid = ~np.ma.getmaskarray(array).ravel()
        if np.issubdtype(array.dtype, np.floating):
            valid &= np.isfinite(array.data.ravel())
    else:
        arr = np.asarray(array)
        valid = np.isfinite(arr.ravel()) if np.issubdtype(arr.dtype, np.floating) else np.ones(arr.size, bool)
    idx_valid = np.flatnonzero(valid)
    n = int(subsample) if subsample > 1 else int(subsample * idx_valid.size)
    n = min(n, idx_valid.size)
    chosen = rng.choice(idx_valid, size=n, replace=False)
    unraveled = np.unravel_index(chosen, np.shape(array))
    if return_indices:
        return unraveled
    return np.asarray(array)[unraveled]


class _SimpleAffine:
    """Minimal north-up affine (a, 0, c, 0, e, f) with e<0, enough for _res / _coords / _interp_points."""

    def __init__(self, a: float, b: float, c: float, d: float, e: float, f: float) -> None:
        self.a, self.b, self.c, self.d, self.e, self.f = a, b, c, d, e, f

    def __iter__(self):  # noqa
        return iter((self.a, self.b, self.c, self.d, self.e, self.f, 0.0, 0.0, 1.0))


def _res(transform: Any) -> tuple[float, float]:
    return (abs(transform.a), abs(transform.e))


def _coords(transform: Any, shape: tuple[int, int], area_or_point: Any = None, grid: bool = True,
            shift_area_or_point: Any = None, force_offset: Any = None) -> tuple[np.ndarray, np.ndarray]:
    """Pixel coordinates; offset convention is irrelevant for the NK path because the same convention is used to
    build the interpolator and the points (affine.py:171-184)."""
    h, w = shape
    xs = transform.c + (np.arange(w) + 0.5) * transform.a
    ys = transform.f + (np.arange(h) + 0.5) * transform.e
    if grid:
        xx, yy = np.meshgrid(xs, ys)
        return xx, yy
    return xs, ys


def _interp_points(array: np.ndarray, transform: Any, points: Any, area_or_point: Any = None,
                   method: str = "linear", dist_nodata_spread: Any = "0.5", return_interpolator: bool = False,
                   **kwargs: Any) -> Any:
    """Bilinear, NaN-propagating point interpolation of a north-up raster (restatement of geoutils._interp_points for
    method="linear": any NaN/out-of-grid cell among the 4 contributing cells gives NaN).  Input points are (y, x)
    when called through the returned interpolator (affine.py:184) and (x, y) when passed as ``points``."""
    from scipy.ndimage import map_coordinates

    if method != "linear":
        raise NotImplementedError("restated interpolator supports method='linear' only")
    arr64 = np.asarray(array, dtype=np.float64)

    def interp_rowcol(rows: np.ndarray, cols: np.ndarray) -> np.ndarray:
        return map_coordinates(arr64, [rows, cols], order=1, mode="constant", cval=np.nan, prefilter=False)

    def yx_to_rowcol(y: np.ndarray, x: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
        cols = (np.asarray(x, dtype=np.float64) - transform.c) / transform.a - 0.5
        rows = (np.asarray(y, dtype=np.float64) - transform.f) / transform.e - 0.5
        return rows, cols

    if return_interpolator:

        def interpolator(yx: tuple[np.ndarray, np.ndarray]) -> np.ndarray:
            rows, cols = yx_to_rowcol(yx[0], yx[1])
            return interp_rowcol(rows, cols)

        return interpolator
    rows, cols = yx_to_rowcol(points[1], points[0])
    return interp_rowcol(rows, cols)


# ---------------------------------------------------------------------------------------------------------------
# Stub installation + loading of the reference's own files
# ---------------------------------------------------------------------------------------------------------------

_LOADED: dict[str, Any] = {}


def _install_stubs() -> None:
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/xdem_b200_numba_cache")  # never write next to the reference

    # Third-party packages that are absent: auto-mocks
    for name in ["affine", "geopandas", "rasterio", "rasterio.warp", "rasterio.transform", "rasterio.crs", "pyproj",
                 "geoutils", "geoutils.raster", "geoutils.raster.array", "geoutils.raster.distributed_computing",
                 "geoutils.raster.georeferencing", "geoutils.raster.geotransformations",
                 "geoutils.raster._geotransformations", "geoutils.raster.raster",
                 "geoutils.stats", "geoutils.stats.sampling", "geoutils.vector", "geoutils.vector.vector",
                 "geoutils.interface", "geoutils.interface.gridding", "geoutils.interface.interpolate",
                 "geoutils.pointcloud", "geoutils.pointcloud.pointcloud", "geoutils._typing", "geoutils.profiler",
                 "geoutils.projtools", "geoutils.raster.sampling"]:
        _mod(name)

    gu = sys.modules["geoutils"]

    class Raster:  # dummy: isinstance(dem, gu.Raster) must be False for ndarrays
        pass

    class Vector:
        pass

    class PointCloud:
        pass

    class GeoDataFrame:
        pass

    def profile(name: str, memprof: bool = False):  # geoutils.profiler.profile -> identity decorator
        def deco(f):
            return f

        return deco

    gu.Raster = Raster
    gu.Vector = Vector
    gu.PointCloud = PointCloud
    sys.modules["geoutils.profiler"].profile = profile
    gu.profiler = sys.modules["geoutils.profiler"]
    r = sys.modules["geoutils.raster"]
    r.Raster, r.RasterType = Raster, Raster
    r.get_array_and_mask = get_array_and_mask
    r.raster = sys.modules["geoutils.raster.raster"]
    r.array = sys.modules["geoutils.raster.array"]
    gu.raster = r
    sys.modules["geoutils.raster.array"].get_array_and_mask = get_array_and_mask
    dc = sys.modules["geoutils.raster.distributed_computing"]
    dc.MultiprocConfig = type("MultiprocConfig", (), {})
    dc.map_overlap_multiproc_save = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("multiproc stub"))
    geo = sys.modules["geoutils.raster.georeferencing"]
    geo._res, geo._coords = _res, _coords
    geo._cast_pixel_interpretation = lambda a, b: a
    sys.modules["geoutils.raster.geotransformations"]._translate = lambda t, xoff, yoff: _SimpleAffine(
        t.a, t.b, t.c + xoff, t.d, t.e, t.f + yoff)
    sys.modules["geoutils.raster._geotransformations"]._resampling_method_from_str = lambda s: s
    st = sys.modules["geoutils.stats"]
    st.nmad = nmad
    st.sampling = sys.modules["geoutils.stats.sampling"]
    gu.stats = st
    sys.modules["geoutils.stats.sampling"].subsample_array = subsample_array
    vv = sys.modules["geoutils.vector.vector"]
    vv.Vector, vv.VectorType = Vector, Vector
    ii = sys.modules["geoutils.interface.interpolate"]
    ii._interp_points = _interp_points
    sys.modules["geoutils.interface.gridding"]._grid_pointcloud = lambda *a, **k: None
    pc = sys.modules["geoutils.pointcloud.pointcloud"]
    pc.PointCloud, pc.PointCloudType = PointCloud, PointCloud
    sys.modules["geoutils._typing"].Number = float
    sys.modules["geopandas"].GeoDataFrame = GeoDataFrame
    sys.modules["affine"].Affine = _SimpleAffine
    sys.modules["rasterio"].transform = sys.modules["rasterio.transform"]
    sys.modules["rasterio"].warp = sys.modules["rasterio.warp"]
    sys.modules["rasterio"].crs = sys.modules["rasterio.crs"]
    sys.modules["rasterio.transform"].Affine = _SimpleAffine

    # The fake top-level package (its real __init__ imports the whole geo stack)
    if "xdem" not in sys.modules or not getattr(sys.modules["xdem"], "_xb_stub", False):
        x = types.ModuleType("xdem")
        x.__path__ = []  # type: ignore
        x.__version__ = "0.2.3"
        x._xb_stub = True  # type: ignore
        sys.modules["xdem"] = x


def _load_file(modname: str, relpath: str) -> Any:
    if modname in _LOADED:
        return _LOADED[modname]
    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(modname, path)
    assert spec is not None and spec.loader is not None
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    sys.dont_write_bytecode, old = True, sys.dont_write_bytecode
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.dont_write_bytecode = old
    _LOADED[modname] = mod
    parent, _, child = modname.rpartition(".")
    if parent in sys.modules:
        setattr(sys.modules[parent], child, mod)
    return mod


def load_reference() -> types.SimpleNamespace:
    """Load the reference's hot-path modules (unmodified source files) and return them in a namespace."""
    if not reference_available():
        raise FileNotFoundError(f"reference not found under {REFERENCE_ROOT}")
    if "ns" in _LOADED:
        return _LOADED["ns"]
    _install_stubs()
    _load_file("xdem._typing", "xdem/_typing.py")
    _load_file("xdem._misc", "xdem/_misc.py")
    _load_file("xdem.fit", "xdem/fit.py")
    spatialstats = _load_file("xdem.spatialstats", "xdem/spatialstats.py")
    pkg_t = types.ModuleType("xdem.terrain")
    pkg_t.__path__ = []  # type: ignore
    sys.modules["xdem.terrain"] = pkg_t
    sys.modules["xdem"].terrain = pkg_t
    _load_file("xdem.terrain.freq", "xdem/terrain/freq.py")
    surfit = _load_file("xdem.terrain.surfit", "xdem/terrain/surfit.py")
    window = _load_file("xdem.terrain.window", "xdem/terrain/window.py")
    terrain = _load_file("xdem.terrain.terrain", "xdem/terrain/terrain.py")
    pkg_c = types.ModuleType("xdem.coreg")
    pkg_c.__path__ = []  # type: ignore
    sys.modules["xdem.coreg"] = pkg_c
    sys.modules["xdem"].coreg = pkg_c
    base = _load_file("xdem.coreg.base", "xdem/coreg/base.py")
    affine = _load_file("xdem.coreg.affine", "xdem/coreg/affine.py")
    ns = types.SimpleNamespace(terrain=terrain, surfit=surfit, window=window, spatialstats=spatialstats,
                               coreg_base=base, coreg_affine=affine, Affine=_SimpleAffine)
    _LOADED["ns"] = ns
    return ns
