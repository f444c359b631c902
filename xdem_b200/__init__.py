"""xdem_b200 -- B200-native (sm_100a) implementation of the dense-raster hot path of GlacioHack/xdem.

* ``xdem_b200.terrain``      -- drop-in for ``xdem.terrain`` (get_terrain_attribute + wrappers), one fused CUDA pass
* ``xdem_b200.surfit/window``-- the reference's engine seam functions (``_get_surface_attributes`` /
                                ``_get_windowed_indexes``)
* ``xdem_b200.spatialstats`` -- ``sample_empirical_variogram`` (all-pairs lag binning on the GPU), ``nd_binning``
* ``xdem_b200.freq``         -- texture shading (our kernels around two cuFFT transforms)
* ``xdem_b200.coreg``        -- ``NuthKaab`` / ``nuth_kaab`` (slope/aspect + aspect-binned medians on the GPU)
* ``xdem_b200.distributed``  -- row-sharded multi-GPU drivers (NCCL halo exchange / histogram all-reduce)
* ``xdem_b200.install()``    -- rebinds the seam inside an importable ``xdem`` so DEM.slope() etc. run on the GPU

There is no CPU fallback: every compute entry point raises if the CUDA library or a CUDA device is missing.
"""

from __future__ import annotations

from . import terrain  # noqa: F401
from ._lib import XdemB200Error  # noqa: F401
from .terrain import get_terrain_attribute  # noqa: F401

__version__ = "0.1.0"

_ORIGINALS: dict[tuple[int, str], tuple[object, object]] = {}  # (id(module), name) -> (module, what install() replaced)


def _rebind(module: object, name: str, new: object) -> None:
    _ORIGINALS.setdefault((id(module), name), (module, getattr(module, name)))
    setattr(module, name, new)


def install() -> None:
    """Rebind the reference's engine seams to the B200 engine (integration mode ii).

    * ``xdem.terrain.terrain`` looks up ``_get_surface_attributes`` / ``_get_windowed_indexes`` as module globals at
      call time (terrain.py:37-38, 574, 606), and ``DEM.slope()`` & co never pass ``engine`` (dem.py:448), so replacing
      those two names makes every reference caller use the GPU path transparently (host rasters are streamed through
      the GPU in row blocks, see ``xdem_b200.surfit``).
    * ``xdem.coreg.affine.nuth_kaab`` (affine.py:539-609) is what ``NuthKaab._fit_rst_pts`` calls (affine.py:2509): it
      is rebound to ``xdem_b200.coreg.nuth_kaab`` for raster-raster inputs; point-cloud inputs (GeoDataFrame) and
      options outside the B200 path (``fit_or_bin="fit"``, a custom ``bin_statistic``) keep the reference function.
    * ``xdem.spatialstats._get_pdist_empirical_variogram`` / ``_get_cdist_empirical_variogram`` (spatialstats.py:1064,
      1186) -- the two places the reference hands the O(N^2) pair work to scikit-gstat -- are rebound to the GPU pair
      kernels, so ``sample_empirical_variogram`` (and ``DEM.estimate_uncertainty``, dem.py:674) keep the reference's
      own glue (coordinates, default bins, aggregation over runs) and only the pair work moves.
    Each part is skipped when its reference module is not importable."""
    import importlib

    ref_terrain = importlib.import_module("xdem.terrain.terrain")  # raises ImportError if the reference is missing

    from .surfit import _get_surface_attributes
    from .window import _get_windowed_indexes

    _rebind(ref_terrain, "_get_surface_attributes", _get_surface_attributes)
    _rebind(ref_terrain, "_get_windowed_indexes", _get_windowed_indexes)

    try:
        ref_affine = importlib.import_module("xdem.coreg.affine")
    except ImportError:
        ref_affine = None
    if ref_affine is not None and hasattr(ref_affine, "nuth_kaab"):
        from . import coreg

        original = _ORIGINALS.get((id(ref_affine), "nuth_kaab"), (None, ref_affine.nuth_kaab))[1]  # idempotent install()
        _rebind(ref_affine, "nuth_kaab", coreg.make_reference_hook(original))

    try:
        ref_ss = importlib.import_module("xdem.spatialstats")
    except ImportError:
        ref_ss = None
    if ref_ss is not None and hasattr(ref_ss, "_get_pdist_empirical_variogram"):
        from . import spatialstats as xs

        _rebind(ref_ss, "_get_pdist_empirical_variogram", xs._get_pdist_empirical_variogram)
        if hasattr(ref_ss, "_get_cdist_empirical_variogram"):
            _rebind(ref_ss, "_get_cdist_empirical_variogram", xs._get_cdist_empirical_variogram)


def uninstall() -> None:
    """Undo ``install()``: put the reference's own functions back."""
    for (_, name), (module, original) in list(_ORIGINALS.items()):
        setattr(module, name, original)
    _ORIGINALS.clear()
