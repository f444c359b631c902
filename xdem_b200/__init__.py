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


def install() -> None:
    """Rebind the reference's engine seam to the B200 engine (integration mode ii).

    ``xdem.terrain.terrain`` looks up ``_get_surface_attributes`` / ``_get_windowed_indexes`` as module globals at call
    time (terrain.py:37-38, 574, 606), and ``DEM.slope()`` & co never pass ``engine`` (dem.py:448), so replacing those
    two names makes every reference caller use the GPU path transparently."""
    import xdem.terrain.terrain as ref_terrain  # noqa: raises ImportError if the reference is not installed

    from .surfit import _get_surface_attributes
    from .window import _get_windowed_indexes

    ref_terrain._get_surface_attributes = _get_surface_attributes
    ref_terrain._get_windowed_indexes = _get_windowed_indexes
