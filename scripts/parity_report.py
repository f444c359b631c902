"""Print, for every surface-fit fixture of the reference (tests/golden/terrain_reference.npz), by how much the CUDA result
violates the strict parity criterion |x - ref| <= 1e-5 |ref| + atol(attr) (tests/parity.py) -- factor <= 1 passes --
against the Numba-engine and SciPy-engine fixtures, without any widening and with the flat-pixel mask applied to aspect
only.  Also reports the reference's own engine-vs-engine spread in the same unit.  GPU box only."""
import sys

import numpy as np

sys.path.insert(0, ".")
import xdem_b200 as xb  # noqa: E402
from oracle import terrain_oracle as to  # noqa: E402
from tests import parity  # noqa: E402

G = parity.load_golden()
SURF = ["slope", "aspect", "hillshade", "curvature", "profile_curvature", "tangential_curvature",
        "planform_curvature", "flowline_curvature", "max_curvature", "min_curvature"]


def viol(x, ref, a, where=None, atol_scale=1.0):
    atol = parity.ATOL.get(a, 1e-6) * atol_scale
    period = 360.0 if a == "aspect" else None
    if where is not None:
        x, ref = np.where(where, x, np.nan), np.where(where, ref, np.nan)
    return parity.max_violation(x, ref, parity.RTOL, atol, period)


for name in ("fractal", "noise", "integer"):
    dem = G[f"in|{name}"]
    for fit in ("Horn", "ZevenbergThorne", "Florinsky"):
        keep = to.get_terrain_attribute(dem.astype(np.float64), "slope", resolution=5.0, surface_fit=fit) > 1e-3
        for cm in ("geometric", "directional"):
            if fit == "Horn" and cm == "directional":
                continue
            attrs = SURF[:3] if fit == "Horn" else SURF
            outs = xb.terrain.get_terrain_attribute(dem, attrs, resolution=5.0, surface_fit=fit, curv_method=cm)
            for a, o in zip(attrs, outs):
                rn = G[f"surf|{name}|numba|{fit}|{cm}|deg|{a}"]
                rs = G[f"surf|{name}|scipy|{fit}|{cm}|deg|{a}"]
                w = keep if a == "aspect" else None
                scale = float(np.nanpercentile(np.abs(rn), 99))
                print(f"{name:8s} {fit:16s} {cm:11s} {a:22s} gpu-vs-numba {viol(o, rn, a, w):10.3g}  "
                      f"gpu-vs-scipy {viol(o, rs, a, w):10.3g}  numba-vs-scipy {viol(rn, rs, a, w):10.3g}  "
                      f"p99|ref| {scale:10.3g}  masks {parity.nanmask_equal(o, rn) and parity.nanmask_equal(o, rs)}",
                      flush=True)
