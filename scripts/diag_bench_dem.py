import sys, numpy as np, torch
sys.path.insert(0, '.')
import bench_data
from oracle import terrain_oracle as to
from xdem_b200 import _engine
from tests import parity
S, res = 32768, 5.0
z = bench_data.device_fractal_dem(S, S, 42, torch.device("cuda"))
amax = int(torch.argmax(z.abs()))
r_hi, c_hi = min(max(amax // S - 32, 0), S - 64), min(max(amax % S - 48, 0), S - 96)
crops = [(0, 0), (0, S - 96), (S - 64, 0), (S - 64, S - 96), (r_hi, c_hi), (4096, 4000), (12345, 23456), (20000, 9999), (30001, 16000), (16384 - 32, 16384 - 48)]
nine = ["slope", "aspect", "hillshade", "profile_curvature", "tangential_curvature", "planform_curvature", "flowline_curvature", "max_curvature", "min_curvature"]
for fit, h in (("Florinsky", 2), ("ZevenbergThorne", 1)):
    out = _engine.terrain_fused(z, res, nine, [], surface_fit=fit, degrees=True, clip_hillshade=True)
    for (r0, c0) in crops:
        crop = z[r0:r0 + 64, c0:c0 + 96].cpu().numpy()
        ref = to.get_terrain_attribute(crop, nine, resolution=res, surface_fit=fit)
        got = out[:, r0:r0 + 64, c0:c0 + 96].cpu().numpy()
        keep = (to.get_terrain_attribute(crop.astype(np.float64), "slope", resolution=res, surface_fit=fit) > 1e-3)[h:-h, h:-h]
        for k, a in enumerate(nine):
            g_, r_ = got[k][h:-h, h:-h], ref[k][h:-h, h:-h]
            v = parity.violation(g_, r_, a, where=keep if a == "aspect" else None)
            if v > 0.5:
                m = np.isfinite(r_) & np.isfinite(g_)
                d = np.abs(g_.astype(np.float64) - r_.astype(np.float64)); d[~m] = 0
                i = np.unravel_index(np.argmax(d / (1e-5 * np.abs(r_) + parity.ATOL[a])), d.shape)
                sl = to.get_terrain_attribute(crop.astype(np.float64), "slope", resolution=res, surface_fit=fit)[h:-h, h:-h][i]
                print(fit, (r0, c0), a, f"viol {v:.3g} at {i}: got {g_[i]!r} ref {r_[i]!r} slope_deg {sl:.4g} zmax {np.abs(crop).max():.1f}", "nanmask", parity.nanmask_equal(g_, r_))
    del out
