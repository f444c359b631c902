import sys, numpy as np, torch
sys.path.insert(0, '.')
import bench_data
from oracle import terrain_oracle as to
from xdem_b200 import _engine
S, res = 32768, 5.0
z = bench_data.device_fractal_dem(S, S, 42, torch.device("cuda"))
nine = ["slope", "aspect", "hillshade", "profile_curvature", "tangential_curvature", "planform_curvature", "flowline_curvature", "max_curvature", "min_curvature"]
out = _engine.terrain_fused(z, res, nine, [], surface_fit="Florinsky", degrees=True, clip_hillshade=True)
for (r0, c0, i, j) in [(30001, 16000, 21 + 2, 27 + 2), (20000, 9999, 1 + 2, 62 + 2)]:
    r, c = r0 + i, c0 + j
    win = z[r - 2:r + 3, c - 2:c + 3].cpu().numpy()
    print("WIN", r, c, win.view(np.uint32).tolist())
    print("GPU", [float(out[k, r, c]) for k in range(9)])
    crop = z[r - 8:r + 9, c - 8:c + 9].cpu().numpy()
    ref = to.get_terrain_attribute(crop, nine, resolution=res, surface_fit="Florinsky")
    print("ORA", [float(ref[k][8, 8]) for k in range(9)])
