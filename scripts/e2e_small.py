"""End-to-end time of the public API on small / medium host rasters (pageable ndarray in, ndarray out)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import bench_data
import xdem_b200
dev = torch.device("cuda")
for S, attrs, fit in ((4096, ["slope"], "Horn"), (4096, ["slope", "aspect", "hillshade", "curvature"], "Florinsky"),
                      (8192, ["slope"], "Horn"), (8192, ["slope", "aspect", "hillshade", "curvature"], "Florinsky")):
    z = bench_data.device_fractal_dem(S, S, 42, dev).cpu().numpy()
    for _ in range(2):
        xdem_b200.terrain.get_terrain_attribute(z, attrs, resolution=5.0, surface_fit=fit)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        r = xdem_b200.terrain.get_terrain_attribute(z, attrs, resolution=5.0, surface_fit=fit)
        ts.append(time.perf_counter() - t0)
        del r
    dt = min(ts)
    print(f"{S}^2 {fit} {len(attrs)} planes: {dt*1e3:7.2f} ms  {S*S/dt/1e6:8.0f} Mpix/s", flush=True)
