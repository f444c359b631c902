"""e2e of the public API at 32768^2 (pageable ndarray in) for the current XDEM_B200_HOST_THREADS."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, '.')
import bench_data
import xdem_b200
S = 32768
z = bench_data.device_fractal_dem(S, S, 42, torch.device("cuda")).cpu().numpy()
torch.cuda.empty_cache()
attrs = ["slope", "aspect", "hillshade", "curvature"]
r = xdem_b200.terrain.get_terrain_attribute(z, attrs, resolution=5.0); del r
ts = []
for _ in range(3):
    t0 = time.perf_counter()
    r = xdem_b200.terrain.get_terrain_attribute(z, attrs, resolution=5.0)
    ts.append(time.perf_counter() - t0); del r
print(f"threads={os.environ.get('XDEM_B200_HOST_THREADS', 'default')}: {min(ts)*1e3:.1f} ms  {S*S/min(ts)/1e6:.0f} Mpix/s", flush=True)
