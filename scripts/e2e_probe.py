"""e2e diagnostics on one GPU: raw pinned H2D / D2H bandwidth (torch copies), then the streamed host path with pinned
and pageable buffers at a few block sizes."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import bench_data
from xdem_b200 import _engine
S = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
dev = torch.device("cuda")
z = bench_data.device_fractal_dem(S, S, 42, dev)
hp = torch.empty((S, S), dtype=torch.float32, pin_memory=True)
for name, fn in (("D2H pinned", lambda: hp.copy_(z, non_blocking=True)), ("H2D pinned", lambda: z.copy_(hp, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3): fn()
    torch.cuda.synchronize()
    print(f"{name}: {3 * S * S * 4 / (time.perf_counter() - t0) / 1e9:.1f} GB/s", flush=True)
hp.copy_(z); torch.cuda.synchronize()
h_in = hp.numpy()
attrs = ["slope", "aspect", "hillshade", "curvature"]
kw = dict(surface_attributes=attrs, surface_fit="Florinsky", degrees=True, clip_hillshade=True)
h_out = _engine.host_planes(4, S, S, np.float32)
for rpb in (0, 256, 768, 2048):
    _engine.terrain_fused_host(h_in, 5.0, out=h_out, rows_per_block=rpb, **kw)
    t0 = time.perf_counter()
    for _ in range(3): _engine.terrain_fused_host(h_in, 5.0, out=h_out, rows_per_block=rpb, **kw)
    dt = (time.perf_counter() - t0) / 3
    print(f"pinned in/out rows_per_block={rpb}: {dt*1e3:.1f} ms  {S*S/dt/1e6:.0f} Mpix/s  D2H {S*S*16/dt/1e9:.1f} GB/s", flush=True)
pg = np.array(h_in)  # pageable copy
for rpb in (0,):
    _engine.terrain_fused_host(pg, 5.0, out=h_out, rows_per_block=rpb, **kw)
    t0 = time.perf_counter()
    for _ in range(3): _engine.terrain_fused_host(pg, 5.0, out=h_out, rows_per_block=rpb, **kw)
    dt = (time.perf_counter() - t0) / 3
    print(f"pageable in, pinned out: {dt*1e3:.1f} ms  {S*S/dt/1e6:.0f} Mpix/s", flush=True)
po = np.empty((4, S, S), dtype=np.float32)
_engine.terrain_fused_host(pg, 5.0, out=po, **kw)
t0 = time.perf_counter()
_engine.terrain_fused_host(pg, 5.0, out=po, **kw)
dt = time.perf_counter() - t0
print(f"pageable in, pageable out: {dt*1e3:.1f} ms  {S*S/dt/1e6:.0f} Mpix/s", flush=True)
assert np.array_equal(po, h_out, equal_nan=True)
import os
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
