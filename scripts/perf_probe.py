import torch, time, sys, json
sys.path.insert(0, '/root/repo')
from xdem_b200 import _engine
size = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dev = torch.device('cuda')
g = torch.Generator(device=dev).manual_seed(42)
z = torch.empty((size, size), device=dev)
carry = torch.zeros((1, size), device=dev)
for r0 in range(0, size, 4096):
    n = torch.randn((min(4096, size - r0), size), generator=g, device=dev)
    blk = torch.cumsum(n, 0) + carry; carry = blk[-1:].clone()
    z[r0:r0 + blk.shape[0]] = 1000 + 0.05 * torch.cumsum(blk, 1)
del n, blk
def bench(name, **kw):
    out = None
    for _ in range(2): out = _engine.terrain_fused(z, 5.0, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): _engine.terrain_fused(z, 5.0, out=out, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    nout = out.shape[0]
    gb = size * size * (4 + 4 * nout) / 1e9
    print(f"{name:55s} {ms:8.3f} ms  {size*size/ms/1e6:8.1f} Gpix/s  {gb/ms*1e3:8.1f} GB/s algorithmic ({nout} planes)", flush=True)
    del out
S4 = ["slope", "aspect", "hillshade", "curvature"]
ALLC = ["slope", "aspect", "hillshade", "profile_curvature", "tangential_curvature", "planform_curvature", "flowline_curvature", "max_curvature", "min_curvature"]
for fit in ["Horn", "ZevenbergThorne", "Florinsky"]:
    bench(f"{fit} slope", surface_attributes=["slope"], surface_fit=fit, degrees=True)
    bench(f"{fit} slope+aspect", surface_attributes=["slope","aspect"], surface_fit=fit, degrees=True)
    if fit != "Horn":
        bench(f"{fit} slope+aspect+curvature", surface_attributes=["slope","aspect","curvature"], surface_fit=fit, degrees=True)
        bench(f"{fit} slope+aspect+hillshade+curvature", surface_attributes=S4, surface_fit=fit, degrees=True, clip_hillshade=True)
        bench(f"{fit} 9 surface attrs", surface_attributes=ALLC, surface_fit=fit, degrees=True, clip_hillshade=True)
        bench(f"{fit} curvature only", surface_attributes=["curvature"], surface_fit=fit)
bench("3x3 TPI", windowed_indexes=["topographic_position_index"])
bench("3x3 TPI+TRI+roughness", windowed_indexes=["topographic_position_index","terrain_ruggedness_index","roughness"])
bench("3x3 TPI+TRI+roughness+rugosity", windowed_indexes=["topographic_position_index","terrain_ruggedness_index","roughness","rugosity"])
bench("5x5 TPI+TRI+roughness", windowed_indexes=["topographic_position_index","terrain_ruggedness_index","roughness"], window_size=5)
bench("ALL 13 (Florinsky + 3x3)", surface_attributes=ALLC, windowed_indexes=["topographic_position_index","terrain_ruggedness_index","roughness","rugosity"], surface_fit="Florinsky", degrees=True, clip_hillshade=True)
