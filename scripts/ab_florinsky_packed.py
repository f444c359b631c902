"""A/B of the packed (f32x2) and scalar sliding Florinsky kernels: xb_set_option("florinsky_packed", 0/1)."""
import sys
import torch
sys.path.insert(0, "/root/repo")
from xdem_b200 import _engine, _lib

for S in (16384, 32768):
    g = torch.Generator(device="cuda").manual_seed(1)
    z = torch.randn((S, S), generator=g, device="cuda")
    z = torch.cumsum(z, 0)
    z = (1000.0 + 0.05 * torch.cumsum(z, 1)).float()
    for attrs in (["slope", "aspect", "hillshade", "curvature"], ["slope", "aspect", "curvature"]):
        outs = {}
        for packed in (0, 1, 0, 1):
            _lib.set_option("florinsky_packed", packed)
            for _ in range(3):
                o = _engine.terrain_fused(z, 5.0, attrs, [], surface_fit="Florinsky", degrees=True, clip_hillshade=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(10):
                o = _engine.terrain_fused(z, 5.0, attrs, [], surface_fit="Florinsky", degrees=True, clip_hillshade=True)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            outs[packed] = o
            print(f"{S}^2 {'+'.join(a[:5] for a in attrs):28s} packed={packed}: {ms:7.3f} ms  {S*S/ms/1e6:7.1f} Gpix/s  "
                  f"{(4+4*len(attrs))*S*S/ms/1e6:7.1f} GB/s", flush=True)
            del o
        a, b = outs[0], outs[1]
        nan_eq = bool(torch.equal(torch.isnan(a), torch.isnan(b)))
        d = (a - b).abs()
        rel = (d / (a.abs() + 1e-3)).nan_to_num(0).amax(dim=(1, 2)).tolist()
        print(f"   packed vs scalar: NaN masks equal={nan_eq}  max rel diff per attribute {['%.2e' % r for r in rel]}", flush=True)
        del outs, a, b, d
_lib.set_option("florinsky_packed", 1)
