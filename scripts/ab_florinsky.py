"""A/B: generic vs sliding Florinsky kernels (values + timing).  Env: XB_FL_GENERIC=1 forces generic; XB_FL_OCC=3."""
import os, sys, subprocess, json
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch, numpy as np
    sys.path.insert(0, '/root/repo')
    from xdem_b200 import _engine
    size = 16384
    dev = torch.device('cuda')
    g = torch.Generator(device=dev).manual_seed(42)
    z = torch.empty((size, size), device=dev); carry = torch.zeros((1, size), device=dev)
    for r0 in range(0, size, 4096):
        n = torch.randn((4096, size), generator=g, device=dev); blk = torch.cumsum(n, 0) + carry; carry = blk[-1:].clone()
        z[r0:r0 + 4096] = 1000 + 0.05 * torch.cumsum(blk, 1)
    z[5000:5003, 7000:7010] = float('nan')
    res = {}
    for name, attrs in (("s4", ["slope", "aspect", "hillshade", "curvature"]), ("slope", ["slope"]), ("all9", ["slope", "aspect", "hillshade", "profile_curvature", "tangential_curvature", "planform_curvature", "flowline_curvature", "max_curvature", "min_curvature"])):
        out = _engine.terrain_fused(z, 5.0, attrs, surface_fit="Florinsky", degrees=True, clip_hillshade=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): _engine.terrain_fused(z, 5.0, attrs, surface_fit="Florinsky", degrees=True, clip_hillshade=True, out=out)
        e1.record(); torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / 5
        torch.save(out[:, 4990:5300, 6900:7300].cpu(), f"/tmp/ab_{os.environ.get('TAG')}_{name}.pt")
    print(json.dumps(res))
else:
    import torch
    outs = {}
    for tag, env in (("generic", {"XB_FL_GENERIC": "1"}), ("slide2", {}), ("slide3", {"XB_FL_OCC": "3"})):
        e = dict(os.environ); e.update(env); e["TAG"] = tag
        r = subprocess.run([sys.executable, __file__, "child"], env=e, capture_output=True, text=True)
        print(tag, r.stdout.strip().split("\n")[-1], r.stderr[-300:] if r.returncode else "")
    for name in ("s4", "slope", "all9"):
        a = torch.load(f"/tmp/ab_generic_{name}.pt"); b = torch.load(f"/tmp/ab_slide2_{name}.pt"); c = torch.load(f"/tmp/ab_slide3_{name}.pt")
        same_nan = bool((torch.isnan(a) == torch.isnan(b)).all())
        d = (torch.nan_to_num(a) - torch.nan_to_num(b)).abs()
        rel = (d / torch.nan_to_num(a).abs().clamp_min(1e-3)).max().item()
        print(name, "nan equal", same_nan, "max abs diff", d.max().item(), "max rel", rel, "slide2==slide3", bool(torch.equal(torch.nan_to_num(b), torch.nan_to_num(c))))
