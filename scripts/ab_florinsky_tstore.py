"""A/B of the TMA-store variant of the packed sliding Florinsky kernel: xb_set_option("florinsky_tma_store", 0/1).
Bit-equality of the planes (incl. NaN rim and ragged sizes) + device-event timing with a >L2 working set."""
import sys
import torch
sys.path.insert(0, "/root/repo")
from xdem_b200 import _engine, _lib


def dem(h, w, seed=1):
    g = torch.Generator(device="cuda").manual_seed(seed)
    z = torch.randn((h, w), generator=g, device="cuda")
    z = torch.cumsum(z, 0)
    return (1000.0 + 0.05 * torch.cumsum(z, 1)).float()


A4 = ["slope", "aspect", "hillshade", "curvature"]
A3 = ["slope", "aspect", "curvature"]
# correctness on ragged shapes (cols multiple of 4 so the plane pitch is 16 B; others must fall back and still agree)
for (h, w) in ((61, 132), (1000, 1028), (123, 64), (777, 1001), (2048, 4096)):
    z = dem(h, w, 5)
    z[h // 2, w // 3] = float("nan")
    for attrs in (A4, A3):
        _lib.set_option("florinsky_tma_store", 0)
        a = _engine.terrain_fused(z, 5.0, attrs, [], surface_fit="Florinsky", degrees=True, clip_hillshade=True)
        _lib.set_option("florinsky_tma_store", 1)
        b = _engine.terrain_fused(z, 5.0, attrs, [], surface_fit="Florinsky", degrees=True, clip_hillshade=True)
        torch.cuda.synchronize()
        ok = torch.equal(torch.isnan(a), torch.isnan(b)) and torch.equal(a.nan_to_num(), b.nan_to_num())
        # shard-like call: only rows [7, h-9) of a buffer with halo rows
        _lib.set_option("florinsky_tma_store", 0)
        a = _engine.terrain_fused(z, 5.0, attrs, [], surface_fit="Florinsky", degrees=True, clip_hillshade=True, row_begin=7, row_end=h - 9)
        _lib.set_option("florinsky_tma_store", 1)
        b = _engine.terrain_fused(z, 5.0, attrs, [], surface_fit="Florinsky", degrees=True, clip_hillshade=True, row_begin=7, row_end=h - 9)
        torch.cuda.synchronize()
        ok2 = torch.equal(torch.isnan(a), torch.isnan(b)) and torch.equal(a.nan_to_num(), b.nan_to_num())
        print(f"{h}x{w} {len(attrs)} planes: bit-equal={ok} row-sliced={ok2} shape={tuple(b.shape)}", flush=True)

for S in (16384, 32768):
    z = dem(S, S)
    for attrs in (A4, A3):
        for ts in (0, 1, 0, 1):
            _lib.set_option("florinsky_tma_store", ts)
            out = None
            for _ in range(3):
                out = _engine.terrain_fused(z, 5.0, attrs, [], surface_fit="Florinsky", degrees=True, clip_hillshade=True, out=out)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(10):
                _engine.terrain_fused(z, 5.0, attrs, [], surface_fit="Florinsky", degrees=True, clip_hillshade=True, out=out)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"{S}^2 {len(attrs)} planes tma_store={ts}: {ms:7.3f} ms  {S*S/ms/1e6:7.1f} Gpix/s  "
                  f"{(4+4*len(attrs))*S*S/ms/1e6:7.1f} GB/s", flush=True)
            del out
_lib.set_option("florinsky_tma_store", 0)
