"""Synthetic inputs shared by bench.py, bench_extra.py and the full-size GPU tests (same generator => the parity tests
check the very rasters the benchmark times).  torch only; nothing here is part of the product."""

from __future__ import annotations

import torch


def device_fractal_dem(rows: int, cols: int, seed: int, device: "torch.device", out: "torch.Tensor | None" = None,
                       chunk: int = 4096) -> "torch.Tensor":
    """z = 1000 + 0.05 * cumsum(cumsum(N(0,1), 0), 1) in float32 (SURVEY 8d's generator family), built on the device
    in row chunks with a carried column sum.  Reaches |z| ~ 1e5 at 32768^2, i.e. ulp(z) up to 2^-7 m."""
    g = torch.Generator(device=device).manual_seed(seed)
    if out is None:
        out = torch.empty((rows, cols), dtype=torch.float32, device=device)
    carry = torch.zeros((1, cols), dtype=torch.float32, device=device)
    for r0 in range(0, rows, chunk):
        r1 = min(rows, r0 + chunk)
        n = torch.randn((r1 - r0, cols), generator=g, device=device)
        blk = torch.cumsum(n, dim=0) + carry
        carry = blk[-1:].clone()
        out[r0:r1] = 1000.0 + 0.05 * torch.cumsum(blk, dim=1)
    return out
