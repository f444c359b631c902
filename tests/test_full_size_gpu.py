"""Size-independent properties at BASELINE.json's full sizes (32768^2 terrain, N = 1e6 variogram samples, 16384^2
Nuth-Kaab pair).  The oracle cannot run at these sizes in seconds, so the checks are: known answers on analytic surfaces,
exactness under integer offsets (differences are taken first), block-tiled == single launch, crops against the oracle,
pair-count conservation, order independence and recovery of an injected shift."""

from __future__ import annotations

import math

import numpy as np
import pytest

from tests import parity

pytestmark = pytest.mark.gpu

ATTRS = ["slope", "aspect", "hillshade", "curvature"]  # BASELINE.json configs[1]


def _free_gb() -> float:
    import torch

    free, _ = torch.cuda.mem_get_info()
    return free / 2**30


@pytest.mark.parametrize("fit", ["Florinsky", "Horn"])
def test_terrain_32768_properties(fit: str) -> None:
    import torch

    from oracle import terrain_oracle as to
    from xdem_b200 import _engine

    if _free_gb() < 60:
        pytest.skip("needs ~45 GB of device memory")
    S = 32768
    res = 5.0
    attrs = ATTRS[:3] if fit == "Horn" else ATTRS  # Horn has no curvatures (surfit.py:1214-1224)
    # (1) integer-valued rough DEM (exact in fp32): outputs are bit-identical under an integer offset
    g = torch.Generator(device="cuda").manual_seed(3)
    z = torch.randint(0, 64, (S, S), generator=g, device="cuda", dtype=torch.int32).float()
    z += torch.arange(S, device="cuda", dtype=torch.float32)[None, :] * 2.0
    z[20000:20003, 123:130] = float("nan")
    out = _engine.terrain_fused(z, res, attrs, [], surface_fit=fit, degrees=True, clip_hillshade=True)
    z += 4096.0
    out2 = _engine.terrain_fused(z, res, attrs, [], surface_fit=fit, degrees=True, clip_hillshade=True)
    assert torch.equal(torch.isnan(out), torch.isnan(out2))
    assert torch.equal(torch.nan_to_num(out), torch.nan_to_num(out2))
    del out2
    # (2) NaN footprint: the window half-width around the raster border and around the hole, nothing else
    h = 2 if fit == "Florinsky" else 1
    nan_expected = 2 * h * S + 2 * h * (S - 2 * h) + (3 + 2 * h) * (7 + 2 * h)
    for k in range(len(attrs)):
        assert int(torch.isnan(out[k]).sum()) == nan_expected
    # (3) crops against the float64 oracle: a corner, the hole, the far corner
    for (r0, c0) in [(0, 0), (19990, 100), (S - 64, S - 96)]:
        crop = z[r0:r0 + 64, c0:c0 + 96].cpu().numpy() - 4096.0
        ref = to.get_terrain_attribute(crop, attrs, resolution=res, surface_fit=fit)
        got = out[:, r0:r0 + 64, c0:c0 + 96].cpu().numpy()
        # exactly flat cells of the integer DEM: the reference's aspect there depends on the sign of a zero
        keep = (to.get_terrain_attribute(crop, "slope", resolution=res, surface_fit=fit) > 1e-3)[h:-h, h:-h]
        for k, a in enumerate(attrs):
            # rows/cols cut by the crop are not raster borders in the full run
            parity.assert_attr_close(got[k][h:-h, h:-h], ref[k][h:-h, h:-h], a, where=keep if a == "aspect" else None,
                                     msg=f"{fit} crop {r0},{c0}")
    # (4) row blocks with halo rows reproduce the single launch bit for bit (the multi-GPU / streaming contract)
    z -= 4096.0
    for r0 in (0, 12345, S - 4096):
        r1 = r0 + 4096
        b0, b1 = max(0, r0 - h), min(S, r1 + h)
        blk = _engine.terrain_fused(z[b0:b1], res, attrs, [], surface_fit=fit, degrees=True, clip_hillshade=True,
                                    row_begin=r0 - b0, row_end=r1 - b0)
        assert torch.equal(torch.nan_to_num(blk), torch.nan_to_num(out[:, r0:r1]))
    del out, blk
    # (5) known answer on a plane: constant slope / aspect, zero curvature, everywhere in the interior
    a_, b_ = 0.75, -0.5  # metres per pixel along x / along rows, exact in fp32 for S < 2^15 * 4
    z = (torch.arange(S, device="cuda", dtype=torch.float32)[None, :] * a_
         + torch.arange(S, device="cuda", dtype=torch.float32)[:, None] * b_)
    pattrs = ["slope", "aspect"] + ([] if fit == "Horn" else ["curvature"])
    out = _engine.terrain_fused(z, res, pattrs, [], surface_fit=fit, degrees=True)
    inner = out[:, h:-h, h:-h]
    slope = math.degrees(math.atan(math.hypot(a_, b_) / res))
    assert float((inner[0] - slope).abs().max()) <= 1e-5 * slope
    ref_plane = to.get_terrain_attribute(z[:8, :8].cpu().numpy(), "aspect", resolution=res, surface_fit=fit)
    aspect = float(ref_plane[4, 4])
    assert float((inner[1] - aspect).abs().max()) <= 1e-5 * aspect + 1e-4
    if fit != "Horn":
        assert float(inner[2].abs().max()) <= 1e-6


@pytest.mark.parametrize("fit", ["Florinsky", "ZevenbergThorne"])
def test_benchmarked_dem_parity_on_crops(fit: str) -> None:
    """The raster bench.py times (bench_data.device_fractal_dem, seed 42, 32768^2, |z| up to ~1e5 m) through the kernels
    the benchmark launches, checked on crops against the float64 oracle at the UNWIDENED criterion
    (|x - ref| <= 1e-5 |ref| + atol; reference accumulation: float64, surfit.py:1044): the four corners, the crop
    holding the largest |z|, the crop holding the smallest |gradient| region sampled, and four interior crops --
    for the headline 4-attribute request and for the 9-attribute request of BASELINE config 4."""
    import torch

    import bench_data
    from oracle import terrain_oracle as to
    from xdem_b200 import _engine

    if _free_gb() < 60:
        pytest.skip("needs ~45 GB of device memory")
    S, res, h = 32768, 5.0, 2 if fit == "Florinsky" else 1
    z = bench_data.device_fractal_dem(S, S, 42, torch.device("cuda"))
    amax = int(torch.argmax(z.abs()))
    r_hi, c_hi = min(max(amax // S - 32, 0), S - 64), min(max(amax % S - 48, 0), S - 96)
    crops = [(0, 0), (0, S - 96), (S - 64, 0), (S - 64, S - 96), (r_hi, c_hi),
             (4096, 4000), (12345, 23456), (20000, 9999), (30001, 16000), (16384 - 32, 16384 - 48)]
    nine = ["slope", "aspect", "hillshade", "profile_curvature", "tangential_curvature", "planform_curvature",
            "flowline_curvature", "max_curvature", "min_curvature"]
    for attrs in (ATTRS, nine):
        out = _engine.terrain_fused(z, res, attrs, [], surface_fit=fit, degrees=True, clip_hillshade=True)
        worst = 0.0
        for (r0, c0) in crops:
            crop = z[r0:r0 + 64, c0:c0 + 96].cpu().numpy()
            ref = to.get_terrain_attribute(crop, attrs, resolution=res, surface_fit=fit)
            got = out[:, r0:r0 + 64, c0:c0 + 96].cpu().numpy()
            keep = (to.get_terrain_attribute(crop.astype(np.float64), "slope", resolution=res, surface_fit=fit)
                    > 1e-3)[h:-h, h:-h]
            for k, a in enumerate(attrs):
                g_, r_ = got[k][h:-h, h:-h], ref[k][h:-h, h:-h]
                parity.assert_attr_close(g_, r_, a, where=keep if a == "aspect" else None,
                                         msg=f"bench DEM {fit} crop {r0},{c0}")
                worst = max(worst, parity.violation(g_, r_, a, where=keep if a == "aspect" else None))
        print(f"bench DEM {fit} {len(attrs)} attrs: worst violation factor {worst:.3g} (<= 1 passes)")
        del out


def test_variogram_1e6_properties() -> None:
    """N = 1e6 samples (5e11 pairs): every pair lands in exactly one class, the counts do not depend on sample order,
    and the sums of squares agree to float64 re-association."""
    import torch

    from xdem_b200 import spatialstats as xs

    g = torch.Generator(device="cuda").manual_seed(9)
    N, S = 1_000_000, 32768
    lin = torch.randperm(S * S // 64, generator=g, device="cuda")[:N].to(torch.int64) * 64 + 11
    x, y = lin % S, lin // S
    v = torch.randn(N, generator=g, device="cuda")
    edges = np.linspace(0, 1.5 * S * 5.0, 31)[1:]
    _, cnt, ssq = xs.pairwise_lag_binning(x, y, v, edges, 5.0)
    assert int(cnt.sum()) == N * (N - 1) // 2
    perm = torch.randperm(N, generator=g, device="cuda")
    _, cnt2, ssq2 = xs.pairwise_lag_binning(x[perm], y[perm], v[perm], edges, 5.0)
    assert np.array_equal(cnt, cnt2) and np.allclose(ssq, ssq2, rtol=1e-9)
    # sum over classes of sum (dv)^2 == N * sum v^2 - (sum v)^2; each dv is rounded to float32 once before squaring
    # (like the reference's float32 values), so the identity holds to the float32 parity tolerance
    v64 = v.double()
    total = float(N * (v64 * v64).sum() - v64.sum() ** 2)
    assert float(ssq.sum()) == pytest.approx(total, rel=1e-5)


def test_nuth_kaab_16384_recovers_shift() -> None:
    """BASELINE configs[3] size: a 16384^2 pair with an injected sub-pixel shift; the fit recovers it, and fitting the
    tba against itself returns zero."""
    import torch

    from xdem_b200 import coreg

    if _free_gb() < 30:
        pytest.skip("needs ~20 GB of device memory")
    S = 16384
    yy = torch.arange(S, device="cuda", dtype=torch.float32)[:, None]
    xx = torch.arange(S, device="cuda", dtype=torch.float32)[None, :]

    def surf(x: "torch.Tensor", y: "torch.Tensor") -> "torch.Tensor":
        return (300.0 * torch.sin(x * (2 * math.pi / 1700.0)) * torch.cos(y * (2 * math.pi / 2300.0))
                + 120.0 * torch.sin((x + 0.6 * y) * (2 * math.pi / 410.0)) + 0.02 * x)

    sx_px, sy_px, dz = 0.37, -0.61, 1.5
    ref = surf(xx, yy)
    tba = surf(xx + sx_px, yy + sy_px) + dz
    res = 10.0
    nkc = coreg.NuthKaab(subsample=1)
    nkc.fit(ref, tba, transform=(res, 0, 0, 0, -res, 0), random_state=1)
    tx, ty, tz = nkc.to_translations()
    # sign convention of the reference: the translation to apply to tba to match ref
    assert abs(abs(tx) - sx_px * res) < 0.05 * res and abs(abs(ty) - abs(sy_px) * res) < 0.05 * res
    assert tz == pytest.approx(-dz, abs=0.05)
    nk0 = coreg.NuthKaab(subsample=1)
    nk0.fit(tba, tba, transform=(res, 0, 0, 0, -res, 0), random_state=1)
    t0 = nk0.to_translations()
    assert max(abs(t) for t in t0) < 1e-3
