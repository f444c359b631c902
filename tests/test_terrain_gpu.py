"""GPU parity tests of the fused terrain kernel, through the public API -> ctypes C ABI -> CUDA.

Compared against (a) fixtures produced by the unmodified reference (both engines), (b) the NumPy oracle on seeded
inputs, (c) the reference's known-answer tests, (d) size-independent properties at large sizes."""

from __future__ import annotations

import numpy as np
import pytest

from tests import parity

pytestmark = pytest.mark.gpu

SURF = ["slope", "aspect", "hillshade", "curvature", "profile_curvature", "tangential_curvature",
        "planform_curvature", "flowline_curvature", "max_curvature", "min_curvature"]
WIN = ["topographic_position_index", "terrain_ruggedness_index", "roughness", "rugosity"]


@pytest.fixture(scope="module")
def G() -> dict[str, np.ndarray]:
    return parity.load_golden()


@pytest.fixture(scope="module")
def xb():
    import xdem_b200

    return xdem_b200


def _keep(dem: np.ndarray, fit: str) -> np.ndarray:
    from oracle import terrain_oracle as to

    s = to.get_terrain_attribute(dem.astype(np.float64), "slope", resolution=5.0, surface_fit=fit)
    return s > 1e-3


@pytest.mark.parametrize("name", ["fractal", "noise", "integer"])
@pytest.mark.parametrize("fit", ["Horn", "ZevenbergThorne", "Florinsky"])
@pytest.mark.parametrize("cm", ["geometric", "directional"])
def test_surface_vs_reference_fixtures(xb, G, name: str, fit: str, cm: str) -> None:
    if fit == "Horn" and cm == "directional":
        pytest.skip("Horn has no curvatures")
    dem = G[f"in|{name}"]
    attrs = SURF[:3] if fit == "Horn" else SURF
    outs = xb.terrain.get_terrain_attribute(dem, attrs, resolution=5.0, surface_fit=fit, curv_method=cm)
    keep = _keep(dem, fit)
    for a, o in zip(attrs, outs):
        where = keep if a == "aspect" else None  # flat pixels are excluded for aspect only (test_terrain.py:168-169)
        ref_n = G[f"surf|{name}|numba|{fit}|{cm}|deg|{a}"]
        ref_s = G[f"surf|{name}|scipy|{fit}|{cm}|deg|{a}"]
        # (1) the stated criterion, unwidened, against the reference's float64 (Numba) engine
        parity.assert_attr_close(o, ref_n, a, where=where, msg=f"{name}/numba/{fit}/{cm}")
        # (2) the SciPy engine rounds its coefficients to float32 (pinned bit-for-bit by the oracle's
        # coef_round=float32 mode, tests/test_oracle_terrain.py).  Same criterion; where the reference's own two
        # engines are further apart than that, the bound is their spread (triangle inequality), never a constant.
        assert parity.nanmask_equal(o, ref_s), f"{name}/scipy/{fit}/{cm}: NaN masks differ"
        v_gs = parity.violation(o, ref_s, a, where=where)
        if v_gs > 1.0:
            v_gn, v_ns = parity.violation(o, ref_n, a, where=where), parity.violation(ref_n, ref_s, a, where=where)
            assert v_gs <= v_gn + v_ns * (1 + 1e-6), (name, fit, cm, a, v_gs, v_gn, v_ns)


def test_radians_and_seam(xb, G) -> None:
    from xdem_b200.surfit import _get_surface_attributes

    dem = G["in|fractal"]
    for fit in ("Horn", "ZevenbergThorne", "Florinsky"):
        out = xb.terrain.get_terrain_attribute(dem, ["slope", "aspect"], resolution=5.0, surface_fit=fit,
                                               degrees=False)
        seam = _get_surface_attributes(dem, 5.0, ["aspect", "hillshade", "slope"], surface_fit=fit)
        assert seam.shape == (3,) + dem.shape and seam.dtype == np.float32
        for a, o in zip(["slope", "aspect"], out):
            ref = G[f"surf|fractal|numba|{fit}|geometric|rad|{a}"]
            parity.assert_attr_close(o, ref, a, degrees=False, where=_keep(dem, fit) if a == "aspect" else None,
                                     msg=f"rad {fit}")
        assert np.array_equal(seam[2], out[0], equal_nan=True)
        assert np.array_equal(seam[0], out[1], equal_nan=True)


def test_hillshade_variants(xb, G) -> None:
    dem = G["in|fractal"]
    for az, alt, zf in ((45.0, 10.0, 1.0), (315.0, 45.0, 3.0), (200.0, 80.0, 0.5)):
        for fit in ("Horn", "Florinsky"):
            o = xb.terrain.hillshade(dem, surface_fit=fit, azimuth=az, altitude=alt, z_factor=zf, resolution=5.0)
            ref = G[f"hs|fractal|numba|{fit}|{az}|{alt}|{zf}"]
            parity.assert_attr_close(o, ref, "hillshade", msg=f"hs {fit} {az} {alt} {zf}")
            assert np.nanmin(o) >= 0 and np.nanmax(o) <= 255


@pytest.mark.parametrize("name", ["fractal", "integer", "small_int"])
@pytest.mark.parametrize("w", [3, 5])
@pytest.mark.parametrize("tm", ["Riley", "Wilson"])
def test_windowed_vs_reference_fixtures(xb, G, name: str, w: int, tm: str) -> None:
    dem = G[f"in|{name}"]
    attrs = WIN if w == 3 else WIN[:3]
    outs = xb.terrain.get_terrain_attribute(dem, attrs, resolution=5.0, window_size=w, tri_method=tm)
    for a, o in zip(attrs, outs):
        ref_n = G[f"win|{name}|numba|{w}|{tm}|{a}"]
        ref_s = G[f"win|{name}|scipy|{w}|{tm}|{a}"]
        assert o.dtype == np.float32
        assert parity.nanmask_equal(o, ref_n) and parity.nanmask_equal(o, ref_s), a
        if a == "roughness":
            assert np.array_equal(o, ref_s, equal_nan=True) and np.array_equal(o, ref_n, equal_nan=True)
        elif a == "rugosity":
            assert np.array_equal(o, ref_s, equal_nan=True), "rugosity must be bit-exact vs the SciPy engine"
        elif a == "topographic_position_index":
            if name != "fractal":
                assert np.array_equal(o, ref_s, equal_nan=True), "TPI must be bit-exact on integer-valued DEMs"
            if w == 3:
                assert np.array_equal(o, ref_n, equal_nan=True)
            assert np.nanmax(np.abs(o - ref_s)) <= 4 * np.spacing(np.float32(np.nanmax(np.abs(dem)) * w * w))
        else:
            assert np.array_equal(o, ref_n, equal_nan=True), "TRI must be bit-exact vs the Numba engine"
            if name == "small_int" or tm == "Wilson":
                assert np.array_equal(o, ref_s, equal_nan=True)


def test_float64_input(xb, G) -> None:
    dem = G["in|fractal64"]
    for fit in ("ZevenbergThorne", "Florinsky"):
        outs = xb.terrain.get_terrain_attribute(dem, SURF, resolution=5.0, surface_fit=fit)
        for a, o in zip(SURF, outs):
            ref = G[f"surf|fractal64|numba|{fit}|geometric|deg|{a}"]
            assert o.dtype == np.float64
            parity.assert_attr_close(o, ref, a, rtol=1e-9, atol_scale=1e-4, msg=f"f64 {fit}")
    outs = xb.terrain.get_terrain_attribute(dem, WIN, resolution=5.0)
    for a, o in zip(WIN, outs):
        ref = G[f"win|fractal64|scipy|3|Riley|{a}"]
        assert o.dtype == np.float64
        parity.assert_attr_close(o, ref, a, rtol=1e-12, atol_scale=1e-6, msg="f64 windowed")


def test_fused_equals_separate_and_order(xb, G) -> None:
    """multi-attribute == single-attribute (test_terrain.py:248-293), any request order, surface+windowed fused."""
    dem = G["in|fractal"]
    req = ["roughness", "slope", "rugosity", "max_curvature", "aspect", "topographic_position_index", "hillshade"]
    outs = xb.terrain.get_terrain_attribute(dem, req, resolution=5.0)
    for a, o in zip(req, outs):
        single = xb.terrain.get_terrain_attribute(dem, a, resolution=5.0)
        assert np.array_equal(o, single, equal_nan=True), a
    # mixed halos: 3x3 fit with 5x5 window and 5x5 fit with 3x3 window
    a1 = xb.terrain.get_terrain_attribute(dem, ["slope", "roughness"], resolution=5.0, surface_fit="Horn",
                                          window_size=5)
    assert np.array_equal(a1[0], xb.terrain.slope(dem, surface_fit="Horn", resolution=5.0), equal_nan=True)
    assert np.array_equal(a1[1], xb.terrain.roughness(dem, window_size=5), equal_nan=True)


def test_reference_doctests(xb) -> None:
    """terrain.py:268-279, 799-813, 1484-1493, 1553-1562."""
    dem = np.repeat(np.arange(3), 3)[::-1].reshape(3, 3)
    s, a = xb.terrain.get_terrain_attribute(dem, ["slope", "aspect"], resolution=1, surface_fit="ZevenbergThorne")
    assert s[1, 1] == np.float32(45.0) and a[1, 1] == np.float32(180.0)
    assert xb.terrain.aspect(np.tile(np.arange(3), (3, 1)), surface_fit="ZevenbergThorne")[1, 1] == np.float32(270.0)
    d3 = np.zeros((3, 3), dtype="int32")
    d3[1, 1] = 1
    assert xb.terrain.topographic_position_index(d3)[1, 1] == np.float32(1.0)
    assert xb.terrain.terrain_ruggedness_index(d3)[1, 1] == np.float32(2.828427)
    assert xb.terrain.roughness(d3)[1, 1] == np.float32(1.0)


def test_rugosity_known_answers(xb) -> None:
    """test_window.py:21-68."""
    dem = np.array([[190, 170, 155], [183, 165, 145], [175, 160, 122]], dtype="float32")
    assert xb.terrain.rugosity(dem, resolution=100.0)[1, 1] == pytest.approx(10280.48 / 10000.0, rel=1e-4)
    for dh in np.linspace(0.01, 100, 3):
        for resolution in np.linspace(0.01, 100, 3):
            d = np.array([[1, 1, 1], [1, 1 + dh, 1], [1, 1, 1]], dtype="float64")
            r = xb.terrain.rugosity(d, resolution=resolution)
            side1 = np.sqrt(2 * resolution**2 + dh**2) / 2.0
            side2 = np.sqrt(resolution**2 + dh**2) / 2.0
            side3 = resolution / 2.0
            s = (side1 + side2 + side3) / 2.0
            A = np.sqrt(s * (s - side1) * (s - side2) * (s - side3))
            assert r[1, 1] == pytest.approx(8 * A / resolution**2, rel=1e-6)


@pytest.mark.parametrize("fit,w", [("Horn", 3), ("ZevenbergThorne", 3), ("Florinsky", 5)])
def test_nan_propagation_surface(xb, fit: str, w: int) -> None:
    """test_surfit.py:467-518: NaN mask == binary_dilation(nanmask, ones(w,w)) + hw-wide border, exactly."""
    from scipy.ndimage import binary_dilation

    rng = np.random.default_rng(42)
    dem = rng.normal(size=(37, 53)).astype(np.float32)
    dem[rng.integers(0, 37, 9), rng.integers(0, 53, 9)] = np.nan
    dem[20, 30] = np.inf
    attrs = SURF[:3] if fit == "Horn" else SURF
    outs = xb.terrain.get_terrain_attribute(dem, attrs, resolution=1.0, surface_fit=fit)
    expected = binary_dilation(~np.isfinite(dem), structure=np.ones((w, w), bool))
    hw = w // 2
    expected[:hw] = expected[-hw:] = True
    expected[:, :hw] = expected[:, -hw:] = True
    for a, o in zip(attrs, outs):
        assert np.array_equal(np.isnan(o), expected), (fit, a)


@pytest.mark.parametrize("w", [3, 5])
def test_nan_propagation_windowed(xb, w: int) -> None:
    """test_window.py:194-239."""
    from scipy.ndimage import binary_dilation

    rng = np.random.default_rng(7)
    dem = rng.normal(size=(29, 41)).astype(np.float32)
    dem[rng.integers(0, 29, 6), rng.integers(0, 41, 6)] = np.nan
    attrs = WIN if w == 3 else WIN[:3]
    outs = xb.terrain.get_terrain_attribute(dem, attrs, resolution=1.0, window_size=w)
    expected = binary_dilation(~np.isfinite(dem), structure=np.ones((w, w), bool))
    hw = w // 2
    expected[:hw] = expected[-hw:] = True
    expected[:, :hw] = expected[:, -hw:] = True
    for a, o in zip(attrs, outs):
        assert np.array_equal(np.isnan(o), expected), a


def test_synthetic_curvature_signs(xb) -> None:
    """test_surfit.py:228-411 (planes -> 0, ridge/trough antisymmetry)."""
    yy, xx = np.mgrid[0:21, 0:21].astype(np.float32)
    plane = (2.0 * xx + 3.0 * yy + 100).astype(np.float32)
    curv = [a for a in SURF if "curvature" in a]
    for fit in ("ZevenbergThorne", "Florinsky"):
        outs = xb.terrain.get_terrain_attribute(plane, curv, resolution=1.0, surface_fit=fit)
        for a, o in zip(curv, outs):
            assert np.nanmax(np.abs(o)) < 1e-4, (fit, a)
        ridge = (-np.abs(xx - 10) * 2 + 50).astype(np.float32)
        r = xb.terrain.get_terrain_attribute(ridge, curv, resolution=1.0, surface_fit=fit)
        t = xb.terrain.get_terrain_attribute(-ridge, curv, resolution=1.0, surface_fit=fit)
        for a, ro, to_ in zip(curv, r, t):
            if a in ("max_curvature", "min_curvature"):
                continue
            assert np.allclose(ro, -to_, equal_nan=True, atol=1e-5), (fit, a)
        rmax = xb.terrain.max_curvature(ridge, resolution=1.0, surface_fit=fit)
        tmin = xb.terrain.min_curvature(-ridge, resolution=1.0, surface_fit=fit)
        assert np.allclose(rmax, -tmin, equal_nan=True, atol=1e-5)


@pytest.mark.parametrize("shape", [(1, 1), (2, 7), (5, 130), (257, 131), (64, 128), (300, 1027)])
def test_odd_shapes_vs_oracle(xb, shape) -> None:
    """Ragged / tiny / unaligned rasters (non-TMA loader, scalar store tails) against the oracle."""
    from oracle import synth
    from oracle import terrain_oracle as to

    dem = synth.fractal_dem(shape, seed=3)
    if dem.size > 50:
        dem = synth.inject_nans(dem, frac=0.002, hole=2)
    req = ["slope", "aspect", "hillshade", "curvature", "topographic_position_index", "roughness"]
    for fit in ("ZevenbergThorne", "Florinsky"):
        outs = xb.terrain.get_terrain_attribute(dem, req, resolution=5.0, surface_fit=fit, window_size=5)
        refs = to.get_terrain_attribute(dem, req, resolution=5.0, surface_fit=fit, window_size=5)
        keep = to.get_terrain_attribute(dem.astype(np.float64), "slope", resolution=5.0, surface_fit=fit) > 1e-3
        for a, o, r in zip(req, outs, refs):
            if a == "topographic_position_index":
                assert parity.nanmask_equal(o, r)
                assert np.array_equal(o, r, equal_nan=True)
            else:
                parity.assert_attr_close(o, r, a, where=keep if a == "aspect" else None, msg=f"{shape} {fit}")


def test_torch_cuda_tensor_in_out(xb) -> None:
    import torch

    from oracle import synth

    dem = synth.fractal_dem((200, 260), seed=5)
    t = torch.from_numpy(dem).cuda()
    s_t = xb.terrain.slope(t, resolution=5.0)
    assert isinstance(s_t, torch.Tensor) and s_t.is_cuda and s_t.dtype == torch.float32
    s_n = xb.terrain.slope(dem, resolution=5.0)
    assert np.array_equal(s_t.cpu().numpy(), s_n, equal_nan=True)
    # non-contiguous view / unaligned base -> cooperative loader path gives identical results
    big = torch.full((204, 271), float("nan"), device="cuda")
    big[2:202, 5:265] = t
    s_v = xb.terrain.slope(big[2:202, 5:265], resolution=5.0)
    assert np.array_equal(s_v.cpu().numpy(), s_n, equal_nan=True)


def test_large_tiled_equals_untiled(xb) -> None:
    """Size-independent property at a large size (test_terrain.py:295-341 analogue): computing row blocks with halo rows
    reproduces the single-launch result bit-for-bit; integer DEM windowed indexes stay bit-exact vs the oracle on a crop."""
    import torch

    from oracle import terrain_oracle as to
    from xdem_b200 import _engine

    g = torch.Generator(device="cuda").manual_seed(11)
    H, W = 4096, 4096
    z = (1000 + 0.05 * torch.cumsum(torch.cumsum(torch.randn((H, W), generator=g, device="cuda"), 0), 1)).float()
    z[1000:1010, 2000:2020] = float("nan")
    surf = ["slope", "aspect", "hillshade", "curvature", "max_curvature"]
    win = ["topographic_position_index", "terrain_ruggedness_index", "roughness"]
    full = _engine.terrain_fused(z, 5.0, surf, win, surface_fit="Florinsky", window_size=5, degrees=True,
                                 clip_hillshade=True)
    depth = 2
    pieces = []
    for r0 in range(0, H, 1000):
        r1 = min(H, r0 + 1000)
        b0, b1 = max(0, r0 - depth), min(H, r1 + depth)
        pieces.append(_engine.terrain_fused(z[b0:b1], 5.0, surf, win, surface_fit="Florinsky", window_size=5,
                                            degrees=True, clip_hillshade=True, row_begin=r0 - b0, row_end=r1 - b0))
    tiled = torch.cat(pieces, dim=1)
    # interior block edges must match exactly; the buffer ends act as raster borders only at the true borders
    assert torch.equal(torch.isnan(full), torch.isnan(tiled))
    assert torch.equal(torch.nan_to_num(full), torch.nan_to_num(tiled))
    crop = z[990:1100, 1990:2120].cpu().numpy()
    o = full[:, 990:1100, 1990:2120].cpu().numpy()[:, 2:-2, 2:-2]
    r = to.get_terrain_attribute(crop, surf + win, resolution=5.0, surface_fit="Florinsky", window_size=5)
    for i, a in enumerate(surf + win):
        rr = r[i][2:-2, 2:-2]
        if a == "topographic_position_index":
            assert np.nanmax(np.abs(o[i] - rr)) < 1e-3
        else:
            parity.assert_attr_close(o[i], rr, a, msg="crop")


def test_math_accuracy_sweep(xb) -> None:
    """The branch-free fp32 atan / atan2 / sqrt / rsqrt cores (csrc/xb_math.cuh): sweep every gradient direction and
    eleven decades of gradient magnitude on planar facets and compare with the float64 oracle at a few fp32 ulp."""
    from oracle import terrain_oracle as to

    n_ang, n_mag, B = 96, 48, 6
    H, W = n_mag * B, n_ang * B
    yy, xx = np.mgrid[0:H, 0:W]
    ang = (xx // B) * (2 * np.pi / n_ang) + 0.0123
    mag = 10.0 ** (-5 + 8 * (yy // B) / (n_mag - 1))
    # planar facet per block, anchored at the block centre so that values stay fp32-friendly
    cx, cy = (xx % B) - B / 2, (yy % B) - B / 2
    dem = (100.0 + mag * (np.cos(ang) * cx + np.sin(ang) * cy)).astype(np.float32)
    for fit in ("Horn", "ZevenbergThorne"):
        out = xb.terrain.get_terrain_attribute(dem, ["slope", "aspect", "hillshade"], resolution=1.0, surface_fit=fit,
                                               hillshade_z_factor=2.0)
        ref = to.get_terrain_attribute(dem.astype(np.float64), ["slope", "aspect", "hillshade"], resolution=1.0,
                                       surface_fit=fit, hillshade_z_factor=2.0)
        for a, o, r in zip(["slope", "aspect", "hillshade"], out, ref):
            m = np.isfinite(r) & (ref[0] > 1e-30)
            d = np.abs(o[m].astype(np.float64) - r[m])
            if a == "aspect":
                d = np.minimum(d, 360.0 - d)
                rel = d / 360.0
            elif a == "hillshade":
                rel = d / 255.0  # values are clipped to [0, 255]; near 0 only the absolute error is meaningful
            else:
                rel = d / np.maximum(np.abs(r[m]), 1e-300)
            assert rel.max() < 6e-7, (fit, a, rel.max())


@pytest.mark.parametrize("name", ["fractal", "small_int"])
def test_generic_windows_and_fractal_vs_reference(xb, G, name: str) -> None:
    """window sizes 7 / 9 (generic odd-window kernel) and fractal roughness against the reference fixtures."""
    dem = G[f"in|{name}"]
    for w in (7, 9):
        for tm in ("Riley", "Wilson"):
            outs = xb.terrain.get_terrain_attribute(dem, WIN[:3], window_size=w, tri_method=tm)
            for a, o in zip(WIN[:3], outs):
                ref_s, ref_n = G[f"win|{name}|scipy|{w}|{tm}|{a}"], G[f"win|{name}|numba|{w}|{tm}|{a}"]
                assert o.dtype == np.float32 and parity.nanmask_equal(o, ref_s)
                if a == "roughness":
                    assert np.array_equal(o, ref_s, equal_nan=True)
                elif a == "terrain_ruggedness_index":
                    assert np.array_equal(o, ref_n, equal_nan=True)
                elif name == "small_int":
                    assert np.array_equal(o, ref_s, equal_nan=True)
                else:
                    assert np.nanmax(np.abs(o - ref_s)) <= 4 * np.spacing(np.float32(np.nanmax(np.abs(dem)) * w * w))
    for wf in (13, 7):
        o = xb.terrain.fractal_roughness(dem, window_size_fractal=wf) if wf == 13 else \
            xb.terrain.get_terrain_attribute(dem, "fractal_roughness", window_size_fractal=wf)
        for engine in ("scipy", "numba"):
            ref = G[f"frac|{name}|{engine}|{wf}"]
            assert parity.nanmask_equal(o, ref), (wf, engine)
            m = np.isfinite(ref)
            assert np.allclose(o[m], ref[m], rtol=1e-5, atol=2e-6)
    o64 = xb.terrain.fractal_roughness(G["in|fractal"].astype(np.float64))
    ref64 = G["frac64|fractal|scipy|13"]
    assert o64.dtype == np.float64 and parity.nanmask_equal(o64, ref64)
    assert np.allclose(o64[np.isfinite(ref64)], ref64[np.isfinite(ref64)], rtol=1e-12)


def test_fractal_roughness_known_answers(xb) -> None:
    """test_window.py:70-89."""
    for setter, expect in ((lambda d: d.__setitem__((1, 1), 6.5), 1.0), (lambda d: d.__setitem__((slice(None), 1), 13), 2.0),
                           (lambda d: d.__setitem__((slice(None), slice(None, 6)), 13), 3.0)):
        dem = np.zeros((13, 13), dtype="float64")
        setter(dem)
        assert np.round(xb.terrain.fractal_roughness(dem)[6, 6], 3) == np.float32(expect)


def test_three_cycle_request_keeps_requested_order(xb, G) -> None:
    """A request whose category permutation is a 3-cycle (windowed, frequency, surface): outs[i] is attribute[i].  The
    reference's re-ordering (terrain.py:648-656) applies the permutation instead of its inverse and would return
    (texture_shading, slope, roughness) here -- a documented, deliberate deviation (terrain.py docstring,
    INTEGRATION.md section 6)."""
    dem = G["in|fractal"]
    req = ["roughness", "texture_shading", "slope"]
    outs = xb.terrain.get_terrain_attribute(dem, req, resolution=5.0)
    assert np.array_equal(outs[0], xb.terrain.roughness(dem), equal_nan=True)
    assert np.array_equal(outs[1], xb.terrain.texture_shading(dem), equal_nan=True)
    assert np.array_equal(outs[2], xb.terrain.slope(dem, resolution=5.0), equal_nan=True)
    # what the reference's index expression yields for this request (pure list logic, no reference import needed)
    grouped = ["slope", "roughness", "texture_shading"]  # surface + windowed + frequency (terrain.py:647)
    ref_order = [grouped[req.index(a)] for a in grouped]
    assert ref_order == ["texture_shading", "slope", "roughness"] and ref_order != req


def test_mixed_request_all_paths(xb, G) -> None:
    """One call mixing the fused kernel, the 3x3 rugosity special case, a generic window and fractal roughness keeps the
    request order and equals the single-attribute calls (terrain.py:651-658)."""
    dem = G["in|fractal"]
    req = ["fractal_roughness", "slope", "rugosity", "roughness", "topographic_position_index", "max_curvature"]
    outs = xb.terrain.get_terrain_attribute(dem, req, resolution=5.0, window_size=7)
    assert np.array_equal(outs[0], xb.terrain.fractal_roughness(dem), equal_nan=True)
    assert np.array_equal(outs[1], xb.terrain.slope(dem, resolution=5.0), equal_nan=True)
    assert np.array_equal(outs[2], xb.terrain.rugosity(dem, resolution=5.0), equal_nan=True)
    assert np.array_equal(outs[3], xb.terrain.roughness(dem, window_size=7), equal_nan=True)
    assert np.array_equal(outs[4], xb.terrain.topographic_position_index(dem, window_size=7), equal_nan=True)
    assert np.array_equal(outs[5], xb.terrain.max_curvature(dem, resolution=5.0), equal_nan=True)
    from xdem_b200.window import _get_windowed_indexes

    seam = _get_windowed_indexes(dem, 7, ["roughness", "fractal_roughness"], 5.0)
    assert seam.shape == (2,) + dem.shape
    assert np.array_equal(seam[0], outs[3], equal_nan=True)


@pytest.mark.parametrize("shape", [(64, 128), (61, 132), (300, 1028), (7, 8), (1200, 516)])
def test_florinsky_sliding_kernel_vs_generic_and_oracle(xb, shape) -> None:
    """The row-feature-reuse Florinsky kernel (xb_terrain_fl.cu, taken for 16-byte aligned float32 rasters when a
    second-derivative attribute is requested) against the generic fused kernel and the float64 oracle; integer-valued
    DEMs must agree bit-for-bit (every stencil sum is exact in both)."""
    import torch

    from oracle import synth
    from oracle import terrain_oracle as to
    from xdem_b200 import _engine, _lib

    attrs = ["slope", "aspect", "hillshade", "curvature", "profile_curvature", "planform_curvature", "max_curvature",
             "min_curvature"]
    dem = synth.inject_nans(synth.fractal_dem(shape, seed=21), frac=0.003, hole=3) if shape[0] > 10 else \
        synth.fractal_dem(shape, seed=21)
    idem = synth.integer_dem(shape, seed=22, high=500)
    for d, exact in ((dem, False), (idem, True)):
        t = torch.from_numpy(d).cuda()
        kw = dict(surface_fit="Florinsky", degrees=True, clip_hillshade=True)
        _lib.set_option("florinsky_generic", 0)
        n0 = _lib.launch_count()
        a = _engine.terrain_fused(t, 5.0, attrs, **kw).cpu().numpy()
        _lib.set_option("florinsky_generic", 1)
        try:
            b = _engine.terrain_fused(t, 5.0, attrs, **kw).cpu().numpy()
        finally:
            _lib.set_option("florinsky_generic", 0)
        assert _lib.launch_count() == n0 + 2
        assert np.array_equal(np.isnan(a), np.isnan(b))
        keep = to.get_terrain_attribute(d.astype(np.float64), "slope", resolution=5.0) > 1e-3
        for i, name in enumerate(attrs):
            if exact and name in ("slope", "hillshade", "curvature"):
                assert np.array_equal(a[i], b[i], equal_nan=True), name
            parity.assert_attr_close(a[i], b[i], name, where=keep, atol_scale=20.0 if exact else 1.0,
                                     msg=f"sliding vs generic {shape}")
        if not exact:
            ref = to.get_terrain_attribute(d, attrs, resolution=5.0)
            for i, name in enumerate(attrs):
                parity.assert_attr_close(a[i], ref[i], name, where=keep, msg=f"sliding vs oracle {shape}")


def test_exact_math_cores(xb) -> None:
    """The branch-free IEEE cores of the 3x3 windowed kernel (xb_terrain_w3.cu) against the CUDA round-to-nearest
    intrinsics, bit for bit: the fast-path square root over EVERY float32 of its range [2^-101, FLT_MAX], and the
    reciprocal-multiply division by L^2 for the resolutions used in the tests over every positive normal float32
    whose quotient stays normal."""
    import ctypes

    import torch

    from xdem_b200 import _lib

    L = _lib.lib()
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    lo, hi = 0x0D000000, 0x7F7FFFFF
    _lib.check(L.xb_probe_exact_math(0, lo, hi - lo + 1, 1.0, 1.0, bad.data_ptr(), None))
    torch.cuda.synchronize()
    assert int(bad.item()) == 0, f"fast sqrt differs from __fsqrt_rn on {int(bad.item())} inputs"
    for res in (5.0, 1.0, 30.0, 0.5, 2.5, 0.1, 100.0, 3.0, 1e-3, 12345.678):
        ll = np.float32(res * res)
        cands = [np.nextafter(np.float32(1.0 / np.float64(ll)), np.float32(d)) for d in (0.0, np.inf)]
        cands.append(np.float32(1.0 / np.float64(ll)))
        y = min(cands, key=lambda c: abs(np.float64(ll) * np.float64(c) - 1.0))
        bad.zero_()
        # numerators 2^-40 .. 2^60: every bit pattern in between (areas are ~L^2)
        lo, hi = 0x2B800000, 0x5D800000
        _lib.check(L.xb_probe_exact_math(1, lo, hi - lo + 1, ctypes.c_float(float(ll)), ctypes.c_float(float(y)),
                                         bad.data_ptr(), None))
        torch.cuda.synchronize()
        assert int(bad.item()) == 0, f"division by {ll} differs from __fdiv_rn on {int(bad.item())} inputs"


@pytest.mark.parametrize("res", [5.0, 1.0, 30.0, 0.37, 1e-7])
@pytest.mark.parametrize("tm", ["Riley", "Wilson"])
def test_window3_sliding_equals_generic(xb, res: float, tm: str) -> None:
    """The row-feature-reuse 3x3 kernel (packed f32x2, shared Jenness segments, fast IEEE sqrt / division) is
    bit-identical to the generic fused kernel on rough data with NaN / inf cells, ragged widths and for every subset
    mask; a resolution outside the fast cores' range (1e-7) must fall back to the generic kernel by itself."""
    import torch

    from xdem_b200 import _engine, _lib

    g = torch.Generator(device="cuda").manual_seed(5)
    H, W = 701, 1028  # width multiple of 4 (TMA eligible) but not of the tile; ragged rows
    z = 1000 + 30 * torch.randn((H, W), generator=g, device="cuda")
    z[::97, ::53] = float("nan")
    z[300, 400] = float("inf")
    z[500:520, 600:640] = 123.0  # flat patch: TRI = 0 exactly
    for attrs in (WIN, WIN[:3], ["rugosity"], ["roughness", "rugosity"], ["terrain_ruggedness_index"]):
        fast = _engine.terrain_fused(z, res, windowed_indexes=attrs, tri_method=tm)
        _lib.set_option("window3_generic", 1)
        try:
            ref = _engine.terrain_fused(z, res, windowed_indexes=attrs, tri_method=tm)
        finally:
            _lib.set_option("window3_generic", 0)
        assert torch.equal(torch.isnan(fast), torch.isnan(ref)), attrs
        assert torch.equal(fast.view(torch.int32), ref.view(torch.int32)), (attrs, res, tm)


@pytest.mark.parametrize("shape", [(61, 132), (777, 1028), (123, 64)])
def test_florinsky_tma_store_variant_equals_default(xb, shape) -> None:
    """The TMA-bulk-store variant of the packed Florinsky kernel (option ``florinsky_tma_store``, kept for A/B: measured
    12 % slower than st.global.cs, profiles/ab_tstore_r02.txt) writes the same bits, including the NaN rim, ragged
    tiles clipped by the tensor maps and row-sliced (shard-like) requests."""
    import torch

    from xdem_b200 import _engine, _lib

    H, W = shape
    g = torch.Generator(device="cuda").manual_seed(9)
    z = (1000 + torch.cumsum(torch.randn((H, W), generator=g, device="cuda"), 1)).float()
    z[H // 2, W // 3] = float("nan")
    for attrs in (["slope", "aspect", "hillshade", "curvature"], ["slope", "aspect", "curvature"]):
        for rb, re in ((0, H), (7, H - 9)):
            kw = dict(surface_fit="Florinsky", degrees=True, clip_hillshade=True, row_begin=rb, row_end=re)
            ref = _engine.terrain_fused(z, 5.0, attrs, [], **kw)
            _lib.set_option("florinsky_tma_store", 1)
            try:
                got = _engine.terrain_fused(z, 5.0, attrs, [], **kw)
            finally:
                _lib.set_option("florinsky_tma_store", 0)
            assert torch.equal(got.view(torch.int32), ref.view(torch.int32)), (attrs, rb, re)


def test_all13_split_equals_separate(xb, G) -> None:
    """BASELINE config 4's request (9 surface attributes + 4 windowed indexes, Florinsky) runs as two specialised
    launches; the planes must equal the ones of separate requests bit for bit."""
    dem = G["in|fractal"]
    surf = [a for a in SURF if a != "curvature"]
    outs = xb.terrain.get_terrain_attribute(dem, surf + WIN, resolution=5.0)
    so = xb.terrain.get_terrain_attribute(dem, surf, resolution=5.0)
    wo = xb.terrain.get_terrain_attribute(dem, WIN, resolution=5.0)
    for a, o, r in zip(surf + WIN, outs, list(so) + list(wo)):
        assert np.array_equal(o, r, equal_nan=True), a
