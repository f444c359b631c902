"""CPU: the NumPy oracle (oracle/terrain_oracle.py) is pinned against fixtures produced by the unmodified reference
(both engines, tests/golden/terrain_reference.npz via oracle/make_golden.py) and against the reference's own
known-answer tests."""

from __future__ import annotations

import numpy as np
import pytest

from oracle import terrain_oracle as to
from tests import parity

G = parity.load_golden()
SURF = to.SURFACE_ATTRS
WIN = to.WINDOW_ATTRS


def _flat_mask(dem: np.ndarray, fit: str) -> np.ndarray:
    """pixels whose gradient is not ~0 (aspect / curvature guards are noise-driven on flats, SURVEY.md hard part 4)."""
    s = to.get_terrain_attribute(dem.astype(np.float64), "slope", resolution=5.0, surface_fit=fit, degrees=True)
    return s > 1e-3


@pytest.mark.parametrize("name", ["fractal", "noise", "integer"])
@pytest.mark.parametrize("engine", ["scipy", "numba"])
@pytest.mark.parametrize("fit", ["Horn", "ZevenbergThorne", "Florinsky"])
@pytest.mark.parametrize("cm", ["geometric", "directional"])
def test_oracle_surface_vs_reference(name: str, engine: str, fit: str, cm: str) -> None:
    if fit == "Horn" and cm == "directional":
        pytest.skip("Horn has no curvatures")
    dem = G[f"in|{name}"]
    attrs = SURF[:3] if fit == "Horn" else SURF
    # scipy.ndimage.convolve rounds each coefficient to the float32 input dtype
    cr = np.float32 if engine == "scipy" else None
    outs = to.get_terrain_attribute(dem, attrs, resolution=5.0, surface_fit=fit, curv_method=cm, coef_round=cr)
    keep = _flat_mask(dem, fit)
    for a, o in zip(attrs, outs):
        ref = G[f"surf|{name}|{engine}|{fit}|{cm}|deg|{a}"]
        # strict criterion (no widening); the flat-pixel mask applies to aspect only (test_terrain.py:168)
        parity.assert_attr_close(o, ref, a, where=keep if a == "aspect" else None, msg=f"{name}/{engine}/{fit}/{cm}")


@pytest.mark.parametrize("name", ["fractal", "integer", "small_int"])
@pytest.mark.parametrize("w", [3, 5])
@pytest.mark.parametrize("tm", ["Riley", "Wilson"])
def test_oracle_windowed_vs_reference(name: str, w: int, tm: str) -> None:
    dem = G[f"in|{name}"]
    attrs = WIN if w == 3 else WIN[:3]
    outs = to.get_terrain_attribute(dem, attrs, resolution=5.0, window_size=w, tri_method=tm)
    for a, o in zip(attrs, outs):
        ref_n = G[f"win|{name}|numba|{w}|{tm}|{a}"]
        ref_s = G[f"win|{name}|scipy|{w}|{tm}|{a}"]
        assert parity.nanmask_equal(o, ref_n) and parity.nanmask_equal(o, ref_s)
        if a == "roughness":
            assert np.array_equal(o, ref_s, equal_nan=True) and np.array_equal(o, ref_n, equal_nan=True)
        elif a == "rugosity":
            assert np.array_equal(o, ref_s, equal_nan=True)  # same float32 op order as the SciPy engine
            if name == "fractal":
                # (on the random integer DEMs relief/resolution ~ 500: float32 Heron areas of needle triangles make the
                # reference's own two engines disagree by 0.4 %, so only the SciPy engine is pinned there)
                parity.assert_attr_close(o, ref_n, a, msg="rugosity vs numba")
        elif a == "topographic_position_index":
            if name != "fractal":
                # integer-valued DEM: bit-exact against the SciPy engine (sums are exact)
                assert np.array_equal(o, ref_s, equal_nan=True)
            if w == 3:
                assert np.array_equal(o, ref_n, equal_nan=True)  # sequential float32 sum, /8 exact
            # float DEM: both engines are within a few float32 ulp of |z| of each other and of the oracle
            assert np.nanmax(np.abs(o - ref_s)) <= 4 * np.spacing(np.float32(np.nanmax(np.abs(dem)) * w * w))
        else:  # TRI
            assert np.array_equal(o, ref_n, equal_nan=True)  # Numba engine: same sequential float32 order
            if name == "small_int" or tm == "Wilson":
                assert np.array_equal(o, ref_s, equal_nan=True)
            else:
                parity.assert_attr_close(o, ref_s, a, msg="TRI vs scipy")


def test_oracle_float64_truth_vs_reference() -> None:
    dem = G["in|fractal64"]
    for fit in ("ZevenbergThorne", "Florinsky"):
        outs = to.get_terrain_attribute(dem, SURF, resolution=5.0, surface_fit=fit)
        for a, o in zip(SURF, outs):
            ref = G[f"surf|fractal64|numba|{fit}|geometric|deg|{a}"]
            assert o.dtype == np.float64
            parity.assert_attr_close(o, ref, a, rtol=1e-9, atol_scale=1e-4, msg=f"f64 {fit}")


def test_reference_doctests() -> None:
    """terrain.py:268-279, 799-813, 1484-1493, 1553-1562."""
    dem = np.repeat(np.arange(3), 3)[::-1].reshape(3, 3)
    s, a = to.get_terrain_attribute(dem, ["slope", "aspect"], resolution=1, surface_fit="ZevenbergThorne")
    assert s[1, 1] == np.float32(45.0) and a[1, 1] == np.float32(180.0)
    dem2 = np.tile(np.arange(3), (3, 1))
    a2 = to.get_terrain_attribute(dem2, "aspect", resolution=1.0, surface_fit="ZevenbergThorne")
    assert a2[1, 1] == np.float32(270.0)
    d3 = np.zeros((3, 3), dtype="int32")
    d3[1, 1] = 1
    assert to.get_terrain_attribute(d3, "topographic_position_index")[1, 1] == np.float32(1.0)
    tri = to.get_terrain_attribute(d3, "terrain_ruggedness_index")[1, 1]
    assert tri == np.float32(2.828427)


def test_rugosity_jenness() -> None:
    """test_window.py:21-36."""
    dem = np.array([[190, 170, 155], [183, 165, 145], [175, 160, 122]], dtype="float32")
    r = to.get_terrain_attribute(dem, "rugosity", resolution=100.0)
    assert r[1, 1] == pytest.approx(10280.48 / 10000.0, rel=1e-4)


@pytest.mark.parametrize("dh", np.linspace(0.01, 100, 3))
@pytest.mark.parametrize("resolution", np.linspace(0.01, 100, 3))
def test_rugosity_simple_cases(dh: float, resolution: float) -> None:
    """test_window.py:38-68."""
    dem = np.array([[1, 1, 1], [1, 1 + dh, 1], [1, 1, 1]], dtype="float64")
    r = to.get_terrain_attribute(dem, "rugosity", resolution=resolution)
    side1 = np.sqrt(2 * resolution**2 + dh**2) / 2.0
    side2 = np.sqrt(resolution**2 + dh**2) / 2.0
    side3 = resolution / 2.0
    s = (side1 + side2 + side3) / 2.0
    A = np.sqrt(s * (s - side1) * (s - side2) * (s - side3))
    assert r[1, 1] == pytest.approx(8 * A / resolution**2, rel=1e-6)


@pytest.mark.needs_reference
def test_oracle_vs_live_reference_random() -> None:
    """Extra pin in the build container: a fresh random DEM through the live reference (numba engine)."""
    import warnings

    from oracle import synth
    from oracle.refload import load_reference

    warnings.filterwarnings("ignore")
    ref = load_reference()
    dem = synth.inject_nans(synth.fractal_dem((33, 47), seed=123))
    r = ref.terrain.get_terrain_attribute(dem, SURF, resolution=2.0, surface_fit="Florinsky", engine="numba")
    o = to.get_terrain_attribute(dem, SURF, resolution=2.0, surface_fit="Florinsky")
    for a, oo, rr in zip(SURF, o, r):
        parity.assert_attr_close(oo, rr, a, msg="live")


@pytest.mark.parametrize("name", ["fractal", "noise", "integer"])
@pytest.mark.parametrize("fit", ["Horn", "ZevenbergThorne", "Florinsky"])
def test_c_oracle_bit_exact_vs_numba_engine(name: str, fit: str) -> None:
    """The plain-C restatement (oracle/terrain_oracle.c, the CPU-baseline arm) reproduces the reference's Numba engine
    bit-for-bit on the committed fixtures."""
    from oracle import c_oracle as co

    dem = G[f"in|{name}"]
    attrs = SURF[:3] if fit == "Horn" else SURF
    for cm in ("geometric", "directional"):
        if fit == "Horn" and cm == "directional":
            continue
        o = co.surface_attributes(dem, 5.0, attrs, fit, curv_method=cm, degrees=True, clip_hillshade=True)
        for i, a in enumerate(attrs):
            ref = G[f"surf|{name}|numba|{fit}|{cm}|deg|{a}"]
            assert np.array_equal(o[i], ref, equal_nan=True), (name, fit, cm, a)
    for w in (3, 5):
        for tm in ("Riley", "Wilson"):
            ww = co.windowed_indexes(dem, w, WIN[:3], tri_method=tm)
            for i, a in enumerate(WIN[:3]):
                assert np.array_equal(ww[i], G[f"win|{name}|numba|{w}|{tm}|{a}"], equal_nan=True), (name, w, tm, a)


@pytest.mark.parametrize("name", ["fractal", "small_int"])
def test_oracle_generic_windows_and_fractal_vs_reference(name: str) -> None:
    """window sizes 7 / 9 and fractal roughness (13, 7) against both reference engines."""
    dem = G[f"in|{name}"]
    for w in (7, 9):
        for tm in ("Riley", "Wilson"):
            outs = to.get_terrain_attribute(dem, WIN[:3], window_size=w, tri_method=tm)
            for a, o in zip(WIN[:3], outs):
                ref_s, ref_n = G[f"win|{name}|scipy|{w}|{tm}|{a}"], G[f"win|{name}|numba|{w}|{tm}|{a}"]
                assert parity.nanmask_equal(o, ref_s) and parity.nanmask_equal(o, ref_n)
                if a == "roughness":
                    assert np.array_equal(o, ref_s, equal_nan=True)
                elif a == "terrain_ruggedness_index":
                    assert np.array_equal(o, ref_n, equal_nan=True)
                elif name == "small_int":
                    assert np.array_equal(o, ref_s, equal_nan=True)  # TPI on integer-valued DEM: exact
    for wf in (13, 7):
        o = to.get_terrain_attribute(dem, "fractal_roughness", window_size_fractal=wf)
        for engine in ("scipy", "numba"):
            ref = G[f"frac|{name}|{engine}|{wf}"]
            assert parity.nanmask_equal(o, ref)
            m = np.isfinite(ref)
            assert np.allclose(o[m], ref[m], rtol=1e-5, atol=2e-6)


def test_fractal_roughness_known_answers() -> None:
    """test_window.py:70-89: line -> 1, plane -> 2, cube -> 3."""
    for setter, expect in ((lambda d: d.__setitem__((1, 1), 6.5), 1.0), (lambda d: d.__setitem__((slice(None), 1), 13), 2.0),
                           (lambda d: d.__setitem__((slice(None), slice(None, 6)), 13), 3.0)):
        dem = np.zeros((13, 13), dtype="float64")
        setter(dem)
        fr = to.get_terrain_attribute(dem, "fractal_roughness")
        assert np.round(fr[6, 6], 3) == np.float32(expect)


def test_c_rugosity_baseline_arm_close_to_numba_fixture() -> None:
    """oracle/terrain_oracle.c:xo_rugosity_f32 is only the CPU arm of the "all attributes" benchmark (the NumPy oracle
    pins rugosity bit-for-bit to the SciPy engine above); it restates the Numba engine's per-pixel function and must
    stay within float32 round-off of that engine's fixture on the smooth DEM."""
    from oracle import c_oracle

    dem = G["in|fractal"]
    o = c_oracle.rugosity(dem, 5.0)
    ref = G["win|fractal|numba|3|Riley|rugosity"]
    assert parity.nanmask_equal(o, ref)
    m = np.isfinite(ref)
    assert np.max(np.abs(o[m] - ref[m]) / np.abs(ref[m])) < 1e-6
