"""CPU: ``xdem_b200.install()`` against the UNMODIFIED reference modules (loaded where they lie under /root/reference
through oracle/refload.py; skipped on the GPU box where the reference does not exist).  The compute entry points are
replaced by recorders, so the test checks what ARRIVES at the seams when the reference's own public functions are
called -- terrain.py:574-583, 606-614; affine.py:2509-2522; spatialstats.py:1054, 1278 -- and that the reference's
own glue still runs around them."""

from __future__ import annotations

import numpy as np
import pytest

from oracle import refload

pytestmark = pytest.mark.skipif(not refload.reference_available(), reason="needs /root/reference")


@pytest.fixture()
def ref():
    return refload.load_reference()


def test_install_routes_reference_terrain_calls_through_the_streamed_seam(ref, monkeypatch) -> None:
    import xdem_b200
    from xdem_b200 import _engine

    calls = []

    def fake_host(dem, resolution, surface_attributes=(), windowed_indexes=(), out=None, **kw):
        calls.append(dict(dem=dem, resolution=resolution, surf=list(surface_attributes), win=list(windowed_indexes),
                          out=out, kw=kw))
        n = len(surface_attributes) + len(windowed_indexes)
        if out is None:
            return np.zeros((n,) + dem.shape, dtype=dem.dtype)
        for o in out:
            o[...] = 1.0
        return out

    monkeypatch.setattr(_engine, "terrain_fused_host", fake_host)
    monkeypatch.setattr(_engine, "host_planes", lambda n, r, c, dt: np.empty((n, r, c), dtype=dt))
    saved = (ref.terrain._get_surface_attributes, ref.terrain._get_windowed_indexes)
    try:
        xdem_b200.install()
        dem = np.arange(48, dtype=np.int32).reshape(6, 8)
        out = ref.terrain.get_terrain_attribute(dem, ["hillshade", "roughness", "slope", "rugosity"], resolution=5.0,
                                                hillshade_azimuth=200.0, hillshade_z_factor=2.0, window_size=5)
    finally:
        xdem_b200.uninstall()
    assert (ref.terrain._get_surface_attributes, ref.terrain._get_windowed_indexes) == saved
    assert len(out) == 4 and all(o.shape == (6, 8) and o.dtype == np.float32 for o in out)
    surf = [c for c in calls if c["surf"]]
    win = [c for c in calls if c["win"]]
    # one streamed call for the surface-fit group, in the reference's canonical order, radians / unclipped at the seam
    assert len(surf) == 1 and surf[0]["surf"] == ["hillshade", "slope"]
    assert surf[0]["dem"].dtype == np.float32 and surf[0]["resolution"] == 5.0  # ints are cast (terrain.py:560-561)
    kw = surf[0]["kw"]
    assert kw["surface_fit"] == "Florinsky" and kw["curv_method"] == "geometric"
    assert kw["degrees"] is False and kw["clip_hillshade"] is False
    assert kw["hillshade_azimuth"] == 200.0 and kw["hillshade_altitude"] == 45.0 and kw["hillshade_z_factor"] == 2.0
    # windowed group: roughness on the 5x5 window; rugosity always on 3x3 (window.py:909-914) -> its own streamed pass,
    # both written straight into planes of the one result array (no stack copy)
    assert sorted(tuple(c["win"]) for c in win) == [("roughness",), ("rugosity",)]
    assert {c["kw"]["window_size"] for c in win} == {5, 3}
    assert all(isinstance(c["out"], list) and c["out"][0].shape == (6, 8) for c in win)


def test_install_rebinds_nuth_kaab_and_variogram_pair_functions(ref, monkeypatch) -> None:
    import sys

    import xdem_b200
    from xdem_b200 import coreg
    from xdem_b200 import spatialstats as xs

    saved = (ref.terrain._get_surface_attributes, ref.terrain._get_windowed_indexes, ref.coreg_affine.nuth_kaab,
             ref.spatialstats._get_pdist_empirical_variogram, ref.spatialstats._get_cdist_empirical_variogram)
    seen = {}
    monkeypatch.setattr(coreg, "nuth_kaab", lambda **kw: seen.setdefault("nk", kw) and ((1.0, 2.0, 3.0), 7))
    try:
        xdem_b200.install()
        assert sys.modules["xdem.coreg.affine"].nuth_kaab is not saved[2]
        assert ref.spatialstats._get_pdist_empirical_variogram is xs._get_pdist_empirical_variogram
        assert ref.spatialstats._get_cdist_empirical_variogram is xs._get_cdist_empirical_variogram
        hook = ref.coreg_affine.nuth_kaab
        a = np.zeros((4, 4), dtype=np.float32)
        common = dict(inlier_mask=np.ones((4, 4), bool), transform=None, crs=None, area_or_point=None, tolerance=0.1,
                      max_iterations=3, params_random={"subsample": 1.0, "random_state": None}, z_name="z")
        # raster-raster with the reference defaults -> GPU function, same keywords as affine.py:2509-2522 passes
        r = hook(ref_elev=a, tba_elev=a, params_fit_or_bin={"fit_or_bin": "bin_and_fit", "bin_sizes": 72,
                                                            "bin_statistic": np.nanmedian}, **common)
        assert r == ((1.0, 2.0, 3.0), 7) and seen["nk"]["max_iterations"] == 3 and seen["nk"]["z_name"] == "z"
        # anything off the B200 path keeps the reference function (here: a custom statistic)
        called = {}
        monkeypatch.setattr(hook, "__wrapped__", lambda **kw: called.setdefault("ref", True) and "reference")
        hook2 = coreg.make_reference_hook(lambda **kw: "reference")
        assert hook2(ref_elev=a, tba_elev=a, params_fit_or_bin={"fit_or_bin": "bin_and_fit", "bin_sizes": 72,
                                                                "bin_statistic": np.nanmean}, **common) == "reference"
        assert hook2(ref_elev=object(), tba_elev=a, params_fit_or_bin={}, **common) == "reference"
    finally:
        xdem_b200.uninstall()
    assert (ref.terrain._get_surface_attributes, ref.terrain._get_windowed_indexes, ref.coreg_affine.nuth_kaab,
            ref.spatialstats._get_pdist_empirical_variogram, ref.spatialstats._get_cdist_empirical_variogram) == saved


def test_equidistant_parameter_split_equals_reference(ref) -> None:
    """`_choose_cdist_equidistant_sampling_parameters` is the reference's own glue (spatialstats.py:1104-1183)."""
    from xdem_b200 import spatialstats as xs

    for shape, sub in (((100, 100), 1000), ((517, 310), 250), ((2000, 3000), 10000), ((50, 60), 12)):
        ext = (0.0, (shape[0] - 1) * 5.0, 0.0, (shape[1] - 1) * 5.0)
        want = ref.spatialstats._choose_cdist_equidistant_sampling_parameters(extent=ext, shape=shape, subsample=sub)
        got = xs._choose_cdist_equidistant_sampling_parameters(extent=ext, shape=shape, subsample=sub)
        assert got == want


def test_uninstall_restores_the_reference_functions(ref) -> None:
    """install() twice then uninstall(): every rebound name is the reference's own object again (install is idempotent:
    the nuth_kaab hook always wraps the ORIGINAL function, never an earlier hook)."""
    import xdem_b200

    names = [(ref.terrain, "_get_surface_attributes"), (ref.terrain, "_get_windowed_indexes"),
             (ref.coreg_affine, "nuth_kaab"), (ref.spatialstats, "_get_pdist_empirical_variogram"),
             (ref.spatialstats, "_get_cdist_empirical_variogram")]
    originals = [getattr(m, n) for m, n in names]
    try:
        xdem_b200.install()
        first_hook = ref.coreg_affine.nuth_kaab
        xdem_b200.install()
        assert ref.coreg_affine.nuth_kaab.__wrapped__ is originals[2] and first_hook.__wrapped__ is originals[2]
        assert all(getattr(m, n) is not o for (m, n), o in zip(names, originals))
    finally:
        xdem_b200.uninstall()
    assert all(getattr(m, n) is o for (m, n), o in zip(names, originals))
    xdem_b200.uninstall()  # a second call is a no-op
