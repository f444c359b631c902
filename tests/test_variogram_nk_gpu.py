"""GPU parity tests of the variogram (K2) and Nuth-Kaab (K3) kernels through the public API -> C ABI."""

from __future__ import annotations

import numpy as np
import pytest

from tests import parity

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------------------------------------- variogram


def _sample(shape, n, seed, nan_frac=0.0):
    rng = np.random.default_rng(seed)
    vals = rng.normal(size=shape).astype(np.float32)
    if nan_frac:
        vals.ravel()[rng.choice(vals.size, int(nan_frac * vals.size), replace=False)] = np.nan
    finite = np.flatnonzero(np.isfinite(vals.ravel()))
    idx = rng.choice(finite, n, replace=False)
    return vals, idx


@pytest.mark.parametrize("shape,n,gsd", [((300, 300), 3000, 5.0), ((257, 257), 1000, 1.0), ((90, 90), 129, 30.0),
                                         ((64, 64), 128, 0.5), ((40, 40), 5, 2.0)])
@pytest.mark.parametrize("bins", ["even", "default"])
def test_pairwise_binning_vs_oracle(shape, n, gsd, bins) -> None:
    """Counts bit-exact, Matheron sums within 1e-5 relative of the float64 oracle (restated scikit-gstat)."""
    import torch

    from oracle import variogram_oracle as vo
    from xdem_b200 import spatialstats as xs

    vals, idx = _sample(shape, n, 44)
    nx, ny = shape
    coords = vo.grid_coords(shape, gsd)  # reference glue: coords[k] for flat sample k
    maxlag = float(np.hypot((nx - 1) * gsd, (ny - 1) * gsd))
    flat = vals.ravel()
    if bins == "even":
        b_o, exp_o, cnt_o = vo.empirical_variogram(coords[idx], flat[idx], "even", n_lags=50, maxlag=maxlag)
        edges_in = None
    else:
        edges_in = np.asarray(vo.default_bins(gsd, maxlag))
        b_o, exp_o, cnt_o = vo.empirical_variogram(coords[idx], flat[idx], edges_in)
    ti = torch.from_numpy(idx).cuda()
    x, y = ti % nx, ti // nx
    v = torch.from_numpy(flat).cuda()[ti]
    edges, cnt, ssq = xs.pairwise_lag_binning(x, y, v, edges_in, gsd, n_lags=50, maxlag=maxlag)
    assert np.array_equal(edges, b_o)
    assert np.array_equal(cnt, cnt_o), (cnt - cnt_o)
    with np.errstate(all="ignore"):
        exp = np.where(cnt > 0, ssq / (2.0 * cnt), np.nan)
    assert np.allclose(exp, exp_o, rtol=1e-5, equal_nan=True)


def test_sample_empirical_variogram_frame() -> None:
    """Public API: frame layout of spatialstats.py:1512-1546 (last bin dropped, dtypes) + oracle values."""
    from oracle import variogram_oracle as vo
    from xdem_b200 import spatialstats as xs

    vals, _ = _sample((200, 200), 10, 3, nan_frac=0.01)
    df = xs.sample_empirical_variogram(vals, gsd=5.0, subsample=1500, subsample_method="pdist_point", random_state=7)
    assert list(df.columns) == ["exp", "lags", "count", "err_exp"]
    assert df["count"].dtype == np.int64 and df["exp"].dtype == np.float64 and df["lags"].dtype == np.float64
    edges = vo.default_bins(5.0, float(np.hypot(199 * 5.0, 199 * 5.0)))
    assert len(df) == len(edges) - 1 and np.allclose(df["lags"].values, edges[:-1])
    assert df["err_exp"].isna().all()
    # several runs: aggregated mean / std / summed counts
    df3 = xs.sample_empirical_variogram(vals, gsd=5.0, subsample=400, subsample_method="pdist_point", n_variograms=3,
                                        random_state=7, bin_func="even", n_lags=20)
    assert len(df3) >= 19 and (df3["count"] > 0).any()
    # the reference's DEFAULT call (cdist_equidistant sampler, spatialstats.py:1301) is on the path
    dfd = xs.sample_empirical_variogram(vals, gsd=5.0, subsample=100, random_state=11)
    assert list(dfd.columns) == ["exp", "lags", "count", "err_exp"] and len(dfd) == len(edges) - 1
    assert dfd["count"].sum() > 0 and np.isfinite(dfd["exp"][dfd["count"] > 0]).all()
    with pytest.raises(ValueError, match="ground sampling distance must be defined"):
        xs.sample_empirical_variogram(vals, subsample=100, subsample_method="pdist_point")


def _edge_golden() -> dict:
    import json
    import os

    with open(os.path.join(parity.GOLDEN, "variogram_edges.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("rule", ["left", "right"])
@pytest.mark.parametrize("gsd", [1.0, 5.0, 0.5])
def test_pair_kernels_on_edge_goldens_both_rules(rule: str, gsd: float) -> None:
    """Pairs sitting EXACTLY on bin edges -- the only pairs on which the two possible conventions of scikit-gstat's lag
    classes differ (PARITY UNPINNED) -- against the hand-checkable exact-integer golden cases, for BOTH conventions and
    BOTH kernels (integer grid kernel, float64-coordinate kernel).  The product's rule is one flag
    (xdem_b200.spatialstats.LAG_EDGE_RULE); whichever skgstat turns out to use, the matching counts are pinned here."""
    from fractions import Fraction

    import torch

    from xdem_b200 import spatialstats as xs

    for c in _edge_golden()["cases"]:
        e_sq = [Fraction(n, d) for n, d in c["edges_sq_num_den"]]
        if not all(float(np.sqrt(float(q))) ** 2 == float(q) for q in e_sq):
            continue  # irrational edges: the float edge is a rounded neighbour of the lattice distance (CPU test)
        edges = np.array([gsd * float(np.sqrt(float(q))) for q in e_sq])
        pts = np.asarray(c["points"], dtype=np.int64)
        vals = np.asarray(c["values"], dtype=np.float32)
        x, y, v = (torch.from_numpy(a).cuda() for a in (pts[:, 0].copy(), pts[:, 1].copy(), vals))
        want_cnt = np.asarray(c[rule]["count"])
        want_ssq = np.asarray([n / d for n, d in c[rule]["sumsq"]])
        _, cnt, ssq = xs.pairwise_lag_binning(x, y, v, edges, gsd, edge_rule=rule)
        assert np.array_equal(cnt, want_cnt) and np.allclose(ssq, want_ssq, rtol=1e-12), (c["name"], rule, "grid")
        ps = xs.PairSet(pts[:, 0] * gsd, pts[:, 1] * gsd, vals.astype(np.float64))
        _, cnt2, ssq2 = xs.pairwise_lag_binning_xy(ps, edges, edge_rule=rule)
        assert np.array_equal(cnt2, want_cnt) and np.allclose(ssq2, want_ssq, rtol=1e-12), (c["name"], rule, "xy")


@pytest.mark.parametrize("estimator", ["matheron", "cressie", "dowd"])
@pytest.mark.parametrize("gsd", [5.0, 0.1, 0.7])
def test_xy_kernel_vs_float64_oracle(estimator: str, gsd: float) -> None:
    """float64-coordinate kernel == the pdist-based oracle: counts identical for ANY spacing (incl. non-dyadic gsd,
    where the integer kernel is only exact off the edges), estimators within float64 / float32-key round-off; both
    through `_get_pdist_empirical_variogram` (the function install() rebinds) and with arbitrary scattered points."""
    from oracle import variogram_oracle as vo
    from xdem_b200 import spatialstats as xs

    rng = np.random.default_rng(21)
    shape = (70, 90)
    coords = vo.grid_coords(shape, gsd)
    vals = rng.normal(size=shape[0] * shape[1]).astype(np.float32)
    idx = rng.choice(vals.size, 900, replace=False)
    maxlag = float(np.hypot(coords[:, 0].max(), coords[:, 1].max()))
    for bins in (vo.default_bins(gsd, maxlag), "even"):
        b_o, exp_o, cnt_o = vo.empirical_variogram(coords[idx], vals[idx], bins, n_lags=17, maxlag=maxlag,
                                                   estimator=estimator)
        df = xs._get_pdist_empirical_variogram(values=vals[idx], coords=coords[idx], bin_func=bins, n_lags=17,
                                               maxlag=maxlag, estimator=estimator, random_state=None)
        assert np.array_equal(df["bins"].values, b_o)
        assert np.array_equal(df["count"].values, cnt_o)
        assert np.allclose(df["exp"].values, exp_o, rtol=2e-6 if estimator == "dowd" else 1e-12, equal_nan=True)
    # scattered float coordinates, two sets (cdist): brute force in NumPy
    a, b = rng.uniform(0, 100, (300, 2)), rng.uniform(0, 100, (450, 2))
    va, vb = rng.normal(size=300), rng.normal(size=450)
    edges = np.linspace(0, 150, 13)[1:]
    d = np.sqrt((a[:, None, 0] - b[None, :, 0]) ** 2 + (a[:, None, 1] - b[None, :, 1]) ** 2)
    df2 = np.abs(va[:, None] - vb[None, :])
    lo = np.concatenate([[0.0], edges[:-1]])
    cnt_w = np.array([int(((d >= lo[k]) & (d < edges[k])).sum()) for k in range(len(edges))])
    _, cnt, third = xs.pairwise_lag_binning_xy(xs.PairSet(a[:, 0], a[:, 1], va, b[:, 0], b[:, 1], vb), edges,
                                               estimator=estimator, edge_rule="left")
    assert np.array_equal(cnt, cnt_w)
    for k in range(len(edges)):
        x = df2[(d >= lo[k]) & (d < edges[k])]
        if x.size == 0:
            continue
        want = {"matheron": np.sum(x**2), "cressie": np.sum(np.sqrt(x)), "dowd": np.median(x)}[estimator]
        assert third[k] == pytest.approx(want, rel=2e-6 if estimator == "dowd" else 1e-12)


def test_cdist_pair_deduplication_and_samplers() -> None:
    """A x B with shared samples: a sample is never paired with itself and a pair of two shared samples counts once;
    cdist_point / cdist_equidistant / pdist_ring / pdist_disk and the 1-D values + coords input run end to end."""
    from xdem_b200 import spatialstats as xs

    rng = np.random.default_rng(5)
    pts = rng.uniform(0, 50, (40, 2))
    v = rng.normal(size=40)
    ia, ib = np.arange(0, 25), np.arange(15, 40)  # samples 15..24 are in both sets
    edges = np.array([1000.0])
    _, cnt, ssq = xs.pairwise_lag_binning_xy(
        xs.PairSet(pts[ia, 0], pts[ia, 1], v[ia], pts[ib, 0], pts[ib, 1], v[ib], ida=ia, idb=ib), edges)
    pairs = {(min(i, j), max(i, j)) for i in ia for j in ib if i != j}
    assert int(cnt[0]) == len(pairs)
    assert ssq[0] == pytest.approx(sum((v[i] - v[j]) ** 2 for i, j in pairs), rel=1e-12)

    vals, _ = _sample((120, 150), 10, 9, nan_frac=0.02)
    for method in ("cdist_equidistant", "cdist_point", "pdist_ring", "pdist_disk", "pdist_point"):
        for est in ("matheron", "dowd"):
            df = xs.sample_empirical_variogram(vals, gsd=2.0, subsample=200, subsample_method=method, estimator=est,
                                               random_state=3)
            assert list(df.columns) == ["exp", "lags", "count", "err_exp"] and df["count"].sum() > 0, (method, est)
            ok = df["count"] > 0
            assert np.isfinite(df["exp"][ok]).all() and (df["exp"][ok] >= 0).all()
    # white noise of unit variance: every well-populated lag class has a semivariance near 1 with the default sampler
    df = xs.sample_empirical_variogram(vals, gsd=2.0, subsample=2000, random_state=1)
    big = df["count"] > 2000
    assert big.any() and np.allclose(df["exp"][big], 1.0, atol=0.15)
    # 1-D values + coords == the same samples passed as a grid (pdist_point with every valid sample drawn)
    small = vals[:20, :30].copy()
    gx, gy = np.meshgrid(np.arange(0, 20 * 2.0, 2.0), np.arange(0, 30 * 2.0, 2.0))
    coords = np.dstack((gx.flatten(), gy.flatten())).squeeze()
    bf = [2.0, 4.0, 7.5, 15.0, 33.3, 80.0]
    d1 = xs.sample_empirical_variogram(small.flatten(), coords=coords, subsample=10**6, subsample_method="pdist_point",
                                       random_state=2, bin_func=bf)
    d2 = xs.sample_empirical_variogram(small, gsd=2.0, subsample=10**6, subsample_method="pdist_point", random_state=2,
                                       bin_func=bf)
    assert np.array_equal(d1["count"].values, d2["count"].values)
    assert np.allclose(d1["exp"].values, d2["exp"].values, rtol=1e-5, equal_nan=True)
    assert np.allclose(d1["lags"].values, d2["lags"].values)


def test_variogram_large_properties() -> None:
    """Size-independent properties at N = 2e5 (2e10 pairs): every pair lands in exactly one class (sum of counts ==
    N(N-1)/2 with a last edge beyond the extent), the result does not depend on sample order, and a constant field has
    zero semivariance."""
    import torch

    from xdem_b200 import spatialstats as xs

    g = torch.Generator(device="cuda").manual_seed(5)
    N, S = 200_000, 20000
    lin = torch.randperm(S * S // 1000, generator=g, device="cuda")[:N].to(torch.int64) * 1000 + 7
    x, y = lin % S, lin // S
    v = torch.randn(N, generator=g, device="cuda")
    edges = np.linspace(0, 1.5 * S * 5.0, 41)[1:]
    e, cnt, ssq = xs.pairwise_lag_binning(x, y, v, edges, 5.0)
    assert int(cnt.sum()) == N * (N - 1) // 2
    perm = torch.randperm(N, generator=g, device="cuda")
    e2, cnt2, ssq2 = xs.pairwise_lag_binning(x[perm], y[perm], v[perm], edges, 5.0)
    assert np.array_equal(cnt, cnt2) and np.allclose(ssq, ssq2, rtol=1e-6)
    _, cnt3, ssq3 = xs.pairwise_lag_binning(x, y, torch.full_like(v, 3.25), edges, 5.0)
    assert np.array_equal(cnt3, cnt) and np.all(ssq3 == 0)
    # white noise: semivariance ~ variance (=1) in every populated class
    with np.errstate(all="ignore"):
        gamma = ssq / (2.0 * cnt)
    assert np.allclose(gamma[cnt > 1e6], 1.0, atol=0.06)  # statistical sanity only


# ------------------------------------------------------------------------------------------------- Nuth & Kaab


def test_nk_aux_bit_exact() -> None:
    import torch

    from xdem_b200 import coreg

    g = parity.load_golden("nk_reference.npz")
    st = coreg._NKState(torch.from_numpy(g["ref"]).cuda(), torch.from_numpy(g["tba"]).cuda(), None)
    ref_st = g["slope_tan"].copy()
    ref_st[np.isclose(ref_st, 0)] = np.nan
    assert np.array_equal(st.slope_tan.cpu().numpy(), ref_st, equal_nan=True)  # IEEE ops in NumPy's order
    asp = st.aspect.cpu().numpy()
    assert parity.nanmask_equal(asp, g["aspect"])
    assert np.nanmax(np.abs(asp - g["aspect"])) <= 4 * np.spacing(np.float32(6.3))  # atan2f: a few ulp


@pytest.mark.parametrize("shape", [(257, 403), (256, 512)])
def test_nk_prepare_mask_count_and_range_candidates(shape) -> None:
    """xb_nk_prepare (one pass: aux variables + validity mask + count + aspect-range candidates), scalar kernel (odd
    width) and 4-pixels-per-thread kernel: slope_tan / aspect identical to xb_nk_aux, mask == inlier & finite(ref, tba,
    slope_tan, aspect) (base.py:653-661), the count, and candidates that really attain the min / max valid aspect."""
    import ctypes

    import torch

    from oracle import synth
    from xdem_b200 import _lib, coreg

    ref, tba = synth.nk_pair(shape, shift_px=(0.3, -0.2), dz=1.0, noise=0.05)
    ref, tba = ref.astype(np.float32), tba.astype(np.float32)
    ref[10:14, 20:40] = np.nan
    tba[100:120, 5:9] = np.nan
    ref[200:210, 300:330] = 7.0  # flat patch: slope_tan == 0 -> NaN (affine.py:578-579)
    inl = np.random.default_rng(3).random(shape) < 0.9
    rt, tt = torch.from_numpy(ref).cuda(), torch.from_numpy(tba).cuda()
    for mask in (None, torch.from_numpy(inl).cuda()):
        st = coreg._NKState(rt, tt, mask)
        L = _lib.lib()
        st_ref = torch.empty_like(rt)
        asp_ref = torch.empty_like(rt)
        _lib.check(L.xb_nk_aux(rt.data_ptr(), shape[0], shape[1], rt.stride(0), 1, 1, 0, shape[0], st_ref.data_ptr(),
                               asp_ref.data_ptr(), shape[1], ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
        assert torch.equal(st.slope_tan.view(torch.int32), st_ref.view(torch.int32))
        assert torch.equal(st.aspect.view(torch.int32), asp_ref.view(torch.int32))
        valid = torch.isfinite(rt) & torch.isfinite(tt) & torch.isfinite(st_ref) & torch.isfinite(asp_ref)
        if mask is not None:
            valid &= mask
        assert torch.equal(st.sub_mask.bool(), valid) and torch.equal(st.valid, valid)
        assert st.n_valid() == int(valid.sum().item())
        rc = st.range_cand.cpu().numpy().view(np.uint32)
        asp = st.aspect.cpu().numpy()
        v = valid.cpu().numpy()
        assert rc[0:1].view(np.float32)[0] == asp[v].min() and rc[1:2].view(np.float32)[0] == asp[v].max()
        assert 1 <= rc[2] <= 64 and 1 <= rc[3] <= 64
        for k in range(int(rc[2])):
            i = int(rc[4 + k])
            assert v.flat[i] and asp.flat[i] == asp[v].min()
        for k in range(int(rc[3])):
            i = int(rc[4 + 64 + k])
            assert v.flat[i] and asp.flat[i] == asp[v].max()


def test_nk_step_pieces_vs_oracle() -> None:
    """dh, its exact median, the per-bin medians / counts and p0 of one iteration against the oracle."""
    import torch

    from oracle import nk_oracle as nk
    from xdem_b200 import coreg

    g = parity.load_golden("nk_reference.npz")
    ref, tba, inl = g["ref"], g["tba"], g["inlier"]
    a_e = (5.0, -5.0)
    slope_tan, aspect = nk.aux_vars(ref)
    valid = np.isfinite(ref) & np.isfinite(tba) & np.isfinite(slope_tan) & np.isfinite(aspect) & inl
    st = coreg._NKState(torch.from_numpy(ref).cuda(), torch.from_numpy(tba).cuda(), torch.from_numpy(inl).cuda())
    assert np.array_equal(st.valid.cpu().numpy(), valid)
    for offs in ((0.0, 0.0), (1.85, 3.05), (-7.3, 12.9), (5.0, -10.0)):
        dx, dy = offs[0] / a_e[0], offs[1] / a_e[1]
        lo, hi, n_fin = st.compute_dh(dx, dy)
        dh_o = nk.dh_at(ref, tba, valid, dx, dy)
        dh_g = st.dh.cpu().numpy().reshape(ref.shape)[valid]
        assert np.array_equal(np.isnan(dh_g), np.isnan(dh_o)), offs
        assert n_fin == int(np.isfinite(dh_o).sum())
        m = np.isfinite(dh_o)
        assert np.allclose(dh_g[m], dh_o[m], rtol=0, atol=2e-4)  # float32 storage of values ~1e3 apart
        med, cnt, _ = st.select_medians(0, 0.0, 0.0, 1.0, 1)
        assert cnt[0] == n_fin
        assert med[0] == pytest.approx(float(np.median(dh_g[m].astype(np.float64))), rel=0, abs=0)  # exact select
        assert med[0] == pytest.approx(float(np.nanmedian(dh_o)), abs=2e-4)
        # aspect range over finite dh
        asp_v = aspect[valid][m]
        assert lo == float(asp_v.min()) and hi == float(asp_v.max())
        # per-bin medians of y
        _, _, dbg = nk.iteration_step((offs[0], offs[1], 0.0), ref, tba, valid, slope_tan, aspect, a_e)
        med_b, cnt_b, mom = st.select_medians(1, dbg["vshift"], lo, hi, 72, want_moments=True)
        assert np.array_equal(cnt_b, dbg["count"].astype(np.int64)), offs
        assert np.allclose(med_b, dbg["median"], rtol=2e-4, atol=2e-4, equal_nan=True)
        n, s1, s2 = mom
        assert n == cnt_b.sum()
        assert s1 / n == pytest.approx(dbg["p0"][2], rel=1e-3, abs=1e-3)


@pytest.mark.parametrize("ncols", [299, 296, 5, 4])
def test_nk_dh_vector_and_scalar_kernels(ncols: int) -> None:
    """xb_nk_dh picks a 4-pixels-per-thread kernel for 16-byte aligned rasters and a scalar one otherwise: both against
    the oracle, including shifts whose 2x2 stencils cross the raster border."""
    import torch

    from oracle import nk_oracle as nk
    from xdem_b200 import coreg

    g = parity.load_golden("nk_reference.npz")
    ref, tba, inl = (np.ascontiguousarray(g[k][:, :ncols]) for k in ("ref", "tba", "inlier"))
    slope_tan, aspect = nk.aux_vars(ref)
    valid = np.isfinite(ref) & np.isfinite(tba) & np.isfinite(slope_tan) & np.isfinite(aspect) & inl
    st = coreg._NKState(torch.from_numpy(ref).cuda(), torch.from_numpy(tba).cuda(), torch.from_numpy(inl).cuda())
    for dx, dy in ((0.0, 0.0), (0.37, -0.61), (-1.46, 2.58), (3.0, -2.0), (-4.75, 0.5)):
        _, _, n_fin = st.compute_dh(dx, dy)
        dh_o = nk.dh_at(ref, tba, valid, dx, dy)
        dh_g = st.dh.cpu().numpy().reshape(ref.shape)[valid]
        assert np.array_equal(np.isnan(dh_g), np.isnan(dh_o)), (ncols, dx, dy)
        assert n_fin == int(np.isfinite(dh_o).sum())
        m = np.isfinite(dh_o)
        assert np.allclose(dh_g[m], dh_o[m], rtol=0, atol=2e-4)
        # everything outside the subsample stays NaN
        assert np.isnan(st.dh.cpu().numpy().reshape(ref.shape)[~valid]).all()


def test_nk_full_fit_vs_reference_fixture() -> None:
    """Whole fit against the per-iteration outputs of the reference's own code (tests/golden/nk_reference.npz)."""
    from xdem_b200 import coreg

    g = parity.load_golden("nk_reference.npz")
    tr = tuple(g["transform"])
    for n_it in (1, 3, 6):
        (e, n, v), n_used = coreg.nuth_kaab(g["ref"], g["tba"], inlier_mask=g["inlier"], transform=tr, tolerance=0.0,
                                            max_iterations=n_it)
        assert n_used == int(g["n_valid"])
        assert np.allclose([e, n, v], g["offsets"][n_it - 1], rtol=1e-4, atol=2e-4), (n_it, (e, n, v))
    nkc = coreg.NuthKaab(max_iterations=6, offset_threshold=0.0, subsample=1.0)
    nkc.fit(g["ref"], g["tba"], inlier_mask=g["inlier"], transform=tr)
    sx, sy, sz = nkc.to_translations()
    # injected: tba = ref shifted by (+0.37, -0.61) px and +1.5 m  ->  shift_x = +0.37*5, shift_y = +0.61*5, z = -1.5
    assert sx == pytest.approx(1.85, abs=0.01) and sy == pytest.approx(3.05, abs=0.01)
    assert sz == pytest.approx(-1.5, abs=0.01)
    assert np.allclose(nkc.to_matrix()[:3, 3], [sx, sy, sz])


def test_nk_subsample_and_large() -> None:
    """Larger pair on the device (2048^2) with the default-style subsample: recovers the injected shift."""
    import torch

    from oracle import synth
    from xdem_b200 import coreg

    ref, tba = synth.nk_pair((1024, 1280), shift_px=(-0.83, 0.42), dz=-2.0, noise=0.02)
    nkc = coreg.NuthKaab(subsample=2e5)
    nkc.fit(torch.from_numpy(ref).cuda(), torch.from_numpy(tba).cuda(), transform=(2.0, 0, 0, 0, -2.0, 0),
            random_state=42)
    sx, sy, sz = nkc.to_translations()
    assert sx == pytest.approx(-0.83 * 2.0, abs=0.02) and sy == pytest.approx(-0.42 * 2.0, abs=0.02)
    assert sz == pytest.approx(2.0, abs=0.02)
    assert nkc.meta["outputs"]["random"]["subsample_final"] == 200000


@pytest.mark.parametrize("case", ["smooth", "holes_and_ties", "subsample"])
def test_nk_bracketed_selection_equals_exhaustive(case: str) -> None:
    """The two-pass bracketed selection (csrc/xb_nk_fast.cu) returns EXACTLY what the exhaustive radix select returns --
    the median of dh, the per-bin medians of y, the per-bin counts, the aspect range, n -- over several shifts, with NaN
    holes, heavy ties (quantised elevations), an odd / even number of valid pixels and a sparse subsample; the moments
    agree to summation order."""
    import torch

    from oracle import synth
    from xdem_b200 import coreg

    ref, tba = synth.nk_pair((1536, 2048), shift_px=(0.37, -0.61), dz=1.5, noise=0.05)
    inl = np.ones(ref.shape, dtype=bool)
    if case == "holes_and_ties":
        ref, tba = np.round(ref * 4) / 4, np.round(tba * 4) / 4  # many identical dh values
        tba[100:400, 300:900] = np.nan
        ref[::37, ::11] = np.nan
        inl[1000:1001, :777] = False  # flips the parity of the valid count
    if case == "subsample":
        inl = np.random.default_rng(1).random(ref.shape) < 0.07
    rt, tt, it = (torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (ref.astype(np.float32),
                                                                              tba.astype(np.float32), inl))
    st = coreg._NKState(rt, tt, it)
    assert st.fast_eligible(72)
    n_fallback = 0
    # floor(dx) mod 4 = 0, 0, 2, 3, 1, 0: every column-shift instantiation of the full dh pass
    for dx, dy in ((0.0, 0.0), (0.37, -0.61), (-1.46, 2.58), (3.0, -2.0), (1.3, 0.75), (0.3712, -0.6093)):
        res = st.iteration_fast(dx, dy, 72)
        if res is None:
            n_fallback += 1
            print("fallback", case, dx, dy, "flags", st.fast_last_flags)
            continue
        lo, hi, n_fin = st.compute_dh(dx, dy)
        med, cnt, _ = st.select_medians(0, 0.0, 0.0, 1.0, 1)
        assert res["n_fin"] == n_fin == cnt[0] and res["lo"] == lo and res["hi"] == hi
        assert res["vshift"] == float(med[0]), (case, dx, dy, res["vshift"], med[0])
        med_b, cnt_b, mom = st.select_medians(1, float(med[0]), lo, hi, 72, want_moments=True)
        assert np.array_equal(res["counts"], cnt_b), (case, dx, dy)
        assert np.array_equal(res["median"], med_b, equal_nan=True), (case, dx, dy, np.nanmax(np.abs(res["median"] - med_b)))
        # n exact; sum y / sum y^2 only seed the optimiser's initial guess (affine.py:384): the fast path accumulates
        # them in float32 per thread and row before going to float64
        assert res["moments"][0] == mom[0] and np.allclose(res["moments"], mom, rtol=1e-5, atol=1e-3 * mom[0] ** 0.5)
    assert n_fallback == 0, "a 4-sigma bracket should not miss on these inputs"
    # whole fits: identical offsets with and without the fast path
    tr = (5.0, 0, 0, 0, -5.0, 0)
    fast = coreg.nuth_kaab(rt, tt, inlier_mask=it, transform=tr, tolerance=0.0, max_iterations=4)
    coreg.NK_FAST = False
    try:
        slow = coreg.nuth_kaab(rt, tt, inlier_mask=it, transform=tr, tolerance=0.0, max_iterations=4)
    finally:
        coreg.NK_FAST = True
    # medians / counts are identical (asserted above); the moments that seed Levenberg-Marquardt's first guess differ in
    # the last float32 digits (atomic accumulation order), so the two fits stop within the optimiser's own tolerance
    # of each other (observed <= 2e-6 relative), three orders below the 1e-3 px convergence threshold
    assert fast[1] == slow[1] and np.allclose(fast[0], slow[0], rtol=1e-5, atol=1e-7), (fast, slow)


def test_nk_point_list_iteration_equals_exhaustive() -> None:
    """Point-list fits (the reference's default: a random subsample of points): the all-device iteration
    (xb_nkf_iteration_points) returns exactly what the exhaustive host-driven radix select returns on the same points --
    n, aspect range, median of dh, per-bin medians and counts -- and the picked points are valid, unique and as many as
    asked for; the whole fit recovers the synthetic shift."""
    import torch

    from oracle import synth
    from xdem_b200 import coreg

    ref, tba = synth.nk_pair((1200, 1604), shift_px=(0.37, -0.61), dz=1.5, noise=0.05)
    tba[100:300, 300:700] = np.nan
    rt, tt = torch.from_numpy(ref).cuda(), torch.from_numpy(tba).cuda()
    st = coreg._NKState(rt, tt, None)
    idx = coreg._pick_valid_points(st, 200_000, st.n_valid(), 7)
    assert idx.numel() == 200_000 and torch.unique(idx).numel() == 200_000
    assert bool(st.sub_mask.view(-1)[idx].all())
    assert torch.equal(idx, coreg._pick_valid_points(st, 200_000, st.n_valid(), 7))  # deterministic for an int seed
    st.set_points(idx)
    for dx, dy in ((0.0, 0.0), (0.37, -0.61), (-1.46, 2.58), (25.0, -13.5)):
        res = st.iteration_points(dx, dy, 72)
        assert res is not None
        lo, hi, n_fin = st.compute_dh(dx, dy)
        med, cnt, _ = st.select_medians(0, 0.0, 0.0, 1.0, 1)
        assert res["n_fin"] == n_fin == cnt[0] and res["lo"] == lo and res["hi"] == hi
        assert res["vshift"] == float(med[0])
        med_b, cnt_b, mom = st.select_medians(1, float(med[0]), lo, hi, 72, want_moments=True)
        assert np.array_equal(res["counts"], cnt_b)
        assert np.array_equal(res["median"], med_b, equal_nan=True)
        assert res["moments"][0] == mom[0] and np.allclose(res["moments"], mom, rtol=1e-9)
    (e, n, vz), used = coreg.nuth_kaab(rt, tt, transform=(5.0, 0, 0, 0, -5.0, 0), tolerance=0.0, max_iterations=6,
                                       params_random={"subsample": 3e5, "random_state": 1})
    assert used == 300_000
    assert abs(e / 5 + 0.37) < 2e-2 and abs(n / 5 + 0.61) < 2e-2 and abs(vz + 1.5) < 2e-2, (e, n, vz)


def test_nk_apply_translation() -> None:
    """`apply`: regrid of the shifted DEM on the input grid == map_coordinates restatement; after
    applying the fitted shift the residual dh median is ~0 and a second fit finds (almost) no shift."""
    from scipy.ndimage import map_coordinates

    from xdem_b200 import coreg

    g = parity.load_golden("nk_reference.npz")
    tr = tuple(g["transform"])
    nkc = coreg.NuthKaab(max_iterations=8, subsample=1.0).fit(g["ref"], g["tba"], inlier_mask=g["inlier"], transform=tr)
    sx, sy, sz = nkc.to_translations()
    applied, tr2 = nkc.apply(g["tba"], transform=tr)
    assert tr2 == tr and applied.dtype == np.float32 and applied.shape == g["tba"].shape
    rows, cols = np.mgrid[0:g["tba"].shape[0], 0:g["tba"].shape[1]].astype(np.float64)
    expect = map_coordinates(g["tba"].astype(np.float64), [rows - sy / tr[4], cols - sx / tr[0]], order=1,
                             mode="constant", cval=np.nan, prefilter=False) + sz
    assert np.array_equal(np.isnan(applied), np.isnan(expect))
    assert np.nanmax(np.abs(applied - expect)) < 3e-4
    resid = g["ref"] - applied
    assert abs(np.nanmedian(resid)) < 0.01
    nk2 = coreg.NuthKaab(max_iterations=5, subsample=1.0).fit(g["ref"], applied, inlier_mask=g["inlier"], transform=tr)
    assert all(abs(v) < 0.05 for v in nk2.to_translations())
    shifted, tr3 = nkc.apply(g["tba"], transform=tr, resample=False)
    assert np.allclose(shifted, g["tba"] + sz, equal_nan=True) and tr3[2] == tr[2] + sx and tr3[5] == tr[5] + sy


@pytest.mark.parametrize("estimator", ["cressie", "dowd"])
@pytest.mark.parametrize("n", [2500, 257])
def test_robust_estimators_vs_oracle(estimator: str, n: int) -> None:
    """Cressie-Hawkins and Dowd against the restated scikit-gstat estimators."""
    import torch

    from oracle import variogram_oracle as vo
    from xdem_b200 import spatialstats as xs

    shape, gsd = (260, 260), 5.0
    vals, idx = _sample(shape, n, 91)
    vals = (vals * 3).astype(np.float32)
    coords = vo.grid_coords(shape, gsd)
    maxlag = float(np.hypot(259 * gsd, 259 * gsd))
    edges_in = np.asarray(vo.default_bins(gsd, maxlag))
    b_o, exp_o, cnt_o = vo.empirical_variogram(coords[idx], vals.ravel()[idx], edges_in, estimator=estimator)
    ti = torch.from_numpy(idx).cuda()
    edges, cnt, third = xs.pairwise_lag_binning(ti % 260, ti // 260, torch.from_numpy(vals.ravel()).cuda()[ti],
                                                edges_in, gsd, estimator=estimator)
    assert np.array_equal(cnt, cnt_o)
    exp = xs.estimate_from_sums(cnt, third, estimator)
    assert np.allclose(exp, exp_o, rtol=2e-5, equal_nan=True), (exp, exp_o)
    df = xs.sample_empirical_variogram(vals, gsd=gsd, subsample=600, subsample_method="pdist_point", random_state=3,
                                       estimator=estimator)
    assert np.isfinite(df["exp"].values[df["count"].values > 0]).all()
