"""`nd_binning`: SciPy-based oracle vs fixtures of the unmodified reference (CPU), the CUDA
path vs both (GPU).  Counts are integers (exact); medians / NMADs are order statistics of float32 data (exact selects),
compared at 1e-6 relative to absorb the float32-vs-float64 mean of two middle values."""

from __future__ import annotations

import itertools

import numpy as np
import pytest

from tests import parity

NAMES = ["slope", "curv", "elev"]
CASES = {
    "one10": (["slope"], 10),
    "two": (["slope", "curv"], (8, 5)),
    "three_edges": (NAMES, (np.array([0, 5, 10, 20, 40, 90], dtype=np.float32),
                            np.array([-5, -1, 0, 1, 5], dtype=np.float32),
                            np.array([800, 1500, 2000, 2700], dtype=np.float32))),
    "three_int": (NAMES, 4),
}


@pytest.fixture(scope="module")
def B() -> dict[str, np.ndarray]:
    return parity.load_golden("binning_reference.npz")


def _combos(n: int) -> list[tuple[int, ...]]:
    c: list[tuple[int, ...]] = [(i,) for i in range(n)]
    if n > 1:
        c += list(itertools.combinations(range(n), 2))
    if n > 2:
        c.append(tuple(range(n)))
    return c


def _close(a: np.ndarray, b: np.ndarray, msg: str) -> None:
    assert a.shape == b.shape, msg
    assert np.array_equal(np.isnan(a), np.isnan(b)), f"{msg}: NaN pattern"
    m = np.isfinite(b)
    assert np.allclose(a[m], b[m], rtol=1e-6, atol=1e-7), f"{msg}: max diff {np.max(np.abs(a[m] - b[m]))}"


def test_oracle_matches_reference_fixtures(B) -> None:
    from oracle import binning_oracle as bo

    for tag, (names, bins) in CASES.items():
        lv = [B[f"in|{n}"] for n in names]
        lb = list(bins) if isinstance(bins, tuple) else [bins] * len(names)
        res = bo.nd_binning(B["in|values"], lv, lb)
        cnt = np.concatenate([res[c]["count"] for c in _combos(len(names))])
        med = np.concatenate([res[c]["median"] for c in _combos(len(names))])
        nm = np.concatenate([res[c]["nmad"] for c in _combos(len(names))])
        assert np.array_equal(cnt, B[f"{tag}|count"]), tag
        _close(med, B[f"{tag}|nanmedian"], f"{tag} median")
        _close(nm, B[f"{tag}|nmad"], f"{tag} nmad")


def test_host_helpers() -> None:
    from xdem_b200 import binning as xb

    e = xb.bin_edges(1.0, 3.0, 4, np.float32)
    assert e.dtype == np.float32 and np.array_equal(e, np.linspace(1, 3, 5, dtype=np.float32))
    assert np.array_equal(xb.bin_edges(2.0, 2.0, 2, np.float32), np.array([1.5, 2.0, 2.5], dtype=np.float32))
    assert np.array_equal(xb.bin_edges(0, 0, [0, 1, 5], np.float32), np.array([0, 1, 5], dtype=np.float32))
    hist = np.array([[0, 3, 0, 2], [1, 1, 1, 1], [0, 0, 0, 0]])
    counts = hist.sum(axis=1)
    digit, below = xb._pick_digit(hist, np.array([3, 1, 0]), counts)
    assert digit.tolist() == [3, 1, 0] and below.tolist() == [3, 1, 0]
    assert np.array_equal(xb._key_to_float(np.array([0x80000000 | np.float32(1.5).view(np.uint32)], dtype=np.uint32)),
                          np.array([1.5], dtype=np.float32))
    with pytest.raises(NotImplementedError, match="arbitrary Python callable"):
        xb._stat_kind(lambda a: 0.0)
    with pytest.raises(NotImplementedError, match="not available"):
        xb._stat_kind("kurtosis")
    assert xb._stat_kind(np.nanstd) == ("nanstd", "std") and xb._stat_kind("mean") == ("mean", "mean")
    assert xb._stat_kind(np.nanmax) == ("nanmax", "max") and xb._stat_kind(np.sum) == ("sum", "sum")
    assert xb._stat_kind(np.nanmedian) == ("nanmedian", "median")
    assert xb._stat_kind(xb.nmad) == ("nmad", "nmad")


def _float_to_key(v: np.ndarray) -> np.ndarray:
    u = v.astype(np.float32).view(np.uint32)
    return np.where(u & np.uint32(0x80000000), ~u, u | np.uint32(0x80000000)).astype(np.uint32)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_radix_select_host_driver_with_numpy_passes(seed: int) -> None:
    """The host side of the exact median (digit picking, ties, even counts, the upper middle value taken from the last
    histogram or from a search pass, empty bins) against np.median, with NumPy stand-ins for the two kernels."""
    from xdem_b200 import binning as xb

    rng = np.random.default_rng(seed)
    n_bins = 7
    n = 4000
    bins = rng.integers(0, n_bins - 1, n)  # the last bin stays empty
    vals = rng.normal(size=n).astype(np.float32)
    vals[rng.choice(n, 800, replace=False)] = np.round(vals[:800] * 2) / 2  # heavy ties, +-0.0
    if seed == 1:  # a bin with two samples whose keys differ in the top digit (needs the search pass)
        sel = np.flatnonzero(bins == 2)
        bins[sel[2:]] = 3
        vals[sel[0]], vals[sel[1]] = np.float32(-3.5e20), np.float32(7.25e-12)
    if seed == 2:  # even count, neighbours in the same 256-key bucket (taken from the last histogram)
        sel = np.flatnonzero(bins == 4)
        bins[sel[4:]] = 0
        vals[sel[:4]] = np.array([1.0, np.nextafter(np.float32(1.0), np.float32(2.0)), 5.0, -5.0], dtype=np.float32)
    keys = _float_to_key(vals)
    calls = {"hist": 0, "next": 0}

    def hist_fn(prefix: np.ndarray, mask: int, shift: int) -> np.ndarray:
        calls["hist"] += 1
        h = np.zeros((n_bins, 256), dtype=np.int64)
        ok = (keys & np.uint32(mask)) == prefix[bins]
        np.add.at(h, (bins[ok], (keys[ok] >> np.uint32(shift)) & np.uint32(255)), 1)
        return h

    def next_fn(sel_keys: np.ndarray) -> np.ndarray:
        calls["next"] += 1
        out = np.full(n_bins, 0xFFFFFFFF, dtype=np.uint32)
        for b in range(n_bins):
            k = keys[(bins == b) & (keys > sel_keys[b])]
            if k.size:
                out[b] = k.min()
        return out

    med, cnt = xb.radix_select_medians(hist_fn, next_fn, n_bins)
    assert calls["hist"] == 4 and calls["next"] <= 1
    for b in range(n_bins):
        sel = vals[bins == b]
        assert cnt[b] == sel.size
        if sel.size == 0:
            assert np.isnan(med[b])
        else:
            assert med[b] == np.median(sel), (b, med[b], np.median(sel))  # float32 median, bit for bit
    if seed == 1:
        assert calls["next"] == 1


# ---------------------------------------------------------------------------------------------------------- GPU


@pytest.mark.gpu
def test_gpu_nd_binning_vs_reference_fixtures(B) -> None:
    from xdem_b200 import spatialstats as xs

    for tag, (names, bins) in CASES.items():
        lv = [B[f"in|{n}"] for n in names]
        df = xs.nd_binning(B["in|values"], lv, list(names), list_var_bins=bins)
        assert np.array_equal(df["nd"].to_numpy(), B[f"{tag}|nd"]), tag
        assert np.array_equal(df["count"].to_numpy(), B[f"{tag}|count"]), tag
        _close(df["nanmedian"].to_numpy(), B[f"{tag}|nanmedian"], f"{tag} median")
        _close(df["nmad"].to_numpy(), B[f"{tag}|nmad"], f"{tag} nmad")
        for n in names:
            left = np.array([iv.left if hasattr(iv, "left") else np.nan for iv in df[n]], dtype=np.float64)
            right = np.array([iv.right if hasattr(iv, "right") else np.nan for iv in df[n]], dtype=np.float64)
            assert np.array_equal(left, B[f"{tag}|{n}|left"], equal_nan=True), (tag, n)
            assert np.array_equal(right, B[f"{tag}|{n}|right"], equal_nan=True), (tag, n)


@pytest.mark.gpu
def test_gpu_statistics_selection_and_edges() -> None:
    import torch

    from oracle import binning_oracle as bo
    from xdem_b200 import spatialstats as xs

    rng = np.random.default_rng(3)
    v = rng.normal(size=5000).astype(np.float32)
    x = rng.uniform(0, 1, 5000).astype(np.float32)
    x[:3] = [0.0, 1.0, 1.0]  # samples on the first and on the (closed) last edge
    df = xs.nd_binning(torch.from_numpy(v).cuda(), [torch.from_numpy(x).cuda()], ["x"], list_var_bins=[[0, 0.25, 0.5, 1.0]],
                       statistics=["count", "median"])
    assert list(df.columns) == ["nd", "count", "median", "x"]
    ref = bo.nd_binning(v, [x], [np.array([0, 0.25, 0.5, 1.0], dtype=np.float32)], with_nmad=False)[(0,)]
    assert np.array_equal(df["count"].to_numpy(), ref["count"]) and int(df["count"].sum()) == 5000
    _close(df["median"].to_numpy(), ref["median"], "median")
    # count is always added; even bins with an even number of samples and heavy ties
    v2 = np.repeat(np.array([1.0, 2.0, 2.0, 7.0], dtype=np.float32), 50)
    x2 = np.tile(np.array([0.1, 0.6], dtype=np.float32), 100)
    df2 = xs.nd_binning(v2, [x2], ["x"], list_var_bins=2, statistics=[np.nanmedian, xs.nmad])
    ref2 = bo.nd_binning(v2, [x2], [2])[(0,)]
    assert list(df2.columns) == ["nd", "count", "nanmedian", "nmad", "x"]
    _close(df2["nanmedian"].to_numpy(), ref2["median"], "ties median")
    _close(df2["nmad"].to_numpy(), ref2["nmad"], "ties nmad")
    assert xs.nmad(v) == pytest.approx(float(bo.nmad(v)), rel=1e-6)
    with pytest.raises(NotImplementedError):
        xs.nd_binning(v, [x], ["x"], statistics=[lambda a: float(np.ptp(a))])
    # the other built-in statistics of scipy.stats.binned_statistic (xb_bin_moments): strings and NumPy callables
    import scipy.stats

    from xdem_b200 import binning as xbin

    rng = np.random.default_rng(5)
    vm = (1000 + 30 * rng.standard_normal(200_000)).astype(np.float32)
    xm = rng.uniform(-1, 1, vm.size).astype(np.float32)
    ym = rng.uniform(0, 5, vm.size).astype(np.float32)
    vm[::97] = np.nan
    stats = ["mean", "std", "sum", "min", "max", np.nanmean, np.nanstd, np.nanmin, np.nanmax, np.nansum]
    dfm = xs.nd_binning(vm, [xm, ym], ["x", "y"], list_var_bins=(7, 5), statistics=stats)
    ok = np.isfinite(vm)
    d1 = dfm[dfm.nd == 1]
    for var, name, nb in ((xm, "x", 7), (ym, "y", 5)):
        sub = d1[d1[name].notna()]
        edges = xbin.bin_edges(float(var[ok].min()), float(var[ok].max()), nb, np.float32)
        for stat in ("mean", "std", "sum", "min", "max"):
            ref = scipy.stats.binned_statistic(var[ok], vm[ok].astype(np.float64), statistic=stat, bins=edges)[0]
            got = sub[stat].to_numpy()
            assert np.allclose(got, ref, rtol=1e-9 if stat in ("min", "max") else 2e-7, atol=0, equal_nan=True), (name, stat)
            assert np.array_equal(sub["nan" + stat].to_numpy(), got, equal_nan=True)
    d2 = dfm[dfm.nd == 2]
    ex = xbin.bin_edges(float(xm[ok].min()), float(xm[ok].max()), 7, np.float32)
    ey = xbin.bin_edges(float(ym[ok].min()), float(ym[ok].max()), 5, np.float32)
    ref2 = scipy.stats.binned_statistic_2d(xm[ok], ym[ok], vm[ok].astype(np.float64), statistic="mean", bins=[ex, ey])[0]
    assert np.allclose(d2["mean"].to_numpy(), ref2.flatten(), rtol=2e-7, equal_nan=True)
    # empty bins: NaN for mean / std / min / max, 0 for sum
    dfe = xs.nd_binning(np.array([1.0, 2.0], dtype=np.float32), [np.array([0.1, 0.2], dtype=np.float32)], ["x"],
                        list_var_bins=[np.array([0.0, 0.5, 1.0])], statistics=["mean", "sum", "max"])
    assert dfe["count"].tolist() == [2, 0] and np.isnan(dfe["mean"].iloc[1]) and dfe["sum"].iloc[1] == 0.0
    assert np.isnan(dfe["max"].iloc[1]) and dfe["max"].iloc[0] == 2.0


@pytest.mark.gpu
def test_gpu_binning_large_properties() -> None:
    """5e7 samples, 20 x 20 bins (global-memory histograms: 400 bins exceed the shared-memory variant): every sample is
    counted once in each binning; medians of a value that only depends on the bin are that value; shuffling the samples
    changes nothing."""
    import torch

    from xdem_b200 import binning as xb

    g = torch.Generator(device="cuda").manual_seed(8)
    n = 50_000_000
    x = torch.rand(n, generator=g, device="cuda")
    y = torch.rand(n, generator=g, device="cuda")
    ex = np.linspace(0, 1, 21, dtype=np.float32)
    ix = torch.clamp((x * 20).floor(), max=19)
    iy = torch.clamp((y * 20).floor(), max=19)
    # recompute the bin from the float32 edges actually used (x*20 can differ from the edge compare by an ulp)
    ext = torch.from_numpy(ex).cuda()
    ix = torch.clamp(torch.bucketize(x, ext, right=True) - 1, 0, 19).float()
    iy = torch.clamp(torch.bucketize(y, ext, right=True) - 1, 0, 19).float()
    v = ix * 100 + iy + 0.25
    st = xb.binned_robust_stats(v, [x, y], [ex, ex], want_nmad=True)
    assert int(st["count"].sum()) == n
    want = (np.arange(20)[:, None] * 100 + np.arange(20)[None, :] + 0.25).ravel().astype(np.float32)
    assert np.array_equal(st["median"], want) and np.all(st["nmad"] == 0)
    perm = torch.randperm(n, generator=g, device="cuda")
    noise = torch.randn(n, generator=g, device="cuda")
    a = xb.binned_robust_stats(noise, [x, y], [ex, ex], want_nmad=True)
    b = xb.binned_robust_stats(noise[perm], [x[perm], y[perm]], [ex, ex], want_nmad=True)
    for k in ("count", "median", "nmad"):
        assert np.array_equal(a[k], b[k]), k
    assert np.allclose(a["median"], 0.0, atol=0.02) and np.allclose(a["nmad"], 1.0, atol=0.02)  # N(0,1): nmad ~ sigma
