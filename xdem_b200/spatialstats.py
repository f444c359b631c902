"""Spatial statistics hot path: drop-in for ``xdem.spatialstats.sample_empirical_variogram`` (spatialstats.py:1295-1546)
for the all-pairs ("pdist_point") sampling mode, with the O(N^2) pairwise distance / squared-difference / lag-binning
work -- done inside scikit-gstat's ``Variogram`` in the reference (spatialstats.py:1064-1101) -- on the GPU.

Glue kept from the reference (same arithmetic, same quirks): coordinates of a 2-D array (:1413-1416), default ``maxlag``
= extent diagonal (:1425-1431), default right bin edges sqrt(2)*gsd*sqrt(2)^k ... maxlag (:1439-1449), child random
states (:1469-1478), aggregation over runs (:1512-1527), last bin dropped (:1541), output dtypes (:1544).
"""

from __future__ import annotations

import ctypes
import math
import os
from fractions import Fraction
from typing import Any, Iterable

import numpy as np
import pandas as pd
import torch

from . import _arrays, _lib
from .binning import nd_binning, nmad  # noqa: F401  (spatialstats.py:76-216 of the reference)

_METHODS = ["cdist_equidistant", "cdist_point", "pdist_point", "pdist_disk", "pdist_ring"]


# ---------------------------------------------------------------------------------------------------------------
# host-side helpers (exact integer thresholds, Morton order, work units)
# ---------------------------------------------------------------------------------------------------------------


def _float_distance(d2: int, gsd: float) -> float:
    """float64 distance the reference obtains for a squared pixel distance d2 (scipy pdist on coords = index*gsd)."""
    return math.sqrt(gsd * gsd * d2)


#: Which side of a lag class is closed.  "left" = edges[k-1] <= d < edges[k], the rule restated from scikit-gstat 1.0.x
#: (``Variogram._calc_groups``); "right" = edges[k-1] < d <= edges[k].  scikit-gstat is absent from the reference tree and
#: from this image, so the rule is UNPINNED: the two only differ for pairs that sit exactly on a bin
#: edge, both are implemented (the rule only changes the integer thresholds built on the host) and pinned by hand-checkable
#: golden cases (tests/golden/variogram_edges.json), and this one flag -- or XDEM_B200_LAG_EDGE_RULE -- selects it.
LAG_EDGE_RULE = os.environ.get("XDEM_B200_LAG_EDGE_RULE", "left")


def edge_thresholds(edges: Iterable[float], gsd: float, rule: str | None = None) -> list[int]:
    """Integer thresholds T_k on the squared pixel distance such that a pair belongs to lag class k  <=>
    T_{k-1} <= d2 < T_k, reproducing the reference's float64 comparison of d = sqrt(gsd^2 d2) with the edges:
    rule "left":  d < edge_k  <=>  d2 < T_k,  T_k = smallest integer d2 whose float64 distance is >= edge_k;
    rule "right": d <= edge_k <=>  d2 < T_k,  T_k = smallest integer d2 whose float64 distance is >  edge_k.
    Exact whenever index*gsd and its squares are exact in float64 (integer or dyadic gsd); for other gsd the
    reference's per-pair rounding of the coordinates can move pairs that sit on an edge to the neighbouring class
    (see ``on_edge_pair_bound``)."""
    rule = rule or LAG_EDGE_RULE
    if rule not in ("left", "right"):
        raise ValueError(f"lag edge rule must be 'left' or 'right', got {rule!r}")
    out = []
    for e in edges:
        e = float(e)
        if not e > 0:
            out.append(0)
            continue
        c = int(math.ceil(Fraction(e) ** 2 / Fraction(gsd) ** 2))
        if rule == "left":
            while c > 0 and _float_distance(c - 1, gsd) >= e:
                c -= 1
            while _float_distance(c, gsd) < e:
                c += 1
        else:
            while c > 0 and _float_distance(c - 1, gsd) > e:
                c -= 1
            while _float_distance(c, gsd) <= e:
                c += 1
        out.append(min(c, (1 << 63) - 1))
    return out


def on_edge_d2(edges: Iterable[float], gsd: float, ulps: int = 4) -> list[int]:
    """Integer squared pixel distances whose float64 distance lies within ``ulps`` ulp of a bin edge: the only pairs
    whose class can depend on how the reference rounds the individual coordinates (non-dyadic gsd) -- with the default
    sqrt(2)-geometric edges these are the lattice distances d2 = 2^k."""
    out = []
    for e in edges:
        e = float(e)
        if not e > 0:
            continue
        c0 = int(round((e / gsd) ** 2))
        for c in range(max(c0 - 1, 0), c0 + 2):
            if abs(_float_distance(c, gsd) - e) <= ulps * math.ulp(e):
                out.append(c)
    return sorted(set(out))


def _spread_bits16(v: torch.Tensor) -> torch.Tensor:
    v = v.to(torch.int64) & 0xFFFFFFFF
    v = (v | (v << 16)) & 0x0000FFFF0000FFFF
    v = (v | (v << 8)) & 0x00FF00FF00FF00FF
    v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0F
    v = (v | (v << 2)) & 0x3333333333333333
    v = (v | (v << 1)) & 0x5555555555555555
    return v


def _prepare_groups(x: torch.Tensor, y: torch.Tensor, v: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor, int]:
    """Sort samples along a Morton curve, pad to whole groups, build the packed point array and the group boxes."""
    G = int(_lib.lib().xb_variogram_group_size())
    n = x.numel()
    order = torch.argsort(_spread_bits16(x) | (_spread_bits16(y) << 1))
    x, y, v = x[order], y[order], v[order]
    n_groups = (n + G - 1) // G
    n_pad = n_groups * G
    pts = torch.zeros((n_pad, 4), dtype=torch.int32, device=x.device)
    pts[:n, 0] = x.to(torch.int32)
    pts[:n, 1] = y.to(torch.int32)
    pts[:n, 2] = v.to(torch.float32).view(torch.int32)
    pts[:n, 3] = torch.arange(n, dtype=torch.int32, device=x.device)
    pts[n:, 3] = -1
    big = torch.iinfo(torch.int32).max
    valid = (pts[:, 3] >= 0).view(n_groups, G)
    px = pts[:, 0].view(n_groups, G)
    py = pts[:, 1].view(n_groups, G)
    gbox = torch.stack([
        torch.where(valid, px, torch.full_like(px, big)).amin(1),
        torch.where(valid, py, torch.full_like(py, big)).amin(1),
        torch.where(valid, px, torch.full_like(px, -big)).amax(1),
        torch.where(valid, py, torch.full_like(py, -big)).amax(1),
    ], dim=1).contiguous()
    return pts.contiguous(), gbox, n_groups


def _unit_prefix(n_groups: int) -> np.ndarray:
    chunk = int(_lib.lib().xb_variogram_chunk())
    i = np.arange(n_groups, dtype=np.int64)
    units = (n_groups - i + chunk - 1) // chunk
    return np.concatenate([[0], np.cumsum(units)]).astype(np.int64)


def pairwise_lag_binning(x: torch.Tensor, y: torch.Tensor, v: torch.Tensor, edges: np.ndarray | None, gsd: float,
                         n_lags: int | None = None, maxlag: float | None = None, group: Any = None,
                         estimator: str = "matheron", distributed: bool = False, edge_rule: str | None = None
                         ) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """All-pairs lag binning of N grid samples (integer pixel coordinates x, y; float32 values v) on the GPU.

    ``edges`` = float64 right bin edges, or None for skgstat's "even" binning with ``n_lags`` classes over
    [0, min(maxlag, largest sampled distance)].  Returns (edges, count int64, third) where ``third`` is, per class,
    sum (v_i-v_j)^2 for "matheron", sum |v_i-v_j|^0.5 for "cressie", or the exact median of |v_i-v_j| for "dowd"
    (4 radix-select passes over all pairs).
    With ``distributed=True`` (every rank of ``group`` calls with the SAME samples) the work units are split across
    the ranks and the per-class results all-reduced (a few hundred bytes over NCCL)."""
    import torch.distributed as dist

    L = _lib.lib()
    dev = x.device
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    pts, gbox, n_groups = _prepare_groups(x, y, v)
    span = int(max((x.max() - x.min()).item(), (y.max() - y.min()).item()))
    wide = 1 if 2 * span * span >= (1 << 32) - 1 else 0
    with torch.cuda.device(dev):
        if edges is None:
            # seed the max-distance search with the extreme samples along 8 directions (a valid lower bound)
            xf, yf = x.to(torch.float64), y.to(torch.float64)
            cand = []
            for ax, ay in ((1, 0), (0, 1), (1, 1), (1, -1)):
                p = ax * xf + ay * yf
                cand += [int(torch.argmax(p)), int(torch.argmin(p))]
            cx, cy = x[cand].to(torch.int64), y[cand].to(torch.int64)
            d2c = (cx[:, None] - cx[None, :]) ** 2 + (cy[:, None] - cy[None, :]) ** 2
            best = torch.tensor([int(d2c.max())], dtype=torch.int64, device=dev)
            _lib.check(L.xb_variogram_maxd2(pts.data_ptr(), gbox.data_ptr(), n_groups, best.data_ptr(), stream))
            dmax = _float_distance(int(best.item()), gsd)
            # skgstat.binning.even_width_lags
            top = dmax if (maxlag is None or maxlag > dmax) else maxlag
            edges = np.linspace(0, top, int(n_lags) + 1)[1:]
        edges = np.asarray(edges, dtype=np.float64)
        e2 = torch.tensor(edge_thresholds(edges, gsd, edge_rule), dtype=torch.int64, device=dev)
        if not bool((e2[1:] >= e2[:-1]).all()):
            raise ValueError("bin edges must be ascending")
        prefix_np = _unit_prefix(n_groups)
        prefix = torch.from_numpy(prefix_np).to(dev)
        n_units = int(prefix_np[-1])
        rank, world = 0, 1
        if distributed and dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(group), dist.get_world_size(group)
        u0 = n_units * rank // world
        u1 = n_units * (rank + 1) // world
        count = torch.zeros(len(edges), dtype=torch.int64, device=dev)
        sumsq = torch.zeros(len(edges), dtype=torch.float64, device=dev)
        if estimator not in ("matheron", "cressie", "dowd"):
            raise NotImplementedError(f"estimator '{estimator}'")
        _lib.check(L.xb_variogram_pairs(pts.data_ptr(), gbox.data_ptr(), n_groups, e2.data_ptr(), len(edges),
                                        prefix.data_ptr(), u0, u1, wide, 1 if estimator == "cressie" else 0,
                                        count.data_ptr(), sumsq.data_ptr(), stream))
        if world > 1:
            dist.all_reduce(count, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(sumsq, op=dist.ReduceOp.SUM, group=group)
        count_h = count.cpu().numpy()
        if estimator != "dowd":
            return edges, count_h, sumsq.cpu().numpy()
        # Dowd: exact per-class median of |diff| by MSD radix select (8 bits per pass) over all pairs
        nb = len(edges)
        pre = np.zeros(nb, dtype=np.uint32)
        mask = 0
        below = np.zeros(nb, dtype=np.int64)
        k_lo = (count_h - 1) // 2
        last = np.zeros(nb, dtype=np.int64)

        def u32(a: np.ndarray) -> torch.Tensor:
            return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint32).view(np.int32).copy()).to(dev)

        for shift in (24, 16, 8, 0):
            hist = torch.zeros(nb * 256, dtype=torch.int64, device=dev)
            _lib.check(L.xb_variogram_median_pass(pts.data_ptr(), gbox.data_ptr(), n_groups, e2.data_ptr(), nb,
                                                  prefix.data_ptr(), u0, u1, 0, u32(pre).data_ptr(), mask, shift,
                                                  hist.data_ptr(), None, stream))
            if world > 1:
                dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
            h = hist.cpu().numpy().reshape(nb, 256)
            cum = np.cumsum(h, axis=1)
            digit = np.array([int(np.searchsorted(cum[g], k_lo[g] - below[g] + 1, side="left")) if count_h[g] > 0
                              else 0 for g in range(nb)], dtype=np.int64)
            digit = np.minimum(digit, 255)
            below += np.where(digit > 0, cum[np.arange(nb), np.maximum(digit - 1, 0)], 0)
            pre = (pre | (digit.astype(np.uint32) << np.uint32(shift))).astype(np.uint32)
            mask |= 255 << shift
            last = h[np.arange(nb), digit]
        lower = pre.view(np.float32).astype(np.float64)
        median = lower.copy()
        need = (count_h % 2 == 0) & (count_h > 0) & (below + last < (count_h // 2 + 1))
        if need.any():
            nxt = u32(np.full(nb, 0xFFFFFFFF, dtype=np.uint32))
            _lib.check(L.xb_variogram_median_pass(pts.data_ptr(), gbox.data_ptr(), n_groups, e2.data_ptr(), nb,
                                                  prefix.data_ptr(), u0, u1, 1, u32(pre).data_ptr(), 0, 0, None,
                                                  nxt.data_ptr(), stream))
            nx = nxt.to(torch.int64) & 0xFFFFFFFF
            if world > 1:
                dist.all_reduce(nx, op=dist.ReduceOp.MIN, group=group)
            upper = nx.cpu().numpy().astype(np.uint32).view(np.float32).astype(np.float64)
            median = np.where(need, 0.5 * (lower + upper), median)
        median = np.where(count_h > 0, median, np.nan)
    return edges, count_h, median


def estimate_from_sums(count: np.ndarray, third: np.ndarray, estimator: str) -> np.ndarray:
    """skgstat.estimators (1.0.x): matheron = sum d^2 / (2n); cressie = (mean |d|^0.5)^4 / (2 (0.457 + 0.494/n +
    0.045/n^2)); dowd = 2.198 median(|d|)^2 / 2.  NaN for empty classes."""
    n = count.astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        if estimator == "matheron":
            exp = third / (2.0 * n)
        elif estimator == "cressie":
            exp = np.power(third / n, 4) / (2.0 * (0.457 + 0.494 / n + 0.045 / n**2))
        else:
            exp = 2.198 * third**2 / 2.0
    return np.where(count > 0, exp, np.nan)


# ---------------------------------------------------------------------------------------------------------------
# public API
# ---------------------------------------------------------------------------------------------------------------


def _draw_valid_subsample(values_flat: torch.Tensor, subsample: int, rng: np.random.Generator) -> torch.Tensor:
    """Uniform subsample (without replacement) of the finite cells; stands in for geoutils.subsample_array
    (spatialstats.py:978) whose RNG stream is third-party."""
    total = values_flat.numel()
    n_valid = int(torch.isfinite(values_flat).sum().item())
    if n_valid == 0:
        raise ValueError("No valid (finite) values to sample.")
    if subsample >= n_valid:
        return torch.nonzero(torch.isfinite(values_flat)).flatten()
    if n_valid == total and subsample > 0.2 * total:
        idx = rng.choice(total, size=subsample, replace=False)
        return torch.from_numpy(np.asarray(idx, dtype=np.int64)).to(values_flat.device)
    picked = np.empty(0, dtype=np.int64)
    while picked.size < subsample:
        need = subsample - picked.size
        draw = rng.integers(0, total, size=int(need * 1.3 * total / n_valid) + 16, dtype=np.int64)
        cand = np.concatenate([picked, draw])
        _, first = np.unique(cand, return_index=True)
        cand = cand[np.sort(first)]  # keep draw order, drop repeats
        t = torch.from_numpy(cand).to(values_flat.device)
        ok = torch.isfinite(values_flat[t]).cpu().numpy()
        picked = cand[ok]
    return torch.from_numpy(picked[:subsample]).to(values_flat.device)


def sample_empirical_variogram(
    values: Any,
    gsd: float | None = None,
    coords: Any = None,
    subsample: int = 1000,
    subsample_method: str = "cdist_equidistant",
    n_variograms: int = 1,
    n_jobs: int = 1,
    random_state: int | np.random.Generator | None = None,
    **kwargs: Any,
) -> pd.DataFrame:
    """Empirical variogram (``exp``, ``lags``, ``count``, ``err_exp``) -- same signature and output frame as
    ``xdem.spatialstats.sample_empirical_variogram``.

    GPU path: ``subsample_method="pdist_point"`` on a 2-D array / Raster-like / CUDA tensor with ``gsd`` (every pair of
    the random subsample is compared).  Supported skgstat keywords: ``estimator`` in {"matheron" (default), "cressie",
    "dowd"} (xDEM's uncertainty pipeline passes "dowd", spatialstats.py:1810), ``bin_func`` = iterable of right edges
    (default: the reference's sqrt(2)-geometric edges) or ``"even"`` with ``n_lags``, ``maxlag``.
    The disk / ring / equidistant samplers (scikit-gstat metric spaces) raise NotImplementedError.
    """
    if _arrays.is_raster_like(values):
        gsd = values.res[0]
        values = values.data
    if isinstance(values, np.ma.MaskedArray):
        values = _arrays.to_host_nan_array(values)
    if not isinstance(values, (np.ndarray, torch.Tensor)):
        raise ValueError("Values must be of type NDArrayf, np.ma.masked_array or Raster subclass.")
    values = values.squeeze()

    # spatialstats.py:1376-1393
    if (gsd is not None or subsample_method in ["cdist_equidistant", "pdist_disk", "pdist_ring"]) and values.ndim == 1:
        raise ValueError(
            'Values array must be 2D when using any of the "cdist_equidistant", "pdist_disk" and '
            '"pdist_ring" methods, or providing a ground sampling distance instead of coordinates.'
        )
    elif coords is not None and values.ndim != 1:
        raise ValueError("Values array must be 1D when providing coordinates.")
    elif coords is not None and (coords.shape[0] != 2 and coords.shape[1] != 2):
        raise ValueError("The coordinates array must have one dimension with length equal to 2")
    elif values.ndim == 2 and gsd is None:
        raise ValueError("The ground sampling distance must be defined when passing a 2D values array.")
    if subsample_method not in _METHODS:
        raise TypeError(
            'The subsampling method must be one of "cdist_equidistant, "cdist_point", "pdist_point", '
            '"pdist_disk" or "pdist_ring".'
        )
    if subsample_method != "pdist_point":
        raise NotImplementedError(
            f"subsample_method='{subsample_method}' relies on scikit-gstat's metric-space samplers and is not on the "
            "B200 hot path yet (SURVEY.md section 8f rank 2); use subsample_method='pdist_point'."
        )
    if coords is not None:
        raise NotImplementedError("the B200 variogram path takes a 2-D array + gsd (grid samples)")
    estimator = kwargs.get("estimator", "matheron")
    if estimator not in ("matheron", "cressie", "dowd"):
        raise NotImplementedError(f"estimator='{estimator}' is not on the B200 hot path (matheron, cressie, dowd)")
    unknown = set(kwargs) - {"estimator", "bin_func", "n_lags", "maxlag"}
    if unknown:
        raise NotImplementedError(f"unsupported skgstat keyword(s) for the B200 variogram path: {sorted(unknown)}")

    dev = _arrays.require_cuda()
    gsd = float(gsd)
    t = values if isinstance(values, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(values))
    t = t.to(device=dev, dtype=torch.float32)
    nx, ny = t.shape
    flat = t.reshape(-1)

    # coordinates of flat sample k: (k % nx, k // nx) in pixels -- np.meshgrid(x, y) with x over shape[0] and y over
    # shape[1] flattened row-major (spatialstats.py:1413-1416; identical to (col, row) for square arrays)
    xs_f = np.arange(0, nx * gsd, gsd)
    ys_f = np.arange(0, ny * gsd, gsd)
    if "maxlag" not in kwargs:
        maxlag = float(np.sqrt((np.max(xs_f) - np.min(xs_f)) ** 2 + (np.max(ys_f) - np.min(ys_f)) ** 2))
    else:
        maxlag = float(kwargs["maxlag"])
    bin_func = kwargs.get("bin_func", None)
    n_lags = int(kwargs.get("n_lags", 10))
    if bin_func is None:
        edges_in: np.ndarray | None = []  # type: ignore
        right = np.sqrt(2) * gsd
        while right < maxlag:
            edges_in.append(right)  # type: ignore
            right *= np.sqrt(2)
        edges_in.append(maxlag)  # type: ignore
        edges_in = np.asarray(edges_in, dtype=np.float64)
    elif isinstance(bin_func, str):
        if bin_func != "even":
            raise NotImplementedError(f"bin_func='{bin_func}' is not supported on the B200 path (use 'even' or edges)")
        edges_in = None
    else:
        edges_in = np.asarray(list(bin_func), dtype=np.float64)

    # child random states (spatialstats.py:1469-1478)
    if random_state is not None:
        rng = np.random.default_rng(random_state)
        list_random_state = list(rng.choice(n_variograms, n_variograms, replace=False))
    else:
        list_random_state = [None for _ in range(n_variograms)]

    runs = []
    for i in range(n_variograms):
        run_rng = np.random.default_rng(list_random_state[i])
        idx = _draw_valid_subsample(flat, int(subsample), run_rng)
        x = idx % nx
        y = idx // nx
        v = flat[idx]
        edges, count, third = pairwise_lag_binning(x, y, v, edges_in, gsd, n_lags=n_lags, maxlag=maxlag,
                                                   estimator=estimator)
        exp = estimate_from_sums(count, third, estimator)
        runs.append(pd.DataFrame().assign(exp=exp, bins=edges, count=count))

    df = pd.concat(runs)
    if n_variograms == 1:
        df = df.rename(columns={"bins": "lags"})
        df["err_exp"] = np.nan
    else:
        grouped = df.groupby("bins", dropna=False)
        df_mean = grouped[["exp"]].mean()
        df_std = grouped[["exp"]].std()
        df_count = grouped[["count"]].sum()
        df_mean["lags"] = df_mean.index.values
        df_mean["err_exp"] = df_std["exp"] / np.sqrt(n_variograms)
        df_mean["count"] = df_count["count"]
        df = df_mean
    df.drop(df.tail(1).index, inplace=True)  # spatialstats.py:1541
    df = df.astype({"exp": "float64", "err_exp": "float64", "lags": "float64", "count": "int64"})
    return df
