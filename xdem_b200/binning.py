"""N-D binned robust statistics on the B200: mirror of ``xdem.spatialstats.nd_binning`` (spatialstats.py:91-216).

For every explanatory variable (1-D), every pair (2-D) and -- with more than two variables -- all of them together
(N-D), the samples are binned like ``scipy.stats.binned_statistic[_2d|_dd]`` does and the requested statistics are
computed per bin.  Bin numbers, exact medians (``np.nanmedian``) and NMADs (``geoutils.stats.nmad``) come from the CUDA
kernels of ``csrc/xb_binning.cu`` (digitize + MSD radix select); the host only builds the edge arrays (the rules of
``scipy/stats/_binned_statistic.py:_bin_edges``), picks radix digits from the 256-counter histograms and assembles the
``pandas.DataFrame`` in the reference's layout.  Supported statistics: ``"count"``, ``"median"`` / ``np.nanmedian`` /
``np.median`` and any callable named ``nmad``; anything else raises ``NotImplementedError`` (an arbitrary Python
callable cannot run per bin on the device).
"""

from __future__ import annotations

import ctypes
import itertools
from typing import Any, Callable, Iterable, Sequence

import numpy as np
import torch

from xdem_b200 import _arrays, _lib

__all__ = ["nd_binning", "nmad", "binned_robust_stats", "bin_edges", "radix_select_medians"]

NMAD_FACTOR = 1.4826


def nmad(data: Any, nfact: float = NMAD_FACTOR) -> float:
    """Normalized median absolute deviation of the finite values (``geoutils.stats.nmad`` semantics) on the device."""
    t = _as_f32_device(data).reshape(-1)
    ones = torch.zeros_like(t)
    st = binned_robust_stats(t, [ones], [np.array([-0.5, 0.5])], want_nmad=True, nfact=nfact)
    return float(st["nmad"][0])


def _as_f32_device(a: Any) -> torch.Tensor:
    dev = _arrays.require_cuda()
    if isinstance(a, torch.Tensor):
        t = a
    else:
        arr = np.asarray(a.filled(np.nan) if isinstance(a, np.ma.MaskedArray) else a)
        t = torch.from_numpy(np.ascontiguousarray(arr))
    if t.device.type != "cuda":
        t = t.to(dev)
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    return t.contiguous()


def _u32(a: np.ndarray, dev: torch.device) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint32).view(np.int32).copy()).to(dev)


def _key_to_float(key: np.ndarray) -> np.ndarray:
    key = key.astype(np.uint32)
    neg = (key & np.uint32(0x80000000)) == 0
    return np.where(neg, ~key, key & np.uint32(0x7FFFFFFF)).astype(np.uint32).view(np.float32)


def bin_edges(vmin: float, vmax: float, bins: int | Sequence[float], dtype: Any, rng: tuple[float, float] | None = None
              ) -> np.ndarray:
    """Edges of one variable as SciPy builds them (_binned_statistic.py:_bin_edges): ``linspace(min, max, n + 1)`` in the
    sample dtype for an integer `bins` (a zero-width range is widened by +-0.5), the given array otherwise."""
    if np.isscalar(bins):
        lo, hi = (float(vmin), float(vmax)) if rng is None else (float(rng[0]), float(rng[1]))
        if hi < lo:
            raise ValueError("In range, start must be <= stop")
        if lo == hi:
            lo, hi = lo - 0.5, hi + 0.5
        return np.linspace(lo, hi, int(bins) + 1, dtype=dtype)
    return np.asarray(bins, dtype)


def _edge_decimal(edges: np.ndarray) -> float:
    d = np.diff(edges)
    dmin = d.min()
    if dmin == 0:
        raise ValueError("The smallest edge difference is numerically 0.")
    return float(10.0 ** (int(-np.log10(dmin)) + 6))


def _pick_digit(hist: np.ndarray, rank: np.ndarray, counts: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """Per bin: the digit whose cumulative count first exceeds the wanted rank, and the count below that digit."""
    cum = np.cumsum(hist, axis=1)
    digit = (cum <= rank[:, None]).sum(axis=1)
    digit = np.where(counts > 0, np.minimum(digit, hist.shape[1] - 1), 0)
    idx = np.arange(hist.shape[0])
    below = np.where(digit > 0, cum[idx, np.maximum(digit - 1, 0)], 0)
    return digit.astype(np.int64), below.astype(np.int64)


def radix_select_medians(hist_fn: Callable[[np.ndarray, int, int], np.ndarray],
                         next_fn: Callable[[np.ndarray], np.ndarray], n_bins: int) -> tuple[np.ndarray, np.ndarray]:
    """Host driver of the per-bin exact median (np.nanmedian on float32 data) by MSD radix select, 4 passes of 8 bits.

    ``hist_fn(prefix[n_bins] u32, prefix_mask, shift)`` returns the (n_bins, 256) histogram of digit
    ``(key >> shift) & 255`` over the keys with ``key & prefix_mask == prefix[bin]``; ``next_fn(sel[n_bins] u32)`` returns
    per bin the smallest key strictly above ``sel[bin]`` (0xFFFFFFFF if none).  Both are one kernel launch on the device
    (`xb_bin_hist`, `xb_bin_next`); the CPU tests drive this function with NumPy stand-ins.  Returns (median float32
    [n_bins], NaN for empty bins; counts int64).  The mean of the two middle values of an even count is formed in
    float32, like NumPy does for float32 input."""
    prefix = np.zeros(n_bins, dtype=np.uint32)
    prefix_mask = 0
    below = np.zeros(n_bins, dtype=np.int64)
    counts = k_lo = last = h = digit = None
    for ip, shift in enumerate((24, 16, 8, 0)):
        h = np.asarray(hist_fn(prefix, prefix_mask, shift)).reshape(n_bins, 256)
        if ip == 0:
            counts = h.sum(axis=1)
            k_lo = (counts - 1) // 2
        digit, below_d = _pick_digit(h, k_lo - below, counts)
        below += below_d
        prefix = (prefix | (digit.astype(np.uint32) << np.uint32(shift))).astype(np.uint32)
        prefix_mask |= 255 << shift
        last = h[np.arange(n_bins), digit]
    # prefix = exact key of the lower middle value; `below` keys are smaller, `last` keys are equal to it
    lower = _key_to_float(prefix)
    median = lower.copy()
    even = (counts % 2 == 0) & (counts > 0)
    need_next = even & (below + last < (counts // 2 + 1))  # the upper middle value is a strictly larger key
    # the last histogram holds every key sharing the lower middle value's upper 24 bits: the next occupied digit is
    # the next larger key; only a value that closes its 256-key bucket needs the search pass
    found = np.zeros(n_bins, dtype=bool)
    upper_key = np.zeros(n_bins, dtype=np.uint32)
    for g in np.flatnonzero(need_next):
        nz = np.flatnonzero(h[g, digit[g] + 1:])
        if nz.size:
            upper_key[g] = (prefix[g] & np.uint32(0xFFFFFF00)) | np.uint32(digit[g] + 1 + nz[0])
            found[g] = True
    need_next = need_next & ~found
    if need_next.any():
        nxt = np.asarray(next_fn(prefix)).astype(np.uint32)
        upper_key = np.where(need_next, nxt, upper_key).astype(np.uint32)
        found |= need_next
    if found.any():
        with np.errstate(invalid="ignore", over="ignore"):
            mid = ((lower + _key_to_float(upper_key)) * np.float32(0.5)).astype(np.float32)
        median = np.where(found, mid, median).astype(np.float32)
    median = np.where(counts > 0, median, np.float32(np.nan)).astype(np.float32)
    return median, counts.astype(np.int64)


def _select_medians(keys: torch.Tensor, bins: torch.Tensor, n_bins: int) -> tuple[np.ndarray, np.ndarray]:
    """`radix_select_medians` with the histogram / next-key passes running as CUDA kernels over (keys, bins)."""
    L = _lib.lib()
    dev = keys.device
    n = int(keys.numel())
    stream = torch.cuda.current_stream(dev).cuda_stream

    def hist_fn(prefix: np.ndarray, prefix_mask: int, shift: int) -> np.ndarray:
        hist = torch.zeros(n_bins * 256, dtype=torch.int64, device=dev)
        pre = _u32(prefix, dev)
        _lib.check(L.xb_bin_hist(keys.data_ptr(), bins.data_ptr(), n, n_bins, pre.data_ptr(), prefix_mask, shift,
                                 hist.data_ptr(), stream))
        return hist.cpu().numpy()

    def next_fn(sel: np.ndarray) -> np.ndarray:
        nxt = _u32(np.full(n_bins, 0xFFFFFFFF, dtype=np.uint32), dev)
        sel_t = _u32(sel, dev)
        _lib.check(L.xb_bin_next(keys.data_ptr(), bins.data_ptr(), n, n_bins, sel_t.data_ptr(), nxt.data_ptr(), stream))
        return nxt.cpu().numpy().view(np.uint32)

    with torch.cuda.device(dev):
        return radix_select_medians(hist_fn, next_fn, n_bins)


def binned_robust_stats(values: torch.Tensor, variables: list[torch.Tensor], edges: list[np.ndarray],
                        want_median: bool = True, want_nmad: bool = False, nfact: float = NMAD_FACTOR,
                        want_moments: bool = False) -> dict[str, np.ndarray]:
    """Counts / medians / NMADs of `values` in the (flattened, C-order) bins spanned by 1-3 `variables` and their edge
    arrays.  All tensors are flat float32 CUDA tensors of one length; samples with a non-finite value or variable, or
    outside the edges, are dropped."""
    L = _lib.lib()
    if not 1 <= len(variables) <= 3:
        raise NotImplementedError("the B200 binning kernels take 1 to 3 explanatory variables at a time")
    dev = values.device
    n = int(values.numel())
    n_bins = int(np.prod([len(e) - 1 for e in edges]))
    if n_bins >= 0xFFFF:
        raise NotImplementedError(f"{n_bins} bins exceed the 16-bit bin numbers of the B200 binning kernels")
    out: dict[str, np.ndarray] = {}
    if n == 0:
        out["count"] = np.zeros(n_bins, dtype=np.int64)
        out["median"] = np.full(n_bins, np.nan, dtype=np.float32)
        out["nmad"] = np.full(n_bins, np.nan, dtype=np.float32)
        for k in ("mean", "std", "min", "max"):
            out[k] = np.full(n_bins, np.nan)
        out["sum"] = np.zeros(n_bins)
        return out
    edges64 = np.concatenate([np.asarray(e, dtype=np.float64) for e in edges])
    edges_t = torch.from_numpy(edges64).to(dev)
    n_edges = (ctypes.c_int32 * len(edges))(*[len(e) for e in edges])
    p10 = (ctypes.c_double * len(edges))(*[_edge_decimal(np.asarray(e)) for e in edges])
    var_ptrs = (ctypes.c_void_p * len(variables))(*[v.data_ptr() for v in variables])
    keys = torch.empty(n, dtype=torch.int32, device=dev)
    bins = torch.empty(n, dtype=torch.int16, device=dev)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(L.xb_bin_keys(values.data_ptr(), var_ptrs, len(variables), n, edges_t.data_ptr(), n_edges, p10,
                                 keys.data_ptr(), bins.data_ptr(), stream))
        median, counts = _select_medians(keys, bins, n_bins)
        out["count"] = counts
        out["median"] = median
        if want_nmad:
            center = torch.from_numpy(np.nan_to_num(median, nan=0.0).astype(np.float32)).to(dev)
            _lib.check(L.xb_bin_absdev_keys(values.data_ptr(), bins.data_ptr(), n, n_bins, center.data_ptr(),
                                            keys.data_ptr(), stream))
            mad, _ = _select_medians(keys, bins, n_bins)
            out["nmad"] = (np.float32(nfact) * mad).astype(np.float32)  # float32 * python float stays float32 (NEP 50)
        if want_moments:
            # 'mean' / 'std' / 'sum' / 'min' / 'max' of scipy.stats.binned_statistic: float64 per-bin moments and
            # order-preserving min / max keys in one pass (xb_bin_moments); empty bins give NaN (0 for 'sum')
            mom = torch.zeros(2 * n_bins, dtype=torch.float64, device=dev)
            kmin = torch.full((n_bins,), -1, dtype=torch.int32, device=dev)  # 0xffffffff
            kmax = torch.zeros(n_bins, dtype=torch.int32, device=dev)
            _lib.check(L.xb_bin_moments(values.data_ptr(), bins.data_ptr(), n, n_bins, mom.data_ptr(),
                                        mom.data_ptr() + 8 * n_bins, kmin.data_ptr(), kmax.data_ptr(), stream))
            mom_h = mom.cpu().numpy()
            s1, s2 = mom_h[:n_bins], mom_h[n_bins:]
            c = counts.astype(np.float64)
            with np.errstate(invalid="ignore", divide="ignore"):
                mean = np.where(c > 0, s1 / c, np.nan)
                var = np.where(c > 0, np.maximum(s2 / c - mean * mean, 0.0), np.nan)
            out["sum"] = s1.copy()
            out["mean"] = mean
            out["std"] = np.sqrt(var)  # population standard deviation, as np.std / SciPy's 'std'
            lo = _key_to_float(kmin.cpu().numpy().view(np.uint32)).astype(np.float64)
            hi = _key_to_float(kmax.cpu().numpy().view(np.uint32)).astype(np.float64)
            out["min"] = np.where(c > 0, lo, np.nan)
            out["max"] = np.where(c > 0, hi, np.nan)
    return out


#: statistics computed from the per-bin moments of xb_bin_moments: SciPy's built-in strings and their NumPy callables
_MOMENT_STATS = {"mean": "mean", "nanmean": "mean", "std": "std", "nanstd": "std", "sum": "sum", "nansum": "sum",
                 "min": "min", "nanmin": "min", "amin": "min", "max": "max", "nanmax": "max", "amax": "max"}


def _stat_kind(stat: Any) -> tuple[str, str]:
    """(column name as the reference builds it -- the string, or the callable's __name__, spatialstats.py:160-175 --,
    kernel statistic)."""
    if isinstance(stat, str):
        if stat == "count":
            return "count", "count"
        if stat == "median":
            return "median", "median"
        if stat in ("mean", "std", "sum", "min", "max"):  # scipy.stats.binned_statistic's built-in names
            return stat, stat
        raise NotImplementedError(f"statistic '{stat}' is not available in the B200 binning (count, median, mean, std, "
                                  "sum, min, max, nmad)")
    name = getattr(stat, "__name__", None)
    if stat is np.nanmedian or stat is np.median or name in ("nanmedian", "median"):
        return name or "nanmedian", "median"
    if name == "nmad":
        return "nmad", "nmad"
    if name in _MOMENT_STATS and getattr(stat, "__module__", "").startswith("numpy"):
        return name, _MOMENT_STATS[name]
    raise NotImplementedError(
        f"statistic {name or stat!r} is an arbitrary Python callable; the B200 binning computes count, median, mean, std, "
        "sum, min, max and nmad on the device (no per-bin Python calls)"
    )


def nd_binning(
    values: Any,
    list_var: list[Any],
    list_var_names: list[str],
    list_var_bins: int | tuple[int, ...] | tuple[Any, ...] | None = None,
    statistics: Iterable[str | Callable[[Any], Any]] = ("count", np.nanmedian, nmad),
    list_ranges: list[tuple[float, float]] | None = None,
) -> Any:
    """N-dimensional binning of `values` by the explanatory variables (spatialstats.py:91-216): 1-D per variable, 2-D per
    pair, N-D for more than two variables; one row per bin with the statistics, the bin intervals and ``nd``.

    `list_ranges`, when given, holds one ``(min, max)`` per variable."""
    import pandas as pd

    if list_var_bins is None:
        list_var_bins = (10,) * len(list_var_names)
    elif isinstance(list_var_bins, (int, np.integer)):
        list_var_bins = (int(list_var_bins),) * len(list_var_names)
    if len(list_var) != len(list_var_names) or len(list_var_bins) != len(list_var_names):
        raise ValueError("list_var, list_var_names and list_var_bins must have the same length")

    statistics = list(statistics)
    if "count" not in statistics:
        statistics.insert(0, "count")  # spatialstats.py:135-137
    kinds = [_stat_kind(s) for s in statistics]
    want_nmad = any(k == "nmad" for _, k in kinds)
    want_moments = any(k in ("mean", "std", "sum", "min", "max") for _, k in kinds)

    vals = _as_f32_device(values).reshape(-1)
    vars_t = [_as_f32_device(v).reshape(-1) for v in list_var]
    for v in vars_t:
        if v.numel() != vals.numel():
            raise ValueError("values and explanatory variables must have the same size")
    # rows with any non-finite entry are removed before every binning (spatialstats.py:128-131)
    valid = torch.isfinite(vals)
    for v in vars_t:
        valid &= torch.isfinite(v)
    vals = torch.where(valid, vals, torch.full_like(vals, float("nan")))
    any_valid = bool(valid.any())

    def var_edges(i: int) -> np.ndarray:
        rng = None if list_ranges is None else list_ranges[i]
        if np.isscalar(list_var_bins[i]) and rng is None:
            if not any_valid:
                raise ValueError("no valid sample to derive the bin range from")
            sel = vars_t[i][valid]
            lo, hi = float(sel.min()), float(sel.max())
        else:
            lo = hi = 0.0
        return bin_edges(lo, hi, list_var_bins[i], np.float32, rng)

    all_edges = [var_edges(i) for i in range(len(vars_t))]

    def frame(idx: tuple[int, ...]) -> Any:
        edges = [all_edges[i] for i in idx]
        st = binned_robust_stats(vals, [vars_t[i] for i in idx], edges, want_nmad=want_nmad, want_moments=want_moments)
        df = pd.DataFrame()
        for (name, kind) in kinds:
            col = st[kind]
            df[name] = col.astype(np.float64)  # SciPy returns float64 statistics
        iis = [pd.IntervalIndex.from_breaks(e, closed="left") for e in edges]
        if len(idx) == 1:
            df[list_var_names[idx[0]]] = iis[0]
        elif len(idx) == 2:
            df[list_var_names[idx[0]]] = [a for a in iis[0] for _ in iis[1]]
            df[list_var_names[idx[1]]] = [b for _ in iis[0] for b in iis[1]]
        else:
            grid = np.meshgrid(*iis)  # the reference's (default "xy") meshgrid layout, spatialstats.py:203-205
            for k, i in enumerate(idx):
                df[list_var_names[i]] = grid[k].flatten()
        df.insert(0, "nd", len(idx))
        return df

    frames = [frame((i,)) for i in range(len(vars_t))]
    if len(vars_t) > 1:
        frames += [frame(c) for c in itertools.combinations(range(len(vars_t)), 2)]
    if len(vars_t) > 2:
        frames.append(frame(tuple(range(len(vars_t)))))
    return pd.concat(frames)
