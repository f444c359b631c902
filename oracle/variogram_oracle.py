"""TEST INFRASTRUCTURE ONLY -- restatement of the empirical-variogram arithmetic behind
``xdem.spatialstats.sample_empirical_variogram`` (subsample_method="pdist_point").

PARITY UNPINNED: the arithmetic lives in scikit-gstat (``skgstat.Variogram``; floor-pinned ``scikit-gstat>=1.0.18`` in
the reference's setup.cfg:56) which is neither vendored in /root/reference nor installed here, and none of the
reference's own tests pins ``exp`` values or the all-pairs mode numerically.  What follows restates
the published algorithm of scikit-gstat 1.0.x, anchored on the reference's call sites:

* spatialstats.py:1091   ``skg.Variogram(coordinates=coords, values=values, normalize=False, fit_method=None, **kw)``
* spatialstats.py:1094-5 ``bins, exp = V.get_empirical(); count = V.bin_count``
* spatialstats.py:1413-16 coordinates of a 2-D array: meshgrid(arange(0, nx*gsd, gsd), arange(0, ny*gsd, gsd))
* spatialstats.py:1425-49 maxlag = extent diagonal; default bins = sqrt(2)*gsd*sqrt(2)^k ... maxlag (right edges)
* spatialstats.py:1541   the last bin is dropped; :1544 dtypes f64/f64/f64/i64

scikit-gstat 1.0.x (restated): distances = scipy pdist(coords) (float64, pairs i<j); diff = |v_i - v_j|;
``bins``: an iterable is used verbatim as right edges; "even" = linspace(0, min(maxlag, max(distances)), n_lags+1)[1:];
lag class k <=> bins[k-1] <= d < bins[k] (lower edge 0 for k=0), pairs with d >= bins[-1] are dropped;
matheron(x) = sum(x^2) / (2 len(x)), NaN for an empty class; bin_count[k] = len(x_k).
"""

from __future__ import annotations

import numpy as np
from scipy.spatial.distance import pdist


def grid_coords(shape: tuple[int, int], gsd: float) -> np.ndarray:
    """spatialstats.py:1413-1416 (x along axis 1 of the meshgrid, 'xy' indexing; values flattened row-major)."""
    x, y = np.meshgrid(np.arange(0, shape[0] * gsd, gsd), np.arange(0, shape[1] * gsd, gsd))
    return np.dstack((x.flatten(), y.flatten())).squeeze()


def default_bins(gsd: float, maxlag: float) -> list[float]:
    """spatialstats.py:1439-1449."""
    bins = []
    right = np.sqrt(2) * gsd
    while right < maxlag:
        bins.append(right)
        right *= np.sqrt(2)
    bins.append(maxlag)
    return bins


def even_bins(distances_max: float, n_lags: int, maxlag: float | None) -> np.ndarray:
    """skgstat.binning.even_width_lags."""
    if maxlag is None or maxlag > distances_max:
        maxlag = distances_max
    return np.linspace(0, maxlag, n_lags + 1)[1:]


def _estimate(x: np.ndarray, estimator: str) -> float:
    """skgstat.estimators (1.0.x): matheron, cressie (Cressie-Hawkins), dowd."""
    if x.size == 0:
        return np.nan
    if estimator == "matheron":
        return float(np.sum(x**2) / (2.0 * x.size))
    if estimator == "cressie":
        n = x.size
        return float(np.power((1.0 / n) * np.sum(np.power(x, 0.5)), 4) / (2 * (0.457 + (0.494 / n) + (0.045 / n**2))))
    if estimator == "dowd":
        return float(2.198 * np.nanmedian(x) ** 2 / 2)
    raise NotImplementedError(estimator)


def empirical_variogram(coords: np.ndarray, values: np.ndarray, bin_func: object = "even", n_lags: int = 10,
                        maxlag: float | None = None, estimator: str = "matheron"
                        ) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(bins, exp, count) of skgstat.Variogram(...).get_empirical() / .bin_count.
    O(N^2) memory: use for N <= ~2e4 (oracle/c_oracle.variogram_pairs is the big-N checker)."""
    coords = np.asarray(coords, dtype=np.float64)
    values = np.asarray(values, dtype=np.float64)
    d = pdist(coords)
    iu = np.triu_indices(len(values), k=1)
    diff = np.abs(values[iu[0]] - values[iu[1]])
    if isinstance(bin_func, str):
        if bin_func != "even":
            raise NotImplementedError(bin_func)
        bins = even_bins(d.max(), n_lags, maxlag)
    else:
        bins = np.asarray(list(bin_func), dtype=np.float64)
    groups = np.full(len(d), -1, dtype=np.int64)
    lower = 0.0
    for k, upper in enumerate(bins):
        groups[(d >= lower) & (d < upper)] = k
        lower = upper
    exp = np.full(len(bins), np.nan)
    count = np.zeros(len(bins), dtype=np.int64)
    for k in range(len(bins)):
        x = diff[groups == k]
        count[k] = x.size
        exp[k] = _estimate(x, estimator)
    return bins, exp, count
