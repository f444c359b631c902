"""TEST INFRASTRUCTURE ONLY -- NumPy/SciPy restatement of the reference's Nuth & Kaab step (checker).

Follows xdem/coreg/affine.py:102-147 (`_iterate_method`), :358-409 (`_nuth_kaab_bin_fit`), :412-474
(`_nuth_kaab_aux_vars`), :477-536 (`_nuth_kaab_iteration_step`), :539-609 (`nuth_kaab`), xdem/coreg/base.py:653-661 (valid
mask), :1006-1045 (`_bin_or_and_fit_nd`, "bin_and_fit") and xdem/spatialstats.py:141-157 (`nd_binning`, 1-D part).
Pinned against the reference itself: tests/golden/nk_reference.npz holds per-iteration outputs of the *unmodified*
reference iteration code run through oracle/refload.py (oracle/make_golden.py).

PARITY UNPINNED for one third-party piece: the bilinear, NaN-propagating interpolator (geoutils `_interp_points`,
geoutils==0.2.5, not installed) is restated with scipy.ndimage.map_coordinates(order=1, cval=nan) both here and in
refload.py; and geoutils `subsample_array`'s RNG stream is not reproduced (tests use subsample=1 or an explicit mask).
"""

from __future__ import annotations

import numpy as np
import scipy.optimize
from scipy.ndimage import map_coordinates
from scipy.stats import binned_statistic


def aux_vars(ref: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """affine.py:433-438, 578-579."""
    gy, gx = np.gradient(ref)
    slope_tan = np.sqrt(gx**2 + gy**2)
    aspect = np.arctan2(-gx, gy)
    aspect += np.pi
    slope_tan[np.isclose(slope_tan, 0)] = np.nan
    return slope_tan, aspect


def dh_at(ref: np.ndarray, tba: np.ndarray, sub_mask: np.ndarray, dx_px: float, dy_px: float) -> np.ndarray:
    """ref[sub] - bilinear(tba)(row + dy_px, col + dx_px) (affine.py:179-184), float64."""
    rows, cols = np.nonzero(sub_mask)
    interp = map_coordinates(tba.astype(np.float64), [rows + dy_px, cols + dx_px], order=1, mode="constant",
                             cval=np.nan, prefilter=False)
    return ref[sub_mask] - interp


def fit_func(xx: np.ndarray, *p: float) -> np.ndarray:
    return p[0] * np.cos(p[1] - xx) + p[2]


def iteration_step(offsets: tuple[float, float, float], ref: np.ndarray, tba: np.ndarray, sub_mask: np.ndarray,
                   slope_tan: np.ndarray, aspect: np.ndarray, a_e: tuple[float, float], bins: int = 72
                   ) -> tuple[tuple[float, float, float], float, dict]:
    """affine.py:477-536 with `_nuth_kaab_bin_fit` (:358-409) inlined."""
    dh = dh_at(ref, tba, sub_mask, offsets[0] / a_e[0], offsets[1] / a_e[1])
    vshift = np.nanmedian(dh)
    dh = dh - vshift
    ok = np.isfinite(dh)
    if not ok.any():
        raise ValueError("The subsample contains no more valid values.")
    st, asp = slope_tan[sub_mask][ok], aspect[sub_mask][ok]
    dh = dh[ok]
    with np.errstate(divide="ignore", invalid="ignore"):
        y = dh / st
    p0 = (3 * np.nanstd(y) / (2**0.5), 0.0, np.nanmean(y))
    valid = np.isfinite(y) & np.isfinite(asp)  # nd_binning drops non-finite pairs (spatialstats.py:130-132)
    med, edges, _ = binned_statistic(asp[valid], y[valid], statistic=np.nanmedian, bins=bins)
    cnt, _, _ = binned_statistic(asp[valid], y[valid], statistic="count", bins=bins)
    mids = 0.5 * (edges[:-1] + edges[1:])
    good = np.isfinite(med) & np.isfinite(mids)
    res = scipy.optimize.curve_fit(f=fit_func, xdata=mids[good], ydata=med[good], sigma=None, absolute_sigma=True,
                                   p0=p0)
    a, b, _c = res[0]
    east, north = a * np.sin(b), a * np.cos(b)
    new = (offsets[0] + east * abs(a_e[0]), offsets[1] + north * abs(a_e[1]), float(vshift))
    dbg = {"vshift": float(vshift), "median": med, "count": cnt, "p0": p0, "edges": edges}
    return new, float(np.sqrt(east**2 + north**2)), dbg


def nuth_kaab(ref: np.ndarray, tba: np.ndarray, inlier_mask: np.ndarray | None = None,
              a_e: tuple[float, float] = (1.0, -1.0), tolerance: float = 0.001, max_iterations: int = 10,
              bins: int = 72) -> tuple[tuple[float, float, float], int, list]:
    slope_tan, aspect = aux_vars(ref)
    valid = np.isfinite(ref) & np.isfinite(tba) & np.isfinite(slope_tan) & np.isfinite(aspect)
    if inlier_mask is not None:
        valid &= inlier_mask
    offsets = (0.0, 0.0, 0.0)
    hist = []
    for i in range(max_iterations):
        offsets, stat, dbg = iteration_step(offsets, ref, tba, valid, slope_tan, aspect, a_e, bins)
        hist.append((offsets, stat))
        if i > 1 and stat < tolerance:
            break
    return offsets, int(valid.sum()), hist
