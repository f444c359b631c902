"""Parity helpers shared by the CPU (oracle vs golden) and GPU (CUDA vs golden/oracle) tests.

Float criterion (DESIGN.md "Parity"):  |x - ref| <= RTOL*|ref| + atol(attr),  RTOL = 1e-5 (BASELINE.json north_star),
with exact NaN-mask equality.  atol(attr) = 1e-6 x the attribute's natural scale covers values that are ~0 where a
relative bound is meaningless (the reference's own two engines differ by more than 1e-5 relative there).
"""

from __future__ import annotations

import os

import numpy as np

RTOL = 1e-5
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

ATOL = {
    "slope": 90e-6, "aspect": 360e-6, "hillshade": 255e-6,
    # curvatures are x100 (1/100 m); scale ~ 1 for the synthetic DEMs
    "curvature": 2e-6, "profile_curvature": 2e-6, "tangential_curvature": 2e-6, "planform_curvature": 2e-5,
    "flowline_curvature": 2e-5, "max_curvature": 2e-6, "min_curvature": 2e-6,
    "terrain_ruggedness_index": 1e-6, "roughness": 0.0, "rugosity": 2e-6,
}


def load_golden(name: str = "terrain_reference.npz") -> dict[str, np.ndarray]:
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


def nanmask_equal(a: np.ndarray, b: np.ndarray) -> bool:
    return bool(np.array_equal(np.isnan(a), np.isnan(b)))


def max_violation(x: np.ndarray, ref: np.ndarray, rtol: float, atol: float, period: float | None = None) -> float:
    """max over finite cells of |x-ref| / (rtol|ref| + atol)  (<= 1 means pass)."""
    m = np.isfinite(ref) & np.isfinite(x)
    if not m.any():
        return 0.0
    d = np.abs(x[m].astype(np.float64) - ref[m].astype(np.float64))
    if period is not None:
        d = np.minimum(d, np.abs(period - d))
    bound = rtol * np.abs(ref[m].astype(np.float64)) + atol
    return float(np.max(d / bound))


def violation(x: np.ndarray, ref: np.ndarray, attr: str, degrees: bool = True, rtol: float = RTOL,
              where: np.ndarray | None = None) -> float:
    """Factor by which x violates the unwidened criterion against ref (<= 1 passes)."""
    atol = ATOL.get(attr, 1e-6)
    if attr in ("slope", "aspect") and not degrees:
        atol *= np.pi / 180
    period = (360.0 if degrees else 2 * np.pi) if attr == "aspect" else None
    if where is not None:
        x, ref = np.where(where, x, np.nan), np.where(where, ref, np.nan)
    return max_violation(x, ref, rtol, atol, period)


def assert_attr_close(x: np.ndarray, ref: np.ndarray, attr: str, degrees: bool = True, rtol: float = RTOL,
                      atol_scale: float = 1.0, where: np.ndarray | None = None, msg: str = "") -> None:
    assert x.shape == ref.shape, f"{msg}: shape {x.shape} vs {ref.shape}"
    assert x.dtype == ref.dtype, f"{msg}: dtype {x.dtype} vs {ref.dtype}"
    assert nanmask_equal(x, ref), f"{msg}: NaN masks differ ({np.isnan(x).sum()} vs {np.isnan(ref).sum()})"
    atol = ATOL.get(attr, 1e-6) * atol_scale
    period = None
    if attr in ("slope", "aspect") and not degrees:
        atol *= np.pi / 180
    if attr == "aspect":
        period = 360.0 if degrees else 2 * np.pi
    xx, rr = x, ref
    if where is not None:
        xx = np.where(where, x, np.nan)
        rr = np.where(where, ref, np.nan)
    v = max_violation(xx, rr, rtol, atol, period)
    assert v <= 1.0, f"{msg}: {attr} violates |x-ref| <= {rtol}|ref| + {atol:g} by factor {v:.3g}"
