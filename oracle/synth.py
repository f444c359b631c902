"""TEST INFRASTRUCTURE -- deterministic synthetic inputs shared by the golden generator, the tests and bench.py.

SURVEY.md section 8(d): fractal-like DEM ``1000 + 0.05*cumsum(cumsum(N(0,1)))`` (seed 42), NaN injection (seed 43),
integer-valued variant for the bit-exact tests, shifted pair for Nuth-Kaab (seed 45).
"""

from __future__ import annotations

import numpy as np


def fractal_dem(shape: tuple[int, int], seed: int = 42, dtype=np.float32) -> np.ndarray:
    rng = np.random.default_rng(seed)
    z = 1000.0 + 0.05 * np.cumsum(np.cumsum(rng.normal(size=shape), axis=0), axis=1)
    return z.astype(dtype)


def inject_nans(dem: np.ndarray, frac: float = 0.001, hole: int = 4, seed: int = 43) -> np.ndarray:
    rng = np.random.default_rng(seed)
    out = dem.copy()
    n = max(1, int(frac * out.size))
    idx = rng.choice(out.size, size=n, replace=False)
    out.ravel()[idx] = np.nan
    h, w = out.shape
    r0, c0 = h // 3, w // 2
    out[r0:r0 + hole, c0:c0 + hole] = np.nan
    return out


def integer_dem(shape: tuple[int, int], seed: int = 7, high: int = 3000) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.integers(0, high, size=shape).astype(np.float32)


def smooth_dem(shape: tuple[int, int], seed: int = 45, dtype=np.float32) -> np.ndarray:
    """Smooth, textured DEM (sum of random sinusoids) used for the Nuth-Kaab pair: it can be evaluated at sub-pixel
    shifted coordinates analytically, so the injected shift is known exactly."""
    return _smooth_eval(shape, 0.0, 0.0, seed).astype(dtype)


def _smooth_eval(shape: tuple[int, int], dx: float, dy: float, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    h, w = shape
    rr, cc = np.meshgrid(np.arange(h, dtype=np.float64) + dy, np.arange(w, dtype=np.float64) + dx, indexing="ij")
    z = np.full(shape, 1500.0)
    for _ in range(12):
        kx, ky = rng.uniform(0.01, 0.12, 2) * rng.choice([-1, 1], 2)
        amp = rng.uniform(5, 40)
        ph = rng.uniform(0, 2 * np.pi)
        z += amp * np.sin(kx * cc + ky * rr + ph)
    return z


def nk_pair(shape: tuple[int, int], shift_px: tuple[float, float] = (0.37, -0.61), dz: float = 1.5,
            noise: float = 0.01, seed: int = 45) -> tuple[np.ndarray, np.ndarray]:
    """(ref, tba): tba is ref sampled at columns+shift_px[0], rows+shift_px[1], plus dz and white noise."""
    ref = _smooth_eval(shape, 0.0, 0.0, seed)
    tba = _smooth_eval(shape, shift_px[0], shift_px[1], seed) + dz
    rng = np.random.default_rng(seed + 1)
    tba = tba + rng.normal(scale=noise, size=shape)
    return ref.astype(np.float32), tba.astype(np.float32)
