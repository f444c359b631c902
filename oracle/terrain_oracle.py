"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference's terrain hot path (checker, never shipped).

Restates, in plain float64 NumPy, what the reference computes in
``xdem/terrain/surfit.py`` and ``xdem/terrain/window.py`` (reference @ 58180c6, v0.2.3).  Every function cites the
reference lines it follows.  It is pinned against the reference itself in two ways (see tests/test_oracle_*.py):

* the committed fixtures under ``tests/golden`` which were produced by running the *unmodified* reference (both its
  "scipy" and "numba" engines) through ``oracle/refload.py`` + ``oracle/make_golden.py``;
* the reference's own known-answer tests (doctests terrain.py:268-279/799-813/1484-1493/1553-1562, Jenness rugosity
  test_window.py:21-36, analytic rugosity :38-68, NaN-propagation test_surfit.py:467-518, test_window.py:194-239).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import
this module.  The product (``xdem_b200``) never does.
"""

from __future__ import annotations

import numpy as np

# ---------------------------------------------------------------------------------------------------------------
# Published stencil tables (Horn 1981; Zevenbergen & Thorne 1987; Florinsky 2009) as listed in surfit.py:61-252.
# They are *convolution* kernels: the reference flips them (surfit.py:966, scipy.ndimage.convolve).
# ---------------------------------------------------------------------------------------------------------------

_K = {
    "h1": [[1, 2, 1], [0, 0, 0], [-1, -2, -1]],
    "h2": [[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]],
    "zt_d": [[0, 1, 0], [0, -2, 0], [0, 1, 0]],
    "zt_e": [[0, 0, 0], [1, -2, 1], [0, 0, 0]],
    "zt_f": [[-1, 0, 1], [0, 0, 0], [1, 0, -1]],
    "zt_g": [[0, 1, 0], [0, 0, 0], [0, -1, 0]],
    "zt_h": [[0, 0, 0], [-1, 0, 1], [0, 0, 0]],
    "fl_r": [[2, -1, -2, -1, 2]] * 5,
    "fl_t": [[2] * 5, [-1] * 5, [-2] * 5, [-1] * 5, [2] * 5],
    "fl_s": [[-4, -2, 0, 2, 4], [-2, -1, 0, 1, 2], [0] * 5, [2, 1, 0, -1, -2], [4, 2, 0, -2, -4]],
    "fl_p": [[31, -44, 0, 44, -31], [-5, -62, 0, 62, 5], [-17, -68, 0, 68, 17], [-5, -62, 0, 62, 5],
             [31, -44, 0, 44, -31]],
    "fl_q": [[-31, 5, 17, 5, -31], [44, 62, 68, 62, 44], [0] * 5, [-44, -62, -68, -62, -44], [31, -5, -17, -5, 31]],
}


def _divider(res: float, coef: str) -> float:
    """surfit.py:278-304."""
    return {
        "zt_d": res**2, "zt_e": res**2, "zt_f": 4 * res**2, "zt_g": 2 * res, "zt_h": 2 * res,
        "h1": 8 * res, "h2": 8 * res,
        "fl_r": 35 * res**2, "fl_t": 35 * res**2, "fl_s": 100 * res**2, "fl_p": 420 * res, "fl_q": 420 * res,
    }[coef]


# derivative -> kernel name, per fit (surfit.py:550-588)
_DERIV = {
    "horn": {"zx": "h2", "zy": "h1"},
    "zevenbergthorne": {"zx": "zt_h", "zy": "zt_g", "zxx": "zt_e", "zyy": "zt_d", "zxy": "zt_f"},
    "florinsky": {"zx": "fl_p", "zy": "fl_q", "zxx": "fl_r", "zyy": "fl_t", "zxy": "fl_s"},
}

SURFACE_ATTRS = ["slope", "aspect", "hillshade", "curvature", "profile_curvature", "tangential_curvature",
                 "planform_curvature", "flowline_curvature", "max_curvature", "min_curvature"]
WINDOW_ATTRS = ["topographic_position_index", "terrain_ruggedness_index", "roughness", "rugosity"]


def _nan_dilation(dem: np.ndarray, w: int) -> np.ndarray:
    """True where the w x w window holds a non-finite cell or leaves the raster (surfit.py:1185-1192 + cval=nan;
    numba path: NaN padding surfit.py:1278-1282)."""
    h = w // 2
    bad = ~np.isfinite(np.pad(dem, h, constant_values=np.nan))
    H, W = dem.shape
    out = np.zeros((H, W), dtype=bool)
    for dr in range(w):
        for dc in range(w):
            out |= bad[dr:dr + H, dc:dc + W]
    return out


def derivative(dem: np.ndarray, res: float, name: str, coef_round: np.dtype | None = None) -> np.ndarray:
    """One derivative coefficient = true convolution with (kernel / divider) in float64 (surfit.py:373-377, 948-969;
    spatialstats.py:2512-2525).  ``coef_round=np.float32`` reproduces scipy.ndimage.convolve's rounding of its output
    to the float32 input dtype."""
    k = np.asarray(_K[name], dtype=np.float64) / _divider(res, name)
    kf = k[::-1, ::-1]  # flipped: effective correlation weights
    w = k.shape[0]
    h = w // 2
    H, W = dem.shape
    pad = np.pad(dem.astype(np.float64), h, constant_values=0.0)  # NaN mask applied separately
    pad[~np.isfinite(pad)] = 0.0
    acc = np.zeros((H, W), dtype=np.float64)
    # same accumulation order as _convolution_numba (m1 outer, m2 inner)
    for m1 in range(w):
        for m2 in range(w):
            if kf[m1, m2] != 0.0:
                acc += pad[m1:m1 + H, m2:m2 + W] * kf[m1, m2]
    if coef_round is not None:
        acc = acc.astype(coef_round).astype(np.float64)
    return acc


def surface_attributes(dem: np.ndarray, resolution: float, attrs: list[str], surface_fit: str = "Florinsky",
                       curv_method: str = "geometric", hillshade_azimuth: float = 315.0,
                       hillshade_altitude: float = 45.0, hillshade_z_factor: float = 1.0,
                       out_dtype: np.dtype = np.float32, coef_round: np.dtype | None = None) -> np.ndarray:
    """Restatement of ``_get_surface_attributes`` (surfit.py:1197-1305): radians, hillshade unclipped,
    shape (n_attr, H, W) in the order of ``attrs``."""
    fit = surface_fit.lower()
    directional = curv_method.lower() == "directional"
    w = 5 if fit == "florinsky" else 3
    need2 = any(a in attrs for a in SURFACE_ATTRS[3:])
    if fit == "horn" and need2:
        raise ValueError("Horn has no curvature coefficients")
    d = {}
    with np.errstate(all="ignore"):
        for key, kname in _DERIV[fit].items():
            if key in ("zx", "zy") or need2:
                d[key] = derivative(dem, resolution, kname, coef_round)
        zx, zy = d["zx"], d["zy"]
        g2 = zx**2 + zy**2
        out = {}
        if any(a in attrs for a in ("slope", "aspect", "hillshade")):
            slope = np.arctan(g2**0.5)  # surfit.py:592
            aspect = (-np.arctan2(-zx, zy)) % (2 * np.pi)  # surfit.py:600
            out["slope"], out["aspect"] = slope, aspect
            if "hillshade" in attrs:  # surfit.py:606-622
                slopemap = np.arctan(np.tan(slope) * hillshade_z_factor) if hillshade_z_factor != 1.0 else slope
                az = np.deg2rad(360 - hillshade_azimuth)
                alt = np.deg2rad(hillshade_altitude)
                out["hillshade"] = 1.5 + 254 * (np.sin(alt) * np.cos(slopemap)
                                                + np.cos(alt) * np.sin(slopemap) * np.sin(az - aspect))
        if need2:
            zxx, zyy, zxy = d["zxx"], d["zyy"], d["zxy"]
            flat0 = g2 == 0.0
            flat_eps = g2 < 10e-15
            if "curvature" in attrs:  # surfit.py:636
                out["curvature"] = -2.0 * (zxx + zyy) * 100
            n1 = zxx * zx**2 + 2 * zxy * zx * zy + zyy * zy**2
            n2 = zxx * zy**2 - 2 * zxy * zx * zy + zyy * zx**2
            n3 = zx * zy * (zxx - zyy) - zxy * (zx**2 - zy**2)
            if "profile_curvature" in attrs:  # surfit.py:644-685
                den = g2 if directional else g2 * np.sqrt((1 + g2) ** 3)
                out["profile_curvature"] = np.where(flat0, 0.0, -n1 / den) * 100
            if "tangential_curvature" in attrs:  # surfit.py:693-734
                den = g2 if directional else g2 * np.sqrt(1 + g2)
                out["tangential_curvature"] = np.where(flat0, 0.0, -n2 / den) * 100
            if "planform_curvature" in attrs:  # surfit.py:742-765
                out["planform_curvature"] = np.where(flat_eps, 0.0, -n2 / np.sqrt(g2**3)) * 100
            if "flowline_curvature" in attrs:  # surfit.py:773-811
                if directional:
                    out["flowline_curvature"] = np.where(flat0, 0.0, n3 / (g2**3) ** 0.5) * 100
                else:
                    out["flowline_curvature"] = np.where(flat_eps, 0.0,
                                                         n3 / ((g2**3) ** 0.5 * (1 + g2) ** 0.5)) * 100
            if "max_curvature" in attrs or "min_curvature" in attrs:  # surfit.py:813-943
                if directional:
                    half = (zxx + zyy) / 2
                    rad = (((zxx - zyy) / 2) ** 2 + zxy**2) ** 0.5
                    out["max_curvature"] = np.where(flat0, 0.0, -(half - rad)) * 100
                    out["min_curvature"] = np.where(flat0, 0.0, -(half + rad)) * 100
                else:
                    mn = (1 + zy**2) * zxx - 2 * zxy * zx * zy + (1 + zx**2) * zyy
                    mean = np.where(flat0, 0.0, -mn / (2 * ((1 + g2) ** 3) ** 0.5))
                    uns = np.where(flat0, 0.0,
                                   ((mn / (2 * ((1 + g2) ** 3) ** 0.5)) ** 2
                                    - (zxx * zyy - zxy**2) / ((1 + g2) ** 2)) ** 0.5)
                    out["max_curvature"] = np.where(flat0, 0.0, mean + uns) * 100
                    out["min_curvature"] = np.where(flat0, 0.0, mean - uns) * 100
    bad = _nan_dilation(dem, w)
    res = np.empty((len(attrs),) + dem.shape, dtype=out_dtype)
    for i, a in enumerate(attrs):
        v = np.asarray(out[a], dtype=np.float64).copy()
        v[bad] = np.nan
        res[i] = v.astype(out_dtype)
    return res


def _windows(dem: np.ndarray, w: int) -> np.ndarray:
    """(w*w, H, W) stack of the NaN-padded window cells in row-major order (window.py:851, mode="constant" cval=nan)."""
    h = w // 2
    H, W = dem.shape
    pad = np.pad(dem, h, constant_values=np.nan)
    return np.stack([pad[dr:dr + H, dc:dc + W] for dr in range(w) for dc in range(w)])


def windowed_indexes(dem: np.ndarray, window_size: int, attrs: list[str], resolution: float = 1.0,
                     tri_method: str = "Riley", out_dtype: np.dtype = np.float32,
                     compute_dtype: np.dtype | None = None) -> np.ndarray:
    """Restatement of ``_get_windowed_indexes`` (window.py:926-1002).  ``compute_dtype=None`` computes in the DEM's own
    dtype, which is what the SciPy engine does (vectorized_filter blocks keep the input dtype, window.py:83-97,
    140-156, 207-222, 275-289, 598-683); pass np.float64 for the float64 "truth"."""
    cd = dem.dtype if compute_dtype is None else np.dtype(compute_dtype)
    z = _windows(dem.astype(cd), window_size)
    n = window_size * window_size
    c = z[n // 2]
    res = np.empty((len(attrs),) + dem.shape, dtype=out_dtype)
    with np.errstate(all="ignore"):
        for i, a in enumerate(attrs):
            if a == "topographic_position_index":  # window.py:216-220
                s = np.sum(z, axis=0, dtype=cd)
                v = c - ((s - c) / cd.type(n - 1)).astype(cd)
            elif a == "terrain_ruggedness_index":
                diff = z - c
                if tri_method.lower() == "riley":  # window.py:94-95
                    v = np.sqrt(np.sum(diff * diff, axis=0, dtype=cd))
                else:  # window.py:150-155
                    v = np.sum(np.abs(diff), axis=0, dtype=cd) / cd.type(n - 1)
            elif a == "roughness":  # window.py:281-287
                v = np.max(z, axis=0) - np.min(z, axis=0)
            elif a == "rugosity":  # window.py:598-683 (3x3 only)
                if window_size != 3:
                    raise ValueError("rugosity is defined on a 3x3 window")
                v = _rugosity(z, cd.type(resolution), cd)
            elif a == "fractal_roughness":  # window.py:402-446
                v = _fractal_roughness(z, window_size)
            else:
                raise ValueError(a)
            v = np.asarray(v)
            if a == "fractal_roughness":
                # only the top-left (w-1) x (w-1) cells of the window are read (window.py:356-358, 431)
                used = z.reshape(window_size, window_size, *z.shape[1:])[: window_size - 1, : window_size - 1]
                v[np.isnan(used).any(axis=(0, 1))] = np.nan
            else:
                v[~np.isfinite(np.sum(z * 0, axis=0))] = np.nan
            res[i] = v.astype(out_dtype)
    return res


def _fractal_roughness(z: np.ndarray, w: int) -> np.ndarray:
    """Taud & Parrot (2005) box counting, vectorized form window.py:402-446 (z is (w*w, H, W) row-major windows)."""
    hw = w // 2
    qs = [q for q in range(1, hw + 1) if hw % q == 0]
    log_q = np.log(np.asarray(qs, dtype=np.int32))
    mx = log_q.mean()
    ss_xx = np.sum(log_q * log_q) - len(qs) * mx * mx
    H, W = z.shape[1:]
    zc = z[(w * w) // 2]
    V = np.clip(z - zc, 0, w).reshape(w, w, H, W)
    ns = []
    for q in qs:
        nq = (w - 1) // q
        blocks = V[: nq * q, : nq * q].reshape(nq, q, nq, q, H, W)
        ns.append(blocks.max(axis=(1, 3)).sum(axis=(0, 1)) / np.int32(q))
    y = np.log(np.stack(ns, axis=-1))
    my = y.mean(axis=-1)
    ss_xy = np.sum(y * log_q, axis=-1) - len(qs) * my * mx
    return -(ss_xy / ss_xx)


def _rugosity(z: np.ndarray, L: float, cd: np.dtype) -> np.ndarray:
    """Jenness (2004) surface-area ratio, window.py:598-683; z is (9,H,W) row-major, centre index 4."""
    zc = z[4]
    nb = [0, 1, 2, 3, 5, 6, 7, 8]
    dz_center = np.stack([zc - z[k] for k in nb])
    dl_center = (np.array([np.sqrt(2), 1, np.sqrt(2), 1, 1, np.sqrt(2), 1, np.sqrt(2)], dtype=cd) * L).astype(cd)
    dz_edges = np.stack([z[0] - z[1], z[1] - z[2], z[6] - z[7], z[7] - z[8],
                         z[0] - z[3], z[3] - z[6], z[2] - z[5], z[5] - z[8]])
    dzs = np.concatenate([dz_center, dz_edges])
    dls = np.concatenate([dl_center, np.full(8, L, dtype=cd)])
    hsl = np.sqrt(dzs * dzs + (dls * dls)[:, None, None]) / cd.type(2)
    tri = [(3, 0, 12), (0, 1, 8), (1, 2, 9), (2, 4, 14), (4, 7, 15), (7, 6, 11), (6, 5, 10), (5, 3, 13)]
    area = 0
    for ia, ib, ic in tri:
        a, b, c = hsl[ia], hsl[ib], hsl[ic]
        s = (a + b + c) / cd.type(2)
        area = area + np.sqrt(s * (s - a) * (s - b) * (s - c))
    return area / (L * L)


def next_fft_length(n: int) -> int:
    """freq.py:33-59: power of two up to 1024, else the next integer whose only prime factors are 2, 3, 5, 7."""
    if n <= 1:
        return 1
    if n <= 1024:
        return int(2 ** np.ceil(np.log2(n)))
    m = int(n)
    while True:
        r = m
        for f in (2, 3, 5, 7):
            while r % f == 0:
                r //= f
        if r == 1:
            return m
        m += 1


def texture_shading(dem: np.ndarray, alpha: float | None = 0.8) -> np.ndarray:
    """Restatement of ``_texture_shading_fft`` (freq.py:62-148): mean-fill of the non-finite cells (:84-94), symmetric
    pad to the FFT size (:96-112), half-spectrum scaled by hypot(fx, fy)**alpha with the DC term set to 0 for alpha > 0
    (:114-131), inverse transform, crop and NaN restore (:133-146).  scipy.fft keeps the input precision (float32
    rasters are transformed in single precision, like the reference)."""
    import scipy.fft as sfft

    if alpha is None:
        alpha = 0.8
    if not 0 <= alpha <= 2:
        raise ValueError(f"Alpha must be between 0 and 2, got {alpha}")
    z = np.array(dem, copy=True)
    ok = np.isfinite(z)
    if not ok.any():
        return np.full_like(z, np.nan)
    if not ok.all():
        z[~ok] = np.nanmean(dem)
    n0, n1 = z.shape
    m0, m1 = next_fft_length(n0), next_fft_length(n1)
    b0, b1 = (m0 - n0) // 2, (m1 - n1) // 2
    zp = np.pad(z, ((b0, m0 - n0 - b0), (b1, m1 - n1 - b1)), mode="symmetric")
    mag = np.hypot(sfft.rfftfreq(m1)[None, :], sfft.fftfreq(m0)[:, None])
    mag[0, 0] = 1.0
    filt = mag ** alpha
    if alpha > 0:
        filt[0, 0] = 0.0
    spec = sfft.rfft2(zp, s=(m0, m1))
    spec *= filt
    back = sfft.irfft2(spec, s=(m0, m1))
    out = back[b0:b0 + n0, b1:b1 + n1].copy()
    out[~ok] = np.nan
    return out


def get_terrain_attribute(dem: np.ndarray, attribute: list[str] | str, resolution: float = 1.0, degrees: bool = True,
                          hillshade_altitude: float = 45.0, hillshade_azimuth: float = 315.0,
                          hillshade_z_factor: float = 1.0, surface_fit: str = "Florinsky",
                          curv_method: str = "geometric", tri_method: str = "Riley", window_size: int = 3,
                          window_size_fractal: int = 13, out_dtype: np.dtype | None = None, coef_round: np.dtype | None = None,
                          window_compute_dtype: np.dtype | None = None,
                          texture_alpha: float = 0.8) -> list[np.ndarray] | np.ndarray:
    """Restatement of ``_get_terrain_attribute`` (terrain.py:528-666) for ndarray input: integer -> float32
    (:560-561), rad2deg in the array dtype (:586-591), hillshade clip (:594-596), request order (:651-658)."""
    single = isinstance(attribute, str)
    attrs = [attribute] if single else list(attribute)
    dem = np.asarray(dem)
    if out_dtype is None:
        out_dtype = np.float32 if np.issubdtype(dem.dtype, np.integer) else dem.dtype
    if np.issubdtype(dem.dtype, np.integer):
        dem = dem.astype(np.float32)
    surf = [a for a in attrs if a in SURFACE_ATTRS]
    win = [a for a in attrs if a in WINDOW_ATTRS]
    frac = [a for a in attrs if a == "fractal_roughness"]
    res = {}
    if surf:
        s = surface_attributes(dem, resolution, surf, surface_fit, curv_method, hillshade_azimuth,
                               hillshade_altitude, hillshade_z_factor, out_dtype, coef_round)
        for i, a in enumerate(surf):
            v = s[i]
            if degrees and a in ("slope", "aspect"):
                v = np.rad2deg(v)
            if a == "hillshade":
                v = np.clip(v, 0, 255)
            res[a] = v
    if win:
        wv = windowed_indexes(dem, window_size, win, resolution, tri_method, out_dtype, window_compute_dtype)
        for i, a in enumerate(win):
            res[a] = wv[i]
    if frac:
        res["fractal_roughness"] = windowed_indexes(dem, window_size_fractal, frac, resolution, tri_method, out_dtype,
                                                    window_compute_dtype)[0]
    if "texture_shading" in attrs:
        res["texture_shading"] = texture_shading(dem, texture_alpha).astype(out_dtype, copy=False)  # terrain.py:641-643
    out = [res[a] for a in attrs]
    return out[0] if single else out
