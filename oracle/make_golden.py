"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz by running the UNMODIFIED reference (both engines) through
oracle/refload.py.  Run in the build container only (needs /root/reference):  python -m oracle.make_golden

The fixtures pin (a) the NumPy oracle (tests/test_oracle_terrain.py, CPU) and (b) the CUDA path (tests -m gpu).
"""

from __future__ import annotations

import os
import warnings

import numpy as np

from oracle import synth
from oracle.refload import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SURF = ["slope", "aspect", "hillshade", "curvature", "profile_curvature", "tangential_curvature",
        "planform_curvature", "flowline_curvature", "max_curvature", "min_curvature"]
WIN = ["topographic_position_index", "terrain_ruggedness_index", "roughness", "rugosity"]


def terrain_inputs() -> dict[str, np.ndarray]:
    frac = synth.inject_nans(synth.fractal_dem((40, 52)), frac=0.004, hole=3)
    rng = np.random.default_rng(42)
    noise = rng.normal(size=(11, 11)).astype(np.float32)  # like test_surfit.py:413-452 (normal DEM with one NaN)
    noise[4, 6] = np.nan
    integer = synth.integer_dem((24, 30))
    integer[7, 9] = np.nan
    small_int = synth.integer_dem((20, 22), seed=9, high=900)  # sums of squared diffs stay < 2^24: order-free
    return {"fractal": frac, "noise": noise, "integer": integer, "small_int": small_int}


def main() -> None:
    warnings.filterwarnings("ignore")
    ref = load_reference()
    os.makedirs(OUT, exist_ok=True)
    gta = ref.terrain.get_terrain_attribute
    inputs = terrain_inputs()
    store: dict[str, np.ndarray] = {f"in|{k}": v for k, v in inputs.items()}
    res = 5.0
    for name, dem in inputs.items():
        for engine in ("scipy", "numba"):
            for fit in ("Horn", "ZevenbergThorne", "Florinsky"):
                for cm in ("geometric", "directional"):
                    attrs = SURF[:3] if fit == "Horn" else SURF
                    if fit == "Horn" and cm == "directional":
                        continue
                    for deg in (True, False):
                        if not deg and not (name == "fractal" and cm == "geometric"):
                            continue
                        outs = gta(dem, attrs, resolution=res, surface_fit=fit, curv_method=cm, engine=engine,
                                   degrees=deg)
                        for a, o in zip(attrs, outs):
                            if not deg and a not in ("slope", "aspect"):
                                continue
                            store[f"surf|{name}|{engine}|{fit}|{cm}|{'deg' if deg else 'rad'}|{a}"] = o
            # hillshade variants
            if name == "fractal":
                for az, alt, zf in ((45.0, 10.0, 1.0), (315.0, 45.0, 3.0), (200.0, 80.0, 0.5)):
                    for fit in ("Horn", "Florinsky"):
                        o = gta(dem, "hillshade", resolution=res, surface_fit=fit, engine=engine,
                                hillshade_azimuth=az, hillshade_altitude=alt, hillshade_z_factor=zf)
                        store[f"hs|{name}|{engine}|{fit}|{az}|{alt}|{zf}"] = o
            for w in (3, 5):
                for tm in ("Riley", "Wilson"):
                    attrs = WIN if w == 3 else WIN[:3]
                    outs = gta(dem, attrs, resolution=res, window_size=w, tri_method=tm, engine=engine)
                    for a, o in zip(attrs, outs):
                        store[f"win|{name}|{engine}|{w}|{tm}|{a}"] = o
    # generic odd windows and fractal roughness (window.py:926-1002, 317-496)
    for name in ("fractal", "small_int"):
        dem = inputs[name]
        for engine in ("scipy", "numba"):
            for w in (7, 9):
                for tm in ("Riley", "Wilson"):
                    outs = gta(dem, WIN[:3], window_size=w, tri_method=tm, engine=engine)
                    for a, o in zip(WIN[:3], outs):
                        store[f"win|{name}|{engine}|{w}|{tm}|{a}"] = o
            for wf in (13, 7):
                store[f"frac|{name}|{engine}|{wf}"] = gta(dem, "fractal_roughness", window_size_fractal=wf, engine=engine)
    store["frac64|fractal|scipy|13"] = gta(inputs["fractal"].astype(np.float64), "fractal_roughness", engine="scipy")
    # float64 input (out_dtype follows the input dtype, terrain.py:328-332)
    dem64 = inputs["fractal"].astype(np.float64) + 0.123456789
    store["in|fractal64"] = dem64
    for fit in ("ZevenbergThorne", "Florinsky"):
        outs = gta(dem64, SURF, resolution=res, surface_fit=fit, engine="numba")
        for a, o in zip(SURF, outs):
            store[f"surf|fractal64|numba|{fit}|geometric|deg|{a}"] = o
    outs = gta(dem64, WIN, resolution=res, window_size=3, engine="scipy")
    for a, o in zip(WIN, outs):
        store[f"win|fractal64|scipy|3|Riley|{a}"] = o
    np.savez_compressed(os.path.join(OUT, "terrain_reference.npz"), **store)
    print(f"terrain_reference.npz: {len(store)} arrays")
    nk_golden(ref)


def texture_inputs() -> dict[str, np.ndarray]:
    frac = synth.inject_nans(synth.fractal_dem((40, 52)), frac=0.004, hole=3)           # pads to 64 x 64
    wide = synth.fractal_dem((30, 1100), seed=5)                                         # 1100 -> 1120 = 2^5 * 5 * 7
    wide[3, 1000:1004] = np.nan
    flat = np.full((17, 9), 123.5, dtype=np.float32)
    allnan = np.full((6, 5), np.nan, dtype=np.float32)
    return {"fractal": frac, "wide": wide, "fractal64": frac.astype(np.float64) + 0.123456789, "flat": flat,
            "allnan": allnan}


def texture_golden(ref) -> None:  # noqa
    """Texture shading (freq.py:62-148) through the reference's public entry point, several exponents."""
    gta = ref.terrain.get_terrain_attribute
    store: dict[str, np.ndarray] = {}
    for name, dem in texture_inputs().items():
        store[f"in|{name}"] = dem
        for alpha in (0.8, 1.5, 0.0, 2.0) if name in ("fractal", "fractal64") else (0.8,):
            store[f"tex|{name}|{alpha}"] = gta(dem, "texture_shading", texture_alpha=alpha)
    np.savez_compressed(os.path.join(OUT, "texture_reference.npz"), **store)
    print(f"texture_reference.npz: {len(store)} arrays")


def binning_inputs() -> dict[str, np.ndarray]:
    """dh-like values with three explanatory variables (slope-, curvature-, elevation-like), NaNs in each."""
    rng = np.random.default_rng(77)
    n = 20000
    slope = rng.gamma(2.0, 8.0, n).astype(np.float32)
    curv = rng.normal(0, 1.5, n).astype(np.float32)
    elev = rng.uniform(900, 2600, n).astype(np.float32)
    vals = (0.02 * slope * rng.normal(size=n) + 0.3 * curv + rng.normal(0, 0.5, n)).astype(np.float32)
    vals[rng.choice(n, 150, replace=False)] = np.nan
    slope[rng.choice(n, 90, replace=False)] = np.nan
    curv[rng.choice(n, 60, replace=False)] = np.nan
    vals[:40] = np.round(vals[:40])  # ties
    return {"values": vals, "slope": slope, "curv": curv, "elev": elev}


def binning_golden(ref) -> None:  # noqa
    """`nd_binning` of the unmodified reference (spatialstats.py:91-216): default statistics (count, nanmedian, nmad),
    integer bin counts and explicit edges, 1, 2 and 3 variables.  DataFrame columns are stored as arrays."""
    nd = ref.spatialstats.nd_binning
    I = binning_inputs()
    store: dict[str, np.ndarray] = {f"in|{k}": v for k, v in I.items()}

    def put(tag: str, df) -> None:  # noqa
        store[f"{tag}|nd"] = df["nd"].to_numpy().astype(np.int64)
        for col in ("count", "nanmedian", "nmad"):
            store[f"{tag}|{col}"] = df[col].to_numpy().astype(np.float64)
        for name in ("slope", "curv", "elev"):
            if name in df.columns:
                left = np.array([iv.left if hasattr(iv, "left") else np.nan for iv in df[name]], dtype=np.float64)
                right = np.array([iv.right if hasattr(iv, "right") else np.nan for iv in df[name]], dtype=np.float64)
                store[f"{tag}|{name}|left"], store[f"{tag}|{name}|right"] = left, right

    put("one10", nd(I["values"], [I["slope"]], ["slope"], list_var_bins=10))
    put("two", nd(I["values"], [I["slope"], I["curv"]], ["slope", "curv"], list_var_bins=(8, 5)))
    edges = (np.array([0, 5, 10, 20, 40, 90], dtype=np.float32), np.array([-5, -1, 0, 1, 5], dtype=np.float32),
             np.array([800, 1500, 2000, 2700], dtype=np.float32))
    put("three_edges", nd(I["values"], [I["slope"], I["curv"], I["elev"]], ["slope", "curv", "elev"],
                          list_var_bins=edges))
    put("three_int", nd(I["values"], [I["slope"], I["curv"], I["elev"]], ["slope", "curv", "elev"], list_var_bins=4))
    np.savez_compressed(os.path.join(OUT, "binning_reference.npz"), **store)
    print(f"binning_reference.npz: {len(store)} arrays")


def terrainbias_golden(ref) -> None:  # noqa
    """The "bin" branch of TerrainBias (biascorr.py:506-620) with the reference's own nd_binning, interp_nd_binning and
    get_perbin_nd_binning (unmodified, spatialstats.py:91-530) on a synthetic pair with a curvature-dependent bias."""
    S = ref.spatialstats
    dem = synth.fractal_dem((200, 260), seed=7)
    attr = ref.terrain.get_terrain_attribute(dem, "max_curvature", resolution=5.0)
    rng = np.random.default_rng(8)
    tba = (dem - 0.8 * np.tanh(attr / 2.0) - 0.3 + rng.normal(scale=0.05, size=dem.shape)).astype(np.float32)
    tba[60:64, 80:100] = np.nan
    valid = np.isfinite(dem) & np.isfinite(tba) & np.isfinite(attr)
    diff = (dem - tba)[valid]
    store = {"ref": dem, "tba": tba, "attr": attr.astype(np.float32)}
    for tag, bins in (("b100", 100), ("edges", np.array([-30, -4, -2, -1, -0.5, 0, 0.5, 1, 2, 4, 30], dtype=np.float32))):
        df = S.nd_binning(values=diff, list_var=[attr[valid]], list_var_names=["max_curvature"],
                          list_var_bins=bins if np.isscalar(bins) else (bins,), statistics=(np.nanmedian, "count"))
        store[f"{tag}|count"] = df["count"].values.astype(np.int64)
        store[f"{tag}|nanmedian"] = df["nanmedian"].values.astype(np.float64)
        store[f"{tag}|left"] = np.array([i.left for i in df["max_curvature"].values], dtype=np.float64)
        store[f"{tag}|right"] = np.array([i.right for i in df["max_curvature"].values], dtype=np.float64)
        fun = S.interp_nd_binning(df=df, list_var_names=["max_curvature"], statistic=np.nanmedian, min_count=0)
        store[f"{tag}|corr_linear"] = fun((attr.flatten(),)).reshape(attr.shape).astype(np.float64)
        store[f"{tag}|corr_perbin"] = S.get_perbin_nd_binning(df=df, list_var=[attr], list_var_names=["max_curvature"],
                                                              statistic=np.nanmedian).astype(np.float64)
    np.savez_compressed(os.path.join(OUT, "terrainbias_reference.npz"), **store)
    print(f"terrainbias_reference.npz: {len(store)} arrays")


def nk_golden(ref) -> None:  # noqa
    """Per-iteration outputs of the reference's own Nuth-Kaab code (affine.py:102-147, 477-609) on a synthetic pair."""
    import scipy.optimize

    A = ref.coreg_affine
    r, t = synth.nk_pair((240, 300))
    r[50:54, 60:70] = np.nan
    t[100, 100] = np.nan
    t[180:183, 20:25] = np.nan
    transform = ref.Affine(5.0, 0, 1000.0, 0, -5.0, 9000.0)
    inl = np.ones(r.shape, bool)
    inl[200:210, 250:260] = False

    class CRS:
        is_projected = True

    pf = {"fit_or_bin": "bin_and_fit", "fit_optimizer": scipy.optimize.curve_fit, "bin_sizes": 72,
          "bin_statistic": np.nanmedian, "fit_func": None, "nd": 1, "bias_var_names": ["aspect"]}
    pr = {"subsample": 1.0, "random_state": None}
    offs = []
    for n_it in range(1, 7):
        out, nfin = A.nuth_kaab(r.copy(), t.copy(), inl, transform, CRS(), "Area", 0.0, n_it, dict(pf), pr, "z")
        offs.append([float(v) for v in out])
    st, asp = A._nuth_kaab_aux_vars(r, t)
    np.savez_compressed(os.path.join(OUT, "nk_reference.npz"), ref=r, tba=t, inlier=inl,
                        offsets=np.array(offs), n_valid=np.array(nfin), slope_tan=st, aspect=asp,
                        transform=np.array([5.0, 0, 1000.0, 0, -5.0, 9000.0]))
    print("nk_reference.npz:", offs[-1], nfin)


if __name__ == "__main__":
    import sys

    warnings.filterwarnings("ignore")
    if len(sys.argv) > 1 and sys.argv[1] == "texture":
        texture_golden(load_reference())
    elif len(sys.argv) > 1 and sys.argv[1] == "binning":
        binning_golden(load_reference())
    elif len(sys.argv) > 1 and sys.argv[1] == "terrainbias":
        terrainbias_golden(load_reference())
    else:
        main()
        texture_golden(load_reference())
        binning_golden(load_reference())
        terrainbias_golden(load_reference())
