"""Terrain bias correction: drop-in for ``xdem.coreg.TerrainBias`` (coreg/biascorr.py:449-620) in its binning form, the
step SURVEY.md section 8f ranks next after the Nuth-Kaab fit: a terrain attribute of the reference DEM (fused stencil
kernel), the exact per-bin median of dh over that attribute (``nd_binning`` on the device: digitize + radix select) and
the correction ``elev + corr(attribute)`` applied raster-wide (``xb_bin_apply_1d``).

What is mirrored: ``BiasCorr.__init__`` validation (biascorr.py:46-170), ``TerrainBias._fit_rst_rst`` / ``_apply_rst``
(:506-620), ``_bin_or_and_fit_nd`` in its "bin" branch (coreg/base.py:906-1005) and the two ways of applying a 1-D
binning, ``interp_nd_binning`` (linear) and ``get_perbin_nd_binning`` (per_bin) (spatialstats.py:237-530).  Fitting
branches ("fit", "bin_and_fit") optimise a user function on the host in the reference and are not on the B200 path."""

from __future__ import annotations

import ctypes
from typing import Any, Callable, Iterable

import numpy as np
import torch

from . import _arrays, _lib, binning, terrain


class TerrainBias:
    """``xdem.coreg.TerrainBias`` for raster-raster fits with ``fit_or_bin="bin"`` (the class default)."""

    def __init__(self, terrain_attribute: str = "max_curvature", fit_or_bin: str = "bin", fit_func: Any = None,
                 fit_optimizer: Any = None, bin_sizes: int | dict[str, int | Iterable[float]] = 100,
                 bin_statistic: Callable[[Any], Any] = np.nanmedian, bin_apply_method: str = "linear",
                 subsample: float | int = 1.0) -> None:
        if fit_or_bin not in ["fit", "bin", "bin_and_fit"]:
            raise ValueError(f"Argument `fit_or_bin` must be 'bin_and_fit', 'fit' or 'bin', got {fit_or_bin}.")
        if fit_or_bin != "bin":
            raise NotImplementedError("the B200 TerrainBias implements fit_or_bin='bin' (the class default); the "
                                      "fitting variants optimise a user function on the host in the reference")
        if not (isinstance(bin_sizes, int) or (isinstance(bin_sizes, dict) and all(
                isinstance(val, (int, Iterable)) for val in bin_sizes.values()))):
            raise TypeError("Argument `bin_sizes` must be an integer, or a dictionary of integers or iterables, "
                            "got {}.".format(type(bin_sizes)))
        if not callable(bin_statistic):
            raise TypeError("Argument `bin_statistic` must be a function (callable), got {}.".format(type(bin_statistic)))
        if not isinstance(bin_apply_method, str):
            raise TypeError("Argument `bin_apply_method` must be the string 'linear' or 'per_bin', "
                            "got {}.".format(type(bin_apply_method)))
        binning._stat_kind(bin_statistic)  # raises for statistics that are arbitrary Python callables
        self._meta: dict[str, Any] = {
            "inputs": {"fitorbin": {"fit_or_bin": "bin", "bin_sizes": bin_sizes, "bin_statistic": bin_statistic,
                                    "bin_apply_method": bin_apply_method, "bias_var_names": [terrain_attribute],
                                    "nd": 1},
                       "random": {"subsample": subsample},
                       "specific": {"terrain_attribute": terrain_attribute}},
            "outputs": {}}
        self._fit_called = False
        self._is_affine = False
        self._needs_vars = False

    @property
    def meta(self) -> dict[str, Any]:
        return self._meta

    # ------------------------------------------------------------------ helpers
    def _attribute(self, elev: torch.Tensor, transform: Any, bias_vars: dict[str, Any] | None) -> torch.Tensor:
        name = self._meta["inputs"]["specific"]["terrain_attribute"]
        if bias_vars is not None:
            if sorted(bias_vars.keys()) != [name]:
                raise ValueError("The keys of `bias_vars` do not match the `bias_var_names` defined during "
                                 "instantiation or fitting: {}.".format([name]))
            t, _ = _arrays.to_device(bias_vars[name])
            return t.to(torch.float32)
        if name == "elevation":
            return elev
        if transform is None:
            raise ValueError("'transform' must be given if the terrain attribute is derived from the elevation.")
        a, e = (float(transform.a), float(transform.e)) if hasattr(transform, "a") else (float(tuple(transform)[0]),
                                                                                        float(tuple(transform)[4]))
        # biascorr.py:530-534: resolution=(transform[0], abs(transform[4]))
        return terrain.get_terrain_attribute(elev, attribute=name, resolution=(a, abs(e)))

    # ------------------------------------------------------------------ fit / apply
    def fit(self, reference_elev: Any, to_be_aligned_elev: Any, inlier_mask: Any = None,
            bias_vars: dict[str, Any] | None = None, transform: Any = None, crs: Any = None,
            random_state: int | np.random.Generator | None = None, **kwargs: Any) -> "TerrainBias":
        """biascorr.py:506-547 + 180-203 + coreg/base.py:980-1005: attribute of the reference, dh = ref - tba on the
        valid (and subsampled) cells, `nd_binning` with (bin_statistic, "count")."""
        ref, _ = _arrays.to_device(reference_elev)
        tba, _ = _arrays.to_device(to_be_aligned_elev)
        ref, tba = ref.to(torch.float32), tba.to(torch.float32)
        if ref.shape != tba.shape or ref.dim() != 2:
            raise ValueError("reference and to-be-aligned elevations must be 2-D rasters of the same shape")
        attr = self._attribute(ref, transform, bias_vars)
        valid = torch.isfinite(ref) & torch.isfinite(tba) & torch.isfinite(attr)
        if inlier_mask is not None:
            m = inlier_mask if isinstance(inlier_mask, torch.Tensor) else torch.from_numpy(np.asarray(inlier_mask))
            valid &= m.to(ref.device).bool()
        n_valid = int(valid.sum().item())
        if n_valid == 0:
            raise ValueError("There is no valid points common to the input and auxiliary data (bias variables, or "
                             "derivatives required for this method, for example slope, aspect, etc).")
        sub = self._meta["inputs"]["random"]["subsample"]
        if sub is not None and sub != 1.0:
            want = int(sub) if sub > 1 else int(sub * n_valid)
            if want < n_valid:
                rng = np.random.default_rng(random_state)
                idx = torch.nonzero(valid.flatten()).flatten()
                pick = torch.from_numpy(rng.choice(n_valid, size=want, replace=False)).to(ref.device)
                m2 = torch.zeros(valid.numel(), dtype=torch.bool, device=ref.device)
                m2[idx[pick]] = True
                valid = m2.view(valid.shape)
                n_valid = want
        nan = torch.full_like(ref, float("nan"))
        diff = torch.where(valid, ref - tba, nan)
        name = self._meta["inputs"]["specific"]["terrain_attribute"]
        bs = self._meta["inputs"]["fitorbin"]["bin_sizes"]
        bins = np.array(bs[name]) if isinstance(bs, dict) else bs
        bins = int(bins) if np.ndim(bins) == 0 else bins
        df = binning.nd_binning(values=diff, list_var=[torch.where(valid, attr, nan)], list_var_names=[name],
                                list_var_bins=(bins,), statistics=(self._meta["inputs"]["fitorbin"]["bin_statistic"],
                                                                   "count"))
        self._meta["outputs"]["fitorbin"] = {"bin_dataframe": df}
        self._meta["outputs"]["random"] = {"subsample_final": n_valid}
        self._fit_called = True
        return self

    def _tables(self, min_count: int) -> tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
        """(mid-points and statistics of the valid bins, all edges, all statistics) of the 1-D binning."""
        df = self._meta["outputs"]["fitorbin"]["bin_dataframe"]
        name = self._meta["inputs"]["specific"]["terrain_attribute"]
        stat = self._meta["inputs"]["fitorbin"]["bin_statistic"]
        stat_name = stat if isinstance(stat, str) else stat.__name__
        sub = df[df.nd == 1] if "nd" in df.columns else df
        iv = sub[name].values
        left = np.array([i.left for i in iv], dtype=np.float64)
        right = np.array([i.right for i in iv], dtype=np.float64)
        vals = sub[stat_name].values.astype(np.float64).copy()
        cnt = sub["count"].values
        vals_l = vals.copy()
        if min_count is not None:
            vals_l[cnt < min_count] = np.nan  # spatialstats.py:332-333
        ok = np.isfinite(vals_l)
        if not ok.any():
            raise ValueError("Dataframe does not contain any valid statistic values.")
        mids = 0.5 * (left + right)  # pd.IntervalIndex.mid
        order = np.argsort(mids[ok])
        edges = np.concatenate([left, right[-1:]])
        return mids[ok][order], vals_l[ok][order], edges, vals

    def apply(self, elev: Any, bias_vars: dict[str, Any] | None = None, transform: Any = None, crs: Any = None,
              **kwargs: Any) -> Any:
        """biascorr.py:592-620 + 259-310: ``elev + corr(attribute(elev))`` as float32 (base.py:491); the kind of ``elev``
        (NumPy array / torch.cuda tensor) is preserved."""
        if not self._fit_called:
            raise AssertionError(".fit() does not seem to have been called yet")
        t, kind = _arrays.to_device(elev)
        t = t.to(torch.float32).contiguous()
        attr = self._attribute(t, transform, bias_vars).contiguous()
        mids, vals_valid, edges, vals_all = self._tables(kwargs.get("min_count", 0))
        method = self._meta["inputs"]["fitorbin"]["bin_apply_method"]
        if method == "linear":
            x, v, m, mode = mids, vals_valid, len(mids), 0
        else:
            x, v, m, mode = edges, vals_all, len(vals_all), 1
        dev = t.device
        xd = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(dev)
        vd = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).to(dev)
        out = torch.empty_like(t)
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().xb_bin_apply_1d(t.data_ptr(), attr.data_ptr(), t.numel(), xd.data_ptr(), vd.data_ptr(),
                                                  int(m), mode, out.data_ptr(), stream))
        return _arrays.from_device(out, kind)
