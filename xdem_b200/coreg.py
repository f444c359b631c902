"""Nuth & Kaab (2011) coregistration hot path: drop-in for ``xdem.coreg.NuthKaab`` / ``xdem.coreg.affine.nuth_kaab``
(affine.py:340-609, 2386-2541) for raster-raster inputs.

What runs on the GPU (csrc/xb_nuthkaab.cu): the auxiliary slope-tangent / aspect rasters (np.gradient semantics,
affine.py:435-438), the zero-slope mask (:578-579), the valid mask (base.py:653-661), every iteration's bilinear dh
(:179-184), ``np.nanmedian(dh)`` (:504), ``dh/slope_tan`` (:381), its mean / std for the initial guess (:384) and the
72-bin ``nanmedian`` over aspect (base.py:1014-1020 -> spatialstats.py:147-149) as exact medians by radix select.
What stays on the host: the 72-point ``scipy.optimize.curve_fit`` of a*cos(b-x)+c (base.py:1038-1045), the iteration
loop and stopping rule (affine.py:102-147).
"""

from __future__ import annotations

import ctypes
import os
import logging
from typing import Any, Callable, Iterable

import numpy as np
import scipy.optimize
import torch

from . import _arrays, _lib

#: bracketed exact selection (csrc/xb_nk_fast.cu) for rasters of >= 2^20 pixels; False forces the exhaustive radix select
NK_FAST = True
NK_SAMPLE = int(os.environ.get("XDEM_B200_NK_SAMPLE", 4_000_000))  # sampled pixels for the selection brackets


def _nuth_kaab_fit_func(xx: np.ndarray, *params: float) -> np.ndarray:
    """y(x) = a * cos(b - x) + c  (affine.py:340-355)."""
    return params[0] * np.cos(params[1] - xx) + params[2]


def _transform_coeffs(transform: Any) -> tuple[float, float]:
    """(a, e) pixel sizes (e < 0 for north-up) from an affine.Affine-like object or a 6/9-tuple (a,b,c,d,e,f)."""
    if transform is None:
        return 1.0, -1.0
    if hasattr(transform, "a") and hasattr(transform, "e"):
        return float(transform.a), float(transform.e)
    t = tuple(transform)
    return float(t[0]), float(t[4])


def _ordered_to_float(key: np.ndarray) -> np.ndarray:
    """inverse of the kernel's order-preserving float32 -> uint32 map."""
    key = key.astype(np.uint32)
    neg = (key & np.uint32(0x80000000)) == 0
    bits = np.where(neg, ~key, key & np.uint32(0x7FFFFFFF)).astype(np.uint32)
    return bits.view(np.float32)


def _u32_tensor(a: np.ndarray, dev: torch.device) -> torch.Tensor:
    """uint32 payload carried in an int32 tensor (bit pattern preserved)."""
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint32).view(np.int32).copy()).to(dev)


class _NKState:
    """Device-resident state of one Nuth-Kaab fit (one GPU / one row shard)."""

    def __init__(self, ref: torch.Tensor, tba: torch.Tensor, inlier_mask: torch.Tensor | None, group: Any = None,
                 ref_halo: tuple[torch.Tensor | None, torch.Tensor | None] | None = None,
                 tba_halo: tuple[torch.Tensor | None, torch.Tensor | None] | None = None):
        """``ref`` / ``tba``: this GPU's rows.  Row-sharded fits (xdem_b200.distributed.sharded_nuth_kaab) pass the
        neighbouring shards' rows: ``ref_halo`` = (1 row above, 1 row below) for np.gradient, ``tba_halo`` = (h rows
        above, h rows below) for the shifted bilinear gather; None at a raster border.  With halos given, the
        histogram / moment / range reductions are all-reduced over ``group``."""
        self.L = _lib.lib()
        self.dev = ref.device
        self.rows, self.cols = ref.shape
        self.group = group
        self.sharded = ref_halo is not None or tba_halo is not None
        self.stream = ctypes.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        n = self.rows * self.cols
        rtop, rbot = ref_halo if ref_halo is not None else (None, None)
        ttop, tbot = tba_halo if tba_halo is not None else (None, None)
        parts = [t for t in (rtop, ref, rbot) if t is not None]
        ref_buf = (torch.cat(parts) if len(parts) > 1 else ref).contiguous()  # no copy without halos
        self._ref_buf = ref_buf
        r0 = 0 if rtop is None else rtop.shape[0]
        self.ref = ref_buf[r0:r0 + self.rows]
        parts = [t for t in (ttop, tba, tbot) if t is not None]
        self.tba_buf = (torch.cat(parts) if len(parts) > 1 else tba).contiguous()
        self.tba_row0 = 0 if ttop is None else ttop.shape[0]
        self.tba_halo_rows = (self.tba_row0, 0 if tbot is None else tbot.shape[0])
        self.tba = self.tba_buf[self.tba_row0:self.tba_row0 + self.rows]
        self.slope_tan = torch.empty((self.rows, self.cols), dtype=torch.float32, device=self.dev)
        self.aspect = torch.empty_like(self.slope_tan)
        # one pass: aux variables + valid mask = inlier & finite(ref, tba, slope_tan, aspect) (base.py:653-661) + its
        # count + the static aspect range with the pixels that attain it (xb_nk_prepare)
        inl = None
        if inlier_mask is not None:
            inl = inlier_mask.to(self.dev)
            inl = (inl if inl.dtype in (torch.bool, torch.uint8) else inl != 0).contiguous()
            if inl.shape != ref.shape:
                raise ValueError("inlier_mask must have the shape of the rasters")
            if inl.dtype == torch.bool:
                inl = inl.view(torch.uint8)
        self.sub_mask = torch.empty((self.rows, self.cols), dtype=torch.uint8, device=self.dev)
        self._n_valid_t = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self.range_cand = torch.zeros(4 + 2 * 64, dtype=torch.int32, device=self.dev)
        with torch.cuda.device(self.dev):
            _lib.check(self.L.xb_nk_prepare(ref_buf.data_ptr(), ref_buf.shape[0], self.cols, ref_buf.stride(0),
                                            int(rtop is None), int(rbot is None), r0, r0 + self.rows,
                                            self.tba.data_ptr(), self.tba_buf.stride(0),
                                            inl.data_ptr() if inl is not None else None, self.slope_tan.data_ptr(),
                                            self.aspect.data_ptr(), self.sub_mask.data_ptr(),
                                            self._n_valid_t.data_ptr(), self.range_cand.data_ptr(), self.stream))
        self._inl_keepalive = inl
        self._valid = None
        self.dh = torch.empty(n, dtype=torch.float32, device=self.dev)
        self.n = n

    @property
    def valid(self) -> torch.Tensor:
        """Boolean valid mask of the fit (before any subsampling), materialised on demand."""
        if self._valid is None:
            self._valid = self.sub_mask.bool()
        return self._valid

    def n_valid(self) -> int:
        return int(self._n_valid_t.item())

    def set_subsample(self, mask_u8: torch.Tensor) -> None:
        """Restrict the fit to a subset of the valid pixels.  The static aspect range of xb_nk_prepare belongs to the
        full valid set, so the per-iteration range falls back to the reduction inside the dh pass."""
        _ = self.valid  # cache the full mask before sub_mask is replaced
        self.sub_mask = mask_u8
        self.range_cand = None

    def set_points(self, idx: torch.Tensor) -> None:
        """Restrict the fit to a list of valid pixels (row-major linear indices, int64) -- the reference's default
        (a random subsample of 5e5 points, affine.py:2405): dh is then evaluated at those points only
        (xb_nk_dh_points) and every later pass works on compact arrays of that length instead of the rasters."""
        self.pt_idx = idx.to(torch.int64).contiguous()
        self.slope_tan = self.slope_tan.view(-1)[self.pt_idx].contiguous()
        self.aspect = self.aspect.view(-1)[self.pt_idx].contiguous()
        self.n = int(self.pt_idx.numel())
        self.dh = torch.empty(self.n, dtype=torch.float32, device=self.dev)
        self.range_cand = None
        self._keys = None

    # ------------------------------------------------------------------ device passes
    def compute_dh(self, dx_px: float, dy_px: float) -> tuple[float, float, int]:
        import torch.distributed as dist

        if self.sharded:
            need = int(np.ceil(abs(dy_px))) + 1
            top, bot = self.tba_halo_rows
            if (top and top < need) or (bot and bot < need):
                raise ValueError(f"row shift of {dy_px:.2f} px exceeds the {min(top or bot, bot or top)}-row halo of the "
                                 "sharded Nuth-Kaab fit; re-run with a larger `halo`")
        mm = _u32_tensor(np.array([0xFFFFFFFF, 0], dtype=np.uint32), self.dev)
        cnt = torch.zeros(1, dtype=torch.int64, device=self.dev)
        with torch.cuda.device(self.dev):
            if getattr(self, "pt_idx", None) is not None:
                _lib.check(self.L.xb_nk_dh_points(self.ref.data_ptr(), self.tba_buf.data_ptr(), self.pt_idx.data_ptr(),
                                                  self.n, self.aspect.data_ptr(), self.rows, self.cols,
                                                  self.ref.stride(0), self.tba_buf.stride(0), self.tba_row0,
                                                  self.tba_buf.shape[0], float(dx_px), float(dy_px), self.dh.data_ptr(),
                                                  mm.data_ptr(), cnt.data_ptr(), self.stream))
            else:
                _lib.check(self.L.xb_nk_dh(self.ref.data_ptr(), self.tba_buf.data_ptr(), self.sub_mask.data_ptr(),
                                           self.aspect.data_ptr(), self.rows, self.cols, self.ref.stride(0),
                                           self.tba_buf.stride(0), self.tba_row0, self.tba_buf.shape[0], float(dx_px),
                                           float(dy_px), self.dh.data_ptr(), mm.data_ptr(), cnt.data_ptr(),
                                           self.stream))
        mm64 = mm.to(torch.int64) & 0xFFFFFFFF
        if self.sharded and dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            lo_t, hi_t = mm64[0:1].clone(), mm64[1:2].clone()
            dist.all_reduce(lo_t, op=dist.ReduceOp.MIN, group=self.group)
            dist.all_reduce(hi_t, op=dist.ReduceOp.MAX, group=self.group)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=self.group)
            mm64 = torch.cat([lo_t, hi_t])
        mmh = mm64.cpu().numpy().astype(np.uint32)
        n_fin = int(cnt.item())
        lo, hi = (float(v) for v in mmh.view(np.float32))
        return lo, hi, n_fin

    def _allreduce(self, t: torch.Tensor, op: Any = None) -> None:
        import torch.distributed as dist

        if self.sharded and dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(t, op=op or dist.ReduceOp.SUM, group=self.group)

    def _allreduce_min_u32(self, t_i32: torch.Tensor) -> np.ndarray:
        """min over ranks of uint32 payloads carried in int32 tensors; returns uint32 host array."""
        import torch.distributed as dist

        v = t_i32.to(torch.int64) & 0xFFFFFFFF
        if self.sharded and dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(v, op=dist.ReduceOp.MIN, group=self.group)
        return v.cpu().numpy().astype(np.uint32)

    @staticmethod
    def _pick_digit(hist: np.ndarray, rank_in_bucket: np.ndarray, counts: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
        """per group: digit whose cumulative count first exceeds the wanted rank, and the count below that digit."""
        cum = np.cumsum(hist, axis=1)
        n_groups, n_digits = hist.shape
        digit = np.array([int(np.searchsorted(cum[g], rank_in_bucket[g] + 1, side="left")) if counts[g] > 0 else 0
                          for g in range(n_groups)], dtype=np.int64)
        digit = np.minimum(digit, n_digits - 1)
        below = np.where(digit > 0, cum[np.arange(n_groups), np.maximum(digit - 1, 0)], 0)
        return digit, below

    # ------------------------------------------------------------------ bracketed selection (csrc/xb_nk_fast.cu)
    def fast_eligible(self, n_bins: int) -> bool:
        """The fast path needs vector-aligned rasters and enough pixels for sampling to pay off."""
        import torch.distributed as dist

        if not NK_FAST or n_bins > 96 or self.cols % 4 or self.ref.stride(0) % 4 or self.tba_buf.stride(0) % 4:
            return False
        if getattr(self, "pt_idx", None) is not None:  # point-list fits never stream the rasters
            return False
        n_global = self.n
        if self.sharded and dist.is_available() and dist.is_initialized():
            t = torch.tensor([self.n], dtype=torch.int64, device=self.dev)
            dist.all_reduce(t, group=self.group)
            n_global = int(t.item())
        self._n_global = n_global
        return n_global >= (1 << 20) and self.rows >= 8

    def _fast_setup(self, n_bins: int) -> None:
        import torch.distributed as dist

        L, dev = self.L, self.dev
        lay = (ctypes.c_int32 * 16)()
        _lib.check(L.xb_nkf_layout(lay))
        names = ["MAXB", "C_SIZE", "K_SIZE", "F_SIZE", "C_NFIN", "C_GBELOW", "C_GNC", "C_BNC", "C_FLAGS", "C_BTOTAL",
                 "C_BBELOW", "K_ASPMIN", "K_GLO", "K_BLO", "F_VSHIFT", "F_MED"]
        self.lay = {k: int(v) for k, v in zip(names, lay)}
        world = (dist.get_world_size(self.group)
                 if (self.sharded and dist.is_available() and dist.is_initialized()) else 1)
        self._world = world
        n_global = self._n_global
        # ~4 M sampled pixels over the whole raster (measured optimum of sample cost vs bracket width at 16384^2): one 4-pixel chunk out of every `stride`, jittered
        self.stride = max(1, int(n_global // NK_SAMPLE))
        n_schunks = (self.rows * (self.cols // 4) + self.stride - 1) // self.stride
        ns = 4 * n_schunks
        if world > 1:
            t = torch.tensor([ns], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            ns = int(t.item())
        self.ns = ns
        m = max(1.0, n_global / self.stride)  # sample size over all ranks
        frac_g = min(1.0, 2.0 * (4.0 * np.sqrt(m) + 8.0) / m)  # bracket = +-(4 sqrt(m) + 8) sample ranks
        mb = max(1.0, m / n_bins)
        frac_b = min(1.0, 2.0 * (4.0 * np.sqrt(mb) + 8.0) / mb)
        self.gcap = int(min(self.n, 3 * frac_g * self.n + 65536))
        self.bcap = int(min(self.n, 2.5 * frac_b * self.n + 262144))
        i32, u8, i64, f64 = torch.int32, torch.uint8, torch.int64, torch.float64
        self.f_sample = torch.zeros(ns, dtype=i32, device=dev)       # keys (0 = empty); reused for the y sample
        self.f_sgrp = torch.full((ns,), 255, dtype=u8, device=dev)
        self.f_gcompact = torch.empty(self.gcap, dtype=i32, device=dev)
        self.f_bkey = torch.empty(self.bcap, dtype=i32, device=dev)
        self.f_bgrp = torch.empty(self.bcap, dtype=u8, device=dev)
        self.f_bins = torch.empty(self.n, dtype=u8, device=dev)
        self.f_cnt = torch.zeros(self.lay["C_SIZE"], dtype=i64, device=dev)
        self.f_keys = torch.zeros(self.lay["K_SIZE"], dtype=i32, device=dev)
        self.f_f64 = torch.full((self.lay["F_SIZE"],), float("nan"), dtype=f64, device=dev)
        maxb = self.lay["MAXB"]
        self.f_hist = torch.zeros(2 * maxb * 256, dtype=i32, device=dev)
        self.f_prefix = torch.zeros(2 * maxb, dtype=i32, device=dev)
        self.f_below = torch.zeros(2 * maxb, dtype=i64, device=dev)
        self.f_rank = torch.zeros(2 * maxb, dtype=i64, device=dev)
        if world > 1:
            self.g_sample = torch.empty(world * ns, dtype=i32, device=dev)
            self.g_sgrp = torch.empty(world * ns, dtype=u8, device=dev)
            self.g_gcompact = torch.empty(world * self.gcap, dtype=i32, device=dev)
            self.g_bkey = torch.empty(world * self.bcap, dtype=i32, device=dev)
            self.g_bgrp = torch.empty(world * self.bcap, dtype=u8, device=dev)
            self.g_counts = torch.zeros(world, dtype=i64, device=dev)
        self._fast_ready = n_bins
        self._fast_iter = 0

    def iteration_fast(self, dx_px: float, dy_px: float, n_bins: int) -> dict[str, Any] | None:
        """One iteration's device work on the bracketed-selection path; returns the statistics the host fit needs, or
        None when a bracket missed / a compact buffer overflowed (the caller then takes the exhaustive path)."""
        import torch.distributed as dist

        if getattr(self, "_fast_ready", None) != n_bins:
            self._fast_setup(n_bins)
        if self.sharded:
            need = int(np.ceil(abs(dy_px))) + 1
            top, bot = self.tba_halo_rows
            if (top and top < need) or (bot and bot < need):
                raise ValueError(f"row shift of {dy_px:.2f} px exceeds the {min(top or bot, bot or top)}-row halo of the "
                                 "sharded Nuth-Kaab fit; re-run with a larger `halo`")
        L, lay, st = self.L, self.lay, self.stream
        world, group = self._world, self.group
        cnt, keys, f64 = self.f_cnt, self.f_keys, self.f_f64
        self._fast_iter += 1
        seed = (self._fast_iter * 0x9E3779B1) & 0xFFFFFFFF
        c8 = cnt.element_size()

        def cptr(i: int) -> int:
            return cnt.data_ptr() + c8 * i

        def kptr(i: int) -> int:
            return keys.data_ptr() + 4 * i

        def fptr(i: int) -> int:
            return f64.data_ptr() + 8 * i

        def dh(sample: int) -> None:
            _lib.check(L.xb_nkf_dh(sample, self.ref.data_ptr(), self.tba_buf.data_ptr(), self.sub_mask.data_ptr(),
                                   self.aspect.data_ptr(), self.rows, self.cols, self.ref.stride(0),
                                   self.tba_buf.stride(0), self.tba_row0, self.tba_buf.shape[0], float(dx_px),
                                   float(dy_px), self.dh.data_ptr(), self.f_sample.data_ptr(), self.stride, seed,
                                   cnt.data_ptr(), keys.data_ptr(), self.f_gcompact.data_ptr(), self.gcap, st))

        def yk(sample: int) -> None:
            _lib.check(L.xb_nkf_y(sample, self.dh.data_ptr(), self.slope_tan.data_ptr(), self.aspect.data_ptr(),
                                  self.f_bins.data_ptr(), self.rows, self.cols, n_bins, self.f_sample.data_ptr(),
                                  self.f_sgrp.data_ptr(), self.stride, seed ^ 0x5BD1E995, cnt.data_ptr(), keys.data_ptr(),
                                  f64.data_ptr(), self.f_bkey.data_ptr(), self.f_bgrp.data_ptr(), self.bcap, st))

        def select(key: torch.Tensor, grp: torch.Tensor | None, n_seg: int, seg_cap: int, seg_count: int | None,
                   G: int, mode: int, ext_total: int | None, ext_below: int | None, out_lo: int | None,
                   out_hi: int | None, out_val: int | None, miss_bit: int) -> None:
            _lib.check(L.xb_nkf_select(key.data_ptr(), grp.data_ptr() if grp is not None else None, n_seg, seg_cap,
                                       seg_count, 1, G, mode, ext_total, ext_below, out_lo, out_hi, out_val,
                                       cptr(lay["C_FLAGS"]), miss_bit, self.f_hist.data_ptr(),
                                       self.f_prefix.data_ptr(), self.f_below.data_ptr(), self.f_rank.data_ptr(), st))

        if world == 1:
            with torch.cuda.device(self.dev):
                _lib.check(L.xb_nkf_iteration(
                    self.ref.data_ptr(), self.tba_buf.data_ptr(), self.sub_mask.data_ptr(), self.slope_tan.data_ptr(),
                    self.aspect.data_ptr(), self.rows, self.cols, self.ref.stride(0), self.tba_buf.stride(0),
                    self.tba_row0, self.tba_buf.shape[0], float(dx_px), float(dy_px), n_bins, self.dh.data_ptr(),
                    self.f_bins.data_ptr(), self.f_sample.data_ptr(), self.f_sgrp.data_ptr(), self.ns, self.stride, seed,
                    self.f_gcompact.data_ptr(), self.gcap, self.f_bkey.data_ptr(), self.f_bgrp.data_ptr(), self.bcap,
                    cnt.data_ptr(), keys.data_ptr(), f64.data_ptr(), self.f_hist.data_ptr(), self.f_prefix.data_ptr(),
                    self.f_below.data_ptr(), self.f_rank.data_ptr(),
                    self.range_cand.data_ptr() if self.range_cand is not None else None, st))
                cnt_h = cnt.cpu().numpy()  # the iteration's only host synchronisation
                f_h = f64.cpu().numpy()
            return self._fast_result(cnt_h, f_h, n_bins)
        with torch.cuda.device(self.dev):
            _lib.check(L.xb_nkf_reset(cnt.data_ptr(), keys.data_ptr(), f64.data_ptr(), self.f_hist.data_ptr(), st))
            # 1-2: sample of dh -> global bracket
            dh(1)
            if world > 1:
                dist.all_gather_into_tensor(self.g_sample, self.f_sample, group=group)
                select(self.g_sample, None, 1, world * self.ns, None, 1, 0, None, None, kptr(lay["K_GLO"]),
                       kptr(lay["K_GLO"] + 1), None, 1)
            else:
                select(self.f_sample, None, 1, self.ns, None, 1, 0, None, None, kptr(lay["K_GLO"]),
                       kptr(lay["K_GLO"] + 1), None, 1)
            # 3: the full dh pass
            dh(0)
            if world > 1:
                dist.all_reduce(cnt[lay["C_NFIN"]:lay["C_GBELOW"] + 1], group=group)
                amin = keys[lay["K_ASPMIN"]:lay["K_ASPMIN"] + 1].to(torch.int64) & 0xFFFFFFFF
                amax = keys[lay["K_ASPMIN"] + 1:lay["K_ASPMIN"] + 2].to(torch.int64) & 0xFFFFFFFF
                dist.all_reduce(amin, op=dist.ReduceOp.MIN, group=group)
                dist.all_reduce(amax, op=dist.ReduceOp.MAX, group=group)
                keys[lay["K_ASPMIN"]] = torch.where(amin >= 2**31, amin - 2**32, amin).to(torch.int32)[0]
                keys[lay["K_ASPMIN"] + 1] = torch.where(amax >= 2**31, amax - 2**32, amax).to(torch.int32)[0]
                dist.all_gather_into_tensor(self.g_gcompact, self.f_gcompact, group=group)
                dist.all_gather_into_tensor(self.g_counts, cnt[lay["C_GNC"]:lay["C_GNC"] + 1], group=group)
            _lib.check(L.xb_nkf_range(keys.data_ptr(), f64.data_ptr(), st))
            # 4: exact median of dh from the compact buffer(s)
            if world > 1:
                select(self.g_gcompact, None, world, self.gcap, self.g_counts.data_ptr(), 1, 1, cptr(lay["C_NFIN"]),
                       cptr(lay["C_GBELOW"]), None, None, fptr(lay["F_VSHIFT"]), 1)
            else:
                select(self.f_gcompact, None, 1, self.gcap, cptr(lay["C_GNC"]), 1, 1, cptr(lay["C_NFIN"]),
                       cptr(lay["C_GBELOW"]), None, None, fptr(lay["F_VSHIFT"]), 1)
            # 5-6: sample of y per aspect bin -> per-bin brackets
            yk(1)
            if world > 1:
                dist.all_gather_into_tensor(self.g_sample, self.f_sample, group=group)
                dist.all_gather_into_tensor(self.g_sgrp, self.f_sgrp, group=group)
                select(self.g_sample, self.g_sgrp, 1, world * self.ns, None, n_bins, 0, None, None, kptr(lay["K_BLO"]),
                       kptr(lay["K_BLO"] + lay["MAXB"]), None, 4)
            else:
                select(self.f_sample, self.f_sgrp, 1, self.ns, None, n_bins, 0, None, None, kptr(lay["K_BLO"]),
                       kptr(lay["K_BLO"] + lay["MAXB"]), None, 4)
            # 7: the full y pass
            yk(0)
            if world > 1:
                dist.all_reduce(cnt[lay["C_BTOTAL"]:lay["C_BBELOW"] + lay["MAXB"]], group=group)
                dist.all_reduce(f64[5:8], group=group)
                dist.all_gather_into_tensor(self.g_bkey, self.f_bkey, group=group)
                dist.all_gather_into_tensor(self.g_bgrp, self.f_bgrp, group=group)
                dist.all_gather_into_tensor(self.g_counts, cnt[lay["C_BNC"]:lay["C_BNC"] + 1], group=group)
                select(self.g_bkey, self.g_bgrp, world, self.bcap, self.g_counts.data_ptr(), n_bins, 2,
                       cptr(lay["C_BTOTAL"]), cptr(lay["C_BBELOW"]), None, None, fptr(lay["F_MED"]), 4)
            else:
                # 8: per-bin exact medians
                select(self.f_bkey, self.f_bgrp, 1, self.bcap, cptr(lay["C_BNC"]), n_bins, 2, cptr(lay["C_BTOTAL"]),
                       cptr(lay["C_BBELOW"]), None, None, fptr(lay["F_MED"]), 4)
            _lib.check(L.xb_nkf_finalize(cnt.data_ptr(), f64.data_ptr(), self.gcap, self.bcap, st))
            if world > 1:
                dist.all_reduce(cnt[lay["C_FLAGS"]:lay["C_FLAGS"] + 1], op=dist.ReduceOp.MAX, group=group)
            cnt_h = cnt.cpu().numpy()  # the iteration's only host synchronisation
            f_h = f64.cpu().numpy()
        return self._fast_result(cnt_h, f_h, n_bins)

    def iteration_points(self, dx_px: float, dy_px: float, n_bins: int) -> dict[str, Any] | None:
        """One iteration of a point-list fit (``set_points``) entirely on the device (xb_nkf_iteration_points): exact
        medians by radix select over ALL points, ranks picked on the device, one host read.  Same result dictionary as
        ``iteration_fast``."""
        L, dev = self.L, self.dev
        if getattr(self, "_points_ready", None) is None:
            lay = (ctypes.c_int32 * 16)()
            _lib.check(L.xb_nkf_layout(lay))
            names = ["MAXB", "C_SIZE", "K_SIZE", "F_SIZE", "C_NFIN", "C_GBELOW", "C_GNC", "C_BNC", "C_FLAGS", "C_BTOTAL",
                     "C_BBELOW", "K_ASPMIN", "K_GLO", "K_BLO", "F_VSHIFT", "F_MED"]
            self.lay = {k: int(v) for k, v in zip(names, lay)}
            i32, u8, i64, f64 = torch.int32, torch.uint8, torch.int64, torch.float64
            maxb = self.lay["MAXB"]
            self.f_cnt = torch.zeros(self.lay["C_SIZE"], dtype=i64, device=dev)
            self.f_keys = torch.zeros(self.lay["K_SIZE"], dtype=i32, device=dev)
            self.f_f64 = torch.full((self.lay["F_SIZE"],), float("nan"), dtype=f64, device=dev)
            self.f_hist = torch.zeros(2 * maxb * 256, dtype=i32, device=dev)
            self.f_prefix = torch.zeros(2 * maxb, dtype=i32, device=dev)
            self.f_below = torch.zeros(2 * maxb, dtype=i64, device=dev)
            self.f_rank = torch.zeros(2 * maxb, dtype=i64, device=dev)
            self.p_key = torch.empty(self.n, dtype=i32, device=dev)
            self.p_grp = torch.empty(self.n, dtype=u8, device=dev)
            self._points_ready = True
        with torch.cuda.device(dev):
            _lib.check(L.xb_nkf_iteration_points(
                self.ref.data_ptr(), self.tba_buf.data_ptr(), self.pt_idx.data_ptr(), self.n, self.slope_tan.data_ptr(),
                self.aspect.data_ptr(), self.rows, self.cols, self.ref.stride(0), self.tba_buf.stride(0), self.tba_row0,
                self.tba_buf.shape[0], float(dx_px), float(dy_px), n_bins, self.dh.data_ptr(), self.p_key.data_ptr(),
                self.p_grp.data_ptr(), self.f_cnt.data_ptr(), self.f_keys.data_ptr(), self.f_f64.data_ptr(),
                self.f_hist.data_ptr(), self.f_prefix.data_ptr(), self.f_below.data_ptr(), self.f_rank.data_ptr(),
                self.stream))
            cnt_h = self.f_cnt.cpu().numpy()  # the iteration's only host synchronisation
            f_h = self.f_f64.cpu().numpy()
        return self._fast_result(cnt_h, f_h, n_bins)

    def _fast_result(self, cnt_h: np.ndarray, f_h: np.ndarray, n_bins: int) -> dict[str, Any] | None:
        lay = self.lay
        self.fast_last_flags = int(cnt_h[lay["C_FLAGS"]])
        if self.fast_last_flags != 0:
            self.fast_fallbacks = getattr(self, "fast_fallbacks", 0) + 1
            logging.info("Nuth-Kaab fast path: flags %d (1 median bracket missed, 2 compact overflow, 4 / 8 the same "
                         "per aspect bin) -> exhaustive selection for this iteration", self.fast_last_flags)
            return None
        return {"n_fin": int(cnt_h[lay["C_NFIN"]]), "vshift": float(f_h[lay["F_VSHIFT"]]), "lo": float(f_h[1]),
                "hi": float(f_h[2]), "moments": f_h[5:8].copy(),
                "median": f_h[lay["F_MED"]:lay["F_MED"] + n_bins].copy(),
                "counts": cnt_h[lay["C_BTOTAL"]:lay["C_BTOTAL"] + n_bins].copy()}

    def select_medians(self, mode: int, vshift: float, lo: float, hi: float, n_groups: int,
                       want_moments: bool = False) -> tuple[np.ndarray, np.ndarray, np.ndarray | None]:
        """Exact medians (np.nanmedian semantics: mean of the two middle values for even counts) of float32 keys by MSD
        radix select.  mode 0: global median of dh (11 + 11 + 10 bit passes, one shared-memory histogram per CTA).
        mode 1: per-aspect-bin medians of y = (dh - vshift)/slope_tan: one pass builds compact (key, bin) pairs, the
        first 8-bit histogram and the moments, three more 8-bit passes stream the 5-byte pairs (n_groups x 256 counters
        per CTA in shared memory).  Returns (median float64 [n_groups], count, moments or None)."""
        L, dev, n = self.L, self.dev, self.n
        moments_t = torch.zeros(3, dtype=torch.float64, device=dev)
        if mode == 0:
            passes = [(21, 2048), (10, 2048), (0, 1024)]
        else:
            passes = [(24, 256), (16, 256), (8, 256), (0, 256)]
            if getattr(self, "_keys", None) is None:
                self._keys = torch.empty(n, dtype=torch.int32, device=dev)
                self._grp = torch.empty(n, dtype=torch.uint8, device=dev)
                self._bin_cache = torch.empty(n, dtype=torch.uint8, device=dev)
                self._bin_range = None
        prefix = np.zeros(n_groups, dtype=np.uint32)
        prefix_mask = 0
        below = np.zeros(n_groups, dtype=np.int64)
        counts = k_lo = last_hist = None
        with torch.cuda.device(dev):
            for ip, (shift, n_digits) in enumerate(passes):
                hist = torch.zeros(n_groups * n_digits, dtype=torch.int64, device=dev)
                if mode == 0:
                    _lib.check(L.xb_nk_hist(self.dh.data_ptr(), n, int(prefix[0]), prefix_mask, shift, n_digits,
                                            hist.data_ptr(), self.stream))
                elif ip == 0:
                    # a pixel's aspect bin only depends on [lo, hi] and n_groups: the kernel caches the bin of every
                    # pixel, later iterations with an identical range read the cache back
                    rng_key = (float(lo), float(hi), int(n_groups))
                    reuse = self._bin_range == rng_key
                    _lib.check(L.xb_nk_make_keys(self.dh.data_ptr(), self.slope_tan.data_ptr(), self.aspect.data_ptr(),
                                                 n, float(vshift), float(lo), float(hi), n_groups,
                                                 self._keys.data_ptr(), self._grp.data_ptr(),
                                                 self._bin_cache.data_ptr(), int(reuse), shift, n_digits,
                                                 hist.data_ptr(), moments_t.data_ptr(), self.stream))
                    self._bin_range = rng_key
                    self._allreduce(moments_t)
                else:
                    pre = _u32_tensor(prefix, dev)
                    _lib.check(L.xb_nk_hist_keys(self._keys.data_ptr(), self._grp.data_ptr(), n, n_groups,
                                                 pre.data_ptr(), prefix_mask, shift, n_digits, hist.data_ptr(),
                                                 self.stream))
                self._allreduce(hist)
                h = hist.cpu().numpy().reshape(n_groups, n_digits)
                if ip == 0:
                    counts = h.sum(axis=1)
                    k_lo = (counts - 1) // 2
                digit, below_d = self._pick_digit(h, k_lo - below, counts)
                below += below_d
                prefix = (prefix | (digit.astype(np.uint32) << np.uint32(shift))).astype(np.uint32)
                prefix_mask |= (n_digits - 1) << shift
                last_hist = h[np.arange(n_groups), digit]
            # prefix = exact key of the lower median; `below` keys are smaller, `last_hist` keys are equal to it
            lower = _ordered_to_float(prefix).astype(np.float64)
            median = lower.copy()
            even = (counts % 2 == 0) & (counts > 0)
            need_next = even & (below + last_hist < (counts // 2 + 1))  # the upper median is a strictly larger key
            # the last pass's histogram already holds every key that shares the lower median's high bits: the next
            # occupied digit IS the next larger key; only a lower median that closes its bucket needs a search pass
            upper_key = np.zeros(n_groups, dtype=np.uint32)
            found = np.zeros(n_groups, dtype=bool)
            last_shift, last_digits = passes[-1]
            for g in np.flatnonzero(need_next):
                nz = np.flatnonzero(h[g, digit[g] + 1:])
                if nz.size:
                    d2 = int(digit[g] + 1 + nz[0])
                    low_mask = np.uint32(((last_digits - 1) << last_shift))
                    upper_key[g] = (prefix[g] & ~low_mask) | np.uint32(d2 << last_shift)
                    found[g] = True
            if found.any():
                up = _ordered_to_float(upper_key).astype(np.float64)
                median = np.where(found, 0.5 * (lower + up), median)
            need_next = need_next & ~found
            if need_next.any():
                nxt = _u32_tensor(np.full(n_groups, 0xFFFFFFFF, dtype=np.uint32), dev)
                if mode == 0:
                    _lib.check(L.xb_nk_next(self.dh.data_ptr(), n, int(prefix[0]), nxt.data_ptr(), self.stream))
                else:
                    sel = _u32_tensor(prefix, dev)
                    _lib.check(L.xb_nk_next_keys(self._keys.data_ptr(), self._grp.data_ptr(), n, n_groups,
                                                 sel.data_ptr(), nxt.data_ptr(), self.stream))
                upper = _ordered_to_float(self._allreduce_min_u32(nxt)).astype(np.float64)
                median = np.where(need_next, 0.5 * (lower + upper), median)
        median = np.where(counts > 0, median, np.nan)
        moments = moments_t.cpu().numpy() if (want_moments and mode == 1) else None
        return median, counts, moments


def _nuth_kaab_bin_fit_gpu(state: _NKState, vshift: float, lo: float, hi: float, bin_sizes: int,
                           fit_optimizer: Callable[..., Any]) -> tuple[float, float, float]:
    """GPU restatement of `_nuth_kaab_bin_fit` + `_bin_or_and_fit_nd("bin_and_fit")` (affine.py:358-409,
    base.py:1006-1045): y = dh/slope_tan; p0 = (3*nanstd(y)/sqrt(2), 0, nanmean(y)); per-aspect-bin nanmedian; fit."""
    med, counts, mom = state.select_medians(1, vshift, lo, hi, int(bin_sizes), want_moments=True)
    return _fit_from_bins(med, mom, lo, hi, bin_sizes, fit_optimizer)


def _fit_from_bins(med: np.ndarray, mom: np.ndarray, lo: float, hi: float, bin_sizes: int,
                   fit_optimizer: Callable[..., Any]) -> tuple[float, float, float]:
    """Host part of `_nuth_kaab_bin_fit` (affine.py:384-409, base.py:1027-1045): p0 from the moments of y, bin mid-points,
    the 72-point curve_fit."""
    n, s1, s2 = mom
    mean = s1 / n
    std = float(np.sqrt(max(s2 / n - mean * mean, 0.0)))
    p0 = (3 * std / (2**0.5), 0.0, float(mean))
    # bin mid-points: pd.IntervalIndex.from_breaks(linspace(lo, hi, n+1)).mid  (spatialstats.py:153, base.py:1027)
    edges = np.linspace(lo, hi, int(bin_sizes) + 1)
    mids = 0.5 * (edges[:-1] + edges[1:])
    ok = np.isfinite(med) & np.isfinite(mids)
    if np.all(~ok):
        raise ValueError("Only NaN values after binning, did you pass the right bin edges?")
    results = fit_optimizer(f=_nuth_kaab_fit_func, xdata=mids[ok], ydata=med[ok], sigma=None, absolute_sigma=True,
                            p0=p0)
    a, b, c = results[0]
    return a * np.sin(b), a * np.cos(b), c  # easting, northing, vertical (affine.py:405-407)


def _nuth_kaab_iteration_step_gpu(coords_offsets: tuple[float, float, float], state: _NKState,
                                  res_xy: tuple[float, float], a_e: tuple[float, float], bin_sizes: int,
                                  fit_optimizer: Callable[..., Any]) -> tuple[tuple[float, float, float], float]:
    """affine.py:477-536."""
    dx_px = coords_offsets[0] / a_e[0]
    dy_px = coords_offsets[1] / a_e[1]
    points = getattr(state, "pt_idx", None) is not None and NK_FAST and int(bin_sizes) <= 96 and not state.sharded
    if getattr(state, "use_fast", None) is None:
        state.use_fast = points or state.fast_eligible(int(bin_sizes))
    if state.use_fast:
        res = (state.iteration_points if points else state.iteration_fast)(dx_px, dy_px, int(bin_sizes))
        if res is not None:
            if res["n_fin"] == 0:
                raise ValueError(
                    "The subsample contains no more valid values. This can happen is the horizontal shift to "
                    "correct is very large, or if the algorithm diverged. To ensure all possible points can "
                    "be used at any iteration step, use subsample=1."
                )
            vshift = res["vshift"]
            easting, northing, _ = _fit_from_bins(res["median"], res["moments"], res["lo"], res["hi"], bin_sizes,
                                                  fit_optimizer)
            new_offsets = (coords_offsets[0] + easting * res_xy[0], coords_offsets[1] + northing * res_xy[1], vshift)
            return new_offsets, float(np.sqrt(easting**2 + northing**2))
    lo, hi, n_fin = state.compute_dh(dx_px, dy_px)
    if n_fin == 0:
        raise ValueError(
            "The subsample contains no more valid values. This can happen is the horizontal shift to "
            "correct is very large, or if the algorithm diverged. To ensure all possible points can "
            "be used at any iteration step, use subsample=1."
        )
    med, _, _ = state.select_medians(0, 0.0, 0.0, 1.0, 1)
    vshift = float(med[0])
    easting, northing, _ = _nuth_kaab_bin_fit_gpu(state, vshift, lo, hi, bin_sizes, fit_optimizer)
    new_offsets = (coords_offsets[0] + easting * res_xy[0], coords_offsets[1] + northing * res_xy[1], vshift)
    return new_offsets, float(np.sqrt(easting**2 + northing**2))


def nuth_kaab(ref_elev: Any, tba_elev: Any, inlier_mask: Any = None, transform: Any = None, crs: Any = None,
              area_or_point: Any = None, tolerance: float = 0.001, max_iterations: int = 10,
              params_fit_or_bin: dict[str, Any] | None = None, params_random: dict[str, Any] | None = None,
              z_name: str = "z", weights: Any = None, **kwargs: Any) -> tuple[tuple[float, float, float], int]:
    """Nuth and Kaab (2011) iterative coregistration on the GPU -- same contract as affine.py:539-609: returns the final
    (easting, northing, vertical) offsets in georeferenced units and the number of points used."""
    logging.info("Running Nuth and Kääb (2011) coregistration")
    if crs is not None and hasattr(crs, "is_projected") and not crs.is_projected:
        raise NotImplementedError(
            f"NuthKaab coregistration only works with a projected CRS, current CRS is {crs}. Reproject "
            f"your DEMs with DEM.reproject() in a local projected CRS such as UTM, that you can find "
            f"using DEM.get_metric_crs()."
        )
    pf = dict(fit_or_bin="bin_and_fit", fit_optimizer=scipy.optimize.curve_fit, bin_sizes=72,
              bin_statistic=np.nanmedian)
    pf.update(params_fit_or_bin or {})
    if pf["fit_or_bin"] not in ["fit", "bin_and_fit"]:
        raise ValueError("Nuth and Kääb method only supports 'fit' or 'bin_and_fit'.")
    if pf["fit_or_bin"] != "bin_and_fit" or pf["bin_statistic"] is not np.nanmedian or not isinstance(
            pf["bin_sizes"], (int, np.integer)):
        raise NotImplementedError("the B200 Nuth-Kaab path implements bin_and_fit with an integer number of aspect "
                                  "bins and bin_statistic=np.nanmedian (the reference defaults)")
    pr = dict(subsample=1.0, random_state=None)
    pr.update(params_random or {})

    ref_t, _ = _arrays.to_device(ref_elev)
    tba_t, _ = _arrays.to_device(tba_elev)
    if ref_t.dtype != torch.float32:
        ref_t = ref_t.to(torch.float32)
    if tba_t.dtype != torch.float32:
        tba_t = tba_t.to(torch.float32)
    if ref_t.shape != tba_t.shape or ref_t.dim() != 2:
        raise ValueError("reference and to-be-aligned elevations must be 2-D rasters of the same shape")
    mask_t = None
    if inlier_mask is not None:
        mask_t = inlier_mask if isinstance(inlier_mask, torch.Tensor) else torch.from_numpy(np.asarray(inlier_mask))
    state = _NKState(ref_t, tba_t, mask_t)
    n_valid = state.n_valid()
    if n_valid == 0:
        raise ValueError(
            "There is no valid points common to the input and auxiliary data (bias variables, or "
            "derivatives required for this method, for example slope, aspect, etc)."
        )
    # subsample among the valid cells (base.py:576-621; geoutils.subsample_array's RNG stream is third-party)
    sub = pr["subsample"]
    if sub is not None and sub != 1.0:
        want = int(sub) if sub > 1 else int(sub * n_valid)
        if want < n_valid:
            state.set_points(_pick_valid_points(state, want, n_valid, pr["random_state"]))
            n_valid = want
    offsets = _iterate_nuth_kaab(state, transform, int(pf["bin_sizes"]), pf["fit_optimizer"], tolerance,
                                 max_iterations)
    return offsets, n_valid


def _pick_valid_points(state: _NKState, want: int, n_valid: int, random_state: Any) -> torch.Tensor:
    """A uniformly random subset of ``want`` valid pixels (sorted linear indices), drawn on the device.  Sparse draws --
    the default 5e5 of a large raster -- use rejection sampling (uniform positions, duplicates removed, invalid ones
    dropped, a random ``want`` of the survivors kept), so nothing of the raster's size is materialised; dense draws fall
    back to a permutation of all valid positions.  Deterministic for an integer ``random_state``."""
    dev, n = state.dev, state.rows * state.cols
    seed = int(np.random.default_rng(random_state).integers(0, 2**62))
    g = torch.Generator(device=dev).manual_seed(seed)
    mask_flat = state.sub_mask.view(-1)
    k = int(want * (n / max(n_valid, 1)) * 1.1) + 4096
    if k < n // 8:
        cand = torch.unique(torch.randint(0, n, (k,), generator=g, device=dev))
        cand = cand[mask_flat[cand] != 0]
        if cand.numel() >= want:
            keep = torch.randperm(cand.numel(), generator=g, device=dev)[:want]
            return cand[keep].sort().values
    idx_valid = torch.nonzero(mask_flat).flatten()
    keep = torch.randperm(idx_valid.numel(), generator=g, device=dev)[:want]
    return idx_valid[keep].sort().values


def make_reference_hook(reference_nuth_kaab: Callable[..., Any]) -> Callable[..., Any]:
    """Wrapper with the signature of ``xdem.coreg.affine.nuth_kaab`` (affine.py:539-553) for ``xdem_b200.install()``:
    raster-raster fits with the reference defaults (bin_and_fit, integer bin count, np.nanmedian) run on the GPU,
    everything else (point clouds, ``fit_or_bin="fit"``, custom statistics, weights) is handed to the reference
    function it replaces -- so the rebinding never narrows what ``NuthKaab.fit`` accepts."""

    def nuth_kaab_hook(ref_elev: Any, tba_elev: Any, inlier_mask: Any, transform: Any, crs: Any, area_or_point: Any,
                       tolerance: float, max_iterations: int, params_fit_or_bin: dict[str, Any],
                       params_random: dict[str, Any], z_name: str, weights: Any = None, **kwargs: Any) -> Any:
        pf = params_fit_or_bin or {}
        on_path = (isinstance(ref_elev, (np.ndarray, torch.Tensor)) and isinstance(tba_elev, (np.ndarray, torch.Tensor))
                   and weights is None and pf.get("fit_or_bin", "bin_and_fit") == "bin_and_fit"
                   and pf.get("bin_statistic", np.nanmedian) is np.nanmedian
                   and isinstance(pf.get("bin_sizes", 72), (int, np.integer)))
        fn = nuth_kaab if on_path else reference_nuth_kaab
        return fn(ref_elev=ref_elev, tba_elev=tba_elev, inlier_mask=inlier_mask, transform=transform, crs=crs,
                  area_or_point=area_or_point, tolerance=tolerance, max_iterations=max_iterations,
                  params_fit_or_bin=params_fit_or_bin, params_random=params_random, z_name=z_name, weights=weights,
                  **kwargs)

    nuth_kaab_hook.__wrapped__ = reference_nuth_kaab  # type: ignore[attr-defined]
    return nuth_kaab_hook


def _iterate_nuth_kaab(state: _NKState, transform: Any, bin_sizes: int, fit_optimizer: Callable[..., Any],
                       tolerance: float, max_iterations: int) -> tuple[float, float, float]:
    """`_iterate_method` (affine.py:102-147) around the GPU iteration step."""
    a_e = _transform_coeffs(transform)
    res = (abs(a_e[0]), abs(a_e[1]))
    offsets = (0.0, 0.0, 0.0)
    for i in range(max_iterations):
        offsets, stat = _nuth_kaab_iteration_step_gpu(offsets, state, res, a_e, bin_sizes, fit_optimizer)
        logging.info("      Iteration #%d - Offset: %s; Magnitude: %s", i + 1, offsets, stat)
        if i > 1 and stat < tolerance:
            logging.info("   Last offset was below the residual offset threshold of %s -> stopping", tolerance)
            break
    return offsets


class NuthKaab:
    """Drop-in for ``xdem.coreg.NuthKaab`` (affine.py:2386-2541) for raster-raster fits.  Results are stored like the
    reference in ``self.meta["outputs"]["affine"]`` (``shift_x``, ``shift_y``, ``shift_z``)."""

    def __init__(self, max_iterations: int = 10, offset_threshold: float = 0.001, bin_before_fit: bool = True,
                 fit_optimizer: Callable[..., Any] = scipy.optimize.curve_fit,
                 bin_sizes: int | dict[str, int | Iterable[float]] = 72,
                 bin_statistic: Callable[[np.ndarray], Any] = np.nanmedian, subsample: int | float = 5e5,
                 vertical_shift: bool = True, initial_shift: Any = None) -> None:
        if not callable(fit_optimizer):
            raise TypeError("Argument `fit_optimizer` must be a function (callable), got {}.".format(type(fit_optimizer)))
        if bin_before_fit and not (isinstance(bin_sizes, int) or isinstance(bin_sizes, dict)):
            raise TypeError("Argument `bin_sizes` must be an integer, or a dictionary of integers or iterables, "
                            "got {}.".format(type(bin_sizes)))
        if bin_before_fit and not callable(bin_statistic):
            raise TypeError("Argument `bin_statistic` must be a function (callable), got {}.".format(type(bin_statistic)))
        if initial_shift is not None:
            raise NotImplementedError("initial_shift is not supported by the B200 Nuth-Kaab path yet")
        self.vertical_shift = vertical_shift
        self._meta: dict[str, Any] = {
            "inputs": {
                "random": {"subsample": subsample, "random_state": None},
                "fitorbin": {"fit_or_bin": "bin_and_fit" if bin_before_fit else "fit", "fit_optimizer": fit_optimizer,
                             "bin_sizes": bin_sizes, "bin_statistic": bin_statistic},
                "iterative": {"max_iterations": max_iterations, "tolerance": offset_threshold},
            },
            "outputs": {},
        }
        self._fit_called = False

    @property
    def meta(self) -> dict[str, Any]:
        return self._meta

    def fit(self, reference_elev: Any, to_be_aligned_elev: Any, inlier_mask: Any = None, bias_vars: Any = None,
            weights: Any = None, subsample: float | int | None = None, transform: Any = None, crs: Any = None,
            area_or_point: Any = None, z_name: str = "z", random_state: Any = None, **kwargs: Any) -> "NuthKaab":
        """Estimate the x/y/z offsets (``Coreg.fit``, base.py:2250-2368, for two rasters on the same grid)."""
        if _arrays.is_raster_like(reference_elev):
            transform = transform or reference_elev.transform
            crs = crs or reference_elev.crs
            reference_elev = reference_elev.data
        if _arrays.is_raster_like(to_be_aligned_elev):
            to_be_aligned_elev = to_be_aligned_elev.data
        pr = dict(self._meta["inputs"]["random"])
        if subsample is not None:
            pr["subsample"] = subsample
        pr["random_state"] = random_state
        (east, north, vert), n_used = nuth_kaab(
            reference_elev, to_be_aligned_elev, inlier_mask=inlier_mask, transform=transform, crs=crs,
            area_or_point=area_or_point, z_name=z_name, weights=weights, params_random=pr,
            params_fit_or_bin=self._meta["inputs"]["fitorbin"],
            max_iterations=self._meta["inputs"]["iterative"]["max_iterations"],
            tolerance=self._meta["inputs"]["iterative"]["tolerance"])
        # affine.py:2526-2530
        self._meta["outputs"]["affine"] = {"shift_x": -east, "shift_y": -north,
                                           "shift_z": vert * self.vertical_shift}
        self._meta["outputs"]["random"] = {"subsample_final": n_used}
        self._fit_called = True
        return self

    def apply(self, elev: Any, bias_vars: Any = None, resample: bool = True, resampling: str = "linear",
              transform: Any = None, crs: Any = None, z_name: str = "z", **kwargs: Any) -> tuple[Any, Any]:
        """Apply the estimated translation to a DEM (``Coreg.apply`` -> ``apply_matrix``, base.py:1686-1766, for a
        translation-only matrix): returns ``(applied_dem, transform)``.  With ``resample=True`` (default) the shifted DEM
        is bilinearly regridded on the input grid (GPU, `xb_shift_resample`); with ``resample=False`` only the vertical
        shift is added and the returned transform is translated (base.py:1567-1570)."""
        if not self._fit_called:
            raise AssertionError(".fit() does not seem to have been called yet")
        if resampling not in ("linear", "bilinear"):
            raise NotImplementedError("the B200 apply step implements bilinear resampling only")
        if _arrays.is_raster_like(elev):
            transform = transform or elev.transform
            elev = elev.data
        sx, sy, sz = self.to_translations()
        a, e = _transform_coeffs(transform)
        t, kind = _arrays.to_device(elev)
        if t.dtype != torch.float32:
            t = t.to(torch.float32)
        if not resample:
            out = t + sz
            tr = tuple(transform) if transform is not None else None
            new_tr = (tr[0], tr[1], tr[2] + sx, tr[3], tr[4], tr[5] + sy) if tr is not None else None
            return _arrays.from_device(out, kind), new_tr
        t = t.contiguous()
        out = torch.empty_like(t)
        stream = ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)
        with torch.cuda.device(t.device):
            _lib.check(_lib.lib().xb_shift_resample(t.data_ptr(), t.shape[0], t.shape[1], t.stride(0), -sx / a,
                                                    -sy / e, float(sz), out.data_ptr(), out.stride(0), stream))
        return _arrays.from_device(out, kind), transform

    def to_matrix(self) -> np.ndarray:
        """affine.py:2532-2541."""
        matrix = np.diag(np.ones(4, dtype=float))
        matrix[0, 3] += self._meta["outputs"]["affine"]["shift_x"]
        matrix[1, 3] += self._meta["outputs"]["affine"]["shift_y"]
        matrix[2, 3] += self._meta["outputs"]["affine"]["shift_z"]
        return matrix

    def to_translations(self) -> tuple[float, float, float]:
        o = self._meta["outputs"]["affine"]
        return o["shift_x"], o["shift_y"], o["shift_z"]


from .biascorr import TerrainBias  # noqa: E402,F401  (xdem.coreg.TerrainBias lives next to NuthKaab)
