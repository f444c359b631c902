"""ctypes binding of libxdem_b200.so (include/xdem_b200.h).  There is NO fallback: if the CUDA library is missing or a
call fails, an exception is raised."""

from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_int, c_int64, c_uint32, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libxdem_b200.so")

_lib: ctypes.CDLL | None = None


class XdemB200Error(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Load the CUDA library (built by ``python -m xdem_b200.build`` / ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise XdemB200Error(
            f"{LIB_PATH} not found: build the CUDA extension first (python -m xdem_b200.build). "
            "xdem_b200 has no CPU fallback."
        )
    L = ctypes.CDLL(LIB_PATH)
    L.xb_last_error.restype = c_char_p
    L.xb_version.restype = c_int
    L.xb_launch_count.restype = c_uint64
    L.xb_terrain_fused.restype = c_int
    L.xb_terrain_fused.argtypes = [c_void_p, c_int, c_int64, c_int64, c_int64, c_int64, c_int64, c_double, c_int,
                                   c_int, c_uint32, c_uint32, c_int, c_int, c_int, c_int, c_double, c_double,
                                   c_double, ctypes.POINTER(c_void_p), c_int64, c_void_p]
    if hasattr(L, "xb_terrain_fused_host"):
        L.xb_terrain_fused_host.restype = c_int
        L.xb_terrain_fused_host.argtypes = [c_void_p, c_int, c_int64, c_int64, c_double, c_int, c_int, c_uint32,
                                            c_uint32, c_int, c_int, c_int, c_int, c_double, c_double, c_double,
                                            ctypes.POINTER(c_void_p), c_int64]
    L.xb_terrain_fused_host_rows.restype = c_int
    L.xb_terrain_fused_host_rows.argtypes = [c_void_p, c_int, c_int64, c_int64, c_int64, c_int64, c_double, c_int, c_int,
                                             c_uint32, c_uint32, c_int, c_int, c_int, c_int, c_double, c_double,
                                             c_double, ctypes.POINTER(c_void_p), c_int64]
    L.xb_release_scratch.restype = c_int
    L.xb_set_option.restype = c_int
    L.xb_set_option.argtypes = [c_char_p, c_int]
    L.xb_windowed_generic.restype = c_int
    L.xb_windowed_generic.argtypes = [c_void_p, c_int, c_int64, c_int64, c_int64, c_int64, c_int64, c_int, c_uint32,
                                      c_int, ctypes.POINTER(c_void_p), c_int64, c_void_p]
    L.xb_variogram_group_size.restype = c_int
    L.xb_variogram_chunk.restype = c_int
    L.xb_variogram_pairs.restype = c_int
    L.xb_variogram_pairs.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_void_p, c_int64, c_int64, c_int,
                                     c_int, c_void_p, c_void_p, c_void_p]
    L.xb_variogram_median_pass.restype = c_int
    L.xb_variogram_median_pass.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_void_p, c_int64, c_int64,
                                           c_int, c_void_p, c_uint32, c_int, c_void_p, c_void_p, c_void_p]
    L.xb_variogram_maxd2.restype = c_int
    L.xb_variogram_maxd2.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]
    L.xb_variogram_pairs_xy.restype = c_int
    L.xb_variogram_pairs_xy.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64,
                                        c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    L.xb_nk_aux.restype = c_int
    L.xb_nk_aux.argtypes = [c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_int64, c_int64, c_void_p, c_void_p,
                            c_int64, c_void_p]
    L.xb_nk_dh.restype = c_int
    L.xb_nk_dh.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64,
                           c_int64, c_double, c_double, c_void_p, c_void_p, c_void_p, c_void_p]
    L.xb_shift_resample.restype = c_int
    L.xb_shift_resample.argtypes = [c_void_p, c_int64, c_int64, c_int64, c_double, c_double, c_double, c_void_p,
                                    c_int64, c_void_p]
    L.xb_nk_hist.restype = c_int
    L.xb_nk_hist.argtypes = [c_void_p, c_int64, c_uint32, c_uint32, c_int, c_int, c_void_p, c_void_p]
    L.xb_nk_next.restype = c_int
    L.xb_nk_next.argtypes = [c_void_p, c_int64, c_uint32, c_void_p, c_void_p]
    L.xb_nk_make_keys.restype = c_int
    L.xb_nk_make_keys.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_double, c_double, c_double, c_int, c_void_p,
                                  c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]
    L.xb_nk_hist_keys.restype = c_int
    L.xb_nk_hist_keys.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_uint32, c_int, c_int, c_void_p,
                                  c_void_p]
    L.xb_nk_next_keys.restype = c_int
    L.xb_nk_next_keys.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]
    for name in ("xb_nkf_layout", "xb_nkf_reset", "xb_nkf_dh", "xb_nkf_range", "xb_nkf_y", "xb_nkf_select",
                 "xb_nkf_finalize", "xb_nkf_iteration"):
        getattr(L, name).restype = c_int
    L.xb_nkf_layout.argtypes = [c_void_p]
    L.xb_nkf_reset.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    L.xb_nkf_dh.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64,
                            c_int64, c_double, c_double, c_void_p, c_void_p, c_int, c_uint32, c_void_p, c_void_p,
                            c_void_p, c_uint64, c_void_p]
    L.xb_nkf_range.argtypes = [c_void_p, c_void_p, c_void_p]
    L.xb_nkf_y.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p,
                           c_int, c_uint32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_uint64, c_void_p]
    L.xb_nkf_select.argtypes = [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int, c_int, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_uint64, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p]
    L.xb_nkf_finalize.argtypes = [c_void_p, c_void_p, c_uint64, c_uint64, c_void_p]
    L.xb_nkf_iteration.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64,
                                   c_int64, c_int64, c_double, c_double, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_int64, c_int, c_uint32, c_void_p, c_uint64, c_void_p, c_void_p, c_uint64, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    L.xb_bin_moments.restype = c_int
    L.xb_bin_moments.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    L.xb_nkf_iteration_points.restype = c_int
    L.xb_nkf_iteration_points.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64,
                                          c_int64, c_int64, c_int64, c_int64, c_double, c_double, c_int, c_void_p,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_void_p]
    L.xb_nk_dh_points.restype = c_int
    L.xb_nk_dh_points.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int64,
                                  c_int64, c_int64, c_double, c_double, c_void_p, c_void_p, c_void_p, c_void_p]
    L.xb_nk_prepare.restype = c_int
    L.xb_nk_prepare.argtypes = [c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_int64, c_int64, c_void_p, c_int64,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    L.xb_texture_prepare.restype = c_int
    L.xb_texture_prepare.argtypes = [c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int64,
                                     c_int64, c_int, c_void_p, c_void_p]
    L.xb_texture_filter.restype = c_int
    L.xb_texture_filter.argtypes = [c_void_p, c_int, c_int64, c_int64, c_double, c_void_p]
    L.xb_texture_finish.restype = c_int
    L.xb_texture_finish.argtypes = [c_void_p, c_int, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int64,
                                    c_int64, c_void_p, c_int64, c_void_p]
    L.xb_bin_keys.restype = c_int
    L.xb_bin_keys.argtypes = [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p]
    L.xb_bin_hist.restype = c_int
    L.xb_bin_hist.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_uint32, c_int, c_void_p, c_void_p]
    L.xb_bin_next.restype = c_int
    L.xb_bin_next.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]
    L.xb_bin_absdev_keys.restype = c_int
    L.xb_bin_absdev_keys.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]
    L.xb_bin_apply_1d.restype = c_int
    L.xb_bin_apply_1d.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]
    L.xb_probe_stream_tma.restype = c_int
    L.xb_probe_stream_tma.argtypes = [c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p]
    L.xb_probe_stream.restype = c_int
    L.xb_probe_stream.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_void_p]
    L.xb_probe_exact_math.restype = c_int
    L.xb_probe_exact_math.argtypes = [c_int, c_uint32, c_uint64, ctypes.c_float, ctypes.c_float, c_void_p, c_void_p]
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().xb_last_error().decode("utf-8", "replace")
        raise XdemB200Error(f"libxdem_b200 error {rc}: {msg}")


def set_option(name: str, value: int) -> None:
    check(lib().xb_set_option(name.encode(), int(value)))


def launch_count() -> int:
    return int(lib().xb_launch_count())


#: every symbol declared in include/xdem_b200.h (checked by tests/test_abi.py)
EXPORTED = ["xb_last_error", "xb_version", "xb_launch_count", "xb_terrain_fused", "xb_terrain_fused_host",
            "xb_variogram_group_size", "xb_variogram_chunk", "xb_variogram_pairs", "xb_variogram_median_pass", "xb_variogram_maxd2", "xb_nk_aux",
            "xb_nk_dh", "xb_nk_hist", "xb_nk_next", "xb_nk_make_keys", "xb_nk_hist_keys", "xb_nk_next_keys",
            "xb_windowed_generic", "xb_set_option", "xb_shift_resample", "xb_texture_prepare", "xb_texture_filter",
            "xb_texture_finish", "xb_bin_keys", "xb_bin_hist", "xb_bin_next", "xb_bin_absdev_keys", "xb_probe_stream", "xb_probe_exact_math",
            "xb_terrain_fused_host_rows", "xb_release_scratch", "xb_variogram_pairs_xy",
            "xb_nkf_layout", "xb_nkf_reset", "xb_nkf_dh", "xb_nkf_range", "xb_nkf_y", "xb_nkf_select", "xb_nkf_finalize", "xb_nkf_iteration", "xb_nk_prepare", "xb_nk_dh_points", "xb_nkf_iteration_points", "xb_bin_moments", "xb_bin_apply_1d", "xb_probe_stream_tma"]

__all__ = ["lib", "check", "launch_count", "set_option", "XdemB200Error", "LIB_PATH", "EXPORTED"]
