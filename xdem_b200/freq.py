"""Frequency-domain terrain attributes on the B200: texture shading (mirror of ``xdem/terrain/freq.py``).

``_texture_shading_fft`` keeps the reference's contract (freq.py:62-148): non-finite cells are filled with the mean of
the valid ones, the raster is padded symmetrically to a 2/3/5/7-smooth size, filtered by ``|f|**alpha`` in the
frequency domain and cropped back, with NaN restored.  The fill/pad, spectral scaling and crop/mask stages are CUDA
kernels of ``libxdem_b200`` (csrc/xb_texture.cu); the two transforms are cuFFT calls through ``torch.fft`` -- plain
library FFTs, like ``scipy.fft`` in the reference.  Everything stays on the device; nothing falls back to the CPU.
"""

from __future__ import annotations

from typing import Any

import numpy as np
import torch

from xdem_b200 import _arrays, _lib

__all__ = ["_nextprod_fft", "_texture_shading_fft", "fft_shape"]


def _nextprod_fft(n: int) -> int:
    """Next transform length the reference would pick (freq.py:33-59): a power of two up to 1024, else the next
    2/3/5/7-smooth integer."""
    n = int(n)
    if n <= 1:
        return 1
    if n <= 1024:
        return 1 << (n - 1).bit_length()
    m = n
    while True:
        rest = m
        for f in (2, 3, 5, 7):
            while rest % f == 0:
                rest //= f
        if rest == 1:
            return m
        m += 1


def fft_shape(shape: tuple[int, int]) -> tuple[int, int, int, int]:
    """(fft_rows, fft_cols, pad_rows, pad_cols) of a raster shape (freq.py:100-106)."""
    rows, cols = int(shape[0]), int(shape[1])
    fr, fc = _nextprod_fft(rows), _nextprod_fft(cols)
    return fr, fc, (fr - rows) // 2, (fc - cols) // 2


def _texture_shading_fft(dem: Any, alpha: float | None = 0.8) -> Any:
    """Texture shading of a 2-D raster (ndarray, masked array or torch tensor; float32 / float64).

    Returns the same container kind as the input (CUDA tensors stay on the device)."""
    on_device = isinstance(dem, torch.Tensor) and dem.is_cuda
    as_tensor = isinstance(dem, torch.Tensor)
    t, _ = _arrays.to_device(dem)
    if t.dim() != 2:
        raise ValueError("texture shading expects a 2-D elevation array")
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float32)
    if t.stride(1) != 1:
        t = t.contiguous()
    out = texture_shading_device(t, alpha)
    if on_device:
        return out
    res = out.cpu()
    return res if as_tensor else res.numpy()


def texture_shading_device(t: torch.Tensor, alpha: float | None) -> torch.Tensor:
    """Device pipeline: prepare (kernels) -> rfft2 (cuFFT) -> filter (kernel) -> irfft2 (cuFFT) -> finish (kernel)."""
    if alpha is None:
        alpha = 0.8  # freq.py:77-78
    if not 0 <= alpha <= 2:
        raise ValueError(f"Alpha must be between 0 and 2, got {alpha}")
    L = _lib.lib()
    rows, cols = int(t.shape[0]), int(t.shape[1])
    fr, fc, pr, pc = fft_shape((rows, cols))
    dt = 1 if t.dtype == torch.float64 else 0
    dev = t.device
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        stats = torch.empty(3, dtype=torch.float64, device=dev)
        padded = torch.empty((fr, fc), dtype=t.dtype, device=dev)
        # alpha > 0 removes the DC term, so the raster can be centred on its mean before the single-precision transform
        _lib.check(L.xb_texture_prepare(t.data_ptr(), dt, rows, cols, t.stride(0), padded.data_ptr(), fr, fc, pr, pc,
                                        1 if alpha > 0 else 0, stats.data_ptr(), stream))
        out = torch.empty((rows, cols), dtype=t.dtype, device=dev)
        if float(stats[2].item()) == 0.0:  # no valid cell at all (freq.py:85-87)
            out.fill_(float("nan"))
            return out
        spec = torch.fft.rfft2(padded)
        del padded
        if not spec.is_contiguous():
            spec = spec.contiguous()
        _lib.check(L.xb_texture_filter(spec.data_ptr(), dt, fr, fc, float(alpha), stream))
        back = torch.fft.irfft2(spec, s=(fr, fc))
        del spec
        if not back.is_contiguous():
            back = back.contiguous()
        _lib.check(L.xb_texture_finish(back.data_ptr(), dt, fr, fc, pr, pc, t.data_ptr(), rows, cols, t.stride(0),
                                       out.data_ptr(), out.stride(0), stream))
    return out
