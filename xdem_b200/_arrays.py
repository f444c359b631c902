"""Host<->device plumbing: NumPy / masked arrays / torch tensors / Raster-like objects -> contiguous CUDA tensors and
back.  torch is used for device memory and streams only."""

from __future__ import annotations

from typing import Any

import numpy as np
import torch


def is_raster_like(obj: Any) -> bool:
    """Duck-typed geoutils.Raster / xdem.DEM (geoutils is optional here)."""
    return hasattr(obj, "transform") and hasattr(obj, "crs") and hasattr(obj, "data") and hasattr(obj, "res")


def array_dtype(dem: Any) -> np.dtype:
    if isinstance(dem, torch.Tensor):
        return np.dtype(str(dem.dtype).replace("torch.", ""))
    return np.dtype(dem.dtype)


def to_host_nan_array(dem: Any) -> np.ndarray:
    """ndarray / masked array / Raster-like -> ndarray with NaN at invalid cells (mirrors
    geoutils.raster.get_array_and_mask as used at terrain.py:558); integer arrays -> float32 (terrain.py:560-561)."""
    if is_raster_like(dem):
        dem = dem.data
    if isinstance(dem, np.ma.MaskedArray):
        data = np.asarray(dem.data)
        if not np.issubdtype(data.dtype, np.floating):
            data = data.astype(np.float32)
        else:
            data = data.copy()
        data[np.ma.getmaskarray(dem)] = np.nan
        arr = data
    else:
        arr = np.asarray(dem)
        if not np.issubdtype(arr.dtype, np.floating):
            arr = arr.astype(np.float32)
    arr = np.squeeze(arr) if arr.ndim == 3 and arr.shape[0] == 1 else arr
    return arr


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("xdem_b200 needs a CUDA device (B200, sm_100a): no CPU fallback exists.")
    return torch.device("cuda", torch.cuda.current_device())


def to_device(dem: Any) -> tuple[torch.Tensor, str]:
    """Returns (2-D contiguous float32/float64 CUDA tensor, kind) with kind in {"torch", "numpy"}."""
    dev = require_cuda()
    if isinstance(dem, torch.Tensor):
        t = dem
        if not t.is_floating_point():
            t = t.to(torch.float32)
        if t.dtype not in (torch.float32, torch.float64):
            t = t.to(torch.float32)
        if t.device.type != "cuda":
            t = t.to(dev)
        return t.contiguous(), "torch"
    arr = to_host_nan_array(dem)
    if arr.dtype not in (np.float32, np.float64):
        arr = arr.astype(np.float32)
    t = torch.from_numpy(np.ascontiguousarray(arr)).to(dev)
    return t, "numpy"


def from_device(t: torch.Tensor, kind: str, out_dtype: Any = None) -> Any:
    if kind == "torch":
        if out_dtype is not None:
            t = t.to(getattr(torch, np.dtype(out_dtype).name))
        return t
    a = t.cpu().numpy()
    if out_dtype is not None and a.dtype != np.dtype(out_dtype):
        a = a.astype(out_dtype)
    return a
