"""Texture shading: oracle vs fixtures of the unmodified reference (CPU) and the CUDA +
cuFFT path vs both (GPU).

Float criterion: the attribute is a global filter, so errors scale with the largest response, not with the local value
(it crosses zero everywhere): |x - ref| <= 1e-5 * max|ref| + 1e-6, exact NaN mask.  The reference transforms float32
rasters in single precision (scipy.fft keeps the input dtype) WITHOUT centring them, so its own float32 output carries
round-off proportional to the elevations (~1e3 m): comparisons with float32 reference fixtures add `noise` =
8 eps32 max|dem|; the float64 fixtures pin the algorithm itself to 1e-9, and the float32 CUDA path is also checked
against the float64 truth of the same input at 2e-5 of the response scale (it centres the raster first)."""

from __future__ import annotations

import numpy as np
import pytest

from tests import parity


EPS32 = float(np.finfo(np.float32).eps)


def _close(x: np.ndarray, ref: np.ndarray, msg: str, rtol: float = 1e-5, noise: float = 0.0,
           same_dtype: bool = True) -> None:
    assert x.shape == ref.shape, msg
    if same_dtype:
        assert x.dtype == ref.dtype, msg
    assert np.array_equal(np.isnan(x), np.isnan(ref)), f"{msg}: NaN masks differ"
    if np.isfinite(ref).any():
        scale = float(np.nanmax(np.abs(ref)))
        err = float(np.nanmax(np.abs(x.astype(np.float64) - ref.astype(np.float64))))
        assert err <= rtol * scale + 1e-6 + noise, f"{msg}: max abs err {err:.3g} vs scale {scale:.3g}"


def _ref_noise(dem: np.ndarray) -> float:
    """Round-off of the reference's own un-centred single-precision transform (0 for float64 rasters)."""
    if dem.dtype != np.float32 or not np.isfinite(dem).any():
        return 0.0
    return 8.0 * EPS32 * float(np.nanmax(np.abs(dem)))


@pytest.fixture(scope="module")
def T() -> dict[str, np.ndarray]:
    return parity.load_golden("texture_reference.npz")


def _cases(T: dict[str, np.ndarray]):
    for k in T:
        if k.startswith("tex|"):
            _, name, alpha = k.split("|")
            yield name, float(alpha), T[f"in|{name}"], T[k]


def test_oracle_matches_reference_fixtures(T) -> None:
    from oracle import terrain_oracle as to

    n = 0
    for name, alpha, dem, ref in _cases(T):
        got = to.get_terrain_attribute(dem, "texture_shading", texture_alpha=alpha)
        _close(got, ref, f"oracle {name} alpha={alpha}", rtol=2e-6)
        n += 1
    assert n == 11


def test_fft_lengths_and_validation() -> None:
    from oracle import terrain_oracle as to
    from xdem_b200 import freq

    for n in (0, 1, 2, 3, 5, 52, 1000, 1024, 1025, 1100, 2049, 4097, 10007, 32768):
        assert freq._nextprod_fft(n) == to.next_fft_length(n)
    assert freq.fft_shape((40, 52)) == (64, 64, 12, 6)
    assert freq.fft_shape((30, 1100)) == (32, 1120, 1, 10)
    with pytest.raises(ValueError, match="Alpha must be between 0 and 2"):
        to.texture_shading(np.zeros((4, 4), np.float32), alpha=2.5)


def test_reference_doctest_properties_oracle() -> None:
    """terrain.py:1816-1831: flat surface -> all zeros; shape preserved."""
    from oracle import terrain_oracle as to

    flat = np.ones((5, 5), dtype=np.float32)
    assert np.all(to.texture_shading(flat, 0.8) == 0)


# ---------------------------------------------------------------------------------------------------------- GPU


@pytest.mark.gpu
def test_gpu_vs_reference_fixtures(T) -> None:
    import xdem_b200

    n = 0
    for name, alpha, dem, ref in _cases(T):
        got = xdem_b200.terrain.get_terrain_attribute(dem, "texture_shading", texture_alpha=alpha)
        if dem.dtype == np.float64:
            _close(got, ref, f"gpu {name} alpha={alpha}", rtol=1e-9)  # pins padding, filter, crop exactly
        else:
            _close(got, ref, f"gpu {name} alpha={alpha}", noise=_ref_noise(dem))
        n += 1
    assert n == 11
    out = xdem_b200.terrain.texture_shading(T["in|fractal"], alpha=1.5)
    _close(out, T["tex|fractal|1.5"], "wrapper", noise=_ref_noise(T["in|fractal"]))
    with pytest.raises(ValueError, match="Alpha must be between 0 and 2"):
        xdem_b200.terrain.texture_shading(T["in|fractal"], alpha=-0.1)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1, 1), (2, 7), (33, 1030), (257, 131), (1500, 1201)])
def test_gpu_vs_oracle_shapes(shape) -> None:
    import torch

    import xdem_b200
    from oracle import synth, terrain_oracle as to

    dem = synth.fractal_dem(shape, seed=shape[1])
    if dem.size > 20:
        dem = synth.inject_nans(dem, frac=0.002, hole=2)
    for alpha in (0.8, 1.0):
        truth = to.texture_shading(dem.astype(np.float64), alpha)  # float64 restatement of the same input
        got = xdem_b200.terrain.get_terrain_attribute(dem, "texture_shading", texture_alpha=alpha)
        assert got.dtype == np.float32
        _close(got, truth, f"{shape} alpha={alpha}", rtol=2e-5, same_dtype=False)
        got64 = xdem_b200.terrain.get_terrain_attribute(dem.astype(np.float64), "texture_shading", texture_alpha=alpha)
        _close(got64, truth, f"{shape} alpha={alpha} f64", rtol=1e-9)
    # device tensors stay on the device; mixed requests keep the request order
    t = torch.from_numpy(dem).cuda()
    outs = xdem_b200.terrain.get_terrain_attribute(t, ["texture_shading", "slope"], resolution=5.0, texture_alpha=0.8)
    assert outs[0].is_cuda and outs[0].shape == t.shape
    _close(outs[0].cpu().numpy(), to.texture_shading(dem.astype(np.float64), 0.8), "tensor", rtol=2e-5,
           same_dtype=False)


@pytest.mark.gpu
def test_gpu_properties_large() -> None:
    """Size-independent properties at 8192 x 6000 (pads to 8192 x 6000 = 2^4 3 5^3): linear in the elevations, blind to
    a constant offset for alpha > 0, identity for alpha = 0, zero on a flat raster."""
    import torch

    from xdem_b200 import freq

    g = torch.Generator(device="cuda").manual_seed(2)
    z = torch.cumsum(torch.cumsum(torch.randn((8192, 6000), generator=g, device="cuda"), 0), 1).float() * 0.01
    a = freq.texture_shading_device(z, 0.8)
    b = freq.texture_shading_device(z * 2.0 + 64.0, 0.8)
    scale = float(a.abs().max())
    assert float((b - 2.0 * a).abs().max()) <= 2e-5 * scale
    ident = freq.texture_shading_device(z, 0.0)
    assert float((ident - z).abs().max()) <= 2e-6 * float(z.abs().max())
    flat = freq.texture_shading_device(torch.full((300, 500), 7.25, device="cuda"), 1.2)
    assert float(flat.abs().max()) <= 1e-5
    nan_in = z[:100, :100].clone()
    nan_in[:] = float("nan")
    assert bool(torch.isnan(freq.texture_shading_device(nan_in, 0.8)).all())
