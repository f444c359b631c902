"""TerrainBias (the bias-correction class on top of the terrain kernel and nd_binning).
CPU: constructor validation mirrors biascorr.py:46-170.  GPU: fit / apply against fixtures produced by the unmodified
reference (`nd_binning`, `interp_nd_binning`, `get_perbin_nd_binning`; oracle/make_golden.py terrainbias)."""

from __future__ import annotations

import numpy as np
import pytest

from tests import parity


def test_terrainbias_validation_without_gpu() -> None:
    from xdem_b200.biascorr import TerrainBias

    with pytest.raises(ValueError, match="must be 'bin_and_fit', 'fit' or 'bin'"):
        TerrainBias(fit_or_bin="nope")
    with pytest.raises(TypeError, match="`bin_sizes` must be an integer"):
        TerrainBias(bin_sizes=2.5)  # type: ignore
    with pytest.raises(TypeError, match="`bin_statistic` must be a function"):
        TerrainBias(bin_statistic="median")  # type: ignore
    with pytest.raises(NotImplementedError):
        TerrainBias(fit_or_bin="fit")
    with pytest.raises(NotImplementedError):
        TerrainBias(bin_statistic=lambda a: 0.0)  # arbitrary callables are not evaluated per bin on the device
    tb = TerrainBias()
    assert tb.meta["inputs"]["specific"]["terrain_attribute"] == "max_curvature"
    assert tb.meta["inputs"]["fitorbin"]["bin_sizes"] == 100 and tb.meta["inputs"]["fitorbin"]["nd"] == 1
    with pytest.raises(AssertionError):
        tb.apply(np.zeros((4, 4), dtype=np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["b100", "edges"])
def test_terrainbias_fit_apply_vs_reference(tag: str) -> None:
    from xdem_b200.biascorr import TerrainBias

    g = parity.load_golden("terrainbias_reference.npz")
    ref, tba, attr = g["ref"], g["tba"], g["attr"]
    bins: object = 100 if tag == "b100" else {"max_curvature": np.concatenate([g[f"{tag}|left"], g[f"{tag}|right"][-1:]])
                                              .astype(np.float32)}
    for method, key in (("linear", "corr_linear"), ("per_bin", "corr_perbin")):
        tb = TerrainBias(bin_sizes=bins, bin_apply_method=method)
        tb.fit(ref, tba, bias_vars={"max_curvature": attr})
        df = tb.meta["outputs"]["fitorbin"]["bin_dataframe"]
        assert np.array_equal(df["count"].values.astype(np.int64), g[f"{tag}|count"])
        assert np.allclose(df["nanmedian"].values, g[f"{tag}|nanmedian"], rtol=1e-6, atol=1e-7, equal_nan=True)
        assert np.allclose([i.left for i in df["max_curvature"].values], g[f"{tag}|left"], rtol=1e-6)
        out = tb.apply(ref, bias_vars={"max_curvature": attr})
        want = (ref.astype(np.float64) + g[f"{tag}|{key}"]).astype(np.float32)
        assert out.dtype == np.float32 and np.array_equal(np.isnan(out), np.isnan(want)), (tag, method)
        m = np.isfinite(want)
        assert np.max(np.abs(out[m] - want[m])) <= 2e-4, (tag, method, np.max(np.abs(out[m] - want[m])))  # ulp(1e3 m)


@pytest.mark.gpu
def test_terrainbias_end_to_end_reduces_the_bias() -> None:
    """Attribute computed by the fused kernel from the transform (biascorr.py:530-534): the curvature-dependent bias
    injected by the fixture generator is removed."""
    import torch

    from xdem_b200.biascorr import TerrainBias

    g = parity.load_golden("terrainbias_reference.npz")
    ref, tba = g["ref"], g["tba"]
    tb = TerrainBias(subsample=0.8)
    tb.fit(ref, tba, transform=(5.0, 0, 0, 0, -5.0, 0), random_state=3)
    assert tb.meta["outputs"]["random"]["subsample_final"] > 0
    corrected = tb.apply(torch.from_numpy(tba).cuda(), transform=(5.0, 0, 0, 0, -5.0, 0))
    assert isinstance(corrected, torch.Tensor) and corrected.dtype == torch.float32
    before = np.nanstd(ref - tba)
    # the correction was fitted on ref's attribute; apply it with the same attribute to tba
    after_t = tb.apply(tba, bias_vars={"max_curvature": g["attr"]})
    after = np.nanstd(ref - after_t)
    assert after < 0.35 * before, (before, after)
    assert abs(np.nanmedian(ref - after_t)) < 0.02
