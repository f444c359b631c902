"""CPU: the Nuth-Kaab oracle is pinned against fixtures produced by the reference's own iteration code; the variogram
oracle (PARITY UNPINNED: scikit-gstat is absent, see oracle/variogram_oracle.py) is checked for internal consistency
(NumPy restatement == plain-C restatement) and against hand-computable cases."""

from __future__ import annotations

import os

import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import nk_oracle as nk
from oracle import variogram_oracle as vo
from tests import parity


def _nk_golden() -> dict[str, np.ndarray]:
    return parity.load_golden("nk_reference.npz")


def test_nk_oracle_aux_bit_exact() -> None:
    g = _nk_golden()
    st, asp = nk.aux_vars(g["ref"])
    ref_st = g["slope_tan"].copy()
    ref_st[np.isclose(ref_st, 0)] = np.nan
    assert np.array_equal(st, ref_st, equal_nan=True)
    assert np.array_equal(asp, g["aspect"], equal_nan=True)


def test_nk_oracle_iterations_vs_reference() -> None:
    g = _nk_golden()
    a_e = (float(g["transform"][0]), float(g["transform"][4]))
    out, n_valid, hist = nk.nuth_kaab(g["ref"], g["tba"], g["inlier"], a_e, tolerance=0.0, max_iterations=6)
    assert n_valid == int(g["n_valid"])
    for i, (offs, _) in enumerate(hist):
        assert np.allclose(offs, g["offsets"][i], rtol=1e-6, atol=1e-7), (i, offs, g["offsets"][i])


def test_variogram_oracles_agree() -> None:
    rng = np.random.default_rng(44)
    shape, gsd = (120, 150), 5.0
    coords = vo.grid_coords(shape, gsd)
    vals = rng.normal(size=shape).flatten()
    idx = rng.choice(vals.size, 2000, replace=False)
    c, v = coords[idx], vals[idx]
    for bins in ("even", vo.default_bins(gsd, np.hypot(119 * gsd, 149 * gsd))):
        b, exp, cnt = vo.empirical_variogram(c, v, bins, n_lags=25)
        cc, ss, dmax = co.variogram_pairs(c, v, b)
        assert np.array_equal(cc, cnt)
        with np.errstate(all="ignore"):
            e2 = np.where(cc > 0, ss / (2.0 * cc), np.nan)
        assert np.allclose(e2, exp, rtol=1e-12, equal_nan=True)
        if bins == "even":
            assert b[-1] == dmax and cnt.sum() == 2000 * 1999 // 2 - 1  # only the farthest pair sits on the last edge


def test_variogram_oracle_hand_case() -> None:
    """4 samples on a unit square: 4 pairs at d=1 (diffs 1,1,1,1... ) and 2 at sqrt(2)."""
    coords = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0]])
    v = np.array([0.0, 1.0, 2.0, 4.0])
    b, exp, cnt = vo.empirical_variogram(coords, v, [1.2, 2.0])
    assert list(cnt) == [4, 2]
    # pairs at d=1: (0,1)=1,(0,2)=2,(1,3)=3,(2,3)=2 -> sum sq = 18 -> 18/(2*4); diagonals: (0,3)=4,(1,2)=1 -> 17/4
    assert exp[0] == pytest.approx(18 / 8) and exp[1] == pytest.approx(17 / 4)


def test_golden_files_are_committed() -> None:
    for f in ("terrain_reference.npz", "nk_reference.npz"):
        assert os.path.exists(os.path.join(parity.GOLDEN, f))
