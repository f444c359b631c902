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


def _edge_golden() -> dict:
    import json

    with open(os.path.join(parity.GOLDEN, "variogram_edges.json")) as f:
        return json.load(f)


def test_variogram_edge_golden_is_reproducible() -> None:
    """tests/golden/variogram_edges.json is what `python -m oracle.variogram_exact` writes (exact integer arithmetic)."""
    from oracle import variogram_exact as ve

    assert ve.build_golden()["cases"] == _edge_golden()["cases"]
    # the hand computation of the first case (3x3 block: 12 pairs at d=1, 8 at sqrt2, 6 at 2, 8 at sqrt5, 2 at sqrt8)
    c0 = _edge_golden()["cases"][0]
    assert c0["left"]["count"] == [0, 20, 16] and c0["right"]["count"] == [12, 14, 10]


@pytest.mark.parametrize("gsd", [1.0, 5.0, 0.5, 30.0])
def test_variogram_float_oracle_follows_left_rule_on_edge_cases(gsd: float) -> None:
    """The float64 (pdist-based) restatement and the exact-integer one agree under the "left" rule on every case with
    pairs exactly on edges, for gsd values whose lattice distances are exact in float64; the product's integer
    thresholds reproduce BOTH rules."""
    from fractions import Fraction

    from xdem_b200.spatialstats import edge_thresholds

    for c in _edge_golden()["cases"]:
        pts = np.asarray(c["points"], dtype=np.float64) * gsd
        vals = np.asarray(c["values"], dtype=np.float64)
        e_sq = [Fraction(n, d) for n, d in c["edges_sq_num_den"]]
        edges = [gsd * float(np.sqrt(float(e))) for e in e_sq]
        if any(abs((e / gsd) ** 2 - float(q)) > 1e-12 * float(q) for e, q in zip(edges, e_sq)):
            continue
        _, exp, cnt = vo.empirical_variogram(pts, vals, edges)
        exact_sqrt = all(float(np.sqrt(float(q))) ** 2 == float(q) for q in e_sq)
        if exact_sqrt:  # integer edges: the float edge IS the lattice distance, the float oracle must follow "left"
            assert list(cnt) == c["left"]["count"], c["name"]
        for rule in ("left", "right"):
            T = edge_thresholds(edges, gsd, rule)
            ipts = np.asarray(c["points"], dtype=np.int64)
            d2 = ((ipts[:, None, :] - ipts[None, :, :]) ** 2).sum(-1)[np.triu_indices(len(ipts), 1)]
            got = [int(((d2 >= (T[k - 1] if k else 0)) & (d2 < T[k])).sum()) for k in range(len(T))]
            if exact_sqrt:
                assert got == c[rule]["count"], (c["name"], rule, gsd)
            else:  # irrational edges (sqrt 2 ...): the float edge is a rounded value next to the lattice distance; the
                # thresholds must then agree with the float64 comparison itself
                d = np.sqrt((gsd * gsd) * d2.astype(np.float64))
                lo = [0.0] + edges[:-1]
                want = [int((((d >= lo[k]) & (d < edges[k])) if rule == "left" else ((d > lo[k]) & (d <= edges[k]))).sum())
                        for k in range(len(edges))]
                assert got == want, (c["name"], rule, gsd)


@pytest.mark.parametrize("gsd", [0.1, 0.3, 0.7])
def test_non_dyadic_gsd_only_moves_on_edge_pairs(gsd: float) -> None:
    """For a gsd that is not exact in binary the reference rounds every coordinate (index*gsd) separately, so pairs whose
    lattice distance coincides with a bin edge (d2 = 2^k for the default sqrt(2)-geometric edges) can land in either
    neighbouring class depending on the pair; the integer thresholds assign all of them to one class.  Every other pair
    is classified identically, and the count deviation per class is bounded by the number of on-edge pairs."""
    from xdem_b200.spatialstats import edge_thresholds, on_edge_d2

    n = 40
    coords = vo.grid_coords((n, n), gsd)
    rng = np.random.default_rng(3)
    idx = rng.choice(n * n, 500, replace=False)
    c = coords[idx]
    maxlag = float(np.hypot(c[:, 0].max() - c[:, 0].min(), c[:, 1].max() - c[:, 1].min()))
    edges = vo.default_bins(gsd, maxlag)
    _, _, cnt = vo.empirical_variogram(c, np.zeros(len(c)), edges)
    ix = np.rint(c / gsd).astype(np.int64)
    d2 = ((ix[:, None, :] - ix[None, :, :]) ** 2).sum(-1)[np.triu_indices(len(ix), 1)]
    T = edge_thresholds(edges, gsd, "left")
    got = np.array([int(((d2 >= (T[k - 1] if k else 0)) & (d2 < T[k])).sum()) for k in range(len(T))])
    amb = on_edge_d2(edges, gsd)
    n_amb = np.array([int(np.isin(d2, [a for a in amb if abs(np.sqrt(gsd * gsd * a) - e) <= 4 * np.spacing(e)]).sum())
                      for e in edges])
    # a pair on edge k can only move between classes k and k+1
    dev = np.abs(got - cnt)
    bound = n_amb + np.concatenate([[0], n_amb[:-1]])
    assert np.all(dev <= bound), (gsd, dev, bound)
    assert dev.sum() <= 2 * n_amb.sum()
    # without the on-edge pairs the two classifications coincide exactly
    keep = ~np.isin(d2, amb)
    d = np.sqrt(((c[:, None, :] - c[None, :, :]) ** 2).sum(-1))[np.triu_indices(len(c), 1)]
    lo = [0.0] + edges[:-1]
    for k in range(len(edges)):
        assert int(((d >= lo[k]) & (d < edges[k]) & keep).sum()) == int(((d2 >= (T[k - 1] if k else 0)) & (d2 < T[k]) & keep).sum())
