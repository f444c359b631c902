"""TEST INFRASTRUCTURE -- CPU restatement of ``xdem.spatialstats.nd_binning`` (spatialstats.py:91-216) for the parity
tests.  Never imported by the product (xdem_b200/).

The binning itself is SciPy's (``scipy.stats.binned_statistic_dd``, scipy/stats/_binned_statistic.py; SciPy 1.18.1 in
this image): the reference calls it once per statistic for every 1-D / 2-D / N-D combination of the explanatory
variables.  This restatement returns plain arrays keyed by the variable combination instead of a DataFrame; it is pinned
against the DataFrame of the unmodified reference in tests/golden/binning_reference.npz (oracle/make_golden.py).
``nmad`` restates geoutils.stats.nmad (third-party, absent here: 1.4826 * nanmedian(|x - nanmedian(x)|))."""

from __future__ import annotations

import itertools

import numpy as np
from scipy.stats import binned_statistic_dd


def nmad(data: np.ndarray, nfact: float = 1.4826) -> float:
    data = np.asarray(data)
    return nfact * np.nanmedian(np.abs(data - np.nanmedian(data)))


def nd_binning(values: np.ndarray, list_var: list[np.ndarray], list_bins: list, with_nmad: bool = True
               ) -> dict[tuple[int, ...], dict[str, np.ndarray]]:
    """{variable-index tuple: {"count", "median", "nmad", "edges"}} with statistics flattened in C order
    (spatialstats.py:139-200)."""
    values = np.asarray(values).ravel()
    list_var = [np.asarray(v).ravel() for v in list_var]
    ok = np.isfinite(values)
    for v in list_var:
        ok &= np.isfinite(v)  # spatialstats.py:128-131
    values = values[ok]
    list_var = [v[ok] for v in list_var]
    combos: list[tuple[int, ...]] = [(i,) for i in range(len(list_var))]
    if len(list_var) > 1:
        combos += list(itertools.combinations(range(len(list_var)), 2))
    if len(list_var) > 2:
        combos.append(tuple(range(len(list_var))))
    out = {}
    for c in combos:
        sample = [list_var[i] for i in c]
        bins = [list_bins[i] for i in c]
        cnt, edges, _ = binned_statistic_dd(sample, values, statistic="count", bins=bins)
        med = binned_statistic_dd(sample, values, statistic=np.nanmedian, bins=bins)[0]
        res = {"count": cnt.ravel(), "median": med.ravel(), "edges": edges}
        if with_nmad:
            res["nmad"] = binned_statistic_dd(sample, values, statistic=nmad, bins=bins)[0].ravel()
        out[c] = res
    return out
