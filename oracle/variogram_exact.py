"""TEST INFRASTRUCTURE ONLY -- second, independent restatement of the lag-class rule of the empirical variogram for GRID
samples, in exact integer / rational arithmetic (no floating-point distance is ever formed).

Why: the pair work of ``xdem.spatialstats._get_pdist_empirical_variogram`` (spatialstats.py:1064-1101) happens inside
scikit-gstat, which is absent here (PARITY UNPINNED, see oracle/variogram_oracle.py).  The first oracle follows the
float64 route of scikit-gstat (scipy ``pdist`` + comparisons).  This one classifies a pair by comparing the exact squared
pixel distance d2 = dx^2 + dy^2 (an integer) with the exact square of each edge (a rational), so it is independent of
sqrt / pdist rounding and settles what the two possible edge conventions give for pairs that sit EXACTLY on a bin edge
-- the only pairs on which they differ:

  rule "left"  (restated from scikit-gstat 1.0.x ``Variogram._calc_groups``):  edges[k-1] <= d <  edges[k]
  rule "right" (the other reading of "upper bin edges"):                       edges[k-1] <  d <= edges[k]

with edges[-1] := 0, pairs beyond the last edge dropped.  ``tests/golden/variogram_edges.json`` (written by
``python -m oracle.variogram_exact``) holds hand-checkable cases under BOTH rules; the product's rule is one flag
(``xdem_b200.spatialstats.LAG_EDGE_RULE``).
"""

from __future__ import annotations

import json
import os
from fractions import Fraction
from typing import Iterable, Sequence


def classify(d2: int, edges_sq: Sequence[Fraction], rule: str) -> int:
    """Lag class of a pair with integer squared pixel distance d2 (in units of gsd^2), or -1 if dropped."""
    lower = Fraction(0)
    for k, upper in enumerate(edges_sq):
        if rule == "left":
            inside = lower <= d2 < upper
        elif rule == "right":
            inside = lower < d2 <= upper
        else:
            raise ValueError(rule)
        if inside:
            return k
        lower = upper
    return -1


def pair_table(points: Sequence[tuple[int, int]], values: Sequence[Fraction], edges_over_gsd_sq: Iterable[Fraction],
               rule: str) -> tuple[list[int], list[Fraction]]:
    """(count, sum of squared differences) per lag class; ``edges_over_gsd_sq`` = (edge / gsd)^2 as exact rationals."""
    e2 = [Fraction(e) for e in edges_over_gsd_sq]
    count = [0] * len(e2)
    sumsq = [Fraction(0)] * len(e2)
    n = len(points)
    for i in range(n):
        for j in range(i + 1, n):
            dx, dy = points[i][0] - points[j][0], points[i][1] - points[j][1]
            k = classify(dx * dx + dy * dy, e2, rule)
            if k >= 0:
                count[k] += 1
                dv = Fraction(values[i]) - Fraction(values[j])
                sumsq[k] += dv * dv
    return count, sumsq


def _cases() -> list[dict]:
    """Small configurations with pairs exactly on bin edges.  Values are small integers so that every quantity can be
    checked by hand; edges are given as (edge/gsd)^2 (exact) and as floats edge = gsd * sqrt(that)."""
    cases = []
    # 3x3 block of a unit grid: 12 pairs at d=1, 8 at sqrt2, 6 at 2, 8 at sqrt5, 2 at sqrt8
    pts = [(x, y) for y in range(3) for x in range(3)]
    vals = [0, 1, 2, 3, 4, 5, 6, 7, 8]
    cases.append({"name": "3x3_integer_edges", "points": pts, "values": vals, "edges_sq": [1, 4, 9],
                  "comment": "edges 1, 2, 3 (in pixels): the d=1 and d=2 pairs sit on edges"})
    # the reference's default sqrt(2)-geometric edges (spatialstats.py:1439-1449): sqrt2, 2, 2 sqrt2, ... -> edge^2 = 2,4,8
    cases.append({"name": "3x3_default_sqrt2_edges", "points": pts, "values": vals, "edges_sq": [2, 4, 8],
                  "comment": "sqrt(2)*gsd*sqrt(2)^k: EVERY edge is a lattice distance (d2 = 2, 4, 8)"})
    # a 5-point cross + far corner; 'even' style edges at multiples of 5/2
    pts2 = [(0, 0), (3, 4), (6, 8), (0, 5), (5, 0), (10, 0)]
    vals2 = [2, -1, 4, 0, 3, 7]
    cases.append({"name": "pythagorean_5_10", "points": pts2, "values": vals2, "edges_sq": [25, 100, 225],
                  "comment": "3-4-5 triangles: many pairs at exactly d=5 and d=10"})
    # no pair on an edge: both rules must agree
    cases.append({"name": "3x3_off_lattice_edges", "points": pts, "values": vals,
                  "edges_sq": [Fraction(3, 2), Fraction(9, 2), Fraction(17, 2)],
                  "comment": "edges between lattice distances: the rules coincide"})
    return cases


def build_golden() -> dict:
    out = {"generator": "python -m oracle.variogram_exact", "cases": []}
    for c in _cases():
        rec = {"name": c["name"], "comment": c["comment"], "points": [list(p) for p in c["points"]],
               "values": list(c["values"]), "edges_sq_num_den": [[Fraction(e).numerator, Fraction(e).denominator]
                                                               for e in c["edges_sq"]]}
        for rule in ("left", "right"):
            cnt, ssq = pair_table(c["points"], [Fraction(v) for v in c["values"]], c["edges_sq"], rule)
            rec[rule] = {"count": cnt, "sumsq": [[s.numerator, s.denominator] for s in ssq]}
        out["cases"].append(rec)
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                        "variogram_edges.json")
    with open(path, "w") as f:
        json.dump(build_golden(), f, indent=1)
    print(path)
