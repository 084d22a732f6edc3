"""Stand-in for the third-party `bezier` package (imported by scenarios/ep_rand_bezier.py:2 and
scenarios/obstacles/o_ep_rand_bezier.py:3; not installed in this image) — TEST INFRASTRUCTURE ONLY.

The reference uses exactly one thing of it: `bezier.Curve(nodes, degree=2).evaluate_multi(pts)`, the points of a Bezier
curve with control points `nodes[:, k]` at parameters `pts`.  That is the published Bernstein form
    B(s) = sum_k C(n, k) (1 - s)^(n - k) s^k P_k
restated here so that the UNMODIFIED reference scenarios run and their goal sequences can be recorded as golden fixtures."""
from math import comb

import numpy as np


class Curve:
    def __init__(self, nodes, degree=None, **kwargs):
        self.nodes = np.asarray(nodes, dtype=np.float64)
        self.degree = self.nodes.shape[1] - 1 if degree is None else int(degree)
        if self.nodes.shape[1] != self.degree + 1:
            raise ValueError("bezier stub: nodes must have degree + 1 columns")

    def evaluate_multi(self, s_vals):
        s = np.asarray(s_vals, dtype=np.float64)
        n = self.degree
        out = np.zeros((self.nodes.shape[0], s.shape[0]))
        for k in range(n + 1):
            out += np.outer(self.nodes[:, k], comb(n, k) * (1.0 - s) ** (n - k) * s ** k)
        return out
