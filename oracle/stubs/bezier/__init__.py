"""Stand-in for `bezier` (imported by scenarios/ep_rand_bezier.py:2 via mix.py:13).
The two bezier scenarios are out of scope; constructing a Curve raises."""


class Curve:
    def __init__(self, *a, **k):
        raise NotImplementedError("bezier is not available in this image (stub)")
