"""vec3 / ray helpers — host-side mirror of reference src/raytrace_clj/util.clj.

Only scene construction runs here (double precision numpy, like vectorz Vector3); every
per-ray use of these helpers in the reference runs on the GPU instead.
"""
from __future__ import annotations

import numpy as np


def vec3(a, b, c) -> np.ndarray:
    """util.clj:5-11 — three doubles (ints / ratios are widened, util_test.clj:7-25)."""
    return np.array([float(a), float(b), float(c)], dtype=np.float64)


def ray(origin, direction, t):
    """util.clj:13-16 — a ray is a map; the direction is NOT normalised."""
    return {"origin": origin, "direction": direction, "time": t}


def point_at_parameter(r, t):
    """util.clj:18-22 — direction * t + origin."""
    return r["direction"] * float(t) + r["origin"]


def magnitude(v) -> float:
    return float(np.sqrt(np.dot(v, v)))


def normalise(v) -> np.ndarray:
    d = magnitude(v)
    return v * (1.0 / d) if d > 0 else v.copy()
