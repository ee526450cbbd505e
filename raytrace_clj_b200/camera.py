"""Camera records and constructors — mirror of reference src/raytrace_clj/camera.clj.

Construction (basis vectors, lower-left corner) stays on the host in double precision;
``get-ray`` (camera.clj:8-16, 35-48) runs on the GPU from the marshalled 24-float record.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from .util import normalise


@dataclass
class PinholeCamera:          # camera.clj:8
    origin: np.ndarray
    lleft: np.ndarray
    horiz: np.ndarray
    vert: np.ndarray


@dataclass
class ThinLensCamera:         # camera.clj:35
    origin: np.ndarray
    lleft: np.ndarray
    horiz: np.ndarray
    vert: np.ndarray
    u: np.ndarray
    v: np.ndarray
    w: np.ndarray
    aperture: float
    t0: float
    t1: float


def pinhole_camera(*, lookfrom, lookat, vup, vfov, aspect) -> PinholeCamera:
    """camera.clj:18-33."""
    theta = vfov * (math.pi / 180.0)
    half_height = math.tan(theta / 2.0)
    half_width = aspect * half_height
    w = normalise(lookfrom - lookat)
    u = normalise(np.cross(vup, w))
    v = np.cross(w, u)
    return PinholeCamera(lookfrom.copy(), lookfrom - (half_width * u + half_height * v + w),
                         2.0 * half_width * u, 2.0 * half_height * v)


def thin_lens_camera(*, lookfrom, lookat, vup, vfov, aspect, aperture, focus_dist, t0, t1) -> ThinLensCamera:
    """camera.clj:50-66."""
    theta = vfov * (math.pi / 180.0)
    half_height = math.tan(theta / 2.0)
    half_width = aspect * half_height
    w = normalise(lookfrom - lookat)
    u = normalise(np.cross(vup, w))
    v = np.cross(w, u)
    lleft = lookfrom - (focus_dist * half_width * u + focus_dist * half_height * v + focus_dist * w)
    return ThinLensCamera(lookfrom.copy(), lleft, 2.0 * focus_dist * half_width * u,
                          2.0 * focus_dist * half_height * v, u, v, w, float(aperture), float(t0), float(t1))
