"""Geometry records — mirror of reference src/raytrace_clj/hitable.clj (spheres + containers).

``hit?`` runs on the GPU as a brute-force closest hit over the flattened leaves
(Hitlist semantics, hitable.clj:15-26).  The BVH (hitable.clj:97-123) is kept only as the
container the scene builders return, so the marshaller has the same tree to flatten that the
JVM-side marshaller would see; rects, boxes, instances, fog and triangles
(hitable.clj:269-581) are outside the accelerated path and are rejected by the marshaller.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, List

import numpy as np

from .util import vec3


@dataclass(eq=False)
class AABB:                   # hitable.clj:36
    vmin: np.ndarray
    vmax: np.ndarray


def make_surrounding_bbox(box0: AABB, box1: AABB) -> AABB:
    """hitable.clj:87-92."""
    return AABB(np.minimum(box0.vmin, box1.vmin), np.maximum(box0.vmax, box1.vmax))


@dataclass(eq=False)
class Sphere:                 # hitable.clj:180
    center: np.ndarray
    radius: float
    material: Any

    def bbox(self, t0, t1):   # hitable.clj:208-211
        r = vec3(self.radius, self.radius, self.radius)
        return AABB(self.center - r, self.center + r)


@dataclass(eq=False)
class UVSphere(Sphere):       # hitable.clj:141
    pass


def center_at_time(center0, t0, center1, t1, t):
    """hitable.clj:219-222 — lerp(center0, center1, (t - t0) / (t1 - t0))."""
    f = (t - t0) / (t1 - t0)
    return center0 * (1.0 - f) + center1 * f


@dataclass(eq=False)
class MovingSphere:           # hitable.clj:224
    center0: np.ndarray
    t0: float
    center1: np.ndarray
    t1: float
    radius: float
    material: Any

    def bbox(self, t_start, t_end):   # hitable.clj:252-259
        r = vec3(self.radius, self.radius, self.radius)
        cs = center_at_time(self.center0, self.t0, self.center1, self.t1, t_start)
        ce = center_at_time(self.center0, self.t0, self.center1, self.t1, t_end)
        return make_surrounding_bbox(AABB(cs - r, cs + r), AABB(ce - r, ce + r))


@dataclass(eq=False)
class Hitlist:                # hitable.clj:15
    items: List[Any]


@dataclass(eq=False)
class BvhNode:                # hitable.clj:97 (record `bvh-node`)
    left: Any
    right: Any
    box: AABB

    def bbox(self, t0, t1):
        return self.box


def sphere(*, center, radius, material):
    return Sphere(center, float(radius), material)


def uv_sphere(*, center, radius, material):
    return UVSphere(center, float(radius), material)


def moving_sphere(*, center0, t0, center1, t1, radius, material):
    return MovingSphere(center0, float(t0), center1, float(t1), float(radius), material)


def hitlist(*, items):
    return Hitlist(list(items))


def make_bvh(hitable_list, t0, t1, rng):
    """hitable.clj:108-123 — random axis, sort by bbox vmin[axis], split at n/2 (left gets
    ceil(n/2)); a 1-element node stores the same object as both children."""
    axis = rng.randrange(3)
    my_list = sorted(hitable_list, key=lambda h: float(h.bbox(t0, t1).vmin[axis]))
    n = len(my_list)
    if n == 1:
        L = my_list[0]
        return BvhNode(L, L, L.bbox(t0, t1))
    if n == 2:
        L, R = my_list
        return BvhNode(L, R, make_surrounding_bbox(L.bbox(t0, t1), R.bbox(t0, t1)))
    k = -(-n // 2)  # (split-at (/ n 2) ...) with a ratio n/2 takes ceil(n/2) items
    L = make_bvh(my_list[:k], t0, t1, rng)
    R = make_bvh(my_list[k:], t0, t1, rng)
    return BvhNode(L, R, make_surrounding_bbox(L.bbox(t0, t1), R.bbox(t0, t1)))
