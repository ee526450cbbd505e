"""Geometry records — mirror of reference src/raytrace_clj/hitable.clj.

``hit?`` runs on the GPU as a closest hit over the flattened leaves (brute force with Hitlist
semantics, hitable.clj:15-26, or the GPU BVH).  The records here are what the scene builders return
and what the marshaller walks: spheres (hitable.clj:141-259), rectangles (:269-363), the wrappers
FlipNormals / Translate / RotateY (:375-486), Box (:491-511), ConstantMedium (:516-543), Triangle
(:548-581) and the containers Hitlist (:15) and bvh-node (:97).  Nothing is intersected on the host.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, List

import math

import numpy as np

from .util import vec3

FLT_MAX = float(np.finfo(np.float32).max)


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


# ------------------------------------------------------------------------------------------------
# rectangles, wrappers, box, fog, triangle (hitable.clj:269-581)
# ------------------------------------------------------------------------------------------------
@dataclass(eq=False)
class RectXY:                 # hitable.clj:272
    x0: float
    y0: float
    x1: float
    y1: float
    k: float
    material: Any

    def bbox(self, t0, t1):   # hitable.clj:294-296
        return AABB(vec3(self.x0, self.y0, self.k - 0.0001), vec3(self.x1, self.y1, self.k + 0.0001))


@dataclass(eq=False)
class RectXZ:                 # hitable.clj:303
    x0: float
    z0: float
    x1: float
    z1: float
    k: float
    material: Any

    def bbox(self, t0, t1):   # hitable.clj:325-327
        return AABB(vec3(self.x0, self.k - 0.0001, self.z0), vec3(self.x1, self.k + 0.0001, self.z1))


@dataclass(eq=False)
class RectYZ:                 # hitable.clj:334
    y0: float
    z0: float
    y1: float
    z1: float
    k: float
    material: Any

    def bbox(self, t0, t1):   # hitable.clj:356-358
        return AABB(vec3(self.k - 0.0001, self.y0, self.z0), vec3(self.k + 0.0001, self.y1, self.z1))


@dataclass(eq=False)
class FlipNormals:            # hitable.clj:375
    item: Any

    def bbox(self, t0, t1):
        return self.item.bbox(t0, t1)


@dataclass(eq=False)
class Translate:              # hitable.clj:391
    item: Any
    offset: np.ndarray

    def bbox(self, t0, t1):   # hitable.clj:401-405
        b = self.item.bbox(t0, t1)
        return AABB(b.vmin + self.offset, b.vmax + self.offset)


@dataclass(eq=False)
class RotateY:                # hitable.clj:410
    obj: Any
    rotated_bbox: AABB
    sin_theta: float
    cos_theta: float

    def bbox(self, t0, t1):
        return self.rotated_bbox


def make_rotate_y(obj, theta):
    """hitable.clj:457-486 — the box of the 8 rotated corners of ``(bbox obj 0 1)``."""
    radians = theta * (math.pi / 180.0)
    cos_th, sin_th = math.cos(radians), math.sin(radians)
    box = obj.bbox(0, 1)
    new_min = vec3(FLT_MAX, FLT_MAX, FLT_MAX)
    new_max = vec3(-FLT_MAX, -FLT_MAX, -FLT_MAX)
    for x in (box.vmin[0], box.vmax[0]):
        for y in (box.vmin[1], box.vmax[1]):
            for z in (box.vmin[2], box.vmax[2]):
                c = vec3(cos_th * x + sin_th * z, y, -(sin_th * x) + cos_th * z)
                new_min = np.minimum(new_min, c)
                new_max = np.maximum(new_max, c)
    return RotateY(obj, AABB(new_min, new_max), sin_th, cos_th)


@dataclass(eq=False)
class Box:                    # hitable.clj:491
    p0: np.ndarray
    p1: np.ndarray
    sides: Hitlist

    def bbox(self, t0, t1):
        return AABB(self.p0, self.p1)


@dataclass(eq=False)
class ConstantMedium:         # hitable.clj:516
    boundary: Any
    density: float
    phase_fn: Any

    def bbox(self, t0, t1):
        return self.boundary.bbox(t0, t1)


@dataclass(eq=False)
class Triangle:               # hitable.clj:548
    v0: np.ndarray
    v1: np.ndarray
    v2: np.ndarray
    material: Any

    def bbox(self, t0, t1):   # hitable.clj:576-581
        e = vec3(0.0001, 0.0001, 0.0001)
        return AABB(np.minimum(np.minimum(self.v0, self.v1), self.v2) - e,
                    np.maximum(np.maximum(self.v0, self.v1), self.v2) + e)


def rect_xy(*, x0, y0, x1, y1, k, material):
    return RectXY(float(x0), float(y0), float(x1), float(y1), float(k), material)


def rect_xz(*, x0, z0, x1, z1, k, material):
    return RectXZ(float(x0), float(z0), float(x1), float(z1), float(k), material)


def rect_yz(*, y0, z0, y1, z1, k, material):
    return RectYZ(float(y0), float(z0), float(y1), float(z1), float(k), material)


def flip_normals(*, item):
    return FlipNormals(item)


def translate(*, item, offset):
    return Translate(item, offset)


def rotate_y(*, item, theta):
    return make_rotate_y(item, float(theta))


def box(*, p0, p1, material):
    """hitable.clj:498-511 — six rectangles in a Hitlist, the three 'low' faces flipped."""
    x0, y0, z0 = (float(v) for v in p0)
    x1, y1, z1 = (float(v) for v in p1)
    return Box(p0, p1, Hitlist([
        RectXY(x0, y0, x1, y1, z1, material),
        FlipNormals(RectXY(x0, y0, x1, y1, z0, material)),
        RectXZ(x0, z0, x1, z1, y1, material),
        FlipNormals(RectXZ(x0, z0, x1, z1, y0, material)),
        RectYZ(y0, z0, y1, z1, x1, material),
        FlipNormals(RectYZ(y0, z0, y1, z1, x0, material)),
    ]))


def constant_medium(*, boundary, density, albedo):
    from . import shader as shad

    return ConstantMedium(boundary, float(density), shad.isotropic(albedo=albedo))


def triangle(*, v0, v1, v2, material):
    return Triangle(v0, v1, v2, material)
