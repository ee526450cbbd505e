"""Texture records — mirror of reference src/raytrace_clj/texture.clj (the subset on the hot path).

``sample`` (texture.clj:8-9) runs on the GPU for every record here.  The Perlin tables
(perlin.clj:6-17: 256 unit vectors + three permutations, drawn from the unseeded RNG at namespace load in
the reference) are marshalled with the scene (``perlin.PerlinTables``), never regenerated on the device.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any

import numpy as np


@dataclass(eq=False)
class Constant:               # texture.clj:14-16
    color: np.ndarray


@dataclass(eq=False)
class UVGradient:             # texture.clj:26-34
    co: np.ndarray
    cu: np.ndarray
    cv: np.ndarray
    cuv: np.ndarray


@dataclass(eq=False)
class Checkerboard:           # texture.clj:44-50
    tex0: Any
    tex1: Any
    scale: float


def constant(*, color):
    return Constant(color)


def uv_gradient(*, co, cu, cv, cuv):
    return UVGradient(co, cu, cv, cuv)


def checkerboard(*, tex0, tex1, scale):
    return Checkerboard(tex0, tex1, float(scale))


@dataclass(eq=False)
class PerlinNoise:            # texture.clj:60-64
    scale: float


@dataclass(eq=False)
class PerlinTurbulence:       # texture.clj:74-78
    scale: float
    depth: int


@dataclass(eq=False)
class Marble:                 # texture.clj:88-93
    scale: float
    depth: int


@dataclass(eq=False)
class FlipTextureU:           # texture.clj:103-106
    tex: Any


@dataclass(eq=False)
class FlipTextureV:           # texture.clj:113-116
    tex: Any


@dataclass(eq=False)
class ImageMap:               # texture.clj:126-133; image = uint8 [h, w, 3], row 0 = top (imagez get-pixel x y)
    image: np.ndarray


def perlin_noise(*, scale):
    return PerlinNoise(float(scale))


def perlin_turbulence(*, scale, depth):
    return PerlinTurbulence(float(scale), int(depth))


def marble(*, scale, depth):
    return Marble(float(scale), int(depth))


def flip_texture_u(*, tex):
    return FlipTextureU(tex)


def flip_texture_v(*, tex):
    return FlipTextureV(tex)


def image_map(*, image=None, filename=None):
    """texture.clj:135-138 loads a file through imagez; here the caller passes the decoded pixels
    (``image``) or a PPM/PNG filename readable by ``ppm.load``."""
    if image is None:
        from . import ppm

        image = ppm.load(filename)
    return ImageMap(np.ascontiguousarray(image, np.uint8))
