"""Texture records — mirror of reference src/raytrace_clj/texture.clj (the subset on the hot path).

``sample`` (texture.clj:8-9) runs on the GPU.  PerlinNoise / PerlinTurbulence / Marble /
FlipTexture / ImageMap (texture.clj:60-138) are outside the accelerated path: the marshaller
rejects them (no CPU fallback).
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
