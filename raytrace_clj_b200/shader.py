"""Material records — mirror of reference src/raytrace_clj/shader.clj.

``scatter`` / ``emitted`` (shader.clj:22-24) run on the GPU.  Isotropic (shader.clj:129-143) is
only reachable through ConstantMedium (hitable.clj:516-543).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any


@dataclass(eq=False)
class Lambertian:             # shader.clj:29-36
    albedo: Any


@dataclass(eq=False)
class Metal:                  # shader.clj:46-59 (fuzz is not clamped)
    albedo: Any
    fuzz: float


@dataclass(eq=False)
class Dielectric:             # shader.clj:76-104
    ri: float


@dataclass(eq=False)
class DiffuseLight:           # shader.clj:114-119
    tex: Any


def lambertian(*, albedo):
    return Lambertian(albedo)


def metal(*, albedo, fuzz):
    return Metal(albedo, float(fuzz))


def dielectric(*, ri):
    return Dielectric(float(ri))


def diffuse_light(*, tex):
    return DiffuseLight(tex)


@dataclass(eq=False)
class Isotropic:              # shader.clj:129-138 (the scattered ray's time is the hit's t, as written there)
    albedo: Any


def isotropic(*, albedo):
    return Isotropic(albedo)
