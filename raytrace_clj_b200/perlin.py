"""Perlin tables — mirror of reference src/raytrace_clj/perlin.clj:6-17.

The reference builds ``random-vectors`` (256 normalised rand-in-unit-sphere points) and ``perm-x/y/z`` (three
shuffles of 0..255) from the unseeded RNG when the namespace loads, so a render's noise pattern exists only in that
JVM's memory: the tables are MARSHALLED with the scene (``rt_scene_ext.perlin_vectors / perlin_perm``), never
regenerated on the device.  ``noise`` / ``turbulence`` (perlin.clj:45-64) themselves run on the GPU.
"""
from __future__ import annotations

import random
from dataclasses import dataclass

import numpy as np


@dataclass
class PerlinTables:
    vectors: np.ndarray   # [256, 3] float64, unit length
    perm: np.ndarray      # [3, 256] int32: perm-x, perm-y, perm-z


def make_tables(rng: random.Random | None = None) -> PerlinTables:
    """perlin.clj:6-17 with a seeded generator (same construction: rejection-sampled ball point, normalised;
    three independent shuffles)."""
    rng = rng or random.Random(0)
    vecs = np.zeros((256, 3))
    for i in range(256):
        while True:
            p = np.array([2.0 * rng.random() - 1.0 for _ in range(3)])
            if not (float(np.dot(p, p)) >= 1.0):
                break
        vecs[i] = p * (1.0 / float(np.sqrt(np.dot(p, p))))
    perm = np.zeros((3, 256), np.int32)
    for a in range(3):
        q = list(range(256))
        rng.shuffle(q)
        perm[a] = q
    return PerlinTables(vecs, perm)


# the tables of this process (the reference's are namespace-level defs); scenes that use Perlin textures marshal these
TABLES = make_tables(random.Random(20170415))
