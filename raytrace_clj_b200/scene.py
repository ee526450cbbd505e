"""Scene builders — mirror of reference src/raytrace_clj/scene.clj (the sphere-only scenes).

In the deployed drop-in these stay in Clojure (north_star: "the scene-building code … stay");
they are restated here because no JVM exists in the build environment and the harness needs the
benchmark scenes.  The reference draws from the unseeded ``clojure.core/rand``; here every
builder takes a ``random.Random`` so scenes are reproducible (same distributions, same draw
order, same rejection rule).

Builders return ``{"camera": …, "world": …}`` like the reference (scene.clj:16,26).
"""
from __future__ import annotations

import math
import random

from . import camera as cam
from . import hitable as hit
from . import shader as shad
from . import texture as tex
from .util import magnitude, vec3


def _sky():
    # scene.clj:336-344 — radius-1000 UVSphere with a DiffuseLight gradient (the "sky dome")
    return hit.uv_sphere(
        center=vec3(0, 0, 0), radius=1000,
        material=shad.diffuse_light(tex=tex.uv_gradient(co=vec3(1, 1, 1), cu=vec3(1, 1, 1),
                                                        cv=vec3(0.5, 0.7, 1.0), cuv=vec3(0.5, 0.7, 1.0))))


def _checker():
    return tex.checkerboard(tex0=tex.constant(color=vec3(0.2, 0.3, 0.1)),
                            tex1=tex.constant(color=vec3(0.9, 0.9, 0.9)), scale=10)


def make_two_spheres(nx, ny, rng=None):
    """scene.clj:9-49 — two touching spheres under the sky dome."""
    rng = rng or random.Random(0)
    return {
        "camera": cam.thin_lens_camera(lookfrom=vec3(13, 2, 3), lookat=vec3(0, 1, 0), vup=vec3(0, 1, 0), vfov=40,
                                       aspect=float(nx) / float(ny), aperture=0.0, focus_dist=10.0, t0=0.0, t1=1.0),
        "world": hit.make_bvh([
            _sky(),
            hit.sphere(center=vec3(0, -10, 0), radius=10, material=shad.lambertian(albedo=_checker())),
            hit.uv_sphere(center=vec3(0, 2, 0), radius=2,
                          material=shad.lambertian(albedo=tex.uv_gradient(co=vec3(0, 1, 0), cu=vec3(0, 1, 1),
                                                                          cv=vec3(1, 0, 1), cuv=vec3(1, 0, 0)))),
        ], 0.0, 1.0, rng),
    }


def _random_scene_camera(nx, ny):
    # scene.clj:321-330 — aperture is 0.0: a thin-lens camera that still draws its RNG values
    return cam.thin_lens_camera(lookfrom=vec3(13, 2, 3), lookat=vec3(0, 0, 0), vup=vec3(0, 1, 0), vfov=20,
                                aspect=float(nx) / float(ny), aperture=0.0, focus_dist=10.0, t0=0.0, t1=1.0)


def _heroes():
    # scene.clj:335-364
    return [
        _sky(),
        hit.sphere(center=vec3(0, -1000, 0), radius=1000, material=shad.lambertian(albedo=_checker())),
        hit.sphere(center=vec3(0, 1, 0), radius=1, material=shad.dielectric(ri=1.5)),
        hit.sphere(center=vec3(-4, 1, 0), radius=1,
                   material=shad.lambertian(albedo=tex.constant(color=vec3(0.4, 0.2, 0.1)))),
        hit.sphere(center=vec3(4, 1, 0), radius=1,
                   material=shad.metal(albedo=tex.constant(color=vec3(0.7, 0.6, 0.5)), fuzz=0.0)),
    ]


def random_scene_objects(n, moving, rng, p_diffuse=0.8, p_metal=0.95):
    """The object list of scene.clj:332-411 before it is wrapped in the BVH.

    Draw order per grid cell (a outer, b inner): cx, cz, choose-mat (the `:let` runs before the
    `:when` filter), then for a kept cell: [c1y] + 6 colour draws (diffuse), 3 colour + fuzz
    (metal), nothing (glass).  ``p_diffuse`` / ``p_metal`` are the 0.8 / 0.95 thresholds.
    """
    rand = rng.random
    objs = _heroes()
    for a in range(-n, n):
        for b in range(-n, n):
            center = vec3(a + 0.9 * rand(), 0.2, b + 0.9 * rand())
            choose_mat = rand()
            if not (magnitude(center - vec3(4, 0.2, 0)) > 0.9):
                continue
            if choose_mat < p_diffuse:
                if moving:
                    center1 = center + vec3(0, 0.5 * rand(), 0)
                    albedo = tex.constant(color=vec3(rand() * rand(), rand() * rand(), rand() * rand()))
                    objs.append(hit.moving_sphere(center0=center, t0=0.0, center1=center1, t1=1.0, radius=0.2,
                                                  material=shad.lambertian(albedo=albedo)))
                else:
                    albedo = tex.constant(color=vec3(rand() * rand(), rand() * rand(), rand() * rand()))
                    objs.append(hit.sphere(center=center, radius=0.2, material=shad.lambertian(albedo=albedo)))
            elif choose_mat < p_metal:
                albedo = tex.constant(color=vec3(0.5 * (1 + rand()), 0.5 * (1 + rand()), 0.5 * (1 + rand())))
                objs.append(hit.sphere(center=center, radius=0.2,
                                       material=shad.metal(albedo=albedo, fuzz=0.5 * rand())))
            else:
                objs.append(hit.sphere(center=center, radius=0.2, material=shad.dielectric(ri=1.5)))
    return objs


def make_random_scene(nx, ny, n=11, moving=True, rng=None):
    """scene.clj:318-412 — the book-cover random-spheres scene (BASELINE configs 1-3)."""
    rng = rng or random.Random(1)
    objs = random_scene_objects(n, moving, rng)
    return {"camera": _random_scene_camera(nx, ny), "world": hit.make_bvh(objs, 0.0, 1.0, rng)}


def make_material_stress_scene(nx, ny, n=11, rng=None):
    """BASELINE config 4: the scene.clj:318-412 generator with the material mix forced to
    ~10 % Lambert / 45 % metal / 45 % glass (divergence and depth-50 stress)."""
    rng = rng or random.Random(4)
    objs = random_scene_objects(n, False, rng, p_diffuse=0.10, p_metal=0.55)
    return {"camera": _random_scene_camera(nx, ny), "world": hit.make_bvh(objs, 0.0, 1.0, rng)}


def make_scale_sweep_scene(nx, ny, n_small, rng=None):
    """BASELINE config 5: the 5 hero objects + ``n_small`` static r=0.2 spheres placed with the
    scene.clj:369-375 rule on a ceil(sqrt(n_small)) grid (same material mix, no motion)."""
    rng = rng or random.Random(5)
    rand = rng.random
    side = int(math.ceil(math.sqrt(n_small)))
    half = side // 2
    objs = _heroes()
    count = 0
    for a in range(-half, side - half):
        for b in range(-half, side - half):
            if count >= n_small:
                break
            center = vec3(a + 0.9 * rand(), 0.2, b + 0.9 * rand())
            choose_mat = rand()
            if not (magnitude(center - vec3(4, 0.2, 0)) > 0.9):
                continue
            count += 1
            if choose_mat < 0.8:
                albedo = tex.constant(color=vec3(rand() * rand(), rand() * rand(), rand() * rand()))
                objs.append(hit.sphere(center=center, radius=0.2, material=shad.lambertian(albedo=albedo)))
            elif choose_mat < 0.95:
                albedo = tex.constant(color=vec3(0.5 * (1 + rand()), 0.5 * (1 + rand()), 0.5 * (1 + rand())))
                objs.append(hit.sphere(center=center, radius=0.2,
                                       material=shad.metal(albedo=albedo, fuzz=0.5 * rand())))
            else:
                objs.append(hit.sphere(center=center, radius=0.2, material=shad.dielectric(ri=1.5)))
    # a flat Hitlist: 100k-leaf BVH construction in Python is pointless for a brute-force renderer
    return {"camera": _random_scene_camera(nx, ny), "world": hit.hitlist(items=objs)}
