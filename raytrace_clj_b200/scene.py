"""Scene builders — mirror of reference src/raytrace_clj/scene.clj (all nine builders + the benchmark variants).

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

import numpy as np

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


# ------------------------------------------------------------------------------------------------
# the other scene builders of scene.clj
# ------------------------------------------------------------------------------------------------
def _camera_13_2_3(nx, ny, vfov, lookat=(0, 1, 0)):
    return cam.thin_lens_camera(lookfrom=vec3(13, 2, 3), lookat=vec3(*lookat), vup=vec3(0, 1, 0), vfov=vfov,
                                aspect=float(nx) / float(ny), aperture=0.0, focus_dist=10.0, t0=0.0, t1=1.0)


def _blue_sky():
    # scene.clj:65-70 etc.: a plain radius-1000 light sphere, colour 0.8 * (0.3, 0.5, 0.8)
    return hit.sphere(center=vec3(0, 0, 0), radius=1000,
                      material=shad.diffuse_light(tex=tex.constant(color=0.8 * vec3(0.3, 0.5, 0.8))))


def make_two_perlin_spheres(nx, ny, rng=None):
    """scene.clj:51-78 — two spheres with Perlin turbulence (scale 4, depth 7)."""
    rng = rng or random.Random(0)
    turb = tex.perlin_turbulence(scale=4, depth=7)
    return {"camera": _camera_13_2_3(nx, ny, 40),
            "world": hit.make_bvh([
                _blue_sky(),
                hit.sphere(center=vec3(0, -1000, 0), radius=1000, material=shad.lambertian(albedo=turb)),
                hit.sphere(center=vec3(0, 2, 0), radius=2, material=shad.lambertian(albedo=turb)),
            ], 0.0, 1.0, rng)}


def make_two_triangles(nx, ny, rng=None):
    """scene.clj:80-114 — two triangles seen down the -z axis under a constant sky."""
    rng = rng or random.Random(0)
    white = tex.constant(color=vec3(0.9, 0.9, 0.9))
    red = tex.constant(color=vec3(0.9, 0, 0))
    return {"camera": cam.thin_lens_camera(lookfrom=vec3(1, 1, -10), lookat=vec3(1, 1, 0), vup=vec3(0, 1, 0), vfov=20,
                                           aspect=float(nx) / float(ny), aperture=0.0, focus_dist=10.0, t0=0.0, t1=1.0),
            "world": hit.make_bvh([
                _blue_sky(),
                hit.triangle(v0=vec3(0, 0, 0), v1=vec3(0, 1, 0), v2=vec3(1, 0, 0), material=shad.lambertian(albedo=red)),
                hit.triangle(v0=vec3(1, 1, 0), v1=vec3(1, 2, 0), v2=vec3(2, 1, 0), material=shad.lambertian(albedo=white)),
            ], 0.0, 1.0, rng)}


def synthetic_earth(w=256, h=128):
    """Stand-in for the reference's "earth.png" (not in its repository: .gitignore:12-13): a deterministic
    land / sea / ice pattern, uint8 [h, w, 3]."""
    v, u = np.meshgrid(np.linspace(0, 1, h, endpoint=False), np.linspace(0, 1, w, endpoint=False), indexing="ij")
    land = (np.sin(9 * u * 2 * np.pi) * np.cos(5 * v * np.pi) + np.sin(3 * u * 2 * np.pi + 1.0) * np.sin(7 * v * np.pi)) > 0.25
    img = np.zeros((h, w, 3), np.uint8)
    img[...] = (20, 60, 160)
    img[land] = (40, 140, 50)
    img[(v < 0.08) | (v > 0.92)] = (240, 240, 250)
    return img


def make_textured_sphere(nx, ny, rng=None, earth=None):
    """scene.clj:116-151 — an image-mapped UVSphere (flip-texture-v of "earth.png"; a synthetic map stands in)."""
    rng = rng or random.Random(0)
    earth_tex = tex.flip_texture_v(tex=tex.image_map(image=synthetic_earth() if earth is None else earth))
    return {"camera": _camera_13_2_3(nx, ny, 15),
            "world": hit.make_bvh([
                _blue_sky(),
                hit.sphere(center=vec3(0, -10, 0), radius=10, material=shad.lambertian(albedo=_checker())),
                hit.uv_sphere(center=vec3(0, 1, 0), radius=1, material=shad.lambertian(albedo=earth_tex)),
            ], 0.0, 1.0, rng)}


def make_subsurface_sphere(nx, ny, rng=None):
    """scene.clj:153-189 — a glass ball filled with a blue constant medium."""
    rng = rng or random.Random(0)
    ball = hit.sphere(center=vec3(0, 2, 0), radius=2, material=shad.dielectric(ri=1.5))
    medium = hit.constant_medium(boundary=ball, density=0.9, albedo=tex.constant(color=vec3(0.2, 0.4, 0.9)))
    return {"camera": _camera_13_2_3(nx, ny, 40),
            "world": hit.make_bvh([
                _blue_sky(),
                hit.sphere(center=vec3(0, -10, 0), radius=10, material=shad.lambertian(albedo=_checker())),
                medium,
                ball,
            ], 0.0, 1.0, rng)}


def make_example_light(nx, ny, rng=None):
    """scene.clj:191-228 — two grey spheres lit by a sphere light and a rectangular area light (no sky)."""
    rng = rng or random.Random(0)
    gray = shad.lambertian(albedo=tex.constant(color=vec3(0.6, 0.6, 0.6)))
    light = shad.diffuse_light(tex=tex.constant(color=vec3(4, 4, 4)))
    return {"camera": _camera_13_2_3(nx, ny, 40),
            "world": hit.make_bvh([
                hit.sphere(center=vec3(0, -1000, 0), radius=1000, material=gray),
                hit.sphere(center=vec3(0, 2, 0), radius=2, material=gray),
                hit.sphere(center=vec3(0, 7, 0), radius=2, material=light),
                hit.rect_xy(x0=3, y0=1, x1=5, y1=3, k=-2, material=light),
            ], 0.0, 1.0, rng)}


def make_cornell_box(nx, ny, classic=True, rng=None):
    """scene.clj:230-316 — the Cornell box; classic: small light + two solid boxes, else big light + two fog boxes."""
    rng = rng or random.Random(0)
    red = shad.lambertian(albedo=tex.constant(color=vec3(0.65, 0.05, 0.05)))
    white = shad.lambertian(albedo=tex.constant(color=vec3(0.73, 0.73, 0.73)))
    green = shad.lambertian(albedo=tex.constant(color=vec3(0.12, 0.45, 0.15)))
    light = shad.diffuse_light(tex=tex.constant(color=vec3(7, 7, 7)))

    def block(p1, theta, offset):
        return hit.translate(item=hit.rotate_y(item=hit.box(p0=vec3(0, 0, 0), p1=vec3(*p1), material=white), theta=theta),
                             offset=vec3(*offset))

    short, tall = block((165, 165, 165), -18.0, (130, 0, 65)), block((165, 330, 165), 15.0, (265, 0, 295))
    if not classic:
        short = hit.constant_medium(boundary=short, density=0.01, albedo=tex.constant(color=vec3(1, 1, 1)))
        tall = hit.constant_medium(boundary=tall, density=0.01, albedo=tex.constant(color=vec3(0, 0, 0)))
    return {"camera": cam.thin_lens_camera(lookfrom=vec3(278, 278, -800), lookat=vec3(278, 278, 0), vup=vec3(0, 1, 0),
                                           vfov=40, aspect=float(nx) / float(ny), aperture=0.0, focus_dist=10.0, t0=0.0, t1=1.0),
            "world": hit.make_bvh([
                hit.flip_normals(item=hit.rect_yz(y0=0, z0=0, y1=555, z1=555, k=555, material=green)),
                hit.rect_yz(y0=0, z0=0, y1=555, z1=555, k=0, material=red),
                (hit.rect_xz(x0=213, z0=227, x1=343, z1=332, k=554, material=light) if classic
                 else hit.rect_xz(x0=113, z0=127, x1=443, z1=432, k=554, material=light)),
                hit.flip_normals(item=hit.rect_xz(x0=0, z0=0, x1=555, z1=555, k=555, material=white)),
                hit.rect_xz(x0=0, z0=0, x1=555, z1=555, k=0, material=white),
                hit.flip_normals(item=hit.rect_xy(x0=0, y0=0, x1=555, y1=555, k=555, material=white)),
                short,
                tall,
            ], 0.0, 1.0, rng)}


def make_final(nx, ny, rng=None, earth=None, nb=20, ns=1000):
    """scene.clj:415-492 — the book-2 final scene (the scene `-main` hard-codes, core.clj:90): a ground of nb x nb
    boxes, an area light, a moving sphere, glass, fuzz-10 metal, a blue subsurface ball, an overall haze, an
    image-mapped "earth" (synthetic map: the reference's earth.png is not in its repository), a marble sphere and
    ns small spheres in a rotated, translated BVH."""
    rng = rng or random.Random(0)
    rand = rng.random
    white = shad.lambertian(albedo=tex.constant(color=vec3(0.73, 0.73, 0.73)))
    ground = shad.lambertian(albedo=tex.constant(color=vec3(0.48, 0.83, 0.53)))
    orange = shad.lambertian(albedo=tex.constant(color=vec3(0.7, 0.3, 0.1)))
    light = shad.diffuse_light(tex=tex.constant(color=vec3(7, 7, 7)))
    glass = shad.dielectric(ri=1.5)
    metal = shad.metal(albedo=tex.constant(color=vec3(0.8, 0.8, 0.9)), fuzz=10)
    bndry = hit.sphere(center=vec3(360, 150, 145), radius=70, material=glass)
    earth_m = shad.lambertian(albedo=tex.flip_texture_v(tex=tex.image_map(image=synthetic_earth() if earth is None else earth)))
    marble = shad.lambertian(albedo=tex.marble(scale=0.1, depth=4))
    boxes = []
    for i in range(nb):
        for j in range(nb):
            w = 100
            p0 = vec3(-1000 + i * w, 0, -1000 + j * w)
            p1 = p0 + vec3(w, 100 * (rand() + 0.01), w)
            boxes.append(hit.box(p0=p0, p1=p1, material=ground))
    small = [hit.sphere(center=165.0 * vec3(rand(), rand(), rand()), radius=10, material=white) for _ in range(ns)]
    return {"camera": cam.thin_lens_camera(lookfrom=vec3(478, 278, -600), lookat=vec3(278, 278, 0), vup=vec3(0, 1, 0),
                                           vfov=40, aspect=float(nx) / float(ny), aperture=0.0, focus_dist=10.0, t0=0.0, t1=1.0),
            "world": hit.make_bvh([
                hit.make_bvh(boxes, 0.0, 1.0, rng),
                hit.rect_xz(x0=123, z0=147, x1=423, z1=412, k=554, material=light),
                hit.moving_sphere(center0=vec3(400, 400, 200), t0=0, center1=vec3(430, 400, 200), t1=1, radius=50, material=orange),
                hit.sphere(center=vec3(260, 150, 45), radius=50, material=glass),
                hit.sphere(center=vec3(0, 150, 145), radius=50, material=metal),
                bndry,
                hit.constant_medium(boundary=bndry, density=0.2, albedo=tex.constant(color=vec3(0.2, 0.4, 0.9))),
                hit.constant_medium(boundary=hit.sphere(center=vec3(0, 0, 0), radius=5000, material=glass),
                                    density=0.0001, albedo=tex.constant(color=vec3(1, 1, 1))),
                hit.uv_sphere(center=vec3(400, 200, 400), radius=100, material=earth_m),
                hit.sphere(center=vec3(220, 280, 300), radius=80, material=marble),
                hit.translate(item=hit.rotate_y(item=hit.make_bvh(small, 0.0, 1.0, rng), theta=15), offset=vec3(-100, 270, 395)),
            ], 0.0, 1.0, rng)}
