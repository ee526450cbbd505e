"""Regenerates the fixtures in this directory:  python tests/golden/make_golden.py

What they are — and are not.  The reference (Clojure on the JVM) cannot run in the build environment, so these
vectors are produced by the double-precision CPU restatement in oracle/ (each function there cites the reference
lines it follows), NOT by the reference itself: they pin the restatement against regression and give the `-m gpu`
parity tests committed numbers to hit without recomputing them.  The only vectors whose expected values come from
the reference's own test files are in `kat_reference_tests.json` (hitable_test.clj:8-59,61-103, util_test.clj:44-49).

  hits_random_scene.npz    4096 rays against make-random-scene (scene seed 1, n = 11, moving): camera rays, rays
                           from surfaces, origins inside the scene volume with ray times, silhouette-grazing rays;
                           closest hit t (float64) and sphere id (Hitlist.hit?, hitable.clj:15-26)
  shade_random_scene.npz   3072 scatter / emitted evaluations with explicit random inputs (shader.clj:29-119)
  image_random_scene.npz   two 64x40 renders at 512 spp with different seeds (linear float sums / spp): the
                           second one measures the Monte-Carlo noise the statistical comparison needs
  kat_reference_tests.json the reference's own known answers for this path
"""
import itertools
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
import raytrace_clj_b200 as rt  # noqa: E402
from helpers import camera_rays  # noqa: E402

FMAX = float(np.finfo(np.float32).max)


def scene():
    sc = rt.scene.make_random_scene(1200, 800, 11, True, random.Random(1))
    flat = rt.native.marshal_world(sc["world"])
    cam_type, cam = rt.native.marshal_camera(sc["camera"])
    return flat, cam_type, cam


def hit_rays(flat, cam, S):
    rng = np.random.default_rng(2026)
    n = 1024
    o1, d1, t1 = camera_rays(cam, 1200, 800, n, rng)
    t, _ = S.hit(o1, d1, t1, 0.001, FMAX)
    o2 = (o1.astype(np.float64) + t[:, None] * d1.astype(np.float64)).astype(np.float32)
    d2 = rng.normal(size=o2.shape).astype(np.float32) * rng.uniform(0.05, 2.0, size=(n, 1)).astype(np.float32)
    o3 = rng.uniform(-15, 15, size=(n, 3)).astype(np.float32)
    o3[:, 1] = rng.uniform(-1, 3, size=n)
    d3 = rng.normal(size=o3.shape).astype(np.float32)
    t3 = rng.random(n).astype(np.float32)
    k = rng.integers(0, flat.n_spheres, n)
    c = flat.center0_r[k, :3].astype(np.float64)
    r = flat.center0_r[k, 3].astype(np.float64)
    o4 = np.tile(np.array([13.0, 2.0, 3.0]), (n, 1)) + rng.normal(scale=0.5, size=(n, 3))
    to_c = c - o4
    perp = np.cross(to_c, rng.normal(size=(n, 3)))
    perp /= np.linalg.norm(perp, axis=1)[:, None]
    target = c + perp * (r * (1.0 + rng.normal(scale=3e-7, size=n)))[:, None]
    d4 = (target - o4) * rng.uniform(0.2, 3.0, size=(n, 1)) / np.linalg.norm(to_c, axis=1)[:, None]
    o = np.concatenate([o1, o2, o3, o4.astype(np.float32)])
    d = np.concatenate([d1, d2, d3, d4.astype(np.float32)])
    tm = np.concatenate([t1, t1, t3, np.zeros(n, np.float32)])
    return o, d, tm


def shade_inputs(flat, cam, S):
    rng = np.random.default_rng(77)
    n = 2048
    o, d, tm = camera_rays(cam, 1200, 800, n, rng)
    t, ids = S.hit(o, d, tm)
    p = (o.astype(np.float64) + (t * (1 + 1e-3))[:, None] * d.astype(np.float64)).astype(np.float32)
    o = np.concatenate([o, p[: n // 2]]); d = np.concatenate([d, d[: n // 2]]); tm = np.concatenate([tm, tm[: n // 2]])
    t, ids = S.hit(o, d, tm)
    ball = rng.uniform(-1, 1, size=(len(o), 3))
    ball *= (rng.random(len(o)) ** (1 / 3) / np.linalg.norm(ball, axis=1))[:, None]
    return o, d, tm, t, ids, ball.astype(np.float32), rng.random(len(o)).astype(np.float32)


def reference_kats():
    """The reference's own known answers for the path, transcribed from its test files (the expected values are
    the reference's assertions, not the oracle's output):
    hitable_test.clj:8-19 grid / directions; :23-47 per centre: ray (centre + dir, -dir) hits, ray (centre + dir,
    dir) misses, three grazing rays and a ray from the centre hit, all with t-range (0.0, Float/MAX_VALUE);
    :97-103 center-at-time; util_test.clj:44-49 point-at-parameter."""
    grid = [[25.0 * x for x in p] for p in itertools.product((-1, 0, 1), repeat=3)]
    dirs = [[5.0 * x for x in p] for p in itertools.product((-1, 0, 1), repeat=3) if any(p)]
    return {
        "source": "gonewest818/raytrace-clj test/raytrace_clj/hitable_test.clj:8-59,61-103 and util_test.clj:44-49",
        "sphere_radius": 1.0, "t_min": 0.0, "t_max": "Float/MAX_VALUE",
        "grid_centres": grid, "directions": dirs,
        "per_centre_and_direction": {"hit": {"origin": "centre + direction", "direction": "-direction"},
                                     "miss": {"origin": "centre + direction", "direction": "direction"}},
        "per_centre_hits": [{"origin_offset": [1, 1, 0], "direction": [-1, 0, 0], "name": "grazing ray x"},
                            {"origin_offset": [1, 1, 0], "direction": [0, -1, 0], "name": "grazing ray y"},
                            {"origin_offset": [1, 0, 1], "direction": [0, 0, -1], "name": "grazing ray z"},
                            {"origin_offset": [0, 0, 0], "direction": [1, 1, 1], "name": "ray from inside"}],
        "point_at_parameter": {"origin": [1, 2, 3], "direction": [4, 5, 6],
                               "cases": [{"t": 0, "p": [1, 2, 3]}, {"t": 1, "p": [5, 7, 9]}, {"t": -1, "p": [-3, -3, -3]}]},
        "center_at_time": {"c0": [0, 0, 0], "t0": 0, "c1": [1, 2, 3], "t1": 1,
                           "cases": [{"t": 0, "c": [0, 0, 0]}, {"t": 1, "c": [1, 2, 3]}, {"t": 0.5, "c": [0.5, 1.0, 1.5]}]},
    }


def main():
    flat, cam_type, cam = scene()
    S = oracle.Scene(flat)
    o, d, tm = hit_rays(flat, cam, S)
    t, ids = S.hit(o, d, tm, 0.001, FMAX)
    np.savez_compressed(os.path.join(HERE, "hits_random_scene.npz"), origins=o, dirs=d, times=tm, t=t, ids=ids)
    so, sd, stm, st, sids, ball, u = shade_inputs(flat, cam, S)
    ref = S.shade_batch(so, sd, stm, sids, ball, u)
    np.savez_compressed(os.path.join(HERE, "shade_random_scene.npz"), origins=so, dirs=sd, times=stm, hit_t=st, hit_id=sids,
                        ball=ball, u01=u, out_origin=ref["origin"].astype(np.float32), out_dir=ref["dir"].astype(np.float32),
                        out_atten=ref["atten"].astype(np.float32), out_emitted=ref["emitted"].astype(np.float32),
                        out_flags=ref["flags"])
    nx, ny, ns = 64, 40, 512
    sc = rt.scene.make_random_scene(nx, ny, 11, True, random.Random(1))
    _, cam_small = rt.native.marshal_camera(sc["camera"])
    a = S.render_accumulate(cam_type, cam_small, nx, ny, 0, ns, 50, seed=11, n_threads=1)[0] / ns
    b = S.render_accumulate(cam_type, cam_small, nx, ny, 0, ns, 50, seed=12, n_threads=1)[0] / ns
    np.savez_compressed(os.path.join(HERE, "image_random_scene.npz"), nx=nx, ny=ny, spp=ns, depth=50,
                        render_a=a.astype(np.float32), render_b=b.astype(np.float32))
    with open(os.path.join(HERE, "kat_reference_tests.json"), "w") as f:
        json.dump(reference_kats(), f, indent=1)
    print("hits", len(t), "hit fraction", float((ids >= 0).mean()), "| shade", len(sids), "| image RMSE a-b",
          float(np.sqrt(((a - b) ** 2).mean())))


if __name__ == "__main__":
    main()
