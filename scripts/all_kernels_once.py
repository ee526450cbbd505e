"""A small pass over every kernel of the library (compute-sanitizer is closed on this GPU pool, so this runs plain; it is
the shape such a run would take): both render variants on the random scene (two lanes, common-origin and general cull, per-CTA tail), a
defocus camera, a tiled scene (> 4096 spheres), the trace / shade / cull-check diagnostics."""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import raytrace_clj_b200 as rt  # noqa: E402
from helpers import camera_rays  # noqa: E402

nx, ny, ns = int(os.environ.get("NX", 320)), int(os.environ.get("NY", 240)), int(os.environ.get("NS", 2))
sc = rt.scene.make_random_scene(nx, ny, 11, True, random.Random(1))
flat = rt.native.marshal_world(sc["world"])
cam_type, cam = rt.native.marshal_camera(sc["camera"])
with rt.native.Renderer([0]) as r:
    r.set_scene(flat)
    r.set_camera(cam_type, cam)
    for variant in (1, 0):
        lin, img = r.render(nx, ny, ns, 50, seed=3, variant=variant)      # 153 600 samples: two lanes + tail
        assert np.isfinite(lin).all() and img.shape == (ny, nx, 3)
    cam2 = np.array(cam, np.float32); cam2[21] = 0.3
    r.set_camera(cam_type, cam2)
    r.render(nx, ny, 1, 50, seed=4, variant=1)
    r.set_camera(cam_type, cam)
    o, d, tm = camera_rays(cam, nx, ny, 3000, np.random.default_rng(1))
    t, ids = r.trace_primary(o, d, tm)
    lost, surv, cand = r.cull_check(o, d, tm)
    assert lost == 0
    ctr = r.counters()
    assert ctr["sphere_tests"] == ctr["rays"] * flat.n_spheres
    big = rt.scene.make_scale_sweep_scene(96, 64, 4500, random.Random(5))
    bflat = rt.native.marshal_world(big["world"])
    bt, bc = rt.native.marshal_camera(big["camera"])
    r.set_scene(bflat)
    r.set_camera(bt, bc)
    r.render(96, 64, 2, 50, seed=5, variant=1)                          # 12 288 samples: single lane, tiled cull
    o, d, tm = camera_rays(bc, 96, 64, 600, np.random.default_rng(2))
    r.trace_primary(o, d, tm)
print("sanitize driver ok")
