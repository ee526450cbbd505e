"""Front end — mirror of reference src/raytrace_clj/core.clj.

`-main` (core.clj:73-115) keeps its shape: parse `[filename] [nx] [ny] [nsamples]`, build the
scene, render, save.  The render block core.clj:99-108 (claypoole `upmap` over 32-pixel chunks
calling `pixel`) is ONE native call here: `Renderer.render` → `rt_render` → sm_100a kernels.
There is no CPU fallback: without the CUDA library or a GPU this raises.
"""
from __future__ import annotations

import math
import os
import random
import sys
import time

from . import native, ppm, scene


def tiled_coords(width, height, chunk_size):
    """core.clj:59-71 — row-major (j outer, i inner) pixel list cut into chunks with completion
    fractions.  Kept as the specification of the reference's work decomposition; the GPU grid
    replaces it."""
    coords = [(i, j) for j in range(height) for i in range(width)]
    chunks = math.ceil(float(width * height) / chunk_size)
    out = []
    for k in range(0, len(coords), chunk_size):
        out.append({"chunk": coords[k:k + chunk_size], "pct": (len(out) + 1) / chunks})
    return out


SCENES = {
    "random": lambda nx, ny, rng: scene.make_random_scene(nx, ny, 11, True, rng),        # core.clj:89
    "random-static": lambda nx, ny, rng: scene.make_random_scene(nx, ny, 11, False, rng),
    "two-spheres": lambda nx, ny, rng: scene.make_two_spheres(nx, ny, rng),              # core.clj:82
    "stress": lambda nx, ny, rng: scene.make_material_stress_scene(nx, ny, 11, rng),
}


def render(camera, world, nx, ny, nr, *, depth=50, seed=1, variant=native.RT_VARIANT_WAVEFRONT,
           device_ids=None, renderer=None):
    """core.clj:99-108 replaced: returns (linear float32 [ny,nx,3] bottom-up, rgb8 [ny,nx,3] top-down,
    counters)."""
    own = renderer is None
    r = renderer or native.Renderer(device_ids)
    try:
        r.set_scene(native.marshal_world(world))
        r.set_camera(*native.marshal_camera(camera))
        lin, img = r.render(nx, ny, nr, depth, seed, variant)
        return lin, img, r.counters()
    finally:
        if own:
            r.close()


def main(argv=None):
    """core.clj:73-115: `[filename] [nx] [ny] [nsamples] [win]`.  Additive, optional knobs come from
    the environment so the 4 documented positional arguments stay byte-for-byte:
    RT_SCENE (random | random-static | two-spheres | stress), RT_SEED, RT_SCENE_SEED,
    RT_VARIANT (1 wavefront [default], 0 megakernel), RT_DEVICES (e.g. "0,1,2,3")."""
    argv = list(sys.argv[1:] if argv is None else argv)
    tstart = time.time()
    filename = argv[0] if len(argv) > 0 else "render.png"
    nx = int(argv[1]) if len(argv) > 1 else 200
    ny = int(argv[2]) if len(argv) > 2 else 100
    nr = int(argv[3]) if len(argv) > 3 else 100
    if nx <= 0 or ny <= 0 or nr <= 0:
        raise ValueError("nx, ny, nsamples must be positive")  # Integer/parseUnsignedInt
    scene_name = os.environ.get("RT_SCENE", "random")
    rng = random.Random(int(os.environ.get("RT_SCENE_SEED", "1")))
    sc = SCENES[scene_name](nx, ny, rng)
    devs = [int(x) for x in os.environ.get("RT_DEVICES", "0").split(",")]
    lin, img, ctr = render(sc["camera"], sc["world"], nx, ny, nr, seed=int(os.environ.get("RT_SEED", "1")),
                           variant=int(os.environ.get("RT_VARIANT", "1")), device_ids=devs)
    ppm.save(filename, img)
    dt = time.time() - tstart
    print("%.2fs, 100%%, %d rays, %d ray-sphere tests" % (dt, ctr["rays"], ctr["sphere_tests"]))
    print("wrote", filename)
    return 0


if __name__ == "__main__":
    sys.exit(main())
