"""Bring-up of the tensor-core cull (option cull_tc) on a GPU box: cull contract, bit-identical frames, timing.
Run under `timeout`:  timeout 300 python scripts/tc_try.py [workload ...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bench                      # noqa: E402
import raytrace_clj_b200 as rt    # noqa: E402
from helpers import camera_rays   # noqa: E402

FMAX = float(np.finfo(np.float32).max)


def check(r, flat, cam):
    rng = np.random.default_rng(21)
    fam = {}
    o, d, tm = camera_rays(cam, 1200, 800, 100_000, rng)
    fam["camera"] = (o, d, tm)
    o2 = rng.uniform(-15, 15, size=(100_000, 3)).astype(np.float32)
    o2[:, 1] = rng.uniform(-1, 3, size=len(o2))
    fam["volume"] = (o2, rng.normal(size=o2.shape).astype(np.float32), rng.random(len(o2)).astype(np.float32))
    for scale in (1.0, 10.0, 100.0, 1000.0):
        for jitter in (3e-7, 1e-4):
            n = 50_000
            k = rng.integers(0, flat.n_spheres, n)
            c = flat.center0_r[k, :3].astype(np.float64)
            rad = flat.center0_r[k, 3].astype(np.float64)
            oo = rng.normal(size=(n, 3)) * scale
            to_c = c - oo
            dist = np.linalg.norm(to_c, axis=1)
            perp = np.cross(to_c, rng.normal(size=(n, 3)))
            perp /= np.linalg.norm(perp, axis=1)[:, None]
            target = c + perp * (rad * (1.0 + rng.normal(scale=jitter, size=n)))[:, None]
            dd = (target - oo) * rng.uniform(0.2, 3.0, size=(n, 1)) / dist[:, None]
            fam[f"grazing |o|~{scale:g} +-{jitter:g}"] = (oo.astype(np.float32), dd.astype(np.float32), rng.random(n).astype(np.float32))
    for tcmode in (0, 1):
        r.set_option("cull_tc", tcmode)
        for name, (o, d, tm) in fam.items():
            lost, surv, cand = r.cull_check(o, d, tm, 0.001, FMAX)
            print(f"cull_tc={tcmode} {name:32s} lost {lost}  survivors {surv}  exact {cand}  ratio {surv / max(1, cand):.3f}", flush=True)


def main():
    names = sys.argv[1:] or ["c2"]
    for name in names:
        nx, ny, spp, depth, scene_name, seed = bench.WORKLOADS[name]
        flat, cam_type, cam = bench.build_scene(scene_name, nx, ny, seed)
        with rt.native.Renderer([0]) as r:
            r.set_scene(flat)
            r.set_camera(cam_type, cam)
            if name == names[0]:
                check(r, flat, cam)
            imgs = {}
            for tcmode in (0, 1):
                r.set_option("cull_tc", tcmode)
                lin, _ = r.render(min(nx, 400), min(ny, 300), 4, depth, seed=5, linear=True, rgb8=False)
                imgs[tcmode] = lin
            same = np.array_equal(imgs[0], imgs[1])
            print(f"{name}: small frame identical with / without the tensor-core cull: {same}  (max |diff| {np.abs(imgs[0] - imgs[1]).max():.3g})", flush=True)
            out = np.empty((ny, nx, 3), np.uint8)
            for tcmode in (0, 1):
                r.set_option("cull_tc", tcmode)
                for per in ((0,) if tcmode == 0 else (0, 8, 16, 32, 64)):
                    r.set_option("tc_tiles_per_cta", per)
                    ts = []
                    for k in range(6):
                        r.reset_counters()
                        t0 = time.perf_counter()
                        r.render(nx, ny, spp, depth, seed=10 + k, linear=False, rgb8=True, out_rgb8=out)
                        ts.append(time.perf_counter() - t0)
                    c = r.counters()
                    print(f"{name}: cull_tc={tcmode} tiles/cta={per:3d}  best {min(ts[1:]) * 1e3:8.3f} ms  median {np.median(ts[1:]) * 1e3:8.3f} ms  "
                          f"device {c['kernel_ns'] * 1e-6:8.3f} ms  candidates/ray {c['candidates'] / max(1, c['rays']):.3f}", flush=True)


if __name__ == "__main__":
    main()
