"""Tensor-core cull on lists over 1024 leaves (several launches per iteration): contract, identical paths, timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import raytrace_clj_b200 as rt
FMAX = float(np.finfo(np.float32).max)
rng = np.random.default_rng(5)
for scene in ("sweep:3000", "sweep:10000"):
    nx, ny = 480, 270
    flat, cam_type, cam = bench.build_scene(scene, nx, ny, 5)
    with rt.native.Renderer([0]) as r:
        r.set_scene(flat); r.set_camera(cam_type, cam)
        n = 20000
        side = np.sqrt(flat.n_spheres) / 2
        o = rng.uniform(-side, side, size=(n, 3)).astype(np.float32); o[:, 1] = rng.uniform(0, 3, n)
        d = rng.normal(size=(n, 3)).astype(np.float32)
        for mode in (0, 1):
            r.set_option("cull_tc", mode)
            print(scene, "cull_tc", mode, "cull_check", r.cull_check(o, d, None, 0.001, FMAX), flush=True)
        pix = rng.integers(0, nx * ny, 100_000).astype(np.int32); smp = rng.integers(0, 16, 100_000).astype(np.int32)
        out = {}
        for mode in (0, 1):
            r.set_option("cull_tc", mode)
            out[mode] = r.trace_paths(nx, ny, pix, smp, 50, seed=9)
        print(scene, "paths identical:", all(np.array_equal(a, b) for a, b in zip(out[0][:3], out[1][:3])), flush=True)
        img = np.empty((ny, nx, 3), np.uint8)
        for mode in (0, 1):
            r.set_option("cull_tc", mode)
            ts = []
            for k in range(3):
                t0 = time.perf_counter(); r.render(nx, ny, 32, 50, seed=3 + k, linear=False, rgb8=True, out_rgb8=img); ts.append(time.perf_counter() - t0)
            print(scene, "cull_tc", mode, f"{nx}x{ny}x32: best {min(ts)*1e3:.1f} ms", flush=True)
