"""C2 device time + the cull_tc 0 / 1 identical-paths check on the library currently in place."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import raytrace_clj_b200 as rt

tag = sys.argv[1] if len(sys.argv) > 1 else ""
nx, ny, spp, depth, scene_name, seed = bench.WORKLOADS["c2"]
flat, cam_type, cam = bench.build_scene(scene_name, nx, ny, seed)
img = np.empty((ny, nx, 3), np.uint8)
with rt.native.Renderer([0]) as r:
    r.set_scene(flat); r.set_camera(cam_type, cam)
    dev = []
    for k in range(6):
        r.reset_counters()
        r.render(nx, ny, spp, depth, seed=10 + k, linear=False, rgb8=True, out_rgb8=img)
        c = r.counters()
        dev.append(c["kernel_ns"] * 1e-6)
    print(tag, "C2 device ms", " ".join(f"{x:.3f}" for x in dev), "rays", c["rays"], "cand", c["candidates"], flush=True)
    fl2, ct2, cm2 = bench.build_scene("random", 320, 200, 1)
    r.set_scene(fl2); r.set_camera(ct2, cm2)
    rng = np.random.default_rng(3)
    n = 200_000
    pix = rng.integers(0, 320 * 200, n).astype(np.int32)
    smp = rng.integers(0, 64, n).astype(np.int32)
    out = {}
    for mode in (0, 1, 0, 1):
        r.set_option("cull_tc", mode)
        r.reset_counters()
        o = r.trace_paths(320, 200, pix, smp, 50, seed=9)
        c = r.counters()
        print(tag, "cull_tc", mode, "rays", c["rays"], "cand", c["candidates"], "direct", c["direct_tests"], "terms", c["term_light"], c["term_absorb"], c["term_depth"], c["term_miss"], flush=True)
        if mode in out:
            print(tag, "  same as the earlier run of this mode:", all(np.array_equal(a, b) for a, b in zip(out[mode][:3], o[:3])))
        out[mode] = o
    d = [int((a != b).any(axis=-1).sum()) if a.ndim > 1 else int((a != b).sum()) for a, b in zip(out[0][:3], out[1][:3])]
    print(tag, "paths differing between cull modes (radiance, nrays, term):", d, flush=True)
