import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import raytrace_clj_b200 as rt
nx, ny, spp, depth, scene_name, seed = bench.WORKLOADS["c2"]
flat, cam_type, cam = bench.build_scene(scene_name, nx, ny, seed)
img = np.empty((ny, nx, 3), np.uint8)
with rt.native.Renderer([0]) as r:
    for k in range(3):
        r.set_scene(flat); r.set_camera(cam_type, cam); r.render(nx, ny, spp, depth, seed=k, linear=False, rgb8=True, out_rgb8=img)
    T = {"set_scene": [], "set_camera": [], "render": [], "device": []}
    for k in range(20):
        t0 = time.perf_counter(); r.set_scene(flat); t1 = time.perf_counter(); r.set_camera(cam_type, cam); t2 = time.perf_counter()
        r.reset_counters()
        t2 = time.perf_counter()
        r.render(nx, ny, spp, depth, seed=10 + k, linear=False, rgb8=True, out_rgb8=img); t3 = time.perf_counter()
        c = r.counters()
        T["set_scene"].append(t1 - t0); T["set_camera"].append(0); T["render"].append(t3 - t2); T["device"].append(c["kernel_ns"] * 1e-9)
    for k, v in T.items(): print(f"{k:12s} median {np.median(v)*1e6:9.1f} us  min {np.min(v)*1e6:9.1f} us")
    print("render - device:", (np.median(T["render"]) - np.median(T["device"])) * 1e6, "us")
