"""A/B of library builds on the C2 frame: python scripts/tc_ab.py lib1.so lib2.so ...  (cull_tc = 1, persistent CTAs)"""
import os, sys, time, subprocess
here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 2 or (len(sys.argv) == 2 and not sys.argv[1].startswith("--one=")):
    for lib in sys.argv[1:]:
        subprocess.run([sys.executable, __file__, "--one=" + lib])
    sys.exit(0)
lib = sys.argv[1][6:]
sys.path.insert(0, here)
import numpy as np
import bench
import raytrace_clj_b200 as rt
rt.native.LIB_PATH = os.path.join(os.path.dirname(rt.native.LIB_PATH), lib)
for name in ("c2", "c4"):
    nx, ny, spp, depth, scene_name, seed = bench.WORKLOADS[name]
    flat, cam_type, cam = bench.build_scene(scene_name, nx, ny, seed)
    img = np.empty((ny, nx, 3), np.uint8)
    with rt.native.Renderer([0]) as r:
        r.set_scene(flat); r.set_camera(cam_type, cam)
        for tcmode in (0, 1):
            r.set_option("cull_tc", tcmode)
            ts = []
            for k in range(6):
                r.reset_counters()
                t0 = time.perf_counter()
                r.render(nx, ny, spp, depth, seed=10 + k, linear=False, rgb8=True, out_rgb8=img)
                ts.append(time.perf_counter() - t0)
            c = r.counters()
            print(f"{lib:32s} {name} cull_tc={tcmode}  best {min(ts[1:])*1e3:8.3f} ms  device {c['kernel_ns']*1e-6:8.3f} ms  cand/ray {c['candidates']/max(1,c['rays']):.3f}", flush=True)
