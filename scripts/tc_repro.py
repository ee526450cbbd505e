import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import raytrace_clj_b200 as rt
if os.environ.get("RT_LIB"): rt.native.LIB_PATH = os.path.join(os.path.dirname(rt.native.LIB_PATH), os.environ["RT_LIB"])
name = sys.argv[1] if len(sys.argv) > 1 else "cornell"
cap = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 18
nx, ny, spp = (int(x) for x in (sys.argv[3:6] if len(sys.argv) > 5 else (200, 200, 16)))
flat, cam_type, cam = bench.build_scene(name, nx, ny, 1)
img = np.empty((ny, nx, 3), np.uint8)
with rt.native.Renderer([0]) as r:
    r.set_option("wave_capacity", cap)
    if len(sys.argv) > 6: r.set_option("cull_tc", int(sys.argv[6]))
    r.set_scene(flat); r.set_camera(cam_type, cam)
    for k in range(2):
        r.render(nx, ny, spp, 50, seed=10 + k, linear=False, rgb8=True, out_rgb8=img)
    print(name, "ok", img.mean())
