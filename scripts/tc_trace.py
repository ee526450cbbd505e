"""Device-side timeline (RT_TRACE) of one C2 frame with the tensor-core cull; one lane or two."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
out = sys.argv[1]
os.environ["RT_TRACE"] = out
import numpy as np
import bench
import raytrace_clj_b200 as rt
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 1
tcmode = int(sys.argv[3]) if len(sys.argv) > 3 else 1
nx, ny, spp, depth, scene_name, seed = bench.WORKLOADS["c2"]
flat, cam_type, cam = bench.build_scene(scene_name, nx, ny, seed)
img = np.empty((ny, nx, 3), np.uint8)
with rt.native.Renderer([0]) as r:
    r.set_scene(flat); r.set_camera(cam_type, cam)
    r.set_option("cull_tc", tcmode); r.set_option("wave_lanes", lanes)
    for k in range(3):
        r.render(nx, ny, spp, depth, seed=10 + k, linear=False, rgb8=True, out_rgb8=img)
    if os.path.exists(out): os.remove(out)
    r.render(nx, ny, spp, depth, seed=20, linear=False, rgb8=True, out_rgb8=img)
    r.counters()
