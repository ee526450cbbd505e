import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import raytrace_clj_b200 as rt
if os.environ.get("RT_LIB"): rt.native.LIB_PATH = os.path.join(os.path.dirname(rt.native.LIB_PATH), os.environ["RT_LIB"])
name = sys.argv[1]
FMAX = float(np.finfo(np.float32).max)
flat, cam_type, cam = bench.build_scene(name, 600, 600, 1)
rng = np.random.default_rng(1)
with rt.native.Renderer([0]) as r:
    r.set_scene(flat); r.set_camera(cam_type, cam)
    r.set_option("cull_tc", int(os.environ.get("TC", "1")))
    for n in [int(x) for x in sys.argv[2:]]:
        o = rng.uniform(-5, 5, size=(n, 3)).astype(np.float32)
        d = rng.normal(size=(n, 3)).astype(np.float32)
        try:
            print(name, n, r.cull_check(o, d, None, 0.001, FMAX), flush=True)
        except Exception as e:
            print(name, n, "FAILED", e, flush=True); break
