"""Cycle accounting of wf_cull_tc per role (needs the -DRT_TC_TIMING build: libraytrace_b200_timing.so). One lane, C2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import raytrace_clj_b200 as rt
rt.native.LIB_PATH = os.path.join(os.path.dirname(rt.native.LIB_PATH), "libraytrace_b200_timing.so")
nx, ny, spp, depth, scene_name, seed = bench.WORKLOADS["c2"]
flat, cam_type, cam = bench.build_scene(scene_name, nx, ny, seed)
img = np.empty((ny, nx, 3), np.uint8)
with rt.native.Renderer([0]) as r:
    r.set_scene(flat); r.set_camera(cam_type, cam)
    r.set_option("cull_tc", 1); r.set_option("wave_lanes", 1)
    r.render(nx, ny, spp, depth, seed=20, linear=False, rgb8=True, out_rgb8=img)
