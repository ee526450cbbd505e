"""Option sweep with the tensor-core cull on (C2, C4, C1): python scripts/tc_sweep.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import raytrace_clj_b200 as rt

def run(name, opts, reps=6):
    nx, ny, spp, depth, scene_name, seed = bench.WORKLOADS[name]
    flat, cam_type, cam = bench.build_scene(scene_name, nx, ny, seed)
    img = np.empty((ny, nx, 3), np.uint8)
    with rt.native.Renderer([0]) as r:
        for k, v in opts.items(): r.set_option(k, v)
        r.set_scene(flat); r.set_camera(cam_type, cam)
        ts = []
        for k in range(reps):
            r.reset_counters()
            t0 = time.perf_counter()
            r.render(nx, ny, spp, depth, seed=10 + k, linear=False, rgb8=True, out_rgb8=img)
            ts.append(time.perf_counter() - t0)
        c = r.counters()
    print(f"{name:6s} {str(opts):90s} best {min(ts[1:])*1e3:8.3f} ms  device {c['kernel_ns']*1e-6:8.3f} ms", flush=True)

if __name__ == "__main__":
  base = {"cull_tc": 1}
  for name in ("c2", "c1", "c4"):
      run(name, {"cull_tc": 0})
      run(name, base)
      for extra in ({"wave_lanes": 1}, {"wave_depth": 3}, {"light_block": 64}, {"light_block": 256}, {"tail_entries": 1 << 18}, {"tail_entries": 1 << 20},
                    {"tail_entries": 1 << 17}, {"wave_capacity": 1 << 22}, {"wave_capacity": 1 << 23}):
          o = dict(base); o.update(extra); run(name, o)
