"""Stage times (one lane, serialised with events) with and without the tensor-core cull."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import raytrace_clj_b200 as rt
import torch

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
nx, ny, spp, depth, scene_name, seed = bench.WORKLOADS[name]
flat, cam_type, cam = bench.build_scene(scene_name, nx, ny, seed)
dev = torch.device("cuda", 0)
with rt.native.Renderer([0]) as r:
    r.set_scene(flat); r.set_camera(cam_type, cam)
    ps = torch.zeros(ny, nx, 3, device=dev, dtype=torch.float32)
    for tcmode in (0, 1):
        r.set_option("cull_tc", tcmode)
        r.set_profile(False)
        r.render_accumulate_device(nx, ny, 0, spp, ps.data_ptr(), max_depth=depth, seed=1, sync=True)
        r.reset_counters(); r.set_profile(True); ps.zero_()
        r.render_accumulate_device(nx, ny, 0, spp, ps.data_ptr(), max_depth=depth, seed=1, sync=True)
        c = r.counters()
        print(f"{name} cull_tc={tcmode}: cull {c['cull_ns']*1e-6:.3f} ms  refine {c['refine_ns']*1e-6:.3f}  tiebreak {c['tiebreak_ns']*1e-6:.3f}  shade {c['shade_ns']*1e-6:.3f}  "
              f"rays {c['rays']}  tests {c['sphere_tests']:.4g}  cull rate {c['sphere_tests']/max(1,c['cull_ns'])*1e9/1e12:.2f} T tests/s", flush=True)
