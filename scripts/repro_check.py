"""Same seed, several runs: do the counters repeat exactly?  (diagnostic)"""
import os, random, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytrace_clj_b200 as rt
nx, ny, ns = 480, 320, 4
moving = os.environ.get("MOVING", "1") == "1"
sc = rt.scene.make_random_scene(nx, ny, 11, moving, random.Random(1))
flat = rt.native.marshal_world(sc["world"])
cam_type, cam = rt.native.marshal_camera(sc["camera"])
with rt.native.Renderer([0]) as r:
    r.set_scene(flat); r.set_camera(cam_type, cam)
    for variant in (0, 1):
        base = None
        for k in range(6):
            r.reset_counters()
            lin, img = r.render(nx, ny, ns, int(os.environ.get("DEPTH", 50)), seed=99, variant=variant)
            c = r.counters()
            row = tuple(c[x] for x in ("rays", "samples", "term_light", "term_absorb", "term_depth", "term_miss", "candidates"))
            if base is None: base, lin0 = row, lin
            d = np.abs(lin - lin0)
            print("variant", variant, "run", k, row, "same" if row == base else "DIFF", "max|dlin| %.3g at %s" % (d.max(), np.unravel_index(d.argmax(), d.shape)), flush=True)
