"""Where do the production kernels and the oracle replay part ways?  (GPU box: python scripts/diag_paths.py [workload] [n])"""
import random
import sys

import numpy as np

sys.path.insert(0, ".")
import oracle  # noqa: E402
import raytrace_clj_b200 as rt  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 1
nx, ny = 1200, 800
sc = rt.scene.make_random_scene(nx, ny, 11, True, random.Random(1)) if wl == "c2" else rt.scene.make_material_stress_scene(nx, ny, 11, random.Random(4))
flat = rt.native.marshal_world(sc["world"])
cam_type, cam = rt.native.marshal_camera(sc["camera"])
S = oracle.Scene(flat)
LB = 8
g = np.random.default_rng(1)
pix = g.integers(0, nx * ny, n).astype(np.int32)
smp = g.integers(0, 10, n).astype(np.int32)
with rt.native.Renderer([0]) as r:
    r.set_scene(flat)
    r.set_camera(cam_type, cam)
    rad, nr, term, log = r.trace_paths(nx, ny, pix, smp, 50, seed=101, variant=variant, log_bounces=LB)
orad, onr, oterm, olog = S.trace_paths(cam_type, cam, nx, ny, pix, smp, 50, seed=101, log_bounces=LB)
same = (nr == onr) & (term == oterm)
d = rad.astype(np.float64) - orad
print(f"{wl} variant {variant}: n {n} same {same.mean():.6f}; mean gpu {rad.mean(axis=0)} oracle {orad.mean(axis=0)}")
print("mean diff all      ", d.mean(axis=0), " +- ", d.std(axis=0) / np.sqrt(n))
print("mean diff same     ", d[same].mean(axis=0) * same.mean(), "(weighted)")
print("mean diff different", d[~same].mean(axis=0) * (~same).mean(), "(weighted)")
close = np.abs(d).max(axis=1) <= 1e-3 * np.maximum(np.abs(orad).max(axis=1), 1e-2)
print("same but radiance differs:", (same & ~close).sum(), " contribution", d[same & ~close].sum(axis=0) / n)
print("same and close contribution", d[same & close].sum(axis=0) / n)
# first bounce where the logged hit differs
idd = log["hit_id"] != olog["hit_id"]
first = np.where(idd.any(axis=1), idd.argmax(axis=1), -1)
print("paths whose logged hit ids differ:", (first >= 0).sum(), "first differing bounce histogram", np.bincount(first[first >= 0], minlength=LB))
mt = flat.mat_type[flat.material_id]
for b in range(1, 4):
    sel = first == b
    if sel.sum():
        prev = olog["hit_id"][sel, b - 1]
        print(f"  diverge at bounce {b}: previous hit material histogram (lambert, metal, glass, light)", np.bincount(mt[prev[prev >= 0]], minlength=4),
              " prev hit is ground:", (prev == 1).sum())
# the same-and-close population: per-material bias of the FIRST hit's contribution
for b in range(0, 3):
    ok = same & (olog["hit_id"][:, b] >= 0) & (log["hit_id"][:, b] == olog["hit_id"][:, b])
    dt = np.abs(log["t"][ok, b] - olog["t"][ok, b])
    do = np.abs(log["o"][ok, b] - olog["o"][ok, b]).max(axis=1)
    dd = np.abs(log["d"][ok, b] - olog["d"][ok, b]).max(axis=1)
    print(f"bounce {b}: rays agree: |do| max {do.max():.3e} mean {do.mean():.3e}; |dd| max {dd.max():.3e} mean {dd.mean():.3e}; |dt| max {dt.max():.3e}")
# radiance differences by the material where the path ENDED / by path length
for L in range(1, 6):
    sel = same & (nr == L)
    print(f"len {L}: n {sel.sum()} mean diff {d[sel].mean(axis=0)} rel {np.abs(d[sel]).max():.3e}")
sky = same & (term == 1)
big = same & (np.abs(d).max(axis=1) > 1e-4)
print("same paths with |diff| > 1e-4:", big.sum(), "of which len1", (big & (nr == 1)).sum(), "len2", (big & (nr == 2)).sum())
j = np.nonzero(big)[0][:10]
for q in j:
    print(" path", q, "nr", nr[q], "rad", rad[q], "orad", orad[q], "hits", log["hit_id"][q, :nr[q]], olog["hit_id"][q, :nr[q]])
