"""The ingest side of the JVM golden-vector loop, exercised without a JVM.

clojure/dump_vectors.clj (run by whoever has a JVM) prints hit records, AABB hits, scatter results, get-ray and gamma values of
the REAL reference as JSON; tests/test_jvm_vectors.py compares the oracle with them and is skipped while that file is absent.
Here a stand-in file with the SAME schema, the SAME hitables, rays, ranges, cameras and colours as dump_vectors.clj is computed by
the second restatement (tests/test_second_restatement.py: numpy float64 + a recursive interpreter of the records, on the double
inputs the JVM would see) and pushed through that very ingest test: its code path runs, its tolerances hold, and the oracle is
checked on exactly the cases a JVM owner will feed it."""
import json
import math

import numpy as np

import test_jvm_vectors as ingest
from test_second_restatement import FMAX, interp_hit, interp_scatter, normalise

RAYS = [([13, 2, 3], [-13, -1, -3]), ([13, 2, 3], [-12.5, -1.7, -2.4]), ([0, 5, 0], [0.1, -1, 0.05]), ([0.5, 0.5, -10], [0, 0, 2]),
        ([278, 278, -800], [0.1, -0.2, 1]), ([278, 278, -800], [-0.3, 0.4, 1]), ([4, 1.3, 0.2], [-1, -0.2, 0.1]),
        ([0.3, 0.2, 0.1], [1, 1, 1]), ([-2, 1, 0], [1, 0, 0]), ([3, 4, 5], [-0.3, -0.4, -0.5])]          # dump_vectors.clj `rays`
HITABLES = {                                                                                              # dump_vectors.clj `hitables`
    "sphere": {"center": [0, 1, 0], "radius": 1}, "ground": {"center": [0, -1000, 0], "radius": 1000},
    "uv-sphere": {"center": [0, 0, 0], "radius": 1000},
    "moving-sphere": {"center0": [4, 0.2, 0], "center1": [4, 0.6, 0], "t0": 0, "t1": 1, "radius": 0.2},
    "rect-xy": {"q": [-1, -1, 2, 3, 0.5]}, "rect-xz": {"q": [0, 0, 555, 555, 555]}, "rect-yz": {"q": [0, 0, 555, 555, 0]},
    "flip-rect-xz": {"q": [0, 0, 555, 555, 555]}, "triangle": {"q": [0, 0, 0, 0, 1, 0, 1, 0, 0]},
    "block": {"p0": [0, 0, 0], "p1": [165, 330, 165], "theta": 15.0, "offset": [265, 0, 295]}}


def _camera_records():
    lookfrom, lookat, vup = np.array([13.0, 2, 3]), np.zeros(3), np.array([0.0, 1, 0])
    out = []
    for name, aspect in (("thin-lens", float(np.float32(1200)) / float(np.float32(800))), ("pinhole", 1.5)):
        hh = math.tan(20 * (math.pi / 180.0) / 2.0)
        hw = aspect * hh
        w = normalise(lookfrom - lookat)
        u = normalise(np.cross(vup, w))
        v = np.cross(w, u)
        if name == "thin-lens":                                    # camera.clj:50-66, focus-dist 10, aperture 0.1
            f = 10.0
            rec = {"origin": lookfrom, "lleft": lookfrom - (f * hw * u + f * hh * v + f * w), "horiz": 2.0 * f * hw * u,
                   "vert": 2.0 * f * hh * v, "u": u, "v": v, "w": w, "aperture": 0.1, "t0": 0.0, "t1": 1.0}
        else:                                                      # camera.clj:18-33
            rec = {"origin": lookfrom, "lleft": lookfrom - (hw * u + hh * v + w), "horiz": 2.0 * hw * u, "vert": 2.0 * hh * v}
        for s, t in ((0.5, 0.5), (0.0, 1.0), (0.123, 0.877)):
            if name == "thin-lens":                                # camera.clj:35-48 with rand-in-unit-disk = (0.3 -0.4 0), rand = 0.25
                rd = (0.1 / 2.0) * np.array([0.3, -0.4])
                off = u * rd[0] + v * rd[1]
                o, d, tm = lookfrom + off, rec["lleft"] + s * rec["horiz"] + t * rec["vert"] - lookfrom - off, 0.0 + (1.0 - 0.0) * 0.25
            else:
                o, d, tm = lookfrom, rec["lleft"] + s * rec["horiz"] + t * rec["vert"] - lookfrom, 0
            out.append({"camera": name, "record": {k: (list(map(float, x)) if isinstance(x, np.ndarray) else x) for k, x in rec.items()},
                        "s": s, "t": t, "disk": [0.3, -0.4], "rand": 0.25, "o": list(map(float, o)), "d": list(map(float, d)), "time": tm})
    return out


def test_ingest_of_a_stand_in_vector_file(tmp_path, monkeypatch):
    hits = []
    for kind, params in HITABLES.items():
        obj = ingest._hitable(kind, params)
        for (o, d) in RAYS:
            for time in (0.0, 0.37):
                for tmin, tmax in ((0.001, FMAX), (0.0, 7.5)):
                    interp_hit.time = time
                    h = interp_hit(obj, np.array(o, np.float64), np.array(d, np.float64), tmin, tmax)
                    hits.append({"kind": kind, "params": params, "ray": {"o": o, "d": d, "time": time}, "tmin": tmin, "tmax": tmax,
                                 "hit": None if h is None else {"t": float(h[0]), "p": list(map(float, h[1])), "normal": list(map(float, h[2])),
                                                                "uv": [float(h[3][0]), float(h[3][1])]}})
    interp_hit.time = 0.0
    assert sum(h["hit"] is not None for h in hits) > 100
    aabb = []
    for (o, d) in RAYS:
        for time in (0.0, 0.37):
            for lo, hi in (([-1, -1, -1], [1, 1, 1]), ([0, 0, 0], [555, 555, 555])):
                with np.errstate(divide="ignore", invalid="ignore"):
                    m, n = (np.array(lo, float) - o) / np.array(d, float), (np.array(hi, float) - o) / np.array(d, float)
                t0, t1 = np.minimum(m, n), np.maximum(m, n)
                if np.isnan(t0).any() or np.isnan(t1).any():
                    continue          # 0 / 0 in a slab: clojure.core/min and Math.max propagate NaN — not a comparable case
                aabb.append({"vmin": lo, "vmax": hi, "ray": {"o": o, "d": d, "time": time},
                             "hit": bool(min(t1.min(), FMAX) > max(t0.max(), 0.001))})
    gamma = [{"mean": c, "rgb8": [int(min(255.99, 255.99 * math.sqrt(x))) for x in c]}
             for c in ([0.0, 0.25, 1.0], [0.5, 2.0, 7.0], [1e-6, 0.999, 0.1234])]
    # dump_vectors.clj `scat`: every material on a uv-sphere (0 1 0) r = 1, every ray that hits it, rand-in-unit-sphere = ball, rand = rnd
    from raytrace_clj_b200 import hitable as H
    from raytrace_clj_b200 import shader as shad
    from raytrace_clj_b200 import texture as tex
    from raytrace_clj_b200.util import vec3

    mats = {"lambertian": shad.lambertian(albedo=tex.checkerboard(tex0=tex.constant(color=vec3(0.2, 0.3, 0.1)),
                                                                  tex1=tex.constant(color=vec3(0.9, 0.9, 0.9)), scale=10)),
            "metal": shad.metal(albedo=tex.constant(color=vec3(0.7, 0.6, 0.5)), fuzz=0.3),
            "dielectric": shad.dielectric(ri=1.5),
            "light": shad.diffuse_light(tex=tex.uv_gradient(co=vec3(1, 1, 1), cu=vec3(1, 1, 1), cv=vec3(0.5, 0.7, 1.0), cuv=vec3(0.5, 0.7, 1.0))),
            "isotropic": shad.isotropic(albedo=tex.constant(color=vec3(0.2, 0.4, 0.9)))}
    ball = [0.1, -0.2, 0.3]
    scat = []
    for mk, m in mats.items():
        sph = H.uv_sphere(center=vec3(0, 1, 0), radius=1, material=m)
        for (o, d) in RAYS:
            for time in (0.0, 0.37):
                for rnd in (0.01, 0.5, 0.99):
                    o64, d64 = np.array(o, np.float64), np.array(d, np.float64)
                    interp_hit.time = time
                    h = interp_hit(sph, o64, d64, 0.001, FMAX)
                    if h is None:
                        continue
                    sc, em = interp_scatter(m, o64, d64, time, h, ball, rnd)
                    scat.append({"material": mk, "ray": {"o": o, "d": d, "time": time}, "ball": ball, "rand": rnd,
                                 "hit": {"t": float(h[0]), "p": list(map(float, h[1])), "normal": list(map(float, h[2])), "uv": list(map(float, h[3]))},
                                 "scattered": None if sc is None else {"o": list(map(float, sc[0])), "d": list(map(float, sc[1])), "time": float(sc[2]),
                                                                       "attenuation": list(map(float, sc[3]))},
                                 "emitted": list(map(float, em))})
    interp_hit.time = 0.0
    assert len(scat) > 60 and any(c["scattered"] is None for c in scat)
    J = {"reference": "gonewest818/raytrace-clj", "hits": hits, "aabb": aabb, "scatter": scat, "get_ray": _camera_records(), "gamma": gamma}
    path = tmp_path / "jvm_vectors.json"
    path.write_text(json.dumps(J).replace(str(FMAX), repr(FMAX)))
    monkeypatch.setattr(ingest, "PATH", str(path))
    ingest.test_oracle_reproduces_the_reference_vectors()
