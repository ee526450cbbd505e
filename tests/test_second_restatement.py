"""A SECOND, independent restatement of the reference's per-ray numerics — vectorised numpy float64, transcribed straight from the
Clojure sources — checked against the C++ oracle on thousands of random inputs.

VERDICT round 1, weak #2: the reference holds no vectors for t / p / normal / uv / scatter / get-ray / gamma, so the oracle's
numerics beyond hit-or-miss were pinned by inspection of ONE restatement.  This file restates the same functions a second time, in
another language and another formulation (whole arrays, the textbook b = 2 oc.d / 4ac form exactly as the Clojure writes it, no
shared helper with oracle.cpp), so a transcription slip in either copy shows up as a disagreement.  Cited lines are the reference's.
No GPU, no product code: oracle (test infrastructure) against numpy.
"""
import numpy as np
import pytest

import oracle
from raytrace_clj_b200.native import (RT_MAT_DIELECTRIC, RT_MAT_DIFFUSE_LIGHT, RT_MAT_LAMBERTIAN, RT_MAT_METAL, RT_SPHERE_MOVING,
                                      RT_SPHERE_UV, RT_TEX_CHECKERBOARD, RT_TEX_CONSTANT, RT_TEX_UV_GRADIENT)

FMAX = float(np.finfo(np.float32).max)


# ---- numpy transcription ---------------------------------------------------------------------------------------------------------
def dot(a, b):
    return (a * b).sum(axis=-1)


def normalise(v):          # core.matrix `normalise`: multiply by 1 / magnitude
    return v * (1.0 / np.sqrt(dot(v, v)))[..., None]


def np_sphere_hit(center, radius, o, d, t_min, t_max):
    """hitable.clj:141-171 / 180-207 / 224-251 (the three records share the body): returns (hit, t, p, normal).
    b = 2 oc.d, c = oc.oc - r^2, discriminant b^2 - 4ac >= 0, near root then far root, strict (t-min, t-max)."""
    oc = o - center
    a = dot(d, d)
    b = 2.0 * dot(oc, d)
    c = dot(oc, oc) - radius * radius
    disc = b * b - 4.0 * a * c
    ok = disc >= 0
    sq = np.sqrt(np.where(ok, disc, 0.0))
    t_near = (-b - sq) / (2.0 * a)
    t_far = (-b + sq) / (2.0 * a)
    near_ok = ok & (t_near > t_min) & (t_near < t_max)
    far_ok = ok & (t_far > t_min) & (t_far < t_max)
    t = np.where(near_ok, t_near, t_far)
    hit = near_ok | far_ok
    p = o + t[..., None] * d                     # util.clj:18-22 point-at-parameter
    with np.errstate(invalid="ignore", divide="ignore"):
        n = normalise(p - center)
    return hit, t, p, n


def np_sphere_uv(n):
    """hitable.clj:128-139 get-sphere-uv."""
    phi = np.arctan2(n[..., 2], n[..., 0])
    theta = np.arcsin(n[..., 1])
    return 1.0 - (phi + np.pi) / (2.0 * np.pi), (theta + np.pi / 2.0) / np.pi


def np_center_at_time(c0, t0, c1, t1, t):
    """hitable.clj:219-222: mat/lerp = a (1 - f) + b f."""
    f = ((t - t0) / (t1 - t0))[..., None]
    return c0 * (1.0 - f) + c1 * f


def np_reflect(v, n):          # shader.clj:6-9
    return v - 2.0 * dot(v, n)[..., None] * n


def np_refract(v, n, ni_over_nt):
    """shader.clj:11-20: (ok, refracted)."""
    uv = normalise(v)
    dt = dot(uv, n)
    disc = 1.0 - ni_over_nt * ni_over_nt * (1 - dt * dt)
    ok = disc > 0
    r = ni_over_nt[..., None] * (uv - n * dt[..., None]) - n * np.sqrt(np.where(ok, disc, 0.0))[..., None]
    return ok, r


def np_schlick(cosine, ri):    # shader.clj:69-74
    r0 = (1.0 - ri) / (1.0 + ri)
    r0 = r0 * r0
    return r0 + (1.0 - r0) * (1.0 - cosine) ** 5


def np_tex_sample(flat, tex, u, v, p):
    """texture.clj:14-50 for ONE point (recursive for the checkerboard)."""
    ty = int(flat.tex_type[tex])
    q = flat.tex_params[tex].astype(np.float64)
    if ty == RT_TEX_CONSTANT:
        return q[0:3]
    if ty == RT_TEX_UV_GRADIENT:
        co, cu, cv, cuv = q[0:3], q[3:6], q[6:9], q[9:12]
        a = cu * (1 - u) + co * u
        b = cuv * (1 - u) + cv * u
        return b * (1 - v) + a * v
    assert ty == RT_TEX_CHECKERBOARD
    sines = np.prod(np.sin(q[0] * p))
    child = flat.tex_children[tex][0] if sines < 0 else flat.tex_children[tex][1]
    return np_tex_sample(flat, int(child), u, v, p)


# ---- the comparisons -------------------------------------------------------------------------------------------------------------
def test_sphere_records_against_numpy():
    """Sphere / UVSphere / MovingSphere hit? on random geometry: same decision, t, p, normal, uv (both roots, origin inside,
    narrow ranges that reject the near or both roots)."""
    rng = np.random.default_rng(11)
    n = 3000
    c0 = rng.uniform(-5, 5, (n, 3))
    c1 = c0 + rng.uniform(-1, 1, (n, 3))
    rad = rng.uniform(0.1, 3.0, n)
    o = rng.uniform(-8, 8, (n, 3))
    target = c0 + rng.normal(size=(n, 3)) * (rad * rng.uniform(0.0, 1.4, n))[:, None]     # most rays hit, some graze or miss
    d = (target - o) * rng.uniform(0.2, 3.0, (n, 1))                                        # un-normalised, as the reference's rays are
    time = rng.uniform(0, 1, n)
    kind = rng.integers(0, 3, n)                                                            # 0 Sphere, 1 UVSphere, 2 MovingSphere
    lo = np.where(rng.random(n) < 0.3, rng.uniform(0.0, 1.5, n), 0.001)                     # ranges that cut roots off
    hi = np.where(rng.random(n) < 0.3, rng.uniform(0.5, 3.0, n), FMAX)
    cen = np.where((kind == 2)[:, None], np_center_at_time(c0, np.zeros(n), c1, np.ones(n), time), c0)
    hit, t, p, nrm = np_sphere_hit(cen, rad, o, d, lo, hi)
    u, v = np_sphere_uv(nrm)
    n_hit = 0
    for i in range(n):
        got = oracle.sphere_hit(c0[i], rad[i], o[i], d[i], lo[i], hi[i], time=time[i], center1=c1[i] if kind[i] == 2 else None,
                                t0=0.0, t1=1.0, uv=kind[i] == 1)
        assert (got is not None) == bool(hit[i]), f"case {i}: hit decision differs"
        if got is None:
            continue
        n_hit += 1
        assert got["t"] == pytest.approx(t[i], rel=1e-12, abs=1e-12)
        assert np.allclose(got["p"], p[i], rtol=1e-11, atol=1e-11)
        assert np.allclose(got["normal"], nrm[i], rtol=1e-10, atol=1e-10)
        if kind[i] == 1:
            assert np.allclose(got["uv"], [u[i], v[i]], rtol=1e-9, atol=1e-9)
        else:
            assert np.all(got["uv"] == 0)                                                   # :uv [0 0], hitable.clj:197
    assert 0.5 * n < n_hit < n


def test_scatter_and_emitted_against_numpy(random_scene_flat):
    """shader.clj:29-119 + texture.clj:14-50 on the benchmark scene's own records, explicit `rand-in-unit-sphere` / `rand` inputs:
    scattered ray, attenuation, emitted, absorb decisions of Lambertian / Metal / Dielectric / DiffuseLight."""
    flat, _, _ = random_scene_flat
    S = oracle.Scene(flat)
    rng = np.random.default_rng(12)
    ns = flat.n_spheres
    per = 12
    ids = np.repeat(np.arange(ns, dtype=np.int32), per)
    n = len(ids)
    c0r = flat.center0_r[ids].astype(np.float64)
    c1 = flat.center1[ids, :3].astype(np.float64)
    tt = flat.t0t1[ids].astype(np.float64)
    flags = flat.sphere_flags[ids]
    rad = np.abs(c0r[:, 3])
    inside = rng.random(n) < 0.25                                                           # start inside: exercises the far root / exit side of glass
    time = rng.uniform(0, 1, n).astype(np.float32)
    moving = (flags & RT_SPHERE_MOVING) != 0
    cen = np.where(moving[:, None], np_center_at_time(c0r[:, :3], tt[:, 0], c1, tt[:, 1], time.astype(np.float64)), c0r[:, :3])
    dirn = normalise(rng.normal(size=(n, 3)))
    o = np.where(inside[:, None], cen + dirn * (rad * rng.uniform(0, 0.9, n))[:, None], cen + dirn * (rad * rng.uniform(1.2, 4.0, n))[:, None])
    target = cen + rng.normal(size=(n, 3)) * (rad * 0.5)[:, None]
    d = (target - o) * rng.uniform(0.3, 2.5, (n, 1))
    o32, d32 = o.astype(np.float32), d.astype(np.float32)                                   # what the oracle is handed
    ball = (normalise(rng.normal(size=(n, 3))) * rng.random((n, 1)) ** (1 / 3)).astype(np.float32)
    u01 = rng.random(n).astype(np.float32)
    got = S.shade_batch(o32, d32, time, ids, ball, u01)

    o64, d64 = o32.astype(np.float64), d32.astype(np.float64)
    cen = np.where(moving[:, None], np_center_at_time(c0r[:, :3], tt[:, 0], c1, tt[:, 1], time.astype(np.float64)), c0r[:, :3])
    hit, t, p, nrm = np_sphere_hit(cen, rad, o64, d64, 0.001, FMAX)
    uu, vv = np_sphere_uv(nrm)
    is_uv = (flags & RT_SPHERE_UV) != 0
    uu, vv = np.where(is_uv, uu, 0.0), np.where(is_uv, vv, 0.0)
    assert np.array_equal(got["flags"] >= 0, hit)
    mat = flat.material_id[ids]
    mtype = flat.mat_type[mat]
    mparam = flat.mat_param[mat].astype(np.float64)
    mtex = flat.mat_tex[mat]
    b64, r64 = ball.astype(np.float64), u01.astype(np.float64)
    seen = {k: 0 for k in (RT_MAT_LAMBERTIAN, RT_MAT_METAL, RT_MAT_DIELECTRIC, RT_MAT_DIFFUSE_LIGHT)}
    absorbed = reflected_by_coin = refracted = total_internal = 0
    for i in np.nonzero(hit)[0]:
        assert got["t"][i] == pytest.approx(t[i], rel=1e-12)
        ty = int(mtype[i])
        seen[ty] += 1
        emitted = np.zeros(3)
        ok, so, sd, att = False, None, None, None
        if ty == RT_MAT_LAMBERTIAN:                     # shader.clj:29-36: target = p + normal + s; direction = target - p
            ok, so, sd = True, p[i], (p[i] + nrm[i] + b64[i]) - p[i]
            att = np_tex_sample(flat, int(mtex[i]), uu[i], vv[i], p[i])
        elif ty == RT_MAT_METAL:                        # shader.clj:46-59
            refl = np_reflect(normalise(d64[i]), nrm[i])
            sd = refl + mparam[i] * b64[i]
            ok = dot(sd, nrm[i]) > 0
            so = p[i]
            att = np_tex_sample(flat, int(mtex[i]), uu[i], vv[i], p[i])
            absorbed += not ok
        elif ty == RT_MAT_DIELECTRIC:                   # shader.clj:76-104
            ri = mparam[i]
            rdn = dot(d64[i], nrm[i])
            mag = np.sqrt(dot(d64[i], d64[i]))
            if rdn > 0:
                outward, nint, cosine = -nrm[i], ri, ri * (rdn / mag)
            else:
                outward, nint, cosine = nrm[i], 1.0 / ri, -(rdn / mag)
            can, refr = np_refract(d64[i], outward, np.float64(nint))
            ok, so, att = True, p[i], np.ones(3)
            if can:
                if r64[i] < np_schlick(cosine, ri):
                    sd = np_reflect(d64[i], nrm[i])     # NOT normalised first: shader.clj:95 reflects ray-direction as given
                    reflected_by_coin += 1
                else:
                    sd = refr
                    refracted += 1
            else:
                sd = np_reflect(d64[i], nrm[i])
                total_internal += 1
        else:                                           # shader.clj:114-119 DiffuseLight: scatter -> nil, emitted = sample tex
            assert ty == RT_MAT_DIFFUSE_LIGHT
            emitted = np_tex_sample(flat, int(mtex[i]), uu[i], vv[i], p[i])
        assert (got["flags"][i] == 1) == bool(ok), f"ray {i} (material {ty}): scatter / absorb decision differs"
        assert np.allclose(got["emitted"][i], emitted, rtol=1e-10, atol=1e-12)
        if ok:
            assert np.allclose(got["origin"][i], so, rtol=1e-10, atol=1e-10)
            assert np.allclose(got["dir"][i], sd, rtol=1e-8, atol=1e-9), f"ray {i} (material {ty})"
            assert np.allclose(got["atten"][i], att, rtol=1e-10, atol=1e-12)
    # the benchmark scene exercises every branch the hot path has
    assert all(v >= 10 for v in seen.values()), seen
    assert absorbed > 0 and reflected_by_coin > 0 and refracted > 0 and total_internal > 0, (absorbed, reflected_by_coin, refracted, total_internal)


def test_cameras_and_get_ray_against_numpy():
    """camera.clj:8-66: both constructors and get-ray with explicit `rand-in-unit-disk` / `rand` values."""
    rng = np.random.default_rng(13)
    for _ in range(40):
        lookfrom, lookat = rng.uniform(-10, 10, 3), rng.uniform(-2, 2, 3)
        vup = normalise(np.array([0.0, 1.0, 0.0]) + rng.normal(scale=0.2, size=3))
        vfov, aspect = rng.uniform(15, 80), rng.uniform(0.5, 2.5)
        aperture, focus, t0, t1 = rng.uniform(0, 0.6), rng.uniform(1, 15), 0.0, rng.uniform(0.2, 1.0)
        theta = vfov * (np.pi / 180.0)
        hh = np.tan(theta / 2.0)
        hw = aspect * hh
        w = normalise(lookfrom - lookat)
        u = normalise(np.cross(vup, w))
        v = np.cross(w, u)
        # thin lens, camera.clj:50-66
        lleft = lookfrom - (focus * hw * u + focus * hh * v + focus * w)
        horiz, vert = 2.0 * focus * hw * u, 2.0 * focus * hh * v
        cam = oracle.thin_lens_camera(lookfrom, lookat, vup, vfov, aspect, aperture, focus, t0, t1)
        for _ in range(8):
            s, t = rng.random(), rng.random()
            disk = rng.uniform(-0.7, 0.7, 2)
            tu = rng.random()
            rd = (aperture / 2.0) * disk                              # camera.clj:38-39
            off = u * rd[0] + v * rd[1]
            want_o = lookfrom + off
            want_d = lleft + s * horiz + t * vert - lookfrom - off
            want_t = t0 + (t1 - t0) * tu
            go, gd, gt = oracle.get_ray(1, cam, s, t, disk=tuple(disk), time_u=tu)
            # the camera record travels as float32 (the ABI's 24 floats): compare at that precision
            assert np.allclose(go, want_o, rtol=2e-6, atol=2e-6) and np.allclose(gd, want_d, rtol=2e-6, atol=2e-5)
            assert gt == pytest.approx(want_t, rel=2e-6, abs=1e-7)
        # pinhole, camera.clj:18-33 / 8-16
        lleft_p = lookfrom - (hw * u + hh * v + w)
        camp = oracle.pinhole_camera(lookfrom, lookat, vup, vfov, aspect)
        s, t = rng.random(), rng.random()
        go, gd, gt = oracle.get_ray(0, camp, s, t)
        assert np.allclose(go, lookfrom, rtol=2e-6, atol=2e-6)
        assert np.allclose(gd, lleft_p + s * (2.0 * hw * u) + t * (2.0 * hh * v) - lookfrom, rtol=2e-6, atol=2e-5)
        assert gt == 0.0


def test_resolve_against_numpy():
    """core.clj:52-57 and :105: mean over samples, sqrt, * 255.99, min with 255.99 (written there as an int cast of the min), rows
    flipped; NaN sums quantise to 0 like the JVM's (int NaN)."""
    rng = np.random.default_rng(14)
    ny, nx, nr = 24, 40, 16
    s = rng.uniform(0, 1.3, (ny, nx, 3)) * nr
    s[3, 5] = [np.nan, 4.0 * nr, 0.0]
    want = np.sqrt(s / nr) * 255.99
    want = np.where(np.isnan(want), 0.0, np.minimum(want, 255.99)).astype(np.uint8)[::-1]
    got = oracle.resolve(s, nr)
    assert np.array_equal(got, want)


# ---- generic leaves and instance wrappers: a tree-walking interpreter of the reference's records ---------------------------------
def interp_hit(obj, o, d, t_min, t_max):
    """hit? of one record of the Python mirror (hitable.py dataclasses = the reference's records), written like the Clojure:
    wrappers RECURSE into their item with a transformed ray (the oracle and the CUDA code instead flatten a leaf's wrappers into a
    chain of ops at marshalling time — a different formulation of the same thing).  Returns None or (t, p, normal, (u, v))."""
    from raytrace_clj_b200 import hitable as H

    if isinstance(obj, (H.RectXY, H.RectXZ, H.RectYZ)):                 # hitable.clj:272-293, 303-324, 334-355
        ax = {H.RectXY: (0, 1, 2), H.RectXZ: (0, 2, 1), H.RectYZ: (1, 2, 0)}[type(obj)]
        if isinstance(obj, H.RectXY):
            a0, b0, a1, b1 = obj.x0, obj.y0, obj.x1, obj.y1
        elif isinstance(obj, H.RectXZ):
            a0, b0, a1, b1 = obj.x0, obj.z0, obj.x1, obj.z1
        else:
            a0, b0, a1, b1 = obj.y0, obj.z0, obj.y1, obj.z1
        ia, ib, ik = ax
        with np.errstate(divide="ignore", invalid="ignore"):
            t = (obj.k - o[ik]) / d[ik]
        if not (t >= t_min and t <= t_max):
            return None
        a = o[ia] + t * d[ia]
        b = o[ib] + t * d[ib]
        if not (a >= a0 and a <= a1 and b >= b0 and b <= b1):
            return None
        n = np.zeros(3)
        n[ik] = 1.0
        return t, o + t * d, n, ((a - a0) / (a1 - a0), (b - b0) / (b1 - b0))
    if isinstance(obj, H.Triangle):                                       # hitable.clj:548-575 (Moeller-Trumbore, one-sided)
        e1, e2 = obj.v1 - obj.v0, obj.v2 - obj.v0
        pvec = np.cross(d, e2)
        det = float(np.dot(e1, pvec))
        if not det > 0.00000001:
            return None
        inv = 1.0 / det
        tvec = o - obj.v0
        u = float(np.dot(tvec, pvec)) * inv
        if not (u > 0 and u <= 1):
            return None
        qvec = np.cross(tvec, e1)
        v = float(np.dot(d, qvec)) * inv
        if not (v > 0 and u + v <= 1):
            return None
        t = float(np.dot(e2, qvec)) * inv
        if not (t >= t_min and t <= t_max):
            return None
        return t, o + t * d, np.cross(e1, e2), (u, v)
    if isinstance(obj, (H.Sphere, H.MovingSphere)):                       # hitable.clj:141-251 (H.UVSphere is a Sphere)
        time = interp_hit.time
        c = np_center_at_time(obj.center0, np.float64(obj.t0), obj.center1, np.float64(obj.t1), np.float64(time)) \
            if isinstance(obj, H.MovingSphere) else np.asarray(obj.center, np.float64)
        hit, t, p, n = np_sphere_hit(c, float(obj.radius), o, d, t_min, t_max)
        if not hit:
            return None
        uv = tuple(float(x) for x in np_sphere_uv(n)) if isinstance(obj, H.UVSphere) else (0.0, 0.0)
        return float(t), p, n, uv
    if isinstance(obj, H.Hitlist):                                        # hitable.clj:15-26: shrinking t-max, later equal hits lose
        best, far = None, t_max
        for item in obj.items:
            h = interp_hit(item, o, d, t_min, far)
            if h is not None:
                best, far = h, h[0]
        return best
    if isinstance(obj, H.Box):                                            # hitable.clj:491-497: the six sides
        return interp_hit(obj.sides, o, d, t_min, t_max)
    if isinstance(obj, H.FlipNormals):                                    # hitable.clj:375-381
        h = interp_hit(obj.item, o, d, t_min, t_max)
        return None if h is None else (h[0], h[1], -h[2], h[3])
    if isinstance(obj, H.Translate):                                      # hitable.clj:391-397
        h = interp_hit(obj.item, o - obj.offset, d, t_min, t_max)
        return None if h is None else (h[0], h[1] + obj.offset, h[2], h[3])
    if isinstance(obj, H.RotateY):                                        # hitable.clj:410-455
        s, c = float(np.float32(obj.sin_theta)), float(np.float32(obj.cos_theta))   # as marshalled (float32)
        ro = np.array([c * o[0] - s * o[2], o[1], s * o[0] + c * o[2]])
        rd = np.array([c * d[0] - s * d[2], d[1], s * d[0] + c * d[2]])
        h = interp_hit(obj.obj, ro, rd, t_min, t_max)
        if h is None:
            return None
        p, n = h[1], h[2]
        return (h[0], np.array([c * p[0] + s * p[2], p[1], -(s * p[0]) + c * p[2]]),
                np.array([c * n[0] + s * n[2], n[1], -(s * n[0]) + c * n[2]]), h[3])
    raise TypeError(type(obj))


interp_hit.time = 0.0          # the ray's time (only moving spheres look at it); set by the caller


def test_generic_leaves_and_wrapper_chains_against_interpreter():
    """Rectangles, triangles and random chains of Translate / RotateY / FlipNormals (hitable.clj:269-455, 548-575): the oracle
    (flattened op chains, as marshalled for the GPU) against a recursive interpreter of the records, on rays aimed at the leaf in
    world space — decision, t, p, normal, uv."""
    import random

    import raytrace_clj_b200 as rt
    from raytrace_clj_b200 import hitable as H
    from raytrace_clj_b200 import shader as shad
    from raytrace_clj_b200 import texture as tex
    from raytrace_clj_b200.util import vec3

    gray = shad.lambertian(albedo=tex.constant(color=vec3(.5, .5, .5)))
    rng = np.random.default_rng(15)
    pyr = random.Random(15)
    n_hits = n_cases = 0
    kinds_hit = set()
    for case in range(60):
        kind = case % 4
        if kind < 3:
            # every parameter float32-representable: the scene travels as float32 (rt_scene_ext), the interpreter sees the same values
            a0, b0 = (float(x) for x in rng.uniform(-2, 0, 2).astype(np.float32))
            a1, b1 = float(np.float32(a0 + rng.uniform(0.5, 3))), float(np.float32(b0 + rng.uniform(0.5, 3)))
            k = float(np.float32(rng.uniform(-2, 2)))
            leaf = [H.rect_xy(x0=a0, y0=b0, x1=a1, y1=b1, k=k, material=gray), H.rect_xz(x0=a0, z0=b0, x1=a1, z1=b1, k=k, material=gray),
                    H.rect_yz(y0=a0, z0=b0, y1=a1, z1=b1, k=k, material=gray)][kind]
            ia, ib, ik = [(0, 1, 2), (0, 2, 1), (1, 2, 0)][kind]
            pts = np.zeros((40, 3))
            pts[:, ia] = rng.uniform(a0 - 0.4, a1 + 0.4, 40)
            pts[:, ib] = rng.uniform(b0 - 0.4, b1 + 0.4, 40)
            pts[:, ik] = k
        else:
            v = rng.uniform(-2, 2, (3, 3)).astype(np.float32).astype(np.float64)
            leaf = H.triangle(v0=vec3(*v[0]), v1=vec3(*v[1]), v2=vec3(*v[2]), material=gray)
            w = rng.dirichlet([1, 1, 1], 40) * rng.uniform(0.6, 1.3, (40, 1))     # some outside the triangle
            pts = w @ v
        obj = leaf
        chain = []
        for _ in range(pyr.randrange(0, 4)):
            op = pyr.choice("trf")
            chain.append(op)
            if op == "t":
                obj = H.translate(item=obj, offset=vec3(*rng.uniform(-3, 3, 3).astype(np.float32).astype(np.float64)))
            elif op == "r":
                obj = H.rotate_y(item=obj, theta=float(rng.uniform(-180, 180)))
            else:
                obj = H.flip_normals(item=obj)
        flat = rt.native.marshal_world(H.hitlist(items=[obj]))
        S = oracle.Scene(flat)

        # the target points in WORLD space: push the leaf-space points through the chain, innermost wrapper first
        def to_world(o_, p_):
            if isinstance(o_, H.Translate):
                return to_world(o_.item, p_) + o_.offset
            if isinstance(o_, H.RotateY):
                q = to_world(o_.obj, p_)
                s, c = o_.sin_theta, o_.cos_theta
                return np.stack([c * q[:, 0] + s * q[:, 2], q[:, 1], -(s * q[:, 0]) + c * q[:, 2]], axis=1)
            if isinstance(o_, H.FlipNormals):
                return to_world(o_.item, p_)
            return p_

        wp = to_world(obj, pts)
        org = wp + rng.normal(size=wp.shape) * 4.0
        dr = (wp - org) * rng.uniform(0.3, 2.0, (len(wp), 1))
        org32, dr32 = org.astype(np.float32), dr.astype(np.float32)
        t, ids, pnuv = S.hit(org32, dr32, None, 0.001, FMAX, details=True)
        for i in range(len(wp)):
            want = interp_hit(obj, org32[i].astype(np.float64), dr32[i].astype(np.float64), 0.001, FMAX)
            n_cases += 1
            # a point within rounding of a rectangle's or a triangle's edge may fall either way in the two formulations
            edge = want is not None and (min(want[3][0], want[3][1], 1 - want[3][0], 1 - want[3][1]) < 1e-9 or
                                         (kind == 3 and 1 - want[3][0] - want[3][1] < 1e-9))
            if edge:
                continue
            assert (ids[i] >= 0) == (want is not None), f"case {case} {chain} ray {i}: hit decision differs"
            if want is None:
                continue
            n_hits += 1
            kinds_hit.add((kind, tuple(chain)))
            assert t[i] == pytest.approx(want[0], rel=1e-9, abs=1e-9), f"case {case} {chain} ray {i}"
            assert np.allclose(pnuv[i][0:3], want[1], rtol=1e-8, atol=1e-8), f"case {case} {chain} ray {i}: p"
            assert np.allclose(pnuv[i][3:6], want[2], rtol=1e-8, atol=1e-8), f"case {case} {chain} ray {i}: normal"
            assert np.allclose(pnuv[i][6:8], want[3], rtol=1e-7, atol=1e-8), f"case {case} {chain} ray {i}: uv"
    assert n_hits > 0.2 * n_cases and len(kinds_hit) > 25, (n_hits, n_cases, len(kinds_hit))


# ---- `pixel` + `color` (core.clj:17-57) as a plain Python loop over the numpy pieces above, on the oracle's replay stream -----------
def _replay_block(seed, pixel, sample, bounce, blk):
    """The four uniforms of Philox block (pixel, sample, bounce << 16 | blk, "RTB2") under key = seed (replay mode)."""
    from helpers import RNG_DOMAIN, philox4x32_10

    ctr = np.array([[pixel, sample, (bounce << 16) | blk, RNG_DOMAIN]], np.uint32)
    key = np.array([[seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF]], np.uint32)
    return (philox4x32_10(ctr, key)[0] >> np.uint32(8)).astype(np.float64) / 16777216.0


def np_color_of_sample(flat, cam_type, cam, nx, ny, pix, smp, seed, max_depth=50):
    """One (pixel, sample): jitter + get-ray (core.clj:49-50, camera.clj:35-48), then `color` (core.clj:17-41) with the
    Hitlist closest hit (hitable.clj:15-26) by brute force over every sphere.  Returns (radiance, rays, termination)."""
    cam = np.asarray(cam, np.float32).astype(np.float64)
    origin, lleft, horiz, vert, cu, cv = (cam[3 * k:3 * k + 3] for k in range(6))
    i, j = pix % nx, pix // nx
    u0 = _replay_block(seed, pix, smp, 0, 0)
    s, t = (float(np.float32(i)) + u0[0]) / nx, (float(np.float32(j)) + u0[1]) / ny
    if cam_type == 1:
        ud = _replay_block(seed, pix, smp, 0, 1)
        rd = (cam[21] / 2.0) * np.array([np.sqrt(ud[0]) * np.cos(2 * np.pi * ud[1]), np.sqrt(ud[0]) * np.sin(2 * np.pi * ud[1])])
        off = cu * rd[0] + cv * rd[1]
        o, d, time = origin + off, lleft + s * horiz + t * vert - origin - off, cam[22] + (cam[23] - cam[22]) * u0[2]
    else:
        o, d, time = origin, lleft + s * horiz + t * vert - origin, 0.0
    ns = flat.n_spheres
    c0r = flat.center0_r[:ns].astype(np.float64)
    c1 = flat.center1[:ns, :3].astype(np.float64)
    tt = flat.t0t1[:ns].astype(np.float64)
    moving = (flat.sphere_flags[:ns] & RT_SPHERE_MOVING) != 0
    is_uv = (flat.sphere_flags[:ns] & RT_SPHERE_UV) != 0
    rad = np.abs(c0r[:, 3])
    atten, accum = np.ones(3), np.zeros(3)
    depth, rays = max_depth, 0
    while True:
        rays += 1
        with np.errstate(invalid="ignore", divide="ignore"):
            f = np.where(moving, (time - tt[:, 0]) / np.where(moving, tt[:, 1] - tt[:, 0], 1.0), 0.0)[:, None]
        cen = np.where(moving[:, None], c0r[:, :3] * (1.0 - f) + c1 * f, c0r[:, :3])
        hit, th, ph, nh = np_sphere_hit(cen, rad, o[None, :], d[None, :], 0.001, FMAX)
        if not hit.any():
            return accum, rays, 4                                        # miss -> accum (core.clj:40-41)
        k = int(np.argmin(np.where(hit, th, np.inf)))                    # closest; the FIRST of equal hits (hitable.clj:17-26)
        p, n = ph[k], nh[k]
        uu, vv = np_sphere_uv(n) if is_uv[k] else (0.0, 0.0)
        m = int(flat.material_id[k])
        ty, param, mt = int(flat.mat_type[m]), float(flat.mat_param[m]), int(flat.mat_tex[m])
        emitted = np_tex_sample(flat, mt, uu, vv, p) if ty == RT_MAT_DIFFUSE_LIGHT else np.zeros(3)
        scat, why = None, 3                                              # why: 1 light, 2 absorbed, 3 depth
        if depth > 0:                                                    # (and (pos? depth) (scatter ...)), core.clj:26-27
            ub = _replay_block(seed, pix, smp, rays, 1)
            br, bz = np.cbrt(ub[0]), 1.0 - 2.0 * ub[1]
            bs = np.sqrt(max(0.0, 1.0 - bz * bz)) * br
            ball = np.array([bs * np.cos(2 * np.pi * ub[2]), bs * np.sin(2 * np.pi * ub[2]), bz * br])
            if ty == RT_MAT_LAMBERTIAN:
                scat = (p, (p + n + ball) - p, np_tex_sample(flat, mt, uu, vv, p))
            elif ty == RT_MAT_METAL:
                sd = np_reflect(normalise(d), n) + param * ball
                scat, why = ((p, sd, np_tex_sample(flat, mt, uu, vv, p)), 0) if dot(sd, n) > 0 else (None, 2)
            elif ty == RT_MAT_DIELECTRIC:
                rdn, mag = dot(d, n), np.sqrt(dot(d, d))
                outward, nint, cosine = (-n, param, param * (rdn / mag)) if rdn > 0 else (n, 1.0 / param, -(rdn / mag))
                can, refr = np_refract(d, outward, np.float64(nint))
                scat = (p, np_reflect(d, n) if (not can or ub[3] < np_schlick(cosine, param)) else refr, np.ones(3))
            else:
                why = 1                                                  # DiffuseLight: scatter -> nil
        if scat is None:
            return accum + atten * emitted, rays, why
        accum = accum + atten * emitted                                  # with the OLD attenuation, core.clj:32-34
        atten = atten * scat[2]
        o, d = scat[0], scat[1]                                          # the ray keeps its time (shader.clj:34, 54, 96)
        depth -= 1


@pytest.mark.parametrize("aperture", [0.0, 0.4])
def test_pixel_and_color_loop_against_numpy(random_scene_flat, aperture):
    """core.clj:17-57 end to end on the benchmark scene: the oracle's replay of chosen (pixel, sample) pairs against the
    plain-Python loop above fed the same Philox uniforms — same number of rays, same termination, same radiance, at depth 50 and
    at depth cutoffs 0 / 1 / 3 (the `(pos? depth)` gate and the emitted-only return)."""
    flat, cam_type, cam = random_scene_flat
    nx, ny = 1200, 800
    if aperture > 0:                                                       # a real lens: the disk sample and the shutter time matter
        cam = oracle.thin_lens_camera([13, 2, 3], [0, 0, 0], [0, 1, 0], 20.0, nx / ny, aperture, 10.0, 0.0, 1.0).astype(np.float32)
        cam_type = 1
    S = oracle.Scene(flat)
    rng = np.random.default_rng(16)
    n = 160
    pix = rng.integers(0, nx * ny, n).astype(np.int32)
    pix[:40] = (rng.integers(300, 500, 40) * nx + rng.integers(300, 900, 40)).astype(np.int32)   # the middle of the frame: the big spheres
    smp = rng.integers(0, 1000, n).astype(np.int32)
    terms = set()
    for depth in (50, 0, 1, 3):
        rad, nr, term, _ = S.trace_paths(cam_type, cam, nx, ny, pix, smp, depth, seed=77)
        for q in range(n if depth == 50 else 40):
            w_rad, w_nr, w_term = np_color_of_sample(flat, cam_type, cam, nx, ny, int(pix[q]), int(smp[q]), 77, depth)
            assert (nr[q], term[q]) == (w_nr, w_term), f"depth {depth} path {q}: rays / termination {nr[q], term[q]} vs {w_nr, w_term}"
            assert np.allclose(rad[q], w_rad, rtol=1e-9, atol=1e-12), f"depth {depth} path {q}"
            terms.add(w_term)
    assert {1, 3} <= terms                                                 # light (the sky dome encloses the scene: no miss) and depth cutoff


def test_color_loop_on_an_open_scene_misses_and_lights():
    """The same comparison on a scene WITHOUT an enclosing sky: paths end by missing everything (accum, core.clj:40-41), on a
    DiffuseLight sphere, by Metal absorption, or at the depth cutoff — every way `color` can return."""
    import raytrace_clj_b200 as rt
    from raytrace_clj_b200 import hitable as H
    from raytrace_clj_b200 import shader as shad
    from raytrace_clj_b200 import texture as tex
    from raytrace_clj_b200.util import vec3

    const = lambda r, g, b: tex.constant(color=vec3(r, g, b))  # noqa: E731
    items = [H.sphere(center=vec3(0, -100.5, -1), radius=100.0, material=shad.lambertian(albedo=tex.checkerboard(
                 tex0=const(.2, .3, .1), tex1=const(.9, .9, .9), scale=10.0))),
             H.sphere(center=vec3(0, 0, -1), radius=0.5, material=shad.lambertian(albedo=const(.8, .3, .3))),
             H.sphere(center=vec3(1, 0, -1), radius=0.5, material=shad.metal(albedo=const(.8, .6, .2), fuzz=0.9)),
             H.sphere(center=vec3(-1, 0, -1), radius=0.5, material=shad.dielectric(ri=1.5)),
             H.sphere(center=vec3(-1, 0, -1), radius=-0.45, material=shad.dielectric(ri=1.5)),     # the book's hollow glass
             H.moving_sphere(center0=vec3(0, 1.2, -1), t0=0.0, center1=vec3(0.3, 1.4, -1), t1=1.0, radius=0.3,
                             material=shad.diffuse_light(tex=const(4, 4, 4)))]
    flat = rt.native.marshal_world(H.hitlist(items=items))
    S = oracle.Scene(flat)
    nx, ny = 200, 100
    cam = oracle.thin_lens_camera([3, 1.5, 2], [0, 0, -1], [0, 1, 0], 35.0, nx / ny, 0.1, 4.0, 0.0, 1.0).astype(np.float32)
    rng = np.random.default_rng(17)
    n = 220
    pix = rng.integers(0, nx * ny, n).astype(np.int32)
    smp = rng.integers(0, 64, n).astype(np.int32)
    rad, nr, term, _ = S.trace_paths(1, cam, nx, ny, pix, smp, 50, seed=5)
    seen = set()
    for q in range(n):
        w_rad, w_nr, w_term = np_color_of_sample(flat, 1, cam, nx, ny, int(pix[q]), int(smp[q]), 5, 50)
        assert (nr[q], term[q]) == (w_nr, w_term), f"path {q}: rays / termination {nr[q], term[q]} vs {w_nr, w_term}"
        assert np.allclose(rad[q], w_rad, rtol=1e-9, atol=1e-12), f"path {q}"
        seen.add(w_term)
    assert {1, 4} <= seen, seen


# ---- Perlin noise, procedural / image textures, the constant medium ----------------------------------------------------------------
def np_perlin_noise(flat, p):
    """perlin.clj:19-50 on the MARSHALLED tables (random-vectors, perm-x / -y / -z)."""
    vec = flat.perlin_vectors.astype(np.float64)
    perm = flat.perlin_perm
    p = np.asarray(p, np.float64)
    ijk = np.floor(p).astype(np.int64)
    uvw = p - ijk
    hh = uvw * uvw * (3 - 2 * uvw)                                       # the Hermite-smoothed blend weights (perlin-interp)
    acc = 0.0
    for i in range(2):
        for j in range(2):
            for k in range(2):
                c = vec[perm[0][(ijk[0] + i) & 255] ^ perm[1][(ijk[1] + j) & 255] ^ perm[2][(ijk[2] + k) & 255]]
                w = uvw - np.array([i, j, k], np.float64)
                acc += ((i * hh[0] + (1.0 - i) * (1.0 - hh[0])) * (j * hh[1] + (1.0 - j) * (1.0 - hh[1])) *
                        (k * hh[2] + (1.0 - k) * (1.0 - hh[2])) * float(np.dot(w, c)))
    return acc


def np_turbulence(flat, p, depth):
    """perlin.clj:52-64."""
    acc, pt, w = 0.0, np.asarray(p, np.float64), 1.0
    for _ in range(depth):
        acc += w * np_perlin_noise(flat, pt)
        pt = 2.0 * pt
        w = w / 2.0
    return abs(acc)


def test_perlin_and_procedural_textures_against_numpy():
    import raytrace_clj_b200 as rt
    from raytrace_clj_b200 import hitable as H
    from raytrace_clj_b200 import shader as shad
    from raytrace_clj_b200 import texture as tex
    from raytrace_clj_b200.util import vec3

    mats = [shad.lambertian(albedo=t) for t in (tex.perlin_noise(scale=3.0), tex.perlin_turbulence(scale=4.0, depth=7),
                                               tex.marble(scale=5.0, depth=6))]
    flat = rt.native.marshal_world(H.hitlist(items=[H.sphere(center=vec3(3 * i, 0, 0), radius=1, material=m) for i, m in enumerate(mats)]))
    S = oracle.Scene(flat)
    rng = np.random.default_rng(18)
    tid = {int(t): i for i, t in enumerate(flat.tex_type)}
    for p in np.concatenate([rng.uniform(-40, 40, (300, 3)), rng.uniform(-1, 1, (100, 3)), [[-0.25, 300.5, -513.75]]]):
        assert S.perlin_noise(p) == pytest.approx(np_perlin_noise(flat, p), rel=1e-10, abs=1e-13)
    for p in rng.uniform(-6, 6, (60, 3)):
        assert S.perlin_turbulence(p, 7) == pytest.approx(np_turbulence(flat, p, 7), rel=1e-10, abs=1e-13)
        # texture.clj:60-98: 0.5 (1 + noise(scale p)), 0.5 (1 + turbulence(scale p, depth)), 0.5 (1 + sin(scale z + 10 turbulence(p, depth)))
        assert np.allclose(S.tex_sample(tid[rt.native.RT_TEX_PERLIN_NOISE], 0, 0, p), 0.5 * (1 + np_perlin_noise(flat, 3.0 * p)), rtol=1e-10)
        assert np.allclose(S.tex_sample(tid[rt.native.RT_TEX_PERLIN_TURB], 0, 0, p), 0.5 * (1 + np_turbulence(flat, 4.0 * p, 7)), rtol=1e-10)
        assert np.allclose(S.tex_sample(tid[rt.native.RT_TEX_MARBLE], 0, 0, p),
                           0.5 * (1 + np.sin(5.0 * p[2] + 10.0 * np_turbulence(flat, p, 6))), rtol=1e-9, atol=1e-12)


def test_image_map_and_flips_against_numpy():
    """texture.clj:103-133: (int (* u width)), (int (* v height)), components / 255, under FlipTextureU / FlipTextureV."""
    import raytrace_clj_b200 as rt
    from raytrace_clj_b200 import hitable as H
    from raytrace_clj_b200 import shader as shad
    from raytrace_clj_b200 import texture as tex
    from raytrace_clj_b200.util import vec3

    rng = np.random.default_rng(19)
    img = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)              # height 37, width 53
    base = tex.image_map(image=img)
    flat = rt.native.marshal_world(H.hitlist(items=[
        H.uv_sphere(center=vec3(0, 0, 0), radius=1, material=shad.lambertian(albedo=tex.flip_texture_v(tex=base))),
        H.uv_sphere(center=vec3(3, 0, 0), radius=1, material=shad.lambertian(albedo=tex.flip_texture_u(tex=base)))]))
    S = oracle.Scene(flat)
    tid = {int(t): i for i, t in enumerate(flat.tex_type)}
    for u, v in rng.uniform(0, 0.999999, (400, 2)):
        px = lambda uu, vv: img[int(vv * 37), int(uu * 53)] / 255.0      # noqa: E731   (get-pixel image i j): column i, row j
        assert np.allclose(S.tex_sample(tid[rt.native.RT_TEX_IMAGE_MAP], u, v, [0, 0, 0]), px(u, v), atol=1e-15)
        assert np.allclose(S.tex_sample(tid[rt.native.RT_TEX_FLIP_V], u, v, [0, 0, 0]), px(u, 1.0 - v), atol=1e-15)
        assert np.allclose(S.tex_sample(tid[rt.native.RT_TEX_FLIP_U], u, v, [0, 0, 0]), px(1.0 - u, v), atol=1e-15)


def test_constant_medium_hit_against_numpy():
    """hitable.clj:516-543 with its `rand` fixed at 0.5 (what the oracle draws without a path context): entry / exit through the
    boundary over (-MAX, MAX) and (t1 + 0.0001, MAX), clamps to [t-min, t-max] and to 0, distance in units of |direction|."""
    import math

    import raytrace_clj_b200 as rt
    from raytrace_clj_b200 import hitable as H
    from raytrace_clj_b200 import shader as shad
    from raytrace_clj_b200 import texture as tex
    from raytrace_clj_b200.util import vec3

    rng = np.random.default_rng(20)
    n_hit = n_all = 0
    for case in range(12):
        c = rng.uniform(-2, 2, 3).astype(np.float32).astype(np.float64)
        r = float(np.float32(rng.uniform(0.5, 2.0)))
        rho = float(np.float32(rng.uniform(0.2, 3.0)))
        ball = H.sphere(center=vec3(*c), radius=r, material=shad.dielectric(ri=1.5))
        flat = rt.native.marshal_world(H.hitlist(items=[H.constant_medium(boundary=ball, density=rho, albedo=tex.constant(color=vec3(1, 1, 1)))]))
        S = oracle.Scene(flat)
        m = 60
        o = (c + rng.normal(size=(m, 3)) * r * rng.uniform(0.0, 2.5, (m, 1))).astype(np.float32)       # inside and outside the boundary
        d = ((c + rng.normal(size=(m, 3)) * r * 0.6 - o) * rng.uniform(0.2, 2.0, (m, 1))).astype(np.float32)
        t_min = np.where(rng.random(m) < 0.3, rng.uniform(0, 2, m), 0.001)
        t_max = np.where(rng.random(m) < 0.3, rng.uniform(0.5, 4, m), FMAX)
        for i in range(m):
            oi, di = o[i].astype(np.float64), d[i].astype(np.float64)
            got_t, got_id = S.hit(o[i:i + 1], d[i:i + 1], None, float(t_min[i]), float(t_max[i]))
            h1, t1, _, _ = np_sphere_hit(c, r, oi, di, -FMAX, FMAX)
            want = None
            if h1:
                h2, t2, _, _ = np_sphere_hit(c, r, oi, di, t1 + 0.0001, FMAX)
                if h2:
                    a, b = (t_min[i] if t1 < t_min[i] else t1), (t_max[i] if t2 > t_max[i] else t2)
                    if a < b:
                        a = 0.0 if a < 0 else a
                        mag = math.sqrt(float(np.dot(di, di)))
                        hd = -(math.log(0.5) / rho)
                        if hd < (b - a) * mag:
                            want = a + hd / mag
            n_all += 1
            assert (got_id[0] >= 0) == (want is not None), f"case {case} ray {i}"
            if want is not None:
                n_hit += 1
                assert got_t[0] == pytest.approx(want, rel=1e-10, abs=1e-12)
    assert 0.2 * n_all < n_hit < 0.95 * n_all


def test_aabb_and_bvh_world_against_numpy(random_scene_flat):
    """AABB.hit? (hitable.clj:36-48: slabs by min / max of (v - o) / d, `(> tmax tmin)`) on random boxes and rays, and the whole
    bvh-node world (hitable.clj:97-106) — the oracle's reference-style BVH traversal must return the brute-force numpy closest
    hit on the benchmark scene (its boxes can only cull: a wrong box or child order would lose hits)."""
    rng = np.random.default_rng(21)
    n_true = 0
    for _ in range(3000):
        lo = rng.uniform(-3, 2, 3)
        hi = lo + rng.uniform(0.1, 3, 3)
        o = rng.uniform(-6, 6, 3)
        d = rng.normal(size=3) * rng.uniform(0.2, 3) if rng.random() < 0.4 else (lo + (hi - lo) * rng.uniform(-0.3, 1.3, 3) - o) * rng.uniform(0.2, 2)
        t_min, t_max = (0.001, FMAX) if rng.random() < 0.6 else (rng.uniform(0, 3), rng.uniform(3, 12))
        m, nn = (lo - o) / d, (hi - o) / d
        want = min(np.maximum(m, nn).min(), t_max) > max(np.minimum(m, nn).max(), t_min)
        assert oracle.aabb_hit(lo, hi, o, d, t_min, t_max) == bool(want)
        n_true += bool(want)
    assert 300 < n_true < 2700
    flat, cam_type, cam = random_scene_flat
    S = oracle.Scene(flat)
    S.build_bvh(0.0, 1.0, seed=3)
    ns = flat.n_spheres
    c0r = flat.center0_r[:ns].astype(np.float64)
    c1 = flat.center1[:ns, :3].astype(np.float64)
    tt = flat.t0t1[:ns].astype(np.float64)
    moving = (flat.sphere_flags[:ns] & RT_SPHERE_MOVING) != 0
    m = 400
    o = rng.uniform(-12, 12, (m, 3)).astype(np.float32)
    o[:, 1] = rng.uniform(0.05, 4, m)
    d = rng.normal(size=(m, 3)).astype(np.float32)
    tm = rng.random(m).astype(np.float32)
    t_bvh, id_bvh = S.hit(o, d, tm, 0.001, FMAX, use_bvh=True)
    for i in range(m):
        f = np.where(moving, (float(tm[i]) - tt[:, 0]) / np.where(moving, tt[:, 1] - tt[:, 0], 1.0), 0.0)[:, None]
        cen = np.where(moving[:, None], c0r[:, :3] * (1.0 - f) + c1 * f, c0r[:, :3])
        hit, th, _, _ = np_sphere_hit(cen, np.abs(c0r[:, 3]), o[i].astype(np.float64)[None], d[i].astype(np.float64)[None], 0.001, FMAX)
        assert hit.any() == (id_bvh[i] >= 0)
        if hit.any():
            k = int(np.argmin(np.where(hit, th, np.inf)))
            assert id_bvh[i] == k and t_bvh[i] == pytest.approx(th[k], rel=1e-12)


# ---- materials and textures as a tree-walking interpreter of the mirror's records (for the JVM-vector stand-in) -----------------------
def interp_tex(t, u, v, p):
    """texture.clj:14-50 on the records of raytrace_clj_b200/texture.py."""
    from raytrace_clj_b200 import texture as T

    if isinstance(t, T.Constant):
        return np.asarray(t.color, np.float64)
    if isinstance(t, T.UVGradient):
        co, cu, cv, cuv = (np.asarray(x, np.float64) for x in (t.co, t.cu, t.cv, t.cuv))
        return (cuv * (1 - u) + cv * u) * (1 - v) + (cu * (1 - u) + co * u) * v
    if isinstance(t, T.Checkerboard):
        return interp_tex(t.tex0 if np.prod(np.sin(t.scale * np.asarray(p, np.float64))) < 0 else t.tex1, u, v, p)
    raise TypeError(type(t))


def interp_scatter(m, o, d, time, h, ball, rnd):
    """shader.clj:29-143 on the records of raytrace_clj_b200/shader.py; h = (t, p, normal, (u, v)).  Returns
    (scattered (o, d, time, attenuation) | None, emitted)."""
    from raytrace_clj_b200 import shader as M

    t, p, n, (u, v) = h
    ball = np.asarray(ball, np.float64)
    zero = np.zeros(3)
    if isinstance(m, M.Lambertian):
        return (p, (p + n + ball) - p, time, interp_tex(m.albedo, u, v, p)), zero
    if isinstance(m, M.Metal):
        sd = np_reflect(normalise(d), n) + m.fuzz * ball
        return ((p, sd, time, interp_tex(m.albedo, u, v, p)) if dot(sd, n) > 0 else None), zero
    if isinstance(m, M.Dielectric):
        rdn, mag = dot(d, n), np.sqrt(dot(d, d))
        outward, nint, cosine = (-n, m.ri, m.ri * (rdn / mag)) if rdn > 0 else (n, 1.0 / m.ri, -(rdn / mag))
        can, refr = np_refract(d, outward, np.float64(nint))
        sd = np_reflect(d, n) if (not can or rnd < np_schlick(cosine, m.ri)) else refr
        return (p, sd, time, np.ones(3)), zero
    if isinstance(m, M.DiffuseLight):
        return None, interp_tex(m.tex, u, v, p)
    if isinstance(m, M.Isotropic):             # shader.clj:129-138: direction = the ball sample, TIME = the hit's t
        return (p, ball, t, interp_tex(m.albedo, u, v, p)), zero
    raise TypeError(type(m))
