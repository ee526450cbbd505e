"""ctypes wrapper of the CPU oracle (oracle/oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, by ``__graft_entry__.smoke()`` and by
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs, never by the product package.

The scene argument of every function is any object with the marshalled SoA attributes
(``center0_r, center1, t0t1, sphere_flags, material_id, mat_type, mat_param, mat_tex,
tex_type, tex_params, tex_children`` as numpy arrays) — the same buffers the C ABI takes.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_u8p = C.POINTER(C.c_uint8)


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (g++ only)."""
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_scene_create.restype = C.c_void_p
        _lib.orc_scene_create_ex.restype = C.c_void_p
        _lib.orc_schlick.restype = C.c_double
        _lib.orc_schlick.argtypes = [C.c_double, C.c_double]
        _lib.orc_max_threads.restype = C.c_int
    return _lib


def _p(a, ty):
    return a.ctypes.data_as(ty) if a is not None else None


def _d3(x):
    return (C.c_double * 3)(*[float(v) for v in x])


PATH_BOUNCE_DTYPE = np.dtype([("o", np.float32, 3), ("time", np.float32), ("d", np.float32, 3), ("hit_id", np.int32),
                              ("t", np.float64)])


class Scene:
    """Owns an oracle scene handle built from marshalled SoA buffers (rt_scene_desc + optional rt_scene_ext)."""

    def __init__(self, flat):
        self.flat = flat
        L = lib()
        c = np.ascontiguousarray
        g = lambda name, ty: None if getattr(flat, name, None) is None else c(getattr(flat, name), ty)  # noqa: E731
        self.n_boundary = int(getattr(flat, "n_boundary", 0) or 0)
        self.n = int(flat.center0_r.shape[0]) - self.n_boundary
        self._keep = [
            c(flat.center0_r, np.float32), c(flat.center1, np.float32), c(flat.t0t1, np.float32),
            c(flat.sphere_flags, np.uint32), c(flat.material_id, np.int32), c(flat.mat_type, np.int32),
            c(flat.mat_param, np.float32), c(flat.mat_tex, np.int32), c(flat.tex_type, np.int32),
            c(flat.tex_params, np.float32), c(flat.tex_children, np.int32),
            g("prim_type", np.int32), g("prim_params", np.float32), g("prim_aux", np.int32), g("prim_xform", np.int32),
            g("xform_ops", np.int32), g("xform_params", np.float32), g("perlin_vectors", np.float32),
            g("perlin_perm", np.int32), g("image_wh", np.int32), g("image_offset", np.int64), g("image_rgb", np.uint8),
        ]
        k = self._keep
        self.h = C.c_void_p(
            L.orc_scene_create_ex(
                C.c_int(self.n), _p(k[0], _f32p), _p(k[1], _f32p), _p(k[2], _f32p), _p(k[3], _u32p), _p(k[4], _i32p),
                C.c_int(len(k[5])), _p(k[5], _i32p), _p(k[6], _f32p), _p(k[7], _i32p),
                C.c_int(len(k[8])), _p(k[8], _i32p), _p(k[9], _f32p), _p(k[10], _i32p),
                C.c_int(self.n_boundary), _p(k[11], _i32p), _p(k[12], _f32p), _p(k[13], _i32p), _p(k[14], _i32p),
                C.c_int(0 if k[15] is None else len(k[15])), _p(k[15], _i32p), _p(k[16], _f32p),
                C.c_int(int(getattr(flat, "tie_rule", 0) or 0)), _p(k[17], _f32p), _p(k[18], _i32p),
                C.c_int(0 if k[19] is None else len(k[19])), _p(k[19], _i32p), _p(k[20], C.POINTER(C.c_int64)),
                _p(k[21], _u8p),
            )
        )
        self.has_bvh = False

    def __del__(self):
        try:
            if self.h:
                lib().orc_scene_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def build_bvh(self, t0=0.0, t1=1.0, seed=1):
        """make-bvh (hitable.clj:108-123) over the world list; returns the node count."""
        n = lib().orc_scene_build_bvh(self.h, C.c_double(t0), C.c_double(t1), C.c_uint64(seed))
        self.has_bvh = True
        return int(n)

    def prim_bbox(self, i, t0=0.0, t1=1.0):
        out = (C.c_double * 6)()
        lib().orc_prim_bbox(self.h, C.c_int(i), C.c_double(t0), C.c_double(t1), out)
        return np.array(list(out))

    # `hit? world` over the flattened leaves (Hitlist, hitable.clj:15-26) or through the reference-style BVH
    def hit(self, origins, dirs, times=None, t_min=0.001, t_max=float(np.finfo(np.float32).max), second=False,
            use_bvh=False, details=False, stats=None):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        n = o.shape[0]
        tm = None if times is None else np.ascontiguousarray(times, np.float32)
        t = np.empty(n, np.float64)
        ids = np.empty(n, np.int32)
        t2 = np.empty(n, np.float64) if second else None
        pnuv = np.zeros((n, 8), np.float64) if details else None
        st = np.zeros(2, np.uint64) if stats is not None else None
        lib().orc_hit_ex(self.h, C.c_int(n), _p(o, _f32p), _p(d, _f32p), _p(tm, _f32p), C.c_double(t_min),
                         C.c_double(t_max), C.c_int(1 if use_bvh else 0), _p(t, _f64p), _p(ids, _i32p), _p(t2, _f64p),
                         _p(pnuv, _f64p), _p(st, _u64p))
        if stats is not None:
            stats["aabb_tests"] = stats.get("aabb_tests", 0) + int(st[0])
            stats["leaf_tests"] = stats.get("leaf_tests", 0) + int(st[1])
        out = (t, ids)
        if second:
            out = out + (t2,)
        if details:
            out = out + (pnuv,)
        return out

    def trace_paths(self, cam_type, cam, nx, ny, pixel, sample, max_depth=50, seed=1, log_bounces=0, n_threads=0):
        """Replay of chosen (pixel, sample) pairs on the Philox counters: (radiance [n,3] f64, nrays, term, log)."""
        camf = np.ascontiguousarray(cam, np.float32)
        pix = np.ascontiguousarray(pixel, np.int32)
        smp = np.ascontiguousarray(sample, np.int32)
        n = len(pix)
        rad = np.zeros((n, 3), np.float64)
        nr = np.zeros(n, np.int32)
        term = np.zeros(n, np.int32)
        log = np.zeros((n, log_bounces), PATH_BOUNCE_DTYPE) if log_bounces > 0 else None
        lib().orc_trace_paths(self.h, C.c_int(cam_type), _p(camf, _f32p), C.c_int(nx), C.c_int(ny), C.c_int(n),
                              _p(pix, _i32p), _p(smp, _i32p), C.c_int(max_depth), C.c_uint64(seed), _p(rad, _f64p),
                              _p(nr, _i32p), _p(term, _i32p), C.c_int(log_bounces),
                              None if log is None else log.ctypes.data_as(C.c_void_p), C.c_int(n_threads))
        return rad, nr, term, log

    def shade_batch(self, origins, dirs, times, hit_id, ball, u01):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        n = o.shape[0]
        tm = np.ascontiguousarray(times, np.float32)
        hid = np.ascontiguousarray(hit_id, np.int32)
        b = np.ascontiguousarray(ball, np.float32).reshape(-1, 3)
        u = np.ascontiguousarray(u01, np.float32)
        oo = np.zeros((n, 3)); od = np.zeros((n, 3)); oa = np.zeros((n, 3)); oe = np.zeros((n, 3))
        fl = np.zeros(n, np.int32)
        ot = np.zeros(n, np.float64)
        lib().orc_shade_batch(self.h, C.c_int(n), _p(o, _f32p), _p(d, _f32p), _p(tm, _f32p), _p(hid, _i32p),
                              _p(b, _f32p), _p(u, _f32p), _p(oo, _f64p), _p(od, _f64p), _p(oa, _f64p),
                              _p(oe, _f64p), _p(fl, _i32p), _p(ot, _f64p))
        return dict(origin=oo, dir=od, atten=oa, emitted=oe, flags=fl, t=ot)

    def tex_sample(self, tex, u, v, p):
        out = (C.c_double * 3)()
        lib().orc_tex_sample(self.h, C.c_int(tex), C.c_double(u), C.c_double(v), _d3(p), out)
        return np.array(list(out))

    # core.clj:43-57 sample loop: returns (sum_rgb float64 [ny, nx, 3] with j = 0 bottom, counters dict)
    def render_accumulate(self, cam_type, cam, nx, ny, s_begin, s_count, max_depth=50, seed=1,
                          row_offset=0, row_stride=1, n_threads=0, sum_rgb=None, replay=False, use_bvh=False):
        """replay: draw from the Philox counters the CUDA kernels use (same paths as the GPU render of that seed);
        use_bvh: traverse the reference-style BVH (build_bvh) — `sphere_tests` then counts leaf tests."""
        camf = np.ascontiguousarray(cam, np.float32)
        assert camf.shape == (24,)
        if sum_rgb is None:
            sum_rgb = np.zeros((ny, nx, 3), np.float64)
        ctr = np.zeros(8, np.uint64)
        lib().orc_render_accumulate_ex(self.h, C.c_int(cam_type), _p(camf, _f32p), C.c_int(nx), C.c_int(ny),
                                       C.c_int(s_begin), C.c_int(s_count), C.c_int(row_offset), C.c_int(row_stride),
                                       C.c_int(max_depth), C.c_uint64(seed), _p(sum_rgb, _f64p), _p(ctr, _u64p),
                                       C.c_int(n_threads), C.c_int((1 if replay else 0) | (2 if use_bvh else 0)))
        names = ["rays", "sphere_tests", "samples", "term_light", "term_absorb", "term_depth", "term_miss", "aabb_tests"]
        return sum_rgb, {k: int(v) for k, v in zip(names, ctr)}

    def perlin_noise(self, p):
        f = lib().orc_perlin_noise
        f.restype = C.c_double
        return float(f(self.h, _d3(p)))

    def perlin_turbulence(self, p, depth):
        f = lib().orc_perlin_turbulence
        f.restype = C.c_double
        return float(f(self.h, _d3(p), C.c_int(depth)))


def resolve(sum_rgb, nr):
    """core.clj:52-57 + the y flip of core.clj:105: float sums -> uint8 [ny, nx, 3], row 0 = top."""
    s = np.ascontiguousarray(sum_rgb, np.float64)
    ny, nx, _ = s.shape
    out = np.empty((ny, nx, 3), np.uint8)
    lib().orc_resolve(_p(s, _f64p), C.c_int(nx), C.c_int(ny), C.c_int(nr), _p(out, _u8p))
    return out


def point_at_parameter(o, d, t):
    out = (C.c_double * 3)()
    lib().orc_point_at_parameter(_d3(o), _d3(d), C.c_double(t), out)
    return np.array(list(out))


def center_at_time(c0, t0, c1, t1, t):
    out = (C.c_double * 3)()
    lib().orc_center_at_time(_d3(c0), C.c_double(t0), _d3(c1), C.c_double(t1), C.c_double(t), out)
    return np.array(list(out))


def sphere_hit(center, radius, o, d, t_min, t_max, time=0.0, center1=None, t0=0.0, t1=1.0, uv=False):
    """Sphere / UVSphere / MovingSphere hit? on doubles; returns None or dict(t, p, normal, uv)."""
    flags = (1 if uv else 0) | (2 if center1 is not None else 0)
    t = C.c_double()
    p = (C.c_double * 3)(); n = (C.c_double * 3)(); uvo = (C.c_double * 2)()
    ok = lib().orc_sphere_hit(_d3(center), _d3(center1) if center1 is not None else None, C.c_double(t0),
                              C.c_double(t1), C.c_double(radius), C.c_uint32(flags), _d3(o), _d3(d),
                              C.c_double(time), C.c_double(t_min), C.c_double(t_max), C.byref(t), p, n, uvo)
    if not ok:
        return None
    return dict(t=t.value, p=np.array(list(p)), normal=np.array(list(n)), uv=np.array(list(uvo)))


def get_sphere_uv(n):
    uv = (C.c_double * 2)()
    lib().orc_get_sphere_uv(_d3(n), uv)
    return np.array(list(uv))


def reflect(v, n):
    out = (C.c_double * 3)()
    lib().orc_reflect(_d3(v), _d3(n), out)
    return np.array(list(out))


def refract(v, n, ni_over_nt):
    out = (C.c_double * 3)()
    ok = lib().orc_refract(_d3(v), _d3(n), C.c_double(ni_over_nt), out)
    return np.array(list(out)) if ok else None


def schlick(cosine, ri):
    return float(lib().orc_schlick(C.c_double(cosine), C.c_double(ri)))


def thin_lens_camera(lookfrom, lookat, vup, vfov, aspect, aperture, focus_dist, t0, t1):
    out = (C.c_double * 24)()
    lib().orc_thin_lens_camera(_d3(lookfrom), _d3(lookat), _d3(vup), C.c_double(vfov), C.c_double(aspect),
                               C.c_double(aperture), C.c_double(focus_dist), C.c_double(t0), C.c_double(t1), out)
    return np.array(list(out))


def pinhole_camera(lookfrom, lookat, vup, vfov, aspect):
    out = (C.c_double * 24)()
    lib().orc_pinhole_camera(_d3(lookfrom), _d3(lookat), _d3(vup), C.c_double(vfov), C.c_double(aspect), out)
    return np.array(list(out))


def get_ray(cam_type, cam, s, t, disk=(0.0, 0.0), time_u=0.0):
    camf = np.ascontiguousarray(cam, np.float32)
    o = (C.c_double * 3)(); d = (C.c_double * 3)(); tm = C.c_double()
    lib().orc_get_ray(C.c_int(cam_type), _p(camf, _f32p), C.c_double(s), C.c_double(t), C.c_double(disk[0]),
                      C.c_double(disk[1]), C.c_double(time_u), o, d, C.byref(tm))
    return np.array(list(o)), np.array(list(d)), tm.value


def aabb_hit(vmin, vmax, o, d, t_min, t_max):
    """AABB.hit? (hitable.clj:36-48)."""
    return bool(lib().orc_aabb_hit(_d3(vmin), _d3(vmax), _d3(o), _d3(d), C.c_double(t_min), C.c_double(t_max)))


def surrounding_bbox(a, b):
    """make-surrounding-bbox (hitable.clj:87-92); boxes as (vmin xyz, vmax xyz)."""
    out = (C.c_double * 6)()
    lib().orc_surrounding_bbox((C.c_double * 6)(*[float(x) for x in a]), (C.c_double * 6)(*[float(x) for x in b]), out)
    return np.array(list(out))


def sample_ball(n, seed=1, replay=False):
    """n points of rand-in-unit-sphere: the reference's rejection loop (util.clj:43-52), or (replay) the closed-form
    map on the Philox counters the CUDA kernels use."""
    out = np.zeros((n, 3), np.float64)
    lib().orc_sample_ball(C.c_int(n), C.c_uint64(seed), C.c_int(1 if replay else 0), _p(out, _f64p))
    return out


def sample_disk(n, seed=1, replay=False):
    out = np.zeros((n, 2), np.float64)
    lib().orc_sample_disk(C.c_int(n), C.c_uint64(seed), C.c_int(1 if replay else 0), _p(out, _f64p))
    return out


def max_threads():
    return int(lib().orc_max_threads())
