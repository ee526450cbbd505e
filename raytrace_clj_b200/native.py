"""Marshaller + ctypes binding of libraytrace_b200.so (the C ABI in include/raytrace_b200.h).

This is the Python stand-in for the JVM-side ``native.clj`` a maintainer would add to the
reference: walk ``world`` (bvh-node tree, hitable.clj:97-123) to its leaves,
de-duplicate by identity (a 1-element node stores the same object twice, hitable.clj:113-114),
and emit structure-of-arrays float32 buffers; any record outside the accelerated path is
rejected — there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import camera as cam
from . import hitable as hit
from . import shader as shad
from . import texture as tex

RT_SPHERE_UV, RT_SPHERE_MOVING = 1, 2
RT_MAT_LAMBERTIAN, RT_MAT_METAL, RT_MAT_DIELECTRIC, RT_MAT_DIFFUSE_LIGHT = 0, 1, 2, 3
RT_TEX_CONSTANT, RT_TEX_UV_GRADIENT, RT_TEX_CHECKERBOARD = 0, 1, 2
RT_CAM_PINHOLE, RT_CAM_THIN_LENS = 0, 1
RT_VARIANT_MEGAKERNEL, RT_VARIANT_WAVEFRONT = 0, 1
RT_CTR_COUNT = 16
COUNTER_NAMES = ["rays", "sphere_tests", "samples", "term_light", "term_absorb", "term_depth", "term_miss",
                 "kernel_ns", "candidates", "kernel_launches", "cull_ns", "refine_ns", "tiebreak_ns", "shade_ns"]


class UnsupportedSceneError(ValueError):
    """A record type outside the accelerated hot path (RT_ERR_UNSUPPORTED on the JVM side)."""


class NativeError(RuntimeError):
    pass


@dataclass
class FlatScene:
    """The marshalled scene: exactly the buffers of ``rt_scene_desc``."""
    center0_r: np.ndarray      # [n,4] f32
    center1: np.ndarray        # [n,4] f32
    t0t1: np.ndarray           # [n,2] f32
    sphere_flags: np.ndarray   # [n]   u32
    material_id: np.ndarray    # [n]   i32
    mat_type: np.ndarray       # [m]   i32
    mat_param: np.ndarray      # [m]   f32
    mat_tex: np.ndarray        # [m]   i32
    tex_type: np.ndarray       # [t]   i32
    tex_params: np.ndarray     # [t,12] f32
    tex_children: np.ndarray   # [t,2] i32

    @property
    def n_spheres(self):
        return int(self.center0_r.shape[0])


def flatten_world(world):
    """Leaves of the container tree, left to right, de-duplicated by identity."""
    out, seen = [], set()

    def walk(h):
        if isinstance(h, hit.BvhNode):
            walk(h.left)
            walk(h.right)
        elif isinstance(h, hit.Hitlist):
            for it in h.items:
                walk(it)
        elif isinstance(h, (hit.Sphere, hit.MovingSphere)):
            if id(h) not in seen:
                seen.add(id(h))
                out.append(h)
        else:
            raise UnsupportedSceneError(
                f"{type(h).__name__} is outside the accelerated path (spheres only); no CPU fallback")

    # iterative-safe for deep trees
    import sys
    old = sys.getrecursionlimit()
    sys.setrecursionlimit(max(old, 10000))
    try:
        walk(world)
    finally:
        sys.setrecursionlimit(old)
    return out


def marshal_world(world) -> FlatScene:
    leaves = flatten_world(world)
    n = len(leaves)
    if n == 0:
        raise UnsupportedSceneError("empty world")
    c0r = np.zeros((n, 4), np.float32)
    c1 = np.zeros((n, 4), np.float32)
    t0t1 = np.zeros((n, 2), np.float32)
    t0t1[:, 1] = 1.0
    flags = np.zeros(n, np.uint32)
    mat_id = np.zeros(n, np.int32)
    mats, mat_index = [], {}
    texs, tex_index = [], {}

    def add_tex(t):
        if id(t) in tex_index:
            return tex_index[id(t)]
        p = np.zeros(12, np.float32)
        ch = [-1, -1]
        if isinstance(t, tex.Constant):
            ty = RT_TEX_CONSTANT
            p[0:3] = t.color
        elif isinstance(t, tex.UVGradient):
            ty = RT_TEX_UV_GRADIENT
            p[0:3], p[3:6], p[6:9], p[9:12] = t.co, t.cu, t.cv, t.cuv
        elif isinstance(t, tex.Checkerboard):
            ty = RT_TEX_CHECKERBOARD
            p[0] = t.scale
            ch = [add_tex(t.tex0), add_tex(t.tex1)]
        else:
            raise UnsupportedSceneError(f"texture {type(t).__name__} is outside the accelerated path")
        tex_index[id(t)] = len(texs)
        texs.append((ty, p, ch))
        return tex_index[id(t)]

    def add_mat(m):
        if id(m) in mat_index:
            return mat_index[id(m)]
        if isinstance(m, shad.Lambertian):
            rec = (RT_MAT_LAMBERTIAN, 0.0, add_tex(m.albedo))
        elif isinstance(m, shad.Metal):
            rec = (RT_MAT_METAL, m.fuzz, add_tex(m.albedo))
        elif isinstance(m, shad.Dielectric):
            rec = (RT_MAT_DIELECTRIC, m.ri, -1)
        elif isinstance(m, shad.DiffuseLight):
            rec = (RT_MAT_DIFFUSE_LIGHT, 0.0, add_tex(m.tex))
        else:
            raise UnsupportedSceneError(f"material {type(m).__name__} is outside the accelerated path")
        mat_index[id(m)] = len(mats)
        mats.append(rec)
        return mat_index[id(m)]

    for i, s in enumerate(leaves):
        if isinstance(s, hit.MovingSphere):
            c0r[i, :3], c0r[i, 3] = s.center0, s.radius
            c1[i, :3] = s.center1
            t0t1[i] = (s.t0, s.t1)
            flags[i] = RT_SPHERE_MOVING
        else:
            c0r[i, :3], c0r[i, 3] = s.center, s.radius
            c1[i, :3] = s.center
            flags[i] = RT_SPHERE_UV if isinstance(s, hit.UVSphere) else 0
        mat_id[i] = add_mat(s.material)

    return FlatScene(
        c0r, c1, t0t1, flags, mat_id,
        np.array([m[0] for m in mats], np.int32), np.array([m[1] for m in mats], np.float32),
        np.array([m[2] for m in mats], np.int32),
        np.array([t[0] for t in texs], np.int32).reshape(-1),
        np.array([t[1] for t in texs], np.float32).reshape(-1, 12),
        np.array([t[2] for t in texs], np.int32).reshape(-1, 2),
    )


def marshal_camera(camera):
    """(cam_type, float32[24]) from a camera record (camera.clj:8,35)."""
    out = np.zeros(24, np.float32)
    if isinstance(camera, cam.ThinLensCamera):
        for k, v in enumerate([camera.origin, camera.lleft, camera.horiz, camera.vert, camera.u, camera.v, camera.w]):
            out[3 * k:3 * k + 3] = v
        out[21], out[22], out[23] = camera.aperture, camera.t0, camera.t1
        return RT_CAM_THIN_LENS, out
    if isinstance(camera, cam.PinholeCamera):
        for k, v in enumerate([camera.origin, camera.lleft, camera.horiz, camera.vert]):
            out[3 * k:3 * k + 3] = v
        return RT_CAM_PINHOLE, out
    raise UnsupportedSceneError(f"camera {type(camera).__name__} is not supported")


# ---------------------------------------------------------------------------------------------
# ctypes binding
# ---------------------------------------------------------------------------------------------

_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_u8p = C.POINTER(C.c_uint8)


class _SceneDesc(C.Structure):
    _fields_ = [
        ("n_spheres", C.c_int32), ("center0_r", _f32p), ("center1", _f32p), ("t0t1", _f32p),
        ("sphere_flags", _u32p), ("material_id", _i32p),
        ("n_materials", C.c_int32), ("mat_type", _i32p), ("mat_param", _f32p), ("mat_tex", _i32p),
        ("n_textures", C.c_int32), ("tex_type", _i32p), ("tex_params", _f32p), ("tex_children", _i32p),
    ]


LIB_NAME = "libraytrace_b200.so"
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)

# every symbol include/raytrace_b200.h declares
ABI_SYMBOLS = [
    "rt_create", "rt_destroy", "rt_last_error", "rt_abi_version", "rt_set_scene", "rt_set_camera", "rt_render",
    "rt_render_accumulate_device", "rt_resolve_device", "rt_trace_primary", "rt_generate_rays", "rt_shade_batch",
    "rt_measure_fp32_peak", "rt_get_counters", "rt_reset_counters", "rt_set_profile", "rt_device_info", "rt_cull_check",
]

_lib = None


def load_library():
    """Load the CUDA library; fails loudly if it was not built (there is no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(f"{LIB_PATH} not built — run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a); the render path has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    L.rt_last_error.restype = C.c_char_p
    L.rt_last_error.argtypes = [C.c_void_p]
    L.rt_create.argtypes = [C.POINTER(C.c_void_p), _i32p, C.c_int]
    L.rt_destroy.argtypes = [C.c_void_p]
    L.rt_destroy.restype = None
    L.rt_set_scene.argtypes = [C.c_void_p, C.POINTER(_SceneDesc)]
    L.rt_set_camera.argtypes = [C.c_void_p, C.c_int, _f32p]
    L.rt_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, _f32p, _u8p]
    L.rt_render_accumulate_device.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                              C.c_int, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    L.rt_resolve_device.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_int]
    L.rt_trace_primary.argtypes = [C.c_void_p, C.c_int, _f32p, _f32p, _f32p, C.c_double, C.c_double, _f64p, _i32p]
    L.rt_cull_check.argtypes = [C.c_void_p, C.c_int, _f32p, _f32p, _f32p, C.c_double, C.c_double, _u64p]
    L.rt_generate_rays.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _i32p, _i32p, C.c_uint64, _f32p, _f32p,
                                   _f32p, _f32p]
    L.rt_reset_counters.argtypes = [C.c_void_p]
    L.rt_set_profile.argtypes = [C.c_void_p, C.c_int]
    L.rt_shade_batch.argtypes = [C.c_void_p, C.c_int, _f32p, _f32p, _f32p, _i32p, _f64p, _f32p, _f32p, _f32p, _f32p,
                                 _f32p, _f32p, _i32p]
    L.rt_measure_fp32_peak.argtypes = [C.c_void_p, _f64p, _f64p]
    L.rt_get_counters.argtypes = [C.c_void_p, _u64p]
    L.rt_device_info.argtypes = [C.c_void_p, _i32p, _i32p, C.c_char_p]
    _lib = L
    return L


def _p(a, ty):
    return a.ctypes.data_as(ty) if a is not None else None


class Renderer:
    """One ``rt_ctx``: the object the front end holds in place of the claypoole pool (core.clj:100)."""

    def __init__(self, device_ids=None):
        self.L = load_library()
        self.h = C.c_void_p()
        ids = np.asarray(device_ids if device_ids is not None else [0], np.int32)
        rc = self.L.rt_create(C.byref(self.h), _p(ids, _i32p), C.c_int(len(ids)))
        if rc != 0:
            msg = self.L.rt_last_error(None)
            raise NativeError(f"rt_create failed ({rc}): {msg.decode() if msg else '?'}")
        self.flat = None

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.rt_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc, what):
        if rc != 0:
            msg = self.L.rt_last_error(self.h)
            raise NativeError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")

    def set_scene(self, flat: FlatScene):
        self.flat = flat
        k = dict(
            c0=np.ascontiguousarray(flat.center0_r, np.float32), c1=np.ascontiguousarray(flat.center1, np.float32),
            tt=np.ascontiguousarray(flat.t0t1, np.float32), fl=np.ascontiguousarray(flat.sphere_flags, np.uint32),
            mi=np.ascontiguousarray(flat.material_id, np.int32), mt=np.ascontiguousarray(flat.mat_type, np.int32),
            mp=np.ascontiguousarray(flat.mat_param, np.float32), mx=np.ascontiguousarray(flat.mat_tex, np.int32),
            tt2=np.ascontiguousarray(flat.tex_type, np.int32), tp=np.ascontiguousarray(flat.tex_params, np.float32),
            tc=np.ascontiguousarray(flat.tex_children, np.int32))
        d = _SceneDesc(len(k["fl"]), _p(k["c0"], _f32p), _p(k["c1"], _f32p), _p(k["tt"], _f32p), _p(k["fl"], _u32p),
                       _p(k["mi"], _i32p), len(k["mt"]), _p(k["mt"], _i32p), _p(k["mp"], _f32p), _p(k["mx"], _i32p),
                       len(k["tt2"]), _p(k["tt2"], _i32p), _p(k["tp"], _f32p), _p(k["tc"], _i32p))
        self._check(self.L.rt_set_scene(self.h, C.byref(d)), "rt_set_scene")

    def set_camera(self, cam_type, camf):
        c = np.ascontiguousarray(camf, np.float32)
        assert c.shape == (24,)
        self._check(self.L.rt_set_camera(self.h, C.c_int(cam_type), _p(c, _f32p)), "rt_set_camera")

    def render(self, nx, ny, nsamples, max_depth=50, seed=1, variant=RT_VARIANT_WAVEFRONT, linear=True, rgb8=True,
               out_linear=None, out_rgb8=None):
        """rt_render with host buffers.  Returns (linear [ny,nx,3] f32 with j=0 bottom | None, rgb8 [ny,nx,3] | None)."""
        lin = (out_linear if out_linear is not None else np.empty((ny, nx, 3), np.float32)) if linear else None
        img = (out_rgb8 if out_rgb8 is not None else np.empty((ny, nx, 3), np.uint8)) if rgb8 else None
        self._check(self.L.rt_render(self.h, nx, ny, nsamples, max_depth, C.c_uint64(seed), variant, _p(lin, _f32p),
                                     _p(img, _u8p)), "rt_render")
        return lin, img

    def render_accumulate_device(self, nx, ny, sample_begin, sample_count, d_sum_ptr, row_offset=0, row_stride=1,
                                 max_depth=50, seed=1, variant=RT_VARIANT_WAVEFRONT, stream=None, sync=True):
        self._check(self.L.rt_render_accumulate_device(
            self.h, nx, ny, sample_begin, sample_count, row_offset, row_stride, max_depth, C.c_uint64(seed), variant,
            C.c_void_p(d_sum_ptr), C.c_void_p(stream or 0), 1 if sync else 0), "rt_render_accumulate_device")

    def resolve_device(self, nx, ny, nsamples_total, d_sum_ptr, d_rgb8_ptr, stream=None, sync=True):
        self._check(self.L.rt_resolve_device(self.h, nx, ny, nsamples_total, C.c_void_p(d_sum_ptr),
                                             C.c_void_p(d_rgb8_ptr), C.c_void_p(stream or 0), 1 if sync else 0),
                    "rt_resolve_device")

    def trace_primary(self, origins, dirs, times=None, t_min=0.001, t_max=float(np.finfo(np.float32).max)):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        n = o.shape[0]
        tm = None if times is None else np.ascontiguousarray(times, np.float32)
        t = np.empty(n, np.float64)
        ids = np.empty(n, np.int32)
        self._check(self.L.rt_trace_primary(self.h, n, _p(o, _f32p), _p(d, _f32p), _p(tm, _f32p), C.c_double(t_min),
                                            C.c_double(t_max), _p(t, _f64p), _p(ids, _i32p)), "rt_trace_primary")
        return t, ids

    def cull_check(self, origins, dirs, times=None, t_min=0.001, t_max=float(np.finfo(np.float32).max)):
        """(pairs the FP32 cull would lose — must be 0, cull survivors, exact candidates) over rays x listed spheres."""
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        tm = None if times is None else np.ascontiguousarray(times, np.float32)
        out = np.zeros(3, np.uint64)
        self._check(self.L.rt_cull_check(self.h, o.shape[0], _p(o, _f32p), _p(d, _f32p), _p(tm, _f32p), C.c_double(t_min),
                                         C.c_double(t_max), _p(out, _u64p)), "rt_cull_check")
        return int(out[0]), int(out[1]), int(out[2])

    def generate_rays(self, nx, ny, ij, s, seed=1):
        ij = np.ascontiguousarray(ij, np.int32).reshape(-1, 2)
        s = np.ascontiguousarray(s, np.int32)
        n = ij.shape[0]
        o = np.empty((n, 3), np.float32); d = np.empty((n, 3), np.float32); t = np.empty(n, np.float32)
        rnd = np.empty((n, 5), np.float32)
        self._check(self.L.rt_generate_rays(self.h, n, nx, ny, _p(ij, _i32p), _p(s, _i32p), C.c_uint64(seed),
                                            _p(o, _f32p), _p(d, _f32p), _p(t, _f32p), _p(rnd, _f32p)),
                    "rt_generate_rays")
        return o, d, t, rnd

    def shade_batch(self, origins, dirs, times, hit_id, hit_t, ball, u01):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        n = o.shape[0]
        tm = np.ascontiguousarray(times, np.float32)
        hid = np.ascontiguousarray(hit_id, np.int32)
        ht = np.ascontiguousarray(hit_t, np.float64)
        b = np.ascontiguousarray(ball, np.float32).reshape(-1, 3)
        u = np.ascontiguousarray(u01, np.float32)
        oo = np.zeros((n, 3), np.float32); od = np.zeros((n, 3), np.float32)
        oa = np.zeros((n, 3), np.float32); oe = np.zeros((n, 3), np.float32)
        fl = np.zeros(n, np.int32)
        self._check(self.L.rt_shade_batch(self.h, n, _p(o, _f32p), _p(d, _f32p), _p(tm, _f32p), _p(hid, _i32p),
                                          _p(ht, _f64p), _p(b, _f32p), _p(u, _f32p), _p(oo, _f32p), _p(od, _f32p),
                                          _p(oa, _f32p), _p(oe, _f32p), _p(fl, _i32p)), "rt_shade_batch")
        return dict(origin=oo, dir=od, atten=oa, emitted=oe, flags=fl)

    def measure_fp32_peak(self):
        a = C.c_double(); b = C.c_double()
        self._check(self.L.rt_measure_fp32_peak(self.h, C.byref(a), C.byref(b)), "rt_measure_fp32_peak")
        return a.value, b.value

    def counters(self):
        out = np.zeros(RT_CTR_COUNT, np.uint64)
        self._check(self.L.rt_get_counters(self.h, _p(out, _u64p)), "rt_get_counters")
        return {k: int(v) for k, v in zip(COUNTER_NAMES, out)}

    def reset_counters(self):
        self._check(self.L.rt_reset_counters(self.h), "rt_reset_counters")

    def set_profile(self, on=True):
        self._check(self.L.rt_set_profile(self.h, 1 if on else 0), "rt_set_profile")

    def device_info(self):
        sm = C.c_int32(); clk = C.c_int32(); name = C.create_string_buffer(64)
        self._check(self.L.rt_device_info(self.h, C.byref(sm), C.byref(clk), name), "rt_device_info")
        return dict(sm_count=sm.value, clock_khz=clk.value, name=name.value.decode())
