"""raytrace_clj_b200 — B200-native drop-in for the per-pixel path-tracing loop of
gonewest818/raytrace-clj (render block core.clj:99-108 and everything it calls per ray).

Package layout (only what the path needs):
  csrc/      hand-written sm_100a CUDA kernels + the C ABI (include/raytrace_b200.h)
  native.py  marshaller (world -> SoA buffers) + ctypes binding of libraytrace_b200.so
  util/camera/hitable/shader/texture/scene/core.py
             host-side mirror of the reference's front end (records, scene builders, CLI)
The importable name uses an underscore (a hyphen is not a valid Python identifier).
"""
from . import camera, core, hitable, native, parallel, ppm, scene, shader, texture, util  # noqa: F401

__all__ = ["camera", "core", "hitable", "native", "parallel", "ppm", "scene", "shader", "texture", "util"]
