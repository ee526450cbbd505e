// rt_device.cuh — per-ray device code shared by every kernel of libraytrace_b200.so (sm_100a).
//
// What runs here replaces, per ray, the reference's protocol calls from `color`
// (core.clj:17-41): hit? (hitable.clj), scatter / emitted (shader.clj), sample (texture.clj),
// get-ray (camera.clj) and the vec3 helpers (util.clj).
//
// Numerics (DESIGN.md "Precision"):
//   * the brute-force loop over ALL spheres is an FP32 conservative cull (17 flop per
//     ray-sphere test) with a slightly inflated radius and deflated a = d.d, so it can only
//     produce false positives;
//   * the few survivors per ray are re-evaluated in FP64 with exactly the reference's formula
//     and operation order (no FMA contraction), so t and the winning sphere are bit-identical
//     to the double-precision oracle on the same float32 inputs;
//   * shading (hit point, normal, scatter direction, textures) is FP32.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <float.h>
#include <math_constants.h>

namespace rt {

// ------------------------------------------------------------------------------------------
// constants shared with the host
// ------------------------------------------------------------------------------------------
constexpr unsigned SPH_UV = 1u, SPH_MOVING = 2u;
constexpr int MAT_LAMBERTIAN = 0, MAT_METAL = 1, MAT_DIELECTRIC = 2, MAT_DIFFUSE_LIGHT = 3, MAT_ISOTROPIC = 4;
constexpr int TEX_CONSTANT = 0, TEX_UV_GRADIENT = 1, TEX_CHECKERBOARD = 2, TEX_PERLIN_NOISE = 3, TEX_PERLIN_TURB = 4,
              TEX_MARBLE = 5, TEX_FLIP_U = 6, TEX_FLIP_V = 7, TEX_IMAGE_MAP = 8;
constexpr int PRIM_SPHERE = 0, PRIM_RECT_XY = 1, PRIM_RECT_XZ = 2, PRIM_RECT_YZ = 3, PRIM_TRIANGLE = 4, PRIM_MEDIUM = 5;
constexpr int XOP_NONE = 0, XOP_TRANSLATE = 1, XOP_ROTATE_Y = 2, XOP_FLIP = 3;
constexpr int XFORM_MAX_OPS = 4;
constexpr int CAM_PINHOLE = 0, CAM_THIN_LENS = 1;

// cull tolerances: the cull direction is scaled up by sqrt(1 + CULL_EPS); see Culler (rt_kernels.cuh)
constexpr float CULL_EPS = 7.62939453125e-6f;  // 2^-17

enum TermReason { TERM_NONE = 0, TERM_LIGHT = 1, TERM_ABSORB = 2, TERM_DEPTH = 3, TERM_MISS = 4 };

// device counter slots (subset of RT_CTR_* that the kernels write)
enum { DC_RAYS = 0, DC_SAMPLES, DC_TERM_LIGHT, DC_TERM_ABSORB, DC_TERM_DEPTH, DC_TERM_MISS, DC_CANDIDATES, DC_DIRECT, DC_BVH_NODES, DC_COUNT = 10 };

// Scene in HBM, every per-sphere array in "cull order": first the n_list spheres the FP32 cull runs over,
// then the n - n_list "direct" spheres that bypass it (enclosing spheres such as a sky dome or a ground
// sphere: the cull would pass them for nearly every ray, so they are tested exactly, once per ray, where the
// closest hit is resolved).  orig_id maps cull order back to the caller's index (ties, reporting).
struct DevScene {
    int n;                    // spheres
    int n_list;               // spheres in the cull list = cull order [0, n_list)
    int n_cull;               // cull records: n_list rounded up to a multiple of 8 (padding never survives)
    const float4* cull_a;     // [n_cull] (-cx, -cy, -cz, W = r2_inflated - c.c); a moving sphere is the bounding
                              //     sphere of its swept volume over the time window the context covers
    const float4* ex_c0r;     // [n] exact centre0 + radius   (float32 as marshalled)
    const float4* ex_c1;      // [n] exact centre1
    const float2* ex_t0t1;    // [n]
    const int* orig_id;       // [n] cull order -> caller's sphere index
    const int* cull_of_orig;  // [n] caller's sphere index -> cull order
    const unsigned* flags;    // [n]
    const int* mat_id;        // [n]
    const int4* shade_rec;    // [n] (material type, texture id, texture type, sphere flags): the material / texture tables
                              //     resolved per sphere at upload, so a hit costs ONE dependent load level instead of four
    const float4* shade_col;  // [n] (material parameter fuzz | ri, then the colour of a CONSTANT texture)
    const int* mat_type;      // [m]
    const float* mat_param;   // [m]
    const int* mat_tex;       // [m]
    const int* tex_type;      // [t]
    const float* tex_params;  // [12 t]
    const int* tex_child;     // [2 t]
    // ---- beyond spheres (rt_scene_ext); generic == 0 for a plain sphere scene: none of the following is touched ----
    int generic;              // 1: some leaf is not a plain sphere, carries a wrapper chain, or a texture / material beyond the
                              //    basic set is present -> the GEN = true instantiations of the kernels (generic refine / hit
                              //    record / textures); a plain scene runs GEN = false kernels that hold none of that code
    int n_total;              // n + boundary primitives of media (indices [n, n_total), caller's order, never hit directly)
    const unsigned* tie_hi;   // [n] high word of the tie-break key: among EXACT ties in t the smallest key wins
                              //     (Hitlist world: first strict leaf, but a later inclusive-range leaf replaces it; bvh-node
                              //     world: the last leaf in flatten order — hitable.clj:17-26 / :99-105)
    const int* prim_type;     // [n_total]
    const float4* prim_q;     // [3 n_total] 12 parameters (rect: a0 b0 a1 b1 k; triangle: v0 v1 v2; medium: density)
    const int2* prim_aux;     // [n_total] medium: (first boundary primitive, count)
    const int* prim_xform;    // [n_total] wrapper chain or -1
    const int* xform_ops;     // [4 x]
    const float4* xform_p;    // [4 x]
    const float4* perlin_vec; // [256] perlin.clj:6-8
    const int* perlin_perm;   // [768] perm-x, perm-y, perm-z
    const int2* image_wh;     // [images]
    const long long* image_off;
    const unsigned char* image_rgb;
    // ---- RT_ACCEL_BVH: flattened BVH over the listed leaves [0, n_list) (the role of hitable.clj:97-123) ----
    const float4* bvh;        // [4 per node]: left child's box lo.xyz / hi.xyz, right child's box lo / hi packed as
                              //   (llo.x llo.y llo.z lhi.x) (lhi.y lhi.z rlo.x rlo.y) (rlo.z rhi.x rhi.y rhi.z) (left, right, -, -);
                              //   a child index < 0 is the leaf ~index (cull order); node 0 is the root
    int bvh_nodes;
    // tensor-core cull (rt_cull_tc.cuh): TF32 feature rows of the listed leaves, tc_tiles tiles of 256 rows x 32 K-slots in
    // the canonical K-major UMMA layout (padding rows never survive); 0 tiles = not built
    const float* cull_tc;
    const int* tc_row_k;          // [tc_tiles * 256] cull index of the leaf in feature row p (a fixed shuffle of the list), -1 = padding
    int tc_tiles;
};

struct DevCamera {
    int type;
    float3 origin, lleft, horiz, vert, u, v, w;
    float lens_radius, t0, t1;
};

// ------------------------------------------------------------------------------------------
// vec3 (util.clj:5-11) in registers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ float3 operator*(float s, float3 a) { return f3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot3(float3 a, float3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
__device__ __forceinline__ float3 normalise3(float3 a) {
    float m = sqrtf(dot3(a, a));
    return (m > 0.f) ? (1.0f / m) * a : a;
}

// ------------------------------------------------------------------------------------------
// Philox4x32-10 counter RNG (stands in for clojure.core/rand; keyed per pixel/sample/bounce)
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint64_t p0 = (uint64_t)M0 * c.x, p1 = (uint64_t)M1 * c.z;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0;
        k.y += W1;
    }
    return c;
}
// uniform float in [0,1) from the top 24 bits
__host__ __device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-8f; }

// counter layout: (pixel index, sample index, (bounce << 16) | block, domain)
constexpr uint32_t RNG_DOMAIN = 0x52544232u;  // "RTB2"
__device__ __forceinline__ uint4 rng_block(uint2 key, uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t blk) {
    return philox4x32_10(make_uint4(pixel, sample, (bounce << 16) | blk, RNG_DOMAIN), key);
}

// util.clj:43-52 rand-in-unit-sphere.  The reference rejects points of [-1,1)^3 until dot < 1, i.e. it draws
// uniformly from the open unit ball; the JVM stream cannot be reproduced anyway, so the same distribution
// is drawn in closed form from ONE Philox block (no divergent retry loop): radius cbrt(u), uniform direction.
__device__ __forceinline__ float3 rand_in_unit_sphere(uint2 key, uint32_t pixel, uint32_t sample, uint32_t bounce,
                                                      uint32_t blk) {
    uint4 r = rng_block(key, pixel, sample, bounce, blk);
    float rad = cbrtf(u01(r.x));
    float z = fmaf(-2.0f, u01(r.y), 1.0f);
    float s = sqrtf(fmaxf(0.f, fmaf(-z, z, 1.0f))) * rad;
    float sn, cs;
    sincospif(2.0f * u01(r.z), &sn, &cs);
    return f3(s * cs, s * sn, z * rad);
}
// util.clj:32-41 rand-in-unit-disk: uniform in the open unit disk, closed form
__device__ __forceinline__ float2 rand_in_unit_disk(uint2 key, uint32_t pixel, uint32_t sample, uint32_t blk) {
    uint4 r = rng_block(key, pixel, sample, 0u, blk);
    float rad = sqrtf(u01(r.x));
    float sn, cs;
    sincospif(2.0f * u01(r.y), &sn, &cs);
    return make_float2(rad * cs, rad * sn);
}

// ------------------------------------------------------------------------------------------
// camera.clj:8-16 / 35-48 get-ray for (pixel i, j, sample s).  core.clj:49-50: u drawn first.
// rnd (optional) receives the raw uniforms (ru, rv, disk.x, disk.y, time_u) for the tests.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void generate_ray(const DevCamera& cam, int nx, int ny, int i, int j, uint32_t pixel,
                                             uint32_t sample, uint2 key, float3& o, float3& d, float& time,
                                             float* rnd) {
    uint4 r0 = rng_block(key, pixel, sample, 0u, 0u);
    float ru = u01(r0.x), rv = u01(r0.y), rt_ = u01(r0.z);
    float s = ((float)i + ru) / (float)nx;
    float t = ((float)j + rv) / (float)ny;
    float3 dir = cam.lleft + s * cam.horiz + t * cam.vert - cam.origin;
    float2 disk = make_float2(0.f, 0.f);
    if (cam.type == CAM_THIN_LENS) {
        float3 offset = f3(0.f, 0.f, 0.f);
        if (cam.lens_radius != 0.f) {  // aperture 0 (every reference scene): the disk draw has no effect
            disk = rand_in_unit_disk(key, pixel, sample, 1u);
            offset = (cam.lens_radius * disk.x) * cam.u + (cam.lens_radius * disk.y) * cam.v;
        }
        o = cam.origin + offset;
        d = dir - offset;
        time = fmaf(cam.t1 - cam.t0, rt_, cam.t0);
    } else {
        o = cam.origin;
        d = dir;
        time = 0.f;
    }
    if (rnd) {
        rnd[0] = ru; rnd[1] = rv; rnd[2] = disk.x; rnd[3] = disk.y; rnd[4] = rt_;
    }
}

// ------------------------------------------------------------------------------------------
// FP64 refine of one cull survivor: Sphere/UVSphere/MovingSphere.hit? (hitable.clj:143-168,
// 182-207, 226-251) with the reference's operation order, no FMA contraction.
// Keeps the closest t; exact ties are resolved by DevScene::tie_hi (hitable.clj:17-26 / :99-105).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dadd_rn(a, -b); }
__device__ __forceinline__ double ddot(double ax, double ay, double az, double bx, double by, double bz) {
    return dadd(dadd(dmul(ax, bx), dmul(ay, by)), dmul(az, bz));
}

// the quadratic of hitable.clj:185-206 given oc = o - centre, d, r: near root then far root, strict range
__device__ __forceinline__ double sphere_roots(double ocx, double ocy, double ocz, double ddx, double ddy, double ddz, double r,
                                               double tmin, double tmax) {
    double a = ddot(ddx, ddy, ddz, ddx, ddy, ddz);
    double b = dmul(2.0, ddot(ocx, ocy, ocz, ddx, ddy, ddz));
    double c = dsub(ddot(ocx, ocy, ocz, ocx, ocy, ocz), dmul(r, r));
    double disc = dsub(dmul(b, b), dmul(dmul(4.0, a), c));
    if (disc >= 0.0) {
        double sq = __dsqrt_rn(disc);
        double two_a = dmul(2.0, a);
        if (tmin >= 0.0) {
            // tmin >= 0 and 2a >= 0: a root whose numerator is <= 0 cannot pass t > tmin, so its division (the costliest
            // part of the test) is skipped — the accepted t is still computed exactly as the reference does
            double num = dsub(-b, sq);
            if (num > 0.0) {
                double t = __ddiv_rn(num, two_a);
                if (t > tmin && t < tmax) return t;
            }
            num = dadd(-b, sq);
            if (num > 0.0) {
                double t = __ddiv_rn(num, two_a);
                if (t > tmin && t < tmax) return t;
            }
        } else {   // a medium's boundary is searched from t = -FLT_MAX (hitable.clj:519)
            double t = __ddiv_rn(dsub(-b, sq), two_a);
            if (t > tmin && t < tmax) return t;
            t = __ddiv_rn(dadd(-b, sq), two_a);
            if (t > tmin && t < tmax) return t;
        }
    }
    return CUDART_INF;
}

// Everything is passed and returned BY VALUE: a reference into a noinline function would force
// the caller's ray registers into local memory.  Returns the accepted t, or +inf.
__device__ __noinline__ double refine_candidate(const float4* __restrict__ ex_c0r, const float4* __restrict__ ex_c1,
                                                const float2* __restrict__ ex_t0t1, const unsigned* __restrict__ flags,
                                                int k, float ox, float oy, float oz, float dx, float dy, float dz,
                                                float time, double tmin, double tmax) {
    float4 c0r = __ldg(&ex_c0r[k]);
    double cx = c0r.x, cy = c0r.y, cz = c0r.z, r = c0r.w;
    if (__ldg(&flags[k]) & SPH_MOVING) {
        float4 c1 = __ldg(&ex_c1[k]);
        float2 tt = __ldg(&ex_t0t1[k]);
        // hitable.clj:219-222: lerp(c0, c1, (time - t0) / (t1 - t0)) = c0*(1-f) + c1*f
        double f = __ddiv_rn(dsub((double)time, (double)tt.x), dsub((double)tt.y, (double)tt.x));
        double g = dsub(1.0, f);
        cx = dadd(dmul(g, cx), dmul(f, (double)c1.x));
        cy = dadd(dmul(g, cy), dmul(f, (double)c1.y));
        cz = dadd(dmul(g, cz), dmul(f, (double)c1.z));
    }
    return sphere_roots(dsub((double)ox, cx), dsub((double)oy, cy), dsub((double)oz, cz), dx, dy, dz, r, tmin, tmax);
}

// ------------------------------------------------------------------------------------------
// Generic leaves (DevScene::generic): rectangles, triangles, media and wrapper chains, FP64, reference operation order
// ------------------------------------------------------------------------------------------
struct DRay {
    double o[3], d[3];
};

// a leaf's wrapper chain applied to the ray on the way in: Translate (hitable.clj:394) and RotateY (:423-428)
__device__ __forceinline__ void xform_ray(const DevScene* sc, int xf, DRay& r) {
    for (int q = 0; q < XFORM_MAX_OPS; ++q) {
        const int op = __ldg(&sc->xform_ops[XFORM_MAX_OPS * xf + q]);
        if (op == XOP_NONE) break;
        const float4 p = __ldg(&sc->xform_p[XFORM_MAX_OPS * xf + q]);
        if (op == XOP_TRANSLATE) {
            r.o[0] = dsub(r.o[0], (double)p.x); r.o[1] = dsub(r.o[1], (double)p.y); r.o[2] = dsub(r.o[2], (double)p.z);
        } else if (op == XOP_ROTATE_Y) {
            const double sn = p.x, cs = p.y;
            const double ox = r.o[0], oz = r.o[2], dx = r.d[0], dz = r.d[2];
            r.o[0] = dsub(dmul(cs, ox), dmul(sn, oz)); r.o[2] = dadd(dmul(sn, ox), dmul(cs, oz));
            r.d[0] = dsub(dmul(cs, dx), dmul(sn, dz)); r.d[2] = dadd(dmul(sn, dx), dmul(cs, dz));
        }
    }
}

// hit? of one non-medium leaf (through its wrappers): the accepted t or +inf.  Spheres: strict range (hitable.clj:197);
// rectangles / triangles: inclusive (hitable.clj:283, :567).
__device__ __noinline__ double leaf_t(const DevScene* sc, int k, DRay r, float time, double tmin, double tmax) {
    const int xf = __ldg(&sc->prim_xform[k]);
    if (xf >= 0) xform_ray(sc, xf, r);
    const int type = __ldg(&sc->prim_type[k]);
    if (type == PRIM_SPHERE) {
        float4 c0r = __ldg(&sc->ex_c0r[k]);
        double cx = c0r.x, cy = c0r.y, cz = c0r.z;
        if (__ldg(&sc->flags[k]) & SPH_MOVING) {
            float4 c1 = __ldg(&sc->ex_c1[k]);
            float2 tt = __ldg(&sc->ex_t0t1[k]);
            double f = __ddiv_rn(dsub((double)time, (double)tt.x), dsub((double)tt.y, (double)tt.x));
            double g = dsub(1.0, f);
            cx = dadd(dmul(g, cx), dmul(f, (double)c1.x));
            cy = dadd(dmul(g, cy), dmul(f, (double)c1.y));
            cz = dadd(dmul(g, cz), dmul(f, (double)c1.z));
        }
        return sphere_roots(dsub(r.o[0], cx), dsub(r.o[1], cy), dsub(r.o[2], cz), r.d[0], r.d[1], r.d[2], (double)c0r.w, tmin, tmax);
    }
    const float4 q0 = __ldg(&sc->prim_q[3 * k]), q1 = __ldg(&sc->prim_q[3 * k + 1]);
    if (type == PRIM_TRIANGLE) {   // hitable.clj:551-575
        const float4 q2 = __ldg(&sc->prim_q[3 * k + 2]);
        const double v0x = q0.x, v0y = q0.y, v0z = q0.z;
        const double e1x = dsub((double)q0.w, v0x), e1y = dsub((double)q1.x, v0y), e1z = dsub((double)q1.y, v0z);   // v0v1
        const double e2x = dsub((double)q1.z, v0x), e2y = dsub((double)q1.w, v0y), e2z = dsub((double)q2.x, v0z);   // v0v2
        const double px = dsub(dmul(r.d[1], e2z), dmul(r.d[2], e2y)), py = dsub(dmul(r.d[2], e2x), dmul(r.d[0], e2z)),
                     pz = dsub(dmul(r.d[0], e2y), dmul(r.d[1], e2x));                                               // d x v0v2
        const double det = ddot(e1x, e1y, e1z, px, py, pz);
        if (!(det > 0.00000001)) return CUDART_INF;
        const double inv_det = __ddiv_rn(1.0, det);
        const double tx = dsub(r.o[0], v0x), ty = dsub(r.o[1], v0y), tz = dsub(r.o[2], v0z);
        const double u = dmul(ddot(tx, ty, tz, px, py, pz), inv_det);
        if (!(u > 0.0 && u <= 1.0)) return CUDART_INF;
        const double qx = dsub(dmul(ty, e1z), dmul(tz, e1y)), qy = dsub(dmul(tz, e1x), dmul(tx, e1z)),
                     qz = dsub(dmul(tx, e1y), dmul(ty, e1x));                                                       // tvec x v0v1
        const double v = dmul(ddot(r.d[0], r.d[1], r.d[2], qx, qy, qz), inv_det);
        if (!(v > 0.0 && dadd(u, v) <= 1.0)) return CUDART_INF;
        const double t = dmul(ddot(e2x, e2y, e2z, qx, qy, qz), inv_det);
        return (t >= tmin && t <= tmax) ? t : CUDART_INF;
    }
    // rectangles, hitable.clj:272-293 / 303-324 / 334-355: q = a0 b0 a1 b1 k
    const int axis = type == PRIM_RECT_XY ? 2 : (type == PRIM_RECT_XZ ? 1 : 0);
    const int A = type == PRIM_RECT_YZ ? 1 : 0, B = type == PRIM_RECT_XY ? 1 : 2;
    const double t = __ddiv_rn(dsub((double)q1.x, r.o[axis]), r.d[axis]);
    if (t >= tmin && t <= tmax) {
        const double a = dadd(r.o[A], dmul(t, r.d[A])), b = dadd(r.o[B], dmul(t, r.d[B]));
        if (a >= (double)q0.x && a <= (double)q0.z && b >= (double)q0.y && b <= (double)q0.w) return t;
    }
    return CUDART_INF;
}

// Hitlist.hit? over a medium's boundary leaves (hitable.clj:15-26): shrinking t-max; a sphere replaces the running hit
// only when strictly closer, a rectangle also at equality — each leaf's own range test does exactly that
__device__ __forceinline__ double boundary_t(const DevScene* sc, int first, int count, const DRay& r, float time, double tmin,
                                             double tmax) {
    double closest = tmax;
    bool any = false;
    for (int i = first; i < first + count; ++i) {
        const double t = leaf_t(sc, i, r, time, tmin, closest);
        if (t < CUDART_INF) { closest = t; any = true; }
    }
    return any ? closest : CUDART_INF;
}

// hit? of world leaf k of a generic scene.  medium_u: the `rand` a ConstantMedium draws inside hit? (hitable.clj:529).
__device__ __noinline__ double refine_generic(const DevScene* sc, int k, float ox, float oy, float oz, float dx, float dy, float dz,
                                              float time, double tmin, double tmax, float medium_u) {
    DRay r;
    r.o[0] = ox; r.o[1] = oy; r.o[2] = oz; r.d[0] = dx; r.d[1] = dy; r.d[2] = dz;
    if (__ldg(&sc->prim_type[k]) != PRIM_MEDIUM) return leaf_t(sc, k, r, time, tmin, tmax);
    // hitable.clj:516-541
    const int xf = __ldg(&sc->prim_xform[k]);
    if (xf >= 0) xform_ray(sc, xf, r);
    const int2 aux = __ldg(&sc->prim_aux[k]);
    const double FM = (double)FLT_MAX;
    double t1 = boundary_t(sc, aux.x, aux.y, r, time, -FM, FM);
    if (!(t1 < CUDART_INF)) return CUDART_INF;
    double t2 = boundary_t(sc, aux.x, aux.y, r, time, dadd(t1, 0.0001), FM);
    if (!(t2 < CUDART_INF)) return CUDART_INF;
    t1 = (t1 < tmin) ? tmin : t1;
    t2 = (t2 > tmax) ? tmax : t2;
    if (t1 < t2) {
        t1 = (t1 < 0.0) ? 0.0 : t1;
        const double mag = __dsqrt_rn(ddot(r.d[0], r.d[1], r.d[2], r.d[0], r.d[1], r.d[2]));
        const double dist_in_boundary = dmul(dsub(t2, t1), mag);
        const double density = (double)__ldg(&sc->prim_q[3 * k]).x;
        const double hit_distance = -__ddiv_rn(log((double)medium_u), density);
        if (hit_distance < dist_in_boundary) return dadd(t1, __ddiv_rn(hit_distance, mag));
    }
    return CUDART_INF;
}
// the uniform a medium at caller index `orig` draws for the ray of (pixel, sample, bounce): block 16 + orig
__device__ __forceinline__ float medium_uniform(uint2 key, uint32_t pixel, uint32_t sample, uint32_t bounce, int orig) {
    return u01(rng_block(key, pixel, sample, bounce, 16u + (uint32_t)orig).x);
}

// ------------------------------------------------------------------------------------------
// texture.clj:14-138 sample (children always have a smaller id: validated on upload)
// ------------------------------------------------------------------------------------------
// perlin.clj:19-50
__device__ __noinline__ float perlin_noise(const DevScene* sc, float px, float py, float pz) {
    const float fi = floorf(px), fj = floorf(py), fk = floorf(pz);
    const int i = (int)fi, j = (int)fj, k = (int)fk;
    const float u = px - fi, v = py - fj, w = pz - fk;
    const float uu = u * u * (3.f - 2.f * u), vv = v * v * (3.f - 2.f * v), ww = w * w * (3.f - 2.f * w);
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int idx = __ldg(&sc->perlin_perm[(i + a) & 255]) ^ __ldg(&sc->perlin_perm[256 + ((j + b) & 255)]) ^
                                __ldg(&sc->perlin_perm[512 + ((k + c) & 255)]);
                const float4 g = __ldg(&sc->perlin_vec[idx]);
                const float wa = a ? uu : 1.f - uu, wb = b ? vv : 1.f - vv, wc = c ? ww : 1.f - ww;
                acc += wa * wb * wc * ((u - a) * g.x + (v - b) * g.y + (w - c) * g.z);
            }
    return acc;
}
// perlin.clj:52-64
__device__ __forceinline__ float perlin_turbulence(const DevScene* sc, float3 p, int depth) {
    float acc = 0.f, w = 1.f;
    for (int i = 0; i < depth; ++i) {
        acc += w * perlin_noise(sc, p.x, p.y, p.z);
        p = 2.0f * p;
        w *= 0.5f;
    }
    return fabsf(acc);
}

// scp = the same scene as a pointer into the kernel's grid-constant parameter space (for the noinline helpers)
template <bool GEN>
__device__ __forceinline__ float3 tex_sample(const DevScene& sc, const DevScene* scp, int id, float u, float v, float3 p) {
    for (int depth = 0; depth < 32 && id >= 0; ++depth) {
        int ty = __ldg(&sc.tex_type[id]);
        const float* P = sc.tex_params + 12 * id;
        if (ty == TEX_CONSTANT) return f3(__ldg(P), __ldg(P + 1), __ldg(P + 2));
        if (ty == TEX_UV_GRADIENT) {
            float3 co = f3(__ldg(P), __ldg(P + 1), __ldg(P + 2)), cu = f3(__ldg(P + 3), __ldg(P + 4), __ldg(P + 5));
            float3 cv = f3(__ldg(P + 6), __ldg(P + 7), __ldg(P + 8)), cuv = f3(__ldg(P + 9), __ldg(P + 10), __ldg(P + 11));
            float3 a = (1.f - u) * cu + u * co;
            float3 b = (1.f - u) * cuv + u * cv;
            return (1.f - v) * b + v * a;
        }
        if (ty == TEX_CHECKERBOARD) {
            float s = __ldg(P);
            float sines = sinf(s * p.x) * sinf(s * p.y) * sinf(s * p.z);
            id = __ldg(&sc.tex_child[2 * id + ((sines < 0.f) ? 0 : 1)]);
            continue;
        }
        if (!GEN) break;
        // the procedural / image textures (texture.clj:60-138; scenes marshalled through rt_set_scene_ex only)
        if (ty == TEX_FLIP_U) { u = 1.0f - u; id = __ldg(&sc.tex_child[2 * id]); continue; }
        if (ty == TEX_FLIP_V) { v = 1.0f - v; id = __ldg(&sc.tex_child[2 * id]); continue; }
        if (ty == TEX_IMAGE_MAP) {
            const int im = (int)__ldg(P);
            const int2 wh = __ldg(&sc.image_wh[im]);
            int i = (int)(u * (float)wh.x), j = (int)(v * (float)wh.y);
            i = min(max(i, 0), wh.x - 1);
            j = min(max(j, 0), wh.y - 1);
            const unsigned char* px = sc.image_rgb + sc.image_off[im] + ((size_t)j * wh.x + i) * 3;
            return f3((float)px[0] / 255.0f, (float)px[1] / 255.0f, (float)px[2] / 255.0f);
        }
        float g;
        const float s = __ldg(P);
        if (ty == TEX_PERLIN_NOISE) g = 0.5f * (1.0f + perlin_noise(scp, s * p.x, s * p.y, s * p.z));
        else if (ty == TEX_PERLIN_TURB) g = 0.5f * (1.0f + perlin_turbulence(scp, s * p, (int)__ldg(P + 1)));
        else g = 0.5f * (1.0f + sinf(s * p.z + 10.0f * perlin_turbulence(scp, p, (int)__ldg(P + 1))));   // marble
        return f3(g, g, g);
    }
    return f3(0.f, 0.f, 0.f);
}

// shader.clj:6-9
__device__ __forceinline__ float3 reflect3(float3 v, float3 n) { return v - (2.0f * dot3(v, n)) * n; }

// shader.clj:69-74
__device__ __forceinline__ float schlick(float cosine, float ri) {
    float r0 = (1.0f - ri) / (1.0f + ri);
    r0 = r0 * r0;
    float m = 1.0f - cosine;
    float m2 = m * m;
    return r0 + (1.0f - r0) * (m2 * m2 * m);
}

// Source of the random inputs of one scatter: the Philox stream of (pixel, sample, bounce), or
// caller-given values (rt_shade_batch).
struct ScatterRng {
    uint2 key;
    uint32_t pixel, sample, bounce;
    const float* ball;  // explicit rand-in-unit-sphere (3 floats) or nullptr
    const float* u;     // explicit rand or nullptr
    // block 1 of the bounce: (ball radius u, ball z u, ball phi u, the Dielectric's rand) — ONE Philox call per scatter
    __device__ __forceinline__ float3 unit_sphere() const {
        return ball ? f3(ball[0], ball[1], ball[2]) : rand_in_unit_sphere(key, pixel, sample, bounce, 1u);
    }
    __device__ __forceinline__ float rand() const { return u ? *u : u01(rng_block(key, pixel, sample, bounce, 1u).w); }
};

// hit record of a generic leaf (hitable.clj:193-201, 284-293, 568-575, 531-540 seen through :381, :398-400, :432-455), FP32.
// Returned by value: (p, u) and (n, v).
struct HitGeom {
    float4 pu, nv;
};
__device__ __noinline__ HitGeom hit_geom_generic(const DevScene* sc, int k, float t, float ox, float oy, float oz, float dx, float dy,
                                                 float dz, float time) {
    float o[3] = {ox, oy, oz}, d[3] = {dx, dy, dz};
    const int xf = __ldg(&sc->prim_xform[k]);
    int nops = 0;
    if (xf >= 0)
        for (; nops < XFORM_MAX_OPS; ++nops) {
            const int op = __ldg(&sc->xform_ops[XFORM_MAX_OPS * xf + nops]);
            if (op == XOP_NONE) break;
            const float4 p = __ldg(&sc->xform_p[XFORM_MAX_OPS * xf + nops]);
            if (op == XOP_TRANSLATE) { o[0] -= p.x; o[1] -= p.y; o[2] -= p.z; }
            else if (op == XOP_ROTATE_Y) {
                const float x = o[0], z = o[2], ex = d[0], ez = d[2];
                o[0] = p.y * x - p.x * z; o[2] = p.x * x + p.y * z;
                d[0] = p.y * ex - p.x * ez; d[2] = p.x * ex + p.y * ez;
            }
        }
    float pt[3] = {t * d[0] + o[0], t * d[1] + o[1], t * d[2] + o[2]};
    float n[3] = {1.f, 0.f, 0.f};
    float u = 0.f, v = 0.f;
    const int type = __ldg(&sc->prim_type[k]);
    if (type == PRIM_SPHERE) {
        const float4 c0r = __ldg(&sc->ex_c0r[k]);
        float3 c = f3(c0r.x, c0r.y, c0r.z);
        const unsigned fl = __ldg(&sc->flags[k]);
        if (fl & SPH_MOVING) {
            const float4 c1 = __ldg(&sc->ex_c1[k]);
            const float2 tt = __ldg(&sc->ex_t0t1[k]);
            const float f = (time - tt.x) / (tt.y - tt.x);
            c = (1.0f - f) * c + f * f3(c1.x, c1.y, c1.z);
        }
        const float3 nn = normalise3(f3(pt[0], pt[1], pt[2]) - c);
        n[0] = nn.x; n[1] = nn.y; n[2] = nn.z;
        if (fl & SPH_UV) {
            const float PI = 3.14159265358979323846f;
            u = 1.0f - (atan2f(nn.z, nn.x) + PI) / (2.0f * PI);
            v = (asinf(fminf(1.0f, fmaxf(-1.0f, nn.y))) + PI / 2.0f) / PI;
        }
    } else if (type == PRIM_TRIANGLE) {
        const float4 q0 = __ldg(&sc->prim_q[3 * k]), q1 = __ldg(&sc->prim_q[3 * k + 1]), q2 = __ldg(&sc->prim_q[3 * k + 2]);
        const float3 v0 = f3(q0.x, q0.y, q0.z), e1 = f3(q0.w, q1.x, q1.y) - v0, e2 = f3(q1.z, q1.w, q2.x) - v0;
        const float3 dd = f3(d[0], d[1], d[2]), tv = f3(o[0], o[1], o[2]) - v0;
        const float3 pv = f3(dd.y * e2.z - dd.z * e2.y, dd.z * e2.x - dd.x * e2.z, dd.x * e2.y - dd.y * e2.x);
        const float inv_det = 1.0f / dot3(e1, pv);
        const float3 qv = f3(tv.y * e1.z - tv.z * e1.y, tv.z * e1.x - tv.x * e1.z, tv.x * e1.y - tv.y * e1.x);
        u = dot3(tv, pv) * inv_det;
        v = dot3(dd, qv) * inv_det;
        n[0] = e1.y * e2.z - e1.z * e2.y; n[1] = e1.z * e2.x - e1.x * e2.z; n[2] = e1.x * e2.y - e1.y * e2.x;
    } else if (type != PRIM_MEDIUM) {
        const float4 q0 = __ldg(&sc->prim_q[3 * k]);
        const int axis = type == PRIM_RECT_XY ? 2 : (type == PRIM_RECT_XZ ? 1 : 0);
        const int A = type == PRIM_RECT_YZ ? 1 : 0, B = type == PRIM_RECT_XY ? 1 : 2;
        u = (pt[A] - q0.x) / (q0.z - q0.x);
        v = (pt[B] - q0.y) / (q0.w - q0.y);
        n[0] = axis == 0 ? 1.f : 0.f; n[1] = axis == 1 ? 1.f : 0.f; n[2] = axis == 2 ? 1.f : 0.f;
    }
    for (int q = nops - 1; q >= 0; --q) {
        const int op = __ldg(&sc->xform_ops[XFORM_MAX_OPS * xf + q]);
        const float4 p = __ldg(&sc->xform_p[XFORM_MAX_OPS * xf + q]);
        if (op == XOP_TRANSLATE) { pt[0] += p.x; pt[1] += p.y; pt[2] += p.z; }
        else if (op == XOP_ROTATE_Y) {
            const float x = pt[0], z = pt[2], nx = n[0], nz = n[2];
            pt[0] = p.y * x + p.x * z; pt[2] = -(p.x * x) + p.y * z;
            n[0] = p.y * nx + p.x * nz; n[2] = -(p.x * nx) + p.y * nz;
        } else { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
    }
    HitGeom g;
    g.pu = make_float4(pt[0], pt[1], pt[2], u);
    g.nv = make_float4(n[0], n[1], n[2], v);
    return g;
}

// One iteration of `color` (core.clj:25-39) after the hit is known: builds the hit record
// (hitable.clj:193-201), evaluates emitted + scatter.  Returns true if the path continues with
// (o, d, time) replaced by the scattered ray and `atten` the attenuation factor; false with `reason`.
template <bool GEN>
__device__ __forceinline__ bool shade_hit(const DevScene& sc, const DevScene* scp, int k, float t, float3& o, float3& d, float& time,
                                          bool allow_scatter, const ScatterRng& rng, float3& atten, float3& emitted,
                                          int& reason) {
    const int4 rec = __ldg(&sc.shade_rec[k]);
    const float4 col = __ldg(&sc.shade_col[k]);
    const unsigned flags = (unsigned)rec.w;
    float3 p, n;
    float u = 0.f, v = 0.f;
    if (GEN) {
        const HitGeom g = hit_geom_generic(scp, k, t, o.x, o.y, o.z, d.x, d.y, d.z, time);
        p = f3(g.pu.x, g.pu.y, g.pu.z); u = g.pu.w;
        n = f3(g.nv.x, g.nv.y, g.nv.z); v = g.nv.w;
    } else {
        float4 c0r = __ldg(&sc.ex_c0r[k]);
        float3 center = f3(c0r.x, c0r.y, c0r.z);
        if (flags & SPH_MOVING) {
            float4 c1 = __ldg(&sc.ex_c1[k]);
            float2 tt = __ldg(&sc.ex_t0t1[k]);
            float f = (time - tt.x) / (tt.y - tt.x);
            center = (1.0f - f) * center + f * f3(c1.x, c1.y, c1.z);
        }
        p = t * d + o;                         // util.clj:18-22
        n = normalise3(p - center);            // hitable.clj:194
        if (flags & SPH_UV) {                  // hitable.clj:128-139
            float phi = atan2f(n.z, n.x);
            float theta = asinf(fminf(1.0f, fmaxf(-1.0f, n.y)));
            const float PI = 3.14159265358979323846f;
            u = 1.0f - (phi + PI) / (2.0f * PI);
            v = (theta + PI / 2.0f) / PI;
        }
    }
    const int type = rec.x, tex = rec.y;
    const bool const_tex = rec.z == TEX_CONSTANT;
    const float param = col.x;
    const float3 const_col = f3(col.y, col.z, col.w);
    emitted = (type == MAT_DIFFUSE_LIGHT) ? (const_tex ? const_col : tex_sample<GEN>(sc, scp, tex, u, v, p)) : f3(0.f, 0.f, 0.f);
    atten = f3(1.f, 1.f, 1.f);
    if (!allow_scatter) {                      // core.clj:26 (pos? depth) fails: scatter is not evaluated
        reason = TERM_DEPTH;
        return false;
    }
    if (type == MAT_LAMBERTIAN) {              // shader.clj:29-36: (p + n + s) - p
        float3 s = rng.unit_sphere();
        d = n + s;
        o = p;
        atten = const_tex ? const_col : tex_sample<GEN>(sc, scp, tex, u, v, p);
        return true;
    }
    if (type == MAT_METAL) {                   // shader.clj:46-59
        float3 refl = reflect3(normalise3(d), n);
        float3 s = rng.unit_sphere();
        float3 nd = refl + param * s;
        if (dot3(nd, n) > 0.f) {
            d = nd;
            o = p;
            atten = const_tex ? const_col : tex_sample<GEN>(sc, scp, tex, u, v, p);
            return true;
        }
        reason = TERM_ABSORB;
        return false;
    }
    if (type == MAT_DIELECTRIC) {              // shader.clj:76-104
        float ri = param;
        float ray_dot_n = dot3(d, n);
        float dmag = sqrtf(dot3(d, d));
        float3 outward;
        float ni_over_nt, cosine;
        if (ray_dot_n > 0.f) {
            outward = -n;
            ni_over_nt = ri;
            cosine = ri * (ray_dot_n / dmag);
        } else {
            outward = n;
            ni_over_nt = 1.0f / ri;
            cosine = -(ray_dot_n / dmag);
        }
        // refract (shader.clj:11-20)
        float3 uv = (dmag > 0.f) ? (1.0f / dmag) * d : d;
        float dt = dot3(uv, outward);
        float disc = 1.0f - ni_over_nt * ni_over_nt * (1.0f - dt * dt);
        float3 nd;
        if (disc > 0.f) {
            if (rng.rand() < schlick(cosine, ri))
                nd = reflect3(d, n);
            else
                nd = ni_over_nt * (uv - dt * outward) - sqrtf(disc) * outward;
        } else {
            nd = reflect3(d, n);
        }
        d = nd;
        o = p;
        return true;
    }
    if (GEN && type == MAT_ISOTROPIC) {        // shader.clj:129-138: the scattered ray's TIME is the hit's t, as written there
        d = rng.unit_sphere();
        o = p;
        time = t;
        atten = const_tex ? const_col : tex_sample<GEN>(sc, scp, tex, u, v, p);
        return true;
    }
    reason = TERM_LIGHT;                       // DiffuseLight: scatter -> nil (shader.clj:116-117)
    return false;
}

}  // namespace rt
