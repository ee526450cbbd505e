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
constexpr int MAT_LAMBERTIAN = 0, MAT_METAL = 1, MAT_DIELECTRIC = 2, MAT_DIFFUSE_LIGHT = 3;
constexpr int TEX_CONSTANT = 0, TEX_UV_GRADIENT = 1, TEX_CHECKERBOARD = 2;
constexpr int CAM_PINHOLE = 0, CAM_THIN_LENS = 1;

// cull tolerances: the cull direction is scaled up by sqrt(1 + CULL_EPS); see Culler (rt_kernels.cuh)
constexpr float CULL_EPS = 7.62939453125e-6f;  // 2^-17

enum TermReason { TERM_NONE = 0, TERM_LIGHT = 1, TERM_ABSORB = 2, TERM_DEPTH = 3, TERM_MISS = 4 };

// device counter slots (subset of RT_CTR_* that the kernels write)
enum { DC_RAYS = 0, DC_SAMPLES, DC_TERM_LIGHT, DC_TERM_ABSORB, DC_TERM_DEPTH, DC_TERM_MISS, DC_CANDIDATES, DC_COUNT = 8 };

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
// Keeps the closest t; exact ties go to the lower caller index (Hitlist, hitable.clj:17-26).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dadd_rn(a, -b); }
__device__ __forceinline__ double ddot(double ax, double ay, double az, double bx, double by, double bz) {
    return dadd(dadd(dmul(ax, bx), dmul(ay, by)), dmul(az, bz));
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
    double ocx = dsub((double)ox, cx), ocy = dsub((double)oy, cy), ocz = dsub((double)oz, cz);
    double ddx = dx, ddy = dy, ddz = dz;
    double a = ddot(ddx, ddy, ddz, ddx, ddy, ddz);
    double b = dmul(2.0, ddot(ocx, ocy, ocz, ddx, ddy, ddz));
    double c = dsub(ddot(ocx, ocy, ocz, ocx, ocy, ocz), dmul(r, r));
    double disc = dsub(dmul(b, b), dmul(dmul(4.0, a), c));
    if (disc >= 0.0) {
        double sq = __dsqrt_rn(disc);
        double two_a = dmul(2.0, a);
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
    }
    return CUDART_INF;
}

// ------------------------------------------------------------------------------------------
// texture.clj:14-50 sample (children of a checkerboard always have a smaller id: validated on upload)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float3 tex_sample(const DevScene& sc, int id, float u, float v, float3 p) {
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
        float s = __ldg(P);
        float sines = sinf(s * p.x) * sinf(s * p.y) * sinf(s * p.z);
        id = __ldg(&sc.tex_child[2 * id + ((sines < 0.f) ? 0 : 1)]);
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
    __device__ __forceinline__ float3 unit_sphere() const {
        return ball ? f3(ball[0], ball[1], ball[2]) : rand_in_unit_sphere(key, pixel, sample, bounce, 1u);
    }
    __device__ __forceinline__ float rand() const { return u ? *u : u01(rng_block(key, pixel, sample, bounce, 0u).x); }
};

// One iteration of `color` (core.clj:25-39) after the hit is known: builds the hit record
// (hitable.clj:193-201), evaluates emitted + scatter.  Returns true if the path continues with
// (o, d) replaced by the scattered ray and `atten` the attenuation factor; false with `reason`.
__device__ __forceinline__ bool shade_hit(const DevScene& sc, int k, float t, float3& o, float3& d, float time,
                                          bool allow_scatter, const ScatterRng& rng, float3& atten, float3& emitted,
                                          int& reason) {
    const int4 rec = __ldg(&sc.shade_rec[k]);
    const float4 col = __ldg(&sc.shade_col[k]);
    const unsigned flags = (unsigned)rec.w;
    float4 c0r = __ldg(&sc.ex_c0r[k]);
    float3 center = f3(c0r.x, c0r.y, c0r.z);
    if (flags & SPH_MOVING) {
        float4 c1 = __ldg(&sc.ex_c1[k]);
        float2 tt = __ldg(&sc.ex_t0t1[k]);
        float f = (time - tt.x) / (tt.y - tt.x);
        center = (1.0f - f) * center + f * f3(c1.x, c1.y, c1.z);
    }
    float3 p = t * d + o;                      // util.clj:18-22
    float3 n = normalise3(p - center);         // hitable.clj:194
    float u = 0.f, v = 0.f;
    if (flags & SPH_UV) {                      // hitable.clj:128-139
        float phi = atan2f(n.z, n.x);
        float theta = asinf(fminf(1.0f, fmaxf(-1.0f, n.y)));
        const float PI = 3.14159265358979323846f;
        u = 1.0f - (phi + PI) / (2.0f * PI);
        v = (theta + PI / 2.0f) / PI;
    }
    const int type = rec.x, tex = rec.y;
    const bool const_tex = rec.z == TEX_CONSTANT;
    const float param = col.x;
    const float3 const_col = f3(col.y, col.z, col.w);
    emitted = (type == MAT_DIFFUSE_LIGHT) ? (const_tex ? const_col : tex_sample(sc, tex, u, v, p)) : f3(0.f, 0.f, 0.f);
    atten = f3(1.f, 1.f, 1.f);
    if (!allow_scatter) {                      // core.clj:26 (pos? depth) fails: scatter is not evaluated
        reason = TERM_DEPTH;
        return false;
    }
    if (type == MAT_LAMBERTIAN) {              // shader.clj:29-36: (p + n + s) - p
        float3 s = rng.unit_sphere();
        d = n + s;
        o = p;
        atten = const_tex ? const_col : tex_sample(sc, tex, u, v, p);
        return true;
    }
    if (type == MAT_METAL) {                   // shader.clj:46-59
        float3 refl = reflect3(normalise3(d), n);
        float3 s = rng.unit_sphere();
        float3 nd = refl + param * s;
        if (dot3(nd, n) > 0.f) {
            d = nd;
            o = p;
            atten = const_tex ? const_col : tex_sample(sc, tex, u, v, p);
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
    reason = TERM_LIGHT;                       // DiffuseLight: scatter -> nil (shader.clj:116-117)
    return false;
}

}  // namespace rt
