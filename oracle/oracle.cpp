// oracle.cpp — CPU restatement of raytrace-clj's per-pixel path-tracing loop.
//
// TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library.  The product
// (raytrace_clj_b200 + libraytrace_b200.so) never does.
//
// Parity status: the reference (Clojure on the JVM) cannot be executed in this
// environment (no java / lein / clojure), so this restatement is
// pinned against (1) every known-answer test the reference's own test files hold
// for the path (test/raytrace_clj/hitable_test.clj:23-59,61-103 hit/miss booleans,
// lerp, :114-141 AABB hits and surrounding boxes; util_test.clj:44-49
// point-at-parameter) and (2) closed-form values derived from the reference
// formulas.  Numeric t / p / normal / uv, scatter, emitted, sample, get-ray,
// color, pixel are "parity unpinned" by the reference itself (it has no such
// tests) — fidelity there is by inspection, each function below citing the
// reference lines it follows; clojure/dump_vectors.clj prints the same vectors
// from the real reference for anyone with a JVM (tests/test_jvm_vectors.py
// ingests them).
//
// Round 2: the remaining leaf primitives and wrappers (hitable.clj:269-581: RectXY/XZ/YZ,
// FlipNormals, Translate, RotateY, Box, ConstantMedium, Triangle), the reference's own
// accelerator (AABB.hit? hitable.clj:36-48, bvh-node.hit? :97-106, make-bvh :108-123),
// Isotropic (shader.clj:129-143), the procedural / image textures (texture.clj:60-138,
// perlin.clj:6-64), and a REPLAY mode in which every random draw comes from the same
// Philox4x32-10 counters the CUDA kernels use (rt_device.cuh rng_block) through the same
// closed-form ball / disk maps, so whole paths can be compared sample by sample
// (orc_trace_paths).
//
// Arithmetic: IEEE double throughout, same operation order as the Clojure source,
// compiled with -ffp-contract=off (the JVM never fuses multiply-add).
// Third-party arithmetic restated from its documented behaviour (sources are not
// under /root/reference): net.mikera/core.matrix 0.52.0 + vectorz-clj 0.44.0
// (element-wise double ops; `normalise` = multiply by 1/magnitude; `lerp` =
// a*(1-f) + b*f), clojure.core/rand = Math.random() (uniform double in [0,1)).

#include <cmath>
#include <cstdint>
#include <cstring>
#include <cfloat>
#include <vector>
#include <algorithm>
#include <random>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct V3 {
    double x, y, z;
};
inline V3 v3(double a, double b, double c) { return V3{a, b, c}; }
inline V3 add(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 sub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 mul(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }  // mat/mul is element-wise
inline V3 mul(double s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
inline V3 neg(V3 a) { return v3(-a.x, -a.y, -a.z); }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline double magnitude(V3 a) { return std::sqrt(dot(a, a)); }
// vectorz Vector3.normalise: d = magnitude; if (d > 0) multiply(1.0 / d)
inline V3 normalise(V3 a) {
    double d = magnitude(a);
    if (d > 0) return mul(1.0 / d, a);
    return a;
}
inline double comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

struct Ray {
    V3 o, d;
    double time;
};

// util.clj:18-22  point-at-parameter = direction * t + origin
inline V3 point_at_parameter(const Ray& r, double t) { return add(mul(t, r.d), r.o); }

// ---- RNG ------------------------------------------------------------------------------------
// Mode A (default): stands in for clojure.core/rand (Math.random); the JVM stream is unseeded and cannot be
// reproduced, so the oracle uses xoshiro256** seeded per (seed, pixel, sample), draws taken in program order and
// the reference's rejection samplers (util.clj:32-52).
// Mode B (replay): Philox4x32-10 keyed exactly like rt_device.cuh: counter (pixel, sample, bounce << 16 | block,
// "RTB2"), key = seed lo / hi; block 0 of bounce 0 = (pixel jitter u, v, shutter u), block 1 of bounce 0 = lens disk,
// block 1 of bounce b >= 1 = (ball radius u, ball z u, ball phi u, dielectric rand), block 16 + i of bounce b = the
// `rand` a ConstantMedium at primitive index i draws inside hit?.  Uniforms are the top 24 bits (exact in float and
// double); ball / disk points come from the closed-form maps of rt_device.cuh, evaluated in double.
struct Xoshiro {
    uint64_t s[4];
    static uint64_t splitmix(uint64_t& x) {
        uint64_t z = (x += 0x9e3779b97f4a7c15ULL);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        return z ^ (z >> 31);
    }
    explicit Xoshiro(uint64_t seed) {
        uint64_t x = seed;
        for (int i = 0; i < 4; ++i) s[i] = splitmix(x);
    }
    static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        uint64_t result = rotl(s[1] * 5, 7) * 9;
        uint64_t t = s[1] << 17;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return result;
    }
    // uniform double in [0,1) with 53 random bits, like java.util.Random.nextDouble
    double rand() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int i = 0; i < 10; ++i) {
        uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += W0;
        k1 += W1;
    }
}
inline double u01_24(uint32_t x) { return (double)(x >> 8) * (1.0 / 16777216.0); }
constexpr uint32_t RNG_DOMAIN = 0x52544232u;  // "RTB2"

struct Rng {
    Xoshiro g;
    bool replay = false;
    uint32_t k0 = 0, k1 = 0, pixel = 0, sample = 0;
    explicit Rng(uint64_t seed) : g(seed) {}
    Rng(uint64_t seed, uint32_t pixel_, uint32_t sample_) : g(0), replay(true), k0((uint32_t)seed), k1((uint32_t)(seed >> 32)),
                                                            pixel(pixel_), sample(sample_) {}
    void block(uint32_t bounce, uint32_t blk, double u[4]) const {
        uint32_t c[4] = {pixel, sample, (bounce << 16) | blk, RNG_DOMAIN};
        philox4x32_10(c, k0, k1);
        for (int i = 0; i < 4; ++i) u[i] = u01_24(c[i]);
    }
    // core.clj:49-50 + camera.clj:47: the three camera uniforms (u drawn first)
    void camera_uniforms(double& ru, double& rv) {
        if (replay) { double u[4]; block(0, 0, u); ru = u[0]; rv = u[1]; }
        else { ru = g.rand(); rv = g.rand(); }
    }
    double shutter_uniform() {
        if (replay) { double u[4]; block(0, 0, u); return u[2]; }
        return g.rand();
    }
    // util.clj:32-41  rand-in-unit-disk: rejection in [-1,1)^2 x {0}, accept when dot < 1
    V3 unit_disk() {
        if (replay) {
            double u[4];
            block(0, 1, u);
            double rad = std::sqrt(u[0]), phi = 2.0 * M_PI * u[1];
            return v3(rad * std::cos(phi), rad * std::sin(phi), 0);
        }
        for (;;) {
            double x = 2.0 * g.rand() - 1.0;
            double y = 2.0 * g.rand() - 1.0;
            V3 p = v3(x, y, 0);
            if (!(dot(p, p) >= 1.0)) return p;
        }
    }
    // util.clj:43-52  rand-in-unit-sphere
    V3 unit_sphere(uint32_t bounce) {
        if (replay) {
            double u[4];
            block(bounce, 1, u);
            double rad = std::cbrt(u[0]), z = 1.0 - 2.0 * u[1];
            double s = std::sqrt(std::max(0.0, 1.0 - z * z)) * rad, phi = 2.0 * M_PI * u[2];
            return v3(s * std::cos(phi), s * std::sin(phi), z * rad);
        }
        for (;;) {
            double x = 2.0 * g.rand() - 1.0;
            double y = 2.0 * g.rand() - 1.0;
            double z = 2.0 * g.rand() - 1.0;
            V3 p = v3(x, y, z);
            if (!(dot(p, p) >= 1.0)) return p;
        }
    }
    // the Dielectric's (rand), shader.clj:94
    double scatter_rand(uint32_t bounce) {
        if (replay) { double u[4]; block(bounce, 1, u); return u[3]; }
        return g.rand();
    }
    // the (rand) inside ConstantMedium.hit?, hitable.clj:529
    double medium_rand(uint32_t bounce, int prim) {
        if (replay) { double u[4]; block(bounce, 16u + (uint32_t)prim, u); return u[0]; }
        return g.rand();
    }
};

// ---- scene tables (the marshalled form of the reference's record graph) ----------------
enum { MAT_LAMBERTIAN = 0, MAT_METAL = 1, MAT_DIELECTRIC = 2, MAT_DIFFUSE_LIGHT = 3, MAT_ISOTROPIC = 4 };
enum { TEX_CONSTANT = 0, TEX_UV_GRADIENT = 1, TEX_CHECKERBOARD = 2, TEX_PERLIN_NOISE = 3, TEX_PERLIN_TURB = 4,
       TEX_MARBLE = 5, TEX_FLIP_U = 6, TEX_FLIP_V = 7, TEX_IMAGE_MAP = 8 };
enum { SPH_UV = 1, SPH_MOVING = 2 };
enum { CAM_PINHOLE = 0, CAM_THIN_LENS = 1 };
enum { PRIM_SPHERE = 0, PRIM_RECT_XY = 1, PRIM_RECT_XZ = 2, PRIM_RECT_YZ = 3, PRIM_TRIANGLE = 4, PRIM_MEDIUM = 5 };
enum { XOP_NONE = 0, XOP_TRANSLATE = 1, XOP_ROTATE_Y = 2, XOP_FLIP = 3 };
enum { TIE_HITLIST = 0, TIE_BVH = 1 };
constexpr int XFORM_MAX_OPS = 4;

struct Prim {
    int type = PRIM_SPHERE;
    // sphere family (hitable.clj:141-259)
    V3 c0{0, 0, 0}, c1{0, 0, 0};
    double r = 0, t0 = 0, t1 = 1;
    uint32_t flags = 0;
    // rect: q[0..4] = a0 b0 a1 b1 k; triangle: q[0..8] = v0 v1 v2; medium: q[0] = density
    double q[12] = {0};
    int aux0 = 0, aux1 = 0;   // medium: first boundary primitive, count
    int xform = -1;
    int mat = 0;
};
struct XForm {
    int op[XFORM_MAX_OPS];
    double p[XFORM_MAX_OPS][4];   // TRANSLATE: offset xyz; ROTATE_Y: sin, cos
};
struct Material {
    int type;
    double param;
    int tex;
};
struct Texture {
    int type;
    double p[12];
    int child[2];
};
struct Image {
    int w = 0, h = 0;
    std::vector<uint8_t> rgb;   // rows top to bottom (imagez get-pixel x y)
};
struct AABB {
    V3 vmin, vmax;
};
struct BvhNode {
    int left, right;      // child node index, or ~prim for a leaf
    AABB box;
};
struct Scene {
    std::vector<Prim> prims;      // [0, n_world) = the world list in flatten order, then boundary primitives of media
    int n_world = 0;
    std::vector<XForm> xforms;
    std::vector<Material> mats;
    std::vector<Texture> texs;
    std::vector<Image> images;
    std::vector<V3> perlin_vec;                  // perlin.clj:6-8 random-vectors (256)
    std::vector<int> perm_x, perm_y, perm_z;     // perlin.clj:10-17
    int tie_rule = TIE_HITLIST;
    std::vector<BvhNode> bvh;                    // reference-style BVH over the world list (orc_scene_build_bvh)
    int bvh_root = -1;
};
struct Camera {
    int type;
    V3 origin, lleft, horiz, vert, u, v, w;
    double aperture, t0, t1;
};

struct HitRec {
    double t;
    V3 p;
    double uv[2];
    V3 normal;
    int mat;
    int id;
};

// hitable.clj:219-222  center-at-time = lerp(center0, center1, (t - t0)/(t1 - t0))
inline V3 center_at_time(V3 c0, double t0, V3 c1, double t1, double t) {
    double f = (t - t0) / (t1 - t0);
    return add(mul(1.0 - f, c0), mul(f, c1));
}

// hitable.clj:128-139  get-sphere-uv
inline void get_sphere_uv(V3 p, double uv[2]) {
    double phi = std::atan2(p.z, p.x);
    double theta = std::asin(p.y);
    uv[0] = 1.0 - (phi + M_PI) / (2.0 * M_PI);
    uv[1] = (theta + M_PI / 2.0) / M_PI;
}

// hitable.clj:182-207 (Sphere), 143-168 (UVSphere), 226-251 (MovingSphere): the quadratic,
// near root then far root, strict range test; normal = normalise(p - centre).
inline bool sphere_hit(const Prim& s, const Ray& r, double t_min, double t_max, HitRec& h) {
    V3 center = (s.flags & SPH_MOVING) ? center_at_time(s.c0, s.t0, s.c1, s.t1, r.time) : s.c0;
    V3 oc = sub(r.o, center);
    double a = dot(r.d, r.d);
    double b = 2.0 * dot(oc, r.d);
    double c = dot(oc, oc) - s.r * s.r;
    double discriminant = b * b - 4.0 * a * c;
    if (discriminant >= 0) {
        double sq = std::sqrt(discriminant);
        double t = (-b - sq) / (2.0 * a);
        if (!(t > t_min && t < t_max)) {
            t = (-b + sq) / (2.0 * a);
            if (!(t > t_min && t < t_max)) return false;
        }
        V3 p = point_at_parameter(r, t);
        V3 cpn = normalise(sub(p, center));
        h.t = t;
        h.p = p;
        h.normal = cpn;
        if (s.flags & SPH_UV) {
            get_sphere_uv(cpn, h.uv);
        } else {
            h.uv[0] = 0;
            h.uv[1] = 0;
        }
        return true;
    }
    return false;
}

// hitable.clj:269-363  RectXY / RectXZ / RectYZ.  axis = the constant coordinate (2, 1, 0); (A, B) the other two in the
// reference's order (x y | x z | y z).  NOTE the inclusive range test (>= t-min, <= t-max), unlike the spheres.
inline bool rect_hit(const Prim& s, const Ray& r, double t_min, double t_max, HitRec& h) {
    const int axis = s.type == PRIM_RECT_XY ? 2 : (s.type == PRIM_RECT_XZ ? 1 : 0);
    const int A = s.type == PRIM_RECT_YZ ? 1 : 0, B = s.type == PRIM_RECT_XY ? 1 : 2;
    const double a0 = s.q[0], b0 = s.q[1], a1 = s.q[2], b1 = s.q[3], k = s.q[4];
    double t = (k - comp(r.o, axis)) / comp(r.d, axis);
    if (t >= t_min && t <= t_max) {
        double a = comp(r.o, A) + t * comp(r.d, A);
        double b = comp(r.o, B) + t * comp(r.d, B);
        if (a >= a0 && a <= a1 && b >= b0 && b <= b1) {
            h.t = t;
            h.p = point_at_parameter(r, t);
            h.uv[0] = (a - a0) / (a1 - a0);
            h.uv[1] = (b - b0) / (b1 - b0);
            h.normal = v3(axis == 0 ? 1 : 0, axis == 1 ? 1 : 0, axis == 2 ? 1 : 0);
            return true;
        }
    }
    return false;
}

// hitable.clj:548-575  Triangle.hit?: Moeller-Trumbore, single-sided (det > 1e-8), u > 0, v > 0 strict,
// inclusive t range, un-normalised normal cross(v0v1, v0v2), uv = [u v].
inline bool triangle_hit(const Prim& s, const Ray& r, double t_min, double t_max, HitRec& h) {
    V3 v0 = v3(s.q[0], s.q[1], s.q[2]), v1 = v3(s.q[3], s.q[4], s.q[5]), v2 = v3(s.q[6], s.q[7], s.q[8]);
    V3 v0v1 = sub(v1, v0), v0v2 = sub(v2, v0);
    V3 pvec = cross(r.d, v0v2);
    double det = dot(v0v1, pvec);
    if (det > 0.00000001) {
        double inv_det = 1.0 / det;
        V3 tvec = sub(r.o, v0);
        double u = dot(tvec, pvec) * inv_det;
        if (u > 0 && u <= 1) {
            V3 qvec = cross(tvec, v0v1);
            double v = dot(r.d, qvec) * inv_det;
            if (v > 0 && (u + v) <= 1) {
                double t = dot(v0v2, qvec) * inv_det;
                if (t >= t_min && t <= t_max) {
                    h.t = t;
                    h.p = point_at_parameter(r, t);
                    h.uv[0] = u;
                    h.uv[1] = v;
                    h.normal = cross(v0v1, v0v2);
                    return true;
                }
            }
        }
    }
    return false;
}

struct HitCtx {            // what ConstantMedium.hit? needs beyond the ray (its `rand`)
    Rng* rng;
    uint32_t bounce;
};

bool prim_hit(const Scene& sc, int id, const Ray& r, double t_min, double t_max, HitRec& h, HitCtx* hc);
// the common leaf — a sphere without wrappers — inline (the brute-force loop of the benchmark scenes: this is the
// CPU baseline bench.py times, so it must not pay for the generality of prim_hit); everything else out of line
inline bool leaf_hit(const Scene& sc, int id, const Ray& r, double t_min, double t_max, HitRec& h, HitCtx* hc) {
    const Prim& s = sc.prims[id];
    if (s.type == PRIM_SPHERE && s.xform < 0) {
        if (!sphere_hit(s, r, t_min, t_max, h)) return false;
        h.mat = s.mat;
        h.id = id;
        return true;
    }
    return prim_hit(sc, id, r, t_min, t_max, h, hc);
}

// hitable.clj:15-26  Hitlist.hit? over primitives [first, first + count): reduce with shrinking t-max.  A sphere
// replaces the running hit only when strictly closer (t < t-max); a rect / triangle also at t == t-max.
inline bool list_hit(const Scene& sc, int first, int count, const Ray& r, double t_min, double t_max, HitRec& out, HitCtx* hc) {
    bool any = false;
    double closest = t_max;
    HitRec h;
    for (int i = first; i < first + count; ++i) {
        if (leaf_hit(sc, i, r, t_min, closest, h, hc)) {
            any = true;
            closest = h.t;
            out = h;
        }
    }
    return any;
}

// hitable.clj:516-543  ConstantMedium.hit?
inline bool medium_hit(const Scene& sc, const Prim& s, int id, const Ray& r, double t_min, double t_max, HitRec& h, HitCtx* hc) {
    HitRec h1, h2;
    if (!list_hit(sc, s.aux0, s.aux1, r, -(double)FLT_MAX, (double)FLT_MAX, h1, nullptr)) return false;
    if (!list_hit(sc, s.aux0, s.aux1, r, h1.t + 0.0001, (double)FLT_MAX, h2, nullptr)) return false;
    double t1 = h1.t, t2 = h2.t;
    t1 = (t1 < t_min) ? t_min : t1;
    t2 = (t2 > t_max) ? t_max : t2;
    if (t1 < t2) {
        t1 = (t1 < 0) ? 0 : t1;
        double mag = magnitude(r.d);
        double dist_in_boundary = (t2 - t1) * mag;
        double u = hc && hc->rng ? hc->rng->medium_rand(hc->bounce, id) : 0.5;
        double hit_distance = -(std::log(u) / s.q[0]);
        if (hit_distance < dist_in_boundary) {
            double new_t = t1 + hit_distance / mag;
            h.t = new_t;
            h.p = point_at_parameter(r, new_t);
            h.uv[0] = 0;
            h.uv[1] = 0;
            h.normal = v3(1, 0, 0);
            return true;
        }
    }
    return false;
}

inline bool base_hit(const Scene& sc, const Prim& s, int id, const Ray& r, double t_min, double t_max, HitRec& h, HitCtx* hc) {
    switch (s.type) {
        case PRIM_SPHERE: return sphere_hit(s, r, t_min, t_max, h);
        case PRIM_RECT_XY:
        case PRIM_RECT_XZ:
        case PRIM_RECT_YZ: return rect_hit(s, r, t_min, t_max, h);
        case PRIM_TRIANGLE: return triangle_hit(s, r, t_min, t_max, h);
        case PRIM_MEDIUM: return medium_hit(sc, s, id, r, t_min, t_max, h, hc);
    }
    return false;
}

// The wrapper chain of one leaf, outermost first: Translate (hitable.clj:391-400), RotateY (:410-455),
// FlipNormals (:375-381).  Each wrapper transforms the ray on the way in and the hit record on the way out.
bool prim_hit(const Scene& sc, int id, const Ray& r, double t_min, double t_max, HitRec& h, HitCtx* hc) {
    const Prim& s = sc.prims[id];
    Ray rr = r;
    int nops = 0;
    const XForm* xf = s.xform >= 0 ? &sc.xforms[s.xform] : nullptr;
    if (xf) {
        for (; nops < XFORM_MAX_OPS && xf->op[nops] != XOP_NONE; ++nops) {
            const double* p = xf->p[nops];
            if (xf->op[nops] == XOP_TRANSLATE) {
                rr.o = sub(rr.o, v3(p[0], p[1], p[2]));
            } else if (xf->op[nops] == XOP_ROTATE_Y) {
                const double sn = p[0], cs = p[1];
                V3 o = rr.o, d = rr.d;
                rr.o = v3(cs * o.x - sn * o.z, o.y, sn * o.x + cs * o.z);
                rr.d = v3(cs * d.x - sn * d.z, d.y, sn * d.x + cs * d.z);
            }
        }
    }
    if (!base_hit(sc, s, id, rr, t_min, t_max, h, hc)) return false;
    for (int i = nops - 1; i >= 0; --i) {
        const double* p = xf->p[i];
        if (xf->op[i] == XOP_TRANSLATE) {
            h.p = add(h.p, v3(p[0], p[1], p[2]));
        } else if (xf->op[i] == XOP_ROTATE_Y) {
            const double sn = p[0], cs = p[1];
            V3 hp = h.p, hn = h.normal;
            h.p = v3(cs * hp.x + sn * hp.z, hp.y, (-(sn * hp.x)) + cs * hp.z);
            h.normal = v3(cs * hn.x + sn * hn.z, hn.y, (-(sn * hn.x)) + cs * hn.z);
        } else if (xf->op[i] == XOP_FLIP) {
            h.normal = neg(h.normal);
        }
    }
    h.mat = s.mat;
    h.id = id;
    return true;
}

// ---- the reference's accelerator ----------------------------------------------------------
// clojure.core/min / max on doubles return NaN when either argument is NaN (Numbers.min / max)
inline double jmin(double a, double b) { return (a != a) ? a : ((b != b) ? b : (a < b ? a : b)); }
inline double jmax(double a, double b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }

// hitable.clj:36-48  AABB.hit?
inline bool aabb_hit(const AABB& bx, const Ray& r, double t_min, double t_max) {
    double m[3], n[3];
    for (int i = 0; i < 3; ++i) {
        m[i] = (comp(bx.vmin, i) - comp(r.o, i)) / comp(r.d, i);
        n[i] = (comp(bx.vmax, i) - comp(r.o, i)) / comp(r.d, i);
    }
    double t0[3], t1[3];
    for (int i = 0; i < 3; ++i) { t0[i] = jmin(m[i], n[i]); t1[i] = jmax(m[i], n[i]); }
    double tmin = jmax(jmax(jmax(t0[0], t0[1]), t0[2]), t_min);   // (mat/maximum t0)
    double tmax = jmin(jmin(jmin(t1[0], t1[1]), t1[2]), t_max);
    return tmax > tmin;
}
// hitable.clj:87-92
inline AABB surrounding(const AABB& a, const AABB& b) {
    return AABB{v3(jmin(a.vmin.x, b.vmin.x), jmin(a.vmin.y, b.vmin.y), jmin(a.vmin.z, b.vmin.z)),
                v3(jmax(a.vmax.x, b.vmax.x), jmax(a.vmax.y, b.vmax.y), jmax(a.vmax.z, b.vmax.z))};
}

// bbox of one leaf (hitable.clj:169-172, 208-211, 252-259, 294-296, 325-327, 356-358, 576-581) seen through its
// wrappers: Translate.bbox (:401-405) shifts, make-rotate-y (:457-486) takes the box of the 8 rotated corners,
// FlipNormals.bbox (:385-386) passes through.  The chain is applied innermost first.
AABB prim_bbox(const Scene& sc, int id, double ts, double te) {
    const Prim& s = sc.prims[id];
    AABB b;
    switch (s.type) {
        case PRIM_SPHERE: {
            V3 rr = v3(s.r, s.r, s.r);
            if (s.flags & SPH_MOVING) {
                V3 cs = center_at_time(s.c0, s.t0, s.c1, s.t1, ts), ce = center_at_time(s.c0, s.t0, s.c1, s.t1, te);
                b = surrounding(AABB{sub(cs, rr), add(cs, rr)}, AABB{sub(ce, rr), add(ce, rr)});
            } else {
                b = AABB{sub(s.c0, rr), add(s.c0, rr)};
            }
            break;
        }
        case PRIM_RECT_XY: b = AABB{v3(s.q[0], s.q[1], s.q[4] - 0.0001), v3(s.q[2], s.q[3], s.q[4] + 0.0001)}; break;
        case PRIM_RECT_XZ: b = AABB{v3(s.q[0], s.q[4] - 0.0001, s.q[1]), v3(s.q[2], s.q[4] + 0.0001, s.q[3])}; break;
        case PRIM_RECT_YZ: b = AABB{v3(s.q[4] - 0.0001, s.q[0], s.q[1]), v3(s.q[4] + 0.0001, s.q[2], s.q[3])}; break;
        case PRIM_TRIANGLE: {
            V3 v0 = v3(s.q[0], s.q[1], s.q[2]), v1 = v3(s.q[3], s.q[4], s.q[5]), v2 = v3(s.q[6], s.q[7], s.q[8]);
            V3 mx = v3(jmax(jmax(v0.x, v1.x), v2.x), jmax(jmax(v0.y, v1.y), v2.y), jmax(jmax(v0.z, v1.z), v2.z));
            V3 mn = v3(jmin(jmin(v0.x, v1.x), v2.x), jmin(jmin(v0.y, v1.y), v2.y), jmin(jmin(v0.z, v1.z), v2.z));
            b = AABB{sub(mn, v3(0.0001, 0.0001, 0.0001)), add(mx, v3(0.0001, 0.0001, 0.0001))};
            break;
        }
        default: {   // medium: the box of its boundary (hitable.clj:542-543); a multi-leaf boundary (a Box) = their union
            b = prim_bbox(sc, s.aux0, ts, te);
            for (int i = 1; i < s.aux1; ++i) b = surrounding(b, prim_bbox(sc, s.aux0 + i, ts, te));
            return b;
        }
    }
    if (s.xform >= 0) {
        const XForm& xf = sc.xforms[s.xform];
        int nops = 0;
        while (nops < XFORM_MAX_OPS && xf.op[nops] != XOP_NONE) ++nops;
        for (int i = nops - 1; i >= 0; --i) {
            const double* p = xf.p[i];
            if (xf.op[i] == XOP_TRANSLATE) {
                b = AABB{add(b.vmin, v3(p[0], p[1], p[2])), add(b.vmax, v3(p[0], p[1], p[2]))};
            } else if (xf.op[i] == XOP_ROTATE_Y) {
                const double sn = p[0], cs = p[1];
                V3 nmin = v3(FLT_MAX, FLT_MAX, FLT_MAX), nmax = v3(-(double)FLT_MAX, -(double)FLT_MAX, -(double)FLT_MAX);
                for (int ix = 0; ix < 2; ++ix)
                    for (int iy = 0; iy < 2; ++iy)
                        for (int iz = 0; iz < 2; ++iz) {
                            double x = ix ? b.vmax.x : b.vmin.x, y = iy ? b.vmax.y : b.vmin.y, z = iz ? b.vmax.z : b.vmin.z;
                            double nx = (cs * x) + (sn * z), nz = (-(sn * x)) + (cs * z);
                            nmin = v3(jmin(nmin.x, nx), jmin(nmin.y, y), jmin(nmin.z, nz));
                            nmax = v3(jmax(nmax.x, nx), jmax(nmax.y, y), jmax(nmax.z, nz));
                        }
                b = AABB{nmin, nmax};
            }
        }
    }
    return b;
}

// hitable.clj:108-123  make-bvh: random axis per node ((rand-int 3)), stable sort by bbox vmin[axis], split at n/2
// (the ratio n/2 takes ceil(n/2) items on the left); a 1-element node holds the same leaf as both children.
int make_bvh(Scene& sc, std::vector<int> items, const std::vector<AABB>& boxes, std::mt19937_64& rng) {
    int axis = (int)(rng() % 3);
    std::stable_sort(items.begin(), items.end(),
                     [&](int a, int b) { return comp(boxes[a].vmin, axis) < comp(boxes[b].vmin, axis); });
    const int n = (int)items.size();
    BvhNode node;
    if (n == 1) {
        node.left = node.right = ~items[0];
        node.box = boxes[items[0]];
    } else if (n == 2) {
        node.left = ~items[0];
        node.right = ~items[1];
        node.box = surrounding(boxes[items[0]], boxes[items[1]]);
    } else {
        const int k = (n + 1) / 2;
        int L = make_bvh(sc, std::vector<int>(items.begin(), items.begin() + k), boxes, rng);
        int R = make_bvh(sc, std::vector<int>(items.begin() + k, items.end()), boxes, rng);
        node.left = L;
        node.right = R;
        node.box = surrounding(sc.bvh[L].box, sc.bvh[R].box);
    }
    sc.bvh.push_back(node);
    return (int)sc.bvh.size() - 1;
}

struct BvhStats {
    uint64_t aabb_tests = 0, leaf_tests = 0;
};

// hitable.clj:97-106  bvh-node.hit?: both children get the SAME (t-min, t-max); the left hit is kept only when
// strictly closer, so the right child wins exact ties.
bool bvh_hit(const Scene& sc, int node, const Ray& r, double t_min, double t_max, HitRec& out, HitCtx* hc, BvhStats* st) {
    const BvhNode& n = sc.bvh[node];
    if (st) st->aabb_tests++;
    if (!aabb_hit(n.box, r, t_min, t_max)) return false;
    HitRec hl, hr;
    bool bl, br;
    if (n.left < 0) { if (st) st->leaf_tests++; bl = leaf_hit(sc, ~n.left, r, t_min, t_max, hl, hc); }
    else bl = bvh_hit(sc, n.left, r, t_min, t_max, hl, hc, st);
    if (n.right < 0) { if (st) st->leaf_tests++; br = leaf_hit(sc, ~n.right, r, t_min, t_max, hr, hc); }
    else br = bvh_hit(sc, n.right, r, t_min, t_max, hr, hc, st);
    if (bl && br) { out = (hl.t < hr.t) ? hl : hr; return true; }
    if (bl) { out = hl; return true; }
    if (br) { out = hr; return true; }
    return false;
}

// `hit? world`: the brute-force closest hit over the flattened world list (the form the GPU path restates), or the
// reference's BVH when one was built.  tie_rule BVH (the world is a bvh-node tree): every leaf sees the root's t-max
// and the right child wins ties, i.e. the LAST leaf in flatten order among exact ties.
inline bool world_hit(const Scene& sc, const Ray& r, double t_min, double t_max, HitRec& out, HitCtx* hc = nullptr,
                      bool use_bvh = false, BvhStats* st = nullptr) {
    if (use_bvh && sc.bvh_root >= 0) return bvh_hit(sc, sc.bvh_root, r, t_min, t_max, out, hc, st);
    if (sc.tie_rule == TIE_BVH) {
        // every leaf of a bvh-node tree sees the root's t-max and an equal t on the right replaces the left one.  Same
        // answer with a shrinking bound (only closer-or-EQUAL leaves build a hit record): a strict-range leaf (sphere,
        // t < t-max) is offered the next double above the running t, an inclusive one (t <= t-max) the running t itself.
        bool any = false;
        double closest = t_max;
        HitRec h;
        for (int i = 0; i < sc.n_world; ++i) {
            const bool strict = sc.prims[i].type == PRIM_SPHERE;
            const double lim = (any && strict) ? std::nextafter(closest, INFINITY) : closest;
            if (leaf_hit(sc, i, r, t_min, lim, h, hc)) {
                any = true;
                closest = h.t;
                out = h;
            }
        }
        return any;
    }
    return list_hit(sc, 0, sc.n_world, r, t_min, t_max, out, hc);
}

// ---- textures -------------------------------------------------------------------------------
// perlin.clj:19-50  noise = perlin-interp(perlin-coefficients(floor p), p - floor p)
double perlin_noise(const Scene& sc, V3 p) {
    double fi = std::floor(p.x), fj = std::floor(p.y), fk = std::floor(p.z);
    int i = (int)fi, j = (int)fj, k = (int)fk;
    V3 uvw = sub(p, v3((double)i, (double)j, (double)k));
    V3 c[8];
    for (int di = 0; di < 2; ++di)
        for (int dj = 0; dj < 2; ++dj)
            for (int dk = 0; dk < 2; ++dk)
                c[4 * di + 2 * dj + dk] = sc.perlin_vec[(size_t)(sc.perm_x[(i + di) & 255] ^ sc.perm_y[(j + dj) & 255] ^
                                                                  sc.perm_z[(k + dk) & 255])];
    double uu = uvw.x * uvw.x * (3 - 2 * uvw.x), vv = uvw.y * uvw.y * (3 - 2 * uvw.y), ww = uvw.z * uvw.z * (3 - 2 * uvw.z);
    double acc = 0;
    bool first = true;
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b)
            for (int cc = 0; cc < 2; ++cc) {
                V3 wv = sub(uvw, v3(a, b, cc));
                double term = (a * uu + (1.0 - a) * (1.0 - uu)) * (b * vv + (1.0 - b) * (1.0 - vv)) *
                              (cc * ww + (1.0 - cc) * (1.0 - ww)) * dot(wv, c[cc + 2 * b + 4 * a]);
                acc = first ? term : acc + term;   // (reduce + seq)
                first = false;
            }
    return acc;
}
// perlin.clj:52-64
double perlin_turbulence(const Scene& sc, V3 p, int depth) {
    double acc = 0, w = 1.0;
    V3 pt = p;
    for (int i = 0; i < depth; ++i) {
        acc = acc + w * perlin_noise(sc, pt);
        pt = mul(2.0, pt);
        w = w / 2.0;
    }
    return std::fabs(acc);
}

// texture.clj:14-138  sample
V3 tex_sample(const Scene& sc, int id, const double uv[2], V3 p) {
    const Texture& t = sc.texs[id];
    switch (t.type) {
        case TEX_CONSTANT:
            return v3(t.p[0], t.p[1], t.p[2]);
        case TEX_UV_GRADIENT: {
            V3 co = v3(t.p[0], t.p[1], t.p[2]), cu = v3(t.p[3], t.p[4], t.p[5]);
            V3 cv = v3(t.p[6], t.p[7], t.p[8]), cuv = v3(t.p[9], t.p[10], t.p[11]);
            double u = uv[0], v = uv[1];
            V3 a = add(mul(1 - u, cu), mul(u, co));
            V3 b = add(mul(1 - u, cuv), mul(u, cv));
            return add(mul(1 - v, b), mul(v, a));
        }
        case TEX_CHECKERBOARD: {
            double scale = t.p[0];
            double sines = std::sin(scale * p.x) * std::sin(scale * p.y) * std::sin(scale * p.z);
            return (sines < 0) ? tex_sample(sc, t.child[0], uv, p) : tex_sample(sc, t.child[1], uv, p);
        }
        case TEX_PERLIN_NOISE: {   // texture.clj:60-64
            double g = 0.5 * (1.0 + perlin_noise(sc, mul(t.p[0], p)));
            return v3(g, g, g);
        }
        case TEX_PERLIN_TURB: {    // texture.clj:74-78
            double g = 0.5 * (1.0 + perlin_turbulence(sc, mul(t.p[0], p), (int)t.p[1]));
            return v3(g, g, g);
        }
        case TEX_MARBLE: {         // texture.clj:88-93
            double g = 0.5 * (1.0 + std::sin(t.p[0] * p.z + 10.0 * perlin_turbulence(sc, p, (int)t.p[1])));
            return v3(g, g, g);
        }
        case TEX_FLIP_U: {         // texture.clj:103-106
            double uv2[2] = {1.0 - uv[0], uv[1]};
            return tex_sample(sc, t.child[0], uv2, p);
        }
        case TEX_FLIP_V: {         // texture.clj:113-116
            double uv2[2] = {uv[0], 1.0 - uv[1]};
            return tex_sample(sc, t.child[0], uv2, p);
        }
        case TEX_IMAGE_MAP: {      // texture.clj:126-133; the reference has no index clamp (u = 1 throws): clamped here
            const Image& im = sc.images[(size_t)t.p[0]];
            int i = (int)(uv[0] * im.w), j = (int)(uv[1] * im.h);
            i = std::min(std::max(i, 0), im.w - 1);
            j = std::min(std::max(j, 0), im.h - 1);
            const uint8_t* px = &im.rgb[((size_t)j * im.w + i) * 3];
            return v3(px[0] / 255.0, px[1] / 255.0, px[2] / 255.0);
        }
    }
    return v3(0, 0, 0);
}

// shader.clj:6-9  reflect = v - (2.0 * dot(v, n)) * n
inline V3 reflect(V3 v, V3 n) { return sub(v, mul(2.0 * dot(v, n), n)); }

// shader.clj:11-20  refract
inline bool refract(V3 v, V3 n, double ni_over_nt, V3& out) {
    V3 uv = normalise(v);
    double dt = dot(uv, n);
    double discriminant = 1.0 - ni_over_nt * ni_over_nt * (1 - dt * dt);
    if (discriminant > 0) {
        out = sub(mul(ni_over_nt, sub(uv, mul(dt, n))), mul(std::sqrt(discriminant), n));
        return true;
    }
    return false;
}

// shader.clj:69-74  schlick
inline double schlick(double cosine, double ri) {
    double r0 = (1.0 - ri) / (1.0 + ri);
    r0 = r0 * r0;
    return r0 + (1.0 - r0) * std::pow(1.0 - cosine, 5);
}

// random inputs of one scatter call: either drawn from the Rng or given by the caller
struct ScatterRand {
    Rng* g;
    uint32_t bounce;
    const double* ball;  // explicit rand-in-unit-sphere result, or null
    const double* u01;   // explicit (rand), or null
    V3 unit_sphere() { return ball ? v3(ball[0], ball[1], ball[2]) : g->unit_sphere(bounce); }
    double rand() { return u01 ? *u01 : g->scatter_rand(bounce); }
};

// shader.clj:29-36, 46-59, 76-104, 114-119, 129-138  scatter; returns false for nil
inline bool scatter(const Scene& sc, const Ray& rin, const HitRec& h, ScatterRand& rr, Ray& scattered,
                    V3& attenuation, int* why_not) {
    const Material& m = sc.mats[h.mat];
    switch (m.type) {
        case MAT_LAMBERTIAN: {
            V3 target = add(add(h.p, h.normal), rr.unit_sphere());
            scattered = Ray{h.p, sub(target, h.p), rin.time};
            attenuation = tex_sample(sc, m.tex, h.uv, h.p);
            return true;
        }
        case MAT_METAL: {
            V3 reflected = reflect(normalise(rin.d), h.normal);
            scattered = Ray{h.p, add(reflected, mul(m.param, rr.unit_sphere())), rin.time};
            if (dot(scattered.d, h.normal) > 0) {
                attenuation = tex_sample(sc, m.tex, h.uv, h.p);
                return true;
            }
            if (why_not) *why_not = 1;  // absorbed
            return false;
        }
        case MAT_DIELECTRIC: {
            double ri = m.param;
            V3 rd = rin.d;
            double ray_dot_n = dot(rd, h.normal);
            V3 outward_normal;
            double ni_over_nt, cosine;
            if (ray_dot_n > 0) {
                outward_normal = neg(h.normal);
                ni_over_nt = ri;
                cosine = ri * (ray_dot_n / magnitude(rd));
            } else {
                outward_normal = h.normal;
                ni_over_nt = 1.0 / ri;
                cosine = -(ray_dot_n / magnitude(rd));
            }
            V3 refr;
            attenuation = v3(1, 1, 1);
            if (refract(rd, outward_normal, ni_over_nt, refr)) {
                if (rr.rand() < schlick(cosine, ri))
                    scattered = Ray{h.p, reflect(rd, h.normal), rin.time};
                else
                    scattered = Ray{h.p, refr, rin.time};
            } else {
                scattered = Ray{h.p, reflect(rd, h.normal), rin.time};
            }
            return true;
        }
        case MAT_ISOTROPIC: {   // shader.clj:129-138: NOTE the scattered ray's TIME is the hit's t (as written in the reference)
            scattered = Ray{h.p, rr.unit_sphere(), h.t};
            attenuation = tex_sample(sc, m.tex, h.uv, h.p);
            return true;
        }
        case MAT_DIFFUSE_LIGHT:
        default:
            if (why_not) *why_not = 0;  // light: scatter -> nil
            return false;
    }
}

inline V3 emitted(const Scene& sc, const HitRec& h) {
    const Material& m = sc.mats[h.mat];
    if (m.type == MAT_DIFFUSE_LIGHT) return tex_sample(sc, m.tex, h.uv, h.p);
    return v3(0, 0, 0);
}

// camera.clj:8-16 / 35-48  get-ray.  Draw order: disk (pairs until accepted), then time.
inline Ray get_ray(const Camera& c, double s, double t, Rng& g) {
    if (c.type == CAM_PINHOLE) {
        V3 d = add(add(add(c.lleft, mul(s, c.horiz)), mul(t, c.vert)), neg(c.origin));
        return Ray{c.origin, d, 0};
    }
    double lens_radius = c.aperture / 2.0;
    V3 rd = mul(lens_radius, g.unit_disk());
    V3 offset = add(mul(rd.x, c.u), mul(rd.y, c.v));
    V3 o = add(c.origin, offset);
    V3 d = add(add(add(add(c.lleft, mul(s, c.horiz)), mul(t, c.vert)), neg(c.origin)), neg(offset));
    double time = c.t0 + (c.t1 - c.t0) * g.shutter_uniform();
    return Ray{o, d, time};
}

struct Counters {
    uint64_t rays = 0, samples = 0, term_light = 0, term_absorb = 0, term_depth = 0, term_miss = 0;
    BvhStats bvh;
};

// one logged bounce of a replayed path (layout shared with rt_path_bounce in include/raytrace_b200.h)
struct PathBounce {
    float o[3];
    float time;
    float d[3];
    int32_t hit_id;
    double t;
};
enum { TERM_LIGHT = 1, TERM_ABSORB = 2, TERM_DEPTH = 3, TERM_MISS = 4 };

// core.clj:17-41  color: iterative loop, depth cutoff, t-range (0.001, Float/MAX_VALUE), miss -> black
inline V3 color(const Scene& sc, Ray r, int depth, Rng& g, Counters& ctr, bool use_bvh = false, int* out_nrays = nullptr,
                int* out_term = nullptr, PathBounce* log = nullptr, int log_n = 0) {
    V3 atten = v3(1, 1, 1), accum = v3(0, 0, 0);
    uint32_t bounce = 0;
    for (;;) {
        ctr.rays++;
        bounce++;
        HitRec h;
        HitCtx hc{&g, bounce};
        const bool hit = world_hit(sc, r, 0.001, (double)FLT_MAX, h, &hc, use_bvh, &ctr.bvh);
        if (log && (int)bounce <= log_n) {
            PathBounce& b = log[bounce - 1];
            b.o[0] = (float)r.o.x; b.o[1] = (float)r.o.y; b.o[2] = (float)r.o.z; b.time = (float)r.time;
            b.d[0] = (float)r.d.x; b.d[1] = (float)r.d.y; b.d[2] = (float)r.d.z;
            b.hit_id = hit ? h.id : -1;
            b.t = hit ? h.t : INFINITY;
        }
        if (hit) {
            Ray scattered;
            V3 attenuation;
            int why = 0;
            ScatterRand rr{&g, bounce, nullptr, nullptr};
            bool scat = false;
            if (depth > 0)
                scat = scatter(sc, r, h, rr, scattered, attenuation, &why);
            else
                why = 2;
            V3 e = emitted(sc, h);
            if (scat) {
                accum = add(accum, mul(atten, e));
                atten = mul(atten, attenuation);
                r = scattered;
                depth--;
            } else {
                if (why == 0) ctr.term_light++;
                else if (why == 1) ctr.term_absorb++;
                else ctr.term_depth++;
                if (out_nrays) *out_nrays = (int)bounce;
                if (out_term) *out_term = why == 0 ? TERM_LIGHT : (why == 1 ? TERM_ABSORB : TERM_DEPTH);
                return add(accum, mul(atten, e));
            }
        } else {
            ctr.term_miss++;
            if (out_nrays) *out_nrays = (int)bounce;
            if (out_term) *out_term = TERM_MISS;
            return accum;
        }
    }
}

inline uint64_t mix_seed(uint64_t seed, uint64_t pixel, uint64_t sample) {
    uint64_t x = seed * 0x9e3779b97f4a7c15ULL + pixel;
    x = Xoshiro::splitmix(x);
    x ^= sample * 0xd1342543de82ef95ULL;
    return Xoshiro::splitmix(x);
}

Camera make_camera(int type, const float cam[24]) {
    Camera c;
    c.type = type;
    auto g = [&](int k) { return v3(cam[3 * k], cam[3 * k + 1], cam[3 * k + 2]); };
    c.origin = g(0);
    c.lleft = g(1);
    c.horiz = g(2);
    c.vert = g(3);
    c.u = g(4);
    c.v = g(5);
    c.w = g(6);
    c.aperture = cam[21];
    c.t0 = cam[22];
    c.t1 = cam[23];
    return c;
}

// one (pixel, sample) of core.clj:43-57 `pixel`: jitter, get-ray, color
inline V3 sample_pixel(const Scene& sc, const Camera& c, int nx, int ny, int i, int j, Rng& g, int max_depth, Counters& ctr,
                       bool use_bvh, int* out_nrays = nullptr, int* out_term = nullptr, PathBounce* log = nullptr, int log_n = 0) {
    // core.clj:49-50: u drawn first, then v; (float i) + rand, divided by nx
    double ru, rv;
    g.camera_uniforms(ru, rv);
    double u = ((double)(float)i + ru) / nx;
    double v = ((double)(float)j + rv) / ny;
    Ray r = get_ray(c, u, v, g);
    ctr.samples++;
    return color(sc, r, max_depth, g, ctr, use_bvh, out_nrays, out_term, log, log_n);
}

}  // namespace

extern "C" {

// ---- small known-answer helpers ---------------------------------------------------------
void orc_point_at_parameter(const double o[3], const double d[3], double t, double out[3]) {
    Ray r{v3(o[0], o[1], o[2]), v3(d[0], d[1], d[2]), 0};
    V3 p = point_at_parameter(r, t);
    out[0] = p.x; out[1] = p.y; out[2] = p.z;
}

void orc_center_at_time(const double c0[3], double t0, const double c1[3], double t1, double t, double out[3]) {
    V3 c = center_at_time(v3(c0[0], c0[1], c0[2]), t0, v3(c1[0], c1[1], c1[2]), t1, t);
    out[0] = c.x; out[1] = c.y; out[2] = c.z;
}

// Sphere.hit? / MovingSphere.hit? on doubles. flags: SPH_UV / SPH_MOVING. Returns 1 on hit.
int orc_sphere_hit(const double c0[3], const double c1[3], double t0, double t1, double radius, uint32_t flags,
                   const double o[3], const double d[3], double time, double t_min, double t_max,
                   double* out_t, double out_p[3], double out_n[3], double out_uv[2]) {
    Prim s;
    s.c0 = v3(c0[0], c0[1], c0[2]);
    s.c1 = c1 ? v3(c1[0], c1[1], c1[2]) : s.c0;
    s.t0 = t0; s.t1 = t1; s.r = radius; s.flags = flags; s.mat = 0;
    Ray r{v3(o[0], o[1], o[2]), v3(d[0], d[1], d[2]), time};
    HitRec h;
    if (!sphere_hit(s, r, t_min, t_max, h)) return 0;
    if (out_t) *out_t = h.t;
    if (out_p) { out_p[0] = h.p.x; out_p[1] = h.p.y; out_p[2] = h.p.z; }
    if (out_n) { out_n[0] = h.normal.x; out_n[1] = h.normal.y; out_n[2] = h.normal.z; }
    if (out_uv) { out_uv[0] = h.uv[0]; out_uv[1] = h.uv[1]; }
    return 1;
}

// AABB.hit? (hitable.clj:36-48) and make-surrounding-bbox (:87-92) on doubles
int orc_aabb_hit(const double vmin[3], const double vmax[3], const double o[3], const double d[3], double t_min, double t_max) {
    AABB b{v3(vmin[0], vmin[1], vmin[2]), v3(vmax[0], vmax[1], vmax[2])};
    Ray r{v3(o[0], o[1], o[2]), v3(d[0], d[1], d[2]), 0};
    return aabb_hit(b, r, t_min, t_max) ? 1 : 0;
}
void orc_surrounding_bbox(const double a[6], const double b[6], double out[6]) {
    AABB s = surrounding(AABB{v3(a[0], a[1], a[2]), v3(a[3], a[4], a[5])}, AABB{v3(b[0], b[1], b[2]), v3(b[3], b[4], b[5])});
    out[0] = s.vmin.x; out[1] = s.vmin.y; out[2] = s.vmin.z; out[3] = s.vmax.x; out[4] = s.vmax.y; out[5] = s.vmax.z;
}

void orc_get_sphere_uv(const double n[3], double uv[2]) { get_sphere_uv(v3(n[0], n[1], n[2]), uv); }

void orc_reflect(const double v[3], const double n[3], double out[3]) {
    V3 r = reflect(v3(v[0], v[1], v[2]), v3(n[0], n[1], n[2]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
int orc_refract(const double v[3], const double n[3], double ni_over_nt, double out[3]) {
    V3 r;
    if (!refract(v3(v[0], v[1], v[2]), v3(n[0], n[1], n[2]), ni_over_nt, r)) return 0;
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
    return 1;
}
double orc_schlick(double cosine, double ri) { return schlick(cosine, ri); }

// camera.clj:50-66 thin-lens-camera / 18-33 pinhole-camera: fills the 24-double camera record
void orc_thin_lens_camera(const double lookfrom[3], const double lookat[3], const double vup[3], double vfov,
                          double aspect, double aperture, double focus_dist, double t0, double t1, double out[24]) {
    V3 lf = v3(lookfrom[0], lookfrom[1], lookfrom[2]), la = v3(lookat[0], lookat[1], lookat[2]);
    V3 up = v3(vup[0], vup[1], vup[2]);
    double theta = vfov * (M_PI / 180.0);
    double half_height = std::tan(theta / 2.0);
    double half_width = aspect * half_height;
    V3 w = normalise(sub(lf, la));
    V3 u = normalise(cross(up, w));
    V3 v = cross(w, u);
    V3 lleft = sub(lf, add(add(mul(focus_dist * half_width, u), mul(focus_dist * half_height, v)), mul(focus_dist, w)));
    V3 horiz = mul(2.0 * focus_dist * half_width, u);
    V3 vert = mul(2.0 * focus_dist * half_height, v);
    V3 all[7] = {lf, lleft, horiz, vert, u, v, w};
    for (int k = 0; k < 7; ++k) { out[3 * k] = all[k].x; out[3 * k + 1] = all[k].y; out[3 * k + 2] = all[k].z; }
    out[21] = aperture; out[22] = t0; out[23] = t1;
}
void orc_pinhole_camera(const double lookfrom[3], const double lookat[3], const double vup[3], double vfov,
                        double aspect, double out[24]) {
    V3 lf = v3(lookfrom[0], lookfrom[1], lookfrom[2]), la = v3(lookat[0], lookat[1], lookat[2]);
    V3 up = v3(vup[0], vup[1], vup[2]);
    double theta = vfov * (M_PI / 180.0);
    double half_height = std::tan(theta / 2.0);
    double half_width = aspect * half_height;
    V3 w = normalise(sub(lf, la));
    V3 u = normalise(cross(up, w));
    V3 v = cross(w, u);
    V3 lleft = sub(lf, add(add(mul(half_width, u), mul(half_height, v)), w));
    V3 horiz = mul(2.0 * half_width, u);
    V3 vert = mul(2.0 * half_height, v);
    V3 all[7] = {lf, lleft, horiz, vert, u, v, w};
    for (int k = 0; k < 7; ++k) { out[3 * k] = all[k].x; out[3 * k + 1] = all[k].y; out[3 * k + 2] = all[k].z; }
    out[21] = 0; out[22] = 0; out[23] = 0;
}

// get-ray with caller-given randoms: disk point (dx, dy) and the time draw
void orc_get_ray(int cam_type, const float cam[24], double s, double t, double disk_x, double disk_y, double time_u,
                 double out_o[3], double out_d[3], double* out_time) {
    Camera c = make_camera(cam_type, cam);
    Ray r;
    if (c.type == CAM_PINHOLE) {
        r = Ray{c.origin, add(add(add(c.lleft, mul(s, c.horiz)), mul(t, c.vert)), neg(c.origin)), 0};
    } else {
        double lens_radius = c.aperture / 2.0;
        V3 rd = mul(lens_radius, v3(disk_x, disk_y, 0));
        V3 offset = add(mul(rd.x, c.u), mul(rd.y, c.v));
        r.o = add(c.origin, offset);
        r.d = add(add(add(add(c.lleft, mul(s, c.horiz)), mul(t, c.vert)), neg(c.origin)), neg(offset));
        r.time = c.t0 + (c.t1 - c.t0) * time_u;
    }
    out_o[0] = r.o.x; out_o[1] = r.o.y; out_o[2] = r.o.z;
    out_d[0] = r.d.x; out_d[1] = r.d.y; out_d[2] = r.d.z;
    *out_time = r.time;
}

// The samplers by themselves: n points of rand-in-unit-sphere (dim 3) / rand-in-unit-disk (dim 2).  replay = 0: the
// reference's rejection loops (util.clj:32-52) on the xoshiro stream; replay = 1: the closed-form maps on the Philox
// counters (pixel = index, sample 0, bounce 1) — the same counters / maps the CUDA kernels use.
void orc_sample_ball(int n, uint64_t seed, int replay, double* out) {
    Rng seq(seed);
    for (int i = 0; i < n; ++i) {
        Rng ph(seed, (uint32_t)i, 0u);
        V3 p = replay ? ph.unit_sphere(1) : seq.unit_sphere(1);
        out[3 * i] = p.x; out[3 * i + 1] = p.y; out[3 * i + 2] = p.z;
    }
}
void orc_sample_disk(int n, uint64_t seed, int replay, double* out) {
    Rng seq(seed);
    for (int i = 0; i < n; ++i) {
        Rng ph(seed, (uint32_t)i, 0u);
        V3 p = replay ? ph.unit_disk() : seq.unit_disk();
        out[2 * i] = p.x; out[2 * i + 1] = p.y;
    }
}

// ---- scene handle ---------------------------------------------------------------------------
// The marshalled form: exactly the buffers of rt_scene_desc (+ rt_scene_ext, all optional).  n = world primitives;
// n_boundary more primitives (the boundaries of media) follow in every per-primitive array.
void* orc_scene_create_ex(int n, const float* c0r, const float* c1, const float* t0t1, const uint32_t* flags,
                          const int32_t* mat_id, int nm, const int32_t* mtype, const float* mparam, const int32_t* mtex,
                          int nt, const int32_t* ttype, const float* tparams, const int32_t* tchild,
                          int n_boundary, const int32_t* prim_type, const float* prim_params, const int32_t* prim_aux,
                          const int32_t* prim_xform, int n_xforms, const int32_t* xform_ops, const float* xform_params,
                          int tie_rule, const float* perlin_vectors, const int32_t* perlin_perm, int n_images,
                          const int32_t* image_wh, const int64_t* image_offset, const uint8_t* image_rgb) {
    Scene* sc = new Scene();
    const int total = n + std::max(0, n_boundary);
    sc->n_world = n;
    sc->tie_rule = tie_rule;
    sc->prims.resize(total);
    for (int i = 0; i < total; ++i) {
        Prim& s = sc->prims[i];
        s.type = prim_type ? prim_type[i] : PRIM_SPHERE;
        s.c0 = v3(c0r[4 * i], c0r[4 * i + 1], c0r[4 * i + 2]);
        s.r = c0r[4 * i + 3];
        s.flags = flags ? flags[i] : 0;
        if ((s.flags & SPH_MOVING) && c1 && t0t1) {
            s.c1 = v3(c1[4 * i], c1[4 * i + 1], c1[4 * i + 2]);
            s.t0 = t0t1[2 * i];
            s.t1 = t0t1[2 * i + 1];
        } else {
            s.c1 = s.c0;
            s.t0 = 0;
            s.t1 = 1;
            s.flags &= ~(uint32_t)SPH_MOVING;
        }
        if (prim_params)
            for (int k = 0; k < 12; ++k) s.q[k] = prim_params[12 * i + k];
        if (prim_aux) { s.aux0 = prim_aux[2 * i]; s.aux1 = prim_aux[2 * i + 1]; }
        s.xform = prim_xform ? prim_xform[i] : -1;
        s.mat = mat_id[i];
    }
    sc->xforms.resize(std::max(0, n_xforms));
    for (int x = 0; x < n_xforms; ++x)
        for (int k = 0; k < XFORM_MAX_OPS; ++k) {
            sc->xforms[x].op[k] = xform_ops[XFORM_MAX_OPS * x + k];
            for (int q = 0; q < 4; ++q) sc->xforms[x].p[k][q] = xform_params[(XFORM_MAX_OPS * x + k) * 4 + q];
        }
    sc->mats.resize(nm);
    for (int i = 0; i < nm; ++i) sc->mats[i] = Material{mtype[i], (double)mparam[i], mtex[i]};
    sc->texs.resize(nt);
    for (int i = 0; i < nt; ++i) {
        Texture& t = sc->texs[i];
        t.type = ttype[i];
        for (int k = 0; k < 12; ++k) t.p[k] = tparams[12 * i + k];
        t.child[0] = tchild[2 * i];
        t.child[1] = tchild[2 * i + 1];
    }
    if (perlin_vectors && perlin_perm) {
        sc->perlin_vec.resize(256);
        sc->perm_x.resize(256); sc->perm_y.resize(256); sc->perm_z.resize(256);
        for (int i = 0; i < 256; ++i) {
            sc->perlin_vec[i] = v3(perlin_vectors[3 * i], perlin_vectors[3 * i + 1], perlin_vectors[3 * i + 2]);
            sc->perm_x[i] = perlin_perm[i]; sc->perm_y[i] = perlin_perm[256 + i]; sc->perm_z[i] = perlin_perm[512 + i];
        }
    }
    sc->images.resize(std::max(0, n_images));
    for (int i = 0; i < n_images; ++i) {
        Image& im = sc->images[i];
        im.w = image_wh[2 * i]; im.h = image_wh[2 * i + 1];
        im.rgb.assign(image_rgb + image_offset[i], image_rgb + image_offset[i] + (size_t)im.w * im.h * 3);
    }
    return sc;
}
void* orc_scene_create(int n, const float* c0r, const float* c1, const float* t0t1, const uint32_t* flags,
                       const int32_t* mat_id, int nm, const int32_t* mtype, const float* mparam,
                       const int32_t* mtex, int nt, const int32_t* ttype, const float* tparams,
                       const int32_t* tchild) {
    return orc_scene_create_ex(n, c0r, c1, t0t1, flags, mat_id, nm, mtype, mparam, mtex, nt, ttype, tparams, tchild, 0, nullptr,
                               nullptr, nullptr, nullptr, 0, nullptr, nullptr, TIE_HITLIST, nullptr, nullptr, 0, nullptr, nullptr,
                               nullptr);
}
void orc_scene_destroy(void* sc) { delete (Scene*)sc; }

// Build the reference's BVH (make-bvh, hitable.clj:108-123) over the world list; returns the node count.
// The axis choices come from a seeded generator ((rand-int 3) in the reference is unseeded).
int orc_scene_build_bvh(void* scene, double t0, double t1, uint64_t seed) {
    Scene& sc = *(Scene*)scene;
    sc.bvh.clear();
    std::mt19937_64 rng(seed);
    std::vector<int> items(sc.n_world);
    std::vector<AABB> boxes(sc.n_world);
    for (int i = 0; i < sc.n_world; ++i) { items[i] = i; boxes[i] = prim_bbox(sc, i, t0, t1); }
    sc.bvh_root = make_bvh(sc, items, boxes, rng);
    return (int)sc.bvh.size();
}

// bbox of world primitive i over [t0, t1] seen through its wrappers: out = vmin xyz, vmax xyz
void orc_prim_bbox(void* scene, int i, double t0, double t1, double out[6]) {
    AABB b = prim_bbox(*(Scene*)scene, i, t0, t1);
    out[0] = b.vmin.x; out[1] = b.vmin.y; out[2] = b.vmin.z; out[3] = b.vmax.x; out[4] = b.vmax.y; out[5] = b.vmax.z;
}

// `hit? world` for n rays (float inputs promoted to double): out_t, out_id (-1 = miss);
// optional out_t2 = second-smallest valid t over the other primitives (inf if none), for the
// "two best within 1e-5" id-ambiguity rule of the parity plan.  use_bvh: traverse the reference-style BVH
// (orc_scene_build_bvh) instead of the flat list; stats[2] (optional) += AABB tests, leaf tests.
// out_pnuv (optional, 8 doubles per ray): hit point, normal, uv.  Media draw their `rand` as 0.5 here (no path context).
void orc_hit_ex(void* scene, int n, const float* origins, const float* dirs, const float* times, double t_min,
                double t_max, int use_bvh, double* out_t, int32_t* out_id, double* out_t2, double* out_pnuv, uint64_t* stats) {
    const Scene& sc = *(Scene*)scene;
    uint64_t st_a = 0, st_l = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : st_a, st_l)
    for (int i = 0; i < n; ++i) {
        Ray r{v3(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]), v3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]),
              times ? (double)times[i] : 0.0};
        HitRec h;
        BvhStats st;
        if (world_hit(sc, r, t_min, t_max, h, nullptr, use_bvh != 0, &st)) {
            out_t[i] = h.t;
            out_id[i] = h.id;
            if (out_pnuv) {
                double* q = out_pnuv + 8 * (size_t)i;
                q[0] = h.p.x; q[1] = h.p.y; q[2] = h.p.z; q[3] = h.normal.x; q[4] = h.normal.y; q[5] = h.normal.z;
                q[6] = h.uv[0]; q[7] = h.uv[1];
            }
        } else {
            out_t[i] = INFINITY;
            out_id[i] = -1;
            if (out_pnuv)
                for (int k = 0; k < 8; ++k) out_pnuv[8 * (size_t)i + k] = 0;
        }
        st_a += st.aabb_tests;
        st_l += st.leaf_tests;
        if (out_t2) {
            double best2 = INFINITY;
            for (int k = 0; k < sc.n_world; ++k) {
                if (k == out_id[i]) continue;
                HitRec h2;
                if (prim_hit(sc, k, r, t_min, t_max, h2, nullptr) && h2.t < best2) best2 = h2.t;
            }
            out_t2[i] = best2;
        }
    }
    if (stats) { stats[0] += st_a; stats[1] += st_l; }
}
void orc_hit(void* scene, int n, const float* origins, const float* dirs, const float* times, double t_min,
             double t_max, double* out_t, int32_t* out_id, double* out_t2) {
    orc_hit_ex(scene, n, origins, dirs, times, t_min, t_max, 0, out_t, out_id, out_t2, nullptr, nullptr);
}

// One scatter + emitted per ray with explicit random inputs (ball = rand-in-unit-sphere result,
// u01 = the dielectric (rand)); the hit record is recomputed from (ray, hit_id) exactly as hit? does.
void orc_shade_batch(void* scene, int n, const float* origins, const float* dirs, const float* times,
                     const int32_t* hit_id, const float* ball, const float* u01, double* out_origin, double* out_dir,
                     double* out_atten, double* out_emitted, int32_t* out_flags, double* out_t) {
    const Scene& sc = *(Scene*)scene;
    for (int i = 0; i < n; ++i) {
        Ray r{v3(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]), v3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]),
              times ? (double)times[i] : 0.0};
        HitRec h;
        out_flags[i] = -1;
        for (int k = 0; k < 3; ++k) out_origin[3 * i + k] = out_dir[3 * i + k] = out_atten[3 * i + k] = out_emitted[3 * i + k] = 0;
        // the t-range of core.clj:25; prim_hit applies each type's own strict / inclusive test
        if (hit_id[i] < 0 || hit_id[i] >= (int)sc.prims.size() ||
            !prim_hit(sc, hit_id[i], r, 0.001, (double)FLT_MAX, h, nullptr))
            continue;
        if (out_t) out_t[i] = h.t;
        double b[3] = {ball[3 * i], ball[3 * i + 1], ball[3 * i + 2]};
        double u = u01[i];
        ScatterRand rr{nullptr, 0u, b, &u};
        Ray sca;
        V3 att = v3(0, 0, 0);
        bool ok = scatter(sc, r, h, rr, sca, att, nullptr);
        V3 e = emitted(sc, h);
        out_emitted[3 * i] = e.x; out_emitted[3 * i + 1] = e.y; out_emitted[3 * i + 2] = e.z;
        out_flags[i] = ok ? 1 : 0;
        if (ok) {
            out_origin[3 * i] = sca.o.x; out_origin[3 * i + 1] = sca.o.y; out_origin[3 * i + 2] = sca.o.z;
            out_dir[3 * i] = sca.d.x; out_dir[3 * i + 1] = sca.d.y; out_dir[3 * i + 2] = sca.d.z;
            out_atten[3 * i] = att.x; out_atten[3 * i + 1] = att.y; out_atten[3 * i + 2] = att.z;
        }
    }
}

void orc_tex_sample(void* scene, int tex, double u, double v, const double p[3], double out[3]) {
    double uv[2] = {u, v};
    V3 c = tex_sample(*(Scene*)scene, tex, uv, v3(p[0], p[1], p[2]));
    out[0] = c.x; out[1] = c.y; out[2] = c.z;
}
double orc_perlin_noise(void* scene, const double p[3]) { return perlin_noise(*(Scene*)scene, v3(p[0], p[1], p[2])); }
double orc_perlin_turbulence(void* scene, const double p[3], int depth) {
    return perlin_turbulence(*(Scene*)scene, v3(p[0], p[1], p[2]), depth);
}

// core.clj:43-57 pixel (the sample loop and the sum) for every pixel of rows j % row_stride == row_offset,
// samples [s_begin, s_begin + s_count): sum_rgb[((j*nx)+i)*3 + c] += SUM of color (double, j = 0 bottom row).
// counters[8]: rays, primitive tests (brute force: rays * n; BVH: leaf tests), samples, term_light, term_absorb,
// term_depth, term_miss, AABB tests (BVH mode).  mode bit 0: replay (Philox counters instead of the xoshiro stream);
// bit 1: traverse the reference-style BVH (orc_scene_build_bvh) instead of the flat list.
void orc_render_accumulate_ex(void* scene, int cam_type, const float cam[24], int nx, int ny, int s_begin, int s_count,
                              int row_offset, int row_stride, int max_depth, uint64_t seed, double* sum_rgb,
                              uint64_t counters[8], int n_threads, int mode) {
    const Scene& sc = *(Scene*)scene;
    Camera c = make_camera(cam_type, cam);
    Counters total;
    const bool replay = (mode & 1) != 0, use_bvh = (mode & 2) != 0 && sc.bvh_root >= 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel
    {
        Counters ctr;
#pragma omp for schedule(dynamic, 16)
        for (int pix = 0; pix < nx * ny; ++pix) {
            int i = pix % nx, j = pix / nx;
            if (row_stride > 1 && (j % row_stride) != row_offset) continue;
            V3 sum = v3(0, 0, 0);
            for (int s = s_begin; s < s_begin + s_count; ++s) {
                Rng g = replay ? Rng(seed, (uint32_t)pix, (uint32_t)s) : Rng(mix_seed(seed, (uint64_t)pix, (uint64_t)s));
                sum = add(sum, sample_pixel(sc, c, nx, ny, i, j, g, max_depth, ctr, use_bvh));
            }
            sum_rgb[3 * pix] += sum.x;
            sum_rgb[3 * pix + 1] += sum.y;
            sum_rgb[3 * pix + 2] += sum.z;
        }
#pragma omp critical
        {
            total.rays += ctr.rays; total.samples += ctr.samples; total.term_light += ctr.term_light;
            total.term_absorb += ctr.term_absorb; total.term_depth += ctr.term_depth; total.term_miss += ctr.term_miss;
            total.bvh.aabb_tests += ctr.bvh.aabb_tests; total.bvh.leaf_tests += ctr.bvh.leaf_tests;
        }
    }
    if (counters) {
        counters[0] += total.rays;
        counters[1] += use_bvh ? total.bvh.leaf_tests : total.rays * (uint64_t)sc.n_world;
        counters[2] += total.samples;
        counters[3] += total.term_light;
        counters[4] += total.term_absorb;
        counters[5] += total.term_depth;
        counters[6] += total.term_miss;
        counters[7] += total.bvh.aabb_tests;
    }
}
void orc_render_accumulate(void* scene, int cam_type, const float cam[24], int nx, int ny, int s_begin, int s_count,
                           int row_offset, int row_stride, int max_depth, uint64_t seed, double* sum_rgb,
                           uint64_t counters[8], int n_threads) {
    orc_render_accumulate_ex(scene, cam_type, cam, nx, ny, s_begin, s_count, row_offset, row_stride, max_depth, seed, sum_rgb,
                             counters, n_threads, 0);
}

// Replay of chosen (pixel, sample) pairs on the Philox counters (the CPU side of rt_trace_paths): per path the
// radiance `color` returns (core.clj:17-41), the number of rays, how the path ended (1 light, 2 absorbed, 3 depth,
// 4 miss) and, optionally, the first log_bounces rays with their hits (PathBounce, 40 bytes each).
void orc_trace_paths(void* scene, int cam_type, const float cam[24], int nx, int ny, int n, const int32_t* pixel,
                     const int32_t* sample, int max_depth, uint64_t seed, double* out_radiance, int32_t* out_nrays,
                     int32_t* out_term, int log_bounces, void* out_log, int n_threads) {
    const Scene& sc = *(Scene*)scene;
    Camera c = make_camera(cam_type, cam);
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    PathBounce* log = (PathBounce*)out_log;
#pragma omp parallel for schedule(dynamic, 64)
    for (int q = 0; q < n; ++q) {
        Counters ctr;
        const int pix = pixel[q], i = pix % nx, j = pix / nx;
        Rng g(seed, (uint32_t)pix, (uint32_t)sample[q]);
        int nr = 0, term = 0;
        PathBounce* lg = (log && log_bounces > 0) ? log + (size_t)q * log_bounces : nullptr;
        if (lg) {
            memset(lg, 0, sizeof(PathBounce) * (size_t)log_bounces);
            for (int b = 0; b < log_bounces; ++b) lg[b].hit_id = -2;   // -2 = no such bounce
        }
        V3 rad = sample_pixel(sc, c, nx, ny, i, j, g, max_depth, ctr, false, &nr, &term, lg, log_bounces);
        out_radiance[3 * q] = rad.x; out_radiance[3 * q + 1] = rad.y; out_radiance[3 * q + 2] = rad.z;
        out_nrays[q] = nr;
        out_term[q] = term;
    }
}

// core.clj:52-57: (sum * (1/nr)) -> sqrt -> * 255.99 -> int(min 255.99 x); row ny-1-j (core.clj:105).
// Java (int) of NaN is 0; negative values truncate toward zero.
void orc_resolve(const double* sum_rgb, int nx, int ny, int nr, uint8_t* rgb8) {
    double inv = 1.0 / nr;
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i)
            for (int ch = 0; ch < 3; ++ch) {
                double x = sum_rgb[((size_t)j * nx + i) * 3 + ch] * inv;
                x = std::sqrt(x);
                x = x * 255.99;
                double m = (x != x) ? x : std::min(255.99, x);  // clojure min propagates NaN
                int v = (m != m) ? 0 : (int)m;
                rgb8[((size_t)(ny - 1 - j) * nx + i) * 3 + ch] = (uint8_t)v;
            }
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
