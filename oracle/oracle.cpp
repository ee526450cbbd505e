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
// lerp; util_test.clj:44-49 point-at-parameter) and (2) closed-form values derived
// from the reference formulas.  Numeric t / p / normal / uv, scatter, emitted,
// sample, get-ray, color, pixel are "parity unpinned" by the reference itself
// (it has no such tests) — fidelity there is by inspection, each function below
// citing the reference lines it follows.
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

struct Ray {
    V3 o, d;
    double time;
};

// util.clj:18-22  point-at-parameter = direction * t + origin
inline V3 point_at_parameter(const Ray& r, double t) { return add(mul(t, r.d), r.o); }

// ---- RNG: stands in for clojure.core/rand (Math.random); the JVM stream is unseeded and
// cannot be reproduced, so the oracle uses xoshiro256** seeded per (seed, pixel, sample).
struct Rng {
    uint64_t s[4];
    static uint64_t splitmix(uint64_t& x) {
        uint64_t z = (x += 0x9e3779b97f4a7c15ULL);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        return z ^ (z >> 31);
    }
    explicit Rng(uint64_t seed) {
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

// util.clj:32-41  rand-in-unit-disk: rejection in [-1,1)^2 x {0}, accept when dot < 1
inline V3 rand_in_unit_disk(Rng& g) {
    for (;;) {
        double x = 2.0 * g.rand() - 1.0;
        double y = 2.0 * g.rand() - 1.0;
        V3 p = v3(x, y, 0);
        if (!(dot(p, p) >= 1.0)) return p;
    }
}
// util.clj:43-52  rand-in-unit-sphere
inline V3 rand_in_unit_sphere(Rng& g) {
    for (;;) {
        double x = 2.0 * g.rand() - 1.0;
        double y = 2.0 * g.rand() - 1.0;
        double z = 2.0 * g.rand() - 1.0;
        V3 p = v3(x, y, z);
        if (!(dot(p, p) >= 1.0)) return p;
    }
}

// ---- scene tables (the marshalled form of the reference's record graph) ----------------
enum { MAT_LAMBERTIAN = 0, MAT_METAL = 1, MAT_DIELECTRIC = 2, MAT_DIFFUSE_LIGHT = 3 };
enum { TEX_CONSTANT = 0, TEX_UV_GRADIENT = 1, TEX_CHECKERBOARD = 2 };
enum { SPH_UV = 1, SPH_MOVING = 2 };
enum { CAM_PINHOLE = 0, CAM_THIN_LENS = 1 };

struct Sphere {
    V3 c0, c1;
    double r, t0, t1;
    uint32_t flags;
    int mat;
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
struct Scene {
    std::vector<Sphere> spheres;
    std::vector<Material> mats;
    std::vector<Texture> texs;
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
inline bool sphere_hit(const Sphere& s, int id, const Ray& r, double t_min, double t_max, HitRec& h) {
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
        h.mat = s.mat;
        h.id = id;
        return true;
    }
    return false;
}

// hitable.clj:15-26  Hitlist.hit?: reduce with shrinking t-max, first item wins ties
inline bool world_hit(const Scene& sc, const Ray& r, double t_min, double t_max, HitRec& out) {
    bool any = false;
    double closest = t_max;
    HitRec h;
    const int n = (int)sc.spheres.size();
    for (int i = 0; i < n; ++i) {
        if (sphere_hit(sc.spheres[i], i, r, t_min, closest, h)) {
            any = true;
            closest = h.t;
            out = h;
        }
    }
    return any;
}

// texture.clj:14-50  sample
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
    const double* ball;  // explicit rand-in-unit-sphere result, or null
    const double* u01;   // explicit (rand), or null
    V3 unit_sphere() { return ball ? v3(ball[0], ball[1], ball[2]) : rand_in_unit_sphere(*g); }
    double rand() { return u01 ? *u01 : g->rand(); }
};

// shader.clj:29-36, 46-59, 76-104, 114-119  scatter; returns false for nil
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
    V3 rd = mul(lens_radius, rand_in_unit_disk(g));
    V3 offset = add(mul(rd.x, c.u), mul(rd.y, c.v));
    V3 o = add(c.origin, offset);
    V3 d = add(add(add(add(c.lleft, mul(s, c.horiz)), mul(t, c.vert)), neg(c.origin)), neg(offset));
    double time = c.t0 + (c.t1 - c.t0) * g.rand();
    return Ray{o, d, time};
}

struct Counters {
    uint64_t rays = 0, samples = 0, term_light = 0, term_absorb = 0, term_depth = 0, term_miss = 0;
};

// core.clj:17-41  color: iterative loop, depth cutoff, t-range (0.001, Float/MAX_VALUE), miss -> black
inline V3 color(const Scene& sc, Ray r, int depth, Rng& g, Counters& ctr) {
    V3 atten = v3(1, 1, 1), accum = v3(0, 0, 0);
    for (;;) {
        ctr.rays++;
        HitRec h;
        if (world_hit(sc, r, 0.001, (double)FLT_MAX, h)) {
            Ray scattered;
            V3 attenuation;
            int why = 0;
            ScatterRand rr{&g, nullptr, nullptr};
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
                return add(accum, mul(atten, e));
            }
        } else {
            ctr.term_miss++;
            return accum;
        }
    }
}

inline uint64_t mix_seed(uint64_t seed, uint64_t pixel, uint64_t sample) {
    uint64_t x = seed * 0x9e3779b97f4a7c15ULL + pixel;
    x = Rng::splitmix(x);
    x ^= sample * 0xd1342543de82ef95ULL;
    return Rng::splitmix(x);
}

Scene* build_scene(int n, const float* c0r, const float* c1, const float* t0t1, const uint32_t* flags,
                   const int32_t* mat_id, int nm, const int32_t* mtype, const float* mparam,
                   const int32_t* mtex, int nt, const int32_t* ttype, const float* tparams,
                   const int32_t* tchild) {
    Scene* sc = new Scene();
    sc->spheres.resize(n);
    for (int i = 0; i < n; ++i) {
        Sphere& s = sc->spheres[i];
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
        s.mat = mat_id[i];
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
    return sc;
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
    Sphere s;
    s.c0 = v3(c0[0], c0[1], c0[2]);
    s.c1 = c1 ? v3(c1[0], c1[1], c1[2]) : s.c0;
    s.t0 = t0; s.t1 = t1; s.r = radius; s.flags = flags; s.mat = 0;
    Ray r{v3(o[0], o[1], o[2]), v3(d[0], d[1], d[2]), time};
    HitRec h;
    if (!sphere_hit(s, 0, r, t_min, t_max, h)) return 0;
    if (out_t) *out_t = h.t;
    if (out_p) { out_p[0] = h.p.x; out_p[1] = h.p.y; out_p[2] = h.p.z; }
    if (out_n) { out_n[0] = h.normal.x; out_n[1] = h.normal.y; out_n[2] = h.normal.z; }
    if (out_uv) { out_uv[0] = h.uv[0]; out_uv[1] = h.uv[1]; }
    return 1;
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

// ---- scene handle ---------------------------------------------------------------------------
void* orc_scene_create(int n, const float* c0r, const float* c1, const float* t0t1, const uint32_t* flags,
                       const int32_t* mat_id, int nm, const int32_t* mtype, const float* mparam,
                       const int32_t* mtex, int nt, const int32_t* ttype, const float* tparams,
                       const int32_t* tchild) {
    return build_scene(n, c0r, c1, t0t1, flags, mat_id, nm, mtype, mparam, mtex, nt, ttype, tparams, tchild);
}
void orc_scene_destroy(void* sc) { delete (Scene*)sc; }

// Hitlist.hit? for n rays (float inputs promoted to double): out_t, out_id (-1 = miss);
// optional out_t2 = second-smallest valid t over the other spheres (inf if none), for the
// "two best within 1e-5" id-ambiguity rule of the parity plan.
void orc_hit(void* scene, int n, const float* origins, const float* dirs, const float* times, double t_min,
             double t_max, double* out_t, int32_t* out_id, double* out_t2) {
    const Scene& sc = *(Scene*)scene;
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; ++i) {
        Ray r{v3(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]), v3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]),
              times ? (double)times[i] : 0.0};
        HitRec h;
        if (world_hit(sc, r, t_min, t_max, h)) {
            out_t[i] = h.t;
            out_id[i] = h.id;
        } else {
            out_t[i] = INFINITY;
            out_id[i] = -1;
        }
        if (out_t2) {
            double best2 = INFINITY;
            for (int k = 0; k < (int)sc.spheres.size(); ++k) {
                if (k == out_id[i]) continue;
                HitRec h2;
                if (sphere_hit(sc.spheres[k], k, r, t_min, t_max, h2) && h2.t < best2) best2 = h2.t;
            }
            out_t2[i] = best2;
        }
    }
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
        if (hit_id[i] < 0 || !sphere_hit(sc.spheres[hit_id[i]], hit_id[i], r, 0.001, (double)FLT_MAX, h)) continue;
        if (out_t) out_t[i] = h.t;
        double b[3] = {ball[3 * i], ball[3 * i + 1], ball[3 * i + 2]};
        double u = u01[i];
        ScatterRand rr{nullptr, b, &u};
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

// core.clj:43-57 pixel (the sample loop and the sum) for every pixel of rows j % row_stride == row_offset,
// samples [s_begin, s_begin + s_count): sum_rgb[((j*nx)+i)*3 + c] += SUM of color (double, j = 0 bottom row).
// counters[8]: rays, sphere_tests, samples, term_light, term_absorb, term_depth, term_miss, 0.
void orc_render_accumulate(void* scene, int cam_type, const float cam[24], int nx, int ny, int s_begin, int s_count,
                           int row_offset, int row_stride, int max_depth, uint64_t seed, double* sum_rgb,
                           uint64_t counters[8], int n_threads) {
    const Scene& sc = *(Scene*)scene;
    Camera c = make_camera(cam_type, cam);
    Counters total;
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
                Rng g(mix_seed(seed, (uint64_t)pix, (uint64_t)s));
                // core.clj:49-50: u drawn first, then v; (float i) + rand, divided by nx
                double u = ((double)(float)i + g.rand()) / nx;
                double v = ((double)(float)j + g.rand()) / ny;
                Ray r = get_ray(c, u, v, g);
                ctr.samples++;
                sum = add(sum, color(sc, r, max_depth, g, ctr));
            }
            sum_rgb[3 * pix] += sum.x;
            sum_rgb[3 * pix + 1] += sum.y;
            sum_rgb[3 * pix + 2] += sum.z;
        }
#pragma omp critical
        {
            total.rays += ctr.rays; total.samples += ctr.samples; total.term_light += ctr.term_light;
            total.term_absorb += ctr.term_absorb; total.term_depth += ctr.term_depth; total.term_miss += ctr.term_miss;
        }
    }
    if (counters) {
        counters[0] += total.rays;
        counters[1] += total.rays * (uint64_t)sc.spheres.size();
        counters[2] += total.samples;
        counters[3] += total.term_light;
        counters[4] += total.term_absorb;
        counters[5] += total.term_depth;
        counters[6] += total.term_miss;
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
