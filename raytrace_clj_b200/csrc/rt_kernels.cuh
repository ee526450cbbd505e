// rt_kernels.cuh — sm_100a kernels of libraytrace_b200.so.
//
//   Intersect<R,K,BLOCK>   brute-force closest hit (Hitlist.hit?, hitable.clj:15-26) for R rays
//                          per thread: sphere list staged in shared memory as float4 SoA,
//                          warp-uniform LDS.128 broadcast, FP32 cull (17 flop / test), survivors
//                          pushed to per-thread shared-memory lists and refined in FP64.
//   mega_kernel            persistent megakernel with per-lane path regeneration
//                          (pixel + color, core.clj:17-57).
//   trace_kernel           closest hit of caller-given rays (rt_trace_primary).
//   shade_kernel / genrays_kernel   diagnostics for the parity tests.
//   resolve_kernel         core.clj:52-57 + the y flip of core.clj:105.
//   ffma_peak_kernel       FP32 roofline denominator measured on the box.
#pragma once

#include "rt_device.cuh"

namespace rt {

#define RT_FOR_R _Pragma("unroll") for (int r = 0; r < R; ++r)

struct RenderParams {
    DevScene sc;
    DevCamera cam;
    int nx, ny;
    int sample_begin, sample_count;
    int row_offset, row_stride, rows_in_shard;
    int max_depth;
    uint2 key;
    float* sum;                           // nx*ny*3, += ; pixel (i, j) at (j*nx+i)*3, j = 0 bottom
    unsigned long long* counters;         // DC_COUNT
    unsigned long long* work_counter;     // next (sample, pixel) work item
    unsigned long long total_work;        // sample_count * nx * rows_in_shard
    int cull_cap;                         // float4 slots of the shared-memory sphere tile
    int preloaded;                        // 1: whole scene fits one tile (loaded once per CTA)
};

// ------------------------------------------------------------------------------------------
// Intersect: R = 4 rays per thread against the whole scene.
//
// FP32 cull, per (ray, sphere), 11 FP32-pipe instructions = the 17-flop test of SURVEY §8(d)
// (a = d.d is folded into a per-ray normalised direction h = d * sqrt(1+eps)/|d|):
//     f   = o + (-c)                      3 FADD   (movers: + 3 FFMA for c(time) = A + time*B)
//     b   = f . h                         1 FMUL + 2 FFMA
//     nc  = r2i - f . f                   3 FFMA
//     key = b * min(b, 0) + nc            1 FMNMX + 1 FFMA
// key >= 0  <=>  (approaching and discriminant >= 0) or (origin inside the sphere): exactly the
// set of spheres that can have a root in front of the origin; r2i is inflated and h is scaled
// up so rounding can only add false positives.  No branch: the sign bits of the R keys are
// funnel-shifted onto the sphere index, the 16-bit entry (k << 4 | signs) is stored
// unconditionally to the thread's shared-memory list and the list pointer advances only if some
// ray survived.  Survivors are refined in FP64 (refine_candidate) when the list fills and at
// the end of each sphere class.
// ------------------------------------------------------------------------------------------
constexpr int LIST_K = 32;       // entries per thread
constexpr int LIST_GUARD = 2;    // spheres between two overflow checks (= unroll group; small bodies stay in the L0 I-cache)

template <int R, int BLOCK>
struct Intersect {
    static_assert(R == 1 || R == 2 || R == 4, "entry layout holds up to 4 sign bits");
    static constexpr unsigned SIGN_MASK = (1u << R) - 1u;
    float ox[R], oy[R], oz[R];     // origin
    float hx[R], hy[R], hz[R];     // cull direction: d * sqrt(1 + eps) / |d|
    float dx[R], dy[R], dz[R];     // direction as given (un-normalised, util.clj:13-16)
    float tm[R];
    double best_t[R];
    int best_k[R], best_orig[R];
    unsigned ncand;
    double tmin, tmax;

    __device__ __forceinline__ void begin() {
        RT_FOR_R {
            float a = fmaf(dz[r], dz[r], fmaf(dy[r], dy[r], dx[r] * dx[r]));
            float s = rsqrtf(a) * (1.0f + 0.5f * CULL_EPS);
            hx[r] = dx[r] * s; hy[r] = dy[r] * s; hz[r] = dz[r] * s;
            // opaque to the optimiser: otherwise ptxas rematerialises h (RSQ + 4 FMUL) per sphere to save registers
            asm volatile("" : "+f"(hx[r]), "+f"(hy[r]), "+f"(hz[r]));
            best_t[r] = CUDART_INF;
            best_k[r] = -1;
            best_orig[r] = 0x7fffffff;
        }
    }

    // a dead slot: a ray that can never produce a candidate (b > 0 and far outside everything)
    __device__ __forceinline__ void kill(int r) {
        ox[r] = 1e18f; oy[r] = 0.f; oz[r] = 0.f;
        dx[r] = 1.f; dy[r] = 0.f; dz[r] = 0.f;
        tm[r] = 0.f;
    }

    // The list pointer is a byte address in the shared window (one LEA less per sphere than indexing).
    static __device__ __forceinline__ void push_entry(unsigned& ptr, unsigned acc) {
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(ptr), "h"((unsigned short)acc) : "memory");
        if ((~acc) & SIGN_MASK) ptr += BLOCK * 2;   // some ray's key has a clear sign bit: keep the entry
    }
    static __device__ __forceinline__ unsigned list_begin(const uint16_t* list) {
        return (unsigned)__cvta_generic_to_shared(list + threadIdx.x);
    }

    // Refine every listed survivor in FP64.  Entries are (k_local << R | sign bits), ray r at bit R-1-r.
    // Each lane walks its own (ray, sphere) pairs, ONE refine call site: the warp runs
    // max-over-lanes(#pairs) iterations instead of (#entries x R) sparsely populated calls.
    __device__ __forceinline__ void flush(const DevScene& sc, const uint16_t* list, unsigned& ptr, int kbase) {
        const int count = (int)(ptr - list_begin(list)) / (BLOCK * 2);
        int i = 0;
        unsigned e = 0, pend = 0;   // pend: rays of the current entry still to refine (bit r = ray r)
        for (;;) {
            while (pend == 0 && i < count) {
                e = list[i * BLOCK + threadIdx.x];
                ++i;
                pend = (~__brev(e) >> (32 - R)) & SIGN_MASK;   // bit R-1-r of e -> bit r, inverted: 1 = survivor
            }
            if (pend == 0) break;
            const int r = __ffs(pend) - 1;
            pend &= pend - 1;
            const int k = kbase + (int)(e >> R);
            float sox = ox[0], soy = oy[0], soz = oz[0], sdx = dx[0], sdy = dy[0], sdz = dz[0], stm = tm[0];
#pragma unroll
            for (int q = 1; q < R; ++q)
                if (r == q) { sox = ox[q]; soy = oy[q]; soz = oz[q]; sdx = dx[q]; sdy = dy[q]; sdz = dz[q]; stm = tm[q]; }
            const double t = refine_candidate(sc.ex_c0r, sc.ex_c1, sc.ex_t0t1, sc.flags, k, sox, soy, soz, sdx, sdy, sdz,
                                              stm, tmin, tmax);
            ncand++;
            if (t < CUDART_INF) {
                const int orig = __ldg(&sc.orig_id[k]);
#pragma unroll
                for (int q = 0; q < R; ++q)   // exact ties go to the lower caller index (hitable.clj:17-26)
                    if (r == q && (t < best_t[q] || (t == best_t[q] && orig < best_orig[q]))) {
                        best_t[q] = t;
                        best_k[q] = k;
                        best_orig[q] = orig;
                    }
            }
        }
        ptr = list_begin(list);
    }

    __device__ __forceinline__ unsigned key_bits(float fx, float fy, float fz, float r2i, int r) const {
        float b = fmaf(fz, hz[r], fmaf(fy, hy[r], fx * hx[r]));
        float nc = fmaf(-fz, fz, fmaf(-fy, fy, fmaf(-fx, fx, r2i)));
        return __float_as_uint(fmaf(b, fminf(b, 0.f), nc));
    }

    // static spheres: s[k] = (-cx, -cy, -cz, r2_inflated)
    __device__ __forceinline__ void cull_static(const DevScene& sc, const float4* __restrict__ s, int count, int kbase,
                                                uint16_t* list, unsigned& ptr) {
        const unsigned limit = list_begin(list) + (LIST_K - LIST_GUARD) * BLOCK * 2;
        int k = 0;
        for (; k + LIST_GUARD <= count; k += LIST_GUARD) {
#pragma unroll
            for (int u = 0; u < LIST_GUARD; ++u) {
                const float4 S = s[k + u];
                unsigned acc = (unsigned)(k + u);
                RT_FOR_R acc = __funnelshift_l(key_bits(ox[r] + S.x, oy[r] + S.y, oz[r] + S.z, S.w, r), acc, 1);
                push_entry(ptr, acc);
            }
            if (ptr > limit) flush(sc, list, ptr, kbase);
        }
        for (; k < count; ++k) {
            const float4 S = s[k];
            unsigned acc = (unsigned)k;
            RT_FOR_R acc = __funnelshift_l(key_bits(ox[r] + S.x, oy[r] + S.y, oz[r] + S.z, S.w, r), acc, 1);
            push_entry(ptr, acc);
        }
        flush(sc, list, ptr, kbase);
    }

    // moving spheres: -centre(time) = nA + time * nB; sa[k] = (nAx, nAy, nAz, r2_inflated), sb[k] = (nBx, nBy, nBz, 0)
    __device__ __forceinline__ void cull_moving(const DevScene& sc, const float4* __restrict__ sa,
                                                const float4* __restrict__ sb, int count, int kbase, uint16_t* list,
                                                unsigned& ptr) {
        const unsigned limit = list_begin(list) + (LIST_K - LIST_GUARD) * BLOCK * 2;
        int k = 0;
        for (; k + LIST_GUARD <= count; k += LIST_GUARD) {
#pragma unroll
            for (int u = 0; u < LIST_GUARD; ++u) {
                const float4 A = sa[k + u];
                const float4 B = sb[k + u];
                unsigned acc = (unsigned)(k + u);
                RT_FOR_R acc = __funnelshift_l(key_bits(fmaf(tm[r], B.x, ox[r] + A.x), fmaf(tm[r], B.y, oy[r] + A.y),
                                                        fmaf(tm[r], B.z, oz[r] + A.z), A.w, r), acc, 1);
                push_entry(ptr, acc);
            }
            if (ptr > limit) flush(sc, list, ptr, kbase);
        }
        for (; k < count; ++k) {
            const float4 A = sa[k];
            const float4 B = sb[k];
            unsigned acc = (unsigned)k;
            RT_FOR_R acc = __funnelshift_l(key_bits(fmaf(tm[r], B.x, ox[r] + A.x), fmaf(tm[r], B.y, oy[r] + A.y),
                                                    fmaf(tm[r], B.z, oz[r] + A.z), A.w, r), acc, 1);
            push_entry(ptr, acc);
        }
        flush(sc, list, ptr, kbase);
    }

    // Whole scene.  In tiled mode every thread of the CTA must call it (tile loads use __syncthreads).
    // One copy of each hot loop: a preloaded scene is a single resident tile (statics at s_cull[0, ns),
    // movers' A at s_cull + ns, B at s_cull + ns + nm); a tiled scene streams tiles of `cap` float4 slots.
    __device__ __forceinline__ void run(const DevScene& sc, float4* s_cull, int cap, bool preloaded, uint16_t* list) {
        begin();
        unsigned ptr = list_begin(list);
        const int ns = sc.n_static, nm = sc.n_moving;
        for (int base = 0; base < ns; base += cap) {
            int count = min(cap, ns - base);
            if (!preloaded) {
                __syncthreads();
                for (int i = threadIdx.x; i < count; i += BLOCK) s_cull[i] = __ldg(&sc.cull_a[base + i]);
                __syncthreads();
            }
            cull_static(sc, s_cull, count, base, list, ptr);
        }
        const int half = cap / 2;
        const float4* sa = preloaded ? s_cull + ns : s_cull;
        const float4* sb = preloaded ? s_cull + ns + nm : s_cull + half;
        const int step = preloaded ? max(nm, 1) : half;
        for (int base = 0; base < nm; base += step) {
            int count = min(step, nm - base);
            if (!preloaded) {
                __syncthreads();
                for (int i = threadIdx.x; i < count; i += BLOCK) {
                    s_cull[i] = __ldg(&sc.cull_a[ns + base + i]);
                    s_cull[half + i] = __ldg(&sc.cull_b[base + i]);
                }
                __syncthreads();
            }
            cull_moving(sc, sa, sb, count, ns + base, list, ptr);
        }
    }
};

__device__ __forceinline__ void preload_scene(const DevScene& sc, float4* s_cull, int block) {
    for (int i = threadIdx.x; i < sc.n; i += block) s_cull[i] = __ldg(&sc.cull_a[i]);
    for (int i = threadIdx.x; i < sc.n_moving; i += block) s_cull[sc.n + i] = __ldg(&sc.cull_b[i]);
    __syncthreads();
}

// ------------------------------------------------------------------------------------------
// Persistent megakernel with path regeneration: every thread owns R path slots; a slot whose
// path ended pulls the next (sample, pixel) work item (warp-aggregated atomic: ballot + prefix
// popc) so the intersect phase always runs with full lanes until the work runs out.
// ------------------------------------------------------------------------------------------
template <int R, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) mega_kernel(const RenderParams P) {
    extern __shared__ float4 smem_f4[];
    float4* s_cull = smem_f4;
    uint16_t* s_list = reinterpret_cast<uint16_t*>(smem_f4 + P.cull_cap);
    __shared__ unsigned s_ctr[DC_COUNT];
    if (threadIdx.x < DC_COUNT) s_ctr[threadIdx.x] = 0;
    if (P.preloaded) preload_scene(P.sc, s_cull, BLOCK);
    __syncthreads();

    Intersect<R, BLOCK> I;
    I.tmin = 0.001;               // core.clj:25
    I.tmax = (double)FLT_MAX;     // Float/MAX_VALUE
    I.ncand = 0;

    float ar[R], ag[R], ab[R];    // attenuation (core.clj:23 `atten`)
    uint32_t pix[R], smp[R];
    int depth[R];
    bool alive[R];
    RT_FOR_R { alive[r] = false; I.kill(r); }
    unsigned n_rays = 0, n_samples = 0;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned pshard = (unsigned)P.nx * (unsigned)P.rows_in_shard;

    for (;;) {
        // ---- generate: refill dead slots -----------------------------------------------------
        RT_FOR_R {
            unsigned need = __ballot_sync(0xffffffffu, !alive[r]);
            if (need) {
                int leader = __ffs(need) - 1;
                unsigned long long base = 0;
                if ((int)lane == leader) base = atomicAdd(P.work_counter, (unsigned long long)__popc(need));
                base = __shfl_sync(0xffffffffu, base, leader);
                unsigned long long w = base + __popc(need & ((1u << lane) - 1u));
                if (!alive[r] && w < P.total_work) {
                    unsigned s_local = (unsigned)(w / pshard);
                    unsigned q = (unsigned)(w - (unsigned long long)s_local * pshard);
                    int row_local = (int)(q / (unsigned)P.nx);
                    int i = (int)(q - (unsigned)row_local * (unsigned)P.nx);
                    int j = P.row_offset + row_local * P.row_stride;
                    pix[r] = (uint32_t)j * (uint32_t)P.nx + (uint32_t)i;
                    smp[r] = (uint32_t)(P.sample_begin + (int)s_local);
                    float3 o, d;
                    float tmv;
                    generate_ray(P.cam, P.nx, P.ny, i, j, pix[r], smp[r], P.key, o, d, tmv, nullptr);
                    I.ox[r] = o.x; I.oy[r] = o.y; I.oz[r] = o.z;
                    I.dx[r] = d.x; I.dy[r] = d.y; I.dz[r] = d.z;
                    I.tm[r] = tmv;
                    ar[r] = ag[r] = ab[r] = 1.0f;
                    depth[r] = P.max_depth;
                    alive[r] = true;
                    n_samples++;
                }
            }
        }
        bool any_alive = false;
        RT_FOR_R any_alive |= alive[r];
        // preloaded scene: warps are independent (no CTA barrier inside the loop) and leave on their own;
        // tiled scene: the tile loads need every thread of the CTA
        if (P.preloaded) {
            if (!__any_sync(0xffffffffu, any_alive)) break;
        } else {
            if (!__syncthreads_or(any_alive)) break;
        }

        // ---- intersect ----------------------------------------------------------------------------
        RT_FOR_R n_rays += alive[r] ? 1u : 0u;
        I.run(P.sc, s_cull, P.cull_cap, P.preloaded != 0, s_list);

        // ---- shade ------------------------------------------------------------------------------
        RT_FOR_R {
            if (alive[r]) {
                if (I.best_k[r] < 0) {                       // core.clj:40-41 miss -> accum (black)
                    alive[r] = false;
                    I.kill(r);
                    atomicAdd(&s_ctr[DC_TERM_MISS], 1u);
                } else {
                    float3 o = f3(I.ox[r], I.oy[r], I.oz[r]), d = f3(I.dx[r], I.dy[r], I.dz[r]);
                    float3 att, em;
                    int reason = TERM_NONE;
                    ScatterRng rng{P.key, pix[r], smp[r], (uint32_t)(P.max_depth - depth[r] + 1), nullptr, nullptr};
                    bool cont = shade_hit(P.sc, I.best_k[r], (float)I.best_t[r], o, d, I.tm[r], depth[r] > 0, rng, att,
                                          em, reason);
                    if (em.x != 0.f || em.y != 0.f || em.z != 0.f) {   // accum += atten * emitted (core.clj:32-34,37-39)
                        float* dst = P.sum + (size_t)pix[r] * 3;
                        atomicAdd(dst + 0, ar[r] * em.x);
                        atomicAdd(dst + 1, ag[r] * em.y);
                        atomicAdd(dst + 2, ab[r] * em.z);
                    }
                    if (cont) {
                        ar[r] *= att.x; ag[r] *= att.y; ab[r] *= att.z;     // core.clj:31
                        I.ox[r] = o.x; I.oy[r] = o.y; I.oz[r] = o.z;
                        I.dx[r] = d.x; I.dy[r] = d.y; I.dz[r] = d.z;
                        depth[r]--;
                    } else {
                        alive[r] = false;
                        I.kill(r);
                        atomicAdd(&s_ctr[reason == TERM_LIGHT ? DC_TERM_LIGHT
                                         : reason == TERM_ABSORB ? DC_TERM_ABSORB : DC_TERM_DEPTH], 1u);
                    }
                }
            }
        }
    }

    atomicAdd(&s_ctr[DC_RAYS], n_rays);
    atomicAdd(&s_ctr[DC_SAMPLES], n_samples);
    atomicAdd(&s_ctr[DC_CANDIDATES], I.ncand);
    __syncthreads();
    if (threadIdx.x < DC_COUNT && s_ctr[threadIdx.x]) atomicAdd(&P.counters[threadIdx.x], (unsigned long long)s_ctr[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------
// Persistent wavefront path tracer: every CTA is an independent wavefront engine.
//
//   queue record (3 x float4 per path, AoS so one thread moves a path with 3 LDG/STG.128):
//       a = (ox, oy, oz, time)   b = (dx, dy, dz, pixel index)   c = (atten r, g, b, sample << 8 | depth)
//   Each CTA owns two ping-pong queues and a hit buffer of `capacity` entries in HBM/L2 and loops
//       G  generate   fill the queue with camera rays; (sample, pixel) work items are claimed from
//                     one global counter, one atomic per warp (ballot + popc prefix)
//       A  intersect  warps claim batches of 32*R queue entries from a shared-memory counter, run the
//                     brute-force cull + FP64 refine (4 rays per thread), write (t, k) per entry;
//                     a nearly empty queue (tail) switches to the 1-ray-per-thread loop
//       B  shade + regenerate + compact   one path per thread: shade the hit; surviving paths and
//                     fresh camera rays are appended to the NEXT queue at positions claimed with
//                     warp ballot + prefix popc + one shared-memory atomic per warp
//   with __syncthreads between phases.  All warps of a CTA therefore execute the same small code
//   region at the same time (an asynchronous megakernel loses ~45 % of its issue slots to
//   instruction-cache misses, profiles/), intersect always runs on a dense compacted queue, and
//   there is no chip-wide barrier (a grid.sync version spent 41 % of its warp-time waiting):
//   two CTAs per SM hide each other's barrier and shade phases.
// ------------------------------------------------------------------------------------------
struct WaveParams {
    RenderParams base;
    float4* queue;             // gridDim.x * 2 * 3 * capacity float4
    float2* hits;              // gridDim.x * capacity: (t as float, k as int bits)
    int capacity;              // entries per CTA queue (multiple of 32)
};

// claim consecutive slots for the lanes whose predicate is set; returns this lane's slot
__device__ __forceinline__ unsigned warp_claim_shared(unsigned* counter, bool pred, unsigned lane) {
    unsigned m = __ballot_sync(0xffffffffu, pred);
    unsigned base = 0;
    if (m) {
        int leader = __ffs(m) - 1;
        if ((int)lane == leader) base = atomicAdd(counter, (unsigned)__popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
    }
    return base + __popc(m & ((1u << lane) - 1u));
}

// pull one (sample, pixel) work item for every lane with `want`; returns false when the work ran out
__device__ __forceinline__ bool fetch_and_generate(const RenderParams& P, bool want, unsigned lane, float4& a, float4& b,
                                                   float4& c) {
    unsigned m = __ballot_sync(0xffffffffu, want);
    if (!m) return false;
    int leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if ((int)lane == leader) base = atomicAdd(P.work_counter, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    unsigned long long w = base + __popc(m & ((1u << lane) - 1u));
    if (!want || w >= P.total_work) return false;
    const unsigned pshard = (unsigned)P.nx * (unsigned)P.rows_in_shard;
    unsigned s_local = (unsigned)(w / pshard);
    unsigned q = (unsigned)(w - (unsigned long long)s_local * pshard);
    int row_local = (int)(q / (unsigned)P.nx);
    int i = (int)(q - (unsigned)row_local * (unsigned)P.nx);
    int j = P.row_offset + row_local * P.row_stride;
    uint32_t pix = (uint32_t)j * (uint32_t)P.nx + (uint32_t)i;
    uint32_t smp = (uint32_t)(P.sample_begin + (int)s_local);
    float3 o, d;
    float tmv;
    generate_ray(P.cam, P.nx, P.ny, i, j, pix, smp, P.key, o, d, tmv, nullptr);
    a = make_float4(o.x, o.y, o.z, tmv);
    b = make_float4(d.x, d.y, d.z, __uint_as_float(pix));
    c = make_float4(1.f, 1.f, 1.f, __uint_as_float((smp << 8) | (uint32_t)P.max_depth));
    return true;
}

// phase A for one CTA: batches of 32*R entries claimed dynamically (preloaded scene) or walked in
// CTA-uniform order (tiled scene: the tile loads inside Intersect::run need every thread)
template <int R, int BLOCK>
__device__ __forceinline__ void wave_intersect(const RenderParams& P, const float4* qc, float2* hits, unsigned n,
                                               float4* s_cull, uint16_t* s_list, unsigned* s_batch, unsigned& n_cand) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, warps = BLOCK / 32;
    const unsigned n_batches = (n + 32 * R - 1) / (32 * R);
    Intersect<R, BLOCK> I;
    I.tmin = 0.001;               // core.clj:25
    I.tmax = (double)FLT_MAX;     // Float/MAX_VALUE
    I.ncand = 0;
    unsigned batch = warp;
    const unsigned uniform_end = (n_batches + warps - 1) / warps * warps;
    for (;;) {
        if (P.preloaded) {
            unsigned b = 0;
            if (lane == 0) b = atomicAdd(s_batch, 1u);
            batch = __shfl_sync(0xffffffffu, b, 0);
            if (batch >= n_batches) break;
        } else {
            if (batch >= uniform_end) break;
        }
        const unsigned wbase = batch * (32 * R) + lane;     // ray r of this lane = entry wbase + 32 r
        RT_FOR_R {
            unsigned idx = wbase + 32 * r;
            if (idx < n) {
                float4 a = qc[3 * (size_t)idx], b = qc[3 * (size_t)idx + 1];
                I.ox[r] = a.x; I.oy[r] = a.y; I.oz[r] = a.z; I.tm[r] = a.w;
                I.dx[r] = b.x; I.dy[r] = b.y; I.dz[r] = b.z;
            } else {
                I.kill(r);
            }
        }
        I.run(P.sc, s_cull, P.cull_cap, P.preloaded != 0, s_list);
        RT_FOR_R {
            unsigned idx = wbase + 32 * r;
            if (idx < n) hits[idx] = make_float2((float)I.best_t[r], __int_as_float(I.best_k[r]));
        }
        batch += warps;
    }
    n_cand += I.ncand;
}

template <int R, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) wave_kernel(const WaveParams W) {
    const RenderParams& P = W.base;
    extern __shared__ float4 smem_f4[];
    float4* s_cull = smem_f4;
    uint16_t* s_list = reinterpret_cast<uint16_t*>(smem_f4 + P.cull_cap);
    __shared__ unsigned s_ctr[DC_COUNT];
    __shared__ unsigned s_qcount[2];
    __shared__ unsigned s_batch;
    if (threadIdx.x < DC_COUNT) s_ctr[threadIdx.x] = 0;
    if (threadIdx.x < 2) s_qcount[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_batch = 0;
    if (P.preloaded) preload_scene(P.sc, s_cull, BLOCK);
    __syncthreads();

    const unsigned lane = threadIdx.x & 31u;
    const unsigned cap = (unsigned)W.capacity;
    float4* q0 = W.queue + (size_t)blockIdx.x * 2 * 3 * cap;
    float4* queue[2] = {q0, q0 + 3 * (size_t)cap};
    float2* hits = W.hits + (size_t)blockIdx.x * cap;
    unsigned n_rays = 0, n_samples = 0, n_cand = 0;

    // ---- phase G: fill queue 0 -------------------------------------------------------------------
    for (unsigned idx = threadIdx.x; idx < cap; idx += BLOCK) {
        float4 a, b, c;
        bool ok = fetch_and_generate(P, true, lane, a, b, c);
        unsigned slot = warp_claim_shared(&s_qcount[0], ok, lane);
        if (ok) {
            float4* q = queue[0] + 3 * (size_t)slot;
            q[0] = a; q[1] = b; q[2] = c;
            n_samples++;
        }
        if (!__any_sync(0xffffffffu, ok)) break;          // work ran out
    }
    __syncthreads();

    int cur = 0;
    for (;;) {
        const unsigned n = s_qcount[cur];
        if (n == 0) break;
        const float4* qc = queue[cur];
        if (threadIdx.x == 0) n_rays += n;

        // ---- phase A: intersect ----------------------------------------------------------------------
        if (n > BLOCK) wave_intersect<R, BLOCK>(P, qc, hits, n, s_cull, s_list, &s_batch, n_cand);
        else           wave_intersect<1, BLOCK>(P, qc, hits, n, s_cull, s_list, &s_batch, n_cand);   // tail: 1 ray/thread
        __syncthreads();
        if (threadIdx.x == 0) { s_batch = 0; s_qcount[cur] = 0; }

        // ---- phase B: shade, compact survivors into the next queue, regenerate -------------------
        float4* qn = queue[cur ^ 1];
        for (unsigned idx0 = threadIdx.x - lane; idx0 < n; idx0 += BLOCK) {   // warp-uniform trip count
            const unsigned idx = idx0 + lane;
            bool have = idx < n;
            bool cont = false;
            float4 a, b, c;
            if (have) {
                a = qc[3 * (size_t)idx]; b = qc[3 * (size_t)idx + 1]; c = qc[3 * (size_t)idx + 2];
                float2 h = hits[idx];
                int k = __float_as_int(h.y);
                uint32_t pix = __float_as_uint(b.w), sd = __float_as_uint(c.w);
                uint32_t smp = sd >> 8;
                int depth = (int)(sd & 255u);
                if (k < 0) {                                        // core.clj:40-41 miss -> accum (black)
                    atomicAdd(&s_ctr[DC_TERM_MISS], 1u);
                } else {
                    float3 o = f3(a.x, a.y, a.z), d = f3(b.x, b.y, b.z);
                    float3 att, em;
                    int reason = TERM_NONE;
                    ScatterRng rng{P.key, pix, smp, (uint32_t)(P.max_depth - depth + 1), nullptr, nullptr};
                    cont = shade_hit(P.sc, k, h.x, o, d, a.w, depth > 0, rng, att, em, reason);
                    if (em.x != 0.f || em.y != 0.f || em.z != 0.f) {   // accum += atten * emitted (core.clj:32-34,37-39)
                        float* dst = P.sum + (size_t)pix * 3;
                        atomicAdd(dst + 0, c.x * em.x);
                        atomicAdd(dst + 1, c.y * em.y);
                        atomicAdd(dst + 2, c.z * em.z);
                    }
                    if (cont) {
                        a = make_float4(o.x, o.y, o.z, a.w);
                        b = make_float4(d.x, d.y, d.z, b.w);
                        c = make_float4(c.x * att.x, c.y * att.y, c.z * att.z, __uint_as_float((smp << 8) | (uint32_t)(depth - 1)));
                    } else {
                        atomicAdd(&s_ctr[reason == TERM_LIGHT ? DC_TERM_LIGHT
                                         : reason == TERM_ABSORB ? DC_TERM_ABSORB : DC_TERM_DEPTH], 1u);
                    }
                }
            }
            // a finished path frees its lane for the next (sample, pixel)
            if (fetch_and_generate(P, have && !cont, lane, a, b, c)) {
                cont = true;
                n_samples++;
            }
            unsigned slot = warp_claim_shared(&s_qcount[cur ^ 1], cont, lane);
            if (cont) {
                float4* q = qn + 3 * (size_t)slot;
                q[0] = a; q[1] = b; q[2] = c;
            }
        }
        __syncthreads();
        cur ^= 1;
    }

    atomicAdd(&s_ctr[DC_RAYS], n_rays);
    atomicAdd(&s_ctr[DC_SAMPLES], n_samples);
    atomicAdd(&s_ctr[DC_CANDIDATES], n_cand);
    __syncthreads();
    if (threadIdx.x < DC_COUNT && s_ctr[threadIdx.x]) atomicAdd(&P.counters[threadIdx.x], (unsigned long long)s_ctr[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------
// rt_trace_primary: closest hit of n caller-given rays (same Intersect as the renderer)
// ------------------------------------------------------------------------------------------
struct TraceParams {
    DevScene sc;
    int n;
    const float* origins;
    const float* dirs;
    const float* times;   // may be null
    double tmin, tmax;
    double* out_t;
    int* out_id;
    int cull_cap, preloaded;
};

template <int R, int BLOCK>
__global__ void __launch_bounds__(BLOCK) trace_kernel(const TraceParams P) {
    extern __shared__ float4 smem_f4[];
    float4* s_cull = smem_f4;
    uint16_t* s_list = reinterpret_cast<uint16_t*>(smem_f4 + P.cull_cap);
    if (P.preloaded) preload_scene(P.sc, s_cull, BLOCK);
    __syncthreads();
    Intersect<R, BLOCK> I;
    I.tmin = P.tmin;
    I.tmax = P.tmax;
    I.ncand = 0;
    for (long long base = (long long)blockIdx.x * BLOCK * R; base < P.n; base += (long long)gridDim.x * BLOCK * R) {
        RT_FOR_R {
            long long idx = base + (long long)r * BLOCK + threadIdx.x;
            if (idx < P.n) {
                I.ox[r] = P.origins[3 * idx]; I.oy[r] = P.origins[3 * idx + 1]; I.oz[r] = P.origins[3 * idx + 2];
                I.dx[r] = P.dirs[3 * idx]; I.dy[r] = P.dirs[3 * idx + 1]; I.dz[r] = P.dirs[3 * idx + 2];
                I.tm[r] = P.times ? P.times[idx] : 0.f;
            } else {
                I.kill(r);
            }
        }
        I.run(P.sc, s_cull, P.cull_cap, P.preloaded != 0, s_list);
        RT_FOR_R {
            long long idx = base + (long long)r * BLOCK + threadIdx.x;
            if (idx < P.n) {
                P.out_t[idx] = I.best_t[r];
                P.out_id[idx] = (I.best_k[r] >= 0) ? I.best_orig[r] : -1;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// diagnostics
// ------------------------------------------------------------------------------------------
__global__ void genrays_kernel(DevCamera cam, int n, int nx, int ny, const int* ij, const int* s, uint2 key, float* out_o,
                               float* out_d, float* out_t, float* out_rnd) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    int i = ij[2 * idx], j = ij[2 * idx + 1];
    float3 o, d;
    float tm;
    float rnd[5];
    generate_ray(cam, nx, ny, i, j, (uint32_t)j * (uint32_t)nx + (uint32_t)i, (uint32_t)s[idx], key, o, d, tm, rnd);
    out_o[3 * idx] = o.x; out_o[3 * idx + 1] = o.y; out_o[3 * idx + 2] = o.z;
    out_d[3 * idx] = d.x; out_d[3 * idx + 1] = d.y; out_d[3 * idx + 2] = d.z;
    out_t[idx] = tm;
    if (out_rnd)
        for (int k = 0; k < 5; ++k) out_rnd[5 * idx + k] = rnd[k];
}

__global__ void shade_kernel(DevScene sc, int n, const float* origins, const float* dirs, const float* times,
                             const int* hit_id, const double* hit_t, const float* ball, const float* u01v, float* out_o,
                             float* out_d, float* out_att, float* out_em, int* out_flags) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    float3 o = f3(origins[3 * idx], origins[3 * idx + 1], origins[3 * idx + 2]);
    float3 d = f3(dirs[3 * idx], dirs[3 * idx + 1], dirs[3 * idx + 2]);
    float3 att = f3(0.f, 0.f, 0.f), em = f3(0.f, 0.f, 0.f);
    int flag = -1;
    int id = hit_id[idx];
    if (id >= 0 && id < sc.n) {
        int k = sc.cull_of_orig[id];
        int reason = TERM_NONE;
        ScatterRng rng{make_uint2(0u, 0u), 0u, 0u, 0u, ball + 3 * idx, u01v + idx};
        bool cont = shade_hit(sc, k, (float)hit_t[idx], o, d, times ? times[idx] : 0.f, true, rng, att, em, reason);
        flag = cont ? 1 : 0;
        if (!cont) {
            o = f3(0.f, 0.f, 0.f); d = o; att = o;
        }
    } else {
        o = f3(0.f, 0.f, 0.f); d = o;
    }
    out_o[3 * idx] = o.x; out_o[3 * idx + 1] = o.y; out_o[3 * idx + 2] = o.z;
    out_d[3 * idx] = d.x; out_d[3 * idx + 1] = d.y; out_d[3 * idx + 2] = d.z;
    out_att[3 * idx] = att.x; out_att[3 * idx + 1] = att.y; out_att[3 * idx + 2] = att.z;
    out_em[3 * idx] = em.x; out_em[3 * idx + 1] = em.y; out_em[3 * idx + 2] = em.z;
    out_flags[idx] = flag;
}

// core.clj:52-57: (sum * (1/nr)) -> sqrt -> * 255.99 -> (int (min 255.99 x)); row ny-1-j (core.clj:105).
// Evaluated in double from the float sums so it matches the oracle's resolve bit for bit.
// `peers` (optional): other devices' sum buffers mapped over NVLink peer access, added first.
struct ResolveParams {
    const float* sum;
    const float* peers[7];
    int n_peers;
    int nx, ny, nr;
    uint8_t* rgb8;       // may be null
    float* mean;         // may be null: (float)(sum / nr), same layout as sum
    float* sum_out;      // may be null: reduced sum written back (multi-device)
};
__global__ void resolve_kernel(const ResolveParams P) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;   // over nx*ny*3
    int total = P.nx * P.ny * 3;
    if (idx >= total) return;
    float s = P.sum[idx];
    for (int p = 0; p < P.n_peers; ++p) s += P.peers[p][idx];
    if (P.sum_out) P.sum_out[idx] = s;
    double x = (double)s * (1.0 / (double)P.nr);
    if (P.mean) P.mean[idx] = (float)x;
    if (P.rgb8) {
        int ch = idx % 3, pixel = idx / 3;
        int i = pixel % P.nx, j = pixel / P.nx;
        double y = sqrt(x) * 255.99;
        int v = (y != y) ? 0 : (int)fmin(255.99, y);   // (int NaN) = 0 on the JVM
        P.rgb8[((size_t)(P.ny - 1 - j) * P.nx + i) * 3 + ch] = (uint8_t)v;
    }
}

// FP32 peak: independent FFMA chains (and the packed FFMA2 form), no memory traffic.
template <bool PACKED>
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, int iters, float a, float b) {
    if (PACKED) {
        float2 x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
        float2 aa = make_float2(a, a), bb = make_float2(b, b);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = __ffma2_rn(x[i], aa, bb);
        }
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += x[i].x + x[i].y;
        out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    } else {
        float x[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3f + i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
        }
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) acc += x[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    }
}

}  // namespace rt
