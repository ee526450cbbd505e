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
// Intersect: R rays per thread against the whole scene.
// ------------------------------------------------------------------------------------------
template <int R, int K, int BLOCK>
struct Intersect {
    float ox[R], oy[R], oz[R], dx[R], dy[R], dz[R], tm[R];
    float ap[R];          // a' = (1 - CULL_EPS) * d.d, hoisted per ray
    double best_t[R];
    int best_k[R], best_orig[R];
    int cnt[R];
    unsigned ncand;
    double tmin, tmax;

    __device__ __forceinline__ void begin() {
        RT_FOR_R {
            ap[r] = (1.0f - CULL_EPS) * fmaf(dz[r], dz[r], fmaf(dy[r], dy[r], dx[r] * dx[r]));
            best_t[r] = CUDART_INF;
            best_k[r] = -1;
            best_orig[r] = 0x7fffffff;
            cnt[r] = 0;
        }
    }

    template <int RR>
    __device__ __forceinline__ void flush(const DevScene& sc, const uint32_t* cand) {
        for (int i = 0; i < cnt[RR]; ++i) {
            int k = (int)cand[(RR * K + i) * BLOCK + threadIdx.x];
            double t = refine_candidate(sc.ex_c0r, sc.ex_c1, sc.ex_t0t1, sc.flags, k, ox[RR], oy[RR], oz[RR], dx[RR],
                                        dy[RR], dz[RR], tm[RR], tmin, tmax);
            if (t <= best_t[RR] && t < CUDART_INF) {   // exact ties go to the lower caller index (hitable.clj:17-26)
                int orig = __ldg(&sc.orig_id[k]);
                if (t < best_t[RR] || orig < best_orig[RR]) {
                    best_t[RR] = t;
                    best_k[RR] = k;
                    best_orig[RR] = orig;
                }
            }
        }
        ncand += cnt[RR];
        cnt[RR] = 0;
    }

    template <int RR>
    __device__ __forceinline__ void push(const DevScene& sc, uint32_t* cand, int k) {
        cand[(RR * K + cnt[RR]) * BLOCK + threadIdx.x] = (uint32_t)k;
        if (++cnt[RR] == K) flush<RR>(sc, cand);
    }

    // survivors of the FP32 test; both roots negative (centre behind, origin outside) are dropped here
    template <int RR>
    __device__ __forceinline__ void consider(const DevScene& sc, uint32_t* cand, int k, const float (&b)[R],
                                             const float (&c)[R], const float (&disc)[R]) {
        if (disc[RR] >= 0.f && !(b[RR] > 0.f && c[RR] > 0.f)) push<RR>(sc, cand, k);
        if constexpr (RR + 1 < R) consider<RR + 1>(sc, cand, k, b, c, disc);
    }

    // static spheres: s[k] = (cx, cy, cz, r2_inflated)
    __device__ __forceinline__ void cull_static(const DevScene& sc, const float4* __restrict__ s, int count, int kbase,
                                                uint32_t* cand) {
#pragma unroll 2
        for (int k = 0; k < count; ++k) {
            const float4 S = s[k];
            float b[R], c[R], disc[R];
            bool any = false;
            RT_FOR_R {
                float fx = ox[r] - S.x, fy = oy[r] - S.y, fz = oz[r] - S.z;           // oc = o - c        (3)
                b[r] = fmaf(fz, dz[r], fmaf(fy, dy[r], fx * dx[r]));                  // oc.d              (5)
                c[r] = fmaf(fz, fz, fmaf(fy, fy, fmaf(fx, fx, -S.w)));                // oc.oc - r^2       (6)
                disc[r] = fmaf(-ap[r], c[r], b[r] * b[r]);                            // b'^2 - a c'       (3)
                any |= (disc[r] >= 0.f);
            }
            if (any) consider<0>(sc, cand, kbase + k, b, c, disc);
        }
    }

    // moving spheres: centre(time) = A + time * B
    __device__ __forceinline__ void cull_moving(const DevScene& sc, const float4* __restrict__ sa,
                                                const float4* __restrict__ sb, int count, int kbase, uint32_t* cand) {
#pragma unroll 2
        for (int k = 0; k < count; ++k) {
            const float4 A = sa[k];
            const float4 B = sb[k];
            float b[R], c[R], disc[R];
            bool any = false;
            RT_FOR_R {
                float fx = ox[r] - fmaf(tm[r], B.x, A.x);
                float fy = oy[r] - fmaf(tm[r], B.y, A.y);
                float fz = oz[r] - fmaf(tm[r], B.z, A.z);
                b[r] = fmaf(fz, dz[r], fmaf(fy, dy[r], fx * dx[r]));
                c[r] = fmaf(fz, fz, fmaf(fy, fy, fmaf(fx, fx, -A.w)));
                disc[r] = fmaf(-ap[r], c[r], b[r] * b[r]);
                any |= (disc[r] >= 0.f);
            }
            if (any) consider<0>(sc, cand, kbase + k, b, c, disc);
        }
    }

    template <int RR>
    __device__ __forceinline__ void flush_all(const DevScene& sc, const uint32_t* cand) {
        flush<RR>(sc, cand);
        if constexpr (RR + 1 < R) flush_all<RR + 1>(sc, cand);
    }

    // Whole scene.  Must be called by every thread of the CTA (tile loads use __syncthreads).
    __device__ __forceinline__ void run(const DevScene& sc, float4* s_cull, int cap, bool preloaded, uint32_t* cand) {
        begin();
        const int ns = sc.n_static, nm = sc.n_moving;
        if (preloaded) {
            cull_static(sc, s_cull, ns, 0, cand);
            cull_moving(sc, s_cull + ns, s_cull + ns + nm, nm, ns, cand);
        } else {
            for (int base = 0; base < ns; base += cap) {
                int count = min(cap, ns - base);
                __syncthreads();
                for (int i = threadIdx.x; i < count; i += BLOCK) s_cull[i] = __ldg(&sc.cull_a[base + i]);
                __syncthreads();
                cull_static(sc, s_cull, count, base, cand);
            }
            const int half = cap / 2;
            for (int base = 0; base < nm; base += half) {
                int count = min(half, nm - base);
                __syncthreads();
                for (int i = threadIdx.x; i < count; i += BLOCK) {
                    s_cull[i] = __ldg(&sc.cull_a[ns + base + i]);
                    s_cull[half + i] = __ldg(&sc.cull_b[base + i]);
                }
                __syncthreads();
                cull_moving(sc, s_cull, s_cull + half, count, ns + base, cand);
            }
        }
        flush_all<0>(sc, cand);
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
template <int R, int K, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) mega_kernel(const RenderParams P) {
    extern __shared__ float4 smem_f4[];
    float4* s_cull = smem_f4;
    uint32_t* s_cand = reinterpret_cast<uint32_t*>(smem_f4 + P.cull_cap);
    __shared__ unsigned s_ctr[DC_COUNT];
    if (threadIdx.x < DC_COUNT) s_ctr[threadIdx.x] = 0;
    if (P.preloaded) preload_scene(P.sc, s_cull, BLOCK);
    __syncthreads();

    Intersect<R, K, BLOCK> I;
    I.tmin = 0.001;               // core.clj:25
    I.tmax = (double)FLT_MAX;     // Float/MAX_VALUE
    I.ncand = 0;

    float ar[R], ag[R], ab[R];    // attenuation (core.clj:23 `atten`)
    uint32_t pix[R], smp[R];
    int depth[R];
    bool alive[R];
    RT_FOR_R alive[r] = false;
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
        if (!__syncthreads_or(any_alive)) break;

        // ---- intersect: uniform over the CTA ---------------------------------------------------
        RT_FOR_R {
            if (!alive[r]) I.ox[r] = __int_as_float(0x7fc00000);   // NaN origin never passes the cull
            else n_rays++;
        }
        I.run(P.sc, s_cull, P.cull_cap, P.preloaded != 0, s_cand);

        // ---- shade ------------------------------------------------------------------------------
        RT_FOR_R {
            if (alive[r]) {
                if (I.best_k[r] < 0) {                       // core.clj:40-41 miss -> accum (black)
                    alive[r] = false;
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

template <int R, int K, int BLOCK>
__global__ void __launch_bounds__(BLOCK) trace_kernel(const TraceParams P) {
    extern __shared__ float4 smem_f4[];
    float4* s_cull = smem_f4;
    uint32_t* s_cand = reinterpret_cast<uint32_t*>(smem_f4 + P.cull_cap);
    if (P.preloaded) preload_scene(P.sc, s_cull, BLOCK);
    __syncthreads();
    Intersect<R, K, BLOCK> I;
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
                I.ox[r] = __int_as_float(0x7fc00000); I.oy[r] = I.oz[r] = 0.f;
                I.dx[r] = I.dy[r] = I.dz[r] = 0.f; I.tm[r] = 0.f;
            }
        }
        I.run(P.sc, s_cull, P.cull_cap, P.preloaded != 0, s_cand);
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
