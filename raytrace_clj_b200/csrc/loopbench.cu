// loopbench.cu — the cull hot loop in isolation (DESIGN.md "Kernels"): which ingredient costs what.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o loopbench loopbench.cu
//   FORM 0: translated form  f = o - c; b = f.h; nc = r2 - f.f            (11 FP32 instr / test, round-1 first version)
//   FORM 1: expanded form    b = P - c.h; s = W + 2 o.c; nc = s - Q       ( 9 FP32 instr / test, shipped)
//   FORM 2: FORM 1 with b min(b,0) written c c - c |c| (no FMNMX: the ALU pipe only sees the funnel shift)
//   FORM 3: all rays share the origin: the sphere's record carries (W + 2 o.c) - |o|^2, 5 FFMA per test (shipped for camera rays)
//   MODE 0: math only (keys summed)   1: + sign funnel, one mask word per 32 spheres per ray (shipped)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define NS 488
#define R 4
__device__ __forceinline__ float4 lds128(unsigned addr) {
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
template <int FORM, int MODE, int BLOCK, int UNROLL>
__global__ void __launch_bounds__(BLOCK) k(const float4* __restrict__ spheres, float* out, int passes) {
    __shared__ float4 s[NS];
    __shared__ uint32_t list[16 * R * BLOCK];
    for (int i = threadIdx.x; i < NS; i += BLOCK) s[i] = spheres[i];
    __syncthreads();
    float ax[R], ay[R], az[R], hx[R], hy[R], hz[R], P[R], Q[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        ax[r] = 0.1f * threadIdx.x + r; ay[r] = 1.0f + r; az[r] = -2.0f * r;
        hx[r] = 0.3f + 0.01f * r; hy[r] = -0.5f; hz[r] = 0.8f; P[r] = 0.7f * r; Q[r] = 3.0f + r;
        asm volatile("" : "+f"(ax[r]), "+f"(ay[r]), "+f"(az[r]), "+f"(hx[r]), "+f"(hy[r]), "+f"(hz[r]), "+f"(P[r]), "+f"(Q[r]));
    }
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(s);
    const unsigned lbase = (unsigned)__cvta_generic_to_shared(list + threadIdx.x);
    float acc_f = 0.f;
    unsigned total = 0;
    for (int p = 0; p < passes; ++p) {
        unsigned sa = sbase, la = lbase;
        for (int g = 0; g < NS / 32; ++g, la += R * BLOCK * 4) {
            unsigned acc[R];
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = 0u;
#pragma unroll UNROLL
            for (int u = 0; u < 32; ++u, sa += 16) {
                const float4 S = lds128(sa);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    float key;
                    if (FORM == 0) {
                        float fx = ax[r] + S.x, fy = ay[r] + S.y, fz = az[r] + S.z;
                        float b = fmaf(fz, hz[r], fmaf(fy, hy[r], fx * hx[r]));
                        float nc = fmaf(-fz, fz, fmaf(-fy, fy, fmaf(-fx, fx, S.w)));
                        key = fmaf(b, fminf(b, 0.f), nc);
                    } else if (FORM == 1) {
                        float b = fmaf(S.x, hx[r], fmaf(S.y, hy[r], fmaf(S.z, hz[r], P[r])));
                        float t = fmaf(S.x, ax[r], fmaf(S.y, ay[r], fmaf(S.z, az[r], S.w)));
                        key = fmaf(b, fminf(b, 0.f), t - Q[r]);
                    } else if (FORM == 3) {   // common-origin form: (W + 2 o.c) - |o|^2 precomputed per sphere in S.w
                        float b = fmaf(S.x, hx[r], fmaf(S.y, hy[r], fmaf(S.z, hz[r], P[r])));
                        key = fmaf(-b, fabsf(b), fmaf(b, b, S.w));
                    } else {   // FORM 2: b min(b,0) = c c - c |c| with c = b / sqrt 2 (h, P prescaled): FMNMX -> FFMA with |.| modifier
                        float b = fmaf(S.x, hx[r], fmaf(S.y, hy[r], fmaf(S.z, hz[r], P[r])));
                        float t = fmaf(S.x, ax[r], fmaf(S.y, ay[r], fmaf(S.z, az[r], S.w)));
                        key = fmaf(-b, fabsf(b), fmaf(b, b, t - Q[r]));
                    }
                    if (MODE == 0) acc_f += key;
                    else acc[r] = __funnelshift_l(__float_as_uint(key), acc[r], 1);
                }
            }
            if (MODE == 1) {
                unsigned all = 0xffffffffu;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(la + (unsigned)(r * BLOCK * 4)), "r"(acc[r]) : "memory");
                    all &= acc[r];
                }
                total |= (all != 0xffffffffu ? 1u : 0u) << g;
            }
        }
        ax[0] += 1e-3f;
    }
    out[blockIdx.x * BLOCK + threadIdx.x] = acc_f + (float)total;
}

template <int FORM, int MODE, int BLOCK, int UNROLL>
void run(const char* name, const float4* sph, float* out, int ctas_per_sm) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int passes = 64, blocks = 148 * ctas_per_sm, ns = NS / 32 * 32;
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<FORM, MODE, BLOCK, UNROLL><<<blocks, BLOCK>>>(sph, out, passes);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    double warps_per_smsp = (double)ctas_per_sm * (BLOCK / 32) / 4.0;
    double cyc_per_test = best * 1e-3 * 1.965e9 / ((double)passes * ns) / warps_per_smsp / R;   // SMSP issue cycles per warp-test
    double tests = (double)blocks * BLOCK * R * passes * ns;
    printf("%-44s block %3d x%d/SM unroll %2d: %7.3f ms  %5.2f cycles per warp-test  %5.1f TFLOP/s (17 flop/test)\n", name, BLOCK,
           ctas_per_sm, UNROLL, best, cyc_per_test, 17.0 * tests / (best * 1e-3) / 1e12);
}

int main() {
    float4 h[NS];
    for (int i = 0; i < NS; ++i) h[i] = make_float4(-(float)(i % 22) + 11.f, -0.2f, -(float)(i / 22) + 11.f, 0.04f);
    float4* sph; cudaMalloc(&sph, sizeof(h)); cudaMemcpy(sph, h, sizeof(h), cudaMemcpyHostToDevice);
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    run<0, 0, 128, 8>("translated form, math only (11+1 instr)", sph, out, 5);
    run<1, 0, 128, 8>("expanded form, math only (9+1 instr)", sph, out, 5);
    run<0, 1, 128, 8>("translated form + funnel + mask words", sph, out, 5);
    run<1, 1, 128, 8>("expanded form + funnel + mask words (shipped)", sph, out, 5);
    run<2, 1, 128, 8>("expanded, all-FMA key c c - c |c| (shipped)", sph, out, 5);
    run<2, 1, 128, 8>("expanded, FMNMX replaced by FFMA |.| (all-FMA key)", sph, out, 4);
    run<3, 1, 128, 8>("common-origin form (5 FFMA + funnel), shipped for camera rays", sph, out, 4);
    run<3, 1, 128, 8>("common-origin form (5 FFMA + funnel)", sph, out, 5);
    run<2, 1, 128, 8>("expanded, all-FMA key", sph, out, 1);
    run<2, 1, 128, 8>("expanded, all-FMA key", sph, out, 2);
    run<2, 1, 128, 8>("expanded, all-FMA key", sph, out, 3);
    run<3, 1, 128, 8>("common-origin form", sph, out, 2);
    run<1, 1, 128, 4>("expanded form + funnel + mask words", sph, out, 5);
    run<1, 1, 128, 16>("expanded form + funnel + mask words", sph, out, 5);
    run<1, 1, 128, 8>("expanded form + funnel + mask words", sph, out, 3);
    run<1, 1, 128, 8>("expanded form + funnel + mask words", sph, out, 4);
    run<1, 1, 128, 8>("expanded form + funnel + mask words", sph, out, 6);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
