// loopbench.cu — the cull hot loop in isolation (DESIGN.md "Roofline"): which ingredient costs what.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o loopbench loopbench.cu
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
// MODE 0: math only (keys summed)      1: + funnel entries, unconditional STS, predicated pointer bump
//      2: + list-full vote every 4     3: mode 1 but any-test per sphere pair (as shipped)
template <int MODE, int BLOCK, int UNROLL>
__global__ void __launch_bounds__(BLOCK) k(const float4* __restrict__ spheres, float* out, int passes) {
    __shared__ float4 s[NS];
    __shared__ uint32_t list[32 * BLOCK];
    for (int i = threadIdx.x; i < NS; i += BLOCK) s[i] = spheres[i];
    __syncthreads();
    float ox[R], oy[R], oz[R], hx[R], hy[R], hz[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        ox[r] = 0.1f * threadIdx.x + r; oy[r] = 1.0f + r; oz[r] = -2.0f * r;
        hx[r] = 0.3f + 0.01f * r; hy[r] = -0.5f; hz[r] = 0.8f;
        asm volatile("" : "+f"(ox[r]), "+f"(oy[r]), "+f"(oz[r]), "+f"(hx[r]), "+f"(hy[r]), "+f"(hz[r]));
    }
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(s);
    const unsigned lbase = (unsigned)__cvta_generic_to_shared(list + threadIdx.x);
    float acc_f = 0.f;
    unsigned total = 0;
    for (int p = 0; p < passes; ++p) {
        unsigned ptr = lbase;
        const unsigned limit = lbase + 28 * BLOCK * 4;
        unsigned sa = sbase;
        for (int k0 = 0; k0 < NS; k0 += UNROLL, sa += 16 * UNROLL) {
#pragma unroll
            for (int u = 0; u < UNROLL; u += 2) {
                unsigned a01[2];
#pragma unroll
                for (int v = 0; v < 2; ++v) {
                    const float4 S = lds128(sa + 16 * (u + v));
                    unsigned acc = (unsigned)(k0 + u + v);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        float fx = ox[r] + S.x, fy = oy[r] + S.y, fz = oz[r] + S.z;
                        float b = fmaf(fz, hz[r], fmaf(fy, hy[r], fx * hx[r]));
                        float nc = fmaf(-fz, fz, fmaf(-fy, fy, fmaf(-fx, fx, S.w)));
                        float key = fmaf(b, fminf(b, 0.f), nc);
                        if (MODE == 0) acc_f += key;
                        else acc = __funnelshift_l(__float_as_uint(key), acc, 1);
                    }
                    a01[v] = acc;
                    if (MODE == 1 || MODE == 2) {
                        asm volatile("st.shared.u32 [%0], %1;" ::"r"(ptr), "r"(acc) : "memory");
                        if ((~acc) & 15u) ptr += BLOCK * 4;
                    }
                }
                if (MODE == 3) {
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(ptr), "r"(a01[0]) : "memory");
                    asm volatile("st.shared.u32 [%0+%2], %1;" ::"r"(ptr), "r"(a01[1]), "n"(BLOCK * 4) : "memory");
                    if (~(a01[0] & a01[1]) & 15u) ptr += BLOCK * 8;
                }
            }
            if (MODE >= 2) {
                if (__any_sync(0xffffffffu, ptr > limit)) { total += (ptr - lbase); ptr = lbase; }
            }
        }
        total += ptr - lbase;
        ox[0] += 1e-3f;
    }
    out[blockIdx.x * BLOCK + threadIdx.x] = acc_f + (float)total;
}

template <int MODE, int BLOCK, int UNROLL>
void run(const char* name, const float4* sph, float* out, int ctas_per_sm) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int passes = 64, blocks = 148 * ctas_per_sm;
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<MODE, BLOCK, UNROLL><<<blocks, BLOCK>>>(sph, out, passes);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    double warps_per_smsp = (double)ctas_per_sm * (BLOCK / 32) / 4.0;
    double cyc_per_sphere = best * 1e-3 * 1.965e9 / ((double)passes * NS) / warps_per_smsp;   // SMSP cycles per warp-sphere (4 rays)
    double tests = (double)blocks * BLOCK * R * passes * NS;
    printf("%-38s block %3d x%d/SM unroll %d: %7.3f ms  %6.1f cycles per warp-sphere  %5.1f TFLOP/s (17 flop/test)\n", name, BLOCK,
           ctas_per_sm, UNROLL, best, cyc_per_sphere, 17.0 * tests / (best * 1e-3) / 1e12);
}

int main() {
    float4 h[NS];
    for (int i = 0; i < NS; ++i) h[i] = make_float4(-(float)(i % 22) + 11.f, -0.2f, -(float)(i / 22) + 11.f, 0.04f);
    float4* sph; cudaMalloc(&sph, sizeof(h)); cudaMemcpy(sph, h, sizeof(h), cudaMemcpyHostToDevice);
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    run<0, 128, 2>("math only", sph, out, 5);
    run<0, 128, 4>("math only", sph, out, 5);
    run<1, 128, 2>("+ entries, STS, ptr bump", sph, out, 5);
    run<1, 128, 4>("+ entries, STS, ptr bump", sph, out, 5);
    run<2, 128, 2>("+ list-full vote", sph, out, 5);
    run<2, 128, 4>("+ list-full vote", sph, out, 5);
    run<3, 128, 4>("paired any-test (no vote)", sph, out, 5);
    run<2, 128, 4>("+ list-full vote", sph, out, 4);
    run<2, 128, 4>("+ list-full vote", sph, out, 8);
    run<2, 256, 4>("+ list-full vote", sph, out, 2);
    run<2, 128, 8>("+ list-full vote", sph, out, 5);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
