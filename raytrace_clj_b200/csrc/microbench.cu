// microbench.cu — issue-slot / pipe measurements that size the intersect loop (DESIGN.md "Roofline").
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu ; run on a B200.
#include <cuda_runtime.h>
#include <cstdio>

#define ITERS 8192

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b, int iters, const float4* __restrict__ gsrc) {
    __shared__ float sm[256 * 8];
    __shared__ float4 sph[64];
    for (int i = threadIdx.x; i < 256 * 8; i += 256) sm[i] = 0.f;
    if (threadIdx.x < 64) sph[threadIdx.x] = gsrc[threadIdx.x];
    __syncthreads();
    float x[16];
    float2 y[8];
    float m[8];
    int ptr = threadIdx.x;
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3f + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) { y[i] = make_float2(x[2 * i], x[2 * i + 1]); m[i] = x[i] * 0.5f; }
    float2 aa = make_float2(a, a), bb = make_float2(b, b);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {            // 16 FFMA
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
        } else if (MODE == 1) {     // 8 FFMA2 (= 16 FMA)
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = __ffma2_rn(y[i], aa, bb);
        } else if (MODE == 2) {     // 8 FFMA2 + 4 FMNMX
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = __ffma2_rn(y[i], aa, bb);
#pragma unroll
            for (int i = 0; i < 4; ++i) m[i] = fminf(m[i], y[(i + 3) & 7].x);
        } else if (MODE == 3) {     // 8 FFMA2 + 8 FMNMX
#pragma unroll
            for (int i = 0; i < 8; ++i) { y[i] = __ffma2_rn(y[i], aa, bb); m[i] = fminf(m[i], y[(i + 3) & 7].x); }
        } else if (MODE == 4) {     // 16 FFMA + 8 FMNMX
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
#pragma unroll
            for (int i = 0; i < 8; ++i) m[i] = fminf(m[i], x[(i + 3) & 15]);
        } else if (MODE == 5) {     // 8 FFMA2 + 16 FMNMX
#pragma unroll
            for (int i = 0; i < 8; ++i) { y[i] = __ffma2_rn(y[i], aa, bb); m[i] = fminf(m[i], y[(i + 3) & 7].x); m[i] = fmaxf(m[i], y[(i + 5) & 7].y); }
        } else if (MODE == 6) {     // 8 FFMA2 + 4 x (FSETP + predicated IADD) + 1 STS
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = __ffma2_rn(y[i], aa, bb);
#pragma unroll
            for (int i = 0; i < 4; ++i) if (y[i].y >= 0.f) ptr += 256;
            sm[ptr & 2047] = y[0].x;
        } else if (MODE == 7) {     // 8 FFMA2 + 2 LDS.128 (broadcast) feeding them
            float4 s0 = sph[it & 63], s1 = sph[(it + 7) & 63];
            float2 c0 = make_float2(s0.x, s0.y), c1 = make_float2(s0.z, s0.w), c2 = make_float2(s1.x, s1.y), c3 = make_float2(s1.z, s1.w);
            y[0] = __ffma2_rn(y[0], aa, c0); y[1] = __ffma2_rn(y[1], aa, c1); y[2] = __ffma2_rn(y[2], aa, c2); y[3] = __ffma2_rn(y[3], aa, c3);
#pragma unroll
            for (int i = 4; i < 8; ++i) y[i] = __ffma2_rn(y[i], aa, bb);
        } else if (MODE == 8) {     // planned mix per 2 ray pairs x 1 static sphere: 22 FFMA2, 8 FMNMX, 4 FSETP, 2 LDS.128, STS, IADD
            float4 s0 = sph[it & 63], s1 = sph[(it + 7) & 63];
            float2 c0 = make_float2(s0.x, s0.y), c1 = make_float2(s0.z, s0.w), c2 = make_float2(s1.x, s1.y), c3 = make_float2(s1.z, s1.w);
            float2 k0, k1;
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                float2 fx = __fadd2_rn(y[4 * p + 0], c0), fy = __fadd2_rn(y[4 * p + 1], c1), fz = __fadd2_rn(y[4 * p + 2], c2);
                float2 bq = __fmul2_rn(fx, aa); bq = __ffma2_rn(fy, bb, bq); bq = __ffma2_rn(fz, aa, bq);
                float2 cq = __ffma2_rn(fx, fx, c3); cq = __ffma2_rn(fy, fy, cq); cq = __ffma2_rn(fz, fz, cq);
                float2 dq = __fmul2_rn(bq, bq); dq = __ffma2_rn(y[4 * p + 3], cq, dq);
                float2 kk = make_float2(fminf(dq.x, fmaxf(-bq.x, -cq.x)), fminf(dq.y, fmaxf(-bq.y, -cq.y)));
                if (p == 0) k0 = kk; else k1 = kk;
            }
            bool any = (k0.x >= 0.f) | (k0.y >= 0.f) | (k1.x >= 0.f) | (k1.y >= 0.f);
            sm[ptr & 2047] = (float)it;
            if (any) ptr += 256;
            m[0] += k0.x + k1.y;
        }
    }
    float acc = (float)ptr;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += x[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += y[i].x + y[i].y + m[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + sm[threadIdx.x];
}

template <int MODE>
void run(const char* name, float* out, int blocks, const float4* gsrc, double fma_per_iter) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, 256>>>(out, 1.0000001f, 1e-7f, ITERS, gsrc);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    double warps_per_smsp = (double)blocks * 8 / 148 / 4;
    double cyc = best * 1e-3 * 1.965e9 / ITERS / warps_per_smsp;
    printf("%-52s %8.3f ms  %6.2f cycles/iter/warp  %5.1f TFLOP/s (FMA-pipe lane-ops x2)\n", name, best, cyc,
           2.0 * fma_per_iter * ITERS * (double)blocks * 256 / (best * 1e-3) / 1e12);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int blocks = p.multiProcessorCount * 8;
    float* out; cudaMalloc(&out, (size_t)blocks * 256 * 4);
    float4* gsrc; cudaMalloc(&gsrc, 64 * 16); cudaMemset(gsrc, 0, 64 * 16);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    run<0>("16 FFMA", out, blocks, gsrc, 16);
    run<1>("8 FFMA2", out, blocks, gsrc, 16);
    run<2>("8 FFMA2 + 4 FMNMX", out, blocks, gsrc, 16);
    run<3>("8 FFMA2 + 8 FMNMX", out, blocks, gsrc, 16);
    run<4>("16 FFMA + 8 FMNMX", out, blocks, gsrc, 16);
    run<5>("8 FFMA2 + 16 FMNMX", out, blocks, gsrc, 16);
    run<6>("8 FFMA2 + 4x(FSETP,@IADD) + STS", out, blocks, gsrc, 16);
    run<7>("8 FFMA2 + 2 LDS.128", out, blocks, gsrc, 16);
    run<8>("planned loop: 22 FFMA2-class + 8 FMNMX + 4 FSETP + ...", out, blocks, gsrc, 44);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
