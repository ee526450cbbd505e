// tcprobe — standalone probe of the tcgen05 (tf32) path on sm_100a, written before the tensor-core cull:
//   1. layout / descriptor check: D[128 x N] = A[128 x 32] . B[N x 32]^T, K-major, no swizzle, 4 chained K=8 MMAs
//   2. numerics: how fp32 operands are narrowed to tf32 (truncate / round), accumulation error of the chained MMAs
//   3. rates per SM: MMA issue, TMEM loads (tcgen05.ld 32x32b.x32) with 4 / 8 / 16 warps, with the two candidate epilogues
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tcprobe tcprobe.cu ; run under `timeout`.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {   // false = gave up (never hang the box)
    for (long long spin = 0; spin < 400000000ll; ++spin)
        if (mbar_try_wait(bar, parity)) return true;
    return false;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= 1ull << 46;                   // descriptor version (sm_100)
    return d;                          // layout type 0 = no swizzle, base offset 0
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);   // f32 acc, tf32 x tf32, K-major
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
#define TMEM_LD32(taddr, v)                                                                                                   \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                   \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                   \
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                   \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),           \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),     \
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),   \
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])    \
                 : "r"(taddr))
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int KTOT = 32;          // tf32 elements per operand row (4 MMAs of K = 8)
// canonical K-major no-swizzle layout of an operand tile [rows x KTOT]: 16-byte chunk kc of row r at
//   kc * (rows * 16) + (r / 8) * 128 + (r % 8) * 16        (core matrix = 8 rows x 16 bytes, contiguous)
// => K-direction stride between core matrices = rows * 16 bytes, row-group stride = 128 bytes
__host__ __device__ inline size_t canon_off(int rows, int r, int k) {
    return (size_t)(k / 4) * ((size_t)rows * 16) + (size_t)(r / 8) * 128 + (size_t)(r % 8) * 16 + (size_t)(k % 4) * 4;
}

struct ProbeOut {
    unsigned long long cyc[16];
    unsigned flags;
};

// ---- 1/2: one CTA, D = A B^T into TMEM, read back -------------------------------------------------
template <int N>
__global__ void __launch_bounds__(128) gemm_probe(const float* A_canon, const float* B_canon, float* D, int swap_lbo_sbo, ProbeOut* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* sA = reinterpret_cast<float*>(smem);
    float* sB = reinterpret_cast<float*>(smem + 128 * KTOT * 4);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < 128 * KTOT; i += 128) sA[i] = A_canon[i];
    for (int i = tid; i < N * KTOT; i += 128) sB[i] = B_canon[i];
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tb = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = make_idesc(128, N);
        uint32_t a_lbo = 128 * 16, a_sbo = 128, b_lbo = N * 16, b_sbo = 128;
        if (swap_lbo_sbo) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
        for (int j = 0; j < KTOT / 8; ++j) {
            const uint64_t ad = make_desc(smem_u32(sA) + j * 2 * 128 * 16, a_lbo, a_sbo);
            const uint64_t bd = make_desc(smem_u32(sB) + j * 2 * N * 16, b_lbo, b_sbo);
            mma_tf32(tb, ad, bd, idesc, j > 0);
        }
        mma_commit(&bar);
    }
    const bool ok = mbar_wait(&bar, 0);
    fence_after();
    if (!ok && tid == 0) out->flags |= 1u;
    if (ok) {
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t v[32];
            TMEM_LD32(tb + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory");
}

// ---- 3: rates ------------------------------------------------------------------------------------
// mode 0: MMA issue rate (thread 0 issues `iters` x 4 MMAs of 128 x 256 x 8, one commit at the end)
// mode 1: TMEM load rate: every warp loops over its 32 lanes x its column range with tcgen05.ld.x32, XOR-consumed
// mode 2: loads + one funnel shift per value (sign collection: the "one value per pair" epilogue)
// mode 3: loads of two values per pair + 2 FFMA + funnel shift (the "two values per pair" epilogue)
// mode 4: mode 2 while thread 0 of an extra warp keeps issuing MMAs into the other half of TMEM
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32 + 32) rate_probe(int mode, int iters, ProbeOut* out, unsigned* sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* sA = reinterpret_cast<float*>(smem);
    float* sB = reinterpret_cast<float*>(smem + 128 * KTOT * 4);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, nthr = WARPS * 32 + 32;
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < 128 * KTOT; i += nthr) sA[i] = 1.0f + (float)(i % 7);
    for (int i = tid; i < 256 * KTOT; i += nthr) sB[i] = 0.5f - (float)(i % 5);
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tb = tmem_base;
    const uint32_t idesc = make_idesc(128, 256);
    // fill all 512 columns with something finite first
    if (tid == 0) {
        for (int h = 0; h < 2; ++h)
            for (int j = 0; j < KTOT / 8; ++j)
                mma_tf32(tb + h * 256, make_desc(smem_u32(sA) + j * 2 * 128 * 16, 128 * 16, 128), make_desc(smem_u32(sB) + j * 2 * 256 * 16, 256 * 16, 128), idesc, j > 0);
        mma_commit(&bar);
    }
    bool ok = mbar_wait(&bar, 0);
    fence_after();
    __syncthreads();
    unsigned acc = 0;
    const long long t0 = clock64();
    if (warp == WARPS) {                                   // the MMA warp
        if ((mode == 0 || mode == 4) && tid == WARPS * 32) {
            for (int it = 0; it < iters; ++it)
                for (int j = 0; j < KTOT / 8; ++j)
                    mma_tf32(tb + (mode == 4 ? 256 : 0), make_desc(smem_u32(sA) + j * 2 * 128 * 16, 128 * 16, 128),
                             make_desc(smem_u32(sB) + j * 2 * 256 * 16, 256 * 16, 128), idesc, 1);
            mma_commit(&bar);
            ok = ok && mbar_wait(&bar, 1);
            fence_after();
        }
    } else if (mode >= 1) {
        // warp w: lanes 32 (w % 4) .., columns [(w / 4) * span, + span) of the first 256
        const int span = 256 / (WARPS / 4);
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const int cbase = (warp >> 2) * span;
        float fa = 0.f;
        for (int it = 0; it < iters; ++it) {
            if (mode == 3) {
                for (int c0 = 0; c0 < span / 2; c0 += 32) {          // "b" values in the first half of the span, "nc" in the second
                    uint32_t vb[32], vn[32];
                    TMEM_LD32(tb + lane_base + (uint32_t)(cbase + c0), vb);
                    TMEM_LD32(tb + lane_base + (uint32_t)(cbase + span / 2 + c0), vn);
                    tmem_wait_ld();
                    unsigned w = 0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float c = __uint_as_float(vb[j]);
                        const float key = fmaf(-c, fabsf(c), fmaf(c, c, __uint_as_float(vn[j])));
                        w = __funnelshift_l(__float_as_uint(key), w, 1);
                    }
                    acc ^= w;
                }
            } else {
                for (int c0 = 0; c0 < span; c0 += 32) {
                    uint32_t v[32];
                    TMEM_LD32(tb + lane_base + (uint32_t)(cbase + c0), v);
                    tmem_wait_ld();
                    if (mode == 5) {                                  // half of the values by funnel shift (ALU pipe), half by FSET + FFMA (FMA pipe?)
                        unsigned w = 0;
                        float af = 0.f;
#pragma unroll
                        for (int j = 0; j < 16; ++j) w = __funnelshift_l(v[j], w, 1);
#pragma unroll
                        for (int j = 16; j < 32; ++j) af = fmaf(af, 2.0f, !(__uint_as_float(v[j]) < 0.0f) ? 1.0f : 0.0f);
                        acc ^= (w << 16) | (unsigned)__float2uint_rz(af);
                    } else if (mode == 6) {                           // all values by FSET + FFMA
                        float a0 = 0.f, a1 = 0.f;
#pragma unroll
                        for (int j = 0; j < 16; ++j) a0 = fmaf(a0, 2.0f, !(__uint_as_float(v[j]) < 0.0f) ? 1.0f : 0.0f);
#pragma unroll
                        for (int j = 16; j < 32; ++j) a1 = fmaf(a1, 2.0f, !(__uint_as_float(v[j]) < 0.0f) ? 1.0f : 0.0f);
                        acc ^= ((unsigned)__float2uint_rz(a0) << 16) | (unsigned)__float2uint_rz(a1);
                    } else if (mode == 1) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) acc ^= v[j] ^ v[j + 1] ^ v[j + 2] ^ v[j + 3];   // LOP3s: ~0.4 per value
                    } else {
                        unsigned w = 0;
#pragma unroll
                        for (int j = 0; j < 32; ++j) w = __funnelshift_l(v[j], w, 1);
                        acc ^= w;
                    }
                }
            }
        }
        acc ^= __float_as_uint(fa);
    }
    const long long t1 = clock64();
    __syncthreads();
    if (tid == 0) { out->cyc[0] = (unsigned long long)(t1 - t0); if (!ok) out->flags |= 2u; }
    if (tid == WARPS * 32) out->cyc[1] = (unsigned long long)(t1 - t0);
    if (acc == 0x12345678u) sink[0] = acc;
    fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory");
}

static float trunc_tf32(float x) { uint32_t b; memcpy(&b, &x, 4); b &= 0xffffe000u; memcpy(&x, &b, 4); return x; }
static float round_tf32(float x) { uint32_t b; memcpy(&b, &x, 4); b = (b + 0x1000u) & 0xffffe000u; memcpy(&x, &b, 4); return x; }

template <int N>
static void run_gemm(const char* what, int mode, int swap) {
    // mode 0: operands already tf32 (products exact) -> layout check + accumulation error
    // mode 1: full fp32 operands -> truncate or round?
    std::mt19937 rng(1234 + mode);
    std::uniform_real_distribution<float> U(-1.f, 1.f);
    std::uniform_int_distribution<int> E(-6, 6);
    std::vector<float> A(128 * KTOT), B((size_t)N * KTOT), Ac(128 * KTOT), Bc((size_t)N * KTOT);
    for (auto& x : A) { x = std::ldexp(U(rng), E(rng)); if (mode == 0) x = trunc_tf32(x); }
    for (auto& x : B) { x = std::ldexp(U(rng), E(rng)); if (mode == 0) x = trunc_tf32(x); }
    for (int r = 0; r < 128; ++r) for (int k = 0; k < KTOT; ++k) Ac[canon_off(128, r, k) / 4] = A[r * KTOT + k];
    for (int r = 0; r < N; ++r) for (int k = 0; k < KTOT; ++k) Bc[canon_off(N, r, k) / 4] = B[r * KTOT + k];
    float *dA, *dB, *dD; ProbeOut* dO;
    CK(cudaMalloc(&dA, Ac.size() * 4)); CK(cudaMalloc(&dB, Bc.size() * 4)); CK(cudaMalloc(&dD, (size_t)128 * N * 4)); CK(cudaMalloc(&dO, sizeof(ProbeOut)));
    CK(cudaMemcpy(dA, Ac.data(), Ac.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, Bc.data(), Bc.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, (size_t)128 * N * 4)); CK(cudaMemset(dO, 0, sizeof(ProbeOut)));
    const size_t smem = (size_t)(128 + N) * KTOT * 4 + 1024;
    CK(cudaFuncSetAttribute(gemm_probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gemm_probe<N><<<1, 128, smem>>>(dA, dB, dD, swap, dO);
    CK(cudaDeviceSynchronize());
    std::vector<float> D((size_t)128 * N); ProbeOut o;
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&o, dO, sizeof o, cudaMemcpyDeviceToHost));
    double worst_t = 0, worst_r = 0, worst_exact = 0; int bad = 0;
    for (int r = 0; r < 128; ++r)
        for (int c = 0; c < N; ++c) {
            double st = 0, sr = 0, se = 0, mag = 0;
            for (int k = 0; k < KTOT; ++k) {
                const float a = A[r * KTOT + k], b = B[(size_t)c * KTOT + k];
                st += (double)trunc_tf32(a) * trunc_tf32(b);
                sr += (double)round_tf32(a) * round_tf32(b);
                se += (double)a * b;
                mag += std::fabs((double)a * b);
            }
            const double d = D[(size_t)r * N + c];
            worst_t = std::max(worst_t, std::fabs(d - st) / mag);
            worst_r = std::max(worst_r, std::fabs(d - sr) / mag);
            worst_exact = std::max(worst_exact, std::fabs(d - se) / mag);
            if (std::fabs(d - st) > 1e-3 * mag) ++bad;
        }
    printf("%-44s N=%3d swap=%d flags=%u  mismatches(>1e-3)=%6d  max|D-ref|/sum|terms| in units of 2^-24: trunc-model %.3g  round-model %.3g  exact-fp32-operands %.3g\n",
           what, N, swap, o.flags, bad, worst_t * 16777216.0, worst_r * 16777216.0, worst_exact * 16777216.0);
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dO);
}

template <int WARPS>
static void run_rate(const char* what, int mode, int iters) {
    ProbeOut* dO; unsigned* dS;
    CK(cudaMalloc(&dO, sizeof(ProbeOut))); CK(cudaMalloc(&dS, 64)); CK(cudaMemset(dO, 0, sizeof(ProbeOut)));
    const size_t smem = (size_t)(128 + 256) * KTOT * 4 + 1024;
    CK(cudaFuncSetAttribute(rate_probe<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rate_probe<WARPS><<<1, WARPS * 32 + 32, smem>>>(mode, iters, dO, dS);
    CK(cudaDeviceSynchronize());
    ProbeOut o; CK(cudaMemcpy(&o, dO, sizeof o, cudaMemcpyDeviceToHost));
    const double cyc = (double)o.cyc[0], cyc_mma = (double)o.cyc[1];
    if (mode == 0) {
        const double macs = (double)iters * 4 * 128.0 * 256 * 8;
        printf("%-60s warps=%2d flags=%u  %.0f cycles  %.1f tf32 MAC/clk/SM  (128x256x8 MMA every %.1f clk)\n", what, WARPS, o.flags, cyc_mma, macs / cyc_mma, cyc_mma / (iters * 4.0));
    } else {
        const double values = (double)iters * 128.0 * 256;           // every (lane, column) of the first 256 columns once per iteration
        const double pairs = mode == 3 ? values / 2 : values;
        printf("%-60s warps=%2d flags=%u  %.0f cycles  %.1f B/clk/SM TMEM read  %.2f pairs/clk/SM", what, WARPS, o.flags, cyc, values * 4 / cyc, pairs / cyc);
        if (mode == 4) printf("  (MMA warp: %.0f cycles, %.1f MAC/clk)", cyc_mma, (double)iters * 4 * 128.0 * 256 * 8 / cyc_mma);
        printf("\n");
    }
    cudaFree(dO); cudaFree(dS);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("%s, %d SMs, cc %d.%d\n", p.name, p.multiProcessorCount, p.major, p.minor);
    run_gemm<128>("tf32 operands (layout + accumulation)", 0, 0);
    run_gemm<256>("tf32 operands (layout + accumulation)", 0, 0);
    run_gemm<128>("fp32 operands (narrowing mode)", 1, 0);
    run_gemm<256>("fp32 operands (narrowing mode)", 1, 0);
    run_rate<4>("MMA issue rate", 0, 512);
    run_rate<4>("TMEM load only (LOP3-consumed)", 1, 256);
    run_rate<8>("TMEM load only (LOP3-consumed)", 1, 256);
    run_rate<16>("TMEM load only (LOP3-consumed)", 1, 256);
    run_rate<4>("load + SHF per value", 2, 256);
    run_rate<8>("load + SHF per value", 2, 256);
    run_rate<16>("load + SHF per value", 2, 256);
    run_rate<8>("two loads + 2 FFMA + SHF per pair", 3, 256);
    run_rate<16>("two loads + 2 FFMA + SHF per pair", 3, 256);
    run_rate<16>("load + 16 SHF + 16 (FSET, FFMA) per 32 values", 5, 256);
    run_rate<16>("load + (FSET, FFMA) per value", 6, 256);
    run_rate<8>("load + SHF per value, MMAs running", 4, 256);
    run_rate<16>("load + SHF per value, MMAs running", 4, 256);
    printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
