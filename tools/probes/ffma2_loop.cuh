// Engine-shaped inner product (one warp = one 10-row x 256-column tile, K = 256 per pass, weights + activations in
// shared memory) in several FFMA2 issue orders.  Compiled twice (ptxas -O1 / -O3) by ffma2_reuse_probe.sh.
#pragma once
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 p_ffma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 p_pack2(float x, float y) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r; }
__device__ __forceinline__ f32x2 p_abs2(f32x2 v) { return v & 0x7fffffff7fffffffull; }
constexpr int P_ROWS = 10, P_S = 260, P_RT = 5, P_W = 256, P_KC = 64;
__device__ __forceinline__ bool p_is_err(int r) { return r == 2; }
__device__ __forceinline__ void p_load_act(float4 (&a)[P_ROWS], const float* p) {
#pragma unroll
    for (int r = 0; r < P_ROWS; ++r) a[r] = *reinterpret_cast<const float4*>(p + r * P_S);
}
template <int ORDER>
__device__ __forceinline__ void p_fma_group(f32x2 (&acc)[P_ROWS][4], const float4 (&a)[P_ROWS], ulonglong2& w0, ulonglong2& w1,
                                            const float* wnext, int wstride, int dcol) {
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
        const ulonglong2 n0 = *reinterpret_cast<const ulonglong2*>(wnext + jj * wstride);
        const ulonglong2 n1 = *reinterpret_cast<const ulonglong2*>(wnext + jj * wstride + dcol);
        const f32x2 wv[4] = {w0.x, w0.y, w1.x, w1.y};
        f32x2 wa[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) wa[c] = p_abs2(wv[c]);
        if (ORDER == 0) {
#pragma unroll
            for (int r = 0; r < P_ROWS; ++r) {
                const float av = jj == 0 ? a[r].x : jj == 1 ? a[r].y : jj == 2 ? a[r].z : a[r].w;
                const f32x2 ap = p_pack2(av, av);
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = p_ffma2(ap, p_is_err(r % P_RT) ? wa[c] : wv[c], acc[r][c]);
            }
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
#pragma unroll
                for (int rr = 0; rr < P_ROWS; ++rr) {
                    const int r = (ORDER == 2 && (c & 1)) ? P_ROWS - 1 - rr : rr;
                    const float av = jj == 0 ? a[r].x : jj == 1 ? a[r].y : jj == 2 ? a[r].z : a[r].w;
                    const f32x2 ap = p_pack2(av, av);
                    acc[r][c] = p_ffma2(ap, p_is_err(r % P_RT) ? wa[c] : wv[c], acc[r][c]);
                }
            }
        }
        w0 = n0; w1 = n1;
    }
}
// smem: [8 warps][10 rows][260] activations, then a [64][256] weight chunk (re-read 4x per pass: K = 256)
template <int ORDER>
__device__ __forceinline__ void p_body(float* out, int iters) {
    extern __shared__ __align__(16) float sm[];
    float* act = sm + (threadIdx.x >> 5) * P_ROWS * P_S;
    float* wch = sm + 8 * P_ROWS * P_S;
    for (int i = threadIdx.x; i < (P_KC + 8) * P_W; i += blockDim.x) wch[i] = 1e-3f * (float)((i * 7) % 13 - 6);
    for (int i = threadIdx.x & 31; i < P_ROWS * P_S; i += 32) act[i] = 1e-2f * (float)((i * 5) % 11 - 5);
    __syncthreads();
    const int cg = threadIdx.x & 31;
    const float* wrow = wch + 4 * cg;
    const int wstride = P_W, dcol = 128;
    f32x2 acc2[P_ROWS][4];
#pragma unroll
    for (int r = 0; r < P_ROWS; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc2[r][c] = 0ull;
    for (int it = 0; it < iters; ++it) {
        for (int k0 = 0; k0 < P_W; k0 += P_KC) {
            const float* arow = act + k0;
            float4 aA[P_ROWS], aB[P_ROWS];
            ulonglong2 w0, w1;
            p_load_act(aA, arow);
            w0 = *reinterpret_cast<const ulonglong2*>(wrow);
            w1 = *reinterpret_cast<const ulonglong2*>(wrow + dcol);
#pragma unroll 2
            for (int j = 0; j < P_KC; j += 8) {
                p_load_act(aB, arow + j + 4);
                p_fma_group<ORDER>(acc2, aA, w0, w1, wrow + (j + 1) * wstride, wstride, dcol);
                p_load_act(aA, arow + j + 8);
                p_fma_group<ORDER>(acc2, aB, w0, w1, wrow + (j + 5) * wstride, wstride, dcol);
            }
        }
    }
    f32x2 s = 0;
#pragma unroll
    for (int r = 0; r < P_ROWS; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) s ^= acc2[r][c];
    if (s == 0x1234567812345678ull) out[threadIdx.x] = 1.f;
}
// register-only FFMA2 streams: REUSE = 1: 16 accumulators share one weight pair per step (slot-B reuse possible);
// REUSE = 0: every FFMA2 has its own weight pair and scalar (5 fresh registers per instruction)
template <int REUSE>
__device__ __forceinline__ void p_regonly(float* out, int iters, float seed) {
    f32x2 acc[16], w[16];
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { acc[i] = p_pack2(seed + i, seed - i); w[i] = p_pack2(0.999f + 1e-6f * (i + threadIdx.x), 0.998f); a[i] = seed * (1.0f + 1e-7f * i); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = p_ffma2(p_pack2(a[i], a[i]), REUSE ? w[rep] : w[(i + rep) & 15], acc[i]);
        }
    }
    f32x2 s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s ^= acc[i];
    if (s == 0x1234567812345678ull) out[threadIdx.x] = 1.f;
}
#define P_CAT_(a, b) a##b
#define P_CAT(a, b) P_CAT_(a, b)
#define P_DEFINE_KERNELS(SFX)                                                                                              \
    extern "C" __global__ void __launch_bounds__(256, 1) P_CAT(k_loop0_, SFX)(float* out, int iters) { p_body<0>(out, iters); }   \
    extern "C" __global__ void __launch_bounds__(256, 1) P_CAT(k_loop1_, SFX)(float* out, int iters) { p_body<1>(out, iters); }   \
    extern "C" __global__ void __launch_bounds__(256, 1) P_CAT(k_loop2_, SFX)(float* out, int iters) { p_body<2>(out, iters); }   \
    extern "C" __global__ void __launch_bounds__(256, 1) P_CAT(k_reg1_, SFX)(float* out, int iters, float s) { p_regonly<1>(out, iters, s); } \
    extern "C" __global__ void __launch_bounds__(256, 1) P_CAT(k_reg0_, SFX)(float* out, int iters, float s) { p_regonly<0>(out, iters, s); }
