// tools/ffma_probe.cu -- development micro-benchmark (GPU box): which formulation of the engine's inner
// product sustains the FP32 pipe on sm_100a?  All variants compute acc[rows][8 cols] += act[rows][k] * W[k][cols]
// for one warp-private activation block and a CTA-shared 32 x 256 weight chunk in shared memory, K = 256 per pass.
//   A: scalar FFMA, 10 rows x 8 cols per thread, float4 activation fragments (the engine's current loop)
//   B: packed FFMA2 (fma.rn.f32x2), column pairs, activation value duplicated into a 64-bit pair by a MOV
//   C: packed FFMA2, 10 rows x 8 cols, k-pairs: acc pair = (even-k sum, odd-k sum), weights k-interleaved
//   D: scalar FFMA, 5 rows x 8 cols (NT = 1), 16 warps
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o build/ffma_probe tools/ffma_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

constexpr int W = 256, S = W + 4, KC = 32;

__device__ __forceinline__ unsigned long long pack2(float x, float y) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& x, float& y) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

template <int ROWS>
__device__ __forceinline__ void load_act(float4 (&a)[ROWS], const float* p) {
#pragma unroll
    for (int r = 0; r < ROWS; ++r) a[r] = *reinterpret_cast<const float4*>(p + r * S);
}

// ---------------- A / D: scalar FFMA --------------------------------------------------------------
template <int ROWS, int SYNC = 0>
__global__ void __launch_bounds__(ROWS == 10 ? 256 : 512, 1) k_scalar(float* out, int iters) {
    extern __shared__ __align__(16) float sm[];
    float* wch = sm;                                   // [KC][W]
    float* act = sm + KC * W + (threadIdx.x >> 5) * ROWS * S;
    for (int i = threadIdx.x; i < KC * W; i += blockDim.x) wch[i] = 1e-3f * (float)((i * 7) % 13 - 6);
    for (int i = threadIdx.x & 31; i < ROWS * S; i += 32) act[i] = 1e-2f * (float)((i * 5) % 11 - 5);
    __syncthreads();
    const int cg = threadIdx.x & 31;
    const float* wrow = wch + 4 * cg;
    float acc[ROWS][8];
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;
    for (int it = 0; it < iters; ++it) {
        for (int k0 = 0; k0 < W; k0 += KC) {
            if (SYNC == 1) __syncthreads();                       // E: CTA barrier per chunk, like acquire_chunk()

            const float* arow = act + k0;
            float4 aA[ROWS], aB[ROWS];
            load_act<ROWS>(aA, arow);
            float4 w0 = *reinterpret_cast<const float4*>(wrow), w1 = *reinterpret_cast<const float4*>(wrow + 128);
            for (int j = 0; j < KC; j += 8) {
                load_act<ROWS>(aB, arow + j + 4);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    if (half == 1) load_act<ROWS>(aA, arow + j + 8);
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int kk = (j + half * 4 + jj + 1) % KC;
                        const float4 n0 = *reinterpret_cast<const float4*>(wrow + kk * W);
                        const float4 n1 = *reinterpret_cast<const float4*>(wrow + kk * W + 128);
                        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                        for (int r = 0; r < ROWS; ++r) {
                            const float4 a4 = half == 0 ? aA[r] : aB[r];
                            const float av = jj == 0 ? a4.x : jj == 1 ? a4.y : jj == 2 ? a4.z : a4.w;
#pragma unroll
                            for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(av, wv[c], acc[r][c]);
                        }
                        w0 = n0; w1 = n1;
                    }
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) s += acc[r][c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---------------- B: FFMA2, column pairs, duplicated activation ------------------------------------
template <int ROWS>
__global__ void __launch_bounds__(ROWS == 10 ? 256 : 512, 1) k_packed_dup(float* out, int iters) {
    extern __shared__ __align__(16) float sm[];
    float* wch = sm;
    float* act = sm + KC * W + (threadIdx.x >> 5) * ROWS * S;
    for (int i = threadIdx.x; i < KC * W; i += blockDim.x) wch[i] = 1e-3f * (float)((i * 7) % 13 - 6);
    for (int i = threadIdx.x & 31; i < ROWS * S; i += 32) act[i] = 1e-2f * (float)((i * 5) % 11 - 5);
    __syncthreads();
    const int cg = threadIdx.x & 31;
    const float* wrow = wch + 4 * cg;
    unsigned long long acc[ROWS][4];
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0ull;
    for (int it = 0; it < iters; ++it) {
        for (int k0 = 0; k0 < W; k0 += KC) {
            const float* arow = act + k0;
            float4 aA[ROWS], aB[ROWS];
            load_act<ROWS>(aA, arow);
            ulonglong2 w0 = *reinterpret_cast<const ulonglong2*>(wrow), w1 = *reinterpret_cast<const ulonglong2*>(wrow + 128);
            for (int j = 0; j < KC; j += 8) {
                load_act<ROWS>(aB, arow + j + 4);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    if (half == 1) load_act<ROWS>(aA, arow + j + 8);
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int kk = (j + half * 4 + jj + 1) % KC;
                        const ulonglong2 n0 = *reinterpret_cast<const ulonglong2*>(wrow + kk * W);
                        const ulonglong2 n1 = *reinterpret_cast<const ulonglong2*>(wrow + kk * W + 128);
                        const unsigned long long wp[4] = {w0.x, w0.y, w1.x, w1.y};
#pragma unroll
                        for (int r = 0; r < ROWS; ++r) {
                            const float4 a4 = half == 0 ? aA[r] : aB[r];
                            const float av = jj == 0 ? a4.x : jj == 1 ? a4.y : jj == 2 ? a4.z : a4.w;
                            const unsigned long long ap = pack2(av, av);
#pragma unroll
                            for (int c = 0; c < 4; ++c) acc[r][c] = ffma2(ap, wp[c], acc[r][c]);
                        }
                        w0 = n0; w1 = n1;
                    }
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) { float x, y; unpack2(acc[r][c], x, y); s += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---------------- C: FFMA2, k-pairs (weights k-interleaved [k/2][col][2]) ---------------------------
template <int ROWS>
__global__ void __launch_bounds__(ROWS == 10 ? 256 : 512, 1) k_packed_kpair(float* out, int iters) {
    extern __shared__ __align__(16) float sm[];
    float* wch = sm;                                   // [KC/2][W][2]
    float* act = sm + KC * W + (threadIdx.x >> 5) * ROWS * S;
    for (int i = threadIdx.x; i < KC * W; i += blockDim.x) wch[i] = 1e-3f * (float)((i * 7) % 13 - 6);
    for (int i = threadIdx.x & 31; i < ROWS * S; i += 32) act[i] = 1e-2f * (float)((i * 5) % 11 - 5);
    __syncthreads();
    const int cg = threadIdx.x & 31;
    const float* wrow = wch + 8 * cg;                  // 4 cols x 2 k = 8 floats; second block at +256
    unsigned long long acc[ROWS][8];
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0ull;
    for (int it = 0; it < iters; ++it) {
        for (int k0 = 0; k0 < W; k0 += KC) {
            const float* arow = act + k0;
            for (int j = 0; j < KC; j += 4) {
                ulonglong2 ap[ROWS];                   // (a(k),a(k+1)), (a(k+2),a(k+3))
#pragma unroll
                for (int r = 0; r < ROWS; ++r) ap[r] = *reinterpret_cast<const ulonglong2*>(arow + r * S + j);
#pragma unroll
                for (int jp = 0; jp < 2; ++jp) {
                    const float* wp = wrow + ((j >> 1) + jp) * (2 * W);
                    const ulonglong2 wa = *reinterpret_cast<const ulonglong2*>(wp), wb = *reinterpret_cast<const ulonglong2*>(wp + 4);
                    const ulonglong2 wc = *reinterpret_cast<const ulonglong2*>(wp + 256), wd = *reinterpret_cast<const ulonglong2*>(wp + 260);
                    const unsigned long long wv[8] = {wa.x, wa.y, wb.x, wb.y, wc.x, wc.y, wd.x, wd.y};
#pragma unroll
                    for (int r = 0; r < ROWS; ++r) {
                        const unsigned long long a2 = jp == 0 ? ap[r].x : ap[r].y;
#pragma unroll
                        for (int c = 0; c < 8; ++c) acc[r][c] = ffma2(a2, wv[c], acc[r][c]);
                    }
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) { float x, y; unpack2(acc[r][c], x, y); s += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K>
static void run(const char* name, K kernel, int warps, int rows, int iters, int sms) {
    const size_t smem = sizeof(float) * (KC * W + (size_t)warps * rows * S) + 64;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    float* out;
    cudaMalloc(&out, sizeof(float) * sms * warps * 32);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a);
        kernel<<<sms, warps * 32, smem>>>(out, iters);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (rep) best = ms < best ? ms : best;
    }
    cudaError_t e = cudaGetLastError();
    const double flop = 2.0 * rows * 256.0 * 256.0 * warps * (double)iters * sms;
    printf("%-28s warps/SM %2d rows/thread %2d : %7.2f TFLOP/s  (%.3f ms) %s\n", name, warps, rows, flop / (best * 1e-3) / 1e12, best,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(out);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, iters = 400;
    run("A scalar 10x8", k_scalar<10>, 8, 10, iters, sms);
    run("E scalar 10x8 + barrier/chunk", k_scalar<10, 1>, 8, 10, iters, sms);
    run("D scalar 5x8", k_scalar<5>, 16, 5, iters, sms);
    run("D scalar 5x8 (8 warps)", k_scalar<5>, 8, 5, iters, sms);
    run("B ffma2 dup 10x8", k_packed_dup<10>, 8, 10, iters, sms);
    run("B ffma2 dup 5x8", k_packed_dup<5>, 16, 5, iters, sms);
    run("C ffma2 kpair 10x8", k_packed_kpair<10>, 8, 10, iters, sms);
    run("C ffma2 kpair 5x8", k_packed_kpair<5>, 16, 5, iters, sms);
    return 0;
}
