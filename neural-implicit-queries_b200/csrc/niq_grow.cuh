// niq_grow.cuh -- bound propagation for the modes whose affine form GROWS: affine_all and affine_truncate
// (reference src/affine.py:164-193 `apply_linear_approx`, :127-162 `truncate_affine`).
//
// One CTA per box; the state matrix S = [aff_1..aff_k] (k x W floats, k grows by `out_dim` per activation),
// base, err live in shared memory (up to 132 KB for the 8x64 nets in affine_all).  A dense layer is the
// row-wise product S@A, done in place: the 4 rows of a register tile are owned by lanes of ONE warp, so
// __syncwarp() orders the read-all / write-back; weights are read through L1 (they are shared by every CTA).
// Radius reduction, (alpha,beta,delta), row scaling, diag(delta) append and the top-n_keep selection are
// block-wide phases separated by __syncthreads().
#pragma once
#include "niq_kernels.cuh"

namespace niq {

struct GrowArgs {
    BoxSource src;
    long long n;
    float offset;
    int truncate;          // 1: affine_truncate, 0: affine_all
    int n_keep;
    int kcap;              // row capacity of the aff matrix
    int W;                 // padded row width (max out_pad / in_pad over layers, multiple of 8)
    int* label; float* lower; float* upper; unsigned char* near_tie;
};

__device__ __forceinline__ int next_pow2(int x) { int p = 1; while (p < x) p <<= 1; return p; }

__global__ void __launch_bounds__(256, 1) k_classify_grow(const __grid_constant__ NetDev net, const GrowArgs g) {
    extern __shared__ __align__(16) float sm[];
    const int W = g.W;
    float* base = sm;                 // [W]
    float* err = base + W;            // [W]
    float* err2 = err + W;            // [W]
    float* alpha = err2 + W;          // [W]
    float* delta = alpha + W;         // [W]
    float* red = delta + W;           // [8][W] partial radius sums
    float* mags = red + 8 * W;        // [kcap]
    int* rank = reinterpret_cast<int*>(mags + g.kcap);   // [kcap]
    float* aff = reinterpret_cast<float*>(rank + g.kcap); // [kcap][W]
    float* tmp = aff + (size_t)g.kcap * W;               // [n_keep][W] (truncate only)
    __shared__ float s_fin[3];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (long long box = blockIdx.x; box < g.n; box += gridDim.x) {
        // ---- input form (reference src/affine.py:109-117, non-interval modes): aff = vecs, err = 0 ----
        int k;
        {
            __syncthreads();
            if (tid == 0) {
                if (g.src.kind == 0) {
                    const float* c = g.src.a + 3 * box;
                    base[0] = c[0]; base[1] = c[1]; base[2] = c[2]; base[3] = 0.f;
                    for (int r = 0; r < g.src.v; ++r) {
                        const float* p = g.src.b + (box * g.src.v + r) * 3;
                        aff[r * W + 0] = p[0]; aff[r * W + 1] = p[1]; aff[r * W + 2] = p[2]; aff[r * W + 3] = 0.f;
                    }
                } else {
                    float4 rows[5];
                    BoxSource s = g.src;
                    s.interval = 0;
                    load_box_rows(s, box, rows);
                    base[0] = rows[0].x; base[1] = rows[0].y; base[2] = rows[0].z; base[3] = 0.f;
                    for (int r = 0; r < 3; ++r) {
                        aff[r * W + 0] = rows[1 + r].x; aff[r * W + 1] = rows[1 + r].y;
                        aff[r * W + 2] = rows[1 + r].z; aff[r * W + 3] = 0.f;
                    }
                }
                err[0] = err[1] = err[2] = err[3] = 0.f;
            }
            k = g.src.kind == 0 ? g.src.v : 3;
            __syncthreads();
        }

        float* e_cur = err;
        float* e_nxt = err2;
        for (int l = 0; l < net.n_layers; ++l) {
            const LayerDev& L = net.layers[l];
            const float* A = net.chunks[L.chunk_begin].src;
            if (!L.last_of_net) {
                const int K = L.in_pad, N = L.out_pad;
                // -- err row: e_nxt = e_cur @ |A| --
                for (int c = tid; c < N; c += blockDim.x) {
                    float s = 0.f;
                    for (int j = 0; j < K; ++j) s = fmaf(e_cur[j], fabsf(__ldg(A + (size_t)j * N + c)), s);
                    e_nxt[c] = s;
                }
                // -- base and aff rows, in place; row R = 0 is base, R >= 1 is aff[R-1] --
                const int cgw = N >> 2;                       // column groups of 4
                const int cgp = next_pow2(cgw) < 32 ? next_pow2(cgw) : 32;
                const int rg_per_warp = 32 / cgp;
                const int my_rg = lane / cgp, my_cg = lane % cgp;
                const int n_rows = k + 1;
                const int rows_per_iter = 8 * rg_per_warp * 4;
                for (int r0 = 0; r0 < n_rows; r0 += rows_per_iter) {
                    const int rbase = r0 + (warp * rg_per_warp + my_rg) * 4;
                    for (int cg0 = 0; cg0 < cgw; cg0 += cgp) {      // cgw > 32 never happens (N <= 128)
                        const int cg = cg0 + my_cg;
                        const bool act_thread = cg < cgw;
                        float acc[4][4];
#pragma unroll
                        for (int r = 0; r < 4; ++r)
#pragma unroll
                            for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
                        const float* rp[4];
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            const int R = rbase + r;
                            rp[r] = R == 0 ? base : (R < n_rows ? aff + (size_t)(R - 1) * W : base);
                        }
                        if (act_thread) {
                            for (int j = 0; j < K; j += 4) {
                                float4 a4[4];
#pragma unroll
                                for (int r = 0; r < 4; ++r) a4[r] = *reinterpret_cast<const float4*>(rp[r] + j);
#pragma unroll
                                for (int jj = 0; jj < 4; ++jj) {
                                    const float4 w = __ldg(reinterpret_cast<const float4*>(A + (size_t)(j + jj) * N + 4 * cg));
#pragma unroll
                                    for (int r = 0; r < 4; ++r) {
                                        const float a = jj == 0 ? a4[r].x : jj == 1 ? a4[r].y : jj == 2 ? a4[r].z : a4[r].w;
                                        acc[r][0] = fmaf(a, w.x, acc[r][0]);
                                        acc[r][1] = fmaf(a, w.y, acc[r][1]);
                                        acc[r][2] = fmaf(a, w.z, acc[r][2]);
                                        acc[r][3] = fmaf(a, w.w, acc[r][3]);
                                    }
                                }
                            }
                        }
                        __syncwarp();
                        if (act_thread) {
#pragma unroll
                            for (int r = 0; r < 4; ++r) {
                                const int R = rbase + r;
                                if (R < n_rows) {
                                    float* dst = R == 0 ? base : aff + (size_t)(R - 1) * W;
                                    float4 o = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
                                    if (R == 0) {
                                        const float4 b = __ldg(reinterpret_cast<const float4*>(L.bias + 4 * cg));
                                        o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                                    }
                                    *reinterpret_cast<float4*>(dst + 4 * cg) = o;
                                }
                            }
                        }
                        __syncwarp();
                    }
                }
                { float* t2 = e_cur; e_cur = e_nxt; e_nxt = t2; }
                __syncthreads();

                if (L.act != ACT_NONE) {
                    // -- radius per neuron: rad[c] = sum_r |aff[r][c]| + err[c] (partials over 8 row strides) --
                    const int parts = 256 / next_pow2(N) >= 8 ? 8 : (256 / next_pow2(N) > 0 ? 256 / next_pow2(N) : 1);
                    {
                        const int c = tid % next_pow2(N), p = tid / next_pow2(N);
                        if (c < N && p < parts) {
                            float s = 0.f;
                            for (int r = p; r < k; r += parts) s += fabsf(aff[(size_t)r * W + c]);
                            red[p * W + c] = s;
                        }
                    }
                    __syncthreads();
                    if (tid < N) {
                        const int c = tid;
                        float rad = 0.f;
                        for (int p = 0; p < parts; ++p) rad += red[p * W + c];
                        rad += e_cur[c];
                        const float b0 = base[c];
                        float al, be, de;
                        if (L.act == ACT_RELU) relu_lin(b0 - rad, b0 + rad, al, be, de);
                        else elu_lin(b0 - rad, b0 + rad, al, be, de);
                        base[c] = al * b0 + be;
                        e_cur[c] = al * e_cur[c];
                        alpha[c] = al;
                        delta[c] = de;
                    }
                    __syncthreads();
                    // -- scale rows by alpha, append diag(delta) (out_dim new rows) --
                    for (int idx = tid; idx < k * N; idx += blockDim.x) {
                        const int r = idx / N, c = idx - r * N;
                        aff[(size_t)r * W + c] = alpha[c] * aff[(size_t)r * W + c];
                    }
                    for (int idx = tid; idx < L.out_dim * N; idx += blockDim.x) {
                        const int r = idx / N, c = idx - r * N;
                        aff[(size_t)(k + r) * W + c] = (r == c) ? delta[c] : 0.f;
                    }
                    k += L.out_dim;
                    __syncthreads();

                    if (g.truncate && k > g.n_keep) {
                        // -- keep the n_keep rows of largest L1 norm, stable (reference src/affine.py:127-162) --
                        for (int r = tid; r < k; r += blockDim.x) {
                            float s = 0.f;
                            for (int c = 0; c < N; ++c) s += fabsf(aff[(size_t)r * W + c]);
                            mags[r] = s;
                        }
                        __syncthreads();
                        for (int r = tid; r < k; r += blockDim.x) {
                            const float m = mags[r];
                            int rk = 0;
                            for (int q = 0; q < k; ++q) rk += (mags[q] > m) || (mags[q] == m && q < r);
                            rank[r] = rk;
                        }
                        __syncthreads();
                        if (tid < N) {
                            float s = 0.f;
                            for (int r = 0; r < k; ++r)
                                if (rank[r] >= g.n_keep) s += fabsf(aff[(size_t)r * W + tid]);
                            e_cur[tid] = e_cur[tid] + s;
                        }
                        for (int idx = tid; idx < k * N; idx += blockDim.x) {
                            const int r = idx / N, c = idx - r * N;
                            if (rank[r] < g.n_keep) tmp[(size_t)rank[r] * W + c] = aff[(size_t)r * W + c];
                        }
                        __syncthreads();
                        for (int idx = tid; idx < g.n_keep * N; idx += blockDim.x) {
                            const int r = idx / N, c = idx - r * N;
                            aff[(size_t)r * W + c] = tmp[(size_t)r * W + c];
                        }
                        k = g.n_keep;
                        __syncthreads();
                    }
                }
            } else {
                // ---- final dot layer: scalar base, k scalar coefficients, scalar err ----
                const int K = L.in_pad;
                float part = 0.f;
                for (int r = tid; r < k; r += blockDim.x) {
                    float s = 0.f;
                    for (int j = 0; j < K; ++j) s = fmaf(aff[(size_t)r * W + j], __ldg(A + j), s);
                    part += fabsf(s);
                }
                // block reduce |coefficients|
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
                if (lane == 0) red[warp] = part;
                if (tid == 32) {
                    float s = 0.f, sa = 0.f;
                    for (int j = 0; j < K; ++j) { s = fmaf(base[j], __ldg(A + j), s); sa = fmaf(fabsf(base[j]), fabsf(__ldg(A + j)), sa); }
                    s_fin[0] = s + __ldg(L.bias);
                    s_fin[2] = sa + fabsf(__ldg(L.bias));
                }
                if (tid == 64) {
                    float s = 0.f;
                    for (int j = 0; j < K; ++j) s = fmaf(e_cur[j], fabsf(__ldg(A + j)), s);
                    s_fin[1] = s;
                }
                __syncthreads();
                if (tid == 0) {
                    float rad = 0.f;
                    for (int w = 0; w < 8; ++w) rad += red[w];
                    rad += s_fin[1];
                    const float lo = s_fin[0] - rad, up = s_fin[0] + rad;
                    if (g.lower) g.lower[box] = lo;
                    if (g.upper) g.upper[box] = up;
                    if (g.label) g.label[box] = label_of(lo, up, g.offset);
                    if (g.near_tie) g.near_tie[box] = bound_near_tie(lo, up, g.offset, s_fin[2], net.tie_rel) ? 1 : 0;
                }
                __syncthreads();
            }
        }
    }
}

}  // namespace niq
