// niq_grow.cuh -- bound propagation for the modes whose affine form GROWS: affine_all and affine_truncate
// (reference src/affine.py:164-193 `apply_linear_approx`, :127-162 `truncate_affine`).
//
// One CTA per box; the state matrix S = [aff_1..aff_k] (k x W floats, k grows by `out_dim` per activation),
// base, err live in shared memory (up to 132 KB for the 8x64 nets in affine_all).  A dense layer is the
// row-wise product S@A, done in place: the 4 rows of a register tile are owned by lanes of ONE warp, so
// __syncwarp() orders the read-all / write-back; weights are read through L1 (they are shared by every CTA).
// The two vector rows (base, err) are K-split over thread groups.  Radius reduction, (alpha,beta,delta), row
// scaling (one warp per row: conflict-free, yields the row's L1 norm for free), diag(delta) append and the stable
// top-n_keep selection (two threads per row for the O(k^2) rank, kept rows moved to a second buffer that then swaps
// with the first) are block-wide phases separated by __syncthreads().  Queries call it on small frontiers
// (<= a few hundred boxes), so the design goal is the latency of ONE box.
#pragma once
#include "niq_kernels.cuh"

namespace niq {

struct GrowArgs {
    BoxSource src;
    long long n;
    float offset;
    int truncate;          // 1: affine_truncate, 0: affine_all / affine_append
    int n_keep;
    int n_append;          // > 0: affine_append (reference src/affine.py:183-191)
    int kcap;              // row capacity of the aff matrix
    int W;                 // padded row width (max out_pad / in_pad over layers, multiple of 8)
    int* label; float* lower; float* upper; unsigned char* near_tie;
};

__device__ __forceinline__ int next_pow2(int x) { int p = 1; while (p < x) p <<= 1; return p; }

// warp-level sum of one value per lane
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

struct GrowCfg {          // what the growing-form propagation needs to know about the mode
    int truncate, n_keep, n_append, kcap, W;
    int wfloats;          // floats of the largest layer's weights (in_pad x out_pad): the staging buffer of grow_forward
};
struct GrowState {        // shared-memory carve-up of one CTA (see grow_carve); aff / tmp swap on every truncation
    float *base, *base2, *err, *err2, *alpha, *delta, *red, *mags;
    int* rank;
    float *aff, *tmp;
    float* wbuf;          // [wfloats] the current layer's weights, prefetched with cp.async while the previous layer's activation runs
};
__host__ __device__ inline size_t grow_state_floats(const GrowCfg& g) {
    return (size_t)22 * g.W + 2 * (size_t)g.kcap + (size_t)g.kcap * g.W * (g.truncate ? 2 : 1) + (size_t)g.wfloats + 16;
}
__device__ __forceinline__ void grow_carve(float* sm, const GrowCfg& g, GrowState& s) {
    const int W = g.W;
    s.base = sm;                  // [W]
    s.base2 = s.base + W;         // [W]
    s.err = s.base2 + W;          // [W]
    s.err2 = s.err + W;           // [W]
    s.alpha = s.err2 + W;         // [W]
    s.delta = s.alpha + W;        // [W]
    s.red = s.delta + W;          // [16][W] partial sums (two vectors x 8 parts)
    s.mags = s.red + 16 * W;      // [kcap]
    s.rank = reinterpret_cast<int*>(s.mags + g.kcap);     // [kcap]
    s.aff = reinterpret_cast<float*>(s.rank + g.kcap);    // [kcap][W]
    s.tmp = s.aff + (size_t)g.kcap * W;                   // [kcap][W] (truncate only): the truncated state is built here, then the two swap
    s.wbuf = s.aff + (size_t)g.kcap * W * (g.truncate ? 2 : 1);
}

// cp.async (16 B per thread and request) of `n_floats` (a multiple of 4, 16-byte aligned source) into shared memory
__device__ __forceinline__ void grow_stage_weights(float* dst, const float* src, int n_floats) {
    const uint32_t d0 = smem_u32(dst);
    for (int i = threadIdx.x * 4; i < n_floats; i += 256 * 4)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + i * 4), "l"(src + i) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void grow_wait_weights() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Bound propagation of ONE general box through layers [l0, l1) of `net` by the whole CTA (256 threads).
// Precondition: the caller wrote the input form -- base[0..3] = centre (4th entry 0), aff rows [0, k) = the box vectors
// (4 floats each), err[0..3] = 0 (reference src/affine.py:109-117, non-interval modes) -- and synchronised the CTA.
// A0 / b0 (optional): weights / bias that replace those of layer l0 (a per-query spatial_transformation, see
// niq_find_any_intersection_batch).  On return every thread holds lower / upper / scale (= sum_j |base_j A_j| + |b| of the
// last dot product, the near-tie yardstick).
// Point rows (optional, hA != nullptr): warp p additionally evaluates f at the point whose coordinates the caller put in
// hA[p*W + 0..3] = (x, y, z, 0), riding on the same staged weights (hA / hB: [8][W] ping-pong rows, private to their warp).
// The arithmetic is that of the engine's point rows (niq_engine.cuh, reference src/mlp.py:253-347): a hidden neuron
// accumulates fma(h_k, A_kc, acc) for ascending k from 0, then + bias, then the activation; the final dot product is split
// over `cg_lanes` (= width class / 8) lanes -- lane cg takes k = cg, cg + CG, ... -- and combined by an xor butterfly, like
// Engine::dot_layer.  pt_f / pt_scale return f and sum_k |h_k w_k| + |b| of point `warp` in every lane of that warp.
__device__ __forceinline__ void grow_forward(const NetDev& net, int l0, int l1, const float* A0, const float* b0, const GrowCfg& g,
                                             GrowState& st, int k, float& lower, float& upper, float& scale,
                                             float* hA = nullptr, float* hB = nullptr, int cg_lanes = 0, float* pt_f = nullptr,
                                             float* pt_scale = nullptr) {
    __shared__ float s_fin[3];
    __shared__ float s_rest;
    const int W = g.W;
    float* base = st.base; float* base2 = st.base2; float* err = st.err; float* err2 = st.err2;
    float* alpha = st.alpha; float* delta = st.delta; float* red = st.red; float* mags = st.mags; int* rank = st.rank;
    float* aff = st.aff; float* tmp = st.tmp;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        float* b_cur = base;
        float* b_nxt = base2;
        float* e_cur = err;
        float* e_nxt = err2;
        // The weights of a layer are read by every thread many times; they come through shared memory: layer l + 1 is
        // prefetched (cp.async) right after the dense phase of layer l, so the copy overlaps the activation / truncation phases.
        float* const A = st.wbuf;
        {
            const LayerDev& F = net.layers[l0];
            grow_stage_weights(A, A0 != nullptr ? A0 : net.chunks[F.chunk_begin].src, F.in_pad * (F.last_of_net ? 1 : F.out_pad));
        }
        for (int l = l0; l < l1; ++l) {
            const LayerDev& L = net.layers[l];
            const float* bias = (l == l0 && b0 != nullptr) ? b0 : L.bias;
            grow_wait_weights();
            __syncthreads();
            if (!L.last_of_net) {
                const int K = L.in_pad, N = L.out_pad;
                const int Np2 = next_pow2(N);
                const int lgN = 31 - __clz(Np2);                 // Np2 is a power of two: (tid % Np2, tid / Np2) by mask and shift
                const int parts = 256 / Np2 >= 8 ? 8 : (256 / Np2 > 0 ? 256 / Np2 : 1);
                const int c_of_tid = tid & (Np2 - 1), p_of_tid = tid >> lgN;
                // -- the two vector rows, K split over `parts` thread groups: base@A and err@|A| (partials in red) --
                {
                    const int c = c_of_tid, p = p_of_tid;
                    if (c < N && p < parts) {
                        float sb = 0.f, se = 0.f;
#pragma unroll 4
                        for (int j = p; j < K; j += parts) {
                            const float w = A[j * N + c];
                            sb = fmaf(b_cur[j], w, sb);
                            se = fmaf(e_cur[j], fabsf(w), se);
                        }
                        red[p * W + c] = sb;
                        red[(8 + p) * W + c] = se;
                    }
                }
                // -- aff rows, in place; a register tile = 4 rows x 4 columns, the 4 rows owned by lanes of ONE warp --
                const int cgw = N >> 2;                       // column groups of 4
                const int cgp = next_pow2(cgw) < 32 ? next_pow2(cgw) : 32;
                const int rg_per_warp = 32 / cgp;
                const int my_rg = lane / cgp, my_cg = lane % cgp;
                const int rows_per_iter = 8 * rg_per_warp * 4;
                for (int r0 = 0; r0 < k; r0 += rows_per_iter) {
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
                        for (int r = 0; r < 4; ++r) rp[r] = aff + (size_t)(rbase + r < k ? rbase + r : 0) * W;
                        if (act_thread) {
                            for (int j = 0; j < K; j += 4) {
                                float4 a4[4];
#pragma unroll
                                for (int r = 0; r < 4; ++r) a4[r] = *reinterpret_cast<const float4*>(rp[r] + j);
#pragma unroll
                                for (int jj = 0; jj < 4; ++jj) {
                                    const float4 w = *reinterpret_cast<const float4*>(A + (j + jj) * N + 4 * cg);
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
                            for (int r = 0; r < 4; ++r)
                                if (rbase + r < k)
                                    *reinterpret_cast<float4*>(aff + (size_t)(rbase + r) * W + 4 * cg) =
                                        make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
                        }
                        __syncwarp();
                    }
                }
                if (hA != nullptr) {
                    // -- point rows: warp p pushes point p through this layer (its rows of hin / hout are private to the warp) --
                    float* hin = (((l - l0) & 1) ? hB : hA) + warp * W;
                    float* hout = (((l - l0) & 1) ? hA : hB) + warp * W;
                    float pacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
                    for (int kk = 0; kk < K; ++kk) {
                        const float a = hin[kk];
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (lane + 32 * i < N) pacc[i] = fmaf(a, A[kk * N + lane + 32 * i], pacc[i]);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int c = lane + 32 * i;
                        if (c < N) {
                            float x = pacc[i] + __ldg(bias + c);
                            if (L.act == ACT_RELU) x = fmaxf(x, 0.f);
                            else if (L.act == ACT_ELU) x = elu_pt(x);
                            else if (L.act == ACT_SIN) x = sinf(x);
                            else if (L.act == ACT_TANH) x = tanhf(x);
                            hout[c] = x;
                        }
                    }
                }
                __syncthreads();
                if (l + 1 < l1) {      // every read of this layer's weights is done: fetch the next layer's behind the phases below
                    const LayerDev& Nx = net.layers[l + 1];
                    grow_stage_weights(A, net.chunks[Nx.chunk_begin].src, Nx.in_pad * (Nx.last_of_net ? 1 : Nx.out_pad));
                }
                // -- finish the vector rows --
                if (tid < N) {
                    float sb = 0.f, se = 0.f;
                    for (int p = 0; p < parts; ++p) { sb += red[p * W + tid]; se += red[(8 + p) * W + tid]; }
                    b_nxt[tid] = sb + __ldg(bias + tid);
                    e_nxt[tid] = se;
                }
                { float* t2 = e_cur; e_cur = e_nxt; e_nxt = t2; t2 = b_cur; b_cur = b_nxt; b_nxt = t2; }
                __syncthreads();

                if (L.act != ACT_NONE) {
                    // -- radius per neuron: rad[c] = sum_r |aff[r][c]| + err[c] (partials over row strides) --
                    {
                        const int c = c_of_tid, p = p_of_tid;
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
                        const float b0 = b_cur[c];
                        float al, be, de;
                        if (L.act == ACT_RELU) relu_lin(b0 - rad, b0 + rad, al, be, de);
                        else if (L.act == ACT_ELU) elu_lin(b0 - rad, b0 + rad, al, be, de);
                        else if (L.act == ACT_TANH) tanh_lin(b0 - rad, b0 + rad, al, be, de);
                        else sin_lin(b0 - rad, b0 + rad, al, be, de);
                        b_cur[c] = al * b0 + be;
                        e_cur[c] = al * e_cur[c];
                        alpha[c] = al;
                        delta[c] = de;
                    }
                    __syncthreads();
                    if (g.n_append > 0) {
                        // ---- affine_append (reference src/affine.py:183-191): scale the rows; the n_append largest
                        // deltas (jax.lax.top_k: descending, lower index first among equals) become single-entry rows;
                        // err += sum(delta) - sum(kept), ONE scalar on every neuron, as the reference writes it ----
                        const int na = g.n_append, w = L.out_dim;
                        int* drank = reinterpret_cast<int*>(red);          // [w] ranks of the deltas (red's partials are consumed)
                        for (int r = warp; r < k; r += 8) {
                            float* row = aff + (size_t)r * W;
                            for (int c = lane; c < N; c += 32) row[c] = alpha[c] * row[c];
                        }
                        if (tid < w) {
                            const float d = delta[tid];
                            int rk = 0;
                            for (int q = 0; q < w; ++q) { const float dq = delta[q]; rk += (dq > d) || (dq == d && q < tid); }
                            drank[tid] = rk;
                            if (rk < na) mags[rk] = d;
                        }
                        for (int r = warp; r < na; r += 8)
                            for (int c = lane; c < N; c += 32) aff[(size_t)(k + r) * W + c] = 0.f;
                        __syncthreads();
                        if (tid < w && drank[tid] < na) aff[(size_t)(k + drank[tid]) * W + tid] = delta[tid];
                        if (warp == 0) {
                            // np.sum order of the oracle (8 strided accumulators, pairwise, sequential remainder) for sum(delta);
                            // the kept values are summed in rank order
                            const int w8 = w & ~7;
                            float s = 0.f;
                            if (lane < 8)
                                for (int c = lane; c < w8; c += 8) s += delta[c];
                            s += __shfl_xor_sync(0xffffffffu, s, 1);
                            s += __shfl_xor_sync(0xffffffffu, s, 2);
                            s += __shfl_xor_sync(0xffffffffu, s, 4);
                            if (lane == 0) {
                                for (int c = w8; c < w; ++c) s += delta[c];
                                float kp = 0.f;
                                for (int r = 0; r < na; ++r) kp += mags[r];
                                s_rest = s - kp;
                            }
                        }
                        __syncthreads();
                        if (tid < w) e_cur[tid] = e_cur[tid] + s_rest;
                        k += na;
                        __syncthreads();
                        continue;
                    }
                    // -- scale rows by alpha (one warp per row: also yields the row's L1 norm), append diag(delta) --
                    const bool trunc = g.truncate && k + L.out_dim > g.n_keep;
                    for (int r = warp; r < k; r += 8) {
                        float* row = aff + (size_t)r * W;
                        for (int c = lane; c < N; c += 32) row[c] = alpha[c] * row[c];
                        if (trunc) {
                            // L1 norm in the summation order of the oracle's np.sum over a contiguous float32 row (n <= 128:
                            // 8 strided accumulators, pairwise combine, sequential remainder), so that near-equal rows rank
                            // the same way on both sides (reference src/affine.py:143-151 sorts by this norm)
                            __syncwarp();
                            const int w = L.out_dim, w8 = w & ~7;
                            float s = 0.f;
                            if (lane < 8)
                                for (int c = lane; c < w8; c += 8) s += fabsf(row[c]);
                            s += __shfl_xor_sync(0xffffffffu, s, 1);
                            s += __shfl_xor_sync(0xffffffffu, s, 2);
                            s += __shfl_xor_sync(0xffffffffu, s, 4);
                            if (lane == 0) {
                                for (int c = w8; c < w; ++c) s += fabsf(row[c]);
                                mags[r] = s;
                            }
                        }
                    }
                    for (int r = warp; r < L.out_dim; r += 8) {              // row r of diag(delta): one warp per row
                        float* row = aff + (size_t)(k + r) * W;
                        for (int c = lane; c < N; c += 32) row[c] = (r == c) ? delta[c] : 0.f;
                    }
                    if (trunc && tid < L.out_dim) mags[k + tid] = fabsf(delta[tid]);   // L1 norm of a diag row
                    k += L.out_dim;
                    __syncthreads();

                    if (trunc) {
                        // -- keep the n_keep rows of largest L1 norm, stable (reference src/affine.py:127-162) --
                        // rank[r] = #rows that sort before r; two threads per row, each scans half of the rows
                        for (int r = tid; r < k; r += blockDim.x) rank[r] = 0;
                        if (tid < 4 && k + tid < g.kcap) mags[k + tid] = -1.f;       // padding of the float4 scan below (magnitudes are >= 0)
                        __syncthreads();
                        for (int r0 = 0; r0 < k; r0 += 128) {
                            const int r = r0 + (tid & 127), half = tid >> 7;
                            if (r < k) {
                                const float m = mags[r];
                                int rk = 0;
                                // kcap is a multiple of 4 and mags is 16-byte aligned: the half a thread scans starts at a multiple
                                // of 4 and is read as float4 (entries >= k hold -1: never counted)
                                const int q0 = half ? (((k + 1) / 2) & ~3) : 0, q1 = half ? k : (((k + 1) / 2) & ~3);
                                for (int q = q0; q < q1; q += 4) {
                                    const float4 mq = *reinterpret_cast<const float4*>(mags + q);
                                    rk += (mq.x > m) || (mq.x == m && q < r);
                                    rk += (mq.y > m) || (mq.y == m && q + 1 < r);
                                    rk += (mq.z > m) || (mq.z == m && q + 2 < r);
                                    rk += (mq.w > m) || (mq.w == m && q + 3 < r);
                                }
                                atomicAdd(&rank[r], rk);
                            }
                        }
                        __syncthreads();
                        // dropped rows fold into err (partials over row strides); kept rows move to tmp at their rank
                        {
                            const int c = c_of_tid, p = p_of_tid;
                            if (c < N && p < parts) {
                                float s = 0.f;
                                for (int r = p; r < k; r += parts)
                                    if (rank[r] >= g.n_keep) s += fabsf(aff[(size_t)r * W + c]);
                                red[p * W + c] = s;
                            }
                        }
                        for (int r = warp; r < k; r += 8) {
                            const int rk = rank[r];
                            if (rk < g.n_keep)
                                for (int c = lane; c < N; c += 32) tmp[(size_t)rk * W + c] = aff[(size_t)r * W + c];
                        }
                        __syncthreads();
                        if (tid < N) {
                            float s = 0.f;
                            for (int p = 0; p < parts; ++p) s += red[p * W + tid];
                            e_cur[tid] = e_cur[tid] + s;
                        }
                        { float* t2 = aff; aff = tmp; tmp = t2; }
                        k = g.n_keep;
                        __syncthreads();
                    }
                }
            } else {
                // ---- final dot layer: scalar base, k scalar coefficients, scalar err (one warp per row) ----
                const int K = L.in_pad;
                float part = 0.f;
                for (int r = warp; r < k; r += 8) {
                    float s = 0.f;
                    for (int j = lane; j < K; j += 32) s = fmaf(aff[(size_t)r * W + j], A[j], s);
                    s = warp_sum(s);
                    part += fabsf(s);
                }
                if (lane == 0) red[warp] = part;
                if (hA != nullptr) {
                    const float* hin = (((l - l0) & 1) ? hB : hA) + warp * W;
                    float po = 0.f, pp = 0.f;
                    if (lane < cg_lanes) {
                        for (int j = lane; j < K; j += cg_lanes) {
                            const float a = hin[j], w = A[j];
                            po = fmaf(a, w, po);
                            pp = fmaf(fabsf(a), fabsf(w), pp);
                        }
                    }
                    for (int off = 1; off < cg_lanes; off <<= 1) {
                        po += __shfl_xor_sync(0xffffffffu, po, off);
                        pp += __shfl_xor_sync(0xffffffffu, pp, off);
                    }
                    const float bb = __ldg(bias);
                    po += bb; pp += fabsf(bb);
                    *pt_f = __shfl_sync(0xffffffffu, po, 0);
                    *pt_scale = __shfl_sync(0xffffffffu, pp, 0);
                }
                if (warp == 0) {
                    float s = 0.f, sa = 0.f, se = 0.f;
                    for (int j = lane; j < K; j += 32) {
                        const float w = A[j];
                        s = fmaf(b_cur[j], w, s);
                        sa = fmaf(fabsf(b_cur[j]), fabsf(w), sa);
                        se = fmaf(e_cur[j], fabsf(w), se);
                    }
                    s = warp_sum(s); sa = warp_sum(sa); se = warp_sum(se);
                    if (lane == 0) {
                        s_fin[0] = s + __ldg(bias);
                        s_fin[2] = sa + fabsf(__ldg(bias));
                        s_fin[1] = se;
                    }
                }
                __syncthreads();
                {
                    float rad = 0.f;
                    for (int w = 0; w < 8; ++w) rad += red[w];
                    rad += s_fin[1];
                    lower = s_fin[0] - rad; upper = s_fin[0] + rad; scale = s_fin[2];
                }
                __syncthreads();
            }
        }
    }
    st.aff = aff; st.tmp = tmp;        // a truncation swapped the two buffers
}

__global__ void __launch_bounds__(256, 1) k_classify_grow(const __grid_constant__ NetDev net, const GrowArgs g) {
    extern __shared__ __align__(16) float sm[];
    GrowCfg cfg{g.truncate, g.n_keep, g.n_append, g.kcap, g.W};
    GrowState st;
    grow_carve(sm, cfg, st);
    const int tid = threadIdx.x;
    for (long long box = blockIdx.x; box < g.n; box += gridDim.x) {
        // ---- input form (reference src/affine.py:109-117, non-interval modes): aff = vecs, err = 0 ----
        __syncthreads();
        if (tid == 0) {
            float* base = st.base; float* aff = st.aff; float* err = st.err;
            const int W = g.W;
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
        __syncthreads();
        float lo, up, sc;
        grow_forward(net, 0, net.n_layers, nullptr, nullptr, cfg, st, g.src.kind == 0 ? g.src.v : 3, lo, up, sc);
        if (tid == 0) {
            if (g.lower) g.lower[box] = lo;
            if (g.upper) g.upper[box] = up;
            if (g.label) g.label[box] = label_of(lo, up, g.offset);
            if (g.near_tie) g.near_tie[box] = bound_near_tie(lo, up, g.offset, sc, net.tie_rel) ? 1 : 0;
        }
    }
}

}  // namespace niq
