// niq_engine.cuh -- the warp-tiled FP32 network engine shared by every query kernel (sm_100a).
//
// What it computes: rows of "state" pushed through the MLP, one dense layer at a time, exactly as the
// reference's affine interpreter does for one box (reference src/mlp.py:96-113 with the rules of
// src/affine_layers.py:11-97): for a box the rows are [base; aff_1..aff_v; err] and a dense layer is
//     base <- base@A + b ;  aff_r <- aff_r@A ;  err <- err@|A|
// followed by the per-neuron Chebyshev linearisation (alpha, beta, delta) of relu / elu, which needs
// rad = sum_r |aff_r| + err of that neuron.  Point rows (plain f(x), src/mlp.py:253-347) ride along.
//
// How it is laid out on a B200 SM:
//   * a CTA is NWARPS warps.  Each warp owns TPW*NT "tiles" (a tile = the RT rows of one box / ray / 8
//     points) in a private shared-memory activation buffer, updated IN PLACE layer by layer, so the only
//     intra-layer synchronisation is __syncwarp().
//   * inside a warp, lane = (tile_in_warp t, column group cg): the thread keeps ALL rows of its NT tiles
//     for 8 output neurons in registers (NT*RT*8 FP32 accumulators), so rad and (alpha,beta,delta) are
//     in-register work -- no shuffles in hidden layers.
//   * weights stream through a ring of shared-memory stages as K-chunks, one cp.async.bulk (TMA bulk
//     copy) per chunk, completion on an mbarrier; a stage is recycled after the __syncthreads() that
//     follows its consumption.  Weight reads are warp-broadcast LDS.128, activation reads LDS.128 along K.
//   * the final (out_dim = 1) layer is a dot product split across the CG lanes of a tile and reduced with
//     warp shuffles.
// All arithmetic is FP32 FFMA (tensor cores would change bounds by >> 1e-5 relative).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef NIQ_CHUNK_FLOATS
#define NIQ_CHUNK_FLOATS 16384
#define NIQ_STAGES 2
#endif

namespace niq {

enum : int { ACT_NONE = 0, ACT_RELU = 1, ACT_ELU = 2, ACT_SIN = 3, ACT_TANH = 4 };

constexpr int kMaxLayers = 32;     // layers of all nets of one launch (cast_rays concatenates funcs)
constexpr int kMaxChunks = 192;
constexpr int kChunkFloats = NIQ_CHUNK_FLOATS;   // per stage: 16384 = 64 KB (a 256-wide layer is 4 chunks; layers up to 128 wide are 1)
constexpr int kStages = NIQ_STAGES;
constexpr int kWarps = 8;          // compute warps per CTA
constexpr int kThreads = kWarps * 32;
constexpr int kMaxSegs = 8;        // weight chunks per layer a zero-skipping consumer can address

struct LayerDev {
    int in_dim, out_dim;      // logical
    int in_pad, out_pad;      // in_pad % 4 == 0 ; out_pad % 8 == 0 (hidden) or 1 (dot layer)
    int act;                  // ACT_*
    int chunk_begin, chunk_end;
    const float* bias;        // device, out_pad floats (zero padded)
    int first_of_net;         // 1: the loader must (re)write the input rows before this layer
    int last_of_net;          // 1: out_dim == 1, finalize after this layer
};

struct ChunkDev {
    const float* src;         // device, 16-byte aligned; rows [k0, k0+kc) of the layer's padded A, contiguous
    unsigned int n_floats;    // multiple of 4
    int k0, kc;
    int smem_off;             // resident mode: float offset of this chunk inside the shared-memory weight region
};

struct NetDev {               // passed by value (__grid_constant__) to every engine kernel
    int n_layers, n_chunks, n_nets;
    int resident;             // 1: ALL weights of the launch fit in shared memory and are loaded once (no ring,
                              //    no CTA barrier in the layer loop: warps run independently); 0: streamed ring
    int w_region_floats;      // resident: size of the weight region (incl. over-read padding)
    float tie_rel;            // near-tie band: 1e-5 (relu-only nets) or 2e-4 (nets with elu)
    int sparse;               // 1: drop exactly-zero columns after relu layers (write_back_sparse); 0: dense K loops
    int dephase;              // streamed ray kernels: cycles of head start the warps 0-3 get over warps 4-7 (same SM sub-partitions)
    unsigned long long* exec_macs;   // optional device counter of executed MACs / RT (see Engine::exec_macs)
    LayerDev layers[kMaxLayers];
    ChunkDev chunks[kMaxChunks];
};

// ------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA bulk copy (global -> shared)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Packed FP32 (Blackwell FFMA2, PTX fma.rn.f32x2): two IEEE fp32 FMAs per instruction on a 64-bit register pair.
// Same results as two fmaf(); half the issue slots and register-file reads per FMA.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float x, float y) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& x, float& y) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
}
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 abs2(f32x2 v) { return v & 0x7fffffff7fffffffull; }

// ------------------------------------------------------------------------------------------------
// Tile descriptors: which rows a thread carries and how the activation couples them.
//   row kinds: B = base, A = affine coefficient, E = interval error (multiplies |A|), P = point
// ------------------------------------------------------------------------------------------------
//   rule: 0 = affine group [base, aff.., err] (+ optional point rows), 1 = point rows only,
//         2 = slope interval [primal, slope centre x3, slope width x3] (reference src/slope_interval_layers.py)
struct TileBox3 {   // [B, A, A, A, E] : one general box with 3 symbols (affine_fixed, v<=3; interval uses E only)
    static constexpr int RT = 5, NT = 2, rule = 0;
    __host__ __device__ static constexpr bool is_err(int r) { return r == 4; }
    __host__ __device__ static constexpr bool has_bias(int r) { return r == 0; }
    __host__ __device__ static constexpr bool is_pt(int) { return false; }
    __host__ __device__ static constexpr bool want_scale(int r) { return r == 0; }
    static constexpr int n_aff = 3, n_pts = 0;
    static constexpr bool has_group = true;
};
struct TileRay {    // [B, A, E, P, P] : one ray step = segment bound (v=1) + f(start), f(start+eps)
    static constexpr int RT = 5, NT = 2, rule = 0;
    __host__ __device__ static constexpr bool is_err(int r) { return r == 2; }
    __host__ __device__ static constexpr bool has_bias(int r) { return r == 0 || r >= 3; }
    __host__ __device__ static constexpr bool is_pt(int r) { return r >= 3; }
    __host__ __device__ static constexpr bool want_scale(int r) { return r == 0 || r >= 3; }
    static constexpr int n_aff = 1, n_pts = 2;
    static constexpr bool has_group = true;
};
struct TileFrustum { // [B, A, A, A, E, P, P] : one frustum step = general-box bound (v=3) + f(start), f(start+eps) on the mid ray
    static constexpr int RT = 7, NT = 1, rule = 0;
    __host__ __device__ static constexpr bool is_err(int r) { return r == 4; }
    __host__ __device__ static constexpr bool has_bias(int r) { return r == 0 || r >= 5; }
    __host__ __device__ static constexpr bool is_pt(int r) { return r >= 5; }
    __host__ __device__ static constexpr bool want_scale(int r) { return r == 0 || r >= 5; }
    static constexpr int n_aff = 3, n_pts = 2;
    static constexpr bool has_group = true;
};
struct TileFrustumSlope { // [P, C, C, C, W, W, W, pt, pt] : one frustum step in slope_interval mode (3 box vectors) + the two points
    static constexpr int RT = 9, NT = 1, rule = 2, n_vec = 3;
    __host__ __device__ static constexpr bool is_err(int r) { return r >= 4 && r <= 6; }
    __host__ __device__ static constexpr bool has_bias(int r) { return r == 0 || r >= 7; }
    __host__ __device__ static constexpr bool is_pt(int r) { return r >= 7; }
    __host__ __device__ static constexpr bool want_scale(int r) { return r == 0 || r >= 7; }
    static constexpr int n_aff = 3, n_pts = 2;
    static constexpr bool has_group = false;
};
struct TilePts {    // [P x 8]
    static constexpr int RT = 8, NT = 1, rule = 1;
    __host__ __device__ static constexpr bool is_err(int) { return false; }
    __host__ __device__ static constexpr bool has_bias(int) { return true; }
    __host__ __device__ static constexpr bool is_pt(int) { return true; }
    __host__ __device__ static constexpr bool want_scale(int) { return true; }
    static constexpr int n_aff = 0, n_pts = 8;
    static constexpr bool has_group = false;
};
struct TileSlope3 { // [P, C, C, C, W, W, W] : primal, slope centres, slope widths of one general box with <= 3 vectors;
                    // the width rows multiply |A| like the err row (reference src/slope_interval_layers.py:11-33)
    static constexpr int RT = 7, NT = 1, rule = 2, n_vec = 3;
    __host__ __device__ static constexpr bool is_err(int r) { return r >= 4; }
    __host__ __device__ static constexpr bool has_bias(int r) { return r == 0; }
    __host__ __device__ static constexpr bool is_pt(int) { return false; }
    __host__ __device__ static constexpr bool want_scale(int r) { return r == 0; }
    static constexpr int n_aff = 3, n_pts = 0;
    static constexpr bool has_group = false;
};
struct TileRaySlope { // [P, C, W, pt, pt] : one ray step in slope_interval mode (v = 1) + f(start), f(start+eps)
    static constexpr int RT = 5, NT = 2, rule = 2, n_vec = 1;
    __host__ __device__ static constexpr bool is_err(int r) { return r == 2; }
    __host__ __device__ static constexpr bool has_bias(int r) { return r == 0 || r >= 3; }
    __host__ __device__ static constexpr bool is_pt(int r) { return r >= 3; }
    __host__ __device__ static constexpr bool want_scale(int r) { return r == 0 || r >= 3; }
    static constexpr int n_aff = 1, n_pts = 2;
    static constexpr bool has_group = false;
};

// the same rows with ONE tile per thread (niq_tree.cuh: the small levels of a tree)
template <class Tile>
struct TileOne : Tile { static constexpr int NT = 1; };

// ------------------------------------------------------------------------------------------------
// scalar activation rules
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float elu_f(float x) { return x > 0.f ? x : expm1f(x); }
// elu of a POINT row: exp(x) - 1 instead of expm1 (a third of the instructions; the point-evaluation kernels of elu
// nets are bound by this epilogue).  For x <= 0 the result differs from expm1 by at most 1 ulp of 1 (1.2e-7 ABSOLUTE),
// the size of the float32 noise the dense sums already carry -- two decades inside the 1e-5 band on point values.
__device__ __forceinline__ float elu_pt(float x) { return x > 0.f ? x : expf(x) - 1.f; }

// n / d without the branch to the IEEE slow path: reciprocal seed + one Newton step on the reciprocal + one
// residual correction of the quotient.  For normal, well-scaled operands (the rules below only divide
// 0 < n <= d or O(1) differences) this is the correctly rounded quotient except in rare 1-ulp cases --
// far inside every tolerance of this backend -- and it keeps the 16 neurons of a thread in
// one basic block.  inf / nan operands propagate to nan like the IEEE quotient inf/inf.
__device__ __forceinline__ float div_nr(float n, float d) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
    r = fmaf(fmaf(-d, r, 1.f), r, r);
    const float q = n * r;
    return fmaf(fmaf(-d, q, n), r, q);
}

// relu linearisation on [l,u]  (reference src/affine_layers.py:34-56)
__device__ __forceinline__ void relu_lin(float l, float u, float& alpha, float& beta, float& delta) {
    // alpha = (relu(u) - relu(l)) / (u - l), then the l >= 0 / u < 0 overrides.  Only the straddling case
    // l < 0 < u needs the quotient, and there relu(u) - relu(l) == u exactly; dividing 1/1 elsewhere keeps
    // the IEEE division on its fast path (0/x and x/0 take the slow subroutine) without changing any result:
    //   l >= 0 -> 1 ; u < 0 -> 0 ; u == 0 (l < 0) -> 0/(0-l) = 0 ; l == u == 0 is covered by l >= 0.
    const bool straddle = (l < 0.f) && (u > 0.f);
    float a = div_nr(straddle ? u : 1.f, straddle ? (u - l) : 1.f);
    if (!straddle) a = 0.f;                    // u <= 0 (and nan bounds: comparisons false -> nan_to_num -> 0)
    if (l >= 0.f) a = 1.f;
    if (u < 0.f) a = 0.f;                      // only reachable with l > u (a NEGATIVE radius: err turns negative after a sin
                                               // layer with alpha < 0, src/affine.py:176); the reference applies it last
    if (a != a) a = 0.f;                       // inf/inf -> nan -> 0 (nan_to_num(nan=0))
    a = fminf(fmaxf(a, 0.f), 1.f);             // also maps +inf -> 1 like nan_to_num + clip
    alpha = a;
    beta = (fmaxf(l, 0.f) - a * l) * 0.5f;
    delta = fabsf(beta);
}

// elu linearisation on [l,u]  (reference src/affine_layers.py:59-97)
__device__ __forceinline__ void elu_lin(float l, float u, float& alpha, float& beta, float& delta) {
    // One expm1 per endpoint serves both elu(x) = x > 0 ? x : expm1(x) and the slope min(exp(x), 1): for x < 0
    // exp(x) = expm1(x) + 1 (within 1 ulp of 1, i.e. <= 1.2e-7 ABSOLUTE of a bound that only clips alpha), for x >= 0
    // the min is 1.  Saves the two expf of the reference formulation; everything else is as written there.
    const float ml = expm1f(fminf(l, 0.f)), mu = expm1f(fminf(u, 0.f));
    const float lF = l > 0.f ? l : ml, uF = u > 0.f ? u : mu;
    const float lS = ml + 1.f, uS = mu + 1.f;
    float a = (uF - lF) / (u - l);
    if (l >= 0.f) a = 1.f;
    if (a != a) a = 0.f;
    a = fminf(fmaxf(a, -3.4028234664e38f), 3.4028234664e38f);   // nan_to_num maps +-inf to +-FLT_MAX
    a = fminf(fmaxf(a, lS), uS);
    const float r_up = lF - a * l;
    const float x_lo = fminf(fmaxf(logf(a), l), u);
    const float r_lo = (a - 1.f) - a * x_lo;
    float b = 0.5f * (r_up + r_lo);
    float d = 0.5f * fabsf(r_up - r_lo);
    if (l >= 0.f) { a = 1.f; b = 0.f; d = 0.f; }
    alpha = a; beta = b; delta = fabsf(d);
}

// bounds of cos on [lower, upper]  (reference src/utils.py:209-231: sin_bound(lower + pi/2, upper + pi/2); the Python
// float constants of the reference become float32 when they meet float32 arrays, so they are float32 literals here)
__device__ __forceinline__ void cos_bound(float lower, float upper, float& out_lo, float& out_hi) {
    const float kHalfPi = 1.5707963267948966f, kTwoPi = 6.283185307179586f;
    float l = lower + kHalfPi, u = upper + kHalfPi;
    const float fl = sinf(l), fu = sinf(u);
    l = l / kTwoPi;
    u = u / kTwoPi;
    const bool has_min = ceilf(l - 0.75f) < (u - 0.75f);
    const bool has_max = ceilf(l - 0.25f) < (u - 0.25f);
    out_lo = has_min ? -1.f : fminf(fl, fu);
    out_hi = has_max ? 1.f : fmaxf(fl, fu);
}
// sin linearisation on [l,u]  (reference src/affine_layers.py:100-137)
__device__ __forceinline__ void sin_lin(float l, float u, float& alpha, float& beta, float& delta) {
    const float kTwoPi = 6.283185307179586f;
    float s_lo, s_hi;
    cos_bound(l, u, s_lo, s_hi);
    float a = 0.5f * (s_lo + s_hi);
    a = fminf(fmaxf(a, -1.f), 1.f);
    const float iA = acosf(a), iB = -iA;
    float locs[6];
    locs[0] = l;
    locs[1] = u;
    locs[2] = kTwoPi * ceilf((l + iA) / kTwoPi) - iA;
    locs[3] = kTwoPi * floorf((u - iA) / kTwoPi) + iA;
    locs[4] = kTwoPi * ceilf((l + iB) / kTwoPi) - iB;
    locs[5] = kTwoPi * floorf((u - iB) / kTwoPi) + iB;
    float r_lo = 0.f, r_hi = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const float x = fminf(fmaxf(locs[i], l), u);
        const float v = sinf(x) - a * x;
        r_lo = i == 0 ? v : fminf(r_lo, v);
        r_hi = i == 0 ? v : fmaxf(r_hi, v);
    }
    const float b = 0.5f * (r_hi + r_lo);
    alpha = a; beta = b; delta = fabsf(r_hi - b);
}

// tanh linearisation on [l,u].  PARITY UNPINNED: the reference registers no tanh rule -- this is the Chebyshev-style
// construction of the paper written from scratch, the same formulas as the CPU checker's tanh coefficients (tests/test_tanh.py compares the two):
//   alpha = secant slope (tanh u - tanh l) / w, w = u - l, evaluated without the cancelling difference through
//           tanh u - tanh l = (1 - tanh u tanh l) tanh(w):  alpha = (1 - tanh u tanh l) * g(w), g = tanh(w)/w (1 - w^2/3 below 1e-3);
//   the residual tanh(x) - alpha x has its extrema over [l,u] at l, u or where tanh'(x) = alpha, x* = +-atanh(sqrt(1 - alpha));
//   beta = mid-range of the residual, delta = its half-range.
// Sound for any alpha in [0,1] because the residual bounds belong to the alpha actually used.
//   The residual is formed RELATIVE to its value at l, d(x) = (tanh x - tanh l) - alpha (x - l), with the tanh difference again in
//   product form, so its rounding noise is ~1 ulp of w instead of 1 ulp of |tanh l|; for intervals narrower than 1e-2, whose true
//   half-range O(w^2 |f''|) even d(x) cannot resolve, the secant-error theorem |f - secant| <= max|f''| w^2 / 8 bounds it instead
//   (one-sided where f'' keeps its sign).  Without these two steps float32 and float64 evaluations of the rule differ by up to
//   3.5e-2 of the bound on the 8 x 256 net (the per-neuron noise is amplified by every later layer); with them by 3.5e-6.
__device__ __forceinline__ float tanh_d(float x, float l, float tl, float a) {
    const float dx = x - l;
    return (1.f - tanhf(x) * tl) * tanhf(dx) - a * dx;
}
__device__ __forceinline__ void tanh_lin(float l, float u, float& alpha, float& beta, float& delta) {
    const float tl = tanhf(l), tu = tanhf(u);
    const float w = u - l;
    const float g = w < 1e-3f ? 1.f - w * w / 3.f : tanhf(w) / w;
    float a = (1.f - tu * tl) * g;
    if (a != a) a = 0.f;
    a = fminf(fmaxf(a, 0.f), 1.f);
    const bool zero = a == 0.f;
    const float du = zero ? tu - tl : tanh_d(u, l, tl, a);
    float d_lo = fminf(du, 0.f), d_hi = fmaxf(du, 0.f);
    const float xs = atanhf(sqrtf(fmaxf(1.f - a, 0.f)));
#pragma unroll
    for (int sgn = 0; sgn < 2; ++sgn) {
        const float x = fminf(fmaxf(sgn == 0 ? xs : -xs, l), u);           // clipped: an end point, already covered
        const float v = zero ? 0.f : tanh_d(x, l, tl, a);
        d_lo = fminf(d_lo, v); d_hi = fmaxf(d_hi, v);
    }
    const float rl = zero ? tl : tl - a * l;
    {
        const float atl = fabsf(tl), atu = fabsf(tu);
        const float t_hi = fmaxf(atl, atu);
        const float t_lo = (l <= 0.f && u >= 0.f) ? 0.f : fminf(atl, atu);
        const float peak = 0.5773502691896258f;
        const float f2lo = 2.f * t_lo * (1.f - t_lo * t_lo), f2hi = 2.f * t_hi * (1.f - t_hi * t_hi);
        const float M = (t_lo <= peak && t_hi >= peak) ? 0.7698003589195010f : fmaxf(f2lo, f2hi);
        const float bnd = M * w * w / 8.f;
        if (w < 1e-2f && !zero) {
            d_hi = l >= 0.f ? bnd : (u <= 0.f ? 0.f : bnd);
            d_lo = l >= 0.f ? 0.f : -bnd;
        }
    }
    alpha = a;
    beta = rl + 0.5f * (d_hi + d_lo);
    delta = fabsf(0.5f * (d_hi - d_lo));
}

// ------------------------------------------------------------------------------------------------
// Engine
// ------------------------------------------------------------------------------------------------
template <int WMAX>
struct Geometry {
    static_assert(WMAX == 32 || WMAX == 64 || WMAX == 128 || WMAX == 256, "hidden width class");
    static constexpr int CG = WMAX / 8;          // column groups (lanes) covering one row
    static constexpr int TPW = 32 / CG;          // tiles per warp pass (per NT slice)
    static constexpr int S = WMAX + 4;           // activation row stride in floats (bank-conflict padding)
    static constexpr int LOG_CG = (CG == 4 ? 2 : CG == 8 ? 3 : CG == 16 ? 4 : 5);
};

template <int WMAX, class Tile>
struct Engine {
    using G = Geometry<WMAX>;
    static constexpr int RT = Tile::RT, NT = Tile::NT;
    static constexpr int ROWS = RT * NT;                       // rows carried by one thread
    static constexpr int SLOTS = G::TPW * NT;                  // tiles per warp
    static constexpr int WARP_ROWS = SLOTS * RT;               // rows in one warp's activation buffer
    // A slot's rows are RT * S floats apart from the next slot's.  When that is a multiple of 32 floats (RT = 8: the point tile)
    // the TPW slots a warp reads together would sit in the SAME shared-memory banks (a 4-way conflict on every activation
    // LDS.128 of the 64-wide nets, measured: profiles/r2_hmc_kernels_ncu_full_summary.txt); one float4 of skew per slot
    // spreads them over distinct bank groups.
    static constexpr int TS = RT * G::S + (((RT * G::S) % 32 == 0 && G::TPW > 1) ? 4 : 0);   // slot stride in floats
    static constexpr int WARP_FLOATS = SLOTS * TS;
    static constexpr int CTA_TILES = kWarps * SLOTS;
    static constexpr int LIST_WORDS = WMAX + 16;               // per-warp live-row list (+ over-read tail)
    static constexpr int SEG_WORDS = 4 * (kMaxSegs + 1);
    // Zero-skipping pays when K is long and a warp holds few tiles: 128- and 256-wide nets (1-2 tiles per column group).
    // For the 32 / 64-wide nets (8-16 tiles per warp, K <= 64) the compaction costs more than it saves (measured).
    static constexpr bool kSparse = WMAX >= 128;             // per-warp (start, n8, cb0, cnt) per chunk of the consuming layer + sentinel

    static constexpr size_t smem_bytes(int w_region_floats = kStages * kChunkFloats) {
        return 64 + sizeof(float) * ((size_t)w_region_floats + kWarps * WARP_FLOATS + kWarps * SLOTS * 8 +
                                     kWarps * LIST_WORDS + kWarps * SEG_WORDS);
    }

    // shared-memory carve-up
    uint64_t* full;        // [kStages] mbarriers: chunk landed in the stage
    int* done;             // [kStages] warps that have finished reading the stage's current chunk
    float* stage;          // [kStages][kChunkFloats]
    float* act;            // this warp's [SLOTS] x ([RT][S] rows + skew)
    float* fin;            // this warp's [SLOTS][8] final scalars / scratch
    uint32_t* lst;         // this warp's [LIST_WORDS]: byte offset (inside its chunk) of the weight row of every live k
    int* seg;              // this warp's [kMaxSegs][4]: start, n8, count-before, count of every chunk of the consuming layer
    const NetDev& net;
    int warp, lane, t, cg;
    bool resident;
    // weight pipeline state: chunks of the stream this warp has acquired so far (every warp acquires the same sequence)
    unsigned int seq_consumed;
    unsigned long long exec_macs;   // tile-row MACs / RT actually executed by this warp (k-steps x out_pad x SLOTS), lane-uniform

    // primary = false: a SECOND engine over the same shared memory (another tile shape for another kind of pass of the same
    // kernel, niq_cp.cuh): it shares the resident weights the first one staged and only lays out its own activation buffers
    // (resident nets only; the two kinds of pass never run at the same time in one CTA).
    __device__ Engine(const NetDev& n, unsigned char* smem_raw, bool primary = true) : net(n) {
        warp = threadIdx.x >> 5;
        lane = threadIdx.x & 31;
        t = lane / G::CG;
        cg = lane % G::CG;
        full = reinterpret_cast<uint64_t*>(smem_raw);
        done = reinterpret_cast<int*>(smem_raw + 32);
        stage = reinterpret_cast<float*>(smem_raw + 64);
        resident = net.resident != 0;
        float* acts = stage + (resident ? net.w_region_floats : kStages * kChunkFloats);
        act = acts + warp * WARP_FLOATS;
        fin = acts + kWarps * WARP_FLOATS + warp * SLOTS * 8;
        lst = reinterpret_cast<uint32_t*>(acts + kWarps * WARP_FLOATS + kWarps * SLOTS * 8) + warp * LIST_WORDS;
        seg = reinterpret_cast<int*>(acts + kWarps * WARP_FLOATS + kWarps * SLOTS * 8 + kWarps * LIST_WORDS) + warp * SEG_WORDS;
        seq_consumed = 0;
        exec_macs = 0;
        if (!primary) return;
        for (int i = lane; i < LIST_WORDS; i += 32) lst[i] = 0u;   // every entry is always a valid in-chunk offset
        if (threadIdx.x == 0) {
            for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); done[s] = 0; }
            fence_barrier_init();
        }
        __syncthreads();
        if (resident) {
            // the whole weight set is staged ONCE: one mbarrier, one bulk copy per chunk
            if (threadIdx.x == 0) {
                uint32_t total = 0;
                for (int c = 0; c < net.n_chunks; ++c) total += net.chunks[c].n_floats * 4u;
                mbar_expect_tx(&full[0], total);
                for (int c = 0; c < net.n_chunks; ++c)
                    tma_bulk_g2s(stage + net.chunks[c].smem_off, net.chunks[c].src, net.chunks[c].n_floats * 4u, &full[0]);
            }
            mbar_wait(&full[0], 0u);
            // De-phase the two warps that share an SM sub-partition (warps w and w+4): every warp repeats the same
            // period (FMA-bound main loop, then the latency-bound activation epilogue); started together they
            // stay in lock-step, so the FMA pipe idles during both epilogues.  Half a layer of head start makes
            // one warp's epilogue overlap the other's main loop, and with no CTA barrier it stays that way.
            if (warp >= 4) {
                const long long t0 = clock64();
                const long long wait_cycles = net.dephase > 0 ? (long long)net.dephase : 40ll * net.layers[net.n_layers > 1 ? 1 : 0].in_pad;
                while (clock64() - t0 < wait_cycles) {}
            }
        } else if (threadIdx.x == 0) {
            for (unsigned s = 0; s < kStages; ++s) issue(s);      // prologue: fill the ring
        }
    }

    // activation row pointer of slot (n, t) row r
    __device__ __forceinline__ float* row_ptr(int n, int tt, int r) const {
        return act + (n * G::TPW + tt) * TS + r * G::S;
    }
    __device__ __forceinline__ float* slot_ptr(int slot) const { return act + slot * TS; }

    // Streamed weights: the chunk sequence 0,1,2,... (cyclic over the launch's chunk table) flows through a ring of
    // kStages shared-memory stages.  Chunk q lands in stage q % kStages by one TMA bulk copy that completes on
    // full[stage].  There is NO CTA barrier in the layer loop: a warp that has finished reading a chunk bumps the
    // stage's `done` counter, and the LAST of the kWarps warps to do so refills the stage with chunk q + kStages.
    // Warps therefore drift apart by up to a chunk, which absorbs their different trip counts (zero-skipping) and
    // lets one warp's activation epilogue overlap another warp's FMA loop.
    __device__ __forceinline__ void issue(unsigned q) {   // one thread
        const ChunkDev& c = net.chunks[q % (unsigned)net.n_chunks];
        const unsigned s = q % kStages;
        const uint32_t bytes = c.n_floats * 4u;
        mbar_expect_tx(&full[s], bytes);
        tma_bulk_g2s(stage + s * kChunkFloats, c.src, bytes, &full[s]);
    }
    // this warp is done with chunk q (all its reads of the stage have been consumed)
    __device__ __forceinline__ void release(unsigned q) {
        __syncwarp();
        if (lane == 0) {
            const unsigned s = q % kStages;
            __threadfence_block();
            const int old = atomicAdd(&done[s], 1);
            if (old == kWarps - 1) {
                done[s] = 0;
                __threadfence_block();
                fence_proxy_async();
                issue(q + kStages);
            }
        }
    }

    // Wait for the next chunk of the stream and return its smem pointer.
    __device__ __forceinline__ const float* acquire_chunk(int ch) {
        if (resident) return stage + net.chunks[ch].smem_off;
        const unsigned q = seq_consumed;
        if (q > 0) release(q - 1);
        const unsigned s = q % kStages;
        mbar_wait(&full[s], (q / kStages) & 1u);
        seq_consumed = q + 1;
        return stage + s * kChunkFloats;
    }

    // Every thread must call this before the kernel exits: no bulk copy may be in flight into a dead CTA.
    __device__ __forceinline__ void drain() {
        if (net.exec_macs != nullptr && lane == 0) atomicAdd(net.exec_macs, exec_macs * (unsigned long long)RT);
        if (resident) return;
        if (seq_consumed > 0) release(seq_consumed - 1);
        for (unsigned q = seq_consumed; q < seq_consumed + kStages; ++q) mbar_wait(&full[q % kStages], (q / kStages) & 1u);
        __syncthreads();
    }

    // activation fragment: 4 consecutive k of every row this thread carries
    __device__ __forceinline__ void load_act(float4 (&a)[ROWS], const float* p) const {
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int r = 0; r < RT; ++r)
                a[n * RT + r] = *reinterpret_cast<const float4*>(p + n * G::TPW * TS + r * G::S);
    }

    // One k-step of the outer product on packed column pairs: acc[r][p] holds columns (2p, 2p+1) of the thread's 8.
    // Issue order is weight-major and serpentine over the rows: FFMA2s that share one 64-bit weight pair are adjacent
    // (the pair sits in the operand reuse cache), and at a column change the activation scalar is the shared
    // operand.  An FFMA2 whose 5 source registers all come from the register file needs 3 cycles on the 2-cycle
    // packed pipe (two banks, B300_MICROARCH.md "RF banking"); with one operand reused it runs at full rate
    // (measured: tools/probes/ffma2_reuse_probe, profiles/r1_ffma2_reuse_probe.txt).  The err rows multiply |A|
    // (reference src/affine_layers.py:27) and follow, activation-major.
    template <int JJ>
    __device__ __forceinline__ void fma_k(f32x2 (&acc)[ROWS][4], const float4 (&a)[ROWS], const ulonglong2& w0,
                                          const ulonglong2& w1) const {
        const f32x2 wv[4] = {w0.x, w0.y, w1.x, w1.y};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int rr = 0; rr < ROWS; ++rr) {
                const int r = (c & 1) ? ROWS - 1 - rr : rr;
                if (!Tile::is_err(r % RT)) {
                    const float av = JJ == 0 ? a[r].x : JJ == 1 ? a[r].y : JJ == 2 ? a[r].z : a[r].w;
                    acc[r][c] = ffma2(pack2(av, av), wv[c], acc[r][c]);
                }
            }
        }
        if (Tile::rule != 1) {
            f32x2 wa[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) wa[c] = abs2(wv[c]);
            int e = 0;
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                if (Tile::is_err(r % RT)) {
                    const float av = JJ == 0 ? a[r].x : JJ == 1 ? a[r].y : JJ == 2 ? a[r].z : a[r].w;
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const int c = (e & 1) ? 3 - cc : cc;
                        acc[r][c] = ffma2(pack2(av, av), wa[c], acc[r][c]);
                    }
                    ++e;
                }
            }
        }
    }
    // 4 k-steps, dense addressing: (w0,w1) hold the weight row of the first k on entry and of the k AFTER the group
    // on exit; wnext points at the weight row of the group's second k.
    __device__ __forceinline__ void fma_group(f32x2 (&acc)[ROWS][4], const float4 (&a)[ROWS], ulonglong2& w0, ulonglong2& w1,
                                              const float* wnext, int wstride, int dcol) const {
        ulonglong2 n0, n1;
#define NIQ_STEP(JJ)                                                                      \
        n0 = *reinterpret_cast<const ulonglong2*>(wnext + JJ * wstride);                  \
        n1 = *reinterpret_cast<const ulonglong2*>(wnext + JJ * wstride + dcol);           \
        fma_k<JJ>(acc, a, w0, w1);                                                        \
        w0 = n0; w1 = n1;
        NIQ_STEP(0) NIQ_STEP(1) NIQ_STEP(2) NIQ_STEP(3)
#undef NIQ_STEP
    }
    // 4 k-steps, list addressing (zero-skipping): o1..o4 are the byte offsets of the NEXT four live weight rows
    __device__ __forceinline__ void fma_group_sp(f32x2 (&acc)[ROWS][4], const float4 (&a)[ROWS], ulonglong2& w0, ulonglong2& w1,
                                                 const char* wb, uint32_t o1, uint32_t o2, uint32_t o3, uint32_t o4,
                                                 int dcol_bytes) const {
        ulonglong2 n0, n1;
#define NIQ_STEP(JJ, OFF)                                                                 \
        n0 = *reinterpret_cast<const ulonglong2*>(wb + OFF);                              \
        n1 = *reinterpret_cast<const ulonglong2*>(wb + OFF + dcol_bytes);                 \
        fma_k<JJ>(acc, a, w0, w1);                                                        \
        w0 = n0; w1 = n1;
        NIQ_STEP(0, o1) NIQ_STEP(1, o2) NIQ_STEP(2, o3) NIQ_STEP(3, o4)
#undef NIQ_STEP
    }

    // bias + activation rule on the thread's NT x 8 neurons (reference src/affine_layers.py:34-97 and
    // src/affine.py:164-182 for the group rows, src/mlp.py:283-293 for the point rows), stage by stage over the 8
    // columns so that independent chains sit next to each other
    template <int ACT>
    __device__ __forceinline__ void epilogue(float (&acc)[ROWS][8], const float (&bias)[8]) const {
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            if constexpr (Tile::rule == 2) {
                // slope interval (reference src/slope_interval_layers.py:35-110, src/slope_interval.py:196-206):
                // slope bounds C -+ W, primal bounds primal -+ sum_v max(upper, -lower), derivative bounds of the
                // activation on them, interval product, re-centred.  Rows: [P, C x n_vec, W x n_vec, ...]
                constexpr int NV = Tile::n_vec;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float p = acc[n * RT][c] + bias[c];
                    float sl[NV], su[NV], prad = 0.f;
#pragma unroll
                    for (int v = 0; v < NV; ++v) {
                        sl[v] = acc[n * RT + 1 + v][c] - acc[n * RT + 1 + NV + v][c];
                        su[v] = acc[n * RT + 1 + v][c] + acc[n * RT + 1 + NV + v][c];
                        prad = prad + fmaxf(su[v], -sl[v]);
                    }
                    if (ACT != ACT_NONE) {
                        const float pl = p - prad, pu = p + prad;
                        float dfl, dfu;
                        if (ACT == ACT_RELU) { dfl = pl > 0.f ? 1.f : 0.f; dfu = pu < 0.f ? 0.f : 1.f; }
                        else if (ACT == ACT_ELU) { dfl = fminf(expf(pl), 1.f); dfu = fminf(expf(pu), 1.f); }
                        else if (ACT == ACT_TANH) {                       // ours (unpinned): 1 - tanh^2, largest nearest to 0
                            const float far = tanhf(fmaxf(fabsf(pl), fabsf(pu))), near = tanhf(fminf(fabsf(pl), fabsf(pu)));
                            dfl = 1.f - far * far;
                            dfu = (pl <= 0.f && pu >= 0.f) ? 1.f : 1.f - near * near;
                        }
                        else cos_bound(pl, pu, dfl, dfu);                 // sin: the derivative can be negative
#pragma unroll
                        for (int v = 0; v < NV; ++v) {
                            float nl = fminf(sl[v] * dfl, sl[v] * dfu), nu = fmaxf(su[v] * dfl, su[v] * dfu);
                            if (ACT == ACT_SIN) {                        // full interval product (:100-104)
                                nl = fminf(fminf(nl, su[v] * dfl), su[v] * dfu);
                                nu = fmaxf(fmaxf(sl[v] * dfl, sl[v] * dfu), nu);
                            }
                            const float nc = 0.5f * (nl + nu);
                            acc[n * RT + 1 + v][c] = nc;
                            acc[n * RT + 1 + NV + v][c] = nu - nc;
                        }
                        acc[n * RT][c] = ACT == ACT_RELU ? fmaxf(p, 0.f) : ACT == ACT_ELU ? elu_f(p) : ACT == ACT_TANH ? tanhf(p) : sinf(p);
                    } else {
                        acc[n * RT][c] = p;
                    }
                }
            }
            if (Tile::has_group) {
                constexpr int ie = Tile::n_aff + 1;      // err row index inside the tile
                float base[8], rad[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) base[c] = acc[n * RT][c] + bias[c];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float r = 0.f;
#pragma unroll
                    for (int k = 1; k <= Tile::n_aff; ++k) r += fabsf(acc[n * RT + k][c]);
                    rad[c] = r + acc[n * RT + ie][c];
                }
                if (ACT != ACT_NONE) {
                    float alpha[8], beta[8], delta[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        if (ACT == ACT_RELU) relu_lin(base[c] - rad[c], base[c] + rad[c], alpha[c], beta[c], delta[c]);
                        else if (ACT == ACT_ELU) elu_lin(base[c] - rad[c], base[c] + rad[c], alpha[c], beta[c], delta[c]);
                        else if (ACT == ACT_TANH) tanh_lin(base[c] - rad[c], base[c] + rad[c], alpha[c], beta[c], delta[c]);
                        else sin_lin(base[c] - rad[c], base[c] + rad[c], alpha[c], beta[c], delta[c]);
                    }
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        base[c] = alpha[c] * base[c] + beta[c];
#pragma unroll
                        for (int k = 1; k <= Tile::n_aff; ++k) acc[n * RT + k][c] = alpha[c] * acc[n * RT + k][c];
                        acc[n * RT + ie][c] = alpha[c] * acc[n * RT + ie][c] + delta[c];
                    }
                }
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[n * RT][c] = base[c];
            }
#pragma unroll
            for (int r = 0; r < RT; ++r) {
                if (Tile::is_pt(r)) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        float x = acc[n * RT + r][c] + bias[c];
                        if (ACT == ACT_RELU) x = fmaxf(x, 0.f);
                        else if (ACT == ACT_ELU) x = elu_pt(x);
                        else if (ACT == ACT_SIN) x = sinf(x);
                        else if (ACT == ACT_TANH) x = tanhf(x);
                        acc[n * RT + r][c] = x;
                    }
                }
            }
        }
    }

    // ---- write-back, dense: every column of every row, in place ---------------------------------------
    __device__ __forceinline__ void write_back_dense(const float (&acc)[ROWS][8], bool active, int col0, int col1) {
        if (active) {
#pragma unroll
            for (int n = 0; n < NT; ++n)
#pragma unroll
                for (int r = 0; r < RT; ++r) {
                    float* dst = row_ptr(n, t, r);
                    const float* a = acc[n * RT + r];
                    *reinterpret_cast<float4*>(dst + col0) = make_float4(a[0], a[1], a[2], a[3]);
                    *reinterpret_cast<float4*>(dst + col1) = make_float4(a[4], a[5], a[6], a[7]);
                }
        }
    }

    // ---- write-back, zero-skipping ------------------------------------------------------------------------
    // After a ReLU layer a neuron that is provably inactive on the whole box has base = aff = err = 0 (alpha = beta
    // = 0, reference src/affine_layers.py:44-49) and relu(point rows) = 0: about half of the columns for the small
    // boxes / ray segments that dominate every query.  A column that is exactly zero in EVERY row of the warp
    // contributes fma(0, w, acc) = acc to the next layer, so the warp drops it: the surviving columns are written
    // compacted IN COLUMN ORDER (same summation order as the dense loop => bit-identical sums), one segment per
    // weight chunk of the consuming layer, each padded to a multiple of 4 with zero columns, together with the list
    // of their weight-row byte offsets.  The next layer's K loop then runs over the list.
    __device__ __forceinline__ void write_back_sparse(const LayerDev& Ln, const float (&acc)[ROWS][8], bool active, int cg_l) {
        unsigned m[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            // nonzero in any row: OR of the bit patterns without their sign bits (-0 counts as zero, nan as nonzero)
            unsigned bits = 0u;
#pragma unroll
            for (int r = 0; r < ROWS; ++r) bits |= __float_as_uint(acc[r][c]);
            const bool nz = active && (bits << 1) != 0u;
            unsigned b = __ballot_sync(0xffffffffu, nz);
#pragma unroll
            for (int s = G::CG; s < 32; s <<= 1) b |= b >> s;          // OR over the tiles (t) that share this column
            m[c] = G::CG == 32 ? b : (b & ((1u << G::CG) - 1u));
        }
        const unsigned lt = (1u << cg) - 1u;
        int tot1 = 0, bef1 = 0, bef2 = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) { tot1 += __popc(m[c]); bef1 += __popc(m[c] & lt); bef2 += __popc(m[c + 4] & lt); }
        // segments: lane i < nch owns chunk i of the consuming layer
        const int nch = Ln.chunk_end - Ln.chunk_begin;
        const int kc0 = net.chunks[Ln.chunk_begin].kc;
        const int split = 4 * cg_l;                                    // first column of the threads' second block
        {
            int cb[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                int k0 = (lane + e) * kc0;
                if (k0 > 2 * split) k0 = 2 * split;
                int g = (k0 < split ? k0 : k0 - split) >> 2;           // column group holding column k0
                const unsigned mk = g >= 32 ? 0xffffffffu : ((1u << g) - 1u);
                int s = 0;
                if (k0 < split) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) s += __popc(m[c] & mk);
                } else {
                    s = tot1;
#pragma unroll
                    for (int c = 4; c < 8; ++c) s += __popc(m[c] & mk);
                }
                cb[e] = s;
            }
            const int cnt = lane < nch ? cb[1] - cb[0] : 0;
            const int n8 = (cnt + 3) & ~3;                             // padded length of the segment (multiple of 4)
            int start = n8;                                            // exclusive prefix sum over the (<= 8) chunks
#pragma unroll
            for (int off = 1; off < 2 * kMaxSegs; off <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, start, off);
                if (lane >= off) start += y;
            }
            start -= n8;
            if (lane < nch) {
                seg[4 * lane] = start; seg[4 * lane + 1] = n8; seg[4 * lane + 2] = cb[0]; seg[4 * lane + 3] = cnt;
            }
            if (lane == nch) seg[4 * lane] = start;                    // total padded length (sentinel)
        }
        __syncwarp();                      // (also: every lane has finished READING this layer's input rows)
        const int row_bytes = Ln.out_pad * 4;                          // out_pad == 1 for the dot layer
        if (active) {
#pragma unroll
            for (int blk = 0; blk < 2; ++blk) {
                const int j0 = blk == 0 ? 4 * cg : split + 4 * cg;     // my 4 columns of this block
                int ch = j0 / kc0;
                if (ch > nch - 1) ch = nch - 1;
                const int k0 = ch * kc0;
                int p = seg[4 * ch] + (blk == 0 ? bef1 : tot1 + bef2) - seg[4 * ch + 2];
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const int c = blk * 4 + c4;
                    if ((m[c] >> cg) & 1u) {
#pragma unroll
                        for (int n = 0; n < NT; ++n)
#pragma unroll
                            for (int r = 0; r < RT; ++r) row_ptr(n, t, r)[p] = acc[n * RT + r][c];
                        if (t == 0) lst[p] = (uint32_t)((j0 + c4 - k0) * row_bytes);
                        ++p;
                    }
                }
            }
        }
        // zero columns that pad every segment to a multiple of 4 (+ the list's over-read tail): lane = (chunk, pad slot)
        {
            const int i = lane >> 2, q = lane & 3;
            if (i < nch) {
                const int4 sg = *reinterpret_cast<const int4*>(seg + 4 * i);      // start, n4, cb0, cnt
                if (q < sg.y - sg.w) {
                    const int pos = sg.x + sg.w + q;
                    for (int sl = 0; sl < SLOTS; ++sl)
#pragma unroll
                        for (int r = 0; r < RT; ++r) act[sl * TS + r * G::S + pos] = 0.f;
                    lst[pos] = 0u;
                }
            }
            if (lane < 8) lst[seg[4 * nch] + lane] = 0u;
        }
    }

    // ---- one hidden layer: acc[rows][8] = act[rows][:K] @ W[:K][my 8 columns] -----------------------
    // sp_in: the input rows are compacted (written by write_back_sparse of the previous layer);
    // sp_out: compact this layer's output for Ln.
    __device__ __forceinline__ void hidden_layer(const LayerDev& L, const LayerDev& Ln, bool sp_in, bool sp_out) {
        f32x2 acc2[ROWS][4];
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc2[r][c] = 0ull;

        const int cg_l = L.out_pad >> 3;                 // active column groups of this layer
        const bool active = cg < cg_l;
        const int col0 = 4 * cg, col1 = 4 * (cg + cg_l); // the thread's two float4 column blocks
        const float* a_base = row_ptr(0, t, 0);
        const int wstride = L.out_pad;
        const int dcol = col1 - col0;

        for (int ch = L.chunk_begin; ch < L.chunk_end; ++ch) {
            const float* w = acquire_chunk(ch);
            const ChunkDev& C = net.chunks[ch];
            const float* wrow = w + col0;
            if (sp_in) {
                const int si = ch - L.chunk_begin;
                const int s0 = seg[4 * si], n8 = seg[4 * si + 1];
                exec_macs += (unsigned long long)(n8 * L.out_pad * SLOTS);
                if (active && n8 > 0) {
                    // the same software pipeline as the dense loop, with the weight row of every k taken from the list
                    const float* arow = a_base + s0;
                    const uint32_t* lp = lst + s0;
                    const char* wb = reinterpret_cast<const char*>(wrow);
                    const int dcol_bytes = dcol * 4;
                    float4 aA[ROWS], aB[ROWS];
                    uint4 oA = *reinterpret_cast<const uint4*>(lp), oB;
                    ulonglong2 w0 = *reinterpret_cast<const ulonglong2*>(wb + oA.x);
                    ulonglong2 w1 = *reinterpret_cast<const ulonglong2*>(wb + oA.x + dcol_bytes);
                    load_act(aA, arow);
                    int j = 0;
#pragma unroll 1
                    for (; j + 8 <= n8; j += 8) {
                        load_act(aB, arow + j + 4);
                        oB = *reinterpret_cast<const uint4*>(lp + j + 4);
                        fma_group_sp(acc2, aA, w0, w1, wb, oA.y, oA.z, oA.w, oB.x, dcol_bytes);
                        load_act(aA, arow + j + 8);
                        oA = *reinterpret_cast<const uint4*>(lp + j + 8);
                        fma_group_sp(acc2, aB, w0, w1, wb, oB.y, oB.z, oB.w, oA.x, dcol_bytes);
                    }
                    // odd number of 4-groups: one more, its fragments were loaded by the prologue / the last trip
                    if (j < n8) fma_group_sp(acc2, aA, w0, w1, wb, oA.y, oA.z, oA.w, oA.w, dcol_bytes);
                }
            } else {
                exec_macs += (unsigned long long)(C.kc * L.out_pad * SLOTS);
                if (active) {
                    const float* arow = a_base + C.k0;
                    if ((C.kc & 7) == 0) {
                        // Software-pipelined main loop, 8 k per trip: register double buffers for the activation
                        // fragments (two float4 sets, 4 k each) and for the weight row of the NEXT k, so every LDS
                        // is issued one stage ahead of the FFMAs that consume it (a warp covers its own latency).
                        // The loads of the last trip run past the chunk / row end into padding or the neighbouring
                        // stage: in-bounds of the CTA's shared memory, never consumed.
                        float4 aA[ROWS], aB[ROWS];
                        ulonglong2 w0, w1;
                        load_act(aA, arow);
                        w0 = *reinterpret_cast<const ulonglong2*>(wrow);
                        w1 = *reinterpret_cast<const ulonglong2*>(wrow + dcol);
#pragma unroll 1
                        for (int j = 0; j < C.kc; j += 8) {
                            load_act(aB, arow + j + 4);
                            fma_group(acc2, aA, w0, w1, wrow + (j + 1) * wstride, wstride, dcol);
                            load_act(aA, arow + j + 8);
                            fma_group(acc2, aB, w0, w1, wrow + (j + 5) * wstride, wstride, dcol);
                        }
                    } else {
                        // short / odd K (the 3-D input layer, in_pad = 4): plain loop
                        for (int j = 0; j < C.kc; j += 4) {
                            float4 a4[ROWS];
                            load_act(a4, arow + j);
                            ulonglong2 w0 = *reinterpret_cast<const ulonglong2*>(wrow + j * wstride);
                            ulonglong2 w1 = *reinterpret_cast<const ulonglong2*>(wrow + j * wstride + dcol);
                            fma_group(acc2, a4, w0, w1, wrow + (j + 1) * wstride, wstride, dcol);
                        }
                    }
                }
            }
        }

        float acc[ROWS][8];
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) unpack2(acc2[r][c], acc[r][2 * c], acc[r][2 * c + 1]);

        // ---- epilogue: bias, activation rule, in-place write-back --------------------------------------
        if (active) {
            float bias[8];
            {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(L.bias + col0));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(L.bias + col1));
                bias[0] = b0.x; bias[1] = b0.y; bias[2] = b0.z; bias[3] = b0.w;
                bias[4] = b1.x; bias[5] = b1.y; bias[6] = b1.z; bias[7] = b1.w;
            }
            // one straight-line instance per activation kind: the 16 neurons of the thread are independent, and
            // without a runtime branch inside the unrolled loops the scheduler interleaves their chains
            if (L.act == ACT_RELU) epilogue<ACT_RELU>(acc, bias);
            else if (L.act == ACT_ELU) epilogue<ACT_ELU>(acc, bias);
            else if (L.act == ACT_SIN) epilogue<ACT_SIN>(acc, bias);
            else if (L.act == ACT_TANH) epilogue<ACT_TANH>(acc, bias);
            else epilogue<ACT_NONE>(acc, bias);
        }
        if (sp_out) {
            write_back_sparse(Ln, acc, active, cg_l);   // orders its reads / writes with its own __syncwarp
        } else {
            __syncwarp();            // every lane has finished READING this layer's input rows
            write_back_dense(acc, active, col0, col1);
        }
        __syncwarp();
    }

    // ---- final layer (out_dim == 1): out[r] = act[r][:K] . w + b ; also |.|-sum for point rows -----------
    // Results land in out[ROWS] (identical in all lanes of the tile group); pscale[ROWS] = sum|h_j w_j| + |b|
    // for point rows and for the base row (the magnitude the output was summed from: near-tie yardstick).
    __device__ __forceinline__ void dot_layer(const LayerDev& L, bool sp_in, float out[ROWS], float pscale[ROWS]) {
#pragma unroll
        for (int r = 0; r < ROWS; ++r) { out[r] = 0.f; pscale[r] = 0.f; }
        const float* a_base = row_ptr(0, t, 0);
        for (int ch = L.chunk_begin; ch < L.chunk_end; ++ch) {
            const float* w = acquire_chunk(ch);
            const ChunkDev& C = net.chunks[ch];
            const int si = ch - L.chunk_begin;
            const int a0 = sp_in ? seg[4 * si] : C.k0;
            const int kn = sp_in ? seg[4 * si + 3] : C.kc;
            exec_macs += (unsigned long long)(kn * SLOTS);
            for (int j = cg; j < kn; j += G::CG) {
                const float wj = sp_in ? w[lst[a0 + j] >> 2] : w[j];
                const float wja = fabsf(wj);
#pragma unroll
                for (int n = 0; n < NT; ++n)
#pragma unroll
                    for (int r = 0; r < RT; ++r) {
                        const float a = a_base[n * G::TPW * TS + r * G::S + a0 + j];
                        const int i = n * RT + r;
                        if (Tile::is_err(r)) out[i] = fmaf(a, wja, out[i]);
                        else out[i] = fmaf(a, wj, out[i]);
                        if (Tile::want_scale(r)) pscale[i] = fmaf(fabsf(a), wja, pscale[i]);
                    }
            }
        }
#pragma unroll
        for (int off = 1; off < G::CG; off <<= 1) {
#pragma unroll
            for (int i = 0; i < ROWS; ++i) {
                out[i] += __shfl_xor_sync(0xffffffffu, out[i], off);
                if (Tile::want_scale(i % RT)) pscale[i] += __shfl_xor_sync(0xffffffffu, pscale[i], off);
            }
        }
        const float b = __ldg(L.bias);
#pragma unroll
        for (int i = 0; i < ROWS; ++i) {
            if (Tile::has_bias(i % RT)) { out[i] += b; pscale[i] += fabsf(b); }
        }
        __syncwarp();
    }

    // A warp that sits a pass out (niq_tree.cuh: small levels use half of the warps so that each runs alone on its scheduler)
    // still takes part in the streamed weight ring: it acquires every chunk of layers [l0, l1) in order and lets go of it at once.
    __device__ __forceinline__ void skip_net(int l0, int l1) {
        if (resident) return;
        for (int l = l0; l < l1; ++l)
            for (int ch = net.layers[l].chunk_begin; ch < net.layers[l].chunk_end; ++ch) acquire_chunk(ch);
    }

    // Run layers [l0, l1) of the stream; the last one must be a dot layer.
    __device__ __forceinline__ void run_net(int l0, int l1, float out[ROWS], float pscale[ROWS]) {
        bool sp = false;
        for (int l = l0; l < l1 - 1; ++l) {
            const LayerDev& L = net.layers[l];
            const bool sp_out = kSparse && net.sparse != 0 && L.act == ACT_RELU;
            hidden_layer(L, net.layers[l + 1], sp, sp_out);
            sp = sp_out;
        }
        dot_layer(net.layers[l1 - 1], sp, out, pscale);
    }
};

}  // namespace niq
