// niq_cp.cuh -- closest_point with the reference's default window (batch_process_size <= 2048) as ONE persistent kernel
// (reference src/kd_tree.py:659-802).  The reference pops the top window of a global LIFO stack per host iteration: one jitted
// pass (classify + 7 samples per node, scatter-min, compaction, push) and a blocking read of the stack top -- 7,290 iterations
// for 256 queries on birdcage_occ.  Here a cooperative kernel runs all of them:
//   phase 1   the window's boxes and its 7 sample points per box go through the engine as two kinds of pass (box tiles and
//             point tiles) dealt over the CTAs of the grid, half-CTA passes when the window fits the grid that way
//   -------   grid barrier
//   phase 2   CTA 0 alone replays the round of the reference on the window -- snapshot of min_dist before the scatter-min,
//             atomicMin on the float bits, "last index wins" location write, ordered compaction, children pushed interleaved
//             at pop + 2*rank -- exactly the logic of k_cp_round_small, with block barriers between its steps
//   -------   grid barrier
// The stack top, the round number and the statistics stay on the device; the host reads them once.  The stack is sized by the
// host; a round that would not fit stops the kernel at a round boundary and the host grows the stack and relaunches.
#pragma once
#include "niq_kernels.cuh"
#include "niq_tree.cuh"

namespace niq {

constexpr int kCpEpt = 8;                    // window entries per thread of CTA 0: 256 x 8 = 2048 = the largest window served

struct CpCtl {
    long long top, max_top, round, status;   // status: 0 finished, 1 the stack needs to grow
    long long stats[4];                      // [0] rounds that did work, [1] node visits, [2] near-tie boxes
    unsigned int bar_count, bar_gen;
};

struct CpArgs {
    float* stack_lo; float* stack_hi; long long* stack_qid; long long cap;
    long long window;                        // B
    const float* query; float* min_dist; float* min_loc; unsigned long long* winner; long long n_query;
    int* label; unsigned char* tie; float* vals;      // window-indexed scratch
    float eps_w;
    int interval;
    CpCtl* ctl;
};

__device__ __forceinline__ void cp_barrier(CpCtl* ctl) {
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile unsigned int* gen = &ctl->bar_gen;
        const unsigned int g = *gen;
        __threadfence();
        if (atomicAdd(&ctl->bar_count, 1u) == gridDim.x - 1) {
            ctl->bar_count = 0u;
            __threadfence();
            atomicAdd(&ctl->bar_gen, 1u);
        } else {
            while (*gen == g) __nanosleep(32);
        }
        __threadfence();
    }
    __syncthreads();
}

template <int WMAX>
__global__ void __launch_bounds__(kThreads, 1) k_cp_persistent(const __grid_constant__ NetDev net, const CpArgs a) {
    using EB = Engine<WMAX, TileBox3>;
    using EP = Engine<WMAX, TilePts>;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ int s_warp[kWarps];
    __shared__ int s_total;
    EB engB(net, smem, true);
    EP engP(net, smem, false);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float INF = __int_as_float(0x7f800000);
    constexpr int PTS_WARP = EP::WARP_ROWS;

    for (;;) {
        const long long top = *reinterpret_cast<volatile long long*>(&a.ctl->top);
        const unsigned long long round = (unsigned long long)*reinterpret_cast<volatile long long*>(&a.ctl->round);
        if (top <= 0) break;
        if (top + a.window > a.cap) {          // the children of this round might not fit: the host grows the stack
            if (blockIdx.x == 0 && tid == 0) a.ctl->status = 1;
            break;
        }
        const long long pop = top - a.window > 0 ? top - a.window : 0;
        const long long nv = top - pop;                                 // valid window entries (reference :685)
        const float* slo = a.stack_lo + 3 * pop;
        const float* shi = a.stack_hi + 3 * pop;

        // ---------------- phase 1: box passes and point passes over the CTAs ----------------
        const long long n_pts = 7 * nv;
        long long nB = (nv + (kWarps / 2) * EB::SLOTS - 1) / ((kWarps / 2) * EB::SLOTS);       // half-CTA passes
        long long nP = (n_pts + (kWarps / 2) * PTS_WARP - 1) / ((kWarps / 2) * PTS_WARP);
        bool halfB = true, halfP = true;
        if (nB + nP > gridDim.x) { halfP = false; nP = (n_pts + kWarps * PTS_WARP - 1) / (kWarps * PTS_WARP); }
        if (nB + nP > gridDim.x) { halfB = false; nB = (nv + EB::CTA_TILES - 1) / EB::CTA_TILES; }
        for (long long pass = blockIdx.x; pass < nB + nP; pass += gridDim.x) {
            if (pass < nB) {
                const int wu = halfB ? kWarps / 2 : kWarps;
                if (warp >= wu) continue;
                const long long warp_box0 = pass * (long long)(wu * EB::SLOTS) + (long long)warp * EB::SLOTS;
                if (lane < EB::SLOTS) {
                    const long long i = warp_box0 + lane;
                    float4 rows[5];
#pragma unroll
                    for (int r = 0; r < 5; ++r) rows[r] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (i < nv) {
                        BoxSource src{};
                        src.kind = 1; src.v = 3; src.a = slo; src.b = shi; src.interval = a.interval;
                        load_box_rows(src, i, rows);
                    }
                    float* dst = engB.slot_ptr(lane);
#pragma unroll
                    for (int r = 0; r < 5; ++r) *reinterpret_cast<float4*>(dst + r * EB::G::S) = rows[r];
                }
                __syncwarp();
                float out[EB::ROWS], ps[EB::ROWS];
                engB.run_net(0, net.n_layers, out, ps);
                if (engB.cg == 0) {
#pragma unroll
                    for (int nn = 0; nn < EB::NT; ++nn) {
                        const long long i = warp_box0 + nn * EB::G::TPW + engB.t;
                        if (i < nv) {
                            const float base = out[nn * 5];
                            const float rad = ((fabsf(out[nn * 5 + 1]) + fabsf(out[nn * 5 + 2])) + fabsf(out[nn * 5 + 3])) + out[nn * 5 + 4];
                            const float lo = base - rad, up = base + rad;
                            a.label[i] = label_of(lo, up, 0.f);
                            a.tie[i] = bound_near_tie(lo, up, 0.f, ps[nn * 5], net.tie_rel) ? 1 : 0;
                        }
                    }
                }
                __syncwarp();
            } else {
                const int wu = halfP ? kWarps / 2 : kWarps;
                if (warp >= wu) continue;
                const long long p0 = (pass - nB) * (long long)(wu * PTS_WARP) + (long long)warp * PTS_WARP;
                PointSource src{};
                src.kind = 2; src.a = slo; src.b = shi; src.sample_scale = -1.f;      // centre +- the node's full extent (:702-704)
                for (int r = lane; r < PTS_WARP; r += 32) {
                    const long long i = p0 + r;
                    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (i < n_pts) x = load_point(src, i);
                    *reinterpret_cast<float4*>(engP.slot_ptr(r >> 3) + (r & 7) * EP::G::S) = x;
                }
                __syncwarp();
                float out[EP::ROWS], ps[EP::ROWS];
                engP.run_net(0, net.n_layers, out, ps);
                if (engP.cg == 0) {
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        const long long i = p0 + engP.t * 8 + r;
                        if (i < n_pts) a.vals[i] = out[r];
                    }
                }
                __syncwarp();
            }
        }
        cp_barrier(a.ctl);

        // ---------------- phase 2: the round of the reference on the window, CTA 0 alone ----------------
        if (blockIdx.x == 0) {
            float l[kCpEpt][3], h[kCpEpt][3], cen[kCpEpt][3], dist[kCpEpt];
            long long qid[kCpEpt];
            bool valid[kCpEpt], need[kCpEpt];
            int n_valid = 0, n_tie = 0;
#pragma unroll
            for (int e = 0; e < kCpEpt; ++e) {
                const long long i = (long long)kCpEpt * tid + e;
                valid[e] = i < nv;
                need[e] = false; dist[e] = INF; qid[e] = 0;
#pragma unroll
                for (int d = 0; d < 3; ++d) { l[e][d] = h[e][d] = cen[e][d] = 0.f; }
                if (valid[e]) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) { l[e][d] = slo[3 * i + d]; h[e][d] = shi[3 * i + d]; }
                    long long q = a.stack_qid[pop + i];
                    if (q < 0 || q >= a.n_query) q = 0;
                    qid[e] = q;
                    const float* qp = a.query + 3 * q;
                    const float ex = h[e][0] - l[e][0], ey = h[e][1] - l[e][1], ez = h[e][2] - l[e][2];
                    const float width = fmaxf(fmaxf(ex, ey), ez);
#pragma unroll
                    for (int d = 0; d < 3; ++d) cen[e][d] = 0.5f * (l[e][d] + h[e][d]);
                    const float off = sqrtf((ex * ex + ey * ey) + ez * ez);
                    const float qx = qp[0] - cen[e][0], qy = qp[1] - cen[e][1], qz = qp[2] - cen[e][2];
                    const float dc = sqrtf((qx * qx + qy * qy) + qz * qz);
                    const bool small = width < a.eps_w;
                    const int lab = a.label[i];
                    const bool outside = lab == SIGN_NEGATIVE || lab == SIGN_POSITIVE;
                    const bool spans = !all_same_sign7(a.vals + 7 * i);
                    const float snap = a.min_dist[q];                 // snapshot before this round's scatter-min (:684)
                    dist[e] = spans ? dc + off : INF;
                    need[e] = !outside && !small && dc < snap;
                    n_valid += 1; n_tie += a.tie[i] ? 1 : 0;
                }
            }
            __syncthreads();                                          // every snapshot is taken before any scatter-min
#pragma unroll
            for (int e = 0; e < kCpEpt; ++e)
                if (valid[e]) atomicMin(reinterpret_cast<int*>(a.min_dist + qid[e]), __float_as_int(dist[e]));   // dist >= 0
            __threadfence();
            __syncthreads();
            const unsigned long long tag0 = (round << 32);
#pragma unroll
            for (int e = 0; e < kCpEpt; ++e)
                if (valid[e] && dist[e] == *reinterpret_cast<volatile float*>(a.min_dist + qid[e]))
                    atomicMax(&a.winner[qid[e]], tag0 | (unsigned long long)((long long)kCpEpt * tid + e + 1));
            __threadfence();
            __syncthreads();
#pragma unroll
            for (int e = 0; e < kCpEpt; ++e)
                if (valid[e] && dist[e] == *reinterpret_cast<volatile float*>(a.min_dist + qid[e]) &&
                    *reinterpret_cast<volatile unsigned long long*>(&a.winner[qid[e]]) == (tag0 | (unsigned long long)((long long)kCpEpt * tid + e + 1)))
                    for (int d = 0; d < 3; ++d) a.min_loc[3 * qid[e] + d] = cen[e][d];
            // exclusive scan of the survivors, in window order
            int mine = 0;
#pragma unroll
            for (int e = 0; e < kCpEpt; ++e) mine += need[e] ? 1 : 0;
            int x = mine;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, off); if (lane >= off) x += y; }
            if (lane == 31) s_warp[warp] = x;
            __syncthreads();
            if (tid == 0) {
                int acc = 0;
                for (int w = 0; w < kWarps; ++w) { const int v = s_warp[w]; s_warp[w] = acc; acc += v; }
                s_total = acc;
            }
            __syncthreads();
            int rank = s_warp[warp] + x - mine;
            // children of the survivors go back on the stack at pop + 2*rank, interleaved [A, B] (reference :732-754)
#pragma unroll
            for (int e = 0; e < kCpEpt; ++e) {
                if (need[e]) {
                    const long long oa = pop + 2ll * rank, ob = oa + 1;
                    const int sd = argmax3_first(h[e][0] - l[e][0], h[e][1] - l[e][1], h[e][2] - l[e][2]);
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        const float mid = 0.5f * (l[e][d] + h[e][d]);
                        a.stack_lo[3 * oa + d] = l[e][d];
                        a.stack_hi[3 * oa + d] = d == sd ? mid : h[e][d];
                        a.stack_lo[3 * ob + d] = d == sd ? mid : l[e][d];
                        a.stack_hi[3 * ob + d] = h[e][d];
                    }
                    a.stack_qid[oa] = qid[e];
                    a.stack_qid[ob] = qid[e];
                    rank += 1;
                }
            }
            {
                int nvv = n_valid, nt = n_tie;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) { nvv += __shfl_xor_sync(0xffffffffu, nvv, off); nt += __shfl_xor_sync(0xffffffffu, nt, off); }
                if (lane == 0) {
                    if (nvv) atomicAdd((unsigned long long*)&a.ctl->stats[1], (unsigned long long)nvv);
                    if (nt) atomicAdd((unsigned long long*)&a.ctl->stats[2], (unsigned long long)nt);
                }
            }
            if (tid == 0) {
                const long long nt2 = pop + 2ll * s_total;
                a.ctl->top = nt2;
                if (nt2 > a.ctl->max_top) a.ctl->max_top = nt2;
                a.ctl->stats[0] += 1;                                 // rounds that did work
                a.ctl->round = (long long)round + 1;                  // round number (winner tags)
                __threadfence();
            }
        }
        cp_barrier(a.ctl);
    }
    engB.drain();
}

}  // namespace niq
