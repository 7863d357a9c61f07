// niq_isect.cuh -- find_any_intersection as ONE persistent kernel for the growing-form modes (affine_truncate / affine_all /
// affine_append), for one query or a whole batch of queries (reference src/kd_tree.py:402-655).
//
// The reference runs a host loop: per round one jitted pass over the frontier (two box classifications and 14 point
// evaluations per node, the found / needs-subdivision logic, a compaction) and a blocking read of the frontier size.  Here
// the frontier of ALL queries lives in one device array (node -> query id; the nodes of a query stay contiguous and in the
// reference's order because the compaction is an ordered scan), and a cooperative kernel runs the rounds:
//   phase 1   work item = (node, shape): one CTA propagates the box through that shape's net (grow_forward) and evaluates
//             the 7 sample points (the point rows of grow_forward)
//   phase 2a  per node: the verdict logic (src/kd_tree.py:449-518); a query's first hit of the round is the lowest node index
//             (atomicMax of a round-tagged key)
//   phase 2b  nodes of queries that were hit are dropped; survivors are counted per tile of 2048 nodes
//   phase 2c  ordered scan + split, children interleaved [A0, B0, A1, B1, ...] (src/kd_tree.py:543-562)
// separated by grid barriers -- or by __syncthreads inside CTA 0 alone when the frontier fits one tile (a single query never
// exceeds a few hundred nodes), which leaves two grid barriers per round.
#pragma once
#include "niq_grow.cuh"
#include "niq_tree.cuh"

namespace niq {

struct IsectCtl {
    long long n_cur, which, round, status, need;     // status: 0 running / finished, 1 frontier buffers too small, 2 round limit
    unsigned int bar_count, bar_gen;
};

struct IsectArgs {
    float* lo[2]; float* hi[2]; int* qid[2];          // frontier double buffer
    long long cap;
    int* lab;                                         // [2][cap]  label per shape
    float* vals;                                      // [2][cap][7] sample values per shape
    unsigned char* tie;                               // [2][cap]
    int* needs;                                       // [cap]  bit 0 = needs subdivision, bit 1 = found
    float* loc;                                       // [cap][3]
    int* tile_cnt;                                    // [2 parities][n_tiles_max]
    long long n_tiles_max;
    long long n_queries;
    unsigned long long* first;                        // [n_queries]  ((round + 1) << 32) | (0xffffffff - node index)
    int* q_found; float* q_loc; long long* q_stats;   // per query: verdict, location, [nodes, rounds, near-tie]
    const float* xf[2];                               // per query layer-0 override of shape s: [n_queries][40] = A0 (4 x 8) + b0 (8), or null
    float eps_w;
    GrowCfg g[2];
    int cg_lanes[2];
    int W;                                            // row width of the point buffers (max of the two nets)
    long long state_floats;                           // size of the larger propagation state (the point buffers follow it)
    int max_rounds;
    IsectCtl* ctl;
};

__device__ __forceinline__ void isect_barrier(IsectCtl* ctl) {
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

__global__ void __launch_bounds__(256) k_isect_persistent(const __grid_constant__ NetDev netA, const __grid_constant__ NetDev netB,
                                                          const IsectArgs a) {
    extern __shared__ __align__(16) float sm[];
    __shared__ long long s_red[8];
    __shared__ int s_warp[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* hA = sm + a.state_floats;
    float* hB = hA + 8 * a.W;
    const long long T = a.n_tiles_max;

    for (;;) {
        // the round's state: published by CTA 0 before the previous round's last barrier
        const long long N = *reinterpret_cast<volatile long long*>(&a.ctl->n_cur);
        const int which = (int)*reinterpret_cast<volatile long long*>(&a.ctl->which);
        const long long round = *reinterpret_cast<volatile long long*>(&a.ctl->round);
        if (N == 0 || *reinterpret_cast<volatile long long*>(&a.ctl->status) != 0) break;
        const float* cur_lo = a.lo[which];
        const float* cur_hi = a.hi[which];
        const int* cur_q = a.qid[which];
        int* cnt = a.tile_cnt + (size_t)(round & 1) * T;
        int* cnt_next = a.tile_cnt + (size_t)((round & 1) ^ 1) * T;

        // ---------------- phase 1: (node, shape) items ----------------
        for (long long item = blockIdx.x; item < 2 * N; item += gridDim.x) {
            const long long node = item >> 1;
            const int s = (int)(item & 1);
            const NetDev& net = s ? netB : netA;
            const GrowCfg& g = a.g[s];
            const int q = cur_q[node];
            const float* A0 = a.xf[s] ? a.xf[s] + (size_t)q * 40 : nullptr;
            const float* b0 = A0 ? A0 + 32 : nullptr;
            GrowState st;
            grow_carve(sm, g, st);
            const float l0 = cur_lo[3 * node], l1 = cur_lo[3 * node + 1], l2 = cur_lo[3 * node + 2];
            const float h0 = cur_hi[3 * node], h1 = cur_hi[3 * node + 1], h2 = cur_hi[3 * node + 2];
            // reference src/implicit_function.py:34-36: center = 0.5*(lo+hi); vec = hi - center; diag(vec)
            const float cx = 0.5f * (l0 + h0), cy = 0.5f * (l1 + h1), cz = 0.5f * (l2 + h2);
            __syncthreads();
            if (tid == 0) {
                const int W = g.W;
                st.base[0] = cx; st.base[1] = cy; st.base[2] = cz; st.base[3] = 0.f;
                st.aff[0] = h0 - cx; st.aff[1] = 0.f; st.aff[2] = 0.f; st.aff[3] = 0.f;
                st.aff[W] = 0.f; st.aff[W + 1] = h1 - cy; st.aff[W + 2] = 0.f; st.aff[W + 3] = 0.f;
                st.aff[2 * W] = 0.f; st.aff[2 * W + 1] = 0.f; st.aff[2 * W + 2] = h2 - cz; st.aff[2 * W + 3] = 0.f;
                st.err[0] = st.err[1] = st.err[2] = st.err[3] = 0.f;
            }
            if (tid >= 32 && tid < 40) {
                // the 7 sample points centre +- eps_w * e_i (src/kd_tree.py:461-464), row 7 repeats the centre
                const int k = tid - 32;
                float c[3] = {cx, cy, cz};
                if (k >= 1 && k <= 3) c[k - 1] = c[k - 1] + a.eps_w;
                if (k >= 4 && k <= 6) c[k - 4] = c[k - 4] + a.eps_w * -1.f;
                float* d = hA + k * g.W;                  // row stride = this net's W (grow_forward indexes hA + warp * W)
                d[0] = c[0]; d[1] = c[1]; d[2] = c[2]; d[3] = 0.f;
            }
            __syncthreads();
            float lo_b, up_b, sc, f, fs;
            grow_forward(net, 0, net.n_layers, A0, b0, g, st, 3, lo_b, up_b, sc, hA, hB, a.cg_lanes[s], &f, &fs);
            if (lane == 0 && warp < 7) a.vals[((size_t)s * a.cap + node) * 7 + warp] = f;
            if (tid == 0) {
                a.lab[(size_t)s * a.cap + node] = label_of(lo_b, up_b, 0.f);
                a.tie[(size_t)s * a.cap + node] = bound_near_tie(lo_b, up_b, 0.f, sc, net.tie_rel) ? 1 : 0;
            }
        }
        isect_barrier(a.ctl);

        // ---------------- phase 2: verdicts, ordered compaction, split ----------------
        const bool single = N <= kTreeTile;                 // one tile: CTA 0 alone, block barriers instead of grid barriers
        const bool part = !single || blockIdx.x == 0;
        const long long cta = single ? 0 : blockIdx.x, n_cta = single ? 1 : gridDim.x;
        const unsigned long long tag = (unsigned long long)(round + 1) << 32;
        if (part) {
            // 2a: per-node verdict (src/kd_tree.py:449-518; the logic of k_isect_logic)
            for (long long i = cta * 256 + tid; i < N; i += n_cta * 256) {
                const float* l = cur_lo + 3 * i;
                const float* h = cur_hi + 3 * i;
                const float* vA = a.vals + (size_t)i * 7;
                const float* vB = a.vals + ((size_t)a.cap + i) * 7;
                const int q = cur_q[i];
                const float width = fmaxf(fmaxf(h[0] - l[0], h[1] - l[1]), h[2] - l[2]);
                const bool small = width < a.eps_w;
                const bool nearA = small && !all_same_sign7(vA);
                const bool nearB = small && !all_same_sign7(vB);
                int iA = 0, iB = 0, iT = 0;
                bool anyA = false, anyB = false, anyT = false;
                for (int k = 6; k >= 0; --k) {
                    if (vA[k] < 0.f) { iA = k; anyA = true; }
                    if (vB[k] < 0.f) { iB = k; anyB = true; }
                    if (vA[k] < 0.f && vB[k] < 0.f) { iT = k; anyT = true; }
                }
                bool found = false;
                float loc[3] = {-777.f, -777.f, -777.f};
                if (small && anyA && anyB) {
                    float pa[3], pb[3];
                    sample_point7(l, h, a.eps_w, iA, pa);
                    sample_point7(l, h, a.eps_w, iB, pb);
                    for (int d = 0; d < 3; ++d) loc[d] = 0.5f * (pa[d] + pb[d]);
                    found = true;
                }
                if (anyT) {
                    sample_point7(l, h, a.eps_w, iT, loc);
                    found = true;
                }
                const int labA = a.lab[i], labB = a.lab[a.cap + i];
                const bool insideA = labA == SIGN_NEGATIVE || (labA == SIGN_UNKNOWN && !nearA);
                const bool insideB = labB == SIGN_NEGATIVE || (labB == SIGN_UNKNOWN && !nearB);
                a.needs[i] = ((insideA && insideB) ? 1 : 0) | (found ? 2 : 0);
                if (found) {
                    for (int d = 0; d < 3; ++d) a.loc[3 * i + d] = loc[d];
                    atomicMax(a.first + q, tag | (unsigned long long)(0xffffffffu - (unsigned int)i));
                }
                atomicAdd(reinterpret_cast<unsigned long long*>(a.q_stats + 3 * q), 1ull);                      // nodes processed
                if (i == 0 || cur_q[i - 1] != q) a.q_stats[3 * q + 1] += 1;                                      // rounds of this query
                const int nt = (int)a.tie[i] + (int)a.tie[a.cap + i];
                if (nt) atomicAdd(reinterpret_cast<unsigned long long*>(a.q_stats + 3 * q + 2), (unsigned long long)nt);
            }
        }
        if (single) __syncthreads(); else isect_barrier(a.ctl);
        if (part) {
            // 2b: drop the nodes of queries that were hit this round (the reference returns at the first hit); count survivors
            for (long long i0 = cta * 256; i0 < N; i0 += n_cta * 256) {
                const long long i = i0 + tid;
                bool keep = false;
                if (i < N) {
                    const int q = cur_q[i];
                    const unsigned long long fq = *reinterpret_cast<volatile unsigned long long*>(a.first + q);
                    const bool hit = (fq >> 32) == (unsigned long long)(round + 1);
                    const int nd = a.needs[i];
                    keep = (nd & 1) && !hit;
                    a.needs[i] = keep ? 1 : 0;
                    if (hit && (nd & 2) && (unsigned int)(fq & 0xffffffffu) == 0xffffffffu - (unsigned int)i) {
                        a.q_found[q] = 1;
                        for (int d = 0; d < 3; ++d) a.q_loc[3 * q + d] = a.loc[3 * i + d];
                    }
                }
                const unsigned b = __ballot_sync(0xffffffffu, keep);
                if (lane == 0 && b) atomicAdd(cnt + (i0 + warp * 32) / kTreeTile, __popc(b));
            }
        }
        if (single) __syncthreads(); else isect_barrier(a.ctl);
        long long status = 0, need = 0, n_next = 0;
        if (part) {
            // 2c: ordered scan over the tiles + split, children interleaved
            const long long n_tiles = (N + kTreeTile - 1) / kTreeTile;
            long long sum = 0;
            for (long long t = tid; t < n_tiles; t += 256) sum += cnt[t];
            const long long total = block_sum_ll(sum, s_red);
            n_next = 2 * total;
            if (n_next > a.cap) { status = 1; need = n_next; }
            else if (round + 1 >= a.max_rounds && n_next > 0) { status = 2; }
            if (status == 0) {
                const long long per = (n_tiles + n_cta - 1) / n_cta;
                const long long t0 = cta * per, t1 = t0 + per < n_tiles ? t0 + per : n_tiles;
                long long pre = 0;
                if (t0 < t1) {
                    long long s0 = 0;
                    for (long long t = tid; t < t0; t += 256) s0 += cnt[t];
                    pre = block_sum_ll(s0, s_red);
                }
                float* out_lo = a.lo[which ^ 1];
                float* out_hi = a.hi[which ^ 1];
                int* out_q = a.qid[which ^ 1];
                for (long long tile = t0; tile < t1; ++tile) {
                    const long long base_i = tile * kTreeTile;
                    int f[8], fsum = 0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const long long i = base_i + tid * 8 + k;
                        f[k] = (i < N && a.needs[i]) ? 1 : 0;
                        fsum += f[k];
                    }
                    int x = fsum;
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, off); if (lane >= off) x += y; }
                    __syncthreads();
                    if (lane == 31) s_warp[warp] = x;
                    __syncthreads();
                    int wbase = 0;
#pragma unroll
                    for (int w = 0; w < 8; ++w) if (w < warp) wbase += s_warp[w];
                    long long run = pre + wbase + x - fsum;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        if (f[k]) {
                            const long long i = base_i + tid * 8 + k;
                            const float l[3] = {cur_lo[3 * i], cur_lo[3 * i + 1], cur_lo[3 * i + 2]};
                            const float h[3] = {cur_hi[3 * i], cur_hi[3 * i + 1], cur_hi[3 * i + 2]};
                            const long long oa = 2 * run, ob = oa + 1;
                            const int sd = argmax3_first(h[0] - l[0], h[1] - l[1], h[2] - l[2]);
#pragma unroll
                            for (int d = 0; d < 3; ++d) {
                                const float mid = 0.5f * (l[d] + h[d]);
                                out_lo[3 * oa + d] = l[d];
                                out_hi[3 * oa + d] = d == sd ? mid : h[d];
                                out_lo[3 * ob + d] = d == sd ? mid : l[d];
                                out_hi[3 * ob + d] = h[d];
                            }
                            out_q[oa] = cur_q[i]; out_q[ob] = cur_q[i];
                            run += 1;
                        }
                    }
                    pre += cnt[tile];
                }
            }
            // the other parity's counters are free (last used two phases ago): zero them for the next round
            for (long long t = cta * 256 + tid; t < T; t += n_cta * 256) cnt_next[t] = 0;
            if (status != 0)
                for (long long t = cta * 256 + tid; t < T; t += n_cta * 256) cnt[t] = 0;
            if (blockIdx.x == 0 && tid == 0) {
                IsectCtl* c = a.ctl;
                c->status = status; c->need = need;
                if (status == 0) { c->n_cur = n_next; c->which = which ^ 1; c->round = round + 1; }
                __threadfence();
            }
        }
        isect_barrier(a.ctl);
    }
}

}  // namespace niq
