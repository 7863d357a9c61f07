// niq_tree.cuh -- the level-set kd-tree as ONE persistent kernel (reference src/kd_tree.py:19-218).
//
// The reference builds the tree level by level from the host: per level a jitted classify of <= 2048-node batches, a
// compaction, and a blocking read of the node count (src/kd_tree.py:137-198).  Here the whole build is a single cooperative
// launch with one CTA per SM:
//   phase 1  every CTA classifies its share of the frontier with the warp-tiled engine (niq_engine.cuh) and writes one label
//            per node; the warps add their UNKNOWN / NEGATIVE / POSITIVE counts to per-tile counters (tile = 2048 nodes)
//   -------  grid barrier
//   phase 2  every CTA takes a contiguous range of tiles, turns the tile counters into prefixes (a block reduction), scans the
//            flags of each tile in node order and writes the children straight into the other frontier buffer -- in the
//            reference's order: per batch of `batch_process_size` nodes [A-children..., B-children...] (src/kd_tree.py:61-96)
//            -- plus the ordered interior / exterior appends (:46-59)
//   -------  grid barrier
// The node counts never leave the device: every CTA derives the next level's size from the tile counters, so there is no
// control block to read and no third barrier.  The host reads the result sizes once, after the kernel.
// Capacity: the buffers are sized by the host; a level that would not fit stops the kernel with its frontier intact, the
// host grows the buffers and relaunches from that level.
#pragma once
#include <type_traits>

#include "niq_kernels.cuh"

namespace niq {

constexpr int kTreeTile = 2048;          // nodes per scan tile (= the reference's default batch_process_size)
constexpr int kTreeMaxLevels = 1024;

struct TreeCtl {                         // device control block; written by CTA 0, read by the host after the kernel
    long long n_cur;                     // nodes of the current frontier
    long long level;                     // next level to process
    long long which;                     // frontier buffer that holds them
    long long bucket;                    // padded array size the reference would hold (src/kd_tree.py:140-148)
    long long n_fin[2];                  // interior / exterior nodes so far
    long long status;                    // 0 finished, 1 frontier buffers too small, 2 interior list, 3 exterior list
    long long need;                      // status != 0: nodes the too-small buffer must hold
    long long n_evals, max_frontier;
    unsigned long long n_tie;            // near-tie boxes of the finished levels
    unsigned long long n_tie_level;      // ... of the level in flight (folded in by CTA 0 once the level is known to fit)
    unsigned int bar_count, bar_gen;     // grid barrier
};

struct TreeArgs {
    float* lo[2]; float* hi[2];          // frontier double buffer
    long long cap;                       // nodes each frontier buffer holds
    float* fin_lo[2]; float* fin_hi[2];  // interior (NEGATIVE) / exterior (POSITIVE) lists
    long long fin_cap[2];
    int* label;                          // cap entries
    int* tile_cnt;                       // [2 parities][3 kinds][n_tiles_max]
    long long n_tiles_max;
    long long* levels;                   // 4 per level: nodes entering, unknown, negative, positive
    long long n_splits, node_thresh, bps;
    float offset;
    int want_neg, want_pos, interval;
    // Multi-GPU subtree partition (ours): every rank builds the top of the tree replicated; of the frontier that ENTERS level
    // deal_level a rank keeps the nodes deal_owner(i) == deal_rank (below) and drops the rest, so below that level it refines its own
    // subtrees only -- in the same launch, without a host round trip.  deal_world <= 1: no deal.
    long long deal_level;
    int deal_rank, deal_world;
    TreeCtl* ctl;
};
constexpr int kTreeDropped = 0x7f;       // label of a node another rank owns (matches no SIGN_* value: every compaction skips it)

// The deal.  Plain round-robin (i mod world) would be the worst choice here: the children of a level are stored
// [A-children..., B-children...], so the LOW bits of a node's index are its FIRST split decisions -- i mod 8 is the octant of the
// domain, and a shape that does not fill the octants evenly (measured on bunny.npz, depth 21, 8 ranks) leaves the fullest rank
// with 1.44x the mean.  Owner = (sum of the base-`world` digits of i) mod world instead: every aligned group of `world`
// consecutive nodes still gives one node to every rank (so a rank's j-th node is computable), but which member a rank gets
// rotates with the higher digits, i.e. with the later splits.
__host__ __device__ __forceinline__ int deal_digit_sum(long long b, int world) {
    long long s = 0;
    while (b > 0) { s += b % world; b /= world; }
    return (int)(s % world);
}
__host__ __device__ __forceinline__ int deal_owner(long long i, int world) { return (int)((i % world + deal_digit_sum(i / world, world)) % world); }
// index of the j-th node owned by `rank` (one per group of `world`)
__host__ __device__ __forceinline__ long long deal_index(long long j, int rank, int world) {
    return j * world + ((rank - deal_digit_sum(j, world)) % world + world) % world;
}
// nodes of a frontier of n that `rank` owns
__host__ __device__ __forceinline__ long long deal_count(long long n, int rank, int world) {
    const long long nb = n / world, rem = n % world;
    return nb + ((rem > 0 && deal_index(nb, rank, world) - nb * world < rem) ? 1 : 0);
}

// sense-reversing grid barrier (cooperative launch: all CTAs are co-resident).  __threadfence() is a gpu-scope fence: it
// orders this CTA's writes before the arrival and invalidates the SM's L1 after the release, so plain loads of data other
// CTAs wrote before the barrier are safe afterwards.
__device__ __forceinline__ void grid_barrier(TreeCtl* ctl) {
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

__device__ __forceinline__ long long tree_next_bucket(long long s) {       // reference src/bucketing.py:7-14
    long long b = 128;
    while (b < s) b <<= 1;
    return b;
}

// block-wide sum of one long long per thread (all threads get the result)
__device__ __forceinline__ long long block_sum_ll(long long v, long long* scratch /* [kThreads/32] shared */) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    long long s = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) s += scratch[w];
    return s;
}

template <int WMAX, class Tile>
__global__ void __launch_bounds__(kThreads, 1)
k_tree_persistent(const __grid_constant__ NetDev net, const TreeArgs a) {
    using E = Engine<WMAX, Tile>;
    constexpr int RT = Tile::RT;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ long long s_red[kThreads / 32];
    __shared__ int s_warp[kThreads / 32];
    E eng(net, smem);
    // Small levels run on a second engine over the same weights (resident set or streamed ring) that holds ONE tile per thread
    // instead of two (half the FFMA2s and half the epilogue per pass): the top of a tree is pure latency, and a warp whose second
    // tile is empty pays for it all the same.  The two engines are never in a pass at the same time; they share the shared-memory
    // activation region, and for streamed weights the position in the chunk sequence is handed back and forth.
    using E1 = Engine<WMAX, TileOne<Tile>>;
    constexpr bool kHasE1 = Tile::NT > 1;
    E1 eng1(net, smem, false);
    const int tid = threadIdx.x, lane = eng.lane, warp = eng.warp;
    // phase-2 scratch: the CTA's activation buffers are idle between passes (every warp is in phase 2 then)
    int* s_scan = reinterpret_cast<int*>(eng.act - warp * E::WARP_FLOATS);       // [kTreeTile + 1]
    static_assert(kWarps * E::WARP_FLOATS >= kTreeTile + 1, "activation region too small for the tile scan");

    // level state, identical in every thread of the grid (derived from the tile counters, never read back from ctl)
    long long N = a.ctl->n_cur, level = a.ctl->level, bucket = a.ctl->bucket;
    int which = (int)a.ctl->which;
    long long n_fin[2] = {a.ctl->n_fin[0], a.ctl->n_fin[1]};
    long long n_evals = a.ctl->n_evals, max_frontier = a.ctl->max_frontier;
    long long status = 0, need = 0;
    unsigned long long n_tie = a.ctl->n_tie;                          // meaningful in CTA 0, thread 0
    const long long T = a.n_tiles_max;

    while (level < a.n_splits) {
        const bool quit_next = (N >= a.node_thresh) || (level + 1 == a.n_splits);
        const long long this_b = a.bps < bucket ? a.bps : bucket;
        const int par = (int)(level & 1);
        int* cnt = a.tile_cnt + (size_t)par * 3 * T;                  // this level's counters (zeroed during the previous level)
        int* cnt_next = a.tile_cnt + (size_t)(par ^ 1) * 3 * T;
        const float* cur_lo = a.lo[which];
        const float* cur_hi = a.hi[which];
        float* out_lo = a.lo[which ^ 1];
        float* out_hi = a.hi[which ^ 1];
        const long long n_tiles = (N + kTreeTile - 1) / kTreeTile;

        // ---------------- phase 1: classify ----------------
        // A level that fits the grid with HALF of the warps uses warps 0-3 only: each then has its scheduler (and its share of
        // the FMA pipe) to itself, which halves the latency of the pass -- and the top of a tree is nothing but latency:
        // a level costs one pass whether it holds one box or a full wave.
        // the deal (TreeArgs): at this one level the rank classifies only its own nodes j -> i = rank + j * world; the others
        // are labelled "dropped" and vanish in the compaction of phase 2
        const bool dealing = a.deal_world > 1 && level == a.deal_level;
        const long long Nc = dealing ? deal_count(N, a.deal_rank, a.deal_world) : N;
        if (dealing)
            for (long long i = (long long)blockIdx.x * kThreads + tid; i < N; i += (long long)gridDim.x * kThreads)
                if (deal_owner(i, a.deal_world) != a.deal_rank) a.label[i] = kTreeDropped;
        auto run_passes = [&](auto& en, int warps_used) {
            using EE = typename std::remove_reference<decltype(en)>::type;
            const long long pass_boxes = (long long)warps_used * EE::SLOTS;
            const long long n_pass = (Nc + pass_boxes - 1) / pass_boxes;
            for (long long pass = blockIdx.x; pass < n_pass; pass += gridDim.x) {
                if (warp >= warps_used) { en.skip_net(0, net.n_layers); continue; }
                const long long warp_box0 = pass * pass_boxes + (long long)warp * EE::SLOTS;
                if (lane < EE::SLOTS) {
                    const long long j = warp_box0 + lane;
                    const long long i = dealing ? deal_index(j, a.deal_rank, a.deal_world) : j;
                    float4 rows[5];
    #pragma unroll
                    for (int r = 0; r < 5; ++r) rows[r] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (j < Nc) {
                        BoxSource src{};
                        src.kind = 1; src.v = 3; src.a = cur_lo; src.b = cur_hi;
                        src.interval = Tile::rule == 2 ? 0 : a.interval;
                        load_box_rows(src, i, rows);
                    }
                    float* dst = en.slot_ptr(lane);
                    if (Tile::rule == 2) {       // slope_interval: [primal, centre x3, width x3 = 0]
    #pragma unroll
                        for (int r = 0; r < 4; ++r) *reinterpret_cast<float4*>(dst + r * EE::G::S) = rows[r];
    #pragma unroll
                        for (int r = 4; r < RT; ++r) *reinterpret_cast<float4*>(dst + r * EE::G::S) = make_float4(0.f, 0.f, 0.f, 0.f);
                    } else {
    #pragma unroll
                        for (int r = 0; r < 5; ++r) *reinterpret_cast<float4*>(dst + r * EE::G::S) = rows[r];
                    }
                }
                __syncwarp();
                float out[EE::ROWS], ps[EE::ROWS];
                en.run_net(0, net.n_layers, out, ps);
                if (en.cg == 0) {      // label + near-tie flag of every slot, handed to lane `slot` through the warp's scratch
    #pragma unroll
                    for (int nn = 0; nn < EE::NT; ++nn) {
                        const long long j = warp_box0 + nn * EE::G::TPW + en.t;
                        const long long i = dealing ? deal_index(j, a.deal_rank, a.deal_world) : j;
                        int code = 0xff;
                        if (j < Nc) {
                            float lo_b, up_b;
                            if (Tile::rule == 2) {     // src/slope_interval.py:201-206
                                float prad = 0.f;
    #pragma unroll
                                for (int v = 0; v < 3; ++v)
                                    prad = prad + fmaxf(out[nn * RT + 1 + v] + out[nn * RT + 4 + v], -(out[nn * RT + 1 + v] - out[nn * RT + 4 + v]));
                                lo_b = out[nn * RT] - prad; up_b = out[nn * RT] + prad;
                            } else {                   // src/affine.py:119-125
                                const float rad = ((fabsf(out[nn * RT + 1]) + fabsf(out[nn * RT + 2])) + fabsf(out[nn * RT + 3])) + out[nn * RT + 4];
                                lo_b = out[nn * RT] - rad; up_b = out[nn * RT] + rad;
                            }
                            const int lab = label_of(lo_b, up_b, a.offset);
                            a.label[i] = lab;
                            code = lab | (bound_near_tie(lo_b, up_b, a.offset, ps[nn * RT], net.tie_rel) ? 0x100 : 0);
                        }
                        en.fin[(nn * EE::G::TPW + en.t) * 8] = __int_as_float(code);
                    }
                }
                __syncwarp();
                {   // lane s < SLOTS counts slot s (all slots of a warp pass lie in one tile: 2048 % CTA_TILES == 0)
                    int lab = 0xff; bool tie = false;
                    if (lane < EE::SLOTS) {
                        const int code = __float_as_int(en.fin[lane * 8]);
                        lab = code & 0xff; tie = (code & 0x100) != 0;
                    }
                    const unsigned b_unk = __ballot_sync(0xffffffffu, lab == SIGN_UNKNOWN);
                    const unsigned b_neg = __ballot_sync(0xffffffffu, lab == SIGN_NEGATIVE);
                    const unsigned b_pos = __ballot_sync(0xffffffffu, lab == SIGN_POSITIVE);
                    const unsigned b_tie = __ballot_sync(0xffffffffu, tie);
                    if (dealing) {
                        // the slots of a warp pass are ~`world` nodes apart: they fall into several tiles, each lane counts its own
                        if (lane < EE::SLOTS && lab != 0xff) {
                            const long long tile = deal_index(warp_box0 + lane, a.deal_rank, a.deal_world) / kTreeTile;
                            if (lab == SIGN_UNKNOWN) atomicAdd(cnt + tile, 1);
                            if (lab == SIGN_NEGATIVE && a.want_neg) atomicAdd(cnt + T + tile, 1);
                            if (lab == SIGN_POSITIVE && a.want_pos) atomicAdd(cnt + 2 * T + tile, 1);
                        }
                        if (lane == 0 && b_tie) atomicAdd(&a.ctl->n_tie_level, (unsigned long long)__popc(b_tie));
                    } else if (lane == 0 && warp_box0 < N) {
                        const long long tile = warp_box0 / kTreeTile;
                        if (b_unk) atomicAdd(cnt + tile, __popc(b_unk));
                        if (b_neg && a.want_neg) atomicAdd(cnt + T + tile, __popc(b_neg));
                        if (b_pos && a.want_pos) atomicAdd(cnt + 2 * T + tile, __popc(b_pos));
                        if (b_tie) atomicAdd(&a.ctl->n_tie_level, (unsigned long long)__popc(b_tie));
                    }
                }
                __syncwarp();
            }
        };
        if (kHasE1 && Nc <= (long long)gridDim.x * (kWarps / 2) * E1::SLOTS) {
            if (E1::kSparse) {
                // the zero-skipping lists of this engine live where the other engine keeps activations: make every entry a valid
                // offset again (the K loop's look-ahead reads run past a segment's end)
                for (int i = lane; i < E1::LIST_WORDS; i += 32) eng1.lst[i] = 0u;
                __syncwarp();
            }
            eng1.seq_consumed = eng.seq_consumed;
            run_passes(eng1, kWarps / 2);
            eng.seq_consumed = eng1.seq_consumed;
        } else {
            const bool half = Nc <= (long long)gridDim.x * (kWarps / 2) * E::SLOTS;
            run_passes(eng, half ? kWarps / 2 : kWarps);
        }
        grid_barrier(a.ctl);

        // ---------------- phase 2: counts -> sizes -> ordered scatter ----------------
        // totals of the level (every CTA computes the same numbers)
        long long tot[3] = {0, 0, 0};
        {
            long long s0 = 0, s1 = 0, s2 = 0;
            for (long long t = tid; t < n_tiles; t += kThreads) {
                s0 += cnt[t];
                if (a.want_neg) s1 += cnt[T + t];
                if (a.want_pos) s2 += cnt[2 * T + t];
            }
            tot[0] = block_sum_ll(s0, s_red);
            if (a.want_neg) tot[1] = block_sum_ll(s1, s_red);
            if (a.want_pos) tot[2] = block_sum_ll(s2, s_red);
        }
        const long long n_out = quit_next ? tot[0] : 2 * tot[0];
        if (n_out > a.cap) { status = 1; need = n_out; }
        else if (a.want_neg && n_fin[0] + tot[1] > a.fin_cap[0]) { status = 2; need = n_fin[0] + tot[1]; }
        else if (a.want_pos && n_fin[1] + tot[2] > a.fin_cap[1]) { status = 3; need = n_fin[1] + tot[2]; }
        if (blockIdx.x == 0 && tid == 0) {      // every add of this level happened before the barrier; the next ones come after the next
            if (status == 0) n_tie += a.ctl->n_tie_level;
            a.ctl->n_tie_level = 0ull;
        }
        if (status != 0) {
            // stop with the frontier intact; the counters of this level are stale for the relaunch, which redoes the level
            for (long long t = (long long)blockIdx.x * kThreads + tid; t < 3 * T; t += (long long)gridDim.x * kThreads) cnt[t] = 0;
            break;
        }
        // contiguous tile range of this CTA
        const long long per = (n_tiles + gridDim.x - 1) / gridDim.x;
        const long long t0 = (long long)blockIdx.x * per;
        const long long t1 = t0 + per < n_tiles ? t0 + per : n_tiles;
        long long pre[3] = {0, 0, 0};                                 // sum of the counters of the tiles before t0
        if (t0 < t1) {
            long long s0 = 0, s1 = 0, s2 = 0;
            for (long long t = tid; t < t0; t += kThreads) {
                s0 += cnt[t];
                if (a.want_neg) s1 += cnt[T + t];
                if (a.want_pos) s2 += cnt[2 * T + t];
            }
            pre[0] = block_sum_ll(s0, s_red);
            if (a.want_neg) pre[1] = block_sum_ll(s1, s_red);
            if (a.want_pos) pre[2] = block_sum_ll(s2, s_red);
        }
        for (long long tile = t0; tile < t1; ++tile) {
            const long long base_i = tile * kTreeTile;
            // three ordered compactions share one pass over the labels: kind 0 unknown (split / copy), 1 negative, 2 positive
            for (int kind = 0; kind < 3; ++kind) {
                if (kind == 1 && !a.want_neg) continue;
                if (kind == 2 && !a.want_pos) continue;
                const int want_lab = kind == 0 ? SIGN_UNKNOWN : kind == 1 ? SIGN_NEGATIVE : SIGN_POSITIVE;
                // exclusive scan of the flags of this tile in node order: thread owns 8 consecutive nodes
                int f[8], sum = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const long long i = base_i + tid * 8 + k;
                    f[k] = (i < N && a.label[i] == want_lab) ? 1 : 0;
                    sum += f[k];
                }
                int x = sum;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, off); if (lane >= off) x += y; }
                __syncthreads();                                      // previous users of s_warp / s_scan are done
                if (lane == 31) s_warp[warp] = x;
                __syncthreads();
                int wbase = 0;
#pragma unroll
                for (int w = 0; w < kThreads / 32; ++w) if (w < warp) wbase += s_warp[w];
                int run = wbase + x - sum;
                if (kind == 0 && !quit_next && this_b < kTreeTile) {
                    // batches smaller than the tile: the scatter needs the scan at batch boundaries inside the tile
#pragma unroll
                    for (int k = 0; k < 8; ++k) { s_scan[tid * 8 + k] = run; run += f[k]; }
                    if (tid == kThreads - 1) s_scan[kTreeTile] = run;
                    __syncthreads();
                    run = wbase + x - sum;
                }
                if (kind == 0) {
                    // tree split in the reference's order: children of batch [b0,b1) at 2*scan[b0] + {rank, cnt + rank}
                    long long bat_base = pre[0], bat_cnt = 0;          // this_b >= tile: the batch is a whole number of tiles
                    if (!quit_next && this_b >= kTreeTile) {
                        const long long tpb = this_b / kTreeTile;                  // tiles per batch
                        const long long bt0 = (tile / tpb) * tpb;
                        const long long bt1 = bt0 + tpb < n_tiles ? bt0 + tpb : n_tiles;
                        long long before = 0, total = 0;
                        for (long long t = bt0; t < bt1; ++t) { const int c = cnt[t]; total += c; if (t < tile) before += c; }
                        bat_base = pre[0] - before; bat_cnt = total;
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        if (f[k]) {
                            const long long i = base_i + tid * 8 + k;
                            const float l[3] = {cur_lo[3 * i], cur_lo[3 * i + 1], cur_lo[3 * i + 2]};
                            const float h[3] = {cur_hi[3 * i], cur_hi[3 * i + 1], cur_hi[3 * i + 2]};
                            if (quit_next) {
                                const long long o = pre[0] + run;
#pragma unroll
                                for (int d = 0; d < 3; ++d) { out_lo[3 * o + d] = l[d]; out_hi[3 * o + d] = h[d]; }
                            } else {
                                long long base, bcnt, rank;
                                if (this_b >= kTreeTile) {
                                    base = bat_base; bcnt = bat_cnt; rank = pre[0] + run - bat_base;
                                } else {
                                    const int j = tid * 8 + k;
                                    const int b0 = (int)((j / this_b) * this_b);
                                    int b1 = (int)(b0 + this_b);
                                    if (base_i + b1 > N) b1 = (int)(N - base_i);
                                    base = pre[0] + s_scan[b0]; bcnt = s_scan[b1] - s_scan[b0]; rank = run - s_scan[b0];
                                }
                                const long long oa = 2 * base + rank, ob = 2 * base + bcnt + rank;
                                const int sd = argmax3_first(h[0] - l[0], h[1] - l[1], h[2] - l[2]);
#pragma unroll
                                for (int d = 0; d < 3; ++d) {
                                    const float mid = 0.5f * (l[d] + h[d]);
                                    out_lo[3 * oa + d] = l[d];
                                    out_hi[3 * oa + d] = d == sd ? mid : h[d];
                                    out_lo[3 * ob + d] = d == sd ? mid : l[d];
                                    out_hi[3 * ob + d] = h[d];
                                }
                            }
                            run += 1;
                        }
                    }
                } else {
                    float* flo = a.fin_lo[kind - 1];
                    float* fhi = a.fin_hi[kind - 1];
                    const long long dst0 = n_fin[kind - 1] + pre[kind];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        if (f[k]) {
                            const long long i = base_i + tid * 8 + k;
                            const long long o = dst0 + run;
#pragma unroll
                            for (int d = 0; d < 3; ++d) { flo[3 * o + d] = cur_lo[3 * i + d]; fhi[3 * o + d] = cur_hi[3 * i + d]; }
                            run += 1;
                        }
                    }
                }
            }
            pre[0] += cnt[tile];
            if (a.want_neg) pre[1] += cnt[T + tile];
            if (a.want_pos) pre[2] += cnt[2 * T + tile];
        }
        // the next level's counters (other parity) are free now: zero them; record the level
        for (long long t = (long long)blockIdx.x * kThreads + tid; t < 3 * T; t += (long long)gridDim.x * kThreads) cnt_next[t] = 0;
        if (blockIdx.x == 0 && tid == 0 && level < kTreeMaxLevels) {
            a.levels[4 * level] = N; a.levels[4 * level + 1] = tot[0]; a.levels[4 * level + 2] = tot[1]; a.levels[4 * level + 3] = tot[2];
        }
        n_evals += Nc;
        if (N > max_frontier) max_frontier = N;
        n_fin[0] += tot[1]; n_fin[1] += tot[2];
        N = n_out;
        which ^= 1;
        bucket = tree_next_bucket(N);
        level += 1;
        grid_barrier(a.ctl);
        if (quit_next || N == 0) break;          // N == 0: nothing left to refine, no later level can add a node
    }
    if (blockIdx.x == 0 && tid == 0) {
        TreeCtl* c = a.ctl;
        c->n_cur = N; c->which = which; c->bucket = bucket; c->n_fin[0] = n_fin[0]; c->n_fin[1] = n_fin[1];
        c->status = status; c->need = need; c->n_evals = n_evals; c->max_frontier = max_frontier; c->n_tie = n_tie;
        c->level = level;                        // = levels processed so far
    }
    if (kHasE1) eng.exec_macs += eng1.exec_macs;   // one drain: it owns the ring and reports the executed-MAC count of both
    eng.drain();
}

}  // namespace niq
