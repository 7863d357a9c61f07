// niq_tu_tree.cu -- host driver of the persistent level-set tree kernel (niq_tree.cuh): buffer sizing, the cooperative
// launch, the capacity-growth relaunch, and the hand-over of the device-resident node lists to the niq_tree object.
// Replaces the per-level host loop of reference src/kd_tree.py:137-198 for the fixed-row modes.
#include "niq_internal.h"
#include "niq_tree.cuh"

namespace {

struct TreeBufs {
    niq_ctx* c;
    float* lo[2] = {nullptr, nullptr}; float* hi[2] = {nullptr, nullptr};
    float* fin_lo[2] = {nullptr, nullptr}; float* fin_hi[2] = {nullptr, nullptr};
    int* label = nullptr; int* tile_cnt = nullptr; long long* levels = nullptr; TreeCtl* ctl = nullptr;
    long long cap = 0, fin_cap[2] = {0, 0}, n_tiles_max = 0;
    explicit TreeBufs(niq_ctx* ctx) : c(ctx) {}
    ~TreeBufs() {
        void* all[] = {lo[0], lo[1], hi[0], hi[1], fin_lo[0], fin_lo[1], fin_hi[0], fin_hi[1], label, tile_cnt, levels, ctl};
        for (void* p : all) if (p) cudaFreeAsync(p, c->stream);
    }
};

template <class T>
int alloc_async(niq_ctx* c, T** p, size_t count) {
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(p), std::max<size_t>(count, 4) * sizeof(T), c->stream);
    if (e != cudaSuccess) { *p = nullptr; return fail(NIQ_ENOMEM, "cudaMallocAsync(%zu) failed: %s", count * sizeof(T), cudaGetErrorString(e)); }
    return NIQ_OK;
}

// (re)size the frontier buffers to `cap` nodes, keeping the first `keep` nodes of buffer `which`
int resize_frontier(niq_ctx* c, TreeBufs& b, long long cap, int which, long long keep) {
    float *nlo[2], *nhi[2];
    for (int k = 0; k < 2; ++k) { TRY(alloc_async(c, &nlo[k], (size_t)cap * 3)); TRY(alloc_async(c, &nhi[k], (size_t)cap * 3)); }
    if (keep > 0) {
        CU(cudaMemcpyAsync(nlo[which], b.lo[which], (size_t)keep * 12, cudaMemcpyDeviceToDevice, c->stream));
        CU(cudaMemcpyAsync(nhi[which], b.hi[which], (size_t)keep * 12, cudaMemcpyDeviceToDevice, c->stream));
    }
    for (int k = 0; k < 2; ++k) {
        if (b.lo[k]) cudaFreeAsync(b.lo[k], c->stream);
        if (b.hi[k]) cudaFreeAsync(b.hi[k], c->stream);
        b.lo[k] = nlo[k]; b.hi[k] = nhi[k];
    }
    if (b.label) cudaFreeAsync(b.label, c->stream);
    if (b.tile_cnt) cudaFreeAsync(b.tile_cnt, c->stream);
    b.label = nullptr; b.tile_cnt = nullptr;
    b.cap = cap;
    b.n_tiles_max = cap / kTreeTile + 2;
    TRY(alloc_async(c, &b.label, (size_t)cap));
    TRY(alloc_async(c, &b.tile_cnt, (size_t)6 * b.n_tiles_max));
    CU(cudaMemsetAsync(b.tile_cnt, 0, (size_t)6 * b.n_tiles_max * sizeof(int), c->stream));
    return NIQ_OK;
}
int resize_fin(niq_ctx* c, TreeBufs& b, int k, long long cap, long long keep) {
    float *nlo = nullptr, *nhi = nullptr;
    TRY(alloc_async(c, &nlo, (size_t)cap * 3)); TRY(alloc_async(c, &nhi, (size_t)cap * 3));
    if (keep > 0) {
        CU(cudaMemcpyAsync(nlo, b.fin_lo[k], (size_t)keep * 12, cudaMemcpyDeviceToDevice, c->stream));
        CU(cudaMemcpyAsync(nhi, b.fin_hi[k], (size_t)keep * 12, cudaMemcpyDeviceToDevice, c->stream));
    }
    if (b.fin_lo[k]) cudaFreeAsync(b.fin_lo[k], c->stream);
    if (b.fin_hi[k]) cudaFreeAsync(b.fin_hi[k], c->stream);
    b.fin_lo[k] = nlo; b.fin_hi[k] = nhi; b.fin_cap[k] = cap;
    return NIQ_OK;
}

template <int WMAX, class Tile>
int launch_tree_wt(niq_ctx* c, NetDev net, int total_floats, const TreeArgs& a) {
    using E = Engine<WMAX, Tile>;
    const size_t smem = place_weights<E>(c, net, total_floats);
    auto kernel = k_tree_persistent<WMAX, Tile>;
    TRY(set_smem(kernel, smem));
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem));
    if (per_sm < 1) return fail(NIQ_ECUDA, "persistent tree kernel does not fit an SM (%zu bytes of shared memory)", smem);
    void* args[] = {(void*)&net, (void*)&a};
    LaunchTimer lt(c, 0);
    CU(cudaLaunchCooperativeKernel((const void*)kernel, dim3(c->prop.multiProcessorCount), dim3(kThreads), args, smem, c->stream));
    return NIQ_OK;
}
template <class Tile>
int launch_tree_t(niq_ctx* c, const niq_mlp* m, const TreeArgs& a) {
    switch (m->wmax) {
        case 32: return launch_tree_wt<32, Tile>(c, m->net, m->total_floats, a);
        case 64: return launch_tree_wt<64, Tile>(c, m->net, m->total_floats, a);
        case 128: return launch_tree_wt<128, Tile>(c, m->net, m->total_floats, a);
        default: return launch_tree_wt<256, Tile>(c, m->net, m->total_floats, a);
    }
}

}  // namespace

int tree_build_persistent(niq_ctx* c, const niq_mlp* m, const niq_mode_cfg* cfg, long long n_roots, const float* lower,
                          const float* upper, int split_depth, long long node_thresh, float offset, int flags, int bps,
                          niq_tree* T, bool* handled, int deal_level, int deal_rank, int deal_world) {
    const bool fixed = cfg->mode == NIQ_MODE_INTERVAL || cfg->mode == NIQ_MODE_AFFINE_FIXED;
    const bool slope = cfg->mode == NIQ_MODE_SLOPE_INTERVAL;
    *handled = false;
    if (!fixed && !slope) return NIQ_OK;
    if (const char* e = getenv("NIQ_TREE_LEGACY")) if (e[0] != '0') return NIQ_OK;     // A/B knob: the per-level host loop
    int coop = 0;
    CU(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device));
    if (!coop) return NIQ_OK;
    *handled = true;

    const long long n_splits = split_depth < 0 ? (1ll << 40) : (long long)split_depth + 1;
    // Frontier capacity: the full tree at split_depth, or twice the terminate threshold (a level that enters with fewer
    // nodes than the threshold at most doubles), whichever is smaller -- bounded at first and grown on demand.
    long long kCapFirst = 4ll << 20;
    if (const char* e = getenv("NIQ_TREE_CAP")) kCapFirst = std::max(1ll, atoll(e));       // test knob: force the growth path
    long long want = kCapFirst;
    if (split_depth >= 0 && split_depth < 40) want = std::min(want, n_roots << std::min(split_depth, 40));
    // a dealt build holds the replicated top frontier or its share of the bottom (x2 for imbalance), whichever is larger
    if (deal_world > 1) want = std::min(want, std::max(want / deal_world * 2, (long long)n_roots << std::min(std::max(deal_level, 0), 40)));
    if (node_thresh < (1ll << 40)) want = std::min(want, std::max(2 * node_thresh, n_roots));
    want = std::max<long long>(std::max(want, n_roots), getenv("NIQ_TREE_CAP") ? 1 : 4096);

    TreeBufs b(c);
    TRY(resize_frontier(c, b, want, 0, 0));
    CU(cudaMemcpyAsync(b.lo[0], lower, (size_t)n_roots * 12, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(b.hi[0], upper, (size_t)n_roots * 12, cudaMemcpyHostToDevice, c->stream));
    const bool want_fin[2] = {(flags & NIQ_TREE_INTERIOR) != 0, (flags & NIQ_TREE_EXTERIOR) != 0};
    for (int k = 0; k < 2; ++k)
        if (want_fin[k]) TRY(resize_fin(c, b, k, std::max<long long>(want, getenv("NIQ_TREE_CAP") ? 1 : (1 << 16)), 0));
    TRY(alloc_async(c, &b.levels, (size_t)4 * kTreeMaxLevels));
    CU(cudaMemsetAsync(b.levels, 0, (size_t)4 * kTreeMaxLevels * sizeof(long long), c->stream));
    TRY(alloc_async(c, &b.ctl, 1));
    TreeCtl h{};
    h.n_cur = n_roots;
    h.bucket = 1;                                                  // padded array size the reference would hold (src/kd_tree.py:140)
    if (n_roots > 1) { long long p = 128; while (p < n_roots) p <<= 1; h.bucket = p; }
    CU(cudaMemcpyAsync(b.ctl, &h, sizeof(h), cudaMemcpyHostToDevice, c->stream));

    for (int attempt = 0; attempt < 64; ++attempt) {
        TreeArgs a{};
        for (int k = 0; k < 2; ++k) { a.lo[k] = b.lo[k]; a.hi[k] = b.hi[k]; a.fin_lo[k] = b.fin_lo[k]; a.fin_hi[k] = b.fin_hi[k]; a.fin_cap[k] = b.fin_cap[k]; }
        a.cap = b.cap; a.label = b.label; a.tile_cnt = b.tile_cnt; a.n_tiles_max = b.n_tiles_max; a.levels = b.levels;
        a.n_splits = n_splits; a.node_thresh = node_thresh; a.bps = bps; a.offset = offset;
        a.want_neg = want_fin[0]; a.want_pos = want_fin[1]; a.interval = cfg->mode == NIQ_MODE_INTERVAL; a.ctl = b.ctl;
        a.deal_level = deal_level; a.deal_rank = deal_rank; a.deal_world = deal_world;
        if (slope) TRY(launch_tree_t<TileSlope3>(c, m, a));
        else TRY(launch_tree_t<TileBox3>(c, m, a));
        timer_mark(c);
        CU(cudaMemcpyAsync(&h, b.ctl, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        if (h.status == 0) break;
        // a level did not fit: grow the buffer it named (x2 beyond the need) and relaunch from that level
        if (h.status == 1) TRY(resize_frontier(c, b, std::max(2 * h.need, 2 * b.cap), (int)h.which, h.n_cur));
        else TRY(resize_fin(c, b, (int)h.status - 2, std::max(2 * h.need, 2 * b.fin_cap[h.status - 2]), h.n_fin[h.status - 2]));
        if (attempt == 63) return fail(NIQ_ENOMEM, "level-set tree: buffers kept growing (frontier of %lld nodes)", h.need);
    }
    // results: the frontier buffer `which` holds the UNKNOWN leaves; the lists are handed to the tree object as they are
    const int w = (int)h.which;
    T->lists[0].lo = b.lo[w]; T->lists[0].hi = b.hi[w]; T->lists[0].n = h.n_cur; T->lists[0].cap = b.cap;
    b.lo[w] = nullptr; b.hi[w] = nullptr;
    for (int k = 0; k < 2; ++k) {
        if (!want_fin[k]) continue;
        T->lists[1 + k].lo = b.fin_lo[k]; T->lists[1 + k].hi = b.fin_hi[k]; T->lists[1 + k].n = h.n_fin[k]; T->lists[1 + k].cap = b.fin_cap[k];
        b.fin_lo[k] = nullptr; b.fin_hi[k] = nullptr;
    }
    T->stats[0] = h.n_evals; T->stats[1] = (long long)h.n_tie; T->stats[2] = h.level; T->stats[3] = h.max_frontier;
    const long long nl = std::min<long long>(h.level, kTreeMaxLevels);
    T->levels.assign((size_t)4 * h.level, 0);
    if (nl > 0) {
        CU(cudaMemcpyAsync(T->levels.data(), b.levels, (size_t)4 * nl * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    return NIQ_OK;
}
