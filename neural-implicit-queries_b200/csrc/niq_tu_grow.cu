// niq_tu_grow.cu -- launchers of the growing-affine-form kernels (affine_all / affine_truncate / affine_append)
#include "niq_internal.h"
#include "niq_grow.cuh"

#include "niq_isect.cuh"
#include "niq_rays_grow.cuh"
#include "niq_frustum_grow.cuh"

// mode -> sizes of the growing-form state (kcap rows of W floats); v = box vectors of the input form
static int make_grow_cfg(niq_ctx* c, const niq_mlp* m, const niq_mode_cfg* cfg, int v, GrowCfg* out) {
    if (m->maxw_pad > 128)
        return fail(NIQ_EUNSUPPORTED, "affine_all / affine_truncate support hidden widths up to 128 (state matrix must fit shared memory)");
    GrowCfg g{};
    g.truncate = cfg->mode == NIQ_MODE_AFFINE_TRUNCATE;
    g.n_keep = g.truncate ? cfg->truncate_count : 0;
    g.n_append = cfg->mode == NIQ_MODE_AFFINE_APPEND ? cfg->truncate_count : 0;
    if (g.truncate && g.n_keep < 0) return fail(NIQ_EINVAL, "affine_truncate: truncate_count must be >= 0");
    if (cfg->mode == NIQ_MODE_AFFINE_APPEND && (g.n_append < 1 || g.n_append > m->min_act_out))
        return fail(NIQ_EINVAL, "affine_append: n_append must be in 1..%d (the narrowest activation layer; jax.lax.top_k needs k <= width)", m->min_act_out);
    g.kcap = g.truncate ? std::max(v, std::min(g.n_keep, v + m->sum_act_out)) + m->max_act_out
             : g.n_append > 0 ? v + g.n_append * m->n_act_layers : v + m->sum_act_out;
    g.kcap = round_up(std::max(g.kcap, 4), 4);   // keeps the aff matrix 16-byte aligned behind mags/rank
    g.W = round_up(m->maxw_pad, 8);
    for (const HostLayer& L : m->layers) g.wfloats = std::max(g.wfloats, round_up(L.in_pad * (L.dot ? 1 : L.out_pad), 4));
    if (grow_state_floats(g) * sizeof(float) > c->prop.sharedMemPerBlockOptin)
        return fail(NIQ_EUNSUPPORTED, "affine state of %zu bytes exceeds shared memory (%zu): network too wide/deep for this mode",
                    grow_state_floats(g) * sizeof(float), (size_t)c->prop.sharedMemPerBlockOptin);
    *out = g;
    return NIQ_OK;
}

int launch_classify_grow(niq_ctx* c, const niq_mlp* m, const niq_mode_cfg* cfg, BoxSource src, long long n,
                         float offset, int* label, float* lower, float* upper, unsigned char* tie) {
    if (n <= 0) return NIQ_OK;
    GrowCfg gc{};
    TRY(make_grow_cfg(c, m, cfg, src.kind == 0 ? src.v : 3, &gc));
    GrowArgs g{};
    g.src = src; g.n = n; g.offset = offset;
    g.truncate = gc.truncate; g.n_keep = gc.n_keep; g.n_append = gc.n_append; g.kcap = gc.kcap; g.W = gc.W;
    g.label = label; g.lower = lower; g.upper = upper; g.near_tie = tie;
    const size_t bytes = grow_state_floats(gc) * sizeof(float);
    TRY(set_smem(k_classify_grow, bytes));
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_classify_grow, 256, bytes);
    per_sm = std::max(per_sm, 1);
    const int grid = (int)std::min<long long>(n, (long long)c->prop.multiProcessorCount * per_sm);
    LaunchTimer lt(c, 0);
    k_classify_grow<<<grid, 256, bytes, c->stream>>>(m->net, g);
    CU(cudaGetLastError());
    return NIQ_OK;
}

static bool is_grow_mode(const niq_mode_cfg* cfg) {
    return cfg->mode == NIQ_MODE_AFFINE_TRUNCATE || cfg->mode == NIQ_MODE_AFFINE_ALL || cfg->mode == NIQ_MODE_AFFINE_APPEND;
}

// R (3x3) + t (3) of a spatial_transformation -> the layer-0 block the kernels read: A0 = inv(R) padded to 4 x 8, b0 = inv(R)@(-t)
// padded to 8 (reference src/affine_layers.py:175-179; the same float32 arithmetic as niq_mlp_create)
static int pack_transforms(const niq_mlp* m, long long n_q, const float* xf, std::vector<float>& out) {
    const HostLayer& L = m->layers[0];
    if (L.in_dim != 3 || L.out_dim != 3 || L.act != ACT_NONE || L.in_pad != 4 || L.out_pad != 8)
        return fail(NIQ_EINVAL, "per-query transforms need an MLP whose first op is a spatial_transformation");
    out.assign((size_t)n_q * 40, 0.f);
    for (long long q = 0; q < n_q; ++q) {
        float inv[9];
        if (!invert3(xf + 12 * q, inv)) return fail(NIQ_EINVAL, "query %lld: R is singular", q);
        float* A0 = out.data() + 40 * q;
        for (int j = 0; j < 3; ++j)
            for (int k = 0; k < 3; ++k) A0[j * 8 + k] = inv[3 * j + k];
        for (int k = 0; k < 3; ++k) {
            float s = 0.f;
            for (int j = 0; j < 3; ++j) s = s + inv[3 * k + j] * (-xf[12 * q + 9 + j]);
            A0[32 + k] = s;
        }
    }
    return NIQ_OK;
}

int isect_grow_batch(niq_ctx* c, const niq_mlp* mA, const niq_mode_cfg* cfgA, const niq_mlp* mB, const niq_mode_cfg* cfgB,
                     long long n_q, const float* xfA, const float* xfB, const float lower[3], const float upper[3], float eps,
                     int32_t* found, float* loc, int64_t* stats, bool* handled) {
    *handled = false;
    if (!is_grow_mode(cfgA) || !is_grow_mode(cfgB)) return NIQ_OK;
    if (const char* e = getenv("NIQ_ISECT_LEGACY")) if (e[0] != '0') return NIQ_OK;       // A/B knob: the per-round host loop
    int coop = 0;
    CU(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device));
    if (!coop) return NIQ_OK;
    *handled = true;
    IsectArgs a{};
    TRY(make_grow_cfg(c, mA, cfgA, 3, &a.g[0]));
    TRY(make_grow_cfg(c, mB, cfgB, 3, &a.g[1]));
    a.cg_lanes[0] = mA->wmax / 8; a.cg_lanes[1] = mB->wmax / 8;
    a.W = std::max(a.g[0].W, a.g[1].W);
    a.state_floats = (long long)std::max(grow_state_floats(a.g[0]), grow_state_floats(a.g[1]));
    a.state_floats = (a.state_floats + 3) / 4 * 4;
    const size_t smem = ((size_t)a.state_floats + 16 * (size_t)a.W) * sizeof(float);
    if (smem > c->prop.sharedMemPerBlockOptin)
        return fail(NIQ_EUNSUPPORTED, "intersection state of %zu bytes exceeds shared memory", smem);
    TRY(set_smem(k_isect_persistent, smem));
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_isect_persistent, 256, smem));
    if (per_sm < 1) return fail(NIQ_ECUDA, "persistent intersection kernel does not fit an SM");
    a.n_queries = n_q;
    a.eps_w = eps / sqrtf(3.0f);                 // reference src/kd_tree.py:446
    a.max_rounds = 512;
    std::vector<float> hxf[2];
    DevBuf dxf0(c), dxf1(c);
    if (xfA) { TRY(pack_transforms(mA, n_q, xfA, hxf[0])); TRY(dxf0.alloc(hxf[0].size() * 4)); CU(cudaMemcpyAsync(dxf0.p, hxf[0].data(), hxf[0].size() * 4, cudaMemcpyHostToDevice, c->stream)); a.xf[0] = dxf0.as<float>(); }
    if (xfB) { TRY(pack_transforms(mB, n_q, xfB, hxf[1])); TRY(dxf1.alloc(hxf[1].size() * 4)); CU(cudaMemcpyAsync(dxf1.p, hxf[1].data(), hxf[1].size() * 4, cudaMemcpyHostToDevice, c->stream)); a.xf[1] = dxf1.as<float>(); }

    // a query's frontier peaks at a few hundred nodes; the buffers hold 1,024 per query to start with and
    // the whole batch is redone with larger ones in the (never observed) case that a round does not fit
    long long cap = std::max<long long>(8192, 1024 * n_q);
    for (int attempt = 0; attempt < 6; ++attempt, cap *= 4) {
        a.cap = cap;
        a.n_tiles_max = cap / kTreeTile + 2;
        DevBuf lo0(c), lo1(c), hi0(c), hi1(c), q0(c), q1(c), lab(c), vals(c), tie(c), needs(c), locs(c), tcnt(c), first(c), qf(c), ql(c), qs(c), ctl(c);
        TRY(lo0.alloc(cap * 12)); TRY(lo1.alloc(cap * 12)); TRY(hi0.alloc(cap * 12)); TRY(hi1.alloc(cap * 12));
        TRY(q0.alloc(cap * 4)); TRY(q1.alloc(cap * 4)); TRY(lab.alloc(cap * 8)); TRY(vals.alloc(cap * 56)); TRY(tie.alloc(cap * 2));
        TRY(needs.alloc(cap * 4)); TRY(locs.alloc(cap * 12)); TRY(tcnt.alloc(a.n_tiles_max * 8));
        TRY(first.alloc(n_q * 8)); TRY(qf.alloc(n_q * 4)); TRY(ql.alloc(n_q * 12)); TRY(qs.alloc(n_q * 24)); TRY(ctl.alloc(sizeof(IsectCtl)));
        a.lo[0] = lo0.as<float>(); a.lo[1] = lo1.as<float>(); a.hi[0] = hi0.as<float>(); a.hi[1] = hi1.as<float>();
        a.qid[0] = q0.as<int>(); a.qid[1] = q1.as<int>(); a.lab = lab.as<int>(); a.vals = vals.as<float>(); a.tie = tie.as<unsigned char>();
        a.needs = needs.as<int>(); a.loc = locs.as<float>(); a.tile_cnt = tcnt.as<int>(); a.first = first.as<unsigned long long>();
        a.q_found = qf.as<int>(); a.q_loc = ql.as<float>(); a.q_stats = qs.as<long long>(); a.ctl = ctl.as<IsectCtl>();
        {   // initial frontier: one root box per query (src/kd_tree.py:581-590), query id = index
            std::vector<float> hl((size_t)n_q * 3), hh((size_t)n_q * 3), hloc((size_t)n_q * 3, -777.f);
            std::vector<int> hq((size_t)n_q);
            for (long long q = 0; q < n_q; ++q) {
                for (int d = 0; d < 3; ++d) { hl[3 * q + d] = lower[d]; hh[3 * q + d] = upper[d]; }
                hq[q] = (int)q;
            }
            CU(cudaMemcpyAsync(a.lo[0], hl.data(), hl.size() * 4, cudaMemcpyHostToDevice, c->stream));
            CU(cudaMemcpyAsync(a.hi[0], hh.data(), hh.size() * 4, cudaMemcpyHostToDevice, c->stream));
            CU(cudaMemcpyAsync(a.qid[0], hq.data(), hq.size() * 4, cudaMemcpyHostToDevice, c->stream));
            CU(cudaMemcpyAsync(a.q_loc, hloc.data(), hloc.size() * 4, cudaMemcpyHostToDevice, c->stream));
            CU(cudaStreamSynchronize(c->stream));      // the staging vectors die with this scope
        }
        CU(cudaMemsetAsync(a.tile_cnt, 0, a.n_tiles_max * 8, c->stream));
        CU(cudaMemsetAsync(a.first, 0, n_q * 8, c->stream));
        CU(cudaMemsetAsync(a.q_found, 0, n_q * 4, c->stream));
        CU(cudaMemsetAsync(a.q_stats, 0, n_q * 24, c->stream));
        IsectCtl h{};
        h.n_cur = n_q;
        CU(cudaMemcpyAsync(a.ctl, &h, sizeof(h), cudaMemcpyHostToDevice, c->stream));
        NetDev nA = mA->net, nB = mB->net;
        nA.exec_macs = nullptr; nB.exec_macs = nullptr;
        // one CTA per (node, shape) item in flight: the grid is what fits the device, at most what the first rounds can use
        const long long want = std::max<long long>(2 * n_q * 512, 2 * 512);
        const int grid = (int)std::min<long long>((long long)per_sm * c->prop.multiProcessorCount, want);
        void* args[] = {(void*)&nA, (void*)&nB, (void*)&a};
        {
            LaunchTimer lt(c, 0);
            CU(cudaLaunchCooperativeKernel((const void*)k_isect_persistent, dim3(grid), dim3(256), args, smem, c->stream));
        }
        timer_mark(c);
        CU(cudaMemcpyAsync(&h, a.ctl, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        if (h.status == 2) return fail(NIQ_ECUDA, "find_any_intersection: no verdict after %d rounds", a.max_rounds);
        if (h.status == 1) continue;               // frontier outgrew the buffers: redo the batch with 4x the capacity
        std::vector<int> hf((size_t)n_q);
        std::vector<long long> hs((size_t)n_q * 3);
        CU(cudaMemcpyAsync(hf.data(), a.q_found, n_q * 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(loc, a.q_loc, n_q * 12, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(hs.data(), a.q_stats, n_q * 24, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        for (long long q = 0; q < n_q; ++q) {
            found[q] = hf[q];
            if (!hf[q]) loc[3 * q] = loc[3 * q + 1] = loc[3 * q + 2] = -777.f;
            if (stats) for (int k = 0; k < 3; ++k) stats[3 * q + k] = hs[3 * q + k];
        }
        return NIQ_OK;
    }
    return fail(NIQ_ENOMEM, "find_any_intersection: frontier kept outgrowing its buffers");
}

// cast_rays in a growing-form mode: one persistent kernel, one CTA per ray in flight (niq_rays_grow.cuh)
int launch_cast_rays_grow(niq_ctx* c, int n_funcs, const niq_mlp* const* mlps, const niq_mode_cfg* cfgs, const NetDev& net,
                          const CastOpts& o, long long n, const float* roots, const float* dirs, float* t, int* hit, int* cnt,
                          unsigned char* tie, unsigned long long* queue) {
    if (n_funcs > kMaxRayFuncs) return fail(NIQ_EUNSUPPORTED, "cast_rays in a growing-form mode supports up to %d funcs", kMaxRayFuncs);
    RayGrowArgs a{};
    a.o = o; a.n = n; a.roots = roots; a.dirs = dirs; a.out_t = t; a.out_hit = hit; a.out_count = cnt; a.out_tie = tie; a.queue = queue;
    a.n_funcs = n_funcs;
    size_t state = 0;
    for (int f = 0; f < n_funcs; ++f) {
        if (!is_grow_mode(&cfgs[f])) return fail(NIQ_EINVAL, "launch_cast_rays_grow: func %d is not in a growing-form mode", f);
        TRY(make_grow_cfg(c, mlps[f], &cfgs[f], 1, &a.g[f]));
        a.cg_lanes[f] = mlps[f]->wmax / 8;
        a.W = std::max(a.W, a.g[f].W);
        state = std::max(state, grow_state_floats(a.g[f]));
    }
    a.state_floats = (long long)((state + 3) / 4 * 4);
    const size_t smem = ((size_t)a.state_floats + 16 * (size_t)a.W) * sizeof(float);
    if (smem > c->prop.sharedMemPerBlockOptin) return fail(NIQ_EUNSUPPORTED, "ray state of %zu bytes exceeds shared memory", smem);
    TRY(set_smem(k_cast_rays_grow, smem));
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cast_rays_grow, 256, smem);
    per_sm = std::max(per_sm, 1);
    const int grid = (int)std::min<long long>(n, (long long)c->prop.multiProcessorCount * per_sm);
    NetDev nd = net;
    nd.exec_macs = nullptr;
    LaunchTimer lt(c, 0);
    k_cast_rays_grow<<<grid, 256, smem, c->stream>>>(nd, a);
    CU(cudaGetLastError());
    return NIQ_OK;
}

// cast_rays_frustum in a growing-form mode: one persistent kernel, one CTA per frustum in flight (niq_frustum_grow.cuh)
int launch_cast_frustum_grow(niq_ctx* c, int n_funcs, const niq_mlp* const* mlps, const niq_mode_cfg* cfgs, const NetDev& net,
                             const CastOpts& o, const FrustCam& cam, const FrustQueue& q, long long n_pixels) {
    if (n_funcs > 4) return fail(NIQ_EUNSUPPORTED, "cast_rays_frustum in a growing-form mode supports up to 4 funcs");
    FrustGrowArgs a{};
    a.o = o; a.q = q; a.n_funcs = n_funcs;
    size_t state = 0;
    for (int f = 0; f < n_funcs; ++f) {
        if (!is_grow_mode(&cfgs[f])) return fail(NIQ_EINVAL, "launch_cast_frustum_grow: func %d is not in a growing-form mode", f);
        TRY(make_grow_cfg(c, mlps[f], &cfgs[f], 3, &a.g[f]));
        a.cg_lanes[f] = mlps[f]->wmax / 8;
        a.W = std::max(a.W, a.g[f].W);
        state = std::max(state, grow_state_floats(a.g[f]));
    }
    a.state_floats = (long long)((state + 3) / 4 * 4);
    const size_t smem = ((size_t)a.state_floats + 16 * (size_t)a.W) * sizeof(float);
    if (smem > c->prop.sharedMemPerBlockOptin) return fail(NIQ_EUNSUPPORTED, "frustum state of %zu bytes exceeds shared memory", smem);
    TRY(set_smem(k_cast_frustum_grow, smem));
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cast_frustum_grow, 256, smem);
    per_sm = std::max(per_sm, 1);
    // idle CTAs poll the queue: never more CTAs than frusta can exist, all of them co-resident (a waiting CTA must not starve a queued one)
    const int grid = (int)std::min<long long>(n_pixels, (long long)c->prop.multiProcessorCount * per_sm);
    NetDev nd = net;
    nd.exec_macs = nullptr;
    LaunchTimer lt(c, 0);
    k_cast_frustum_grow<<<grid, 256, smem, c->stream>>>(nd, cam, a);
    CU(cudaGetLastError());
    return NIQ_OK;
}
