// niq_tu_grow.cu -- launchers of the growing-affine-form kernels (affine_all / affine_truncate / affine_append)
#include "niq_internal.h"
#include "niq_grow.cuh"

int launch_classify_grow(niq_ctx* c, const niq_mlp* m, const niq_mode_cfg* cfg, BoxSource src, long long n,
                                float offset, int* label, float* lower, float* upper, unsigned char* tie) {
    if (n <= 0) return NIQ_OK;
    if (m->maxw_pad > 128)
        return fail(NIQ_EUNSUPPORTED, "affine_all / affine_truncate support hidden widths up to 128 (state matrix must fit shared memory)");
    GrowArgs g{};
    g.src = src; g.n = n; g.offset = offset;
    g.truncate = cfg->mode == NIQ_MODE_AFFINE_TRUNCATE;
    g.n_keep = g.truncate ? cfg->truncate_count : 0;
    g.n_append = cfg->mode == NIQ_MODE_AFFINE_APPEND ? cfg->truncate_count : 0;
    const int v = src.kind == 0 ? src.v : 3;
    if (g.truncate && g.n_keep < 0) return fail(NIQ_EINVAL, "affine_truncate: truncate_count must be >= 0");
    if (cfg->mode == NIQ_MODE_AFFINE_APPEND && (g.n_append < 1 || g.n_append > m->min_act_out))
        return fail(NIQ_EINVAL, "affine_append: n_append must be in 1..%d (the narrowest activation layer; jax.lax.top_k needs k <= width)", m->min_act_out);
    g.kcap = g.truncate ? std::max(v, std::min(g.n_keep, v + m->sum_act_out)) + m->max_act_out
             : g.n_append > 0 ? v + g.n_append * m->n_act_layers : v + m->sum_act_out;
    g.kcap = round_up(std::max(g.kcap, 4), 4);   // keeps the aff matrix 16-byte aligned behind mags/rank
    g.W = round_up(m->maxw_pad, 8);
    g.label = label; g.lower = lower; g.upper = upper; g.near_tie = tie;
    const size_t floats = (size_t)22 * g.W + 2 * (size_t)g.kcap + (size_t)g.kcap * g.W * (g.truncate ? 2 : 1) + 16;
    const size_t bytes = floats * sizeof(float);
    if (bytes > c->prop.sharedMemPerBlockOptin)
        return fail(NIQ_EUNSUPPORTED, "affine state of %zu bytes exceeds shared memory (%zu): network too wide/deep for this mode", bytes, (size_t)c->prop.sharedMemPerBlockOptin);
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
