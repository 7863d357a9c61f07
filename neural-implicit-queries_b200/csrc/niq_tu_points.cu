// niq_tu_points.cu -- launcher of k_eval_points (plain f(x) rows)
#include "niq_internal.h"

template <int WMAX>
static int launch_eval_points_w(niq_ctx* c, NetDev net, int total_floats, const PointSource& src, long long n, float* f, float* scale) {
    using E = Engine<WMAX, TilePts>;
    const size_t smem = place_weights<E>(c, net, total_floats);
    TRY(set_smem(k_eval_points<WMAX>, smem));
    const long long per = kWarps * E::WARP_ROWS;
    LaunchTimer lt(c, 0);
    k_eval_points<WMAX><<<grid_for(c, (n + per - 1) / per), kThreads, smem, c->stream>>>(net, src, n, f, scale);
    CU(cudaGetLastError());
    return NIQ_OK;
}
int launch_eval_points(niq_ctx* c, const niq_mlp* m, const PointSource& src, long long n, float* f, float* scale) {
    if (n <= 0) return NIQ_OK;
    switch (m->wmax) {
        case 32: return launch_eval_points_w<32>(c, m->net, m->total_floats, src, n, f, scale);
        case 64: return launch_eval_points_w<64>(c, m->net, m->total_floats, src, n, f, scale);
        case 128: return launch_eval_points_w<128>(c, m->net, m->total_floats, src, n, f, scale);
        default: return launch_eval_points_w<256>(c, m->net, m->total_floats, src, n, f, scale);
    }
}
