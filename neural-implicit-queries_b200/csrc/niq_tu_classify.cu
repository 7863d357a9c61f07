// niq_tu_classify.cu -- launchers of k_classify_fixed / k_classify_slope (interval, affine_fixed, slope_interval box classification)
#include "niq_internal.h"

template <int WMAX>
static int launch_classify_fixed_w(niq_ctx* c, NetDev net, int total_floats, const BoxSource& src, long long n, float offset,
                                   int* label, float* lower, float* upper, unsigned char* tie) {
    using E = Engine<WMAX, TileBox3>;
    const size_t smem = place_weights<E>(c, net, total_floats);
    TRY(set_smem(k_classify_fixed<WMAX>, smem));
    const long long n_pass = (n + E::CTA_TILES - 1) / E::CTA_TILES;
    LaunchTimer lt(c, 0);
    k_classify_fixed<WMAX><<<grid_for(c, n_pass), kThreads, smem, c->stream>>>(net, src, n, offset, label, lower, upper, tie);
    CU(cudaGetLastError());
    return NIQ_OK;
}
template <int WMAX>
static int launch_classify_slope_w(niq_ctx* c, NetDev net, int total_floats, const BoxSource& src, long long n, float offset,
                                   int* label, float* lower, float* upper, unsigned char* tie, float* raw, float* raw_scale) {
    using E = Engine<WMAX, TileSlope3>;
    const size_t smem = place_weights<E>(c, net, total_floats);
    TRY(set_smem(k_classify_slope<WMAX>, smem));
    const long long n_pass = (n + E::CTA_TILES - 1) / E::CTA_TILES;
    LaunchTimer lt(c, 0);
    k_classify_slope<WMAX><<<grid_for(c, n_pass), kThreads, smem, c->stream>>>(net, src, n, offset, label, lower, upper, tie, raw, raw_scale);
    CU(cudaGetLastError());
    return NIQ_OK;
}
int launch_classify_slope(niq_ctx* c, const niq_mlp* m, const BoxSource& src, long long n, float offset,
                                 int* label, float* lower, float* upper, unsigned char* tie, float* raw, float* raw_scale) {
    if (n <= 0) return NIQ_OK;
    switch (m->wmax) {
        case 32: return launch_classify_slope_w<32>(c, m->net, m->total_floats, src, n, offset, label, lower, upper, tie, raw, raw_scale);
        case 64: return launch_classify_slope_w<64>(c, m->net, m->total_floats, src, n, offset, label, lower, upper, tie, raw, raw_scale);
        case 128: return launch_classify_slope_w<128>(c, m->net, m->total_floats, src, n, offset, label, lower, upper, tie, raw, raw_scale);
        default: return launch_classify_slope_w<256>(c, m->net, m->total_floats, src, n, offset, label, lower, upper, tie, raw, raw_scale);
    }
}
int launch_classify_fixed(niq_ctx* c, const niq_mlp* m, const BoxSource& src, long long n, float offset,
                                 int* label, float* lower, float* upper, unsigned char* tie) {
    if (n <= 0) return NIQ_OK;
    switch (m->wmax) {
        case 32: return launch_classify_fixed_w<32>(c, m->net, m->total_floats, src, n, offset, label, lower, upper, tie);
        case 64: return launch_classify_fixed_w<64>(c, m->net, m->total_floats, src, n, offset, label, lower, upper, tie);
        case 128: return launch_classify_fixed_w<128>(c, m->net, m->total_floats, src, n, offset, label, lower, upper, tie);
        default: return launch_classify_fixed_w<256>(c, m->net, m->total_floats, src, n, offset, label, lower, upper, tie);
    }
}
