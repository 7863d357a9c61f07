// niq_tu_rays.cu -- launcher of the persistent k_cast_rays (interval / affine_fixed / slope_interval)
#include "niq_internal.h"

template <int WMAX, class Tile>
static int launch_cast_rays_wt(niq_ctx* c, NetDev net, int total_floats, const CastOpts& o, long long n, int interval,
                               const float* roots, const float* dirs, float* t, int* hit, int* cnt,
                               unsigned char* tie, unsigned long long* queue) {
    using E = Engine<WMAX, Tile>;
    const size_t smem = place_weights<E>(c, net, total_floats);
    TRY(set_smem(k_cast_rays<WMAX, Tile>, smem));
    const long long n_pass = (n + E::CTA_TILES - 1) / E::CTA_TILES;
    LaunchTimer lt(c, 0);
    k_cast_rays<WMAX, Tile><<<grid_for(c, n_pass), kThreads, smem, c->stream>>>(net, o, n, interval, roots, dirs, t, hit, cnt, tie, queue);
    CU(cudaGetLastError());
    return NIQ_OK;
}
template <int WMAX>
static int launch_cast_rays_w(niq_ctx* c, NetDev net, int total_floats, const CastOpts& o, long long n, int interval,
                              const float* roots, const float* dirs, float* t, int* hit, int* cnt,
                              unsigned char* tie, unsigned long long* queue, bool slope = false) {
    if (slope) return launch_cast_rays_wt<WMAX, TileRaySlope>(c, net, total_floats, o, n, 0, roots, dirs, t, hit, cnt, tie, queue);
    return launch_cast_rays_wt<WMAX, TileRay>(c, net, total_floats, o, n, interval, roots, dirs, t, hit, cnt, tie, queue);
}
int launch_cast_rays(niq_ctx* c, int wmax, const NetDev& net, int total_floats, const CastOpts& o, long long n, int interval,
                     const float* roots, const float* dirs, float* t, int* hit, int* cnt, unsigned char* tie,
                     unsigned long long* queue, bool slope) {
    switch (wmax) {
        case 32: return launch_cast_rays_w<32>(c, net, total_floats, o, n, interval, roots, dirs, t, hit, cnt, tie, queue, slope);
        case 64: return launch_cast_rays_w<64>(c, net, total_floats, o, n, interval, roots, dirs, t, hit, cnt, tie, queue, slope);
        case 128: return launch_cast_rays_w<128>(c, net, total_floats, o, n, interval, roots, dirs, t, hit, cnt, tie, queue, slope);
        default: return launch_cast_rays_w<256>(c, net, total_floats, o, n, interval, roots, dirs, t, hit, cnt, tie, queue, slope);
    }
}
