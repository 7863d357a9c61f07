// niq_tu_frustum.cu -- launcher of the persistent k_cast_frustum
#include "niq_internal.h"

template <int WMAX, class Tile>
static int launch_cast_frustum_wt(niq_ctx* c, NetDev net, int total_floats, const CastOpts& o, const FrustCam& cam, int interval,
                                  const FrustQueue& q, long long n_pixels) {
    using E = Engine<WMAX, Tile>;
    const size_t smem = place_weights<E>(c, net, total_floats);
    TRY(set_smem(k_cast_frustum<WMAX, Tile>, smem));
    const long long n_pass = (n_pixels + E::CTA_TILES - 1) / E::CTA_TILES;
    LaunchTimer lt(c, 0);
    k_cast_frustum<WMAX, Tile><<<grid_for(c, n_pass), kThreads, smem, c->stream>>>(net, o, cam, interval, q);
    CU(cudaGetLastError());
    return NIQ_OK;
}
template <int WMAX>
static int launch_cast_frustum_w(niq_ctx* c, NetDev net, int total_floats, const CastOpts& o, const FrustCam& cam, int interval,
                                 const FrustQueue& q, long long n_pixels, bool slope) {
    if (slope) return launch_cast_frustum_wt<WMAX, TileFrustumSlope>(c, net, total_floats, o, cam, 0, q, n_pixels);
    return launch_cast_frustum_wt<WMAX, TileFrustum>(c, net, total_floats, o, cam, interval, q, n_pixels);
}
int launch_cast_frustum(niq_ctx* c, int wmax, const NetDev& net, int total_floats, const CastOpts& o, const FrustCam& cam,
                        int interval, const FrustQueue& q, long long n_pixels, bool slope) {
    switch (wmax) {
        case 32: return launch_cast_frustum_w<32>(c, net, total_floats, o, cam, interval, q, n_pixels, slope);
        case 64: return launch_cast_frustum_w<64>(c, net, total_floats, o, cam, interval, q, n_pixels, slope);
        case 128: return launch_cast_frustum_w<128>(c, net, total_floats, o, cam, interval, q, n_pixels, slope);
        default: return launch_cast_frustum_w<256>(c, net, total_floats, o, cam, interval, q, n_pixels, slope);
    }
}
