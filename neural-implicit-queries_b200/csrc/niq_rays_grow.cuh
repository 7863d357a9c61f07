// niq_rays_grow.cuh -- cast_rays for the growing-form modes (affine_truncate / affine_all / affine_append) as one persistent
// kernel (reference src/queries.py:39-175).  One CTA marches one ray from its root to termination -- the segment bound through
// grow_forward (v = 1 general box = the ray segment, src/queries.py:55-58), f(start) and f(start + eps) through the point rows of grow_forward
// (:67-70), the step update in the reference's operation order -- then takes the next ray from a global atomic queue.  No host
// round trip per iteration, no bucket padding; N_evals is replayed by the host from the per-ray step counts like the fixed modes.
#pragma once
#include "niq_grow.cuh"

namespace niq {

constexpr int kMaxRayFuncs = 4;

struct RayGrowArgs {
    CastOpts o;
    long long n;
    const float* roots; const float* dirs;
    float* out_t; int* out_hit; int* out_count; unsigned char* out_tie;
    unsigned long long* queue;
    int n_funcs;
    GrowCfg g[kMaxRayFuncs];
    int cg_lanes[kMaxRayFuncs];
    int W;                       // row width of the point buffers
    long long state_floats;      // size of the largest propagation state (the point buffers follow it)
};

__global__ void __launch_bounds__(256) k_cast_rays_grow(const __grid_constant__ NetDev net, const RayGrowArgs a) {
    extern __shared__ __align__(16) float sm[];
    __shared__ long long s_ray;
    __shared__ float s_pt[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* hA = sm + a.state_floats;
    float* hB = hA + 8 * a.W;
    const CastOpts& o = a.o;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_ray = (long long)atomicAdd(a.queue, 1ull);
        __syncthreads();
        const long long ray = s_ray;
        if (ray >= a.n) break;
        // ray state, identical in every thread
        const float rx = a.roots[3 * ray], ry = a.roots[3 * ray + 1], rz = a.roots[3 * ray + 2];
        const float dx = a.dirs[3 * ray], dy = a.dirs[3 * ray + 1], dz = a.dirs[3 * ray + 2];
        float t = 0.f, step = o.init_step;
        int count = 0, sub = 0, hit_id = 0;
        bool tie = false;
        for (;;) {
            // ---- one (sub)step: all funcs share t; can_step = AND, hit_id = last func whose signs differ (src/queries.py:52-77) ----
            bool can_step = true, is_hit = false;
            hit_id = 0;
            int l0 = 0;
            for (int f = 0; f < a.n_funcs; ++f) {
                int l1 = l0;
                while (!net.layers[l1].last_of_net) ++l1;
                ++l1;
                const GrowCfg& g = a.g[f];
                GrowState st;
                grow_carve(sm, g, st);
                // reference src/queries.py:55-58, 67-70
                const float psx = rx + t * dx, psy = ry + t * dy, psz = rz + t * dz;
                const float hs = 0.5f * step;
                const float hx = hs * dx, hy = hs * dy, hz = hs * dz;
                const float te = t + o.hit_eps;
                __syncthreads();
                if (tid == 0) {
                    st.base[0] = psx + hx; st.base[1] = psy + hy; st.base[2] = psz + hz; st.base[3] = 0.f;
                    st.aff[0] = hx; st.aff[1] = hy; st.aff[2] = hz; st.aff[3] = 0.f;
                    st.err[0] = st.err[1] = st.err[2] = st.err[3] = 0.f;
                }
                if (tid >= 32 && tid < 40) {
                    const int k = tid - 32;
                    float* d = hA + k * g.W;                  // row stride = this net's W (grow_forward indexes hA + warp * W)
                    d[0] = k == 0 ? psx : k == 1 ? rx + te * dx : 0.f;
                    d[1] = k == 0 ? psy : k == 1 ? ry + te * dy : 0.f;
                    d[2] = k == 0 ? psz : k == 1 ? rz + te * dz : 0.f;
                    d[3] = 0.f;
                }
                __syncthreads();
                float lo_b, up_b, sc, fv, fs;
                grow_forward(net, l0, l1, nullptr, nullptr, g, st, 1, lo_b, up_b, sc, hA, hB, a.cg_lanes[f], &fv, &fs);
                if (lane == 0 && warp < 2) { s_pt[warp] = fv; s_pt[2 + warp] = fs; }
                __syncthreads();
                const float v0 = s_pt[0], v1 = s_pt[1];
                const int lab = label_of(lo_b, up_b, 0.f);
                can_step = can_step && (lab == SIGN_POSITIVE || lab == SIGN_NEGATIVE);
                const int s0 = (v0 > 0.f) - (v0 < 0.f), s1 = (v1 > 0.f) - (v1 < 0.f);
                const bool this_hit = (s0 != s1) || (v0 != v0) || (v1 != v1);   // sign(nan)=nan != anything
                if (this_hit) hit_id = f + 1;
                is_hit = is_hit || this_hit;
                if (a.out_tie)
                    tie = tie || bound_near_tie(lo_b, up_b, 0.f, sc, net.tie_rel) || fabsf(v0) <= kNearTieRel * s_pt[2] ||
                          fabsf(v1) <= kNearTieRel * s_pt[3];
                l0 = l1;
            }
            // ---- step update + termination (reference src/queries.py:79-90, 113-126) ----
            count += 1;
            sub += 1;
            const float this_step = can_step ? step : o.hit_eps;
            if (!is_hit) t = t + this_step * o.safety;
            step = can_step ? step * o.grow : step * o.shrink;
            step = fmaxf(step, o.hit_eps);
            bool done = is_hit;
            if (!done && sub >= o.n_substeps) {
                sub = 0;
                done = (t > o.max_dist) || (count >= o.n_max_step);
            }
            if (done) break;
        }
        if (tid == 0) {
            a.out_t[ray] = t;
            a.out_hit[ray] = hit_id;
            a.out_count[ray] = count;
            if (a.out_tie) a.out_tie[ray] = tie ? 1 : 0;
        }
    }
}

}  // namespace niq
