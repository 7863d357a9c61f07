// niq_frustum_grow.cuh -- cast_rays_frustum for the growing-form modes (affine_truncate / affine_all / affine_append) as one
// persistent kernel (reference src/queries.py:178-587).  The protocol is that of k_cast_frustum (niq_kernels.cuh) with a CTA in
// the place of a warp slot: a CTA holds one frustum of pixels [x0,x1) x [y0,y1); the general-box bound (v = 3) goes through
// grow_forward, f(start) / f(start + eps) on the mid ray through its point rows; a split keeps child A and pushes child B on the
// device queue; an idle CTA takes a TICKET and adopts the record once it is published; the counter of outstanding frusta ends
// the kernel.  Every frustum carries its own iteration index k, and the per-iteration termination / split counts that the
// reference's N_evals depends on are histogrammed by k, exactly as in the fixed-row kernel (same FrustQueue, same host driver).
// Every thread of the CTA keeps the (identical) frustum state and geometry; thread 0 alone talks to the queue, and whatever it
// reads from global memory is broadcast through shared memory so that the CTA never diverges.
#pragma once
#include "niq_grow.cuh"

namespace niq {

struct FrustGrowArgs {
    CastOpts o;
    FrustQueue q;
    int n_funcs;
    GrowCfg g[4];
    int cg_lanes[4];
    int W;
    long long state_floats;
};

__global__ void __launch_bounds__(256) k_cast_frustum_grow(const __grid_constant__ NetDev net, const __grid_constant__ FrustCam cam,
                                                           const FrustGrowArgs a) {
    extern __shared__ __align__(16) float sm[];
    __shared__ long long s_ticket;
    __shared__ int s_flag;
    __shared__ FrustRec s_rec;
    __shared__ float s_pt[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* hA = sm + a.state_floats;
    float* hB = hA + 8 * a.W;
    const CastOpts& o = a.o;
    const FrustQueue& q = a.q;
    volatile unsigned long long* ctrl = q.ctrl;

    bool live = false, tie = false;
    long long ticket = -1;
    int x0 = 0, y0 = 0, x1 = 1, y1 = 1, k = 0;
    float t = 0.f, step = 0.f, count = 0.f;
    int sub = 0, n_inner = 0, hit_id = 0;
    bool is_hit = false, demands = false;

    for (;;) {
        // ---- an idle CTA takes a ticket, then adopts the record once it is published ----
        if (!live) {
            __syncthreads();
            if (tid == 0) {
                if (ticket < 0) ticket = (long long)atomicAdd(q.ctrl, 1ull);
                s_ticket = ticket;
                int flag = 0;                               // 0 wait, 1 adopted, 2 everything has finished
                if (ticket < q.cap && *((volatile int*)(q.ready + ticket)) != 0) {
                    __threadfence();
                    const int4 ra = __ldcg(reinterpret_cast<const int4*>(q.rec + ticket));
                    const float4 rb = __ldcg(reinterpret_cast<const float4*>(q.rec + ticket) + 1);
                    s_rec.x0 = ra.x; s_rec.y0 = ra.y; s_rec.x1 = ra.z; s_rec.y1 = ra.w;
                    s_rec.t = rb.x; s_rec.step = rb.y; s_rec.count = rb.z; s_rec.k_tie = __float_as_int(rb.w);
                    flag = 1;
                } else if (ctrl[2] == 0ull) {
                    flag = 2;
                }
                s_flag = flag;
            }
            __syncthreads();
            ticket = s_ticket;
            const int flag = s_flag;
            if (flag == 2) break;
            if (flag == 0) { __nanosleep(512); continue; }
            x0 = s_rec.x0; y0 = s_rec.y0; x1 = s_rec.x1; y1 = s_rec.y1;
            t = s_rec.t; step = s_rec.step; count = s_rec.count;
            k = s_rec.k_tie & 0x3fffffff; tie = (s_rec.k_tie >> 30) & 1;
            live = true; ticket = -1;
            sub = 0; n_inner = 0; hit_id = 0; is_hit = false; demands = false;
        }

        // ---- frustum geometry (src/queries.py:272-315), in the oracle's operation order ----
        float mid[3], rf[3], uf[3], t_adj;
        {
            const float dx = (float)cam.res_x + 1.f, dy = (float)cam.res_y + 1.f;
            const float xc_lo = (2.f * (float)x0) / dx - 1.f, xc_up = (2.f * (float)(x1 - 1)) / dx - 1.f;
            const float yc_lo = (2.f * (float)y0) / dy - 1.f, yc_up = (2.f * (float)(y1 - 1)) / dy - 1.f;
            float r_uu[3], r_lu[3], r_ul[3], r_ll[3];
            frustum_cam_ray(cam, xc_up, yc_up, r_uu);
            frustum_cam_ray(cam, xc_lo, yc_up, r_lu);
            frustum_cam_ray(cam, xc_up, yc_lo, r_ul);
            frustum_cam_ray(cam, xc_lo, yc_lo, r_ll);
#pragma unroll
            for (int d = 0; d < 3; ++d) mid[d] = 0.5f * (r_uu[d] + r_ll[d]);
            const float len = sqrtf((mid[0] * mid[0] + mid[1] * mid[1]) + mid[2] * mid[2]);
#pragma unroll
            for (int d = 0; d < 3; ++d) mid[d] = mid[d] / len;
            const float expand = 1.f / len;                 // the spherical cap reaches a little beyond the flat box
            t_adj = (t + step) * expand;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                rf[d] = ((r_uu[d] - r_lu[d]) * t_adj) / 2.f;
                uf[d] = ((r_uu[d] - r_ul[d]) * t_adj) / 2.f;
            }
        }

        // ---- one substep: all funcs share t; can_step = AND, hit_id = last func whose signs differ ----
        bool can_step = !is_hit;
        if (!is_hit) n_inner += 1;
        int l0 = 0;
        for (int f = 0; f < a.n_funcs; ++f) {
            int l1 = l0;
            while (!net.layers[l1].last_of_net) ++l1;
            ++l1;
            const GrowCfg& g = a.g[f];
            GrowState st;
            grow_carve(sm, g, st);
            const float cm = 0.5f * (t + t_adj), cv = 0.5f * (t_adj - t), te = t + o.hit_eps;
            __syncthreads();
            if (tid == 0) {
                const int W = g.W;
                st.base[0] = cam.root[0] + cm * mid[0]; st.base[1] = cam.root[1] + cm * mid[1]; st.base[2] = cam.root[2] + cm * mid[2]; st.base[3] = 0.f;
                st.aff[0] = cv * mid[0]; st.aff[1] = cv * mid[1]; st.aff[2] = cv * mid[2]; st.aff[3] = 0.f;
                st.aff[W] = rf[0]; st.aff[W + 1] = rf[1]; st.aff[W + 2] = rf[2]; st.aff[W + 3] = 0.f;
                st.aff[2 * W] = uf[0]; st.aff[2 * W + 1] = uf[1]; st.aff[2 * W + 2] = uf[2]; st.aff[2 * W + 3] = 0.f;
                st.err[0] = st.err[1] = st.err[2] = st.err[3] = 0.f;
            }
            if (tid >= 32 && tid < 40) {
                const int kk = tid - 32;
                const float tt = kk == 0 ? t : te;
                float* d = hA + kk * g.W;
                d[0] = kk < 2 ? cam.root[0] + tt * mid[0] : 0.f;
                d[1] = kk < 2 ? cam.root[1] + tt * mid[1] : 0.f;
                d[2] = kk < 2 ? cam.root[2] + tt * mid[2] : 0.f;
                d[3] = 0.f;
            }
            __syncthreads();
            float lo_b, up_b, sc, fv, fs;
            grow_forward(net, l0, l1, nullptr, nullptr, g, st, 3, lo_b, up_b, sc, hA, hB, a.cg_lanes[f], &fv, &fs);
            if (lane == 0 && warp < 2) { s_pt[warp] = fv; s_pt[2 + warp] = fs; }
            __syncthreads();
            const float v0 = s_pt[0], v1 = s_pt[1];
            const int lab = label_of(lo_b, up_b, 0.f);
            can_step = can_step && (lab == SIGN_POSITIVE || lab == SIGN_NEGATIVE);
            const int s0 = (v0 > 0.f) - (v0 < 0.f), s1 = (v1 > 0.f) - (v1 < 0.f);
            const bool this_hit = (s0 != s1) || (v0 != v0) || (v1 != v1);   // sign(nan)=nan != anything
            if (this_hit) hit_id = f + 1;
            is_hit = is_hit || this_hit;
            tie = tie || bound_near_tie(lo_b, up_b, 0.f, sc, net.tie_rel) || fabsf(v0) <= kNearTieRel * s_pt[2] ||
                  fabsf(v1) <= kNearTieRel * s_pt[3];
            l0 = l1;
        }

        // ---- step update (src/queries.py:249-266), then the end-of-iteration logic (:339-365) ----
        const bool single = (x0 + 1 == x1) && (y0 + 1 == y1);
        const float this_step = can_step ? step : (single ? o.hit_eps : 0.f);   // larger frusta may not inch forward
        if (!is_hit) t = t + this_step * o.safety;
        step = can_step ? step * o.grow : step * o.shrink;
        demands = demands || (step < o.hit_eps) || is_hit;
        step = fmaxf(step, o.hit_eps);
        sub += 1;
        if (is_hit) {
            // the remaining substeps re-evaluate the same points: t and hit_id stay, the step keeps shrinking
            for (; sub < o.n_substeps; ++sub) step = fmaxf(step * o.shrink, o.hit_eps);
        }
        if (sub >= o.n_substeps) {
            const int w = x1 - x0, h = y1 - y0, area = w * h;
            count = count + (float)n_inner * (1.0f / (float)area);
            const bool done = (is_hit && area == 1) || (t > o.max_dist) || (k * o.n_substeps >= o.n_max_step);
            const int kb = k < q.n_bins ? k : q.n_bins - 1;
            if (done) {
                if (tid == 0) {
                    const unsigned long long fi = atomicAdd(q.ctrl + 3, 1ull);
                    FrustFin fr;
                    fr.x0 = x0; fr.y0 = y0; fr.x1 = x1; fr.y1 = y1; fr.t = t; fr.hit_id = hit_id; fr.count = (int)count; fr.tie = tie ? 1 : 0;
                    q.fin[fi] = fr;
                    atomicAdd(q.hist_term + kb, 1u);
                    __threadfence();
                    atomicAdd(q.ctrl + 2, ~0ull);                       // outstanding -= 1
                }
                live = false;
            } else {
                const float wx = (2.f * sinf((cam.half_fov_x * (float)w) / (float)cam.res_x)) * t;
                const float wy = (2.f * sinf((cam.half_fov_y * (float)h) / (float)cam.res_y)) * t;
                const float lim = cam.refine_fac * step;
                const bool refine = (wx > lim || wy > lim || demands) && (w > 1 || h > 1);
                k += 1;
                if (refine) {
                    // split the longer pixel axis, x on ties (src/queries.py:371-432); B goes to the queue
                    int bx0 = x0, by0 = y0, bx1 = x1, by1 = y1;
                    if (w >= h) { const int xm = (x0 + x1) / 2; bx0 = xm; x1 = xm; }
                    else { const int ym = (y0 + y1) / 2; by0 = ym; y1 = ym; }
                    if (tid == 0) {
                        atomicAdd(q.hist_ref + kb, 1u);
                        const int k_tie = k | (tie ? (1 << 30) : 0);
                        atomicAdd(q.ctrl + 2, 1ull);                    // outstanding += 1, before the record can be adopted
                        const unsigned long long bi = atomicAdd(q.ctrl + 1, 1ull);
                        if ((long long)bi < q.cap) {
                            reinterpret_cast<int4*>(q.rec + bi)[0] = make_int4(bx0, by0, bx1, by1);
                            reinterpret_cast<float4*>(q.rec + bi)[1] = make_float4(t, step, count, __int_as_float(k_tie));
                            __threadfence();
                            *((volatile int*)(q.ready + bi)) = 1;
                        } else {
                            atomicAdd(q.ctrl + 4, 1ull);                // cannot happen: at most one frustum per pixel
                            atomicAdd(q.ctrl + 2, ~0ull);
                        }
                    }
                }
                sub = 0; n_inner = 0; hit_id = 0; is_hit = false; demands = false;
            }
        }
    }
}

}  // namespace niq
