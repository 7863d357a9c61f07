// niq_internal.h -- host-side plumbing shared by the translation units of libniq.so (contexts, error codes, stream-ordered
// temporaries, launch timers, the packed MLP handle, and the launchers each kernel family's TU exports).  Not part of the ABI.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/niq.h"
#include "niq_kernels.cuh"

using namespace niq;

int niq_fail(int code, const char* fmt, ...);
#define fail niq_fail

#define CU(expr)                                                                                    \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail(e_ == cudaErrorMemoryAllocation ? NIQ_ENOMEM : NIQ_ECUDA, "%s failed: %s (%s:%d)", #expr, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                                \
    } while (0)
#define TRY(expr)            \
    do {                     \
        int r_ = (expr);     \
        if (r_ != NIQ_OK) return r_; \
    } while (0)

struct TimedLaunch { cudaEvent_t a, b; int family; };

struct niq_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;          // side stream: the second shape of find_any_intersection runs beside the first
    cudaEvent_t fork = nullptr, join = nullptr;
    cudaDeviceProp prop{};
    long long launches = 0;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    bool timing = false;
    std::vector<TimedLaunch> pending;
    std::vector<cudaEvent_t> event_pool;
    double fam_ms[2] = {0, 0};
    long long fam_launches[2] = {0, 0};
    long long* pinned = nullptr;     // small pinned read-back area (64 x int64)
    bool timer_armed = false, timer_started = false;   // niq_ctx_timer_start .. _stop bracket (see timer_touch / timer_mark)
    unsigned long long* d_exec = nullptr;   // executed-MAC counter of the engine kernels (zero-skipping accounting)
    bool count_exec = false;
    long long mc_points_evaluated = 0, mc_points_lattice = 0;   // marching cubes: lattice points evaluated / the reference's count
};

struct DevBuf {   // stream-ordered temporary
    niq_ctx* ctx; void* p = nullptr;
    explicit DevBuf(niq_ctx* c) : ctx(c) {}
    int alloc(size_t bytes) {
        if (bytes == 0) bytes = 16;
        cudaError_t e = cudaMallocAsync(&p, bytes, ctx->stream);
        if (e != cudaSuccess) { p = nullptr; return fail(NIQ_ENOMEM, "cudaMallocAsync(%zu) failed: %s", bytes, cudaGetErrorString(e)); }
        return NIQ_OK;
    }
    ~DevBuf() { if (p) cudaFreeAsync(p, ctx->stream); }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
};

cudaEvent_t niq_get_event(niq_ctx* c);
#define get_event niq_get_event
struct LaunchTimer {   // brackets one kernel launch with events when timing is on
    niq_ctx* c; int fam; cudaEvent_t a = nullptr, b = nullptr;
    LaunchTimer(niq_ctx* ctx, int family) : c(ctx), fam(family) {
        c->launches++;
        if (c->timing && c->pending.size() < 8192) { a = get_event(c); b = get_event(c); cudaEventRecord(a, c->stream); }
    }
    ~LaunchTimer() {
        if (a) { cudaEventRecord(b, c->stream); c->pending.push_back({a, b, fam}); }
    }
};
void niq_resolve_timers(niq_ctx* c);

struct HostLayer { int in_dim, out_dim, in_pad, out_pad, act; bool dot; size_t w_off, b_off; };

struct niq_mlp {
    niq_ctx* ctx = nullptr;
    std::vector<HostLayer> layers;
    float* d_weights = nullptr;
    float* d_bias = nullptr;
    NetDev net{};
    int wmax = 32;         // width class of the fixed-row engine
    int maxw_pad = 8;      // widest padded row (grow engine)
    int64_t macs = 0;
    int total_floats = 0;  // packed weights of all layers
    int sum_act_out = 0;   // sum of out_dim over activation layers (affine_all growth)
    int max_act_out = 0;
    int min_act_out = 1 << 30, n_act_layers = 0;
};

static int round_up(int x, int m) { return (x + m - 1) / m * m; }
constexpr int kResidentPad = 512;   // floats after the resident weights: the pipelined loop over-reads one weight row

// Decide where the weights of a launch live: resident in shared memory when everything fits beside the
// activation buffers, otherwise streamed through the ring.  Returns the dynamic shared-memory size.
template <class E>
static size_t place_weights(niq_ctx* c, NetDev& net, int total_floats) {
    net.exec_macs = c->count_exec ? c->d_exec : nullptr;
    {   // development knob: head start (cycles) of warps 0-3 over warps 4-7 (streamed ray kernels; resident nets: overrides the
        // default of half a layer, niq_engine.cuh)
        const char* e = getenv("NIQ_DEPHASE");
        net.dephase = e ? atoi(e) : 0;
    }
    const size_t res = E::smem_bytes(total_floats + kResidentPad);
    if (res <= c->prop.sharedMemPerBlockOptin) {
        net.resident = 1;
        net.w_region_floats = total_floats + kResidentPad;
        return res;
    }
    net.resident = 0;
    return E::smem_bytes();
}

template <class K>
static int set_smem(K kernel, size_t bytes) {
    CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return NIQ_OK;
}
static int grid_for(niq_ctx* c, long long n_pass) {
    return (int)std::max<long long>(1, std::min<long long>(n_pass, c->prop.multiProcessorCount));
}

struct NodeList { float* lo = nullptr; float* hi = nullptr; long long n = 0, cap = 0; };

struct niq_tree {
    niq_ctx* ctx = nullptr;
    NodeList lists[3];     // 0 unknown leaves, 1 interior, 2 exterior
    long long stats[4] = {0, 0, 0, 0};
    std::vector<long long> levels;   // 4 per level: nodes entering, unknown, negative, positive
};


static void timer_touch(niq_ctx* c) {
    if (c->timer_armed && !c->timer_started) { cudaEventRecord(c->t0, c->stream); c->timer_started = true; }
}
static void timer_mark(niq_ctx* c) {
    if (c->timer_armed && c->timer_started) cudaEventRecord(c->t1, c->stream);
}
#define FINAL_SYNC(c) do { timer_mark(c); CU(cudaStreamSynchronize((c)->stream)); } while (0)

static bool invert3(const float* R, float* inv) {   // float32 Gauss-Jordan with partial pivoting
    float a[3][6];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { a[i][j] = R[3 * i + j]; a[i][3 + j] = i == j ? 1.f : 0.f; }
    for (int col = 0; col < 3; ++col) {
        int piv = col;
        for (int r = col + 1; r < 3; ++r) if (std::fabs(a[r][col]) > std::fabs(a[piv][col])) piv = r;
        if (a[piv][col] == 0.f) return false;
        if (piv != col) for (int j = 0; j < 6; ++j) std::swap(a[piv][j], a[col][j]);
        const float d = a[col][col];
        for (int j = 0; j < 6; ++j) a[col][j] = a[col][j] / d;
        for (int r = 0; r < 3; ++r) {
            if (r == col) continue;
            const float f = a[r][col];
            for (int j = 0; j < 6; ++j) a[r][j] = a[r][j] - f * a[col][j];
        }
    }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) inv[3 * i + j] = a[i][3 + j];
    return true;
}


// launchers, one translation unit per kernel family (compiled in parallel; see __graft_entry__.build)
int launch_classify_fixed(niq_ctx* c, const niq_mlp* m, const BoxSource& src, long long n, float offset,
                          int* label, float* lower, float* upper, unsigned char* tie);
int launch_classify_slope(niq_ctx* c, const niq_mlp* m, const BoxSource& src, long long n, float offset,
                          int* label, float* lower, float* upper, unsigned char* tie, float* raw = nullptr, float* raw_scale = nullptr);
int launch_eval_points(niq_ctx* c, const niq_mlp* m, const PointSource& src, long long n, float* f, float* scale);
int launch_cast_rays(niq_ctx* c, int wmax, const NetDev& net, int total_floats, const CastOpts& o, long long n, int interval,
                     const float* roots, const float* dirs, float* t, int* hit, int* cnt, unsigned char* tie,
                     unsigned long long* queue, bool slope);
int launch_cast_frustum(niq_ctx* c, int wmax, const NetDev& net, int total_floats, const CastOpts& o, const FrustCam& cam,
                        int interval, const FrustQueue& q, long long n_pixels, bool slope);
int launch_classify_grow(niq_ctx* c, const niq_mlp* m, const niq_mode_cfg* cfg, BoxSource src, long long n,
                         float offset, int* label, float* lower, float* upper, unsigned char* tie);
// the whole level-set tree in one cooperative launch (niq_tree.cuh); *handled = false when the mode has no persistent kernel
// deal_world > 1: the frontier entering level deal_level is dealt round-robin, this rank keeps i = deal_rank (mod deal_world)
int tree_build_persistent(niq_ctx* c, const niq_mlp* m, const niq_mode_cfg* cfg, long long n_roots, const float* lower,
                          const float* upper, int split_depth, long long node_thresh, float offset, int flags, int bps,
                          niq_tree* T, bool* handled, int deal_level = -1, int deal_rank = 0, int deal_world = 1);
// find_any_intersection for the growing-form modes as one persistent cooperative kernel (niq_isect.cuh), for a batch of n_q
// queries that differ in the rigid transforms prepended to the two shapes (xfA / xfB: HOST (n_q, 12) = R (3x3) + t (3) per
// query, or NULL: the handle's own first layer).  *handled = false when a mode is not a growing-form mode.
int isect_grow_batch(niq_ctx* c, const niq_mlp* mA, const niq_mode_cfg* cfgA, const niq_mlp* mB, const niq_mode_cfg* cfgB,
                     long long n_q, const float* xfA, const float* xfB, const float lower[3], const float upper[3], float eps,
                     int32_t* found, float* loc, int64_t* stats, bool* handled);
int launch_cast_rays_grow(niq_ctx* c, int n_funcs, const niq_mlp* const* mlps, const niq_mode_cfg* cfgs, const NetDev& net,
                          const CastOpts& o, long long n, const float* roots, const float* dirs, float* t, int* hit, int* cnt,
                          unsigned char* tie, unsigned long long* queue);
int launch_cast_frustum_grow(niq_ctx* c, int n_funcs, const niq_mlp* const* mlps, const niq_mode_cfg* cfgs, const NetDev& net,
                             const CastOpts& o, const FrustCam& cam, const FrustQueue& q, long long n_pixels);
namespace niq { struct CpArgs; }
int launch_cp_persistent(niq_ctx* c, const niq_mlp* m, const niq::CpArgs& a, bool* fits);
