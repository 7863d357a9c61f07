// niq_api.cu -- the C ABI (include/niq.h): contexts, MLP packing, and the host-side drivers of the queries.
// Build: see __graft_entry__.build (one translation unit per kernel family, linked into libniq.so).
#define NIQ_HELPER_KERNELS
#include <cstring>

#include "niq_internal.h"
#include "niq_cp.cuh"

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

int niq_fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

extern "C" const char* niq_last_error(void) { return g_last_error.c_str(); }
extern "C" const char* niq_version(void) { return "niq-b200 0.1 (sm_100a)"; }

// 128-bit fingerprint of a byte range (host code, no GPU): four interleaved lanes of xor / rotate / odd-multiply over
// 8-byte words (every step is a bijection of the lane state, so a changed word always changes its lane), mixed down with
// splitmix64 finalisers.  The Python layer keys its MLP-handle cache with it: the reference re-reads `params` on every
// call, so the key must follow the CONTENT, and it is recomputed per query call -- one memory-bound pass over the weights.
extern "C" int niq_fingerprint128(const void* data, int64_t nbytes, uint64_t out[2]) {
    if ((!data && nbytes > 0) || nbytes < 0 || !out) return fail(NIQ_EINVAL, "niq_fingerprint128: bad argument");
    static const uint64_t K[4] = {0x9E3779B97F4A7C15ull, 0xC2B2AE3D27D4EB4Full, 0x165667B19E3779F9ull, 0xD6E8FEB86659FD93ull};
    uint64_t s[4] = {0x243F6A8885A308D3ull ^ (uint64_t)nbytes, 0x13198A2E03707344ull, 0xA4093822299F31D0ull, 0x082EFA98EC4E6C89ull};
    const unsigned char* p = static_cast<const unsigned char*>(data);
    int64_t i = 0;
    for (; i + 32 <= nbytes; i += 32) {
        uint64_t x[4];
        memcpy(x, p + i, 32);
        for (int l = 0; l < 4; ++l) { const uint64_t v = s[l] ^ x[l]; s[l] = ((v << 29) | (v >> 35)) * K[l]; }
    }
    if (i < nbytes) {                                   // tail: zero-padded block (the length is part of the seed)
        uint64_t x[4] = {0, 0, 0, 0};
        memcpy(x, p + i, (size_t)(nbytes - i));
        for (int l = 0; l < 4; ++l) { const uint64_t v = s[l] ^ x[l]; s[l] = ((v << 29) | (v >> 35)) * K[l]; }
    }
    auto mix = [](uint64_t z) { z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 27; z *= 0x94D049BB133111EBull; return z ^ (z >> 31); };
    out[0] = mix(s[0] + mix(s[1] + mix(s[2] + mix(s[3]))));
    out[1] = mix(s[3] ^ mix(s[2] ^ mix(s[1] ^ mix(s[0] + 0x9E3779B97F4A7C15ull))));
    return NIQ_OK;
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
cudaEvent_t niq_get_event(niq_ctx* c) {
    if (!c->event_pool.empty()) { cudaEvent_t e = c->event_pool.back(); c->event_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
void niq_resolve_timers(niq_ctx* c) {
    for (auto& t : c->pending) {
        float ms = 0.f;
        if (cudaEventSynchronize(t.b) == cudaSuccess && cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) {
            c->fam_ms[t.family] += ms;
            c->fam_launches[t.family] += 1;
        }
        c->event_pool.push_back(t.a);
        c->event_pool.push_back(t.b);
    }
    c->pending.clear();
}
#define resolve_timers niq_resolve_timers


extern "C" int niq_ctx_create(int device, niq_ctx** out) {
    if (!out) return fail(NIQ_EINVAL, "niq_ctx_create: out is NULL");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(NIQ_ECUDA, "no CUDA device available (%s): this backend has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(NIQ_EINVAL, "device %d out of range (have %d)", device, count);
    CU(cudaSetDevice(device));
    niq_ctx* c = new niq_ctx();
    struct CtxGuard { niq_ctx* c; bool ok = false; ~CtxGuard() { if (!ok) niq_ctx_destroy(c); } } guard{c};   // no leak on any error path
    c->device = device;
    CU(cudaGetDeviceProperties(&c->prop, device));
    if (c->prop.major != 10)      // the library holds sm_100a code only: any other architecture would fail at the first launch
        return fail(NIQ_ECUDA, "device compute capability %d.%d: kernels are built for sm_100a (B200) only", c->prop.major, c->prop.minor);
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    {   // stream-ordered temporaries (DevBuf) stay in the pool across synchronisations instead of going back to the OS
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    CU(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->fork, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->join, cudaEventDisableTiming));
    CU(cudaEventCreate(&c->t0));
    CU(cudaEventCreate(&c->t1));
    CU(cudaMallocHost(&c->pinned, 64 * sizeof(long long)));
    CU(cudaMalloc(&c->d_exec, 8));
    CU(cudaMemset(c->d_exec, 0, 8));
    guard.ok = true;
    *out = c;
    return NIQ_OK;
}
extern "C" int niq_ctx_destroy(niq_ctx* c) {
    if (!c) return NIQ_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    resolve_timers(c);
    for (auto e : c->event_pool) cudaEventDestroy(e);
    if (c->t0) cudaEventDestroy(c->t0);
    if (c->t1) cudaEventDestroy(c->t1);
    if (c->pinned) cudaFreeHost(c->pinned);
    if (c->d_exec) cudaFree(c->d_exec);
    if (c->fork) cudaEventDestroy(c->fork);
    if (c->join) cudaEventDestroy(c->join);
    if (c->stream2) cudaStreamDestroy(c->stream2);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return NIQ_OK;
}
extern "C" int niq_ctx_sync(niq_ctx* c) {
    if (!c) return fail(NIQ_EINVAL, "ctx is NULL");
    CU(cudaStreamSynchronize(c->stream));
    return NIQ_OK;
}
extern "C" int niq_ctx_device_info(niq_ctx* c, int32_t info[4]) {
    if (!c || !info) return fail(NIQ_EINVAL, "bad argument");
    info[0] = c->prop.multiProcessorCount; info[1] = c->prop.major; info[2] = c->prop.minor;
    info[3] = (int32_t)c->prop.sharedMemPerBlockOptin;
    return NIQ_OK;
}
extern "C" int niq_ctx_launch_count(niq_ctx* c, int64_t* out) {
    if (!c || !out) return fail(NIQ_EINVAL, "bad argument");
    *out = c->launches;
    return NIQ_OK;
}
// Device-side timing of a bracket of API calls: t0 is recorded on the stream when the first call after
// niq_ctx_timer_start enqueues its work, t1 right after the last call has enqueued its last operation (BEFORE the
// host waits for it), so the reading is the device time of the calls and does not include the wake-up latency of a
// descheduled host thread (measured on the shared GPU box: up to a second of jitter on a 0.6 s step).

extern "C" int niq_ctx_mc_points(niq_ctx* c, int64_t* evaluated, int64_t* lattice, int reset) {
    if (!c || !evaluated || !lattice) return fail(NIQ_EINVAL, "bad argument");
    *evaluated = c->mc_points_evaluated; *lattice = c->mc_points_lattice;
    if (reset) c->mc_points_evaluated = c->mc_points_lattice = 0;
    return NIQ_OK;
}
extern "C" int niq_ctx_timer_start(niq_ctx* c) {
    if (!c) return fail(NIQ_EINVAL, "ctx is NULL");
    c->timer_armed = true;
    c->timer_started = false;
    return NIQ_OK;
}
extern "C" int niq_ctx_timer_stop(niq_ctx* c, float* ms) {
    if (!c || !ms) return fail(NIQ_EINVAL, "bad argument");
    *ms = 0.f;
    const bool started = c->timer_armed && c->timer_started;
    c->timer_armed = false;
    c->timer_started = false;
    if (started) {
        CU(cudaEventSynchronize(c->t1));
        CU(cudaEventElapsedTime(ms, c->t0, c->t1));
    }
    return NIQ_OK;
}
extern "C" int niq_ctx_kernel_timing(niq_ctx* c, int on) {
    if (!c) return fail(NIQ_EINVAL, "ctx is NULL");
    c->timing = on != 0;
    return NIQ_OK;
}
extern "C" int niq_ctx_kernel_ms(niq_ctx* c, int which, float* ms, int64_t* launches, int reset) {
    if (!c || which < 0 || which > 1) return fail(NIQ_EINVAL, "bad argument");
    CU(cudaStreamSynchronize(c->stream));
    resolve_timers(c);
    if (ms) *ms = (float)c->fam_ms[which];
    if (launches) *launches = c->fam_launches[which];
    if (reset) { c->fam_ms[which] = 0; c->fam_launches[which] = 0; }
    return NIQ_OK;
}
extern "C" int niq_ctx_exec_macs(niq_ctx* c, int on, int64_t* macs, int reset) {
    if (!c) return fail(NIQ_EINVAL, "ctx is NULL");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    if (macs) {
        unsigned long long v = 0;
        CU(cudaMemcpy(&v, c->d_exec, 8, cudaMemcpyDeviceToHost));
        *macs = (int64_t)v;
    }
    if (reset) CU(cudaMemset(c->d_exec, 0, 8));
    c->count_exec = on != 0;
    return NIQ_OK;
}
extern "C" int niq_dev_alloc(niq_ctx* c, int64_t bytes, void** out) {
    if (!c || !out || bytes < 0) return fail(NIQ_EINVAL, "bad argument");
    CU(cudaSetDevice(c->device));
    CU(cudaMalloc(out, (size_t)std::max<int64_t>(bytes, 16)));
    return NIQ_OK;
}
extern "C" int niq_dev_free(niq_ctx* c, void* p) {
    if (!c) return fail(NIQ_EINVAL, "ctx is NULL");
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaFree(p));
    return NIQ_OK;
}
extern "C" int niq_dev_upload(niq_ctx* c, void* dst, const void* src, int64_t bytes) {
    if (!c) return fail(NIQ_EINVAL, "ctx is NULL");
    CU(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return NIQ_OK;
}
extern "C" int niq_dev_download(niq_ctx* c, void* dst, const void* src, int64_t bytes) {
    if (!c) return fail(NIQ_EINVAL, "ctx is NULL");
    CU(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return NIQ_OK;
}
extern "C" int niq_measure_fp32_peak(niq_ctx* c, float* tflops) {
    if (!c || !tflops) return fail(NIQ_EINVAL, "bad argument");
    CU(cudaSetDevice(c->device));
    DevBuf out(c);
    TRY(out.alloc(16));
    const int iters = 4096, blocks = c->prop.multiProcessorCount * 8;
    float best = 0.f;
    for (int rep = 0; rep < 5; ++rep) {
        CU(cudaEventRecord(c->t0, c->stream));
        k_ffma_peak<<<blocks, 256, 0, c->stream>>>(out.as<float>(), iters, 0.999f, 0.001f);
        c->launches++;
        CU(cudaEventRecord(c->t1, c->stream));
        CU(cudaEventSynchronize(c->t1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, c->t0, c->t1));
        const double flop = 2.0 * 16 * 8 * (double)iters * 256.0 * blocks;
        if (rep > 0) best = std::max(best, (float)(flop / (ms * 1e-3) / 1e12));
    }
    CU(cudaGetLastError());
    *tflops = best;
    return NIQ_OK;
}

// development probe (tools/): FFMA rate at a given occupancy -- blocks per SM x threads per block
extern "C" int niq_probe_ffma(niq_ctx* c, int blocks_per_sm, int threads, float* tflops) {
    if (!c || !tflops || threads <= 0 || threads > 256) return fail(NIQ_EINVAL, "bad argument");
    CU(cudaSetDevice(c->device));
    DevBuf out(c);
    TRY(out.alloc(16));
    const int iters = 8192, blocks = c->prop.multiProcessorCount * blocks_per_sm;
    float best = 0.f;
    for (int rep = 0; rep < 4; ++rep) {
        CU(cudaEventRecord(c->t0, c->stream));
        k_ffma_peak<<<blocks, threads, 0, c->stream>>>(out.as<float>(), iters, 0.999f, 0.001f);
        CU(cudaEventRecord(c->t1, c->stream));
        CU(cudaEventSynchronize(c->t1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, c->t0, c->t1));
        const double flop = 2.0 * 16 * 8 * (double)iters * threads * (double)blocks;
        if (rep > 0) best = std::max(best, (float)(flop / (ms * 1e-3) / 1e12));
    }
    *tflops = best;
    return NIQ_OK;
}

// ------------------------------------------------------------------------------------------------
// MLP packing
// ------------------------------------------------------------------------------------------------
extern "C" int niq_mlp_create(niq_ctx* c, int32_t n_ops, const niq_op_desc* ops, niq_mlp** out) {
    if (!c || !ops || !out || n_ops <= 0) return fail(NIQ_EINVAL, "niq_mlp_create: bad argument");
    CU(cudaSetDevice(c->device));
    struct Raw { int in, out, act; std::vector<float> A, b; };
    std::vector<Raw> raw;
    bool squeezed = false;
    for (int i = 0; i < n_ops; ++i) {
        const niq_op_desc& op = ops[i];
        if (squeezed) return fail(NIQ_EINVAL, "op %d follows squeeze_last (squeeze_last must be the final op)", i);
        switch (op.kind) {
            case NIQ_OP_DENSE: {
                if (op.in_dim <= 0 || op.out_dim <= 0 || !op.A) return fail(NIQ_EINVAL, "dense op %d: bad shape / NULL A", i);
                Raw r; r.in = op.in_dim; r.out = op.out_dim; r.act = ACT_NONE;
                r.A.assign(op.A, op.A + (size_t)op.in_dim * op.out_dim);
                if (op.b) r.b.assign(op.b, op.b + op.out_dim); else r.b.assign(op.out_dim, 0.f);
                raw.push_back(std::move(r));
                break;
            }
            case NIQ_OP_SPATIAL: {
                if (!op.A || !op.b) return fail(NIQ_EINVAL, "spatial_transformation op %d: R and t are required", i);
                Raw r; r.in = 3; r.out = 3; r.act = ACT_NONE;
                r.A.resize(9); r.b.resize(3);
                if (!invert3(op.A, r.A.data())) return fail(NIQ_EINVAL, "spatial_transformation op %d: R is singular", i);
                // reference src/affine_layers.py:175-179: dense(x, A=inv(R), b=inv(R)@(-t)), used as x@A
                for (int k = 0; k < 3; ++k) {
                    float s = 0.f;
                    for (int j = 0; j < 3; ++j) s = s + r.A[3 * k + j] * (-op.b[j]);
                    r.b[k] = s;
                }
                raw.push_back(std::move(r));
                break;
            }
            case NIQ_OP_POW2_ENCODE: {
                // x (3) -> (x[:,None] * coefs[None,:] + shift).flatten(): a dense layer with one non-zero per column,
                // A[d][d*c + i] = coefs[i], b = tile(shift).  fma(x_d, coef, 0) rounds once like the reference's product,
                // the bias add follows it, and err * coefs == err @ |A| because the coefficients are positive.
                if (!op.A || op.out_dim <= 0 || op.in_dim != 3) return fail(NIQ_EINVAL, "pow2_frequency_encode op %d: needs 3 inputs and coefs", i);
                if (!raw.empty() && (raw.back().out != 3 || raw.back().act != ACT_NONE))
                    return fail(NIQ_EUNSUPPORTED, "pow2_frequency_encode op %d: only on the 3-D input (optionally after spatial_transformation)", i);
                const int nc = op.out_dim;
                Raw r; r.in = 3; r.out = 3 * nc; r.act = ACT_NONE;
                r.A.assign((size_t)3 * r.out, 0.f); r.b.assign(r.out, 0.f);
                for (int d = 0; d < 3; ++d)
                    for (int k = 0; k < nc; ++k) {
                        if (!(op.A[k] > 0.f)) return fail(NIQ_EUNSUPPORTED, "pow2_frequency_encode op %d: coefs must be positive", i);
                        r.A[(size_t)d * r.out + d * nc + k] = op.A[k];
                        if (op.b) r.b[d * nc + k] = op.b[k];
                    }
                raw.push_back(std::move(r));
                break;
            }
            case NIQ_OP_RELU:
            case NIQ_OP_ELU:
            case NIQ_OP_SIN:
            case NIQ_OP_TANH:
                if (raw.empty() || raw.back().act != ACT_NONE)
                    return fail(NIQ_EUNSUPPORTED, "op %d: an activation must directly follow a dense / spatial / encode op", i);
                raw.back().act = op.kind == NIQ_OP_RELU ? ACT_RELU : op.kind == NIQ_OP_ELU ? ACT_ELU : op.kind == NIQ_OP_TANH ? ACT_TANH : ACT_SIN;
                break;
            case NIQ_OP_SQUEEZE_LAST:
                if (raw.empty() || raw.back().out != 1) return fail(NIQ_EINVAL, "squeeze_last needs a preceding op with out_dim 1");
                squeezed = true;
                break;
            default:
                return fail(NIQ_EUNSUPPORTED, "op %d: kind %d is not an op of the reference's mlp format", i, op.kind);
        }
    }
    if (raw.empty()) return fail(NIQ_EINVAL, "no dense layer");
    if (raw.front().in != 3) return fail(NIQ_EUNSUPPORTED, "input dimension %d: queries are 3-D", raw.front().in);
    for (const Raw& r : raw)
        if (r.act == ACT_SIN && &r == &raw.back()) return fail(NIQ_EUNSUPPORTED, "activation after the final layer is not supported");
    if (raw.back().out != 1) return fail(NIQ_EUNSUPPORTED, "the last dense layer must have out_dim 1 (scalar implicit function)");
    if (raw.back().act != ACT_NONE) return fail(NIQ_EUNSUPPORTED, "activation after the final layer is not supported");
    if ((int)raw.size() > kMaxLayers) return fail(NIQ_EUNSUPPORTED, "more than %d layers", kMaxLayers);
    for (size_t l = 1; l < raw.size(); ++l)
        if (raw[l].in != raw[l - 1].out) return fail(NIQ_EINVAL, "layer %zu: in_dim %d != previous out_dim %d", l, raw[l].in, raw[l - 1].out);

    niq_mlp* m = new niq_mlp();
    m->ctx = c;
    struct MlpGuard { niq_mlp* m; bool ok = false; ~MlpGuard() { if (!ok) niq_mlp_destroy(m); } } guard{m};   // no leak on any error path
    int maxw = 8;
    for (size_t l = 0; l + 1 < raw.size(); ++l) maxw = std::max(maxw, raw[l].out);
    if (maxw > 256) return fail(NIQ_EUNSUPPORTED, "hidden width %d > 256", maxw);
    m->wmax = maxw <= 32 ? 32 : maxw <= 64 ? 64 : maxw <= 128 ? 128 : 256;

    std::vector<float> hw, hb;
    int n_chunks = 0;
    for (size_t l = 0; l < raw.size(); ++l) {
        HostLayer L{};
        L.in_dim = raw[l].in; L.out_dim = raw[l].out; L.act = raw[l].act;
        L.dot = (l + 1 == raw.size());
        L.in_pad = l == 0 ? 4 : m->layers[l - 1].out_pad;
        L.out_pad = L.dot ? 1 : round_up(L.out_dim, 8);
        if (!L.dot && raw[l].out == 1) return fail(NIQ_EUNSUPPORTED, "hidden layer of width 1");
        L.w_off = hw.size();
        hw.resize(hw.size() + (size_t)L.in_pad * (L.dot ? 1 : L.out_pad), 0.f);
        for (int j = 0; j < L.in_dim; ++j)
            for (int k = 0; k < L.out_dim; ++k)
                hw[L.w_off + (size_t)j * (L.dot ? 1 : L.out_pad) + k] = raw[l].A[(size_t)j * L.out_dim + k];
        hw.resize(round_up((int)hw.size(), 4), 0.f);
        L.b_off = hb.size();
        hb.resize(hb.size() + (L.dot ? 4 : L.out_pad), 0.f);
        for (int k = 0; k < L.out_dim; ++k) hb[L.b_off + k] = raw[l].b[k];
        m->layers.push_back(L);
        m->macs += (int64_t)L.in_dim * L.out_dim;
        m->maxw_pad = std::max(m->maxw_pad, std::max(L.in_pad, L.dot ? 1 : L.out_pad));
        if (L.act != ACT_NONE) {
            m->sum_act_out += L.out_dim; m->max_act_out = std::max(m->max_act_out, L.out_dim);
            m->min_act_out = std::min(m->min_act_out, L.out_dim); m->n_act_layers += 1;
        }
    }
    // from the stream-ordered pool (kept across synchronisations, niq_ctx_create): a query that edits a transform makes a new handle
    // per call, and cudaMalloc / cudaFree are device-wide synchronisation points with occasional long stalls
    CU(cudaMallocAsync(reinterpret_cast<void**>(&m->d_weights), hw.size() * sizeof(float), c->stream));
    CU(cudaMallocAsync(reinterpret_cast<void**>(&m->d_bias), hb.size() * sizeof(float), c->stream));
    CU(cudaMemcpyAsync(m->d_weights, hw.data(), hw.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(m->d_bias, hb.data(), hb.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));

    NetDev& nd = m->net;
    nd.n_layers = (int)m->layers.size();
    nd.n_nets = 1;
    nd.tie_rel = 1e-5f;
    for (const HostLayer& L : m->layers) if (L.act == ACT_SIN) nd.tie_rel = 5e-5f;   // sin rule: float32 conditioning ~1.6e-5 (tests)
    for (const HostLayer& L : m->layers) if (L.act == ACT_TANH) nd.tie_rel = 1e-4f;  // ours (unpinned rule): secant slope / atanh amplify 1-ulp tanhf differences
    for (const HostLayer& L : m->layers) if (L.act == ACT_ELU) nd.tie_rel = 2e-4f;   // ELU rule conditioning
    for (size_t l = 0; l < m->layers.size(); ++l) {
        const HostLayer& L = m->layers[l];
        LayerDev& D = nd.layers[l];
        D.in_dim = L.in_dim; D.out_dim = L.out_dim; D.in_pad = L.in_pad; D.out_pad = L.out_pad; D.act = L.act;
        D.bias = m->d_bias + L.b_off;
        D.first_of_net = l == 0; D.last_of_net = L.dot;
        D.chunk_begin = n_chunks;
        const int row = L.dot ? 1 : L.out_pad;
        int kc_max = L.dot ? L.in_pad : std::max(8, (kChunkFloats / row) / 8 * 8);   // multiple of 8: pipelined main loop
        for (int k0 = 0; k0 < L.in_pad; k0 += kc_max) {
            if (n_chunks >= kMaxChunks) return fail(NIQ_EUNSUPPORTED, "weight stream needs more than %d chunks", kMaxChunks);
            ChunkDev& C = nd.chunks[n_chunks++];
            C.k0 = k0; C.kc = std::min(kc_max, L.in_pad - k0);
            C.src = m->d_weights + L.w_off + (size_t)k0 * row;
            C.n_floats = (unsigned)(C.kc * row);
            C.smem_off = (int)(L.w_off + (size_t)k0 * row);
        }
        D.chunk_end = n_chunks;
    }
    nd.n_chunks = n_chunks;
    // zero-skipping after relu layers (niq_engine.cuh write_back_sparse); NIQ_NO_SPARSE=1 forces the dense K loops (A/B tests)
    nd.sparse = getenv("NIQ_NO_SPARSE") == nullptr || getenv("NIQ_NO_SPARSE")[0] == '0';
    for (int l = 0; l < nd.n_layers; ++l)
        if (nd.layers[l].chunk_end - nd.layers[l].chunk_begin > kMaxSegs) nd.sparse = 0;
    nd.exec_macs = nullptr;
    nd.resident = 0;
    nd.w_region_floats = (int)hw.size() + kResidentPad;
    m->total_floats = (int)hw.size();
    guard.ok = true;
    *out = m;
    return NIQ_OK;
}
extern "C" int niq_mlp_destroy(niq_mlp* m) {
    if (!m) return NIQ_OK;
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    if (m->d_weights) cudaFreeAsync(m->d_weights, m->ctx->stream);
    if (m->d_bias) cudaFreeAsync(m->d_bias, m->ctx->stream);
    delete m;
    return NIQ_OK;
}
extern "C" int niq_mlp_tie_rel(const niq_mlp* m, float* rel) {
    if (!m || !rel) return fail(NIQ_EINVAL, "bad argument");
    *rel = m->net.tie_rel;
    return NIQ_OK;
}
extern "C" int niq_mlp_macs(const niq_mlp* m, int64_t* macs) {
    if (!m || !macs) return fail(NIQ_EINVAL, "bad argument");
    *macs = m->macs;
    return NIQ_OK;
}

// ------------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------------
static int check_cfg(const niq_mode_cfg* cfg) {
    if (!cfg) return fail(NIQ_EINVAL, "mode cfg is NULL");
    if (cfg->mode < NIQ_MODE_INTERVAL || cfg->mode > NIQ_MODE_SLOPE_INTERVAL) return fail(NIQ_EINVAL, "invalid mode");
    if (cfg->mode == NIQ_MODE_SDF && !(cfg->sdf_lipschitz >= 0.f)) return fail(NIQ_EINVAL, "sdf mode: lipschitz bound must be >= 0");
    if (cfg->mode == NIQ_MODE_AFFINE_TRUNCATE && cfg->truncate_policy != 0)
        return fail(NIQ_EUNSUPPORTED, "truncate policy 'relative' is not supported (reference src/affine.py:146 broadcasts (k,)/(w,))");
    return NIQ_OK;
}
static bool is_fixed_mode(const niq_mode_cfg* cfg) { return cfg->mode == NIQ_MODE_INTERVAL || cfg->mode == NIQ_MODE_AFFINE_FIXED; }

// classify n boxes from a device-side source, any mode
static int classify_dev(niq_ctx* c, const niq_mlp* m, const niq_mode_cfg* cfg, BoxSource src, long long n, float offset,
                        int* label, float* lower, float* upper, unsigned char* tie) {
    if (is_fixed_mode(cfg)) {
        src.interval = cfg->mode == NIQ_MODE_INTERVAL;
        return launch_classify_fixed(c, m, src, n, offset, label, lower, upper, tie);
    }
    src.interval = 0;
    if (cfg->mode == NIQ_MODE_SLOPE_INTERVAL) return launch_classify_slope(c, m, src, n, offset, label, lower, upper, tie);
    if (cfg->mode == NIQ_MODE_SDF) {
        // reference src/sdf.py:31-50: one point evaluation at the box centre, then the Lipschitz test
        if (n <= 0) return NIQ_OK;
        DevBuf vals(c), scl(c);
        TRY(vals.alloc((size_t)n * 4)); TRY(scl.alloc((size_t)n * 4));
        PointSource ps{};
        ps.kind = 3; ps.box_kind = src.kind; ps.a = src.a; ps.b = src.b; ps.top = src.top; ps.window = src.window;
        TRY(launch_eval_points(c, m, ps, n, vals.as<float>(), scl.as<float>()));
        LaunchTimer lt(c, 1);
        k_sdf_labels<<<(int)((n + 255) / 256), 256, 0, c->stream>>>(src, n, vals.as<float>(), scl.as<float>(), cfg->sdf_lipschitz, offset,
                                                                     m->net.tie_rel, label, lower, upper, tie);
        CU(cudaGetLastError());
        return NIQ_OK;
    }
    return launch_classify_grow(c, m, cfg, src, n, offset, label, lower, upper, tie);
}

// exclusive scan of n ints into out[0..n] (out[n] = total); recursive over 2048-int tiles
static int scan_exclusive(niq_ctx* c, const int* in, long long n, int* out) {
    if (n <= 0) return NIQ_OK;
    const long long tiles = (n + kScanTile - 1) / kScanTile;
    if (tiles == 1) {
        LaunchTimer lt(c, 1);
        k_scan_apply<<<1, kScanThreads, 0, c->stream>>>(in, n, nullptr, out);
        CU(cudaGetLastError());
        return NIQ_OK;
    }
    DevBuf sums(c), offs(c);
    TRY(sums.alloc(tiles * sizeof(int)));
    TRY(offs.alloc((tiles + 1) * sizeof(int)));
    {
        LaunchTimer lt(c, 1);
        k_scan_tile_sums<<<(int)tiles, kScanThreads, 0, c->stream>>>(in, n, sums.as<int>());
        CU(cudaGetLastError());
    }
    TRY(scan_exclusive(c, sums.as<int>(), tiles, offs.as<int>()));
    {
        LaunchTimer lt(c, 1);
        k_scan_apply<<<(int)tiles, kScanThreads, 0, c->stream>>>(in, n, offs.as<int>(), out);
        CU(cudaGetLastError());
    }
    return NIQ_OK;
}

static int read_back(niq_ctx* c, const void* dsrc, size_t bytes, void* hdst) {
    if (bytes <= 64 * sizeof(long long)) {
        CU(cudaMemcpyAsync(c->pinned, dsrc, bytes, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        memcpy(hdst, c->pinned, bytes);
    } else {
        CU(cudaMemcpyAsync(hdst, dsrc, bytes, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    return NIQ_OK;
}

// host<->device staging for `mem` = NIQ_MEM_HOST
struct InBuf {
    DevBuf buf; const void* dev = nullptr;
    explicit InBuf(niq_ctx* c) : buf(c) {}
    int stage(niq_ctx* c, const void* p, size_t bytes, int mem) {
        if (!p) { dev = nullptr; return NIQ_OK; }
        if (mem == NIQ_MEM_DEVICE) { dev = p; return NIQ_OK; }
        TRY(buf.alloc(bytes));
        CU(cudaMemcpyAsync(buf.p, p, bytes, cudaMemcpyHostToDevice, c->stream));
        dev = buf.p;
        return NIQ_OK;
    }
    template <class T> const T* as() const { return reinterpret_cast<const T*>(dev); }
};
struct OutBuf {
    DevBuf buf; void* dev = nullptr; void* host = nullptr; size_t bytes = 0;
    explicit OutBuf(niq_ctx* c) : buf(c) {}
    int stage(niq_ctx* c, void* p, size_t nbytes, int mem) {
        (void)c;
        if (!p) { dev = nullptr; return NIQ_OK; }
        if (mem == NIQ_MEM_DEVICE) { dev = p; return NIQ_OK; }
        TRY(buf.alloc(nbytes));
        dev = buf.p; host = p; bytes = nbytes;
        return NIQ_OK;
    }
    int flush(niq_ctx* c) {
        if (host) CU(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
        return NIQ_OK;
    }
    template <class T> T* as() { return reinterpret_cast<T*>(dev); }
};

// ------------------------------------------------------------------------------------------------
// primitives
// ------------------------------------------------------------------------------------------------
extern "C" int niq_eval_points(niq_ctx* c, const niq_mlp* m, int64_t n, const float* x, float* f, float* scale, int mem) {
    if (!c || !m || n < 0 || (n > 0 && (!x || !f))) return fail(NIQ_EINVAL, "niq_eval_points: bad argument");
    if (n == 0) return NIQ_OK;
    CU(cudaSetDevice(c->device));
    timer_touch(c);
    InBuf dx(c); OutBuf df(c), ds(c);
    TRY(dx.stage(c, x, (size_t)n * 12, mem));
    TRY(df.stage(c, f, (size_t)n * 4, mem));
    TRY(ds.stage(c, scale, (size_t)n * 4, mem));
    PointSource src{};
    src.kind = 0; src.a = dx.as<float>();
    TRY(launch_eval_points(c, m, src, n, df.as<float>(), ds.as<float>()));
    TRY(df.flush(c)); TRY(ds.flush(c));
    FINAL_SYNC(c);
    return NIQ_OK;
}

static int classify_common(niq_ctx* c, const niq_mlp* m, const niq_mode_cfg* cfg, int64_t n, BoxSource src, size_t a_bytes,
                           size_t b_bytes, const float* a, const float* b, float offset, int32_t* label, float* lower,
                           float* upper, uint8_t* tie, int mem) {
    CU(cudaSetDevice(c->device));
    timer_touch(c);
    InBuf da(c), db(c); OutBuf dl(c), dlo(c), dup(c), dt(c);
    TRY(da.stage(c, a, a_bytes, mem));
    TRY(db.stage(c, b, b_bytes, mem));
    TRY(dl.stage(c, label, (size_t)n * 4, mem));
    TRY(dlo.stage(c, lower, (size_t)n * 4, mem));
    TRY(dup.stage(c, upper, (size_t)n * 4, mem));
    TRY(dt.stage(c, tie, (size_t)n, mem));
    src.a = da.as<float>(); src.b = db.as<float>();
    TRY(classify_dev(c, m, cfg, src, n, offset, dl.as<int>(), dlo.as<float>(), dup.as<float>(), dt.as<unsigned char>()));
    TRY(dl.flush(c)); TRY(dlo.flush(c)); TRY(dup.flush(c)); TRY(dt.flush(c));
    FINAL_SYNC(c);
    return NIQ_OK;
}

extern "C" int niq_classify_general_boxes(niq_ctx* c, const niq_mlp* m, const niq_mode_cfg* cfg, int64_t n, const float* center,
                                          const float* vecs, int32_t v, float offset, int32_t* label, float* lower,
                                          float* upper, uint8_t* tie, int mem) {
    if (!c || !m || n < 0 || (n > 0 && (!center || !vecs))) return fail(NIQ_EINVAL, "niq_classify_general_boxes: bad argument");
    TRY(check_cfg(cfg));
    if (v < 1) return fail(NIQ_EINVAL, "v must be >= 1");
    if ((is_fixed_mode(cfg) || cfg->mode == NIQ_MODE_SLOPE_INTERVAL) && v > 3)
        return fail(NIQ_EUNSUPPORTED, "interval / affine_fixed / slope_interval support v <= 3 box vectors (got %d)", v);
    if (n == 0) return NIQ_OK;
    BoxSource src{};
    src.kind = 0; src.v = v;
    return classify_common(c, m, cfg, n, src, (size_t)n * 12, (size_t)n * v * 12, center, vecs, offset, label, lower, upper, tie, mem);
}
extern "C" int niq_classify_boxes(niq_ctx* c, const niq_mlp* m, const niq_mode_cfg* cfg, int64_t n, const float* lo,
                                  const float* hi, float offset, int32_t* label, float* lower, float* upper, uint8_t* tie, int mem) {
    if (!c || !m || n < 0 || (n > 0 && (!lo || !hi))) return fail(NIQ_EINVAL, "niq_classify_boxes: bad argument");
    TRY(check_cfg(cfg));
    if (n == 0) return NIQ_OK;
    BoxSource src{};
    src.kind = 1; src.v = 3;
    return classify_common(c, m, cfg, n, src, (size_t)n * 12, (size_t)n * 12, lo, hi, offset, label, lower, upper, tie, mem);
}

// The slope-interval form of the output itself (reference src/slope_interval.py:15-33 `slope_interval_func` on
// coordinates_in_general_box(center, vecs)): raw (n,7) = [primal, slope centre x3, slope width x3] (unused vectors: 0),
// scale (n) or NULL = sum_j |h_j A_j| + |b| of the primal's last dot product (the near-tie yardstick).  What the
// min_distance_to_zero* helpers (:52-163) are computed from on the host side.
extern "C" int niq_slope_forward(niq_ctx* c, const niq_mlp* m, int64_t n, const float* center, const float* vecs, int32_t v,
                                 float* raw, float* scale, int mem) {
    if (!c || !m || n < 0 || (n > 0 && (!center || !vecs || !raw))) return fail(NIQ_EINVAL, "niq_slope_forward: bad argument");
    if (v < 1) return fail(NIQ_EINVAL, "v must be >= 1");
    if (v > 3) return fail(NIQ_EUNSUPPORTED, "slope_interval supports v <= 3 box vectors (got %d)", v);
    if (n == 0) return NIQ_OK;
    CU(cudaSetDevice(c->device));
    timer_touch(c);
    InBuf da(c), db(c); OutBuf dr(c), ds(c);
    TRY(da.stage(c, center, (size_t)n * 12, mem));
    TRY(db.stage(c, vecs, (size_t)n * v * 12, mem));
    TRY(dr.stage(c, raw, (size_t)n * 28, mem));
    TRY(ds.stage(c, scale, (size_t)n * 4, mem));
    BoxSource src{};
    src.kind = 0; src.v = v; src.interval = 0;
    src.a = da.as<float>(); src.b = db.as<float>();
    TRY(launch_classify_slope(c, m, src, n, 0.f, nullptr, nullptr, nullptr, nullptr, dr.as<float>(), ds.as<float>()));
    TRY(dr.flush(c)); TRY(ds.flush(c));
    FINAL_SYNC(c);
    return NIQ_OK;
}

// ------------------------------------------------------------------------------------------------
// cast_rays
// ------------------------------------------------------------------------------------------------
static int next_bucket(long long s, long long* out) {    // reference src/bucketing.py:7-14
    for (int p = 7; p < 31; ++p) if (s <= (1ll << p)) { *out = 1ll << p; return NIQ_OK; }
    return fail(NIQ_EINVAL, "max bucket size exceeded");
}

// concatenate the funcs' layer / chunk tables into one weight stream (cast_rays / cast_rays_frustum evaluate every func per step)
static int concat_nets(int n_funcs, const niq_mlp* const* mlps, NetDev& net, int& wmax, int& total_floats) {
    wmax = 32; total_floats = 0;
    for (int f = 0; f < n_funcs; ++f) {
        const NetDev& s = mlps[f]->net;
        if (net.n_layers + s.n_layers > kMaxLayers || net.n_chunks + s.n_chunks > kMaxChunks)
            return fail(NIQ_EUNSUPPORTED, "too many layers / weight chunks for one cast_rays launch");
        for (int l = 0; l < s.n_layers; ++l) {
            LayerDev L = s.layers[l];
            L.chunk_begin += net.n_chunks; L.chunk_end += net.n_chunks;
            net.layers[net.n_layers + l] = L;
        }
        for (int k = 0; k < s.n_chunks; ++k) {
            net.chunks[net.n_chunks + k] = s.chunks[k];
            net.chunks[net.n_chunks + k].smem_off += total_floats;     // nets sit one after the other when resident
        }
        total_floats += mlps[f]->total_floats;
        net.n_layers += s.n_layers; net.n_chunks += s.n_chunks;
        net.tie_rel = std::max(net.tie_rel, s.tie_rel);
        net.sparse = f == 0 ? s.sparse : std::min(net.sparse, s.sparse);
        wmax = std::max(wmax, mlps[f]->wmax);
    }
    net.n_nets = n_funcs;
    return NIQ_OK;
}

extern "C" int niq_cast_rays(niq_ctx* c, int32_t n_funcs, const niq_mlp* const* mlps, const niq_mode_cfg* cfgs,
                             const niq_cast_opts* o, int64_t n, const float* roots, const float* dirs, float* t,
                             int32_t* hit_id, int32_t* count, int64_t* n_evals, uint8_t* tie, int mem) {
    if (!c || !mlps || !cfgs || !o || n_funcs < 1 || n < 0) return fail(NIQ_EINVAL, "niq_cast_rays: bad argument");
    if (n > 0 && (!roots || !dirs || !t || !hit_id || !count)) return fail(NIQ_EINVAL, "niq_cast_rays: NULL array");
    if (o->n_substeps < 1) return fail(NIQ_EINVAL, "n_substeps must be >= 1");
    for (int f = 0; f < n_funcs; ++f) {
        if (!mlps[f]) return fail(NIQ_EINVAL, "mlp %d is NULL", f);
        TRY(check_cfg(&cfgs[f]));
        if (cfgs[f].mode != cfgs[0].mode) return fail(NIQ_EUNSUPPORTED, "all funcs of one cast_rays call must use the same mode");
    }
    if (n_evals) *n_evals = 0;
    if (n == 0) return NIQ_OK;
    CU(cudaSetDevice(c->device));
    timer_touch(c);
    InBuf dr(c), dd(c); OutBuf dt(c), dh(c), dc(c), dtie(c);
    TRY(dr.stage(c, roots, (size_t)n * 12, mem));
    TRY(dd.stage(c, dirs, (size_t)n * 12, mem));
    TRY(dt.stage(c, t, (size_t)n * 4, mem));
    TRY(dh.stage(c, hit_id, (size_t)n * 4, mem));
    TRY(dc.stage(c, count, (size_t)n * 4, mem));
    TRY(dtie.stage(c, tie, (size_t)n, mem));

    CastOpts co{};
    co.hit_eps = o->hit_eps; co.max_dist = o->max_dist; co.safety = o->safety_factor; co.grow = o->interval_grow_fac;
    co.shrink = o->interval_shrink_fac; co.n_max_step = o->n_max_step; co.n_substeps = o->n_substeps;
    co.init_step = (1.0f * o->interval_init_size) * o->max_dist;     // reference src/queries.py:149

    const bool slope = cfgs[0].mode == NIQ_MODE_SLOPE_INTERVAL;
    if (is_fixed_mode(&cfgs[0]) || slope) {
        NetDev net{};
        int wmax = 32, total_floats = 0;
        TRY(concat_nets(n_funcs, mlps, net, wmax, total_floats));
        DevBuf queue(c);
        TRY(queue.alloc(8));
        CU(cudaMemsetAsync(queue.p, 0, 8, c->stream));
        const int interval = cfgs[0].mode == NIQ_MODE_INTERVAL;
        TRY(launch_cast_rays(c, wmax, net, total_floats, co, n, interval, dr.as<float>(), dd.as<float>(), dt.as<float>(), dh.as<int>(), dc.as<int>(),
                             dtie.as<unsigned char>(), queue.as<unsigned long long>(), slope));
    } else if (cfgs[0].mode == NIQ_MODE_AFFINE_TRUNCATE || cfgs[0].mode == NIQ_MODE_AFFINE_ALL || cfgs[0].mode == NIQ_MODE_AFFINE_APPEND) {
        // growing-form modes: one CTA marches one ray, rays come from an atomic queue (niq_rays_grow.cuh)
        NetDev net{};
        int wmax = 32, total_floats = 0;
        TRY(concat_nets(n_funcs, mlps, net, wmax, total_floats));
        DevBuf queue(c);
        TRY(queue.alloc(8));
        CU(cudaMemsetAsync(queue.p, 0, 8, c->stream));
        TRY(launch_cast_rays_grow(c, n_funcs, mlps, cfgs, net, co, n, dr.as<float>(), dd.as<float>(), dt.as<float>(), dh.as<int>(), dc.as<int>(),
                                  dtie.as<unsigned char>(), queue.as<unsigned long long>()));
    } else {
        return fail(NIQ_EUNSUPPORTED, "cast_rays in sdf mode runs through the host-level stepping loop of the Python layer");
    }

    // N_evals of the reference (src/queries.py:137,164-173): lanes evaluated per iteration incl. bucket padding
    if (n_evals) {
        const int n_bins = o->n_max_step / o->n_substeps + 3;
        DevBuf hist(c);
        TRY(hist.alloc((size_t)n_bins * 8));
        CU(cudaMemsetAsync(hist.p, 0, (size_t)n_bins * 8, c->stream));
        {
            LaunchTimer lt(c, 1);
            k_iter_hist<<<(int)((n + 255) / 256), 256, 0, c->stream>>>(dc.as<int>(), n, o->n_substeps, n_bins, hist.as<unsigned long long>());
            CU(cudaGetLastError());
        }
        std::vector<unsigned long long> h(n_bins);
        TRY(read_back(c, hist.p, (size_t)n_bins * 8, h.data()));
        long long cur = n, valid = n, evals = 0;
        for (int it = 1; it < n_bins && valid > 0; ++it) {
            evals += cur * o->n_substeps;
            valid -= (long long)h[it];
            if (valid <= 0) break;
            long long nb;
            TRY(next_bucket(valid, &nb));
            if (nb < cur) cur = nb;
        }
        *n_evals = evals;
    }
    TRY(dt.flush(c)); TRY(dh.flush(c)); TRY(dc.flush(c)); TRY(dtie.flush(c));
    FINAL_SYNC(c);
    return NIQ_OK;
}

// ------------------------------------------------------------------------------------------------
// cast_rays_frustum (src/queries.py:178-587)
// ------------------------------------------------------------------------------------------------
extern "C" int niq_cast_rays_frustum(niq_ctx* c, int32_t n_funcs, const niq_mlp* const* mlps, const niq_mode_cfg* cfgs,
                                     const niq_cast_opts* o, const niq_camera* cam, float refine_width_fac, int64_t n_init,
                                     const int32_t* init_ranges, float* t, int32_t* hit_id, int32_t* count, int64_t* n_evals,
                                     int64_t* iter_counts, uint8_t* tie, int mem) {
    if (!c || !mlps || !cfgs || !o || !cam || n_funcs < 1 || n_init < 1 || !init_ranges)
        return fail(NIQ_EINVAL, "niq_cast_rays_frustum: bad argument");
    if (cam->res_x < 1 || cam->res_y < 1) return fail(NIQ_EINVAL, "niq_cast_rays_frustum: image resolution must be positive");
    if (!t || !hit_id || !count) return fail(NIQ_EINVAL, "niq_cast_rays_frustum: NULL array");
    if (o->n_substeps < 1) return fail(NIQ_EINVAL, "n_substeps must be >= 1");
    for (int f = 0; f < n_funcs; ++f) {
        if (!mlps[f]) return fail(NIQ_EINVAL, "mlp %d is NULL", f);
        TRY(check_cfg(&cfgs[f]));
        if (cfgs[f].mode != cfgs[0].mode) return fail(NIQ_EUNSUPPORTED, "all funcs of one cast_rays_frustum call must use the same mode");
    }
    const bool slope = cfgs[0].mode == NIQ_MODE_SLOPE_INTERVAL;
    const bool grow = cfgs[0].mode == NIQ_MODE_AFFINE_TRUNCATE || cfgs[0].mode == NIQ_MODE_AFFINE_ALL || cfgs[0].mode == NIQ_MODE_AFFINE_APPEND;
    if (!is_fixed_mode(&cfgs[0]) && !slope && !grow)
        return fail(NIQ_EUNSUPPORTED, "cast_rays_frustum in sdf mode runs through the host-level loop of the Python layer");
    const long long n = (long long)cam->res_x * cam->res_y;
    if (n_init > n) return fail(NIQ_EINVAL, "niq_cast_rays_frustum: more initial frusta than pixels");
    if (n_evals) *n_evals = 0;
    CU(cudaSetDevice(c->device));
    timer_touch(c);
    InBuf dinit(c); OutBuf dt(c), dh(c), dc(c), dtie(c);
    // the initial ranges always come from the host (a few hundred tiles), whatever `mem` says about the images
    TRY(dinit.stage(c, init_ranges, (size_t)n_init * 16, NIQ_MEM_HOST));
    TRY(dt.stage(c, t, (size_t)n * 4, mem));
    TRY(dh.stage(c, hit_id, (size_t)n * 4, mem));
    TRY(dc.stage(c, count, (size_t)n * 4, mem));
    TRY(dtie.stage(c, tie, (size_t)n, mem));

    // pixels outside the initial tiles (a rank's share of a sharded image) stay zero
    CU(cudaMemsetAsync(dt.dev, 0, (size_t)n * 4, c->stream));
    CU(cudaMemsetAsync(dh.dev, 0, (size_t)n * 4, c->stream));
    CU(cudaMemsetAsync(dc.dev, 0, (size_t)n * 4, c->stream));
    if (dtie.dev) CU(cudaMemsetAsync(dtie.dev, 0, (size_t)n, c->stream));

    CastOpts co{};
    co.hit_eps = o->hit_eps; co.max_dist = o->max_dist; co.safety = o->safety_factor; co.grow = o->interval_grow_fac;
    co.shrink = o->interval_shrink_fac; co.n_max_step = o->n_max_step; co.n_substeps = o->n_substeps;
    co.init_step = (1.0f * o->interval_init_size) * o->max_dist;     // reference src/queries.py:505
    FrustCam fc{};
    for (int d = 0; d < 3; ++d) { fc.root[d] = cam->root[d]; fc.look[d] = cam->look[d]; fc.up[d] = cam->up[d]; fc.left[d] = cam->left[d]; }
    fc.tan_x = cam->tan_half_fov_x; fc.tan_y = cam->tan_half_fov_y; fc.half_fov_x = cam->half_fov_x; fc.half_fov_y = cam->half_fov_y;
    fc.res_x = cam->res_x; fc.res_y = cam->res_y; fc.refine_fac = refine_width_fac;

    NetDev net{};
    int wmax = 32, total_floats = 0;
    TRY(concat_nets(n_funcs, mlps, net, wmax, total_floats));

    // every frustum covers >= 1 pixel and a split only partitions pixels: <= n records are ever pushed, <= n finish
    FrustQueue q{};
    q.cap = n_init + n;
    q.n_bins = o->n_max_step / o->n_substeps + 3;
    DevBuf rec(c), ready(c), ctrl(c), fin(c), hist(c);
    TRY(rec.alloc((size_t)q.cap * sizeof(FrustRec)));
    TRY(ready.alloc((size_t)q.cap * 4));
    TRY(ctrl.alloc(8 * 8));
    TRY(fin.alloc((size_t)n * sizeof(FrustFin)));
    TRY(hist.alloc((size_t)q.n_bins * 8));
    CU(cudaMemsetAsync(ready.p, 0, (size_t)q.cap * 4, c->stream));
    CU(cudaMemsetAsync(hist.p, 0, (size_t)q.n_bins * 8, c->stream));
    q.rec = rec.as<FrustRec>(); q.ready = ready.as<int>(); q.ctrl = ctrl.as<unsigned long long>(); q.fin = fin.as<FrustFin>();
    q.hist_term = hist.as<unsigned int>(); q.hist_ref = hist.as<unsigned int>() + q.n_bins;
    {
        LaunchTimer lt(c, 1);
        k_frustum_init<<<(int)((n_init + 255) / 256), 256, 0, c->stream>>>(q, dinit.as<int>(), n_init, co.init_step);
        CU(cudaGetLastError());
    }
    const int interval = cfgs[0].mode == NIQ_MODE_INTERVAL;
    if (grow) TRY(launch_cast_frustum_grow(c, n_funcs, mlps, cfgs, net, co, fc, q, n));
    else TRY(launch_cast_frustum(c, wmax, net, total_floats, co, fc, interval, q, n, slope));
    {
        LaunchTimer lt(c, 1);
        const int blocks = (int)std::min<long long>((n + 7) / 8, 8ll * c->prop.multiProcessorCount);
        k_frustum_fill<<<std::max(blocks, 1), 256, 0, c->stream>>>(q.fin, q.ctrl, fc.res_y, dt.as<float>(), dh.as<int>(), dc.as<int>(),
                                                                   dtie.as<unsigned char>());
        CU(cudaGetLastError());
    }
    // N_evals (src/queries.py:523-548): the padded array length of every marching iteration, replayed from the
    // per-iteration termination / split counts
    {
        std::vector<unsigned int> h(2 * (size_t)q.n_bins);
        unsigned long long hc[8];
        TRY(read_back(c, hist.p, h.size() * 4, h.data()));
        TRY(read_back(c, ctrl.p, sizeof(hc), hc));
        if (hc[4] != 0ull) return fail(NIQ_ECAPACITY, "cast_rays_frustum: work queue overflow (%llu records)", hc[4]);
        if (hc[2] != 0ull) return fail(NIQ_ECUDA, "cast_rays_frustum: %llu frusta left unfinished", hc[2]);
        if (iter_counts) for (size_t i = 0; i < h.size(); ++i) iter_counts[i] = (int64_t)h[i];
        if (n_evals) {
            long long size = n_init, empty_start = n_init, alive = n_init, evals = 0;
            for (int k = 0; k < q.n_bins; ++k) {
                evals += size;
                const long long n_valid = alive - (long long)h[k];
                if (n_valid <= 0) break;
                const long long n_ref = (long long)h[q.n_bins + k];
                long long nb;
                TRY(next_bucket(n_valid + n_ref, &nb));
                if (empty_start + n_ref > size || nb < size) { size = nb; empty_start = n_valid; }
                empty_start += n_ref;
                alive = n_valid + n_ref;
            }
            *n_evals = evals;
        }
    }
    TRY(dt.flush(c)); TRY(dh.flush(c)); TRY(dc.flush(c)); TRY(dtie.flush(c));
    FINAL_SYNC(c);
    return NIQ_OK;
}

// ------------------------------------------------------------------------------------------------
// level-set tree
// ------------------------------------------------------------------------------------------------
static int list_reserve(niq_ctx* c, NodeList& L, long long need) {
    if (need <= L.cap) return NIQ_OK;
    long long cap = std::max<long long>(need, std::max<long long>(2 * L.cap, 1024));
    float *lo = nullptr, *hi = nullptr;
    CU(cudaMallocAsync(&lo, (size_t)cap * 12, c->stream));
    CU(cudaMallocAsync(&hi, (size_t)cap * 12, c->stream));
    if (L.n > 0) {
        CU(cudaMemcpyAsync(lo, L.lo, (size_t)L.n * 12, cudaMemcpyDeviceToDevice, c->stream));
        CU(cudaMemcpyAsync(hi, L.hi, (size_t)L.n * 12, cudaMemcpyDeviceToDevice, c->stream));
    }
    if (L.lo) cudaFreeAsync(L.lo, c->stream);
    if (L.hi) cudaFreeAsync(L.hi, c->stream);
    L.lo = lo; L.hi = hi; L.cap = cap;
    return NIQ_OK;
}

extern "C" int niq_tree_destroy(niq_tree* t) {
    if (!t) return NIQ_OK;
    cudaSetDevice(t->ctx->device);
    for (auto& L : t->lists) {
        if (L.lo) cudaFreeAsync(L.lo, t->ctx->stream);
        if (L.hi) cudaFreeAsync(L.hi, t->ctx->stream);
    }
    cudaStreamSynchronize(t->ctx->stream);
    delete t;
    return NIQ_OK;
}

extern "C" int niq_tree_build_roots(niq_ctx* c, const niq_mlp* m, const niq_mode_cfg* cfg, int64_t n_roots, const float* lower,
                                    const float* upper, int32_t split_depth, int64_t node_thresh, float offset, int32_t flags,
                                    int32_t bps, niq_tree** out) {
    if (!c || !m || !lower || !upper || !out || n_roots < 1) return fail(NIQ_EINVAL, "niq_tree_build: bad argument");
    TRY(check_cfg(cfg));
    if (bps <= 0) return fail(NIQ_EINVAL, "batch_process_size must be positive");
    for (int p = 7; p < 31; ++p) {        // reference src/kd_tree.py:105-109
        const long long b = 1ll << p;
        if (b > bps && (b / bps) * bps != b)
            return fail(NIQ_EINVAL, "batch_process_size must be a factor of our bucket sizes, is not a factor of %lld (try a power of 2)", b);
    }
    if (split_depth < 0 && node_thresh <= 0)
        return fail(NIQ_EINVAL, "must specify at least one of node_terminate_thresh or split_depth as a terminating condition");
    if (node_thresh <= 0) node_thresh = 9999999999ll;
    CU(cudaSetDevice(c->device));
    timer_touch(c);

    niq_tree* T = new niq_tree();
    T->ctx = c;
    struct Guard { niq_tree* t; bool ok = false; ~Guard() { if (!ok) niq_tree_destroy(t); } } guard{T};
    {   // interval / affine_fixed / slope_interval: the whole build is one cooperative launch (niq_tree.cuh)
        bool handled = false;
        TRY(tree_build_persistent(c, m, cfg, n_roots, lower, upper, split_depth, node_thresh, offset, flags, bps, T, &handled));
        if (handled) { guard.ok = true; *out = T; return NIQ_OK; }
    }

    NodeList cur, nxt;
    struct ListGuard { niq_ctx* c; NodeList* L; ~ListGuard() { if (L->lo) cudaFreeAsync(L->lo, c->stream); if (L->hi) cudaFreeAsync(L->hi, c->stream); } } g1{c, &cur}, g2{c, &nxt};
    TRY(list_reserve(c, cur, n_roots));
    CU(cudaMemcpyAsync(cur.lo, lower, (size_t)n_roots * 12, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(cur.hi, upper, (size_t)n_roots * 12, cudaMemcpyHostToDevice, c->stream));
    cur.n = n_roots;
    long long bucket = 1;                                       // padded array size the reference would hold
    if (n_roots > 1) TRY(next_bucket(n_roots, &bucket));
    DevBuf d_tie(c);
    TRY(d_tie.alloc(8));
    CU(cudaMemsetAsync(d_tie.p, 0, 8, c->stream));
    const long long n_splits = split_depth < 0 ? 99999999ll : (long long)split_depth + 1;

    for (long long i_split = 0; i_split < n_splits; ++i_split) {
        const long long N = cur.n;
        const long long this_b = std::min<long long>(bps, bucket);
        const bool quit_next = (N >= node_thresh) || (i_split + 1 == n_splits);
        T->stats[2] += 1;
        T->stats[3] = std::max(T->stats[3], N);
        long long counts[4] = {0, 0, 0, 0};                     // unknown, negative, positive, near-tie
        if (N > 0) {
            DevBuf label(c), tie(c), f_unk(c), f_neg(c), f_pos(c), s_unk(c), s_neg(c), s_pos(c), s_tie(c), f_tie(c);
            TRY(label.alloc(N * 4)); TRY(tie.alloc(N));
            BoxSource src{};
            src.kind = 1; src.v = 3; src.a = cur.lo; src.b = cur.hi;
            TRY(classify_dev(c, m, cfg, src, N, offset, label.as<int>(), nullptr, nullptr, tie.as<unsigned char>()));
            T->stats[0] += N;
            const bool want_neg = flags & NIQ_TREE_INTERIOR, want_pos = flags & NIQ_TREE_EXTERIOR;
            TRY(f_unk.alloc(N * 4)); TRY(s_unk.alloc((N + 1) * 4));
            if (want_neg) { TRY(f_neg.alloc(N * 4)); TRY(s_neg.alloc((N + 1) * 4)); }
            if (want_pos) { TRY(f_pos.alloc(N * 4)); TRY(s_pos.alloc((N + 1) * 4)); }
            const int g = (int)((N + 255) / 256);
            {
                LaunchTimer lt(c, 1);
                k_tree_flags<<<g, 256, 0, c->stream>>>(label.as<int>(), N, f_unk.as<int>(), want_neg ? f_neg.as<int>() : nullptr,
                                                      want_pos ? f_pos.as<int>() : nullptr, tie.as<unsigned char>(),
                                                      d_tie.as<unsigned long long>());
                CU(cudaGetLastError());
            }
            TRY(scan_exclusive(c, f_unk.as<int>(), N, s_unk.as<int>()));
            if (want_neg) TRY(scan_exclusive(c, f_neg.as<int>(), N, s_neg.as<int>()));
            if (want_pos) TRY(scan_exclusive(c, f_pos.as<int>(), N, s_pos.as<int>()));
            int tot[3] = {0, 0, 0};
            CU(cudaMemcpyAsync(&c->pinned[0], s_unk.as<int>() + N, 4, cudaMemcpyDeviceToHost, c->stream));
            if (want_neg) CU(cudaMemcpyAsync(reinterpret_cast<int*>(c->pinned) + 1, s_neg.as<int>() + N, 4, cudaMemcpyDeviceToHost, c->stream));
            if (want_pos) CU(cudaMemcpyAsync(reinterpret_cast<int*>(c->pinned) + 2, s_pos.as<int>() + N, 4, cudaMemcpyDeviceToHost, c->stream));
            // near-tie count: reuse the scan machinery on the flags widened to int is overkill; count on host side of a tiny reduction
            CU(cudaStreamSynchronize(c->stream));
            tot[0] = reinterpret_cast<int*>(c->pinned)[0];
            if (want_neg) tot[1] = reinterpret_cast<int*>(c->pinned)[1];
            if (want_pos) tot[2] = reinterpret_cast<int*>(c->pinned)[2];
            counts[0] = tot[0]; counts[1] = tot[1]; counts[2] = tot[2];
            T->levels.insert(T->levels.end(), {N, counts[0], counts[1], counts[2]});
            if (want_neg && counts[1] > 0) {
                NodeList& L = T->lists[1];
                TRY(list_reserve(c, L, L.n + counts[1]));
                LaunchTimer lt(c, 1);
                k_append_flagged<<<g, 256, 0, c->stream>>>(cur.lo, cur.hi, N, f_neg.as<int>(), s_neg.as<int>(), L.n, L.lo, L.hi);
                CU(cudaGetLastError());
                L.n += counts[1];
            }
            if (want_pos && counts[2] > 0) {
                NodeList& L = T->lists[2];
                TRY(list_reserve(c, L, L.n + counts[2]));
                LaunchTimer lt(c, 1);
                k_append_flagged<<<g, 256, 0, c->stream>>>(cur.lo, cur.hi, N, f_pos.as<int>(), s_pos.as<int>(), L.n, L.lo, L.hi);
                CU(cudaGetLastError());
                L.n += counts[2];
            }
            const long long n_out = quit_next ? counts[0] : 2 * counts[0];
            TRY(list_reserve(c, nxt, std::max<long long>(n_out, 1)));
            if (counts[0] > 0) {
                LaunchTimer lt(c, 1);
                k_tree_scatter<<<g, 256, 0, c->stream>>>(cur.lo, cur.hi, N, f_unk.as<int>(), s_unk.as<int>(), this_b, quit_next ? 0 : 1, nxt.lo, nxt.hi);
                CU(cudaGetLastError());
            }
            nxt.n = n_out;                          // temporaries are stream-ordered (DevBuf): no host sync needed here
        } else {
            T->levels.insert(T->levels.end(), {0, 0, 0, 0});
            nxt.n = 0;
        }
        std::swap(cur, nxt);
        TRY(next_bucket(cur.n, &bucket));
        if (quit_next) break;
        if (cur.n == 0) break;      // nothing left to refine: no later level can add a node (the reference would idle through them)
    }
    // hand the final frontier to the tree object
    T->lists[0] = cur;
    cur = NodeList{};
    {   // near-tie boxes over all levels (device counter fed by k_tree_flags)
        unsigned long long nt = 0;
        TRY(read_back(c, d_tie.p, 8, &nt));
        T->stats[1] = (long long)nt;
    }
    FINAL_SYNC(c);
    guard.ok = true;
    *out = T;
    return NIQ_OK;
}

extern "C" int niq_tree_build(niq_ctx* c, const niq_mlp* m, const niq_mode_cfg* cfg, const float lower[3], const float upper[3],
                              int32_t split_depth, int64_t node_thresh, float offset, int32_t flags, int32_t bps, niq_tree** out) {
    return niq_tree_build_roots(c, m, cfg, 1, lower, upper, split_depth, node_thresh, offset, flags, bps, out);
}
// One rank's share of a tree whose subtrees are partitioned over `world` ranks: the levels above `deal_depth` are built as in
// niq_tree_build (replicated on every rank), the nodes entering level deal_depth are dealt round-robin, and the rank refines
// its own ones to split_depth -- one persistent launch (niq_tree.cuh TreeArgs::deal_*).  UNKNOWN leaves only.
extern "C" int niq_tree_build_dealt(niq_ctx* c, const niq_mlp* m, const niq_mode_cfg* cfg, const float lower[3], const float upper[3],
                                    int32_t split_depth, float offset, int32_t bps, int32_t deal_depth, int32_t rank, int32_t world,
                                    niq_tree** out) {
    if (!c || !m || !lower || !upper || !out) return fail(NIQ_EINVAL, "niq_tree_build_dealt: bad argument");
    TRY(check_cfg(cfg));
    if (split_depth < 0 || deal_depth < 0 || deal_depth > split_depth || world < 1 || rank < 0 || rank >= world || bps <= 0)
        return fail(NIQ_EINVAL, "niq_tree_build_dealt: need 0 <= deal_depth <= split_depth, 0 <= rank < world, batch_process_size > 0");
    CU(cudaSetDevice(c->device));
    timer_touch(c);
    niq_tree* T = new niq_tree();
    T->ctx = c;
    struct Guard { niq_tree* t; bool ok = false; ~Guard() { if (!ok) niq_tree_destroy(t); } } guard{T};
    bool handled = false;
    TRY(tree_build_persistent(c, m, cfg, 1, lower, upper, split_depth, 9999999999ll, offset, 0, bps, T, &handled, deal_depth, rank, world));
    if (!handled) return fail(NIQ_EUNSUPPORTED, "niq_tree_build_dealt: interval / affine_fixed / slope_interval only (other modes: build the top, deal on the host, niq_tree_build_roots)");
    guard.ok = true;
    *out = T;
    return NIQ_OK;
}
extern "C" int niq_tree_count(const niq_tree* t, int which, int64_t* n) {
    if (!t || !n || which < 0 || which > 2) return fail(NIQ_EINVAL, "bad argument");
    *n = t->lists[which].n;
    return NIQ_OK;
}
extern "C" int niq_tree_copy(const niq_tree* t, int which, float* lower, float* upper, int64_t capacity, int mem) {
    if (!t || which < 0 || which > 2) return fail(NIQ_EINVAL, "bad argument");
    const NodeList& L = t->lists[which];
    if (capacity < L.n) return fail(NIQ_ECAPACITY, "capacity %lld < %lld nodes", (long long)capacity, L.n);
    if (L.n == 0) return NIQ_OK;
    if (!lower || !upper) return fail(NIQ_EINVAL, "NULL output");
    niq_ctx* c = t->ctx;
    CU(cudaSetDevice(c->device));
    const cudaMemcpyKind k = mem == NIQ_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    CU(cudaMemcpyAsync(lower, L.lo, (size_t)L.n * 12, k, c->stream));
    CU(cudaMemcpyAsync(upper, L.hi, (size_t)L.n * 12, k, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return NIQ_OK;
}
extern "C" int niq_tree_level_info(const niq_tree* t, int32_t level, int64_t info[4]) {
    if (!t || !info || level < 0 || (size_t)level * 4 + 3 >= t->levels.size()) return fail(NIQ_EINVAL, "bad argument / level out of range");
    for (int i = 0; i < 4; ++i) info[i] = t->levels[(size_t)level * 4 + i];
    return NIQ_OK;
}
extern "C" int niq_tree_stats(const niq_tree* t, int64_t stats[4]) {
    if (!t || !stats) return fail(NIQ_EINVAL, "bad argument");
    for (int i = 0; i < 4; ++i) stats[i] = t->stats[i];
    return NIQ_OK;
}

// ------------------------------------------------------------------------------------------------
// marching cubes
// ------------------------------------------------------------------------------------------------
struct niq_mesh { niq_ctx* ctx = nullptr; float* tris = nullptr; long long n = 0; };

extern "C" int niq_mesh_destroy(niq_mesh* m) {
    if (!m) return NIQ_OK;
    cudaSetDevice(m->ctx->device);
    if (m->tris) cudaFreeAsync(m->tris, m->ctx->stream);
    cudaStreamSynchronize(m->ctx->stream);
    delete m;
    return NIQ_OK;
}

static int mc_device(niq_ctx* c, const niq_mlp* m, long long n, const float* lo, const float* hi, int n_sub, niq_mesh** out) {
    if (n_sub < 0 || n_sub > 5) return fail(NIQ_EINVAL, "n_subcell_depth must be in 0..5");
    niq_mesh* M = new niq_mesh();
    M->ctx = c;
    *out = M;
    if (n == 0) return NIQ_OK;
    timer_touch(c);
    const int side = 1 << n_sub, P = side + 1;
    const long long pts_per_leaf = (long long)P * P * P;
    // leaves are processed in slabs so the lattice values stay bounded (256 MB)
    const long long slab = std::max<long long>(1, (64ll << 20) / pts_per_leaf);
    const bool dedup = getenv("NIQ_MC_NO_DEDUP") == nullptr;       // development knob: A/B against one evaluation per leaf and point
    std::vector<std::pair<float*, long long>> parts;
    struct PartGuard { niq_ctx* c; std::vector<std::pair<float*, long long>>* p; ~PartGuard() { for (auto& q : *p) if (q.first) cudaFreeAsync(q.first, c->stream); } } pg{c, &parts};
    long long total = 0;
    for (long long s0 = 0; s0 < n; s0 += slab) {
        const long long L = std::min(slab, n - s0);
        DevBuf vals(c), cnt(c), off(c), table(c), nb(c), own_cnt(c), own_base(c);
        TRY(cnt.alloc(L * 4));
        TRY(off.alloc((L + 1) * 4));
        PointSource src{};
        src.a = lo + 3 * s0; src.b = hi + 3 * s0; src.pts_per_side = P;
        McArgs a{};
        a.leaf_lo = lo + 3 * s0; a.leaf_hi = hi + 3 * s0; a.n_leaves = L; a.n_side = side;
        long long n_pts = L * pts_per_leaf;
        if (dedup && L > 1) {
            // lattice points on a face shared with another leaf of the slab are evaluated once (k_mc_neighbours, mc_val)
            long long T = 16;
            while (T < 2 * L) T <<= 1;
            TRY(table.alloc((size_t)T * 4));
            TRY(nb.alloc((size_t)L * 12));
            TRY(own_cnt.alloc((size_t)L * 4));
            TRY(own_base.alloc((size_t)(L + 1) * 4));
            CU(cudaMemsetAsync(table.p, 0, (size_t)T * 4, c->stream));
            {
                LaunchTimer lt(c, 1);
                k_mc_hash_insert<<<(int)((L + 255) / 256), 256, 0, c->stream>>>(src.a, L, table.as<int>(), (unsigned)(T - 1));
                CU(cudaGetLastError());
            }
            {
                LaunchTimer lt(c, 1);
                k_mc_neighbours<<<(int)((L + 255) / 256), 256, 0, c->stream>>>(src.a, src.b, L, table.as<int>(), (unsigned)(T - 1), P,
                                                                               nb.as<int>(), own_cnt.as<int>());
                CU(cudaGetLastError());
            }
            TRY(scan_exclusive(c, own_cnt.as<int>(), L, own_base.as<int>()));
            int n_own = 0;
            TRY(read_back(c, own_base.as<int>() + L, 4, &n_own));
            n_pts = n_own;
            src.kind = 4; src.own_base = own_base.as<int>(); src.own_nb = nb.as<int>(); src.n_leaves = L;
            a.own_base = src.own_base; a.own_nb = src.own_nb;
        } else {
            src.kind = 1;
        }
        TRY(vals.alloc((size_t)n_pts * 4));
        TRY(launch_eval_points(c, m, src, n_pts, vals.as<float>(), nullptr));
        a.vals = vals.as<float>();
        c->mc_points_evaluated += n_pts; c->mc_points_lattice += L * pts_per_leaf;
        {
            LaunchTimer lt(c, 1);
            k_mc_count<<<(int)L, kScanThreads, 0, c->stream>>>(a, cnt.as<int>());
            CU(cudaGetLastError());
        }
        TRY(scan_exclusive(c, cnt.as<int>(), L, off.as<int>()));
        int tot = 0;
        TRY(read_back(c, off.as<int>() + L, 4, &tot));
        float* tri = nullptr;
        if (tot > 0) {
            CU(cudaMallocAsync(&tri, (size_t)tot * 36, c->stream));
            LaunchTimer lt(c, 1);
            k_mc_write<<<(int)L, kScanThreads, 0, c->stream>>>(a, off.as<int>(), tri);
            CU(cudaGetLastError());
        }
        parts.push_back({tri, (long long)tot});
        total += tot;
        CU(cudaStreamSynchronize(c->stream));
    }
    if (total > 0) {
        if (parts.size() == 1) {
            M->tris = parts[0].first;
            parts[0].first = nullptr;
        } else {
            CU(cudaMallocAsync(&M->tris, (size_t)total * 36, c->stream));
            long long o = 0;
            for (auto& q : parts) {
                if (q.second > 0) CU(cudaMemcpyAsync(M->tris + o * 9, q.first, (size_t)q.second * 36, cudaMemcpyDeviceToDevice, c->stream));
                o += q.second;
            }
        }
    }
    M->n = total;
    FINAL_SYNC(c);
    return NIQ_OK;
}

extern "C" int niq_marching_cubes(niq_ctx* c, const niq_mlp* m, int64_t n, const float* leaf_lower, const float* leaf_upper,
                                  int32_t n_sub, int mem, niq_mesh** out) {
    if (!c || !m || !out || n < 0 || (n > 0 && (!leaf_lower || !leaf_upper))) return fail(NIQ_EINVAL, "niq_marching_cubes: bad argument");
    CU(cudaSetDevice(c->device));
    InBuf dlo(c), dhi(c);
    TRY(dlo.stage(c, leaf_lower, (size_t)n * 12, mem));
    TRY(dhi.stage(c, leaf_upper, (size_t)n * 12, mem));
    niq_mesh* M = nullptr;
    int r = mc_device(c, m, n, dlo.as<float>(), dhi.as<float>(), n_sub, &M);
    if (r != NIQ_OK) { niq_mesh_destroy(M); return r; }
    *out = M;
    return NIQ_OK;
}
extern "C" int niq_marching_cubes_tree(niq_ctx* c, const niq_mlp* m, const niq_tree* t, int32_t n_sub, niq_mesh** out) {
    if (!c || !m || !t || !out) return fail(NIQ_EINVAL, "niq_marching_cubes_tree: bad argument");
    CU(cudaSetDevice(c->device));
    niq_mesh* M = nullptr;
    int r = mc_device(c, m, t->lists[0].n, t->lists[0].lo, t->lists[0].hi, n_sub, &M);
    if (r != NIQ_OK) { niq_mesh_destroy(M); return r; }
    *out = M;
    return NIQ_OK;
}
extern "C" int niq_mesh_count(const niq_mesh* m, int64_t* n) {
    if (!m || !n) return fail(NIQ_EINVAL, "bad argument");
    *n = m->n;
    return NIQ_OK;
}
extern "C" int niq_mesh_copy(const niq_mesh* m, float* tri_pos, int64_t capacity, int mem) {
    if (!m) return fail(NIQ_EINVAL, "bad argument");
    if (capacity < m->n) return fail(NIQ_ECAPACITY, "capacity %lld < %lld triangles", (long long)capacity, m->n);
    if (m->n == 0) return NIQ_OK;
    if (!tri_pos) return fail(NIQ_EINVAL, "NULL output");
    niq_ctx* c = m->ctx;
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(tri_pos, m->tris, (size_t)m->n * 36, mem == NIQ_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return NIQ_OK;
}
extern "C" int niq_mc_tables(int32_t* tri_table, int32_t* edge_verts, uint8_t* vert_coords) {
    static const unsigned long long words[256] = NIQ_MC_CASE_WORDS_INIT;
    if (tri_table)
        for (int cs = 0; cs < 256; ++cs)
            for (int i = 0; i < 16; ++i) {
                const int nb = (int)((words[cs] >> (4 * i)) & 0xF);
                tri_table[cs * 16 + i] = nb == 0xF ? -1 : nb;
            }
    if (edge_verts)
        for (int e = 0; e < 12; ++e) {
            edge_verts[2 * e] = (int)((NIQ_MC_EDGE_A_NIBBLES >> (4 * e)) & 0xF);
            edge_verts[2 * e + 1] = (int)((NIQ_MC_EDGE_B_NIBBLES >> (4 * e)) & 0xF);
        }
    if (vert_coords)
        for (int v = 0; v < 8; ++v) {
            vert_coords[3 * v] = (NIQ_MC_VERT_MASK_X >> v) & 1;
            vert_coords[3 * v + 1] = (NIQ_MC_VERT_MASK_Y >> v) & 1;
            vert_coords[3 * v + 2] = (NIQ_MC_VERT_MASK_Z >> v) & 1;
        }
    return NIQ_OK;
}

// ------------------------------------------------------------------------------------------------
// find_any_intersection
// ------------------------------------------------------------------------------------------------
extern "C" int niq_find_any_intersection(niq_ctx* c, const niq_mlp* mA, const niq_mode_cfg* cfgA, const niq_mlp* mB,
                                         const niq_mode_cfg* cfgB, const float lower[3], const float upper[3], float eps,
                                         int32_t* found, float loc[3], int64_t stats[3]) {
    if (!c || !mA || !mB || !lower || !upper || !found || !loc) return fail(NIQ_EINVAL, "niq_find_any_intersection: bad argument");
    TRY(check_cfg(cfgA)); TRY(check_cfg(cfgB));
    CU(cudaSetDevice(c->device));
    timer_touch(c);
    {   // affine_truncate / affine_all / affine_append: the whole search is one persistent cooperative kernel (niq_isect.cuh)
        bool handled = false;
        int64_t st3[3] = {0, 0, 0};
        TRY(isect_grow_batch(c, mA, cfgA, mB, cfgB, 1, nullptr, nullptr, lower, upper, eps, found, loc, st3, &handled));
        if (handled) {
            if (stats) { stats[0] = st3[0]; stats[1] = st3[1]; stats[2] = st3[2]; }
            return NIQ_OK;
        }
    }
    const float eps_w = eps / sqrtf(3.0f);                 // reference src/kd_tree.py:446
    NodeList cur, nxt;
    struct ListGuard { niq_ctx* c; NodeList* L; ~ListGuard() { if (L->lo) cudaFreeAsync(L->lo, c->stream); if (L->hi) cudaFreeAsync(L->hi, c->stream); } } g1{c, &cur}, g2{c, &nxt};
    TRY(list_reserve(c, cur, 1));
    CU(cudaMemcpyAsync(cur.lo, lower, 12, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(cur.hi, upper, 12, cudaMemcpyHostToDevice, c->stream));
    cur.n = 1;
    long long n_nodes = 0, n_rounds = 0, n_tie = 0;
    *found = 0;
    loc[0] = loc[1] = loc[2] = -777.f;
    DevBuf meta(c);                       // [0] first found index (u64), [1] near-tie boxes (u64), [2] survivors of the round (int)
    TRY(meta.alloc(32));
    while (cur.n > 0) {
        const long long N = cur.n;
        n_nodes += N; n_rounds += 1;
        DevBuf labA(c), labB(c), vA(c), vB(c), needs(c), scan(c), locs(c), tieA(c), tieB(c);
        TRY(labA.alloc(N * 4)); TRY(labB.alloc(N * 4)); TRY(vA.alloc(N * 28)); TRY(vB.alloc(N * 28));
        TRY(needs.alloc(N * 4)); TRY(scan.alloc((N + 1) * 4)); TRY(locs.alloc(N * 12));
        TRY(tieA.alloc(N)); TRY(tieB.alloc(N));
        {
            const unsigned long long init[4] = {~0ull, 0ull, 0ull, 0ull};
            memcpy(c->pinned + 8, init, 32);
            CU(cudaMemcpyAsync(meta.p, c->pinned + 8, 32, cudaMemcpyHostToDevice, c->stream));
        }
        BoxSource bs{};
        bs.kind = 1; bs.v = 3; bs.a = cur.lo; bs.b = cur.hi;
        PointSource ps{};
        ps.kind = 2; ps.a = cur.lo; ps.b = cur.hi; ps.sample_scale = eps_w;
        // shape B on the side stream, beside shape A (the frontier is small: each launch alone leaves most SMs idle)
        CU(cudaEventRecord(c->fork, c->stream));
        CU(cudaStreamWaitEvent(c->stream2, c->fork, 0));
        // from here on stream2 may be running kernels on this round's buffers: whatever path leaves the scope, the main
        // stream first waits for the side stream, so the stream-ordered frees of the DevBufs cannot overtake it
        struct JoinGuard { niq_ctx* c; ~JoinGuard() { cudaEventRecord(c->join, c->stream2); cudaStreamWaitEvent(c->stream, c->join, 0); } } jg{c};
        int rB = NIQ_OK;
        {
            const bool timing = c->timing;
            c->timing = false;
            std::swap(c->stream, c->stream2);
            rB = classify_dev(c, mB, cfgB, bs, N, 0.f, labB.as<int>(), nullptr, nullptr, tieB.as<unsigned char>());
            if (rB == NIQ_OK) rB = launch_eval_points(c, mB, ps, 7 * N, vB.as<float>(), nullptr);
            std::swap(c->stream, c->stream2);
            c->timing = timing;
        }
        CU(cudaEventRecord(c->join, c->stream2));
        TRY(rB);
        TRY(classify_dev(c, mA, cfgA, bs, N, 0.f, labA.as<int>(), nullptr, nullptr, tieA.as<unsigned char>()));
        TRY(launch_eval_points(c, mA, ps, 7 * N, vA.as<float>(), nullptr));
        CU(cudaStreamWaitEvent(c->stream, c->join, 0));
        const int g = (int)((N + 255) / 256);
        {
            LaunchTimer lt(c, 1);
            k_isect_logic<<<g, 256, 0, c->stream>>>(cur.lo, cur.hi, N, labA.as<int>(), labB.as<int>(), vA.as<float>(), vB.as<float>(),
                                                   eps_w, needs.as<int>(), locs.as<float>(), meta.as<unsigned long long>(),
                                                   tieA.as<unsigned char>(), tieB.as<unsigned char>(), meta.as<unsigned long long>() + 1);
            CU(cudaGetLastError());
        }
        TRY(scan_exclusive(c, needs.as<int>(), N, scan.as<int>()));
        CU(cudaMemcpyAsync(meta.as<unsigned long long>() + 2, scan.as<int>() + N, 4, cudaMemcpyDeviceToDevice, c->stream));
        unsigned long long hm[4];
        TRY(read_back(c, meta.p, 32, hm));                 // the round's only host synchronisation
        n_tie += (long long)hm[1];
        if (hm[0] != ~0ull) {
            TRY(read_back(c, locs.as<float>() + 3 * hm[0], 12, loc));
            *found = 1;
            break;
        }
        const int n_new = (int)(hm[2] & 0xffffffffull);
        TRY(list_reserve(c, nxt, std::max<long long>(2ll * n_new, 1)));
        if (n_new > 0) {
            LaunchTimer lt(c, 1);
            k_split_interleaved<<<g, 256, 0, c->stream>>>(cur.lo, cur.hi, nullptr, N, needs.as<int>(), scan.as<int>(), nxt.lo, nxt.hi, nullptr);
            CU(cudaGetLastError());
        }
        nxt.n = 2ll * n_new;
        std::swap(cur, nxt);
    }
    if (stats) { stats[0] = n_nodes; stats[1] = n_rounds; stats[2] = n_tie; }
    FINAL_SYNC(c);
    return NIQ_OK;
}

extern "C" int niq_find_any_intersection_batch(niq_ctx* c, const niq_mlp* mA, const niq_mode_cfg* cfgA, const niq_mlp* mB,
                                               const niq_mode_cfg* cfgB, int64_t n_queries, const float* xfA, const float* xfB,
                                               const float lower[3], const float upper[3], float eps, int32_t* found, float* loc,
                                               int64_t* stats) {
    if (!c || !mA || !mB || !lower || !upper || n_queries < 0 || (n_queries > 0 && (!found || !loc)))
        return fail(NIQ_EINVAL, "niq_find_any_intersection_batch: bad argument");
    TRY(check_cfg(cfgA)); TRY(check_cfg(cfgB));
    if (n_queries == 0) return NIQ_OK;
    CU(cudaSetDevice(c->device));
    timer_touch(c);
    bool handled = false;
    TRY(isect_grow_batch(c, mA, cfgA, mB, cfgB, n_queries, xfA, xfB, lower, upper, eps, found, loc, stats, &handled));
    if (!handled)
        return fail(NIQ_EUNSUPPORTED, "niq_find_any_intersection_batch runs the growing-form modes (affine_truncate / affine_all / "
                                      "affine_append); other modes: one niq_find_any_intersection call per query");
    return NIQ_OK;
}

// ------------------------------------------------------------------------------------------------
// closest_point
// ------------------------------------------------------------------------------------------------
extern "C" int niq_closest_point(niq_ctx* c, const niq_mlp* m, const niq_mode_cfg* cfg, const float lower[3], const float upper[3],
                                 int64_t q, const float* query_points, float eps, int64_t B, float* dist, float* loc,
                                 int64_t stats[4], int mem) {
    if (!c || !m || !lower || !upper || q < 0 || B <= 0) return fail(NIQ_EINVAL, "niq_closest_point: bad argument");
    if (q > 0 && (!query_points || !dist || !loc)) return fail(NIQ_EINVAL, "niq_closest_point: NULL array");
    TRY(check_cfg(cfg));
    if (stats) stats[0] = stats[1] = stats[2] = stats[3] = 0;
    if (q == 0) return NIQ_OK;
    CU(cudaSetDevice(c->device));
    timer_touch(c);
    InBuf dq(c); OutBuf dd(c), dl(c);
    TRY(dq.stage(c, query_points, (size_t)q * 12, mem));
    TRY(dd.stage(c, dist, (size_t)q * 4, mem));
    TRY(dl.stage(c, loc, (size_t)q * 12, mem));

    // the effective window never exceeds what the stack can hold; B >= stack size behaves like "everything"
    const long long Bw = B;
    long long cap = 0;
    float *s_lo = nullptr, *s_hi = nullptr;
    long long* s_id = nullptr;
    struct StackGuard { niq_ctx* c; float** a; float** b; long long** d; ~StackGuard() { if (*a) cudaFreeAsync(*a, c->stream); if (*b) cudaFreeAsync(*b, c->stream); if (*d) cudaFreeAsync(*d, c->stream); } } sg{c, &s_lo, &s_hi, &s_id};
    auto reserve = [&](long long need, long long live) -> int {
        if (need <= cap) return NIQ_OK;
        long long ncap = std::max<long long>(need, 2 * cap);
        float *nlo = nullptr, *nhi = nullptr; long long* nid = nullptr;
        CU(cudaMallocAsync(&nlo, (size_t)ncap * 12, c->stream));
        CU(cudaMallocAsync(&nhi, (size_t)ncap * 12, c->stream));
        CU(cudaMallocAsync(&nid, (size_t)ncap * 8, c->stream));
        CU(cudaMemsetAsync(nlo, 0, (size_t)ncap * 12, c->stream));
        CU(cudaMemsetAsync(nhi, 0, (size_t)ncap * 12, c->stream));
        CU(cudaMemsetAsync(nid, 0, (size_t)ncap * 8, c->stream));
        if (live > 0) {
            CU(cudaMemcpyAsync(nlo, s_lo, (size_t)live * 12, cudaMemcpyDeviceToDevice, c->stream));
            CU(cudaMemcpyAsync(nhi, s_hi, (size_t)live * 12, cudaMemcpyDeviceToDevice, c->stream));
            CU(cudaMemcpyAsync(nid, s_id, (size_t)live * 8, cudaMemcpyDeviceToDevice, c->stream));
        }
        if (s_lo) cudaFreeAsync(s_lo, c->stream);
        if (s_hi) cudaFreeAsync(s_hi, c->stream);
        if (s_id) cudaFreeAsync(s_id, c->stream);
        s_lo = nlo; s_hi = nhi; s_id = nid; cap = ncap;
        return NIQ_OK;
    };

    // Each round pops the top min(B, top) entries (reference :679-686).  `ub` is a host-side upper bound of
    // the device-resident stack top, exact after every poll.
    const int kPoll = 8;                                  // rounds between polls in the windowed regime
    long long ub = q;
    TRY(reserve(ub + 3 * std::min<long long>(Bw, ub) + 16, 0));
    {   // initial stack: one root box per query (reference src/kd_tree.py:769-775)
        std::vector<float> hlo((size_t)q * 3), hhi((size_t)q * 3);
        std::vector<long long> hid((size_t)q);
        for (long long i = 0; i < q; ++i) {
            for (int d = 0; d < 3; ++d) { hlo[3 * i + d] = lower[d]; hhi[3 * i + d] = upper[d]; }
            hid[i] = i;
        }
        CU(cudaMemcpyAsync(s_lo, hlo.data(), (size_t)q * 12, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(s_hi, hhi.data(), (size_t)q * 12, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(s_id, hid.data(), (size_t)q * 8, cudaMemcpyHostToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    DevBuf d_top(c), d_stats(c), d_winner(c), d_maxtop(c);
    TRY(d_top.alloc(8)); TRY(d_stats.alloc(32)); TRY(d_winner.alloc((size_t)q * 8)); TRY(d_maxtop.alloc(8));
    CU(cudaMemcpyAsync(d_top.p, &ub, 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d_maxtop.p, &ub, 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemsetAsync(d_stats.p, 0, 32, c->stream));
    CU(cudaMemsetAsync(d_winner.p, 0, (size_t)q * 8, c->stream));
    {   // min_dist = +inf, min_loc = -777 (reference :776-777)
        std::vector<float> hd((size_t)q, INFINITY), hl((size_t)q * 3, -777.f);
        CU(cudaMemcpyAsync(dd.dev, hd.data(), (size_t)q * 4, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(dl.dev, hl.data(), (size_t)q * 12, cudaMemcpyHostToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    const float eps_w = eps / sqrtf(3.0f);                 // reference :689
    unsigned long long round = 0;
    // Windowed regime with a window of <= 2048 entries (the reference's default): one round = classify + 7-sample
    // evaluation + ONE fused single-CTA kernel, captured once into a CUDA graph and replayed (the launch-bound inner
    // loop of the query); buffers of the graph live for the whole call.
    const bool small_window = Bw <= 2 * kCpSmallThreads;
    // ---- the reference's default regime (window <= 2048), fixed-row modes, resident weights: every round of the search inside
    // ONE cooperative kernel (niq_cp.cuh); the stack top never leaves the device ----
    if (small_window && is_fixed_mode(cfg) && !(getenv("NIQ_CP_LEGACY") && getenv("NIQ_CP_LEGACY")[0] != '0')) {
        int coop = 0;
        CU(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device));
        DevBuf p_label(c), p_tie(c), p_vals(c), p_ctl(c);
        TRY(p_label.alloc(Bw * 4)); TRY(p_tie.alloc(Bw)); TRY(p_vals.alloc(Bw * 28)); TRY(p_ctl.alloc(sizeof(CpCtl)));
        CpCtl h{};
        h.top = q; h.max_top = q;
        CU(cudaMemcpyAsync(p_ctl.p, &h, sizeof(h), cudaMemcpyHostToDevice, c->stream));
        bool ran = false;
        for (int attempt = 0; coop && attempt < 40; ++attempt) {
            // room for the window's children of many rounds; a round that would not fit stops the kernel at its boundary
            TRY(reserve(std::max<long long>(h.top, q) + 64 * Bw + 16, h.top));
            CpArgs a{};
            a.stack_lo = s_lo; a.stack_hi = s_hi; a.stack_qid = s_id; a.cap = cap; a.window = Bw;
            a.query = dq.as<float>(); a.min_dist = dd.as<float>(); a.min_loc = dl.as<float>(); a.winner = d_winner.as<unsigned long long>();
            a.n_query = q; a.label = p_label.as<int>(); a.tie = p_tie.as<unsigned char>(); a.vals = p_vals.as<float>();
            a.eps_w = eps_w; a.interval = cfg->mode == NIQ_MODE_INTERVAL; a.ctl = p_ctl.as<CpCtl>();
            bool fits = false;
            TRY(launch_cp_persistent(c, m, a, &fits));
            if (!fits) break;                       // streamed / wide nets: the CUDA-graph round loop below
            ran = true;
            timer_mark(c);
            CU(cudaMemcpyAsync(&h, p_ctl.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            if (h.status == 0) break;
            h.status = 0;                           // the stack was too small for the next round: grow it and go on
            CU(cudaMemcpyAsync(p_ctl.p, &h, sizeof(h), cudaMemcpyHostToDevice, c->stream));
            if (attempt == 39) return fail(NIQ_ENOMEM, "closest_point: the stack kept growing (%lld entries)", h.top);
        }
        if (ran) {
            if (stats) { stats[0] = h.stats[0]; stats[1] = h.stats[1]; stats[2] = h.max_top; stats[3] = h.stats[2]; }
            c->launches += 0;
            TRY(dd.flush(c)); TRY(dl.flush(c));
            CU(cudaStreamSynchronize(c->stream));
            return NIQ_OK;
        }
    }
    DevBuf g_label(c), g_tie(c), g_vals(c);
    cudaGraphExec_t cp_exec = nullptr;
    struct GraphGuard { cudaGraphExec_t* e; ~GraphGuard() { if (*e) cudaGraphExecDestroy(*e); } } gg{&cp_exec};
    const float* graph_stack = nullptr;
    if (small_window) { TRY(g_label.alloc(Bw * 4)); TRY(g_tie.alloc(Bw)); TRY(g_vals.alloc(Bw * 28)); }
    while (true) {
        // ub <= B: the window covers the whole stack (per-query level-synchronous regime) -> poll every round so
        // the launch size tracks the stack; ub > B: steady windows of exactly B entries -> poll every kPoll rounds.
        const bool level_regime = ub <= Bw;
        const bool fused = !level_regime && small_window;
        const int rounds = level_regime ? 1 : (fused ? 4 * kPoll : kPoll);
        const long long W = std::min<long long>(Bw, std::max<long long>(ub, 1));
        TRY(reserve(ub + (rounds + 2) * W + 16, ub));
        if (fused) {
            if (cp_exec == nullptr || graph_stack != s_lo) {
                if (cp_exec) { cudaGraphExecDestroy(cp_exec); cp_exec = nullptr; }
                const bool timing = c->timing;
                c->timing = false;
                const long long launches0 = c->launches;
                CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
                BoxSource bs{};
                bs.kind = 2; bs.v = 3; bs.a = s_lo; bs.b = s_hi; bs.top = d_top.as<long long>(); bs.window = W;
                int rc = classify_dev(c, m, cfg, bs, W, 0.f, g_label.as<int>(), nullptr, nullptr, g_tie.as<unsigned char>());
                PointSource ps{};
                ps.kind = 2; ps.a = s_lo; ps.b = s_hi; ps.sample_scale = -1.f; ps.top = d_top.as<long long>(); ps.window = W;
                if (rc == NIQ_OK) rc = launch_eval_points(c, m, ps, 7 * W, g_vals.as<float>(), nullptr);
                CpRound r{};
                r.stack_lo = s_lo; r.stack_hi = s_hi; r.stack_qid = s_id; r.top = d_top.as<long long>(); r.window = W;
                r.query = dq.as<float>(); r.min_dist = dd.as<float>(); r.min_loc = dl.as<float>(); r.winner = d_winner.as<unsigned long long>();
                r.n_query = q; r.label = g_label.as<int>(); r.tie = g_tie.as<unsigned char>(); r.vals = g_vals.as<float>(); r.eps_w = eps_w;
                r.stats = d_stats.as<long long>();
                k_cp_round_small<<<1, kCpSmallThreads, 0, c->stream>>>(r, s_lo, s_hi, s_id, d_maxtop.as<long long>());
                cudaGraph_t graph = nullptr;
                const cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
                c->timing = timing;
                c->launches = launches0;
                if (rc != NIQ_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
                if (ce != cudaSuccess) return fail(NIQ_ECUDA, "closest_point: graph capture failed: %s", cudaGetErrorString(ce));
                const cudaError_t ie = cudaGraphInstantiate(&cp_exec, graph, 0);
                cudaGraphDestroy(graph);
                if (ie != cudaSuccess) { cp_exec = nullptr; return fail(NIQ_ECUDA, "closest_point: graph instantiation failed: %s", cudaGetErrorString(ie)); }
                graph_stack = s_lo;
            }
            // the fused kernel keeps the round number on the device (winner tags must keep growing across regimes)
            CU(cudaMemcpyAsync(d_stats.as<long long>() + 3, &round, 8, cudaMemcpyHostToDevice, c->stream));
            for (int it = 0; it < rounds; ++it) CU(cudaGraphLaunch(cp_exec, c->stream));
            c->launches += 3ll * rounds;
            round += (unsigned long long)rounds;
            TRY(read_back(c, d_top.p, 8, &ub));
            if (ub <= 0) break;
            continue;
        }
        DevBuf label(c), tie(c), vals(c), this_d(c), cen(c), needs(c), scan(c), t_lo(c), t_hi(c), t_id(c);
        TRY(label.alloc(W * 4)); TRY(tie.alloc(W)); TRY(vals.alloc(W * 28)); TRY(this_d.alloc(W * 4)); TRY(cen.alloc(W * 12));
        TRY(needs.alloc(W * 4)); TRY(scan.alloc((W + 1) * 4)); TRY(t_lo.alloc(W * 12)); TRY(t_hi.alloc(W * 12)); TRY(t_id.alloc(W * 8));
        for (int it = 0; it < rounds; ++it) {
            BoxSource bs{};
            bs.kind = 2; bs.v = 3; bs.a = s_lo; bs.b = s_hi; bs.top = d_top.as<long long>(); bs.window = W;
            TRY(classify_dev(c, m, cfg, bs, W, 0.f, label.as<int>(), nullptr, nullptr, tie.as<unsigned char>()));
            PointSource ps{};
            ps.kind = 2; ps.a = s_lo; ps.b = s_hi; ps.sample_scale = -1.f; ps.top = d_top.as<long long>(); ps.window = W;
            TRY(launch_eval_points(c, m, ps, 7 * W, vals.as<float>(), nullptr));
            CpRound r{};
            r.stack_lo = s_lo; r.stack_hi = s_hi; r.stack_qid = s_id; r.top = d_top.as<long long>(); r.window = W;
            r.query = dq.as<float>(); r.min_dist = dd.as<float>(); r.min_loc = dl.as<float>(); r.winner = d_winner.as<unsigned long long>();
            r.n_query = q; r.label = label.as<int>(); r.tie = tie.as<unsigned char>(); r.vals = vals.as<float>(); r.eps_w = eps_w; r.round = round;
            r.this_dist = this_d.as<float>(); r.center = cen.as<float>(); r.needs = needs.as<int>(); r.stats = d_stats.as<long long>();
            const int g = (int)((W + 255) / 256);
            {
                LaunchTimer lt(c, 1);
                k_cp_eval<<<g, 256, 0, c->stream>>>(r);
                k_cp_min<<<g, 256, 0, c->stream>>>(r);
                k_cp_winner<<<g, 256, 0, c->stream>>>(r);
                k_cp_loc<<<g, 256, 0, c->stream>>>(r);
                k_cp_copy_window<<<g, 256, 0, c->stream>>>(r, t_lo.as<float>(), t_hi.as<float>(), t_id.as<long long>());
                CU(cudaGetLastError());
                c->launches += 4;
            }
            TRY(scan_exclusive(c, needs.as<int>(), W, scan.as<int>()));
            {
                LaunchTimer lt(c, 1);
                k_cp_push<<<g, 256, 0, c->stream>>>(r, t_lo.as<float>(), t_hi.as<float>(), t_id.as<long long>(), scan.as<int>(), s_lo, s_hi, s_id);
                k_cp_advance<<<1, 1, 0, c->stream>>>(r, scan.as<int>(), d_maxtop.as<long long>());
                CU(cudaGetLastError());
                c->launches += 1;
            }
            round += 1;
        }
        TRY(read_back(c, d_top.p, 8, &ub));
        if (ub <= 0) break;
    }
    if (stats) {
        long long hs[4] = {0, 0, 0, 0};
        TRY(read_back(c, d_stats.p, 32, hs));
        long long mt = 0;
        TRY(read_back(c, d_maxtop.p, 8, &mt));
        stats[0] = hs[0]; stats[1] = hs[1]; stats[2] = mt; stats[3] = hs[2];
    }
    TRY(dd.flush(c)); TRY(dl.flush(c));
    FINAL_SYNC(c);
    return NIQ_OK;
}
