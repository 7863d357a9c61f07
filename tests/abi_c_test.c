/* abi_c_test.c -- the C ABI of include/niq.h exercised from plain C (no Python struct mirrors): builds an MLP handle from a
 * raw weight blob, classifies boxes (affine_fixed) and casts rays, and prints the results as text for the calling test
 * (tests/test_gpu_parity.py::test_c_abi_program) to compare bit for bit with what the ctypes binding returns.
 *
 *   blob (all little-endian): int32 n_dense; per dense layer: int32 in, int32 out, float32 A[in*out], float32 b[out]
 *                             (relu between the dense layers, squeeze_last at the end: the fox.npz topology)
 *                             int32 n_boxes; float32 lo[n*3], hi[n*3]; int32 n_rays; float32 roots[n*3], dirs[n*3]
 *   build: gcc -std=c99 -I include tests/abi_c_test.c -L <pkg> -lniq -Wl,-rpath,<pkg> -o abi_c_test
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "niq.h"

#define CHECK(call)                                                                   \
    do {                                                                              \
        int rc_ = (call);                                                             \
        if (rc_ != NIQ_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, niq_last_error()); return 1; } \
    } while (0)

static void* read_n(FILE* f, size_t bytes) {
    void* p = malloc(bytes ? bytes : 1);
    if (fread(p, 1, bytes, f) != bytes) { fprintf(stderr, "short read\n"); exit(2); }
    return p;
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s blob\n", argv[0]); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror("blob"); return 2; }
    int32_t n_dense = 0;
    if (fread(&n_dense, 4, 1, f) != 1 || n_dense < 1 || n_dense > 16) return 2;
    niq_op_desc ops[64];
    int n_ops = 0;
    memset(ops, 0, sizeof(ops));
    for (int l = 0; l < n_dense; ++l) {
        int32_t dims[2];
        if (fread(dims, 4, 2, f) != 2) return 2;
        ops[n_ops].kind = NIQ_OP_DENSE; ops[n_ops].in_dim = dims[0]; ops[n_ops].out_dim = dims[1];
        ops[n_ops].A = (const float*)read_n(f, (size_t)dims[0] * dims[1] * 4);
        ops[n_ops].b = (const float*)read_n(f, (size_t)dims[1] * 4);
        ++n_ops;
        if (l + 1 < n_dense) ops[n_ops++].kind = NIQ_OP_RELU;
    }
    ops[n_ops++].kind = NIQ_OP_SQUEEZE_LAST;
    int32_t n_boxes = 0, n_rays = 0;
    if (fread(&n_boxes, 4, 1, f) != 1) return 2;
    float* lo = (float*)read_n(f, (size_t)n_boxes * 12);
    float* hi = (float*)read_n(f, (size_t)n_boxes * 12);
    if (fread(&n_rays, 4, 1, f) != 1) return 2;
    float* roots = (float*)read_n(f, (size_t)n_rays * 12);
    float* dirs = (float*)read_n(f, (size_t)n_rays * 12);
    fclose(f);

    niq_ctx* ctx = NULL;
    niq_mlp* mlp = NULL;
    CHECK(niq_ctx_create(0, &ctx));
    CHECK(niq_mlp_create(ctx, n_ops, ops, &mlp));
    int64_t macs = 0;
    CHECK(niq_mlp_macs(mlp, &macs));
    printf("macs %lld\n", (long long)macs);

    niq_mode_cfg cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.mode = NIQ_MODE_AFFINE_FIXED;
    int32_t* label = (int32_t*)malloc((size_t)n_boxes * 4);
    float* lower = (float*)malloc((size_t)n_boxes * 4);
    float* upper = (float*)malloc((size_t)n_boxes * 4);
    uint8_t* tie = (uint8_t*)malloc((size_t)n_boxes);
    CHECK(niq_classify_boxes(ctx, mlp, &cfg, n_boxes, lo, hi, 0.f, label, lower, upper, tie, NIQ_MEM_HOST));
    for (int i = 0; i < n_boxes; ++i) {
        uint32_t a, b;
        memcpy(&a, &lower[i], 4); memcpy(&b, &upper[i], 4);
        printf("box %d %d %08x %08x %d\n", i, label[i], a, b, (int)tie[i]);
    }

    niq_cast_opts o;
    o.hit_eps = 0.001f; o.max_dist = 10.f; o.n_max_step = 512; o.n_substeps = 1; o.safety_factor = 0.98f;
    o.interval_grow_fac = 1.5f; o.interval_shrink_fac = 0.5f; o.interval_init_size = 0.1f;      /* src/queries.py:23-36 */
    float* t = (float*)malloc((size_t)n_rays * 4);
    int32_t* hit = (int32_t*)malloc((size_t)n_rays * 4);
    int32_t* cnt = (int32_t*)malloc((size_t)n_rays * 4);
    int64_t n_evals = 0;
    const niq_mlp* mlps[1];
    mlps[0] = mlp;
    CHECK(niq_cast_rays(ctx, 1, mlps, &cfg, &o, n_rays, roots, dirs, t, hit, cnt, &n_evals, NULL, NIQ_MEM_HOST));
    printf("n_evals %lld\n", (long long)n_evals);
    for (int i = 0; i < n_rays; ++i) {
        uint32_t a;
        memcpy(&a, &t[i], 4);
        printf("ray %d %08x %d %d\n", i, a, hit[i], cnt[i]);
    }
    /* error path: a NULL cfg must come back as NIQ_EINVAL with a message, not a crash */
    if (niq_classify_boxes(ctx, mlp, NULL, n_boxes, lo, hi, 0.f, label, NULL, NULL, NULL, NIQ_MEM_HOST) != NIQ_EINVAL) return 3;
    printf("einval %s\n", niq_last_error());
    CHECK(niq_mlp_destroy(mlp));
    CHECK(niq_ctx_destroy(ctx));
    printf("ok\n");
    return 0;
}
