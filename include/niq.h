/*
 * niq.h -- C ABI of the B200-native range-analysis backend for neural implicit queries.
 *
 * The reference (nmwsharp/neural-implicit-queries) has no native / FFI layer: its boundary is the Python
 * function API of its src/ modules, backed by XLA.  This header is the boundary a binding would target instead;
 * every entry point cites the reference function it replaces (paths relative to the reference repo).
 * The Python modules in neural-implicit-queries_b200/ bind it with ctypes and re-expose the reference's
 * module / function names.
 *
 * Conventions
 *   - plain C types only; all arrays are dense, row-major, float32 / int32 / uint8 as stated
 *   - every function returns 0 (NIQ_OK) or a negative niq_status; niq_last_error() gives the message of
 *     the last failure on the calling thread; nothing throws or aborts across the boundary
 *   - `mem` says where the caller's buffers live: NIQ_MEM_HOST (copied in/out on the context stream,
 *     synchronous on return) or NIQ_MEM_DEVICE (device pointers of the context's GPU; the call is
 *     enqueued on the context stream and the function returns after the stream has been synchronised
 *     unless stated otherwise)
 *   - one host thread per context; a context owns one CUDA device, one stream and its scratch memory
 *   - there is NO CPU fallback: without a CUDA device niq_ctx_create fails with NIQ_ECUDA
 */
#ifndef NIQ_H
#define NIQ_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum niq_status {
    NIQ_OK = 0,
    NIQ_EINVAL = -1,       /* bad argument / unsupported op sequence (reference: ValueError)      */
    NIQ_ENOMEM = -2,       /* host or device allocation failed                                     */
    NIQ_ECUDA = -3,        /* CUDA runtime error, or no usable device                              */
    NIQ_ECAPACITY = -4,    /* a caller-provided output buffer is too small                         */
    NIQ_EUNSUPPORTED = -5  /* valid in the reference but outside this backend (reference: RuntimeError) */
} niq_status;

typedef enum niq_mem { NIQ_MEM_HOST = 0, NIQ_MEM_DEVICE = 1 } niq_mem;

/* SIGN_* of src/implicit_function.py:11-13 */
enum { NIQ_SIGN_UNKNOWN = 0, NIQ_SIGN_POSITIVE = 1, NIQ_SIGN_NEGATIVE = 2 };

/* ops of the MLP dict format, src/mlp.py:14-24 ("%04d.<op>.<arg>" keys) */
typedef enum niq_op_kind {
    NIQ_OP_DENSE = 0,        /* src/mlp.py:218-258, src/affine_layers.py:11-31 : A (in,out), optional b (out) */
    NIQ_OP_RELU = 1,         /* src/mlp.py:283-287, src/affine_layers.py:34-56                              */
    NIQ_OP_ELU = 2,          /* src/mlp.py:289-293, src/affine_layers.py:59-97                              */
    NIQ_OP_SQUEEZE_LAST = 3, /* src/mlp.py:328-332, src/affine_layers.py:164-172                            */
    NIQ_OP_SPATIAL = 4,      /* src/mlp.py:335-347, src/affine_layers.py:175-179 : A = R (3,3), b = t (3)    */
    NIQ_OP_SIN = 5,          /* src/mlp.py:296-300, src/affine_layers.py:100-137, src/slope_interval_layers.py:85-110 */
    NIQ_OP_POW2_ENCODE = 6,  /* src/mlp.py:304-322, src/affine_layers.py:140-161: in_dim = 3, out_dim = #coefs (per input
                                coordinate), A = coefs (out_dim), b = shift (out_dim) or NULL; must be the first op         */
    NIQ_OP_TANH = 7          /* OURS, PARITY UNPINNED: the reference's README names TanH MLPs but its code registers no tanh op
                                or rule; Chebyshev-style linearisation written from the paper's construction
                                (csrc/niq_engine.cuh tanh_lin = oracle/niq_oracle/net.py _tanh_coeffs), key "<i>.tanh._"     */
} niq_op_kind;

typedef struct niq_op_desc {
    int32_t kind;      /* niq_op_kind */
    int32_t in_dim;    /* dense: rows of A; spatial: 3; otherwise ignored */
    int32_t out_dim;   /* dense: cols of A; spatial: 3; otherwise ignored */
    const float* A;    /* HOST pointer, row-major (in_dim, out_dim); spatial: R */
    const float* b;    /* HOST pointer (out_dim) or NULL; spatial: t */
} niq_op_desc;

/* modes of src/affine.py:62-76 (AffineContext) selectable in src/implicit_mlp_utils.py:24-57 */
typedef enum niq_mode {
    NIQ_MODE_INTERVAL = 0,
    NIQ_MODE_AFFINE_FIXED = 1,
    NIQ_MODE_AFFINE_TRUNCATE = 2,
    NIQ_MODE_AFFINE_ALL = 3,
    NIQ_MODE_AFFINE_APPEND = 4,  /* src/affine.py:183-191: per activation keep the n_append largest deltas as new terms */
    NIQ_MODE_SDF = 5,            /* src/sdf.py:17-50 WeakSDFImplicitFunction: f(centre) against lipschitz * box radius */
    NIQ_MODE_SLOPE_INTERVAL = 6  /* src/slope_interval.py + src/slope_interval_layers.py: primal + slope centre / width */
} niq_mode;

typedef struct niq_mode_cfg {
    int32_t mode;            /* niq_mode */
    int32_t truncate_count;  /* affine_truncate: rows kept (src/affine.py:133); affine_append: n_append; else ignored */
    int32_t truncate_policy; /* 0 = 'absolute' (only policy supported; 'relative' -> NIQ_EUNSUPPORTED) */
    float sdf_lipschitz;     /* NIQ_MODE_SDF: the Lipschitz bound of src/sdf.py:19 (kwarg sdf_lipschitz); else ignored */
} niq_mode_cfg;

/* opts of src/queries.py:23-36 that cast_rays reads */
typedef struct niq_cast_opts {
    float hit_eps;
    float max_dist;
    int32_t n_max_step;
    int32_t n_substeps;
    float safety_factor;
    float interval_grow_fac;
    float interval_shrink_fac;
    float interval_init_size;
} niq_cast_opts;

typedef struct niq_ctx niq_ctx;
typedef struct niq_mlp niq_mlp;
typedef struct niq_tree niq_tree;   /* result of niq_tree_build (variable size; two-phase read-out) */
typedef struct niq_mesh niq_mesh;   /* result of niq_marching_cubes                                  */

const char* niq_last_error(void);
const char* niq_version(void);
/* Host-only helper (no GPU): 128-bit content fingerprint of a byte range -- the key of the Python layer's MLP-handle
 * cache (the reference looks `params` up afresh on every call, so handles are cached by content, never by identity). */
int niq_fingerprint128(const void* data, int64_t nbytes, uint64_t out[2]);

/* ---- context ------------------------------------------------------------------------------- */
int niq_ctx_create(int device, niq_ctx** out);
int niq_ctx_destroy(niq_ctx* ctx);
int niq_ctx_sync(niq_ctx* ctx);
/* device properties the host side wants to report: [0]=SM count, [1]=cc major, [2]=cc minor, [3]=smem/block optin */
int niq_ctx_device_info(niq_ctx* ctx, int32_t info[4]);
/* kernel launches issued through this context since creation (bench.py's gpu_launches evidence) */
int niq_ctx_launch_count(niq_ctx* ctx, int64_t* out);
/* CUDA-event timing on the context stream: niq_ctx_timer_start(); ...calls...; niq_ctx_timer_stop(&ms) */
int niq_ctx_timer_start(niq_ctx* ctx);
int niq_ctx_timer_stop(niq_ctx* ctx, float* ms);
/* accumulated device time (ms) of the dominant kernel family since the last reset, by CUDA events:
 * which = 0 bound/point network passes (FP32 FMA), 1 compaction / split / triangle write (HBM)        */
int niq_ctx_kernel_ms(niq_ctx* ctx, int which, float* ms, int64_t* launches, int reset);
/* per-launch CUDA-event timing of the kernel families above is off by default (events cost launch latency) */
int niq_ctx_kernel_timing(niq_ctx* ctx, int on);
/* executed-MAC accounting of the network kernels: they skip columns that are exactly zero after a relu layer
 * (exact: fma(0,w,acc) == acc), so executed work < algorithmic work (report both).  `on` switches
 * the device counter on/off for subsequent launches; *macs (optional) returns the row-level multiply-adds counted
 * so far (padded rows / idle slots included); reset != 0 clears it.                                          */
int niq_ctx_exec_macs(niq_ctx* ctx, int on, int64_t* macs, int reset);
/* marching cubes accounting: lattice points actually evaluated / the (2^n+1)^3 per leaf the reference evaluates (points on
 * a face shared by two leaves are evaluated once; the values -- hence the triangles -- are unchanged).  reset != 0 clears. */
int niq_ctx_mc_points(niq_ctx* ctx, int64_t* evaluated, int64_t* lattice, int reset);
/* device memory helpers so a host language can keep inputs resident (bench `value` leg)              */
int niq_dev_alloc(niq_ctx* ctx, int64_t bytes, void** out);
int niq_dev_free(niq_ctx* ctx, void* p);
int niq_dev_upload(niq_ctx* ctx, void* dst, const void* src, int64_t bytes);
int niq_dev_download(niq_ctx* ctx, void* dst, const void* src, int64_t bytes);
/* measured FP32 FFMA peak of this device (TFLOP/s): a register-only FFMA kernel, CUDA-event timed.  */
int niq_measure_fp32_peak(niq_ctx* ctx, float* tflops);

/* ---- MLP handle: replaces src/mlp.py:96-144 (op-list interpreter) + :173-185 (load) ---------- */
/* Supported op sequences: an optional spatial_transformation, an optional pow2_frequency_encode on the 3-D input, then
 * dense ops each optionally followed by ONE of relu / elu / sin / tanh, with an optional trailing squeeze_last (requires
 * out_dim 1; no activation after the last dense).  Input dimension 3, hidden widths <= 256.  Weights are copied and
 * packed once.                                                                                              */
int niq_mlp_create(niq_ctx* ctx, int32_t n_ops, const niq_op_desc* ops, niq_mlp** out);
int niq_mlp_destroy(niq_mlp* mlp);
/* MACs of one row through the net (sum of in*out over dense/spatial layers) -- SURVEY.md 8(d) "M"     */
int niq_mlp_macs(const niq_mlp* mlp, int64_t* macs);
/* relative half-width of the near-tie band this MLP's bounds are flagged with: 1e-5 for relu-only nets, 2e-4
 * for nets with elu (the reference's elu rule, src/affine_layers.py:59-97, is ill-conditioned in float32:
 * delta = |r_upper - r_lower|/2 cancels O(1) terms).  A bound is near-tie iff it lies within
 * rel * (max(|lower|,|upper|) + sum_j|base_j A_j| + |b|) of +-offset.                                      */
int niq_mlp_tie_rel(const niq_mlp* mlp, float* rel);

/* ---- primitives ------------------------------------------------------------------------------ */
/* f(x): src/mlp.py:96-113 with the 'default' rules.  x (n,3) -> f (n).  `scale` (n) optional (NULL): the
 * magnitude sum_j|h_j A_j|+|b| of the last dot product, used for near-tie bands of sign tests.        */
int niq_eval_points(niq_ctx* ctx, const niq_mlp* mlp, int64_t n, const float* x, float* f, float* scale,
                    int mem);

/* classify_general_box: src/affine.py:34-55 (+ :109-125).  center (n,3), vecs (n,v,3), v in 1..3 for the
 * fixed-row modes (interval / affine_fixed), any v >= 1 for affine_all / affine_truncate.
 * Outputs (each optional, NULL to skip): label (n) int32 SIGN_*, lower/upper (n) float32 bounds,
 * near_tie (n) uint8 = bound within the band of niq_mlp_tie_rel of +-offset.                             */
int niq_classify_general_boxes(niq_ctx* ctx, const niq_mlp* mlp, const niq_mode_cfg* cfg, int64_t n,
                               const float* center, const float* vecs, int32_t v, float offset,
                               int32_t* label, float* lower, float* upper, uint8_t* near_tie, int mem);

/* classify_box: src/implicit_function.py:28-37 -- axis-aligned boxes lo/hi (n,3).                     */
int niq_classify_boxes(niq_ctx* ctx, const niq_mlp* mlp, const niq_mode_cfg* cfg, int64_t n,
                       const float* lo, const float* hi, float offset,
                       int32_t* label, float* lower, float* upper, uint8_t* near_tie, int mem);
/* The slope-interval form of f over n general boxes (src/slope_interval.py:15-33 `slope_interval_func` applied to
 * coordinates_in_general_box(center, vecs), :181-194): raw (n,7) = [primal, slope centre x3, slope width x3] (unused
 * vectors 0), scale (n) or NULL = the magnitude the primal was summed from.  The reference's min_distance_to_zero /
 * min_distance_to_zero_in_direction (:52-163) are closed-form in these numbers (slope_interval.py of this package). */
int niq_slope_forward(niq_ctx* ctx, const niq_mlp* mlp, int64_t n, const float* center, const float* vecs, int32_t v,
                      float* raw, float* scale, int mem);

/* ---- cast_rays: src/queries.py:39-175 --------------------------------------------------------- */
/* roots, dirs (n,3) -> t (n) f32, hit_id (n) i32, count (n) i32.  n_evals = the reference's N_evals
 * (lanes evaluated INCLUDING bucket padding, src/queries.py:137,164), reconstructed from the per-ray
 * counts.  near_tie (n) uint8 optional: some decision of that ray was inside the 1e-5 band.
 * All funcs must use the same mode family; n_funcs >= 1.                                               */
int niq_cast_rays(niq_ctx* ctx, int32_t n_funcs, const niq_mlp* const* mlps, const niq_mode_cfg* cfgs,
                  const niq_cast_opts* opts, int64_t n, const float* roots, const float* dirs,
                  float* t, int32_t* hit_id, int32_t* count, int64_t* n_evals, uint8_t* near_tie, int mem);

/* ---- cast_rays_frustum: src/queries.py:178-587 -------------------------------------------------- */
/* cam_params of src/queries.py:197 (root_pos, look_dir, up_dir, left_dir, fov_x, fov_y, res_x, res_y).  The caller
 * evaluates the transcendental constants once in float32: tan(deg2rad(fov)/2) (src/render.py:17-24) and
 * deg2rad(fov)/2 (src/queries.py:352-353).                                                              */
typedef struct niq_camera {
    float root[3], look[3], up[3], left[3];
    float tan_half_fov_x, tan_half_fov_y;
    float half_fov_x, half_fov_y;
    int32_t res_x, res_y;
} niq_camera;

/* Frusta of pixels [x0, y0, x1, y1) marched together and split when they grow wider than refine_width_fac * step or
 * stall (src/queries.py:350-360); init_ranges (n_init,4) int32 = the initial tiles (src/queries.py:495-501).
 * Outputs are (res_x, res_y) row-major images as the reference returns them (pixel (x,y) at x*res_y + y):
 * t f32, hit_id i32, count i32 (the truncated fractional step count), near_tie u8 optional.  n_evals = the
 * reference's N_evals (padded array length of every marching iteration, src/queries.py:523), replayed from per-iteration
 * termination / split counts.  iter_counts (HOST, optional): those counts, int64[2 * (n_max_step / n_substeps + 3)] =
 * terminated per iteration, then split per iteration (a sharded caller sums them over the ranks before the replay).
 * Pixels outside the initial tiles are left zero.  interval, affine_fixed and slope_interval modes (one persistent
 * kernel with a device work queue).                                                                                    */
int niq_cast_rays_frustum(niq_ctx* ctx, int32_t n_funcs, const niq_mlp* const* mlps, const niq_mode_cfg* cfgs,
                          const niq_cast_opts* opts, const niq_camera* cam, float refine_width_fac,
                          int64_t n_init, const int32_t* init_ranges,
                          float* t, int32_t* hit_id, int32_t* count, int64_t* n_evals, int64_t* iter_counts,
                          uint8_t* near_tie, int mem);

/* ---- level-set kd-tree: src/kd_tree.py:19-218 -------------------------------------------------- */
enum { NIQ_TREE_INTERIOR = 1, NIQ_TREE_EXTERIOR = 2 };  /* flags: also collect NEGATIVE / POSITIVE nodes */
/* split_depth < 0 means "None"; node_terminate_thresh <= 0 means "None" (at least one must be given).
 * Node ORDER follows the reference (per batch of batch_process_size nodes: A-children then B-children). */
int niq_tree_build(niq_ctx* ctx, const niq_mlp* mlp, const niq_mode_cfg* cfg, const float lower[3],
                   const float upper[3], int32_t split_depth, int64_t node_terminate_thresh, float offset,
                   int32_t flags, int32_t batch_process_size, niq_tree** out);
/* The same tree grown from n_roots root boxes lower/upper (n_roots,3) at once (ours: the unit of the multi-GPU
 * partition -- a rank refines all frontier boxes it was dealt in ONE level-synchronous build; n_roots = 1 is
 * niq_tree_build).                                                                                     */
int niq_tree_build_roots(niq_ctx* ctx, const niq_mlp* mlp, const niq_mode_cfg* cfg, int64_t n_roots, const float* lower,
                         const float* upper, int32_t split_depth, int64_t node_terminate_thresh, float offset,
                         int32_t flags, int32_t batch_process_size, niq_tree** out);
/* Ours (multi-GPU subtree partition, no reference counterpart: the reference is single-device): rank `rank` of `world`
 * builds the levels above deal_depth like niq_tree_build, keeps every world-th node of the frontier entering level
 * deal_depth (i = rank mod world) and refines those to split_depth, all in one launch.  The union of the ranks' UNKNOWN
 * leaves is the leaf set of niq_tree_build(split_depth); leaf order is per rank.  Modes: interval, affine_fixed,
 * slope_interval (NIQ_EUNSUPPORTED otherwise: deal on the host and use niq_tree_build_roots).               */
int niq_tree_build_dealt(niq_ctx* ctx, const niq_mlp* mlp, const niq_mode_cfg* cfg, const float lower[3],
                         const float upper[3], int32_t split_depth, float offset, int32_t batch_process_size,
                         int32_t deal_depth, int32_t rank, int32_t world, niq_tree** out);
/* which: 0 unknown leaves, 1 interior (NEGATIVE) nodes, 2 exterior (POSITIVE) nodes                    */
int niq_tree_count(const niq_tree* tree, int which, int64_t* n);
int niq_tree_copy(const niq_tree* tree, int which, float* lower, float* upper, int64_t capacity, int mem);
/* stats[0]=boxes classified, [1]=near-tie boxes, [2]=levels processed, [3]=max frontier                */
int niq_tree_stats(const niq_tree* tree, int64_t stats[4]);
/* per level (0 .. stats[2]-1): info[0]=nodes entering the level, [1]=UNKNOWN, [2]=NEGATIVE, [3]=POSITIVE.
 * The host side replays the reference's array-growth rule (src/kd_tree.py:156-164) from these to return
 * interior/exterior arrays of the reference's padded sizes.                                              */
int niq_tree_level_info(const niq_tree* tree, int32_t level, int64_t info[4]);
int niq_tree_destroy(niq_tree* tree);

/* ---- hierarchical marching cubes: src/kd_tree.py:338-399 + src/extract_cell.py:314-421 ---------- */
/* leaves lo/hi (n,3) (e.g. from niq_tree_copy(which=0)); triangles come out ordered node, subcell, slot. */
int niq_marching_cubes(niq_ctx* ctx, const niq_mlp* mlp, int64_t n, const float* leaf_lower,
                       const float* leaf_upper, int32_t n_subcell_depth, int mem, niq_mesh** out);
int niq_marching_cubes_tree(niq_ctx* ctx, const niq_mlp* mlp, const niq_tree* tree, int32_t n_subcell_depth,
                            niq_mesh** out);
int niq_mesh_count(const niq_mesh* mesh, int64_t* n_tris);
int niq_mesh_copy(const niq_mesh* mesh, float* tri_pos /* (n,3,3) */, int64_t capacity, int mem);
int niq_mesh_destroy(niq_mesh* mesh);
/* the case tables of src/extract_cell.py:12-302: tri (256,16) i32, edge_verts (12,2) i32, vert_coords (8,3) u8 */
int niq_mc_tables(int32_t* tri_table, int32_t* edge_verts, uint8_t* vert_coords);

/* ---- find_any_intersection: src/kd_tree.py:402-655 ---------------------------------------------- */
/* found (1) i32, loc (3) f32 (= -777 when not found); stats[0]=nodes processed, [1]=rounds, [2]=near-tie */
int niq_find_any_intersection(niq_ctx* ctx, const niq_mlp* mlpA, const niq_mode_cfg* cfgA,
                              const niq_mlp* mlpB, const niq_mode_cfg* cfgB, const float lower[3],
                              const float upper[3], float eps, int32_t* found, float loc[3], int64_t stats[3]);

/* A BATCH of pairwise queries that differ only in the rigid transforms of the two shapes (the reference's GUI re-runs the
 * query whenever a gizmo moves, src/main_intersection.py:171-183; SURVEY.md 8(d) config 3 is a list of such transforms):
 * xfA / xfB: HOST (n_queries, 12) float32 = R (3x3 row-major) then t (3) per query -- the values the reference would store in
 * params["0000.spatial_transformation.R" / ".t"] -- or NULL to use the handle's own first op.  A handle that receives
 * transforms must have been created from params whose FIRST op is a spatial_transformation.  All queries run in ONE persistent
 * cooperative kernel (every round processes the frontiers of all live queries); each query's verdict, location and
 * statistics equal those of its own niq_find_any_intersection call.  found (n) i32, loc (n,3) f32 (-777 when not found),
 * stats (n,3) i64 optional.  affine_truncate / affine_all / affine_append; other modes -> NIQ_EUNSUPPORTED.             */
int niq_find_any_intersection_batch(niq_ctx* ctx, const niq_mlp* mlpA, const niq_mode_cfg* cfgA,
                                    const niq_mlp* mlpB, const niq_mode_cfg* cfgB, int64_t n_queries,
                                    const float* xfA, const float* xfB, const float lower[3], const float upper[3],
                                    float eps, int32_t* found, float* loc, int64_t* stats);

/* ---- closest_point: src/kd_tree.py:659-802 ------------------------------------------------------- */
/* query_points (q,3) -> dist (q) (inf if no surface found), loc (q,3) (contract only where dist is finite).
 * batch_process_size = the reference's global LIFO window (results depend on it).
 * stats[0]=rounds, [1]=node visits, [2]=max stack, [3]=near-tie boxes                                  */
int niq_closest_point(niq_ctx* ctx, const niq_mlp* mlp, const niq_mode_cfg* cfg, const float lower[3],
                      const float upper[3], int64_t q, const float* query_points, float eps,
                      int64_t batch_process_size, float* dist, float* loc, int64_t stats[4], int mem);

#ifdef __cplusplus
}
#endif
#endif /* NIQ_H */
