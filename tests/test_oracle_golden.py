"""Pins the CPU oracle (oracle/niq_oracle) to the golden vectors produced by the UNMODIFIED reference
sources run on the NumPy `jax` stand-in (oracle/tools/gen_golden.py).  CPU only.

Tolerance (BASELINE.json north_star): labels / topology / verdicts exact outside a counted band of
boxes whose bound lies within 1e-5 (relative) of the level set; values within 1e-5 relative.
"""
import numpy as np
import pytest

from conftest import golden, sample_params
from niq_oracle import net, rays, tree

RTOL = 1e-5
SAMPLES = ("fox", "bunny", "hammer", "birdcage_occ")
MODES = ("interval", "affine_fixed", "affine_truncate", "affine_all", "affine_append")


def ctx_for(mode, n_trunc):
    if mode == "sdf":                      # tree fixtures store the Lipschitz bound under the same key
        return net.AffineContext(mode, sdf_lipschitz=float(n_trunc))
    if mode == "affine_append":            # the golden files store the count under the same key
        return net.AffineContext(mode, n_append=int(n_trunc))
    return net.AffineContext(mode, truncate_count=int(n_trunc))


def assert_bounds_close(lo, up, glo, gup, sc, rtol=RTOL):
    scale = net.tol_scale(glo, gup, sc)
    assert np.all(np.abs(lo.astype(np.float64) - glo) <= rtol * scale + 1e-30)
    assert np.all(np.abs(up.astype(np.float64) - gup) <= rtol * scale + 1e-30)


def assert_labels(lab, glab, lo, up, sc, offset=0.0):
    tie = net.bound_near_tie(lo, up, offset, sc)
    bad = (lab != glab) & ~tie
    assert not bad.any(), f"{bad.sum()} label mismatches outside the near-tie band"


@pytest.mark.parametrize("name", SAMPLES)
def test_point_values(name):
    g = golden(f"points_{name}")
    f = net.eval_points(sample_params(name), g["points"])
    s = rays.point_scale(sample_params(name), g["points"])
    assert np.all(np.abs(f - g["values"]) <= RTOL * s)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("name", SAMPLES)
def test_classify(name, mode):
    g = golden(f"classify_{name}_{mode}")
    p = sample_params(name)
    ctx = ctx_for(mode, g["n_trunc"])
    rtol = net.mode_rel(p, ctx) if mode == "affine_append" else RTOL      # see net.mode_rel: ill-conditioned reference formula
    lab, lo, up, sc = net.classify_box(p, ctx, g["box_lower"], g["box_upper"], return_scale=True)
    assert_bounds_close(lo, up, g["lower"], g["upper"], sc, rtol)
    assert_labels(lab, g["label"], lo, up, sc)
    assert_labels(net.labels_from_bounds(lo, up, 0.05), g["label_offset005"], lo, up, sc, 0.05)
    # ray segments (v = 1)
    lab, lo, up, sc = net.classify_general_box(p, ctx, g["seg_center"], g["seg_vecs"], return_scale=True)
    assert_bounds_close(lo, up, g["seg_lower"], g["seg_upper"], sc, rtol)
    assert_labels(lab, g["seg_label"], lo, up, sc)
    # rigid transform prepended
    p2 = net.prepend_op(p, net.spatial_transformation(g["xf_R"], g["xf_t"]))
    lab, lo, up, sc = net.classify_box(p2, ctx, g["box_lower"][9:18], g["box_upper"][9:18], return_scale=True)
    assert_bounds_close(lo, up, g["xf_lower"], g["xf_upper"], sc, rtol)
    assert_labels(lab, g["xf_label"], lo, up, sc)
    f = net.eval_points(p2, g["xf_points"])
    assert np.all(np.abs(f - g["xf_values"]) <= RTOL * rays.point_scale(p2, g["xf_points"]))


def test_append_is_ill_conditioned_in_float32():
    """Evidence for net.mode_rel: float32 vs float64 evaluation of the oracle's own formulas.  affine_append loses
    1.5-2.5 more digits than affine_fixed (the reference's `sum(delta) - sum(kept)` cancellation, src/affine.py:191)."""
    rng = np.random.default_rng(13)
    n = 600
    c = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    h = (2.0 ** rng.uniform(-9, 0, (n, 1)) * rng.uniform(0.5, 1.0, (n, 3))).astype(np.float32)
    for name in ("fox", "birdcage_occ"):
        p = sample_params(name)
        worst = {}
        for mode, kw in (("affine_append", dict(n_append=4)), ("affine_fixed", {})):
            ctx = net.AffineContext(mode, **kw)
            _, lo, up, _ = net.classify_box(p, ctx, c - h, c + h, return_scale=True)
            with net.precision(np.float64):
                _, lo6, up6, sc6 = net.classify_box(p, ctx, c - h, c + h, return_scale=True)
            s = net.tol_scale(lo6, up6, sc6)
            worst[mode] = float(np.max(np.maximum(np.abs(lo - lo6), np.abs(up - up6)) / s))
        assert worst["affine_fixed"] < 1e-5 < 10 * worst["affine_fixed"] < worst["affine_append"] < net.mode_rel(p, net.AffineContext("affine_append", n_append=4))


@pytest.mark.parametrize("name", SAMPLES)
def test_classify_sdf(name):
    """sdf.WeakSDFImplicitFunction (src/sdf.py:31-50).  Labels are exact unless |f(centre)| is within
    the point-value band of lipschitz*radius or of the offset."""
    g = golden(f"classify_{name}_sdf")
    p = sample_params(name)
    ctx = net.AffineContext("sdf", sdf_lipschitz=float(g["lipschitz"]))

    def check(lab, glab, center, vecs, offset):
        val, reach = net.sdf_center_value_and_reach(p, ctx, center, vecs)
        band = RTOL * (rays.point_scale(p, center) + reach)
        tie = (np.abs(np.abs(val) - reach) <= band) | (np.abs(val - offset) <= band) | (np.abs(val + offset) <= band)
        assert np.all((lab == glab) | tie)
        assert tie.mean() < 0.1

    c, v = net.box_to_general(g["box_lower"], g["box_upper"])
    check(net.classify_box(p, ctx, g["box_lower"], g["box_upper"]), g["label"], c, v, 0.0)
    check(net.classify_box(p, ctx, g["box_lower"], g["box_upper"], offset=0.05), g["label_offset005"], c, v, 0.05)
    check(net.classify_general_box(p, ctx, g["gen_center"], g["gen_vecs"]), g["gen_label"], g["gen_center"], g["gen_vecs"], 0.0)
    check(net.classify_general_box(p, ctx, g["gen_center"], g["gen_vecs"][:, :1]), g["gen_label_v1"], g["gen_center"], g["gen_vecs"][:, :1], 0.0)
    assert name == "hammer" or len(np.unique(g["label"])) >= 2      # (hammer's values dwarf L*radius: all POSITIVE)


@pytest.mark.parametrize("name", SAMPLES)
def test_classify_slope_interval(name):
    """the slope_interval bounder (src/slope_interval.py:29-50 with the rules of
    src/slope_interval_layers.py) -- labels and may-contain bounds of axis-aligned boxes, v=2 / v=1 general boxes and
    a prepended rigid transform, against the run of the unmodified reference."""
    g = golden(f"classify_{name}_slope_interval")
    p = sample_params(name)
    ctx = net.AffineContext("slope_interval")
    rtol = net.tie_rel(p)
    lab, lo, up, sc = net.classify_box(p, ctx, g["box_lower"], g["box_upper"], return_scale=True)
    assert_bounds_close(lo, up, g["lower"], g["upper"], sc, rtol)
    assert_labels(lab, g["label"], lo, up, sc)
    assert_labels(net.labels_from_bounds(lo, up, 0.05), g["label_offset005"], lo, up, sc, 0.05)
    for tag, vecs in (("gen", g["gen_vecs"]), ("gen1", g["gen_vecs"][:, :1])):
        lab, lo, up, sc = net.classify_general_box(p, ctx, g["gen_center"], vecs, return_scale=True)
        assert_bounds_close(lo, up, g[f"{tag}_lower"], g[f"{tag}_upper"], sc, rtol)
        assert_labels(lab, g[f"{tag}_label"], lo, up, sc)
    p2 = net.prepend_op(p, net.spatial_transformation(g["xf_R"], g["xf_t"]))
    lab, lo, up, sc = net.classify_box(p2, ctx, g["box_lower"][9:18], g["box_upper"][9:18], return_scale=True)
    assert_bounds_close(lo, up, g["xf_lower"], g["xf_upper"], sc, rtol)
    assert_labels(lab, g["xf_label"], lo, up, sc)


def check_min_distance(p, g, box_fn, dir_fn, rtol, report=None):
    """Shared by the oracle (here) and the CUDA backend (tests/test_gpu_parity.py): the distance helpers of the slope-interval
    bounder against the goldens of the unmodified reference.  A source value must lie within the band of the magnitude it was
    summed from; a distance is value / slope, so its error is the value's band over the slope = distance * scale / |value|
    (+ the slope's own band): checked with a factor 4 of head room.  "Contains zero" verdicts must agree unless a source bound
    lies within the band of 0."""
    f64 = np.float64

    def close(a, b, tol):
        a, b = np.asarray(a, f64), np.asarray(b, f64)
        return bool(np.all((np.isnan(a) == np.isnan(b)) & (np.isnan(a) | (np.abs(a - b) <= tol))))

    worst = {}
    prim, dist = box_fn(p, g["box_center"], g["box_axis_vec"])
    sc = rays.point_scale(p, g["box_center"]).astype(f64)
    assert close(prim, g["box_primal"], rtol * sc)
    cap = np.abs(g["box_axis_vec"]).min(axis=1)
    gd = np.abs(g["box_distance"]).astype(f64)
    assert close(dist, g["box_distance"], 4 * rtol * gd * (2 + sc / np.maximum(np.abs(g["box_primal"]), 1e-30)))
    assert np.all(np.asarray(dist) <= cap) and np.all(np.asarray(dist) >= 0)
    worst["box"] = float(np.max(np.abs(np.asarray(dist, f64) - g["box_distance"]) / np.maximum(gd, 1e-30)))
    val, d = dir_fn(p, g["source"], g["bound_vec"], None)
    ssc = rays.point_scale(p, g["source"]).astype(f64)
    assert close(val, g["ray_value"], rtol * ssc)
    gd = np.abs(g["ray_distance"]).astype(f64)
    assert close(d, g["ray_distance"], 4 * rtol * gd * (2 + ssc / np.maximum(np.abs(g["ray_value"]), 1e-30)))
    assert np.all(np.asarray(d, f64) <= np.linalg.norm(g["bound_vec"].astype(f64), axis=1) * (1 + 1e-6))
    worst["ray"] = float(np.max(np.abs(np.asarray(d, f64) - g["ray_distance"]) / np.maximum(gd, 1e-30)))
    for k, key in ((1, "swept1"), (2, "swept2")):
        lo, up, d = dir_fn(p, g["source"], g["bound_vec"], g["source_range"][:, :k])
        G = g[key].astype(f64)
        bsc = net.tol_scale(G[:, 0].astype(np.float32), G[:, 1].astype(np.float32), ssc.astype(np.float32)).astype(f64)
        assert close(lo, G[:, 0], rtol * bsc) and close(up, G[:, 1], rtol * bsc)
        zero = (G[:, 0] <= 0) & (G[:, 1] >= 0)
        v = np.minimum(np.abs(G[:, 0]), np.abs(G[:, 1]))
        near = v <= rtol * bsc
        assert np.array_equal((np.asarray(d) == 0)[~near], (G[:, 2] == 0)[~near])
        sel = ~zero & ~near
        assert close(np.asarray(d)[sel], G[sel, 2], 4 * rtol * np.abs(G[sel, 2]) * (2 + bsc[sel] / np.maximum(v[sel], 1e-30)))
        worst[key] = float(np.max(np.abs(np.asarray(d, f64)[sel] - G[sel, 2]) / np.maximum(np.abs(G[sel, 2]), 1e-30))) if sel.any() else 0.0
    if report is not None:
        report.update(worst)
    return worst


@pytest.mark.parametrize("name", ["fox", "bunny"])
def test_slope_min_distance_helpers(name):
    """The distance helpers of the slope-interval bounder (src/slope_interval.py:52-163: min_distance_to_zero for axis-aligned
    boxes, min_distance_to_zero_in_direction for a ray and for a swept box with 1 / 2 source-range vectors): the oracle's
    restatement against the run of the unmodified reference."""
    p = sample_params(name)
    check_min_distance(p, golden(f"mindist_{name}_slope"), net.slope_min_distance_to_zero,
                       lambda pp, s, b, r: net.slope_min_distance_to_zero_in_direction(pp, s, b, r), net.tie_rel(p))


PE_MODES = ("interval", "affine_fixed", "affine_truncate", "affine_all", "slope_interval")


def pe_params(g):
    return {k.split("/", 1)[1]: g[k] for k in g if k.startswith("params/")}


@pytest.mark.parametrize("mode", PE_MODES)
def test_positional_encoding_ops(mode):
    """the sin and pow2_frequency_encode ops (src/mlp.py:296-322, src/affine_layers.py:100-161,
    src/slope_interval_layers.py:85-126) on a positional-encoding MLP, against the run of the unmodified reference."""
    g = golden(f"pe_{mode}")
    p = pe_params(g)
    ctx = ctx_for(mode, g["n_trunc"])
    lab, lo, up, sc = net.classify_box(p, ctx, g["box_lower"], g["box_upper"], return_scale=True)
    assert_bounds_close(lo, up, g["lower"], g["upper"], sc)
    assert_labels(lab, g["label"], lo, up, sc)
    f = net.eval_points(p, g["points"])
    assert np.all(np.abs(f - g["values"]) <= RTOL * rays.point_scale(p, g["points"]))
    assert len(np.unique(g["label"])) == 3


def test_classify_truncate64():
    g = golden("classify_hammer_affine_truncate64")
    ctx = ctx_for("affine_truncate", g["n_trunc"])
    lab, lo, up, sc = net.classify_box(sample_params("hammer"), ctx, g["box_lower"], g["box_upper"], return_scale=True)
    assert_bounds_close(lo, up, g["lower"], g["upper"], sc)
    assert_labels(lab, g["label"], lo, up, sc)


RAY_CASES = {
    "rays_fox_fixed_r12": (("fox",), "affine_fixed"),
    "rays_fox_interval_r6": (("fox",), "interval"),
    "rays_fox_all_r6": (("fox",), "affine_all"),
    "rays_fox_fixed_r8_sub3": (("fox",), "affine_fixed"),
    "rays_fox_bunny_fixed_r8": (("fox", "bunny"), "affine_fixed"),
}


@pytest.mark.parametrize("case", sorted(RAY_CASES))
def test_cast_rays(case):
    names, mode = RAY_CASES[case]
    g = golden(case)
    # the camera generator restatement reproduces the reference's rays
    look, up, _ = rays.look_at(g["eye"])
    roots, dirs = rays.generate_camera_rays(g["eye"], look, up, res=int(g["res"]), fov_deg=30.0)
    np.testing.assert_allclose(dirs, g["dirs"], rtol=0, atol=2e-7)
    np.testing.assert_array_equal(roots, g["roots"])

    opts = rays.get_default_cast_opts()
    opts["n_substeps"] = int(g["n_substeps"])
    ctxs = tuple(ctx_for(mode, g["n_trunc"]) for _ in names)
    ps = tuple(sample_params(n) for n in names)
    t, hit, cnt, n_evals, tie = rays.cast_rays(ctxs, ps, g["roots"], g["dirs"], opts, return_near_tie=True)
    ok = ~tie
    assert ok.mean() > 0.9
    np.testing.assert_array_equal(hit[ok], g["out_hit_id"][ok])
    np.testing.assert_array_equal(cnt[ok], g["out_count"][ok])
    np.testing.assert_allclose(t[ok], g["out_t"][ok], rtol=RTOL, atol=0)
    if not tie.any():
        assert n_evals == int(g["n_evals"])


TREE_CASES = {
    "tree_fox_fixed_d12": ("fox", "affine_fixed"),
    "tree_bunny_all_d9": ("bunny", "affine_all"),
    "tree_fox_trunc_d9": ("fox", "affine_truncate"),
    "tree_fox_append_d9": ("fox", "affine_append"),
    "tree_fox_sdf_d12": ("fox", "sdf"),
    "tree_fox_slope_d12": ("fox", "slope_interval"),
    "tree_fox_fixed_thresh": ("fox", "affine_fixed"),
    "tree_fox_fixed_b128": ("fox", "affine_fixed"),
}


@pytest.mark.parametrize("case", sorted(TREE_CASES))
def test_tree(case):
    name, mode = TREE_CASES[case]
    g = golden(case)
    kw = {}
    for k in g:
        if k.startswith("kw_"):
            v = g[k].item()
            kw[k[3:]] = v
    stats = {}
    out = tree.construct_uniform_unknown_levelset_tree(
        ctx_for(mode, g["n_trunc"]), sample_params(name), np.full(3, -1, np.float32), np.full(3, 1, np.float32),
        stats=stats, **kw)
    assert stats["n_near_tie"] == 0, "pick a different golden case: near-tie boxes make topology ambiguous"
    for tag in ("unknown", "interior", "exterior"):
        if f"{tag}_node_valid" not in g:
            continue
        gv = g[f"{tag}_node_valid"]
        v = out[f"{tag}_node_valid"]
        assert v.shape == gv.shape                      # same padded bucket size
        np.testing.assert_array_equal(v, gv)
        np.testing.assert_array_equal(out[f"{tag}_node_lower"][v], g[f"{tag}_node_lower"][gv])   # order too
        np.testing.assert_array_equal(out[f"{tag}_node_upper"][v], g[f"{tag}_node_upper"][gv])


@pytest.mark.parametrize("case,name", [("mc_fox_d4_s2", "fox"), ("mc_bunny_d4_s3", "bunny")])
def test_marching_cubes(case, name):
    g = golden(case)
    tri = tree.hierarchical_marching_cubes(ctx_for("affine_fixed", 0), sample_params(name),
                                           np.full(3, -1, np.float32), np.full(3, 1, np.float32),
                                           int(g["depth"]), n_subcell_depth=int(g["n_sub"]))
    assert tri.shape == g["tri_pos"].shape
    assert tri.shape[0] > 0
    np.testing.assert_allclose(tri, g["tri_pos"], rtol=0, atol=2e-5)


@pytest.mark.parametrize("case,mode", [("isect_fixed", "affine_fixed"), ("isect_trunc64", "affine_truncate")])
def test_find_any_intersection(case, mode):
    g = golden(case)
    pA = sample_params("hammer")
    ctx = ctx_for(mode, g["n_trunc"])
    lo, hi = np.full(3, -1, np.float32), np.full(3, 1, np.float32)
    for i in range(g["R"].shape[0]):
        pB = net.prepend_op(sample_params("bunny"), net.spatial_transformation(g["R"][i], g["t"][i]))
        found, ia, ib, loc = tree.find_any_intersection((ctx, ctx), (pA, pB), lo, hi, float(g["eps"]))
        assert bool(found) == bool(g["found"][i])
        np.testing.assert_allclose(loc, g["loc"][i], rtol=0, atol=1e-6)
        assert (ia, ib) == ((1, 2) if found else (0, 0))


@pytest.mark.parametrize("case", ["closest_fox_B4", "closest_fox_B256"])
def test_closest_point(case):
    g = golden(case)
    d, loc = tree.closest_point(ctx_for("affine_fixed", 0), sample_params("fox"),
                                np.full(3, -1, np.float32), np.full(3, 1, np.float32),
                                g["query_points"], eps=float(g["eps"]), batch_process_size=int(g["B"]))
    np.testing.assert_allclose(d, g["dist"], rtol=RTOL)
    fin = np.isfinite(g["dist"])
    assert fin.any()
    np.testing.assert_allclose(loc[fin], g["loc"][fin], rtol=0, atol=1e-6)


def test_elu_rule_conditioning():
    """Why elu nets are compared at 2e-4 relative (net.tie_rel) and relu nets at 1e-5: the oracle's OWN float32
    evaluation differs from the float64 evaluation of the same formulas by > 1e-5 relative for the elu net
    (bunny) but not for the relu nets -- the elu rule's delta = |r_upper - r_lower|/2 cancels O(1) terms."""
    rng = np.random.default_rng(11)
    n = 4000
    c = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    h = (2.0 ** rng.uniform(-9, 0, (n, 1)) * rng.uniform(0.5, 1.0, (n, 3))).astype(np.float32)
    worst = {}
    for name in ("fox", "bunny", "hammer"):
        p = sample_params(name)
        ctx = ctx_for("affine_fixed", 0)
        _, lo, up, sc = net.classify_box(p, ctx, c - h, c + h, return_scale=True)
        with net.precision(np.float64):
            _, lo6, up6, sc6 = net.classify_box(p, ctx, c - h, c + h, return_scale=True)
        assert lo6.dtype == np.float64
        worst[name] = float((np.maximum(np.abs(lo - lo6), np.abs(up - up6)) / net.tol_scale(lo6, up6, sc6)).max())
    assert worst["fox"] < 1e-5 and worst["hammer"] < 1e-5
    assert 1e-5 < worst["bunny"] < net.NEAR_TIE_REL_ELU
    assert net.tie_rel(sample_params("bunny")) == net.NEAR_TIE_REL_ELU and net.tie_rel(sample_params("fox")) == 1e-5


def test_mc_tables_match_reference_hash():
    import hashlib

    from niq_oracle import mc_tables
    tri, edges, vc = mc_tables.unpack()
    assert hashlib.sha256(tri.astype(np.int8).tobytes()).hexdigest() == mc_tables.TRI_TABLE_SHA256
    assert (tri[:, 15] == -1).all() and tri.max() == 11
    assert edges.shape == (12, 2) and vc.shape == (8, 3)


def test_render_image():
    """The caller of cast_rays (src/render.py:53-165): camera rays -> cast_rays -> finite-difference
    normals -> 'normal' shading.  Normals are differences of 4 point values eps=1e-3 apart, so the colour tolerance
    is the point-value band (1e-5 relative of the evaluation scale) divided by the difference magnitude."""
    g = golden("render_fox_fixed_r10")
    eye = np.array((2., 1., 2.), np.float32)
    look, up, _ = rays.look_at(eye)
    img, depth, cnt, hit, n_eval = rays.render_image((net.AffineContext("affine_fixed"),), (sample_params("fox"),), eye, look, up,
                                                     int(g["res"]), 30.0, rays.get_default_cast_opts())
    np.testing.assert_array_equal(hit, g["hit_ids"])
    np.testing.assert_array_equal(cnt, g["counts"])
    assert n_eval == int(g["n_eval"])
    np.testing.assert_allclose(depth, g["depth"], rtol=RTOL, atol=0)
    assert (g["hit_ids"] != 0).sum() > 10
    np.testing.assert_allclose(img, g["img"], rtol=0, atol=2e-3)
    assert np.all(img[g["hit_ids"] == 0] == 1.0)


def test_render_image_frustum_branch():
    """render_image(frustum=True) of the unmodified reference: the frustum images transposed into ray order, normals from
    the generate_camera_rays directions (src/render.py:116-132)."""
    g = golden("render_frustum_fox_fixed_r14")
    eye = np.array((2., 1., 2.), np.float32)
    look, up, left = rays.look_at(eye)
    opts = rays.get_default_cast_opts()
    opts["n_side_init"] = int(g["n_side"])
    img, depth, cnt, hit, n_eval = rays.render_image((net.AffineContext("affine_fixed"),), (sample_params("fox"),), eye, look, up,
                                                     int(g["res"]), 30.0, opts, frustum=True, left_dir=left)
    np.testing.assert_array_equal(hit, g["hit_ids"])
    np.testing.assert_array_equal(cnt, g["counts"])
    assert n_eval == int(g["n_eval"]) and (g["hit_ids"] != 0).sum() > 10
    np.testing.assert_allclose(depth, g["depth"], rtol=RTOL, atol=0)
    np.testing.assert_allclose(img, g["img"], rtol=0, atol=2e-3)


FRUSTUM_CASES = {
    "frust_fox_fixed_r12_s4": (("fox",), "affine_fixed"),
    "frust_hammer_fixed_r9_s3_sub3": (("hammer",), "affine_fixed"),
    "frust_fox_bunny_interval_r10_s2_sub2": (("fox", "bunny"), "interval"),
    "frust_fox_fixed_r13x9_s3": (("fox",), "affine_fixed"),           # res_x != res_y and fov_x != fov_y
    "frust_fox_slope_r10_s2": (("fox",), "slope_interval"),
    "frust_fox_trunc_r8_s2": (("fox",), "affine_truncate"),
}


def frustum_inputs(g):
    opts = rays.get_default_cast_opts()
    opts["n_side_init"] = int(g["n_side"])
    opts["n_substeps"] = int(g["n_substeps"])
    res = int(g["res"])
    res_y = int(g["res_y"]) if "res_y" in g else res
    fov_y = float(g["fov_y"]) if "fov_y" in g else 30.0
    return (g["eye"], g["look"], g["up"], g["left"], 30.0, fov_y, res, res_y), opts


@pytest.mark.parametrize("case", sorted(FRUSTUM_CASES))
def test_cast_rays_frustum(case):
    """queries.cast_rays_frustum (src/queries.py:178-587): (res_x, res_y) images of t / hit id /
    truncated fractional step count and N_evals, against the unmodified reference."""
    names, mode = FRUSTUM_CASES[case]
    g = golden(case)
    cam, opts = frustum_inputs(g)
    iters = []
    t, hit, cnt, n_evals, tie = rays.cast_rays_frustum(tuple(ctx_for(mode, g.get("n_trunc", 8)) for _ in names), tuple(sample_params(n) for n in names),
                                                       cam, opts, return_near_tie=True, iter_counts=iters)
    ok = ~tie
    assert ok.mean() > 0.6 and t.shape == g["out_t"].shape
    np.testing.assert_array_equal(hit[ok], g["out_hit_id"][ok])
    np.testing.assert_array_equal(cnt[ok], g["out_count"][ok])
    np.testing.assert_allclose(t[ok], g["out_t"][ok], rtol=RTOL, atol=0)
    if mode == "interval":                                   # the loose interval bounds crawl: every frustum ends on the step limit
        assert (g["out_count"] > 100).all()
    else:
        assert (g["out_hit_id"] != 0).any()
    if not tie.any():
        assert n_evals == int(g["n_evals"])
    # host logic of the product (no GPU needed): initial tiles and the N_evals replay from per-iteration counts
    import queries
    n_side = int(g["n_side"])
    init = queries._initial_frusta(cam[6], cam[7], n_side)
    assert init.shape == (n_side * n_side, 4) and init.dtype == np.int32
    assert ((init[:, 2] - init[:, 0]) * (init[:, 3] - init[:, 1])).sum() == cam[6] * cam[7]
    assert queries._frustum_n_evals(n_side * n_side, [a for a, _ in iters], [b for _, b in iters]) == n_evals


def test_initial_frusta_uneven_tiles():
    """src/queries.py:495-501 with a resolution the tile count does not divide: floor of the float32 linspace."""
    import queries
    init = queries._initial_frusta(50, 37, 16)
    xt = np.floor(np.linspace(0, 50, 17, dtype=np.float32)).astype(np.int32)
    yt = np.floor(np.linspace(0, 37, 17, dtype=np.float32)).astype(np.int32)
    np.testing.assert_array_equal(init[:16, 0], xt[:-1])
    np.testing.assert_array_equal(init[:16, 2], xt[1:])
    np.testing.assert_array_equal(init[::16, 1], yt[:-1])
    np.testing.assert_array_equal(init[::16, 3], yt[1:])
    assert ((init[:, 2] - init[:, 0]) * (init[:, 3] - init[:, 1])).sum() == 50 * 37


def test_mc_lattice_coordinates_on_shared_faces_are_bit_identical():
    """The premise of the product's shared-face evaluation (csrc: k_mc_neighbours / mc_val): when the low face of a leaf
    coincides exactly with the high face of another leaf (float-equal bounds), the lattice coordinates the reference's
    linspace formula (src/extract_cell.py:372-381) gives the two leaves on that face are the same bits -- first sample = lo
    exactly, last sample = hi itself, the other two axes from identical inputs -- so evaluating them once changes nothing.
    Checked on the leaves of a real tree (dyadic bounds) and on leaves with non-dyadic bounds."""
    from niq_oracle import mc
    p = sample_params("fox")
    r = tree.construct_uniform_unknown_levelset_tree(net.AffineContext("affine_fixed"), p, np.full(3, -1, np.float32),
                                                     np.full(3, 1, np.float32), split_depth=9)
    v = r["unknown_node_valid"]
    lo, hi = r["unknown_node_lower"][v], r["unknown_node_upper"][v]
    rng = np.random.default_rng(0)
    base = rng.uniform(-1, 1, (40, 3)).astype(np.float32)
    ext = rng.uniform(0.01, 0.3, (40, 3)).astype(np.float32)
    lo2 = np.concatenate((base, base + np.array([1, 0, 0], np.float32) * ext))          # second half: x-neighbours of the first
    hi2 = np.concatenate((base + ext, (base + np.array([1, 0, 0], np.float32) * ext) + ext)).astype(np.float32)
    hi2[:40, 0] = lo2[40:, 0]                                                           # make the shared face exact
    hi2[40:, 1:] = hi2[:40, 1:]
    n_pairs = 0
    for lo_, hi_ in ((lo, hi), (lo2.astype(np.float32), hi2)):
        grid = mc.lattice_points(lo_, hi_, 3)                                           # (L, 9, 9, 9, 3)
        key = {tuple(a.tolist()): i for i, a in enumerate(lo_)}
        for i in range(lo_.shape[0]):
            for d in range(3):
                k = lo_[i].copy()
                k[d] = lo_[i, d] - (hi_[i, d] - lo_[i, d])
                j = key.get(tuple(k.tolist()))
                if j is None or j == i or hi_[j, d] != lo_[i, d] or any(hi_[j, e] != hi_[i, e] for e in range(3) if e != d):
                    continue
                mine = np.take(grid[i], 0, axis=d)                                      # my low face
                theirs = np.take(grid[j], 8, axis=d)                                    # their high face
                assert np.array_equal(mine.view(np.uint32), theirs.view(np.uint32))
                n_pairs += 1
    assert n_pairs > 100
