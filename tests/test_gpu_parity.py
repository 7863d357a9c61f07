"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the reference-shaped
Python modules -> ctypes -> C ABI (include/niq.h), against (a) the committed golden vectors produced by
the unmodified reference on the NumPy jax stand-in and (b) the CPU oracle on seeded inputs.

Tolerance (BASELINE.json north_star): labels / topology / hit flags / verdicts bit-exact except for boxes
whose bound lies within 1e-5 (relative) of the level set -- counted; values within 1e-5 relative, where
"relative" is measured against the magnitude the bound was summed from (net.tol_scale)."""
import numpy as np
import pytest

from conftest import golden, parity_report, sample_params
from niq_oracle import net, rays, tree as otree

pytestmark = pytest.mark.gpu

RTOL = 1e-5
SAMPLES = ("fox", "bunny", "hammer", "birdcage_occ")
MODES = ("interval", "affine_fixed", "affine_truncate", "affine_all")
LO = np.full(3, -1, np.float32)
HI = np.full(3, 1, np.float32)


def make(params, mode, n_trunc=8):
    import implicit_mlp_utils
    kw = dict(affine_n_truncate=int(n_trunc), affine_truncate_policy="absolute") if mode == "affine_truncate" else {}
    if mode == "affine_append":
        kw = dict(affine_n_append=int(n_trunc))
    if mode == "sdf":
        kw = dict(sdf_lipschitz=float(n_trunc))
    return implicit_mlp_utils.generate_implicit_from_params(params, mode, **kw)


def octx(mode, n_trunc=8):
    if mode == "sdf":
        return net.AffineContext(mode, sdf_lipschitz=float(n_trunc))
    if mode == "affine_append":
        return net.AffineContext(mode, n_append=int(n_trunc))
    return net.AffineContext(mode, truncate_count=int(n_trunc))


def check_bounds(lo, up, glo, gup, sc, rel=RTOL):
    """rel = net.tie_rel(params): 1e-5 for relu nets, 2e-4 for elu nets (the reference's float32 elu rule is
    itself only conditioned to ~5e-5, tests/test_oracle_golden.py::test_elu_rule_conditioning)."""
    scale = net.tol_scale(glo, gup, sc)
    assert np.all(np.abs(lo.astype(np.float64) - glo) <= rel * scale + 1e-30), \
        f"lower off by {np.max(np.abs(lo - glo) / scale):.3e} rel"
    assert np.all(np.abs(up.astype(np.float64) - gup) <= rel * scale + 1e-30), \
        f"upper off by {np.max(np.abs(up - gup) / scale):.3e} rel"


def check_labels(lab, glab, glo, gup, sc, offset=0.0, rel=RTOL):
    tie = net.bound_near_tie(glo, gup, offset, sc, rel=rel)
    bad = (lab != glab) & ~tie
    assert not bad.any(), f"{bad.sum()} label mismatches outside the near-tie band"
    return int(tie.sum())


def random_boxes(seed, n, smin=-9, smax=0):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    h = (2.0 ** rng.uniform(smin, smax, (n, 1)) * rng.uniform(0.5, 1.0, (n, 3))).astype(np.float32)
    return c - h, c + h


# ---------------------------------------------------------------------------------------------------
# primitives
# ---------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("name", SAMPLES)
def test_point_values_golden(name):
    import mlp
    g = golden(f"points_{name}")
    p = sample_params(name)
    f, s = mlp.eval_points(p, g["points"], return_scale=True)
    assert np.all(np.abs(f - g["values"]) <= RTOL * rays.point_scale(p, g["points"]))
    np.testing.assert_allclose(s, rays.point_scale(p, g["points"]), rtol=1e-4)


@pytest.mark.parametrize("n", [0, 1, 7, 1000, 70001])
def test_point_values_ragged_sizes(n):
    import mlp
    p = sample_params("bunny")
    x = np.random.default_rng(n).uniform(-1, 1, (n, 3)).astype(np.float32)
    f = mlp.eval_points(p, x)
    assert f.shape == (n,)
    if n:
        assert np.all(np.abs(f - net.eval_points(p, x)) <= RTOL * rays.point_scale(p, x))


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("name", SAMPLES)
def test_classify_golden(name, mode):
    g = golden(f"classify_{name}_{mode}")
    p = sample_params(name)
    rel = net.tie_rel(p)
    func = make(p, mode, g["n_trunc"])
    # scale of the golden bounds (oracle run gives the yardstick only)
    _, _, _, sc = net.classify_box(p, octx(mode, g["n_trunc"]), g["box_lower"], g["box_upper"], return_scale=True)
    lab, lo, up, tie = func.bound_box(p, g["box_lower"], g["box_upper"])
    check_bounds(lo, up, g["lower"], g["upper"], sc, rel)
    check_labels(lab, g["label"], g["lower"], g["upper"], sc, rel=rel)
    lab5 = func.classify_box(p, g["box_lower"], g["box_upper"], offset=0.05)
    check_labels(lab5, g["label_offset005"], g["lower"], g["upper"], sc, 0.05, rel=rel)
    # v = 1 general boxes
    _, _, _, sc = net.classify_general_box(p, octx(mode, g["n_trunc"]), g["seg_center"], g["seg_vecs"], return_scale=True)
    lab, lo, up, tie = func.bound_general_box(p, g["seg_center"], g["seg_vecs"])
    check_bounds(lo, up, g["seg_lower"], g["seg_upper"], sc, rel)
    check_labels(lab, g["seg_label"], g["seg_lower"], g["seg_upper"], sc, rel=rel)
    # rigid transform prepended
    import mlp
    p2 = mlp.prepend_op(p, mlp.spatial_transformation())
    p2["0000.spatial_transformation.R"] = g["xf_R"]
    p2["0000.spatial_transformation.t"] = g["xf_t"]
    op2 = net.prepend_op(p, net.spatial_transformation(g["xf_R"], g["xf_t"]))
    _, _, _, sc = net.classify_box(op2, octx(mode, g["n_trunc"]), g["box_lower"][9:18], g["box_upper"][9:18], return_scale=True)
    lab, lo, up, tie = func.bound_box(p2, g["box_lower"][9:18], g["box_upper"][9:18])
    check_bounds(lo, up, g["xf_lower"], g["xf_upper"], sc, rel)
    check_labels(lab, g["xf_label"], g["xf_lower"], g["xf_upper"], sc, rel=rel)
    f = func(p2, g["xf_points"])
    assert np.all(np.abs(f - g["xf_values"]) <= RTOL * rays.point_scale(op2, g["xf_points"]))


APPEND_N = 4


@pytest.mark.parametrize("name", SAMPLES)
def test_classify_append_golden_and_oracle(name):
    """the affine_append bounder (src/affine.py:183-191) against the golden vectors of the unmodified
    reference and against the oracle on random boxes.  Tolerance net.mode_rel (10x the hot-path band: the reference's
    `err + (sum(delta) - sum(kept))` cancels catastrophically; the oracle itself only matches the reference run to
    2e-5 / 4e-4 relative, tests/test_oracle_golden.py).  Soundness is checked exactly like the other modes."""
    import mlp
    p = sample_params(name)
    ctx = octx("affine_append", APPEND_N)
    rel = net.mode_rel(p, ctx)
    func = make(p, "affine_append", APPEND_N)
    g = golden(f"classify_{name}_affine_append")
    _, _, _, sc = net.classify_box(p, ctx, g["box_lower"], g["box_upper"], return_scale=True)
    lab, lo, up, tie = func.bound_box(p, g["box_lower"], g["box_upper"])
    check_bounds(lo, up, g["lower"], g["upper"], sc, rel)
    check_labels(lab, g["label"], g["lower"], g["upper"], sc, rel=rel)
    _, _, _, sc = net.classify_general_box(p, ctx, g["seg_center"], g["seg_vecs"], return_scale=True)
    lab, lo, up, tie = func.bound_general_box(p, g["seg_center"], g["seg_vecs"])
    check_bounds(lo, up, g["seg_lower"], g["seg_upper"], sc, rel)
    check_labels(lab, g["seg_label"], g["seg_lower"], g["seg_upper"], sc, rel=rel)
    # random boxes: held to the float64 value of the same formulas, and statistically no further from it than the
    # float32 oracle is (the yardstick of an ill-conditioned formula, like the elu case of test_classify_synthetic_widths)
    lo_b, hi_b = random_boxes(13, 1200)
    olab, olo, oup, sc = net.classify_box(p, ctx, lo_b, hi_b, return_scale=True)
    lab, lo, up, tie = func.bound_box(p, lo_b, hi_b)
    check_bounds(lo, up, olo, oup, sc, rel)
    with net.precision(np.float64):
        _, lo6, up6, sc6 = net.classify_box(p, ctx, lo_b, hi_b, return_scale=True)
    s6 = net.tol_scale(lo6, up6, sc6)
    e_gpu = np.maximum(np.abs(lo - lo6), np.abs(up - up6)) / s6
    e_o32 = np.maximum(np.abs(olo - lo6), np.abs(oup - up6)) / s6
    for q in (50, 90, 99, 100):
        assert np.percentile(e_gpu, q) <= 6 * np.percentile(e_o32, q) + 1e-5, f"P{q}: gpu {np.percentile(e_gpu, q):.2e} oracle {np.percentile(e_o32, q):.2e}"
    assert check_labels(lab, olab, olo, oup, sc, rel=rel) < 0.1 * 1200
    # soundness
    rng = np.random.default_rng(2)
    u = rng.uniform(0, 1, (1200, 8, 3)).astype(np.float32)
    x = lo_b[:, None, :] + u * (hi_b - lo_b)[:, None, :]
    f = mlp.eval_points(p, x)
    slack = 1e-5 * np.maximum(np.abs(lo), np.abs(up))[:, None] + 1e-6
    assert np.all(f >= lo[:, None] - slack) and np.all(f <= up[:, None] + slack)
    # argument checking: jax.lax.top_k needs k <= width
    with pytest.raises(ValueError):
        make(p, "affine_append", 1000).classify_box(p, LO, HI)


@pytest.mark.parametrize("name", SAMPLES)
def test_classify_sdf_golden_and_oracle(name):
    """the sdf bounder (src/sdf.py:31-50: f(centre) against lipschitz * radius) against the labels of
    the unmodified reference and against the oracle on random boxes; (lower, upper) = f -+ L*radius within 1e-5."""
    p = sample_params(name)
    g = golden(f"classify_{name}_sdf")
    L = float(g["lipschitz"])
    ctx = octx("sdf", L)
    func = make(p, "sdf", L)

    def check(lab, tie, glab, center, vecs, offset):
        val, reach = net.sdf_center_value_and_reach(p, ctx, center, vecs)
        band = RTOL * (rays.point_scale(p, center) + reach)
        otie = (np.abs(np.abs(val) - reach) <= band) | (np.abs(val - offset) <= band) | (np.abs(val + offset) <= band)
        assert np.all((lab == glab) | otie)
        assert np.all((lab == glab) | tie)                 # the device flag covers every disagreement too

    c, v = net.box_to_general(g["box_lower"], g["box_upper"])
    lab, lo, up, tie = func.bound_box(p, g["box_lower"], g["box_upper"])
    check(lab, tie, g["label"], c, v, 0.0)
    lab5, _, _, tie5 = func.bound_box(p, g["box_lower"], g["box_upper"], offset=0.05)
    check(lab5, tie5, g["label_offset005"], c, v, 0.05)
    labg, _, _, tieg = func.bound_general_box(p, g["gen_center"], g["gen_vecs"])
    check(labg, tieg, g["gen_label"], g["gen_center"], g["gen_vecs"], 0.0)
    lab1, _, _, tie1 = func.bound_general_box(p, g["gen_center"], g["gen_vecs"][:, :1])
    check(lab1, tie1, g["gen_label_v1"], g["gen_center"], g["gen_vecs"][:, :1], 0.0)
    assert np.array_equal(func.classify_box(p, g["box_lower"], g["box_upper"]), lab)
    # random boxes vs the oracle (labels + our bounds)
    lo_b, hi_b = random_boxes(17, 5000)
    c, v = net.box_to_general(lo_b, hi_b)
    olab, olo, oup, sc = net.classify_box(p, ctx, lo_b, hi_b, return_scale=True)
    lab, lo, up, tie = func.bound_box(p, lo_b, hi_b)
    check(lab, tie, olab, c, v, 0.0)
    assert np.all(np.abs(lo - olo) <= RTOL * sc) and np.all(np.abs(up - oup) <= RTOL * sc)
    assert tie.mean() < 0.01


@pytest.mark.parametrize("name", SAMPLES)
def test_classify_slope_interval_golden_and_oracle(name):
    """the slope_interval bounder on the engine's 7-row tile (k_classify_slope) against the run of the
    unmodified reference (labels + may-contain bounds: axis-aligned, offset, v=2 / v=1 general boxes, rigid transform),
    against the oracle on random boxes, and the soundness property."""
    import mlp
    p = sample_params(name)
    g = golden(f"classify_{name}_slope_interval")
    ctx = octx("slope_interval")
    rel = net.tie_rel(p)
    func = make(p, "slope_interval")
    _, _, _, sc = net.classify_box(p, ctx, g["box_lower"], g["box_upper"], return_scale=True)
    lab, lo, up, tie = func.bound_box(p, g["box_lower"], g["box_upper"])
    check_bounds(lo, up, g["lower"], g["upper"], sc, rel)
    check_labels(lab, g["label"], g["lower"], g["upper"], sc, rel=rel)
    lab5 = func.classify_box(p, g["box_lower"], g["box_upper"], offset=0.05)
    check_labels(lab5, g["label_offset005"], g["lower"], g["upper"], sc, 0.05, rel=rel)
    for tag, vecs in (("gen", g["gen_vecs"]), ("gen1", g["gen_vecs"][:, :1])):
        _, _, _, sc = net.classify_general_box(p, ctx, g["gen_center"], vecs, return_scale=True)
        lab, lo, up, tie = func.bound_general_box(p, g["gen_center"], vecs)
        check_bounds(lo, up, g[f"{tag}_lower"], g[f"{tag}_upper"], sc, rel)
        check_labels(lab, g[f"{tag}_label"], g[f"{tag}_lower"], g[f"{tag}_upper"], sc, rel=rel)
    p2 = mlp.prepend_op(p, mlp.spatial_transformation())
    p2["0000.spatial_transformation.R"] = g["xf_R"]
    p2["0000.spatial_transformation.t"] = g["xf_t"]
    op2 = net.prepend_op(p, net.spatial_transformation(g["xf_R"], g["xf_t"]))
    _, _, _, sc = net.classify_box(op2, ctx, g["box_lower"][9:18], g["box_upper"][9:18], return_scale=True)
    lab, lo, up, tie = func.bound_box(p2, g["box_lower"][9:18], g["box_upper"][9:18])
    check_bounds(lo, up, g["xf_lower"], g["xf_upper"], sc, rel)
    check_labels(lab, g["xf_label"], g["xf_lower"], g["xf_upper"], sc, rel=rel)
    # random boxes vs the oracle
    lo_b, hi_b = random_boxes(19, 8000)
    olab, olo, oup, sc = net.classify_box(p, ctx, lo_b, hi_b, return_scale=True)
    lab, lo, up, tie = func.bound_box(p, lo_b, hi_b)
    check_bounds(lo, up, olo, oup, sc, rel)
    assert check_labels(lab, olab, olo, oup, sc, rel=rel) < 0.01 * 8000
    assert np.all(tie | (lab == olab))
    # soundness
    rng = np.random.default_rng(2)
    u = rng.uniform(0, 1, (8000, 4, 3)).astype(np.float32)
    x = lo_b[:, None, :] + u * (hi_b - lo_b)[:, None, :]
    f = mlp.eval_points(p, x)
    slack = 1e-5 * np.maximum(np.abs(lo), np.abs(up))[:, None] + 1e-6
    assert np.all(f >= lo[:, None] - slack) and np.all(f <= up[:, None] + slack)


@pytest.mark.parametrize("width,act", [(256, "relu"), (128, "elu"), (40, "relu")])
def test_classify_slope_interval_synthetic_widths(width, act):
    """The 7-row slope tile on the other width classes (256: streamed weights + zero-skipping K loops)."""
    p = net.random_mlp([3] + [width] * 8 + [1], act, seed=0)
    lo_b, hi_b = random_boxes(3, 2000, smin=-12, smax=-2)
    ctx = octx("slope_interval")
    olab, olo, oup, sc = net.classify_box(p, ctx, lo_b, hi_b, return_scale=True)
    lab, lo, up, _ = make(p, "slope_interval").bound_box(p, lo_b, hi_b)
    rel = net.tie_rel(p)
    check_bounds(lo, up, olo, oup, sc, rel)
    check_labels(lab, olab, olo, oup, sc, rel=rel)


PE_MODES = ("interval", "affine_fixed", "affine_truncate", "affine_all", "slope_interval")


@pytest.mark.parametrize("mode", PE_MODES)
def test_positional_encoding_ops(mode):
    """sin + pow2_frequency_encode (positional-encoding MLPs of src/main_fit_implicit.py:113-115): the
    encode op is folded into a dense layer on the host, sin is a third activation rule of the engine / grow kernel.
    Against the run of the unmodified reference (labels, bounds, point values), the oracle on random boxes, soundness,
    and a small cast_rays through the persistent kernel."""
    import mlp
    import queries
    import render
    g = golden(f"pe_{mode}")
    p = {k.split("/", 1)[1]: g[k] for k in g if k.startswith("params/")}
    func = make(p, mode, g["n_trunc"])
    ctx = octx(mode, g["n_trunc"])
    rel = net.tie_rel(p)                     # 5e-5 for nets with a sin layer (float32 conditioning of the sin rule)
    _, _, _, sc = net.classify_box(p, ctx, g["box_lower"], g["box_upper"], return_scale=True)
    lab, lo, up, tie = func.bound_box(p, g["box_lower"], g["box_upper"])
    okg = np.ones(lab.shape[0], bool)
    if mode == "affine_truncate":
        # sin saturates to (alpha, beta, delta) = (0, 0, 1) on boxes that span whole periods: many new rows of EQUAL L1
        # norm, where a 1-ulp difference in sin decides which are kept -- truncation near-ties, set aside as elsewhere
        c, v = net.box_to_general(g["box_lower"], g["box_upper"])
        okg = ~net.truncate_rank_near_tie(p, ctx, c, v, rel=rel)
        assert okg.sum() >= 5
    check_bounds(lo[okg], up[okg], g["lower"][okg], g["upper"][okg], sc[okg], rel)
    check_labels(lab[okg], g["label"][okg], g["lower"][okg], g["upper"][okg], sc[okg], rel=rel)
    f = func(p, g["points"])
    assert np.all(np.abs(f - g["values"]) <= RTOL * rays.point_scale(p, g["points"]))
    n = 4000 if mode in ("interval", "affine_fixed", "slope_interval") else 600
    lo_b, hi_b = random_boxes(23, n, smin=-8, smax=-1)
    olab, olo, oup, sc = net.classify_box(p, ctx, lo_b, hi_b, return_scale=True)
    lab, lo, up, tie = func.bound_box(p, lo_b, hi_b)
    ok = np.ones(n, bool)
    if mode == "affine_truncate":
        c, v = net.box_to_general(lo_b, hi_b)
        ok = ~net.truncate_rank_near_tie(p, ctx, c, v, rel=rel)
        assert ok.mean() > 0.2
    check_bounds(lo[ok], up[ok], olo[ok], oup[ok], sc[ok], rel)
    assert check_labels(lab[ok], olab[ok], olo[ok], oup[ok], sc[ok], rel=rel) < 0.01 * n
    rng = np.random.default_rng(2)
    u = rng.uniform(0, 1, (n, 4, 3)).astype(np.float32)
    x = lo_b[:, None, :] + u * (hi_b - lo_b)[:, None, :]
    fx = mlp.eval_points(p, x)
    slack = 1e-5 * np.maximum(np.abs(lo), np.abs(up))[:, None] + 1e-6
    if mode != "interval":
        # (interval mode enters the sin layer with err > 0 and the reference then forms `alpha * err + delta` with a
        # possibly NEGATIVE alpha, src/affine.py:176 -- its bounds are not sound there, and parity reproduces them)
        assert np.all(fx >= lo[:, None] - slack) and np.all(fx <= up[:, None] + slack)
    if mode == "affine_fixed":
        eye = np.array((2., 1., 2.), np.float32)
        look, upd, _ = render.look_at(eye)
        roots, dirs = render.generate_camera_rays(eye, look, upd, res=24, fov_deg=30.)
        opts = queries.get_default_cast_opts()
        opts["n_max_step"] = 96
        t, hit, cnt, n_evals, gtie = queries.cast_rays((func,), (p,), roots, dirs, opts, return_near_tie=True)
        ot, ohit, ocnt, on_evals, otie = rays.cast_rays((ctx,), (p,), roots, dirs, opts, return_near_tie=True)
        okr = ~(gtie | otie)
        assert okr.mean() > 0.9
        assert np.array_equal(hit[okr], ohit[okr]) and np.array_equal(cnt[okr], ocnt[okr])
        np.testing.assert_allclose(t[okr], ot[okr], rtol=RTOL)


def test_classify_truncate64_golden():
    g = golden("classify_hammer_affine_truncate64")
    p = sample_params("hammer")
    func = make(p, "affine_truncate", 64)
    _, _, _, sc = net.classify_box(p, octx("affine_truncate", 64), g["box_lower"], g["box_upper"], return_scale=True)
    lab, lo, up, _ = func.bound_box(p, g["box_lower"], g["box_upper"])
    check_bounds(lo, up, g["lower"], g["upper"], sc)
    check_labels(lab, g["label"], g["lower"], g["upper"], sc)


@pytest.mark.parametrize("mode,n", [("interval", 20000), ("affine_fixed", 20000), ("affine_truncate", 1500), ("affine_all", 1500)])
@pytest.mark.parametrize("name", SAMPLES)
def test_classify_vs_oracle_random(name, mode, n):
    p = sample_params(name)
    lo_b, hi_b = random_boxes(11, n)
    func = make(p, mode, 16)
    olab, olo, oup, sc = net.classify_box(p, octx(mode, 16), lo_b, hi_b, return_scale=True)
    lab, lo, up, tie = func.bound_box(p, lo_b, hi_b)
    rel = net.tie_rel(p)
    ok = np.ones(n, bool)
    if mode == "affine_truncate":
        # keep/drop decisions between rows of (nearly) equal L1 norm flip with float32 summation order and change the
        # affine form, not just its last bits: those boxes are near-ties of the truncation, counted and set aside
        c, v = net.box_to_general(lo_b, hi_b)
        rank_tie = net.truncate_rank_near_tie(p, octx(mode, 16), c, v, rel=rel)
        assert rank_tie.mean() < (0.1 if rel > 1e-5 else 0.01), f"{rank_tie.sum()} truncation near-ties"   # elu nets: 2e-4 band
        ok = ~rank_tie
    check_bounds(lo[ok], up[ok], olo[ok], oup[ok], sc[ok], rel)
    n_tie = check_labels(lab[ok], olab[ok], olo[ok], oup[ok], sc[ok], rel=rel)
    assert n_tie < 0.01 * n
    # the device's own near-tie flag covers every label disagreement
    assert np.all((tie | (lab == olab))[ok])
    assert tie.mean() < 0.01


@pytest.mark.parametrize("n", [0, 1, 3, 129, 4099])
def test_classify_ragged_sizes(n):
    p = sample_params("fox")
    lo_b, hi_b = random_boxes(5, n)
    for mode in ("affine_fixed", "affine_all"):
        lab, lo, up, tie = make(p, mode).bound_box(p, lo_b, hi_b)
        assert lab.shape == (n,)
        if n:
            olab, olo, oup, sc = net.classify_box(p, octx(mode), lo_b, hi_b, return_scale=True)
            check_bounds(lo, up, olo, oup, sc)
            check_labels(lab, olab, olo, oup, sc)
            assert np.all(tie | (lab == olab))


@pytest.mark.parametrize("width,act", [(256, "relu"), (128, "elu"), (40, "relu"), (64, "relu")])
def test_classify_synthetic_widths(width, act):
    """Random-init MLPs of other width classes (config 5's 8x256; odd widths exercise the padding)."""
    p = net.random_mlp([3] + [width] * 8 + [1], act, seed=0)
    lo_b, hi_b = random_boxes(3, 3000, smin=-12, smax=-2)
    func = make(p, "affine_fixed")
    olab, olo, oup, sc = net.classify_box(p, octx("affine_fixed"), lo_b, hi_b, return_scale=True)
    lab, lo, up, _ = func.bound_box(p, lo_b, hi_b)
    if act == "elu":
        # A random-init elu net has a large layer gain: the float32 noise of the reference's elu rule (delta is a
        # difference of O(1) terms) is amplified until it dominates the radius of tiny boxes.  No float32
        # implementation can agree with another to 1e-5 there, so the GPU result is held to the float64 value
        # of the same formulas, and must be statistically no further from it than the float32 oracle is.
        with net.precision(np.float64):
            _, lo6, up6, sc6 = net.classify_box(p, octx("affine_fixed"), lo_b, hi_b, return_scale=True)
        s6 = net.tol_scale(lo6, up6, sc6)
        e_gpu = np.maximum(np.abs(lo - lo6), np.abs(up - up6)) / s6
        e_o32 = np.maximum(np.abs(olo - lo6), np.abs(oup - up6)) / s6
        for q in (50, 90, 99, 100):
            assert np.percentile(e_gpu, q) <= 6 * np.percentile(e_o32, q) + 1e-5, f"P{q}: gpu {np.percentile(e_gpu, q):.2e} oracle {np.percentile(e_o32, q):.2e}"
        far = np.minimum(np.abs(lo6), np.abs(up6)) > 8 * np.maximum(e_gpu, e_o32) * s6 + 1e-5 * s6
        assert np.all(lab[far] == olab[far])
    else:
        check_bounds(lo, up, olo, oup, sc)
        check_labels(lab, olab, olo, oup, sc)
    x = np.random.default_rng(1).uniform(-1, 1, (5000, 3)).astype(np.float32)
    f = func(p, x)
    assert np.all(np.abs(f - net.eval_points(p, x)) <= RTOL * rays.point_scale(p, x))
    if width <= 128 and act == "relu":
        funca = make(p, "affine_all")
        olab, olo, oup, sc = net.classify_box(p, octx("affine_all"), lo_b[:300], hi_b[:300], return_scale=True)
        lab, lo, up, _ = funca.bound_box(p, lo_b[:300], hi_b[:300])
        check_bounds(lo, up, olo, oup, sc)
        check_labels(lab, olab, olo, oup, sc)


@pytest.mark.parametrize("mode", MODES)
def test_soundness_property(mode):
    """Size-independent property: f(x) sampled inside a box lies within the computed bounds."""
    import mlp
    rng = np.random.default_rng(2)
    for name in SAMPLES:
        p = sample_params(name)
        n = 4000 if mode in ("interval", "affine_fixed") else 400
        lo_b, hi_b = random_boxes(9, n, smin=-7, smax=-1)
        _, lo, up, _ = make(p, mode, 16).bound_box(p, lo_b, hi_b)
        u = rng.uniform(0, 1, (n, 8, 3)).astype(np.float32)
        x = lo_b[:, None, :] + u * (hi_b - lo_b)[:, None, :]
        f = mlp.eval_points(p, x)
        slack = 1e-5 * np.maximum(np.abs(lo), np.abs(up))[:, None] + 1e-6
        assert np.all(f >= lo[:, None] - slack) and np.all(f <= up[:, None] + slack)


def test_params_are_read_afresh_each_call():
    """The reference's GUI mutates params between calls (src/main_intersection.py:171-183)."""
    import mlp
    p = mlp.prepend_op(sample_params("bunny"), mlp.spatial_transformation())
    func = make(p, "affine_fixed")
    x = np.array([[0.1, 0.2, -0.1]], np.float32)
    f0 = func(p, x)
    p["0000.spatial_transformation.t"] = np.array([0.3, 0.0, 0.0], np.float32)
    f1 = func(p, x)
    f1_ref = func(sample_params("bunny"), x - np.array([[0.3, 0, 0]], np.float32))
    assert f0[0] != f1[0]
    np.testing.assert_allclose(f1, f1_ref, rtol=1e-5)


@pytest.mark.parametrize("name", ["fox", "bunny"])
def test_slope_min_distance_helpers(name):
    """SlopeIntervalImplicitFunction.min_distance_to_zero / min_distance_to_zero_in_direction (src/slope_interval.py:52-163) on
    the CUDA path (niq_slope_forward + the reference's closed-form arithmetic) against the goldens of the unmodified reference,
    with the checker the oracle is pinned by; a single box / ray must return scalars like the reference."""
    import implicit_mlp_utils
    from test_oracle_golden import check_min_distance
    p = sample_params(name)
    f = implicit_mlp_utils.generate_implicit_from_params(p, "slope_interval")
    g = golden(f"mindist_{name}_slope")
    rep = {}
    check_min_distance(p, g, lambda pp, c, a: f.min_distance_to_zero(pp, c, a, return_source_value=True),
                       lambda pp, s, b, r: f.min_distance_to_zero_in_direction(pp, s, b, source_range=r, return_source_value=True),
                       net.tie_rel(p), report=rep)
    parity_report(f"slope_min_distance[{name}]", max_rel_distance_error=rep)
    d1 = f.min_distance_to_zero(p, g["box_center"][3], g["box_axis_vec"][3])
    assert np.ndim(d1) == 0 and d1 == f.min_distance_to_zero(p, g["box_center"], g["box_axis_vec"])[3]
    v1, r1 = f.min_distance_to_zero_in_direction(p, g["source"][5], g["bound_vec"][5], return_source_value=True)
    assert np.ndim(r1) == 0 and r1 == f.min_distance_to_zero_in_direction(p, g["source"], g["bound_vec"])[5]


def test_unsupported_and_invalid_arguments():
    import _niq
    import implicit_mlp_utils
    p = sample_params("fox")
    with pytest.raises(RuntimeError):
        implicit_mlp_utils.generate_implicit_from_params(p, "tanh_interval")          # not a mode of the reference
    with pytest.raises(_niq.NiqError):                                              # more than 3 box vectors: forward + 3 source-range vectors
        implicit_mlp_utils.generate_implicit_from_params(p, "slope_interval").min_distance_to_zero_in_direction(
            p, LO, HI, source_range=np.eye(3, dtype=np.float32) * 0.01)
    f = implicit_mlp_utils.generate_implicit_from_params(p, "affine_truncate", affine_n_truncate=8,
                                                         affine_truncate_policy="relative")
    with pytest.raises(_niq.NiqError):
        f.classify_box(p, LO, HI)
    bad = dict(p)
    bad["0001.softplus._"] = bad.pop("0001.relu._")
    with pytest.raises(RuntimeError):
        make(bad, "affine_fixed").classify_box(bad, LO, HI)
    import kd_tree
    with pytest.raises(ValueError):
        kd_tree.construct_uniform_unknown_levelset_tree(make(p, "affine_fixed"), p, LO, HI)
    with pytest.raises(ValueError):
        kd_tree.construct_uniform_unknown_levelset_tree(make(p, "affine_fixed"), p, LO, HI, split_depth=3, batch_process_size=100)
    with pytest.raises(ValueError):
        kd_tree.find_any_intersection((make(p, "affine_fixed"),), (p,), LO, HI, 1e-3)


# ---------------------------------------------------------------------------------------------------
# cast_rays
# ---------------------------------------------------------------------------------------------------

RAY_CASES = {
    "rays_fox_fixed_r12": (("fox",), "affine_fixed"),
    "rays_fox_interval_r6": (("fox",), "interval"),
    "rays_fox_all_r6": (("fox",), "affine_all"),
    "rays_fox_fixed_r8_sub3": (("fox",), "affine_fixed"),
    "rays_fox_bunny_fixed_r8": (("fox", "bunny"), "affine_fixed"),
}


@pytest.mark.parametrize("case", sorted(RAY_CASES))
def test_cast_rays_golden(case):
    import queries
    import render
    names, mode = RAY_CASES[case]
    g = golden(case)
    look, up, _ = render.look_at(g["eye"])
    roots, dirs = render.generate_camera_rays(g["eye"], look, up, res=int(g["res"]), fov_deg=30.0)
    np.testing.assert_array_equal(roots, g["roots"])
    np.testing.assert_allclose(dirs, g["dirs"], rtol=0, atol=2e-7)
    opts = queries.get_default_cast_opts()
    opts["n_substeps"] = int(g["n_substeps"])
    ps = tuple(sample_params(n) for n in names)
    funcs = tuple(make(p, mode, g["n_trunc"]) for p in ps)
    t, hit, cnt, n_evals, tie = queries.cast_rays(funcs, ps, g["roots"], g["dirs"], opts, return_near_tie=True)
    assert t.dtype == np.float32 and hit.dtype == np.int32 and cnt.dtype == np.int32
    ok = ~tie
    assert ok.mean() > 0.9
    np.testing.assert_array_equal(hit[ok], g["out_hit_id"][ok])
    np.testing.assert_array_equal(cnt[ok], g["out_count"][ok])
    np.testing.assert_allclose(t[ok], g["out_t"][ok], rtol=RTOL, atol=0)
    if not tie.any():
        assert n_evals == int(g["n_evals"])


@pytest.mark.parametrize("name,mode,res", [("fox", "affine_fixed", 96), ("bunny", "affine_fixed", 48), ("hammer", "interval", 32),
                                           ("fox", "slope_interval", 24), ("fox", "sdf", 24)])
def test_cast_rays_vs_oracle(name, mode, res):
    import queries
    import render
    p = sample_params(name)
    eye = np.array((2., 1., 2.), np.float32)
    look, up, _ = render.look_at(eye)
    roots, dirs = render.generate_camera_rays(eye, look, up, res=res, fov_deg=30.)
    opts = queries.get_default_cast_opts()
    t, hit, cnt, n_evals, tie = queries.cast_rays((make(p, mode),), (p,), roots, dirs, opts, return_near_tie=True)
    ot, ohit, ocnt, on_evals, otie = rays.cast_rays((octx(mode),), (p,), roots, dirs, opts, return_near_tie=True)
    ok = ~(tie | otie)
    # a ray is flagged when ANY of its ~90 steps had a decision inside the band (2e-4 for the elu net)
    assert ok.mean() > (0.97 if net.tie_rel(p) == 1e-5 else 0.85), f"{(~ok).sum()} near-tie rays of {ok.size}"
    np.testing.assert_array_equal(hit[ok], ohit[ok])
    np.testing.assert_array_equal(cnt[ok], ocnt[ok])
    np.testing.assert_allclose(t[ok], ot[ok], rtol=RTOL, atol=0)
    assert (hit == 0).any()
    assert mode in ("interval", "sdf") or (hit > 0).any()      # interval bounds are so loose that rays crawl to the step limit
    if ok.all():
        assert n_evals == on_evals


@pytest.mark.parametrize("mode,n_trunc,n_sub", [("affine_truncate", 8, 1), ("affine_all", 8, 1), ("affine_append", 4, 1), ("affine_truncate", 64, 3)])
def test_cast_rays_grow_kernel_equals_host_loop(monkeypatch, mode, n_trunc, n_sub):
    """The persistent ray kernel of the growing-form modes (csrc/niq_rays_grow.cuh) against the host-level iteration it replaces
    (NIQ_RAYS_HOST_LOOP=1): same propagation code per segment, same summation order of the two point values -> identical."""
    import queries
    import render
    pf, pb = sample_params("fox"), sample_params("bunny")
    funcs = (make(pf, mode, n_trunc), make(pb, mode, n_trunc))
    eye = np.array((2., 1., 2.), np.float32)
    look, up, _ = render.look_at(eye)
    roots, dirs = render.generate_camera_rays(eye, look, up, res=20, fov_deg=30.)
    opts = queries.get_default_cast_opts()
    opts["n_substeps"] = n_sub
    res = []
    for host in ("", "1"):
        if host:
            monkeypatch.setenv("NIQ_RAYS_HOST_LOOP", host)
        else:
            monkeypatch.delenv("NIQ_RAYS_HOST_LOOP", raising=False)
        res.append(queries.cast_rays(funcs, (pf, pb), roots, dirs, opts, return_near_tie=True))
    a, b = res
    for x, y in zip(a[:4], b[:4]):                 # t, hit_id, count, N_evals
        np.testing.assert_array_equal(x, y)
    # near-tie flags: the kernel bands every func's bounds with the WIDEST tie_rel of the call (2e-4 here: bunny has elu), the
    # host loop with each func's own -> the kernel flags a superset
    assert not (b[4] & ~a[4]).any()
    parity_report(f"cast_rays_grow_kernel_vs_host_loop[{mode}-{n_trunc}-{n_sub}]", rays=int(a[0].shape[0]), flagged_kernel=int(a[4].sum()),
                  flagged_host_loop=int(b[4].sum()))
    assert (a[1] == 1).any() and (a[1] == 2).any() and (a[1] == 0).any()


def test_cast_rays_empty_and_single():
    import queries
    p = sample_params("fox")
    f = make(p, "affine_fixed")
    opts = queries.get_default_cast_opts()
    t, hit, cnt, n = queries.cast_rays((f,), (p,), np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32), opts)
    assert t.shape == (0,) and n == 0
    r = np.array([[2., 1., 2.]], np.float32)
    d = -r / np.linalg.norm(r)
    t, hit, cnt, n = queries.cast_rays((f,), (p,), r, d.astype(np.float32), opts)
    ot, ohit, ocnt, on = rays.cast_rays((octx("affine_fixed"),), (p,), r, d.astype(np.float32), opts)
    assert hit[0] == ohit[0] and cnt[0] == ocnt[0] and n == on
    np.testing.assert_allclose(t, ot, rtol=RTOL)


def test_cast_rays_full_size_properties():
    """Config 1 size (fox 512x512): properties that need no oracle run -- every ray terminates with a legal
    state, hits lie on a sign change of f within hit_eps, misses are beyond max_dist."""
    import mlp
    import queries
    import render
    p = sample_params("fox")
    eye = np.array((2., 1., 2.), np.float32)
    look, up, _ = render.look_at(eye)
    roots, dirs = render.generate_camera_rays(eye, look, up, res=512, fov_deg=30.)
    opts = queries.get_default_cast_opts()
    t, hit, cnt, n_evals, tie = queries.cast_rays((make(p, "affine_fixed"),), (p,), roots, dirs, opts, return_near_tie=True)
    assert np.all((cnt >= 1) & (cnt <= opts["n_max_step"]))
    miss = hit == 0
    assert np.all((t[miss] > opts["max_dist"]) | (cnt[miss] >= opts["n_max_step"]))
    h = hit == 1
    assert 0.1 < h.mean() < 0.5
    x0 = roots[h] + t[h, None] * dirs[h]
    x1 = roots[h] + (t[h] + np.float32(opts["hit_eps"]))[:, None] * dirs[h]
    f0, f1 = mlp.eval_points(p, x0), mlp.eval_points(p, x1)
    assert np.mean(np.sign(f0) != np.sign(f1)) > 0.999      # position is recomputed on the host: allow ulp ties
    assert n_evals >= int(cnt.sum())
    assert tie.mean() < 0.02


@pytest.mark.parametrize("width", [128, 256])
def test_cast_rays_wide_relu_vs_oracle(width):
    """The 128 / 256-wide engine (streamed weights, zero-skipping K loops) through queries.cast_rays against the
    oracle: the synthetic config-5 network family, few rays, step limit cut so the oracle finishes in seconds."""
    import queries
    import render
    p = net.random_mlp([3] + [width] * 8 + [1], "relu", seed=0)
    eye = np.array((2., 1., 2.), np.float32)
    look, up, _ = render.look_at(eye)
    roots, dirs = render.generate_camera_rays(eye, look, up, res=20, fov_deg=30.)
    opts = queries.get_default_cast_opts()
    opts["n_max_step"] = 48
    t, hit, cnt, n_evals, tie = queries.cast_rays((make(p, "affine_fixed"),), (p,), roots, dirs, opts, return_near_tie=True)
    ot, ohit, ocnt, on_evals, otie = rays.cast_rays((octx("affine_fixed"),), (p,), roots, dirs, opts, return_near_tie=True)
    ok = ~(tie | otie)
    assert ok.mean() > 0.9
    assert np.array_equal(hit[ok], ohit[ok]) and np.array_equal(cnt[ok], ocnt[ok])
    np.testing.assert_allclose(t[ok], ot[ok], rtol=RTOL)


def test_zero_skipping_equals_dense(monkeypatch):
    """Engine A/B: the list-driven K loops (columns that are exactly zero after a relu layer are dropped) against the
    dense loops (NIQ_NO_SPARSE=1).  Hidden layers are bit-identical; the lane-split dot product of the last layer
    sums in a different order, so values may differ by a few ulp of the summed magnitude."""
    import _niq
    import mlp
    import queries
    import render
    p = net.random_mlp([3] + [256] * 8 + [1], "relu", seed=0)
    func = make(p, "affine_fixed")
    rng = np.random.default_rng(5)
    c = rng.uniform(-1, 1, (4001, 3)).astype(np.float32)
    h = (2.0 ** rng.uniform(-12, -1, (4001, 1)) * rng.uniform(0.5, 1, (4001, 3))).astype(np.float32)
    x = (c + 0.5 * h).astype(np.float32)
    eye = np.array((2., 1., 2.), np.float32)
    look, up, _ = render.look_at(eye)
    roots, dirs = render.generate_camera_rays(eye, look, up, res=24, fov_deg=30.)
    opts = queries.get_default_cast_opts()
    opts["n_max_step"] = 40
    res = {}
    for tag, env in (("dense", "1"), ("sparse", "0")):
        monkeypatch.setenv("NIQ_NO_SPARSE", env)
        ctx = _niq.Context(0)
        try:
            ctx.exec_macs(on=True, reset=True)
            lab, lo, upb, tie = func.bound_box(p, c - h, c + h, ctx=ctx)
            f, sc = mlp.eval_points(p, x, return_scale=True, ctx=ctx)
            macs = ctx.exec_macs(on=False, reset=True)
            r = queries.cast_rays((func,), (p,), roots, dirs, opts, return_near_tie=True, ctx=ctx)
            res[tag] = (lab, lo, upb, tie, f, sc, r, macs)
        finally:
            ctx.close()
    d, s_ = res["dense"], res["sparse"]
    mag = np.maximum(np.abs(d[1]), np.abs(d[2])) + 1e-30
    assert np.all(np.abs(d[1] - s_[1]) <= 4e-6 * mag) and np.all(np.abs(d[2] - s_[2]) <= 4e-6 * mag)
    assert np.all((d[0] == s_[0]) | d[3].astype(bool) | s_[3].astype(bool))
    assert np.all(np.abs(d[4] - s_[4]) <= 4e-6 * d[5])
    ok = ~(d[6][4] | s_[6][4])
    assert ok.mean() > 0.9 and np.array_equal(d[6][1][ok], s_[6][1][ok]) and np.array_equal(d[6][2][ok], s_[6][2][ok])
    np.testing.assert_allclose(d[6][0][ok], s_[6][0][ok], rtol=RTOL)
    # the accounting: dense executes every (padded) column, zero-skipping fewer (these boxes are spatially unrelated, so
    # the two tiles of a warp share few dead columns; coherent workloads reach ~0.55, see bench.py)
    assert s_[7] < 0.97 * d[7]


def test_weight_ring_stress_skewed_trip_counts(monkeypatch):
    """Stress of the barrier-free streamed weight ring (niq_engine.cuh acquire_chunk / release): within every CTA the warps
    0-3 get tiny boxes (most columns dead after each relu -> short list-driven K loops) and the warps 4-7 huge ones (no dead
    column -> full loops), so the warps drift as far apart as the ring allows while the last one to release a stage refills
    it.  A race between a refill and a late reader would corrupt weights: repeated launches must be bit-identical to each
    other, equal to the dense loops (NIQ_NO_SPARSE=1) within the summation-order noise, and equal to the oracle."""
    import _niq
    p = net.random_mlp([3] + [256] * 8 + [1], "relu", seed=0)
    func = make(p, "affine_fixed")
    rng = np.random.default_rng(9)
    n = 16 * 148 * 6 + 5                               # six passes per CTA and a ragged tail
    c = rng.uniform(-0.8, 0.8, (n, 3)).astype(np.float32)
    warp = (np.arange(n) % 16) // 2                     # 16 boxes per CTA pass, 2 per warp (Engine<256, TileBox3>)
    h = np.where(warp[:, None] < 4, 2.0 ** -14, 0.9).astype(np.float32) * rng.uniform(0.5, 1, (n, 3)).astype(np.float32)
    runs = {}
    for tag, env in (("sparse", "0"), ("dense", "1")):
        monkeypatch.setenv("NIQ_NO_SPARSE", env)
        ctx = _niq.Context(0)
        try:
            runs[tag] = [func.bound_box(p, c - h, c + h, ctx=ctx) for _ in range(6 if tag == "sparse" else 1)]
        finally:
            ctx.close()
    first = runs["sparse"][0]
    for r in runs["sparse"][1:]:
        for a, b in zip(first, r):
            np.testing.assert_array_equal(a, b)
    d = runs["dense"][0]
    mag = np.maximum(np.abs(d[1]), np.abs(d[2]))
    mag = np.maximum(mag, np.median(mag))        # a bound that cancels to ~0 still carries the summation noise of its terms
    assert np.all(np.abs(d[1] - first[1]) <= 4e-6 * mag) and np.all(np.abs(d[2] - first[2]) <= 4e-6 * mag)
    sub = slice(0, 600)
    olab, olo, oup, osc = net.classify_box(p, octx("affine_fixed"), (c - h)[sub], (c + h)[sub], return_scale=True)
    check_bounds(first[1][sub], first[2][sub], olo, oup, osc)
    check_labels(first[0][sub], olab, olo, oup, osc)


def test_exec_macs_counter_and_device_timer():
    """niq_ctx_exec_macs counts the multiply-adds the network kernels issue; an elu net has no exact zeros, so the
    count is the padded dense count (>= algorithmic 5*M per box).  niq_ctx_timer_* brackets device time only."""
    import _niq
    ctx = _niq.Context(0)
    try:
        p = sample_params("bunny")                      # elu, 8x64
        func = make(p, "affine_fixed")
        lo_b, hi_b = random_boxes(4, 2048, smin=-8, smax=-3)
        ctx.exec_macs(on=True, reset=True)
        ctx.timer_start()
        func.bound_box(p, lo_b, hi_b, ctx=ctx)
        ms = ctx.timer_stop()
        macs = ctx.exec_macs(on=False, reset=True)
        M = ctx.mlp(p).macs
        assert 5 * M * 2048 <= macs <= 1.5 * 5 * M * 2048
        assert 0.0 < ms < 1000.0
        ctx.timer_start()
        assert ctx.timer_stop() == 0.0                  # no call in the bracket
        func.bound_box(p, lo_b, hi_b, ctx=ctx)
        assert ctx.exec_macs(on=False) == 0             # counter off: nothing counted
    finally:
        ctx.close()


# ---------------------------------------------------------------------------------------------------
# level-set tree, marching cubes
# ---------------------------------------------------------------------------------------------------

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
def test_tree_golden(case):
    import kd_tree
    name, mode = TREE_CASES[case]
    g = golden(case)
    kw = {k[3:]: g[k].item() for k in g if k.startswith("kw_")}
    stats = {}
    p = sample_params(name)
    out = kd_tree.construct_uniform_unknown_levelset_tree(make(p, mode, g["n_trunc"]), p, LO, HI, stats=stats, **kw)
    assert stats["n_near_tie"] == 0
    for tag in ("unknown", "interior", "exterior"):
        if f"{tag}_node_valid" not in g:
            assert f"{tag}_node_valid" not in out
            continue
        gv, v = g[f"{tag}_node_valid"], out[f"{tag}_node_valid"]
        assert v.shape == gv.shape                       # the reference's padded bucket size
        np.testing.assert_array_equal(v, gv)
        np.testing.assert_array_equal(out[f"{tag}_node_lower"][v], g[f"{tag}_node_lower"][gv])    # order too
        np.testing.assert_array_equal(out[f"{tag}_node_upper"][v], g[f"{tag}_node_upper"][gv])


def _canon(lo, hi):
    a = np.concatenate((lo, hi), axis=1)
    return a[np.lexsort(a.T[::-1])]


@pytest.mark.parametrize("name,depth", [("bunny", 15), ("birdcage_occ", 14), ("hammer", 13)])
def test_tree_vs_oracle(name, depth):
    import kd_tree
    p = sample_params(name)
    st, ost = {}, {}
    out = kd_tree.construct_uniform_unknown_levelset_tree(make(p, "affine_fixed"), p, LO, HI, split_depth=depth,
                                                          with_interior_nodes=True, with_exterior_nodes=True, stats=st)
    ref = otree.construct_uniform_unknown_levelset_tree(octx("affine_fixed"), p, LO, HI, split_depth=depth,
                                                        with_interior_nodes=True, with_exterior_nodes=True, stats=ost)
    n_tie = ost["n_near_tie"] + st["n_near_tie"]
    parity_report(f"tree_vs_oracle[{name}-d{depth}]", boxes=ost["n_evals"], near_tie_gpu=st["n_near_tie"], near_tie_oracle=ost["n_near_tie"])
    assert n_tie <= 1e-3 * ost["n_evals"] + 2, "more than 0.1 % of the boxes inside the near-tie band"
    if n_tie == 0:
        assert st["n_evals"] == ost["n_evals"]
        for tag in ("unknown", "interior", "exterior"):
            v, rv = out[f"{tag}_node_valid"], ref[f"{tag}_node_valid"]
            assert v.shape == rv.shape
            np.testing.assert_array_equal(out[f"{tag}_node_lower"][v], ref[f"{tag}_node_lower"][rv])
            np.testing.assert_array_equal(out[f"{tag}_node_upper"][v], ref[f"{tag}_node_upper"][rv])
    else:
        # topology may differ only below near-tie boxes: every leaf of one tree that is missing from the other must lie
        # inside a near-tie box's subtree, so the symmetric difference is bounded by the flagged boxes' descendants
        a = {r.tobytes() for r in _canon(out["unknown_node_lower"][out["unknown_node_valid"]], out["unknown_node_upper"][out["unknown_node_valid"]])}
        b = {r.tobytes() for r in _canon(ref["unknown_node_lower"][ref["unknown_node_valid"]], ref["unknown_node_upper"][ref["unknown_node_valid"]])}
        n_diff = len(a ^ b)
        parity_report(f"tree_vs_oracle[{name}-d{depth}]:topology", leaves=len(b), leaves_differing=n_diff)
        assert n_diff <= n_tie * 2 ** 4


def test_tree_deep_properties():
    """bunny split_depth 18 (~100k boxes): leaves are disjoint dyadic boxes of equal size, surface samples
    found by ray casting all fall inside some UNKNOWN leaf, interior/exterior boxes carry the right sign."""
    import kd_tree
    import mlp
    p = sample_params("bunny")
    func = make(p, "affine_fixed")
    st = {}
    out = kd_tree.construct_uniform_unknown_levelset_tree(func, p, LO, HI, split_depth=18, with_interior_nodes=True,
                                                          with_exterior_nodes=True, stats=st)
    v = out["unknown_node_valid"]
    lo, hi = out["unknown_node_lower"][v], out["unknown_node_upper"][v]
    assert lo.shape[0] > 10000
    ext = hi - lo
    assert np.all(ext == ext[0])                         # uniform depth
    key = np.round((lo + 1) / ext[0]).astype(np.int64)
    assert np.unique(key, axis=0).shape[0] == lo.shape[0]    # no duplicates
    for tag, sign in (("interior", -1), ("exterior", 1)):
        m = out[f"{tag}_node_valid"]
        c = 0.5 * (out[f"{tag}_node_lower"][m] + out[f"{tag}_node_upper"][m])
        if c.shape[0]:
            assert np.all(np.sign(mlp.eval_points(p, c)) == sign)
    # volume is conserved: unknown + interior + exterior tile the domain
    vol = lambda a, b: np.prod((b - a).astype(np.float64), axis=1).sum()
    total = vol(lo, hi) + sum(vol(out[f"{t}_node_lower"][out[f"{t}_node_valid"]], out[f"{t}_node_upper"][out[f"{t}_node_valid"]])
                              for t in ("interior", "exterior"))
    np.testing.assert_allclose(total, 8.0, rtol=1e-9)
    assert st["n_evals"] > 50000


def _canon(lo, hi):
    a = np.concatenate((lo, hi), axis=1)
    return a[np.lexsort(a.T[::-1])]


def test_tree_multi_root_and_sharded_equal_single_tree():
    """niq_tree_build_roots (the unit of the multi-GPU subtree partition): refining all frontier boxes of a depth-6
    tree in ONE build to total depth 15 yields exactly the leaves of the single depth-15 tree (as a set: the order
    differs), and so does sharding.tree_sharded at world size 1."""
    import kd_tree
    import sharding
    p = sample_params("bunny")
    func = make(p, "affine_fixed")
    full = kd_tree.construct_uniform_unknown_levelset_tree(func, p, LO, HI, split_depth=15)
    v = full["unknown_node_valid"]
    want = _canon(full["unknown_node_lower"][v], full["unknown_node_upper"][v])
    top = kd_tree.construct_uniform_unknown_levelset_tree(func, p, LO, HI, split_depth=6)
    tv = top["unknown_node_valid"]
    tree = kd_tree.build_tree(func, p, top["unknown_node_lower"][tv], top["unknown_node_upper"][tv], split_depth=9)
    try:
        lo, hi = tree.nodes(0)
        assert tree.stats()["n_levels"] == 10
    finally:
        tree.close()
    assert np.array_equal(_canon(lo, hi), want)
    slo, shi = sharding.tree_sharded(func, p, LO, HI, 15)
    assert np.array_equal(_canon(slo, shi), want)
    with pytest.raises(ValueError):
        kd_tree.build_tree(func, p, np.zeros((2, 2), np.float32), np.ones((2, 2), np.float32), split_depth=2)


@pytest.mark.parametrize("name,mode,depth,deal,world,kw", [
    ("bunny", "affine_fixed", 15, 6, 3, {}),
    ("bunny", "affine_fixed", 16, 12, 8, {}),            # the default top depth of sharding.tree_sharded
    ("fox", "interval", 13, 0, 2, {}),                   # deal at the root level: one rank owns the whole tree, the other nothing
    ("hammer", "slope_interval", 12, 5, 4, dict(offset=0.01)),
    ("fox", "affine_fixed", 14, 14, 5, {}),              # deal at the last level (classification only, no split)
    ("bunny", "affine_fixed", 14, 7, 3, dict(batch_process_size=128)),
])
def test_tree_dealt_builds_partition_the_single_tree(monkeypatch, name, mode, depth, deal, world, kw):
    """niq_tree_build_dealt (one persistent launch: replicated top, round-robin deal of the frontier entering level `deal`,
    own subtrees below): the ranks' leaf sets are disjoint and their union is exactly the leaf set of the single tree; the
    boxes classified over all ranks = the single tree's below the deal + world x the replicated top.  Also through the
    capacity-growth relaunch (NIQ_TREE_CAP)."""
    import _niq
    import kd_tree
    p = sample_params(name)
    func = make(p, mode)
    st = {}
    full = kd_tree.construct_uniform_unknown_levelset_tree(func, p, LO, HI, split_depth=depth, stats=st, **kw)
    v = full["unknown_node_valid"]
    want = _canon(full["unknown_node_lower"][v], full["unknown_node_upper"][v])
    for cap in (None, "64"):
        if cap:
            monkeypatch.setenv("NIQ_TREE_CAP", cap)
        parts, evals = [], 0
        for rank in range(world):
            t = kd_tree.build_tree_dealt(func, p, LO, HI, depth, deal, rank, world, **kw)
            try:
                parts.append(t.nodes(0))
                evals += t.stats()["n_evals"]
            finally:
                t.close()
        lo = np.concatenate([a for a, _ in parts]); hi = np.concatenate([b for _, b in parts])
        got = _canon(lo, hi)
        assert got.shape == want.shape and np.array_equal(got, want)           # same multiset: no leaf twice, none missing
        top_boxes = sum(st["level_sizes"][:deal])
        assert evals == st["n_evals"] + (world - 1) * top_boxes
    f2 = make(p, "affine_truncate")
    with pytest.raises(_niq.NiqError):
        kd_tree.build_tree_dealt(f2, p, LO, HI, 6, 3, 0, 2)                   # no persistent kernel for the growing-form modes
    with pytest.raises(ValueError):
        kd_tree.build_tree_dealt(func, p, LO, HI, 6, 7, 0, 2)                  # deal below the split depth


@pytest.mark.parametrize("name,mode,kw", [
    ("fox", "affine_fixed", dict(split_depth=11, with_interior_nodes=True, with_exterior_nodes=True, batch_process_size=128)),
    ("bunny", "interval", dict(split_depth=13, with_interior_nodes=True)),
    ("hammer", "slope_interval", dict(split_depth=12, with_exterior_nodes=True, offset=0.01)),
    ("birdcage_occ", "affine_fixed", dict(node_terminate_thresh=3000, with_interior_nodes=True, with_exterior_nodes=True)),
    ("bunny", "affine_fixed", dict(split_depth=17, batch_process_size=4096)),
    ("fox", "affine_fixed", dict(split_depth=2)),
])
def test_tree_persistent_kernel_equals_level_loop(monkeypatch, name, mode, kw):
    """The one-launch cooperative tree kernel (csrc/niq_tree.cuh) against the per-level host loop it replaces
    (NIQ_TREE_LEGACY=1): the same engine classifies the same boxes, so every array, its ORDER, the per-level counts and the
    statistics must be bit-identical."""
    import kd_tree
    p = sample_params(name)
    func = make(p, mode)
    res = []
    for legacy in ("0", "1"):
        monkeypatch.setenv("NIQ_TREE_LEGACY", legacy)
        st = {}
        out = kd_tree.construct_uniform_unknown_levelset_tree(func, p, LO, HI, stats=st, **kw)
        res.append((out, st))
    (a, sa), (b, sb) = res
    assert sa == sb, (sa, sb)
    assert sorted(a) == sorted(b)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    assert int(a["unknown_node_valid"].sum()) > 0


@pytest.mark.parametrize("layers,act,mode,depth", [
    ([3, 128, 128, 1], "relu", "affine_fixed", 13),        # resident weights + zero-skipping lists + the one-tile engine
    ([3, 256, 256, 1], "relu", "affine_fixed", 12),        # 256-wide, still resident
    ([3, 256, 256, 256, 256, 1], "relu", "interval", 11),  # streamed ring shared by the two engines
    ([3, 128, 128, 128, 1], "elu", "affine_fixed", 11),    # 128-wide without zero-skipping (elu)
])
def test_tree_small_levels_on_the_one_tile_engine(monkeypatch, layers, act, mode, depth):
    """The tree kernel runs small levels on a second engine with ONE tile per thread (niq_tree.cuh; for streamed weights the
    two engines hand the ring position back and forth, for the zero-skipping width classes the one-tile engine's lists are
    re-initialised before every use).  Random-init nets of the width classes the sample MLPs do not reach: the persistent
    kernel must equal the per-level host loop bit for bit (arrays, order, per-level counts), and a dealt build -- whose
    frontier SHRINKS at the deal level, so the one-tile engine is used again after the two-tile one -- must partition it."""
    import kd_tree
    import mlp
    p = mlp.initialize_params(mlp.build_spec(mlp.quick_mlp_spec(layers, act)), 3)
    func = make(p, mode)
    res = []
    for legacy in ("0", "1"):
        monkeypatch.setenv("NIQ_TREE_LEGACY", legacy)
        st = {}
        out = kd_tree.construct_uniform_unknown_levelset_tree(func, p, LO, HI, split_depth=depth, stats=st)
        res.append((out, st))
    monkeypatch.setenv("NIQ_TREE_LEGACY", "0")
    (a, sa), (b, sb) = res
    assert sa == sb, (sa, sb)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    v = a["unknown_node_valid"]
    assert int(v.sum()) > 0
    want = _canon(a["unknown_node_lower"][v], a["unknown_node_upper"][v])
    parts = []
    for rank in range(4):
        t = kd_tree.build_tree_dealt(func, p, LO, HI, depth, depth - 3, rank, 4)
        try:
            parts.append(t.nodes(0))
        finally:
            t.close()
    got = _canon(np.concatenate([x for x, _ in parts]), np.concatenate([y for _, y in parts]))
    assert got.shape == want.shape and np.array_equal(got, want)


def test_tree_persistent_kernel_multi_root_and_growth(monkeypatch):
    """Multi-root builds (the unit of the subtree sharding) through the cooperative kernel, and a build whose frontier
    outgrows the first buffer (relaunch from the level that did not fit): both equal the level loop."""
    import kd_tree
    p = sample_params("bunny")
    func = make(p, "affine_fixed")
    top = kd_tree.construct_uniform_unknown_levelset_tree(func, p, LO, HI, split_depth=7)
    tv = top["unknown_node_valid"]
    rl, ru = top["unknown_node_lower"][tv], top["unknown_node_upper"][tv]
    outs = []
    for legacy in ("0", "1"):
        monkeypatch.setenv("NIQ_TREE_LEGACY", legacy)
        tree = kd_tree.build_tree(func, p, rl, ru, split_depth=9, with_interior_nodes=True)
        try:
            outs.append((tree.nodes(0), tree.nodes(1), tree.stats(), tree.level_info()))
        finally:
            tree.close()
    for x, y in zip(outs[0][:2], outs[1][:2]):
        np.testing.assert_array_equal(x[0], y[0])
        np.testing.assert_array_equal(x[1], y[1])
    assert outs[0][2] == outs[1][2] and outs[0][3] == outs[1][3]
    # growth: NIQ_TREE_CAP (test knob) makes the first buffers 3,000 nodes, so the frontier and the exterior list outgrow them
    monkeypatch.setenv("NIQ_TREE_LEGACY", "0")
    monkeypatch.setenv("NIQ_TREE_CAP", "3000")
    st = {}
    deep = kd_tree.construct_uniform_unknown_levelset_tree(func, p, LO, HI, split_depth=20, with_exterior_nodes=True, stats=st)
    monkeypatch.setenv("NIQ_TREE_LEGACY", "1")
    st2 = {}
    deep2 = kd_tree.construct_uniform_unknown_levelset_tree(func, p, LO, HI, split_depth=20, with_exterior_nodes=True, stats=st2)
    assert st == st2 and int(deep["exterior_node_valid"].sum()) > 65536 and st["max_frontier"] > 100000
    for k in deep:
        np.testing.assert_array_equal(deep[k], deep2[k], err_msg=k)


@pytest.mark.parametrize("case,name", [("mc_fox_d4_s2", "fox"), ("mc_bunny_d4_s3", "bunny")])
def test_marching_cubes_golden(case, name):
    import kd_tree
    g = golden(case)
    p = sample_params(name)
    tri = kd_tree.hierarchical_marching_cubes(make(p, "affine_fixed"), p, LO, HI, int(g["depth"]), n_subcell_depth=int(g["n_sub"]))
    assert tri.shape == g["tri_pos"].shape and tri.dtype == np.float32
    np.testing.assert_allclose(tri, g["tri_pos"], rtol=0, atol=2e-5)


def test_marching_cubes_vs_oracle_and_properties():
    import extract_cell
    import kd_tree
    import mlp
    from niq_oracle import mc
    p = sample_params("bunny")
    func = make(p, "affine_fixed")
    tri = kd_tree.hierarchical_marching_cubes(func, p, LO, HI, 6, n_subcell_depth=3)
    otri = otree.hierarchical_marching_cubes(octx("affine_fixed"), p, LO, HI, 6, n_subcell_depth=3)
    assert tri.shape == otri.shape and tri.shape[0] > 10000
    np.testing.assert_allclose(tri, otri, rtol=0, atol=2e-5)
    # vertices lie near the level set: |f| small relative to the local gradient scale
    f = mlp.eval_points(p, tri.reshape(-1, 3))
    assert np.percentile(np.abs(f), 99) < 0.02
    # single-cell API keeps the reference's padded layout
    mc_data = extract_cell.get_mc_data()
    tp, tv = extract_cell.extract_triangles_from_subcells(func, p, mc_data, 2, np.array([-.25, -.25, -.25], np.float32),
                                                          np.array([.25, .25, .25], np.float32))
    assert tp.shape == (5 * 64, 3, 3) and tv.shape == (5 * 64,)
    ref = mc.extract_mesh_from_leaves(p, np.array([[-.25, -.25, -.25]], np.float32), np.array([[.25, .25, .25]], np.float32), 2)
    np.testing.assert_allclose(tp[tv], ref, rtol=0, atol=2e-5)


# ---------------------------------------------------------------------------------------------------
# find_any_intersection, closest_point
# ---------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("case,mode", [("isect_fixed", "affine_fixed"), ("isect_trunc64", "affine_truncate")])
def test_find_any_intersection_golden(case, mode):
    import kd_tree
    import mlp
    g = golden(case)
    pA = sample_params("hammer")
    pB = mlp.prepend_op(sample_params("bunny"), mlp.spatial_transformation())
    fA, fB = make(pA, mode, g["n_trunc"]), make(pB, mode, g["n_trunc"])
    n_cmp = n_flag = n_bad = 0
    for i in range(g["R"].shape[0]):
        pB["0000.spatial_transformation.R"] = g["R"][i]
        pB["0000.spatial_transformation.t"] = g["t"][i]
        st = {}
        found, ia, ib, loc = kd_tree.find_any_intersection((fA, fB), (pA, pB), LO, HI, float(g["eps"]), stats=st)
        n_cmp += 1
        if st["n_near_tie"] == 0:
            assert bool(found) == bool(g["found"][i])
            np.testing.assert_allclose(loc, g["loc"][i], rtol=0, atol=1e-6)
        else:
            n_flag += 1
            n_bad += bool(found) != bool(g["found"][i])
        assert (ia, ib) == ((1, 2) if found else (0, 0))
    parity_report(f"find_any_intersection_golden[{case}]", queries=n_cmp, flagged=n_flag, flagged_verdict_mismatch=n_bad)
    assert n_flag <= max(1, n_cmp // 2) and n_bad == 0


def test_find_any_intersection_vs_oracle_list():
    """Seeded rigid transforms (config 3 recipe, affine_fixed for oracle speed): verdict + location parity."""
    import kd_tree
    import mlp
    rng = np.random.default_rng(0)
    pA = sample_params("hammer")
    pB = mlp.prepend_op(sample_params("bunny"), mlp.spatial_transformation())
    fA, fB = make(pA, "affine_fixed"), make(pB, "affine_fixed")
    n_found = n_flag = n_bad = n_diff = n_nodes = n_tie = 0
    for i in range(12):
        th = rng.uniform(0, 2 * np.pi)
        R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]], np.float32)
        t = rng.uniform(-1.5, 1.5, 3).astype(np.float32)
        pB["0000.spatial_transformation.R"] = R
        pB["0000.spatial_transformation.t"] = t
        opB = net.prepend_op(sample_params("bunny"), net.spatial_transformation(R, t))
        st, ost = {}, {}
        found, _, _, loc = kd_tree.find_any_intersection((fA, fB), (pA, pB), LO, HI, 1e-3, stats=st)
        ofound, _, _, oloc = otree.find_any_intersection((octx("affine_fixed"),) * 2, (pA, opB), LO, HI, 1e-3, stats=ost)
        same = (bool(found) == bool(ofound) and st["n_nodes"] == ost["n_nodes"] and st["n_rounds"] == ost["n_rounds"]
                and np.allclose(loc, oloc, rtol=0, atol=1e-6))
        flagged = st["n_near_tie"] > 0 or ost["n_near_tie"] > 0
        assert same or flagged, f"transform {i}: differs with no near-tie box on either side"
        n_flag += flagged
        n_diff += not same
        n_bad += bool(found) != bool(ofound)
        n_nodes += ost["n_nodes"]
        n_tie += max(st["n_near_tie"], ost["n_near_tie"])
        n_found += bool(found)
    parity_report("find_any_intersection_vs_oracle_list", queries=12, nodes=n_nodes, near_tie_boxes=n_tie, queries_flagged=n_flag,
                  flagged_differing=n_diff, verdict_mismatch=n_bad)
    # every query is compared; a difference is tolerated only where a box sat inside the band, and those boxes stay rare
    assert n_tie <= 0.01 * n_nodes and n_bad == 0 and n_diff <= 2
    assert 0 < n_found < 12


@pytest.mark.parametrize("case", ["closest_fox_B4", "closest_fox_B256"])
def test_closest_point_golden(case):
    import kd_tree
    g = golden(case)
    p = sample_params("fox")
    d, loc = kd_tree.closest_point(make(p, "affine_fixed"), p, LO, HI, g["query_points"], eps=float(g["eps"]),
                                   batch_process_size=int(g["B"]))
    np.testing.assert_allclose(d, g["dist"], rtol=RTOL)
    fin = np.isfinite(g["dist"])
    assert fin.any()
    np.testing.assert_allclose(loc[fin], g["loc"][fin], rtol=0, atol=1e-6)


@pytest.mark.parametrize("B", [64, 2048, 2 ** 14])
def test_closest_point_vs_oracle(B):
    import kd_tree
    p = sample_params("fox")
    q = np.random.default_rng(4).uniform(-1, 1, (24, 3)).astype(np.float32)
    st, ost = {}, {}
    d, loc = kd_tree.closest_point(make(p, "affine_fixed"), p, LO, HI, q, eps=0.01, batch_process_size=B, stats=st)
    od, oloc = otree.closest_point(octx("affine_fixed"), p, LO, HI, q, eps=0.01, batch_process_size=B, stats=ost)
    ok = np.abs(d - od) <= RTOL * od
    ok |= ~np.isfinite(od) & ~np.isfinite(d)
    parity_report(f"closest_point_vs_oracle[B={B}]", queries=24, near_tie_boxes_gpu=st["n_near_tie"], near_tie_boxes_oracle=ost["n_near_tie"],
                  visits=ost["n_visits"], dist_within_1e5=int(ok.sum()))
    if st["n_near_tie"] == 0 and ost["n_near_tie"] == 0:
        assert st["n_visits"] == ost["n_visits"] and st["n_rounds"] == ost["n_rounds"]
        np.testing.assert_allclose(d, od, rtol=RTOL)
        fin = np.isfinite(od)
        np.testing.assert_allclose(loc[fin], oloc[fin], rtol=0, atol=1e-6)
    else:
        # a label flipped inside the band re-orders the shared LIFO stack: the flagged boxes must stay rare
        # and the distances must still agree for most queries
        assert st["n_near_tie"] + ost["n_near_tie"] <= 1e-3 * ost["n_visits"] + 2
        assert ok.mean() > 0.7


@pytest.mark.parametrize("name,mode,B,Q,eps", [("fox", "affine_fixed", 256, 40, 0.01), ("birdcage_occ", "affine_fixed", 2048, 24, 1e-3),
                                               ("fox", "interval", 64, 12, 0.05), ("hammer", "affine_fixed", 2048, 3000, 0.02)])
def test_closest_point_persistent_kernel_equals_graph_loop(monkeypatch, name, mode, B, Q, eps):
    """The one-launch cooperative search (csrc/niq_cp.cuh: every round of the LIFO window inside the kernel) against the
    CUDA-graph round loop it replaces (NIQ_CP_LEGACY=1): same engine passes, same round logic -> distances, locations and
    statistics identical.  Q = 3000 > window: the stack starts longer than the window; eps 0.02 keeps it short."""
    import kd_tree
    p = sample_params(name)
    q = np.random.default_rng(8).uniform(-1, 1, (Q, 3)).astype(np.float32)
    res = []
    for legacy in ("0", "1"):
        monkeypatch.setenv("NIQ_CP_LEGACY", legacy)
        st = {}
        d, loc = kd_tree.closest_point(make(p, mode), p, LO, HI, q, eps=eps, batch_process_size=B, stats=st)
        res.append((d, loc, st))
    (d0, l0, s0), (d1, l1, s1) = res
    assert s0 == s1, (s0, s1)
    np.testing.assert_array_equal(d0, d1)
    fin = np.isfinite(d1)
    np.testing.assert_array_equal(l0[fin], l1[fin])
    assert fin.any() and s0["n_rounds"] > 3


# ---------------------------------------------------------------------------------------------------
# tree consumers
# ---------------------------------------------------------------------------------------------------

def test_sample_surface_and_bulk_properties_vs_oracle():
    """kd_tree.sample_surface / bulk_properties (src/kd_tree.py:220-292, 804-863) with the same NumPy generator as the
    oracle: the draws are identical, so the accepted samples agree one by one except where |f| is within the
    point-value band of the acceptance threshold; mass / centroid agree to Monte-Carlo-free precision."""
    import kd_tree
    p = sample_params("fox")
    func = make(p, "affine_fixed")
    width = 0.01
    pts = kd_tree.sample_surface(func, p, LO, HI, 3000, width, 7)
    opts = otree.sample_surface(octx("affine_fixed"), p, LO, HI, 3000, width, np.random.default_rng(7))
    f = net.eval_points(p, pts)
    assert np.all(np.abs(f) < width + 1e-5 * rays.point_scale(p, pts))
    same = np.all(pts == opts, axis=1)
    assert same.mean() > 0.99 or np.array_equal(pts[:100], opts[:100])
    # the tree-free sibling (src/kd_tree.py:296-336): every sample inside the band and inside the domain
    upts = kd_tree.sample_surface_uniform(func, p, LO, HI, 500, width, 7)
    assert upts.shape == (500, 3) and upts.dtype == np.float32 and np.all(np.abs(upts) <= 1.0)
    assert np.all(np.abs(net.eval_points(p, upts)) < width + 1e-5 * rays.point_scale(p, upts))
    mass, cen = kd_tree.bulk_properties(func, p, LO, HI, 11, n_expand=2000, n_sample=200000)
    omass, ocen = otree.bulk_properties(octx("affine_fixed"), p, LO, HI, np.random.default_rng(11), n_expand=2000, n_sample=200000)
    assert abs(mass - omass) <= 2e-4 * omass and np.all(np.abs(cen - ocen) <= 2e-4)
    assert 0.01 < mass < 8.0 and np.all(np.abs(cen) < 1.0)


def test_marching_cubes_shared_faces_evaluated_once(monkeypatch):
    """Lattice points on a face shared by two leaves are evaluated once (k_mc_neighbours / mc_val): fewer evaluations, and
    the triangle soup is BIT-identical to one evaluation per leaf and point (NIQ_MC_NO_DEDUP=1), for a uniform-depth tree,
    for caller-supplied leaves of mixed sizes (no exact face match -> nothing shared) and across the golden case."""
    import _niq
    import extract_cell
    import kd_tree
    ctx = _niq.default_context()
    p = sample_params("bunny")
    func = make(p, "affine_fixed")
    ctx.mc_points(reset=True)
    tri = kd_tree.hierarchical_marching_cubes(func, p, LO, HI, 7, n_subcell_depth=3)
    ev, lat = ctx.mc_points(reset=True)
    assert lat == 4096 * 729 or lat % 729 == 0
    assert 0.6 * lat < ev < 0.9 * lat, (ev, lat)
    monkeypatch.setenv("NIQ_MC_NO_DEDUP", "1")
    tri_ref = kd_tree.hierarchical_marching_cubes(func, p, LO, HI, 7, n_subcell_depth=3)
    ev2, lat2 = ctx.mc_points(reset=True)
    assert ev2 == lat2 == lat
    monkeypatch.delenv("NIQ_MC_NO_DEDUP")
    assert tri.shape == tri_ref.shape and tri.shape[0] > 50000
    np.testing.assert_array_equal(tri, tri_ref)
    # mixed leaf sizes / a duplicate leaf / non-dyadic bounds through the leaf-list entry point
    lo = np.array([[-.25, -.25, -.25], [0.0, -.25, -.25], [0.25, -.25, -.25], [-.25, -.25, -.25], [0.1, 0.05, -0.3], [0.35, 0.05, -0.3]], np.float32)
    hi = np.array([[0.0, 0.0, 0.0], [0.25, 0.0, 0.0], [0.75, 0.25, 0.25], [0.0, 0.0, 0.0], [0.35, 0.3, -0.05], [0.6, 0.3, -0.05]], np.float32)
    a = extract_cell.extract_mesh_from_cells(func, p, lo, hi, 2)
    monkeypatch.setenv("NIQ_MC_NO_DEDUP", "1")
    b = extract_cell.extract_mesh_from_cells(func, p, lo, hi, 2)
    monkeypatch.delenv("NIQ_MC_NO_DEDUP")
    np.testing.assert_array_equal(a, b)


def test_marching_cubes_sharded_entry_point_world1():
    """sharding.hierarchical_marching_cubes_sharded on one device (top levels, then ONE multi-root build of the frontier,
    marching cubes over its leaves): the same triangles as kd_tree.hierarchical_marching_cubes, in a different order."""
    import kd_tree
    import sharding
    p = sample_params("bunny")
    func = make(p, "affine_fixed")
    tri = kd_tree.hierarchical_marching_cubes(func, p, LO, HI, 6, n_subcell_depth=3)
    stri = sharding.hierarchical_marching_cubes_sharded(func, p, LO, HI, 6, n_subcell_depth=3, top_depth=5)
    canon = lambda a: np.unique(a.reshape(-1, 9), axis=0)
    assert stri.shape == tri.shape and tri.shape[0] > 10000
    np.testing.assert_array_equal(canon(stri), canon(tri))


# ---------------------------------------------------------------------------------------------------
# the caller of cast_rays: render.render_image
# ---------------------------------------------------------------------------------------------------

def test_render_image_golden_and_oracle():
    """render.render_image (src/render.py:94-150; frustum=False, shading='normal'): hit ids / counts / depth as cast_rays,
    colours from finite-difference normals (4 GPU point evaluations eps=1e-3 apart per hit: tolerance 2e-3 absolute =
    the point-value band over the difference magnitude)."""
    import queries
    import render
    import _niq
    p = sample_params("fox")
    func = make(p, "affine_fixed")
    g = golden("render_fox_fixed_r10")
    eye = np.array((2., 1., 2.), np.float32)
    look, up, left = render.look_at(eye)
    opts = queries.get_default_cast_opts()
    img, depth, cnt, hit, n_eval, _ = render.render_image(func, p, eye, look, up, left, int(g["res"]), 30.0, False, opts)
    assert img.shape == (10, 10, 3) and img.dtype == np.float32 and hit.dtype == np.int32
    np.testing.assert_array_equal(hit, g["hit_ids"])
    np.testing.assert_array_equal(cnt, g["counts"])
    assert n_eval == int(g["n_eval"])
    np.testing.assert_allclose(depth, g["depth"], rtol=RTOL, atol=0)
    np.testing.assert_allclose(img, g["img"], rtol=0, atol=2e-3)
    # a larger image against the oracle
    res = 64
    img, depth, cnt, hit, n_eval, _ = render.render_image(func, p, eye, look, up, left, res, 30.0, False, opts)
    oimg, odepth, ocnt, ohit, on = rays.render_image((octx("affine_fixed"),), (p,), eye, look, up, res, 30.0, opts)
    same = (hit == ohit) & (cnt == ocnt)
    assert same.mean() > 0.995
    np.testing.assert_allclose(depth[same], odepth[same], rtol=RTOL, atol=0)
    np.testing.assert_allclose(img[same], oimg[same], rtol=0, atol=2e-3)
    x = np.linspace(0, 1, 33, dtype=np.float32)                      # tonemap (src/render.py:152-158) is a pure formula
    ref = ((x.astype(np.float64) * (1 + x.astype(np.float64) / 0.75 ** 2)) / (1 + x.astype(np.float64))) ** (1 / 2.2)
    np.testing.assert_allclose(render.tonemap_image(x), ref, rtol=1e-5, atol=1e-7)
    with pytest.raises(ValueError):
        render.render_image((func, func), (p,), eye, look, up, left, 8, 30.0, False, opts)


# ---------------------------------------------------------------------------------------------------
# cast_rays_frustum
# ---------------------------------------------------------------------------------------------------
FRUSTUM_CASES = {
    "frust_fox_fixed_r12_s4": (("fox",), "affine_fixed"),
    "frust_hammer_fixed_r9_s3_sub3": (("hammer",), "affine_fixed"),
    "frust_fox_bunny_interval_r10_s2_sub2": (("fox", "bunny"), "interval"),
    "frust_fox_slope_r10_s2": (("fox",), "slope_interval"),          # persistent kernel over the 9-row slope tile
    "frust_fox_trunc_r8_s2": (("fox",), "affine_truncate"),          # host-level loop over the grow kernel
}


def _frustum_inputs(g):
    import queries
    opts = queries.get_default_cast_opts()
    opts["n_side_init"] = int(g["n_side"])
    opts["n_substeps"] = int(g["n_substeps"])
    res = int(g["res"])
    res_y = int(g["res_y"]) if "res_y" in g else res
    fov_y = float(g["fov_y"]) if "fov_y" in g else 30.0
    return (g["eye"], g["look"], g["up"], g["left"], 30.0, fov_y, res, res_y), opts


@pytest.mark.parametrize("case", sorted(FRUSTUM_CASES))
def test_cast_rays_frustum_golden(case):
    """The persistent frustum kernel (k_cast_frustum) and the host-level loop, both against the unmodified reference."""
    import queries
    names, mode = FRUSTUM_CASES[case]
    g = golden(case)
    cam, opts = _frustum_inputs(g)
    ps = tuple(sample_params(n) for n in names)
    funcs = tuple(make(p, mode, int(g.get("n_trunc", 8))) for p in ps)
    import _niq
    host_loop = lambda *a: queries._cast_rays_frustum_host_loop(_niq.default_context(), *a)
    for impl in (queries.cast_rays_frustum, host_loop):
        t, hit, cnt, n_evals, tie = impl(funcs, ps, cam, opts, True)
        assert t.shape == g["out_t"].shape and t.dtype == np.float32 and hit.dtype == np.int32 and cnt.dtype == np.int32
        ok = ~tie          # ~490 crawling steps x 2 funcs per pixel in the interval case: many chains touch the band once
        assert ok.mean() > (0.3 if mode == "interval" else 0.6)
        np.testing.assert_array_equal(hit[ok], g["out_hit_id"][ok])
        np.testing.assert_array_equal(cnt[ok], g["out_count"][ok])
        np.testing.assert_allclose(t[ok], g["out_t"][ok], rtol=RTOL, atol=0)
        if not tie.any():
            assert n_evals == int(g["n_evals"])


@pytest.mark.parametrize("name,mode,res,n_side,n_sub", [("fox", "affine_fixed", 64, 16, 1), ("bunny", "affine_fixed", 40, 8, 1),
                                                        ("birdcage_occ", "interval", 24, 4, 2), ("fox", "affine_truncate", 16, 4, 1),
                                                        ("fox", "slope_interval", 16, 4, 1), ("bunny", "slope_interval", 32, 8, 2)])
def test_cast_rays_frustum_vs_oracle(name, mode, res, n_side, n_sub):
    """Larger images against the oracle.  A pixel is compared when no decision of its frustum chain was inside the 1e-5
    band in EITHER implementation (a flipped split changes the whole subtree of pixels)."""
    import queries
    import render
    p = sample_params(name)
    func = make(p, mode)
    eye = np.array((2., 1., 2.), np.float32)
    look, up, left = render.look_at(eye)
    opts = queries.get_default_cast_opts()
    opts["n_side_init"] = n_side
    opts["n_substeps"] = n_sub
    cam = (eye, look, up, left, 30.0, 30.0, res, res)
    t, hit, cnt, n_evals, tie = queries.cast_rays_frustum((func,), (p,), cam, opts, return_near_tie=True)
    ot, ohit, ocnt, on, otie = rays.cast_rays_frustum((octx(mode),), (p,), cam, opts, return_near_tie=True)
    ok = ~(tie | otie)          # a child inherits the flag of its whole chain of parents; bunny (ELU) has the wider band
    assert ok.mean() > 0.8
    np.testing.assert_array_equal(hit[ok], ohit[ok])
    np.testing.assert_array_equal(cnt[ok], ocnt[ok])
    np.testing.assert_allclose(t[ok], ot[ok], rtol=RTOL, atol=0)
    if mode != "interval":      # the loose interval bounds crawl from the start: every pixel ends on the step limit
        assert (hit != 0).any() and (hit == 0).any()
    if not (tie | otie).any():
        assert n_evals == on


@pytest.mark.parametrize("mode,n_trunc,n_sub", [("affine_truncate", 8, 2), ("affine_all", 8, 1), ("affine_append", 4, 1)])
def test_cast_rays_frustum_grow_kernel_equals_host_loop(monkeypatch, mode, n_trunc, n_sub):
    """The persistent frustum kernel of the growing-form modes (csrc/niq_frustum_grow.cuh: one CTA per frustum, device work
    queue) against the host-level iteration it replaces (NIQ_RAYS_HOST_LOOP=1): images, N_evals identical; two funcs."""
    import queries
    import render
    pf, pb = sample_params("fox"), sample_params("bunny")
    funcs = (make(pf, mode, n_trunc), make(pb, mode, n_trunc))
    eye = np.array((2., 1., 2.), np.float32)
    look, up, left = render.look_at(eye)
    opts = queries.get_default_cast_opts()
    opts["n_side_init"] = 4
    opts["n_substeps"] = n_sub
    cam = (eye, look, up, left, 30.0, 30.0, 24, 20)
    res = []
    for host in ("", "1"):
        if host:
            monkeypatch.setenv("NIQ_RAYS_HOST_LOOP", host)
        else:
            monkeypatch.delenv("NIQ_RAYS_HOST_LOOP", raising=False)
        res.append(queries.cast_rays_frustum(funcs, (pf, pb), cam, opts, return_near_tie=True))
    a, b = res
    for x, y in zip(a[:4], b[:4]):                 # t, hit_id, count images, N_evals
        np.testing.assert_array_equal(x, y)
    assert not (b[4] & ~a[4]).any()                # the kernel bands with the widest tie_rel of the call: a superset
    assert (a[1] != 0).any() and (a[1] == 0).any()


def test_cast_rays_frustum_wide_streamed_and_render():
    """A 256-wide net (weights streamed through the ring: every warp of a CTA leaves together) against the oracle, the
    frustum branch of render.render_image, and argument errors."""
    import queries
    import render
    p = net.random_mlp([3, 256, 256, 256, 1], "relu", seed=5)
    func = make(p, "affine_fixed")
    eye = np.array((2., 1., 2.), np.float32)
    look, up, left = render.look_at(eye)
    opts = queries.get_default_cast_opts()
    opts["n_side_init"] = 4
    opts["n_max_step"] = 24
    cam = (eye, look, up, left, 30.0, 30.0, 20, 20)
    t, hit, cnt, n_evals, tie = queries.cast_rays_frustum((func,), (p,), cam, opts, return_near_tie=True)
    ot, ohit, ocnt, on, otie = rays.cast_rays_frustum((octx("affine_fixed"),), (p,), cam, opts, return_near_tie=True)
    ok = ~(tie | otie)
    assert ok.mean() > 0.9
    np.testing.assert_array_equal(hit[ok], ohit[ok])
    np.testing.assert_array_equal(cnt[ok], ocnt[ok])
    np.testing.assert_allclose(t[ok], ot[ok], rtol=RTOL, atol=0)

    pf = sample_params("fox")
    ff = make(pf, "affine_fixed")
    o2 = queries.get_default_cast_opts()
    img, depth, counts, hit_ids, n_eval, _ = render.render_image(ff, pf, eye, look, up, left, 32, 30.0, True, o2)
    t2, h2, c2, n2 = queries.cast_rays_frustum((ff,), (pf,), (eye, look, up, left, 30.0, 30.0, 32, 32), o2)
    np.testing.assert_array_equal(hit_ids, h2.T)            # render transposes the (res_x, res_y) images (src/render.py:124-126)
    np.testing.assert_array_equal(depth, t2.T)
    assert img.shape == (32, 32, 3) and n_eval == n2 and np.all(img[hit_ids == 0] == 1.0)
    with pytest.raises(ValueError):
        queries.cast_rays_frustum((ff,), (pf,), (eye, look, up, left, 30.0, 30.0, 8, 8), o2)      # n_side_init 16 > res
    with pytest.raises(ValueError):
        queries.cast_rays_frustum((ff, ff), (pf,), (eye, look, up, left, 30.0, 30.0, 32, 32), o2)


def test_cast_rays_frustum_tile_shares_equal_whole_image():
    """The multi-GPU partition of frustum casting (sharding.cast_rays_frustum_sharded) on one device: marching the initial
    tiles in two disjoint shares and summing the images / iteration counts gives the single-call result bit for bit."""
    import queries
    import render
    import sharding
    p = sample_params("fox")
    func = make(p, "affine_fixed")
    eye = np.array((2., 1., 2.), np.float32)
    look, up, left = render.look_at(eye)
    opts = queries.get_default_cast_opts()
    opts["n_side_init"] = 8
    cam = (eye, look, up, left, 30.0, 30.0, 48, 40)
    t, hit, cnt, n_evals = queries.cast_rays_frustum((func,), (p,), cam, opts)
    init = queries._initial_frusta(48, 40, 8)
    acc = [np.zeros((48, 40), np.float32), np.zeros((48, 40), np.int32), np.zeros((48, 40), np.int32)]
    n_bins = 512 + 3
    counts = np.zeros((2, n_bins), np.int64)
    for r in range(2):
        it = []
        part = queries.cast_rays_frustum((func,), (p,), cam, opts, init_ranges=init[r::2], iter_counts=it)
        for a, b in zip(acc, part[:3]):
            a += b
        for k, (x, y) in enumerate(it):
            counts[0, k] += x; counts[1, k] += y
    np.testing.assert_array_equal(acc[0], t)
    np.testing.assert_array_equal(acc[1], hit)
    np.testing.assert_array_equal(acc[2], cnt)
    assert queries._frustum_n_evals(64, counts[0].tolist(), counts[1].tolist()) == n_evals
    # world size 1 through the sharding entry point
    st, sh, sc, sn = sharding.cast_rays_frustum_sharded((func,), (p,), cam, opts)
    np.testing.assert_array_equal(st, t)
    np.testing.assert_array_equal(sh, hit)
    assert sn == n_evals and (hit != 0).any()


def test_render_image_frustum_branch_golden():
    """render.render_image(frustum=True) against the unmodified reference (transposition into ray order, normals, shading)."""
    import queries
    import render
    g = golden("render_frustum_fox_fixed_r14")
    p = sample_params("fox")
    func = make(p, "affine_fixed")
    eye = np.array((2., 1., 2.), np.float32)
    look, up, left = render.look_at(eye)
    opts = queries.get_default_cast_opts()
    opts["n_side_init"] = int(g["n_side"])
    img, depth, cnt, hit, n_eval, _ = render.render_image(func, p, eye, look, up, left, int(g["res"]), 30.0, True, opts)
    np.testing.assert_array_equal(hit, g["hit_ids"])
    np.testing.assert_array_equal(cnt, g["counts"])
    assert n_eval == int(g["n_eval"])
    np.testing.assert_allclose(depth, g["depth"], rtol=RTOL, atol=0)
    np.testing.assert_allclose(img, g["img"], rtol=0, atol=2e-3)


def test_cast_rays_frustum_uneven_tiles_non_square():
    """res_x != res_y and a tile count that divides neither (initial tiles of 3-4 x 2-3 pixels) against the oracle."""
    import queries
    import render
    p = sample_params("fox")
    func = make(p, "affine_fixed")
    eye = np.array((2., 1., 2.), np.float32)
    look, up, left = render.look_at(eye)
    opts = queries.get_default_cast_opts()
    cam = (eye, look, up, left, 30.0, 30.0, 50, 37)
    t, hit, cnt, n_evals, tie = queries.cast_rays_frustum((func,), (p,), cam, opts, return_near_tie=True)
    ot, ohit, ocnt, on, otie = rays.cast_rays_frustum((octx("affine_fixed"),), (p,), cam, opts, return_near_tie=True)
    assert t.shape == (50, 37)
    ok = ~(tie | otie)
    assert ok.mean() > 0.9 and (hit != 0).any()
    np.testing.assert_array_equal(hit[ok], ohit[ok])
    np.testing.assert_array_equal(cnt[ok], ocnt[ok])
    np.testing.assert_allclose(t[ok], ot[ok], rtol=RTOL, atol=0)
    if not (tie | otie).any():
        assert n_evals == on


def test_cast_rays_frustum_non_square_two_fov_golden():
    """res_x != res_y and fov_x != fov_y (13x9 pixels, 30 / 22 degrees) against the unmodified reference: pins which axis every
    camera constant belongs to.  Kernel and host-level loop."""
    import _niq
    import queries
    g = golden("frust_fox_fixed_r13x9_s3")
    cam, opts = _frustum_inputs(g)
    assert cam[6] == 13 and cam[7] == 9 and cam[5] == 22.0
    p = sample_params("fox")
    funcs = (make(p, "affine_fixed"),)
    host_loop = lambda *a: queries._cast_rays_frustum_host_loop(_niq.default_context(), *a)
    for impl in (queries.cast_rays_frustum, host_loop):
        t, hit, cnt, n_evals, tie = impl(funcs, (p,), cam, opts, True)
        assert t.shape == (13, 9)
        ok = ~tie
        assert ok.mean() > 0.6
        np.testing.assert_array_equal(hit[ok], g["out_hit_id"][ok])
        np.testing.assert_array_equal(cnt[ok], g["out_count"][ok])
        np.testing.assert_allclose(t[ok], g["out_t"][ok], rtol=RTOL, atol=0)
        if not tie.any():
            assert n_evals == int(g["n_evals"])


# ---------------------------------------------------------------------------------------------------
# the C ABI from plain C (no Python struct mirrors in between)
# ---------------------------------------------------------------------------------------------------

def test_c_abi_program(tmp_path):
    """tests/abi_c_test.c includes include/niq.h, links libniq.so, builds the fox MLP from a raw weight blob and runs
    niq_classify_boxes + niq_cast_rays; its printed results equal the ctypes binding's bit for bit."""
    import os
    import struct
    import subprocess
    import queries
    import render
    from conftest import PKG, ROOT
    p = sample_params("fox")
    dense = sorted(k for k in p if k.endswith(".dense.A"))
    lo, hi = random_boxes(21, 40)
    eye = np.array((2., 1., 2.), np.float32)
    look, up, _ = render.look_at(eye)
    roots, dirs = render.generate_camera_rays(eye, look, up, res=12, fov_deg=30.)
    blob = tmp_path / "fox.blob"
    with open(blob, "wb") as f:
        f.write(struct.pack("<i", len(dense)))
        for k in dense:
            A = np.ascontiguousarray(p[k], np.float32)
            b = np.ascontiguousarray(p[k.replace(".A", ".b")], np.float32)
            f.write(struct.pack("<ii", *A.shape)); f.write(A.tobytes()); f.write(b.tobytes())
        f.write(struct.pack("<i", lo.shape[0])); f.write(lo.tobytes()); f.write(hi.tobytes())
        f.write(struct.pack("<i", roots.shape[0])); f.write(roots.tobytes()); f.write(dirs.tobytes())
    exe = tmp_path / "abi_c_test"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "abi_c_test.c"),
                    "-L", PKG, "-lniq", f"-Wl,-rpath,{PKG}", "-o", str(exe)], check=True)
    out = subprocess.run([str(exe), str(blob)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    assert lines[-1] == "ok" and lines[-2].startswith("einval ")
    func = make(p, "affine_fixed")
    lab, blo, bup, tie = func.bound_box(p, lo, hi)
    t, hit, cnt, n_evals = queries.cast_rays((func,), (p,), roots, dirs, queries.get_default_cast_opts())
    boxes = [ln.split() for ln in lines if ln.startswith("box ")]
    rays_ = [ln.split() for ln in lines if ln.startswith("ray ")]
    assert len(boxes) == lo.shape[0] and len(rays_) == roots.shape[0]
    for i, b in enumerate(boxes):
        assert int(b[2]) == lab[i] and int(b[3], 16) == blo[i:i + 1].view(np.uint32)[0] and int(b[4], 16) == bup[i:i + 1].view(np.uint32)[0]
        assert int(b[5]) == int(tie[i])
    for i, r in enumerate(rays_):
        assert int(r[2], 16) == t[i:i + 1].view(np.uint32)[0] and int(r[3]) == hit[i] and int(r[4]) == cnt[i]
    assert int([ln for ln in lines if ln.startswith("n_evals ")][0].split()[1]) == n_evals
