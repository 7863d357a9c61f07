"""GPU parity at the SIZES of BASELINE.json's configs 3, 4 and 5 (run with -m gpu), against oracle outputs committed as
fixtures (tests/golden/cfg*.npz, made by oracle/tools/gen_config_fixtures.py: minutes of CPU time each, so the GPU box replays
them).  Decisions inside the near-tie band are compared too, COUNTED, bounded and reported (conftest.parity_report) --
nothing is skipped silently."""
import numpy as np
import pytest

from conftest import golden, parity_report, sample_params
from niq_oracle import net

pytestmark = pytest.mark.gpu

RTOL = 1e-5
LO = np.full(3, -1, np.float32)
HI = np.full(3, 1, np.float32)
LAYERS5 = [3] + [256] * 8 + [1]


def make(params, mode, **kw):
    import implicit_mlp_utils
    return implicit_mlp_utils.generate_implicit_from_params(params, mode, **kw)


# ---------------------------------------------------------------------------------------------------
# config 3: hammer x bunny under 64 seeded rigid transforms, affine_truncate (n_keep 64, 'absolute'), eps 1e-3
# (reference src/main_intersection.py:86,95-99 + src/kd_tree.py:567-655)
# ---------------------------------------------------------------------------------------------------

def test_cfg3_intersection_truncate64_list():
    import kd_tree
    import mlp
    g = golden("cfg3_isect_trunc64_list")
    n = g["R"].shape[0]
    assert n >= 64
    pA = sample_params("hammer")
    pB = mlp.prepend_op(sample_params("bunny"), mlp.spatial_transformation())
    kw = dict(affine_n_truncate=int(g["n_trunc"]), affine_truncate_policy="absolute")
    fA, fB = make(pA, "affine_truncate", **kw), make(pB, "affine_truncate", **kw)
    n_clean = n_flag = bad_clean = bad_flag = nodes_differ = 0
    for i in range(n):
        pB["0000.spatial_transformation.R"] = g["R"][i]
        pB["0000.spatial_transformation.t"] = g["t"][i]
        st = {}
        found, ia, ib, loc = kd_tree.find_any_intersection((fA, fB), (pA, pB), LO, HI, float(g["eps"]), stats=st)
        assert (ia, ib) == ((1, 2) if found else (0, 0))
        same = bool(found) == bool(g["found"][i])
        if st["n_near_tie"] == 0 and g["n_near_tie"][i] == 0:
            n_clean += 1
            bad_clean += not same
            assert same, f"transform {i}: verdict differs with no near-tie box on either side"
            assert st["n_nodes"] == g["n_nodes"][i] and st["n_rounds"] == g["n_rounds"][i], f"transform {i}: traversal differs"
            np.testing.assert_allclose(loc, g["loc"][i], rtol=0, atol=1e-6)
        else:
            n_flag += 1
            bad_flag += not same
            nodes_differ += st["n_nodes"] != g["n_nodes"][i]
    parity_report("cfg3_intersection_truncate64", transforms=n, found=int(g["found"].sum()), clean=n_clean, flagged=n_flag,
                  flagged_verdict_mismatch=bad_flag, flagged_traversal_differs=nodes_differ)
    assert n_flag <= n // 2, "more than half of the queries touch a near-tie box"
    assert bad_flag <= 1, "verdicts of flagged queries may differ only exceptionally"


def _cfg3_pair(mode="affine_truncate", **kw):
    import mlp
    pA = sample_params("hammer")
    pB = mlp.prepend_op(sample_params("bunny"), mlp.spatial_transformation())
    if mode == "affine_truncate":
        kw = dict(affine_n_truncate=64, affine_truncate_policy="absolute")
    return make(pA, mode, **kw), make(pB, mode, **kw), pA, pB


@pytest.mark.parametrize("mode", ["affine_truncate", "affine_all"])
def test_intersection_persistent_kernel_equals_round_loop(monkeypatch, mode):
    """The one-launch cooperative search (csrc/niq_isect.cuh) against the per-round host loop it replaces (NIQ_ISECT_LEGACY=1):
    same propagation code, same summation order of the sample points -> verdict, location and statistics identical."""
    import kd_tree
    g = golden("cfg3_isect_trunc64_list")
    fA, fB, pA, pB = _cfg3_pair(mode)
    hit = np.nonzero(g["found"])[0]
    miss = np.nonzero(~g["found"])[0]
    picks = list(hit[:3]) + list(miss[:5]) if mode == "affine_truncate" else [hit[0], hit[1], miss[0], miss[1]]
    n_found = 0
    for i in picks:
        pB["0000.spatial_transformation.R"] = g["R"][i]
        pB["0000.spatial_transformation.t"] = g["t"][i]
        res = []
        for legacy in ("0", "1"):
            monkeypatch.setenv("NIQ_ISECT_LEGACY", legacy)
            st = {}
            out = kd_tree.find_any_intersection((fA, fB), (pA, pB), LO, HI, 1e-3, stats=st)
            res.append((bool(out[0]), out[1], out[2], np.asarray(out[3]), st))
        a, b = res
        assert a[0] == b[0] and a[1:3] == b[1:3] and a[4] == b[4], (i, a, b)
        np.testing.assert_array_equal(a[3], b[3])
        n_found += a[0]
    assert 0 < n_found < len(picks)


def test_cfg3_intersection_batch_equals_single_queries_and_fixture():
    """All 64 config-3 transforms in ONE persistent kernel (niq_find_any_intersection_batch): every query's verdict, location
    and statistics equal its own single call, and the batch agrees with the oracle fixture like the single calls do."""
    import kd_tree
    g = golden("cfg3_isect_trunc64_list")
    fA, fB, pA, pB = _cfg3_pair()
    st = {}
    found, loc = kd_tree.find_any_intersection_batch((fA, fB), (pA, pB), LO, HI, float(g["eps"]), R_B=g["R"], t_B=g["t"], stats=st)
    for i in range(0, 64, 5):
        pB["0000.spatial_transformation.R"] = g["R"][i]
        pB["0000.spatial_transformation.t"] = g["t"][i]
        s1 = {}
        f1, _, _, l1 = kd_tree.find_any_intersection((fA, fB), (pA, pB), LO, HI, float(g["eps"]), stats=s1)
        assert bool(f1) == bool(found[i]) and s1["n_nodes"] == st["n_nodes"][i] and s1["n_rounds"] == st["n_rounds"][i]
        assert s1["n_near_tie"] == st["n_near_tie"][i]
        np.testing.assert_array_equal(np.asarray(l1), loc[i])
    clean = (st["n_near_tie"] == 0) & (g["n_near_tie"] == 0)
    parity_report("cfg3_intersection_batch", transforms=64, clean=int(clean.sum()), flagged=int((~clean).sum()),
                  verdict_mismatch=int((found != g["found"]).sum()), traversal_differs=int((st["n_nodes"] != g["n_nodes"]).sum()))
    np.testing.assert_array_equal(found[clean], g["found"][clean])
    np.testing.assert_array_equal(st["n_nodes"][clean], g["n_nodes"][clean])
    np.testing.assert_array_equal(st["n_rounds"][clean], g["n_rounds"][clean])
    np.testing.assert_allclose(loc[clean], g["loc"][clean], rtol=0, atol=1e-6)
    assert int((found != g["found"])[~clean].sum()) <= 1
    # transforms on shape A as well, and a mode without the persistent kernel (one call per query behind the same function)
    import mlp
    pA2 = mlp.prepend_op(sample_params("hammer"), mlp.spatial_transformation())
    fA2 = make(pA2, "affine_truncate", affine_n_truncate=64, affine_truncate_policy="absolute")
    eye = np.tile(np.eye(3, dtype=np.float32), (4, 1, 1))
    f2, l2 = kd_tree.find_any_intersection_batch((fA2, fB), (pA2, pB), LO, HI, 1e-3, R_A=eye, t_A=np.zeros((4, 3), np.float32),
                                                 R_B=g["R"][:4], t_B=g["t"][:4])
    np.testing.assert_array_equal(f2, found[:4])
    np.testing.assert_array_equal(l2, loc[:4])
    ffA, ffB = make(pA, "affine_fixed"), make(pB, "affine_fixed")
    f3, _ = kd_tree.find_any_intersection_batch((ffA, ffB), (pA, pB), LO, HI, 1e-3, R_B=g["R"][:3], t_B=g["t"][:3])
    assert f3.shape == (3,)


def test_cfg3_more_than_cache_size_transforms():
    """> 64 distinct params sets through one context: the MLP handle cache must never close a handle the call still uses
    (advisor finding on _niq.Context.mlp)."""
    import kd_tree
    import mlp
    pA = sample_params("fox")
    pB = mlp.prepend_op(sample_params("fox"), mlp.spatial_transformation())
    fA, fB = make(pA, "affine_fixed"), make(pB, "affine_fixed")
    n_found = 0
    for i in range(80):
        pB["0000.spatial_transformation.t"] = np.array((0.02 * i, 0., 0.), np.float32)
        n_found += bool(kd_tree.find_any_intersection((fA, fB), (pA, pB), LO, HI, 1e-2)[0])
    assert 0 < n_found <= 80


# ---------------------------------------------------------------------------------------------------
# config 4: birdcage_occ closest_point, Q = 256, eps 1e-3, at the reference's default window and at window >= stack
# (reference src/kd_tree.py:765-802; SURVEY.md F6: results depend on the window)
# ---------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("case,min_agree", [("cfg4_closest_birdcage_B2048", 0.85), ("cfg4_closest_birdcage_B2p21", 0.95)])
def test_cfg4_closest_point_birdcage(case, min_agree):
    import kd_tree
    g = golden(case)
    p = sample_params("birdcage_occ")
    st = {}
    d, loc = kd_tree.closest_point(make(p, "affine_fixed"), p, LO, HI, g["query_points"], eps=float(g["eps"]),
                                   batch_process_size=int(g["B"]), stats=st)
    od, oloc = g["dist"], g["loc"]
    assert np.isfinite(od).all() and np.isfinite(d).all()
    rel = np.abs(d - od) / od
    ok = rel <= RTOL
    loc_ok = np.abs(loc - oloc).max(axis=1) <= 1e-6
    parity_report(case, queries=d.shape[0], window=int(g["B"]), visits_gpu=st["n_visits"], visits_oracle=int(g["n_visits"]),
                  rounds_gpu=st["n_rounds"], rounds_oracle=int(g["n_rounds"]), max_stack_gpu=st["max_stack"],
                  max_stack_oracle=int(g["max_stack"]), near_tie_boxes_gpu=st["n_near_tie"], near_tie_boxes_oracle=int(g["n_near_tie"]),
                  queries_with_near_tie_oracle=int(g["tie_query"].sum()), dist_within_1e5=int(ok.sum()), loc_equal=int((ok & loc_ok).sum()),
                  max_rel_diff=float(rel.max()))
    # 15 M boxes are classified; ~1e-4 of them lie inside the band, and with the reference's shared LIFO window ONE flipped
    # label re-orders every later window, so the traversals need not be identical: the flagged boxes must stay
    # rare, the visit counts close, and the distances equal for the stated majority and near-equal for all
    assert st["n_near_tie"] <= 2e-4 * st["n_visits"]
    assert abs(st["n_visits"] - int(g["n_visits"])) <= 0.02 * int(g["n_visits"])
    assert ok.mean() >= min_agree, f"only {ok.mean():.3f} of the distances within 1e-5"
    assert rel.max() < 0.05
    if st["n_near_tie"] == 0 and int(g["n_near_tie"]) == 0:
        assert ok.all() and st["n_visits"] == int(g["n_visits"]) and st["n_rounds"] == int(g["n_rounds"])


# ---------------------------------------------------------------------------------------------------
# config 5: random-init 3 -> 256 x 8 -> 1 ReLU MLP (NumPy seed 0): depth-14 tree, cast_rays at the full 512 steps
# ---------------------------------------------------------------------------------------------------

def _params5():
    import mlp
    return mlp.initialize_params(mlp.build_spec(mlp.quick_mlp_spec(LAYERS5, "relu")), 0)


def test_cfg5_tree_depth14_vs_oracle():
    import kd_tree
    g = golden("cfg5_tree_d14")
    p = _params5()
    st = {}
    out = kd_tree.construct_uniform_unknown_levelset_tree(make(p, "affine_fixed"), p, LO, HI, split_depth=14, stats=st)
    v = out["unknown_node_valid"]
    parity_report("cfg5_tree_depth14", boxes=st["n_evals"], leaves=int(v.sum()), near_tie_gpu=st["n_near_tie"], near_tie_oracle=int(g["n_near_tie"]))
    assert st["n_near_tie"] == 0 and int(g["n_near_tie"]) == 0       # the full tree of this net has no box near the level set
    assert st["n_evals"] == int(g["n_evals"]) == 2 ** 15 - 1
    assert st["level_sizes"] == g["level_sizes"].tolist()
    assert v.shape[0] == int(g["padded_size"]) and int(v.sum()) == int(g["n_valid"])
    np.testing.assert_array_equal(out["unknown_node_lower"][v], g["lower"])     # order too
    np.testing.assert_array_equal(out["unknown_node_upper"][v], g["upper"])


def test_cfg5_cast_rays_full_512_steps_vs_oracle():
    import queries
    g = golden("cfg5_rays_256x512")
    p = _params5()
    opts = queries.get_default_cast_opts()
    t, hit, cnt, n_evals, tie = queries.cast_rays((make(p, "affine_fixed"),), (p,), g["roots"], g["dirs"], opts, return_near_tie=True)
    flagged = tie | g["near_tie"]
    okm = ~flagged
    parity_report("cfg5_cast_rays_512_steps", rays=t.shape[0], ray_steps=int(cnt.sum()), flagged_gpu=int(tie.sum()),
                  flagged_oracle=int(g["near_tie"].sum()), flagged_differing=int(((cnt != g["count"]) | (hit != g["hit"]))[flagged].sum()))
    assert flagged.mean() <= 0.05
    np.testing.assert_array_equal(hit[okm], g["hit"][okm])
    np.testing.assert_array_equal(cnt[okm], g["count"][okm])
    np.testing.assert_allclose(t[okm], g["t"][okm], rtol=RTOL, atol=0)
    assert int(cnt.min()) == 512 and n_evals == int(g["n_evals"])     # every ray of this net runs to the step limit


# ---------------------------------------------------------------------------------------------------
# the tolerance itself: strict 1e-5 band with max(|lower|,|upper|) = |base| + rad as the yardstick, beside the wider
# yardstick the backend flags with (net.tol_scale adds the magnitude of the last dot product); elu nets at 1e-5 and 2e-4
# ---------------------------------------------------------------------------------------------------

def _random_boxes(seed, n, smin=-9, smax=0):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    h = (2.0 ** rng.uniform(smin, smax, (n, 1)) * rng.uniform(0.5, 1.0, (n, 3))).astype(np.float32)
    return c - h, c + h


@pytest.mark.parametrize("name", ["fox", "hammer", "birdcage_occ", "bunny"])
def test_strict_band_accounting(name):
    p = sample_params(name)
    f = make(p, "affine_fixed")
    lo, hi = _random_boxes(11, 20000)
    lab, blo, bup, _ = f.bound_box(p, lo, hi)
    olab, olo, oup, osc = net.classify_box(p, net.AffineContext("affine_fixed"), lo, hi, return_scale=True)
    strict_scale = np.maximum(np.abs(olo), np.abs(oup)).astype(np.float64)          # |base| + rad
    wide_scale = net.tol_scale(olo, oup, osc)
    err = np.maximum(np.abs(blo.astype(np.float64) - olo), np.abs(bup.astype(np.float64) - oup))
    rel_wide = net.tie_rel(p)
    counts = dict(boxes=lo.shape[0],
                  values_beyond_1e5_strict=int((err > 1e-5 * strict_scale).sum()),
                  values_beyond_1e5_wide=int((err > 1e-5 * wide_scale).sum()),
                  values_beyond_band_wide=int((err > rel_wide * wide_scale).sum()),
                  max_err_over_strict_scale=float((err / np.maximum(strict_scale, 1e-30)).max()))
    mism = lab != olab
    tie_strict = net.bound_near_tie(olo, oup, 0.0, None, rel=1e-5)                  # yardstick max(|lower|,|upper|) only
    tie_wide_1e5 = net.bound_near_tie(olo, oup, 0.0, osc, rel=1e-5)
    tie_wide = net.bound_near_tie(olo, oup, 0.0, osc, rel=rel_wide)
    counts.update(label_mismatch=int(mism.sum()), near_tie_strict_1e5=int(tie_strict.sum()), near_tie_wide_1e5=int(tie_wide_1e5.sum()),
                  near_tie_band=int(tie_wide.sum()), band_rel=rel_wide,
                  mismatch_outside_strict_1e5=int((mism & ~tie_strict).sum()), mismatch_outside_wide_1e5=int((mism & ~tie_wide_1e5).sum()),
                  mismatch_outside_band=int((mism & ~tie_wide).sum()))
    parity_report(f"strict_band_accounting[{name}]", **counts)
    assert counts["mismatch_outside_band"] == 0 and counts["values_beyond_band_wide"] == 0
    if name != "bunny":
        # relu-only nets meet north_star's literal tolerance: every bound within 1e-5 of |base| + rad, no label differs outside
        # the strict band (measured: max error 1.2e-6 of the strict yardstick, 0 mismatches)
        assert counts["mismatch_outside_strict_1e5"] == 0
        assert counts["values_beyond_1e5_strict"] == 0
    else:
        # elu (bunny): the reference's float32 rule is itself conditioned to ~5e-5 (tests/test_oracle_golden.py::
        # test_elu_rule_conditioning), hence the 2e-4 band on the wider yardstick; the counts at 1e-5 are reported beside it
        assert counts["max_err_over_strict_scale"] < 5e-3
