"""TanH MLPs (BASELINE.json north_star names "ReLU/TanH" for config 5).  PARITY UNPINNED: the reference registers no tanh op
or rule (README.md:24 says "ReLU or TanH", src/affine_layers.py has relu :34, elu :59, sin :100 only), so there is
nothing of the reference to compare with.  What is checked instead: (CPU) the self-written oracle rule is SOUND -- sampled f(x)
inside a box lies within the computed bounds, in every mode; (GPU) the CUDA rule reproduces the oracle's within the band, and is
sound on its own."""
import numpy as np
import pytest

from niq_oracle import net, rays

LO = np.full(3, -1, np.float32)
HI = np.full(3, 1, np.float32)
MODES = ("interval", "affine_fixed", "affine_truncate", "affine_all", "affine_append", "slope_interval")


def tanh_net(width=32, depth=3, seed=1, gain=3.0):
    p = net.random_mlp([3] + [width] * depth + [1], "tanh", seed=seed)
    for k in p:
        if k.endswith("dense.A"):
            p[k] = (p[k] * np.float32(gain)).astype(np.float32)      # glorot-normal tanh nets are almost linear: steepen them
    return p


def boxes(seed, n, smin=-8, smax=0):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    h = (2.0 ** rng.uniform(smin, smax, (n, 1)) * rng.uniform(0.5, 1.0, (n, 3))).astype(np.float32)
    return c - h, c + h


def octx(mode):
    return net.AffineContext(mode, truncate_count=8, n_append=4)


def sound(params, lo, hi, lower, upper, seed=0, n_samp=128):
    rng = np.random.default_rng(seed)
    bad = 0
    for i in range(lo.shape[0]):
        x = (lo[i] + (hi[i] - lo[i]) * rng.uniform(0, 1, (n_samp, 3))).astype(np.float32)
        f = net.eval_points(params, x).astype(np.float64)
        tol = 1e-5 * max(abs(float(lower[i])), abs(float(upper[i])), 1.0)
        bad += int((f < lower[i] - tol).sum() + (f > upper[i] + tol).sum())
    return bad


@pytest.mark.parametrize("mode", MODES)
def test_oracle_tanh_rule_is_sound(mode):
    p = tanh_net()
    lo, hi = boxes(3, 300)
    lab, lower, upper = net.classify_box(p, octx(mode), lo, hi, return_bounds=True)
    assert sound(p, lo, hi, lower, upper) == 0
    assert (lab != net.SIGN_UNKNOWN).any() and (lab == net.SIGN_UNKNOWN).any()


def test_oracle_tanh_coefficients_bracket_the_function():
    """(alpha, beta, delta) of the rule on random intervals: |tanh(x) - (alpha x + beta)| <= delta (+ rounding) on a dense
    sample of [l, u], including the narrow-interval branch and intervals far out in the saturated tails."""
    rng = np.random.default_rng(5)
    mid = rng.uniform(-6, 6, 4000).astype(np.float32)
    rad = (10.0 ** rng.uniform(-6, 1, 4000)).astype(np.float32)
    base, aff, err = mid[None, :], np.zeros((1, 0, 4000), np.float32), rad[None, :]
    alpha, beta, delta = net._tanh_coeffs(base, aff, err)
    l, u = (mid - rad).astype(np.float64), (mid + rad).astype(np.float64)
    for s in np.linspace(0, 1, 65):
        x = l + s * (u - l)
        resid = np.abs(np.tanh(x) - (alpha[0].astype(np.float64) * x + beta[0]))
        assert np.all(resid <= delta[0] + 2e-6 * (1 + np.abs(x) * alpha[0])), float((resid - delta[0]).max())
    assert np.all((alpha >= 0) & (alpha <= 1))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("width,depth", [(32, 3), (64, 8)])
def test_gpu_tanh_classify_vs_oracle_and_sound(mode, width, depth):
    import implicit_mlp_utils
    from conftest import parity_report
    p = tanh_net(width, depth, seed=2, gain=2.0 if depth > 4 else 3.0)
    kw = dict(affine_n_truncate=8, affine_truncate_policy="absolute") if mode == "affine_truncate" else {}
    if mode == "affine_append":
        kw = dict(affine_n_append=4)
    f = implicit_mlp_utils.generate_implicit_from_params(p, mode, **kw)
    lo, hi = boxes(7, 600)
    lab, lower, upper, tie = f.bound_box(p, lo, hi)
    olab, olo, oup, osc = net.classify_box(p, octx(mode), lo, hi, return_scale=True)
    rel = net.mode_rel(p, octx(mode))
    scale = net.tol_scale(olo, oup, osc)
    err = np.maximum(np.abs(lower.astype(np.float64) - olo), np.abs(upper.astype(np.float64) - oup)) / scale
    otie = net.bound_near_tie(olo, oup, 0.0, osc, rel=rel)
    mism = lab != olab
    parity_report(f"tanh_classify[{mode}-{width}x{depth}] (unpinned)", boxes=lo.shape[0], max_err_over_scale=float(err.max()), band=rel,
                  near_tie=int(otie.sum()), label_mismatch=int(mism.sum()), mismatch_outside_band=int((mism & ~otie).sum()))
    assert err.max() <= rel and not (mism & ~otie).any()
    assert sound(p, lo[:150], hi[:150], lower[:150], upper[:150]) == 0
    # point values: within 1e-5 of the magnitude the last dot product was summed from
    import mlp
    x = np.random.default_rng(1).uniform(-1, 1, (2000, 3)).astype(np.float32)
    fv, fsc = mlp.eval_points(p, x, return_scale=True)
    assert np.all(np.abs(fv - net.eval_points(p, x)) <= 1e-5 * fsc)


@pytest.mark.gpu
def test_gpu_tanh_queries_vs_oracle():
    """cast_rays, the level-set tree and marching cubes on a tanh net against the oracle (affine_fixed)."""
    import implicit_mlp_utils
    import kd_tree
    import queries
    import render
    from conftest import parity_report
    from niq_oracle import tree as otree
    p = tanh_net(32, 3, seed=5, gain=2.0)
    xs = np.random.default_rng(0).uniform(-1, 1, (20000, 3)).astype(np.float32)
    last = sorted(k for k in p if k.endswith("dense.b"))[-1]
    p[last] = (p[last] - np.float32(np.median(net.eval_points(p, xs)))).astype(np.float32)     # the level set crosses the domain
    f = implicit_mlp_utils.generate_implicit_from_params(p, "affine_fixed")
    eye = np.array((2., 1., 2.), np.float32)
    look, up, _ = render.look_at(eye)
    roots, dirs = render.generate_camera_rays(eye, look, up, res=32, fov_deg=30.)
    opts = queries.get_default_cast_opts()
    t, hit, cnt, n_ev, tie = queries.cast_rays((f,), (p,), roots, dirs, opts, return_near_tie=True)
    ot, ohit, ocnt, on_ev, otie = rays.cast_rays((octx("affine_fixed"),), (p,), roots, dirs, opts, return_near_tie=True)
    ok = ~(tie | otie)
    both = (hit > 0) & (ohit > 0)
    parity_report("tanh_cast_rays (unpinned)", rays=int(t.shape[0]), flagged=int((~ok).sum()), hits=int((ohit > 0).sum()),
                  hit_flag_mismatch=int((hit != ohit).sum()), count_mismatch=int((cnt != ocnt).sum()),
                  max_t_diff_on_hits=float(np.abs(t - ot)[both].max()) if both.any() else 0.0)
    # rays whose every decision was outside the band: exact; all rays (a ray that comes to rest ON the surface is always
    # inside the band of the 1e-4 tanh yardstick): hit flags and depths agree to the hit tolerance of the march
    np.testing.assert_array_equal(hit[ok], ohit[ok])
    np.testing.assert_array_equal(cnt[ok], ocnt[ok])
    np.testing.assert_allclose(t[ok], ot[ok], rtol=1e-5, atol=0)
    assert (hit != ohit).mean() <= 0.02 and (ohit > 0).any() and (ohit == 0).any()
    assert np.all(np.abs(t - ot)[both] <= 2 * opts["hit_eps"])
    st, ost = {}, {}
    out = kd_tree.construct_uniform_unknown_levelset_tree(f, p, LO, HI, split_depth=12, stats=st)
    ref = otree.construct_uniform_unknown_levelset_tree(octx("affine_fixed"), p, LO, HI, split_depth=12, stats=ost)
    parity_report("tanh_tree_d12 (unpinned)", boxes=ost["n_evals"], near_tie_gpu=st["n_near_tie"], near_tie_oracle=ost["n_near_tie"])
    if st["n_near_tie"] == 0 and ost["n_near_tie"] == 0:
        v, rv = out["unknown_node_valid"], ref["unknown_node_valid"]
        np.testing.assert_array_equal(out["unknown_node_lower"][v], ref["unknown_node_lower"][rv])
    else:
        assert abs(int(out["unknown_node_valid"].sum()) - int(ref["unknown_node_valid"].sum())) <= 16 * (st["n_near_tie"] + ost["n_near_tie"])
    assert st["n_near_tie"] + ost["n_near_tie"] <= 1e-3 * ost["n_evals"] + 2
