#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference sources on the NumPy `jax` stand-in.

Build-container only (reads /root/reference; needs oracle/jaxshim).  Usage:

    python oracle/tools/gen_golden.py [case ...]        # default: all cases, in parallel

Every case stores its inputs and the reference's outputs; tests/test_oracle_golden.py replays the inputs
through oracle/niq_oracle and compares.  Cases are small because the stand-in executes `vmap` as a
Python loop.  The weights come from /root/reference/sample_inputs/*.npz, which are NOT copied: tests
that need them on the GPU box use tests/golden/mlps.npz (written here: the four sample MLPs'
arrays under '<name>/<key>' keys, 96k floats in total) -- fixtures derived from the reference inputs.
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"
SAMPLES = ("fox", "bunny", "hammer", "birdcage_occ")


def _ref_modules():
    sys.path[:0] = [os.path.join(ROOT, "oracle", "jaxshim"), os.path.join(ROOT, "oracle", "jaxshim", "stubs"),
                    os.path.join(REF, "src")]
    import jax  # noqa: F401  (the stand-in)
    import jax.numpy as jnp
    import affine
    import implicit_mlp_utils
    import kd_tree
    import mlp
    import queries
    import render
    return dict(jnp=jnp, affine=affine, imu=implicit_mlp_utils, kd_tree=kd_tree, mlp=mlp, queries=queries,
                render=render)


def _load(m, name, mode, **kw):
    return m["imu"].generate_implicit_from_file(f"{REF}/sample_inputs/{name}.npz", mode, **kw)


def _mode_kwargs(mode, n_trunc):
    if mode == "affine_truncate":
        return dict(affine_n_truncate=n_trunc, affine_truncate_policy="absolute")
    if mode == "affine_append":
        return dict(affine_n_append=n_trunc)
    if mode == "sdf":
        return dict(sdf_lipschitz=n_trunc)
    return {}


def _boxes(seed, n_per_scale=3, scales=range(0, 11)):
    rng = np.random.default_rng(seed)
    lo, hi = [], []
    for s in scales:
        half = np.float32(2.0 ** (-s))
        c = rng.uniform(-1, 1, (n_per_scale, 3)).astype(np.float32) * np.float32(1.0 - 0.5 * half)
        h = (half * rng.uniform(0.5, 1.0, (n_per_scale, 3))).astype(np.float32)
        lo.append(c - h)
        hi.append(c + h)
    return np.concatenate(lo).astype(np.float32), np.concatenate(hi).astype(np.float32)


# ------------------------------------------------------------------------------------------------

def case_mlps():
    out = {}
    for name in SAMPLES:
        with np.load(f"{REF}/sample_inputs/{name}.npz") as d:
            for k in d.files:
                out[f"{name}/{k}"] = d[k]
    return out


def case_classify(name, mode, n_trunc=8):
    """labels + bounds of axis-aligned boxes (v=3), ray segments (v=1), an offset, and a rigid transform."""
    m = _ref_modules()
    jnp, affine = m["jnp"], m["affine"]
    func, params = _load(m, name, mode, **_mode_kwargs(mode, n_trunc))
    import zlib
    lo, hi = _boxes(seed=zlib.crc32(f"{name}-{mode}".encode()) % 1000)
    out = dict(box_lower=lo, box_upper=hi, n_trunc=n_trunc)

    def bounds(params, center, vecs):
        import dataclasses
        ctx = dataclasses.replace(func.ctx, affine_domain_terms=vecs.shape[0])
        inp = affine.coordinates_in_general_box(ctx, center, vecs)
        res = func.affine_func(params, inp, {"ctx": ctx})
        return affine.may_contain_bounds(ctx, res)

    lab, lb, ub = [], [], []
    lab_off = []
    for i in range(lo.shape[0]):
        l, u = jnp.array(lo[i]), jnp.array(hi[i])
        lab.append(int(func.classify_box(params, l, u)))
        lab_off.append(int(func.classify_box(params, l, u, offset=0.05)))
        c = 0.5 * (l + u)
        b = bounds(params, c, jnp.diag(u - c))
        lb.append(float(b[0]))
        ub.append(float(b[1]))
    out.update(label=np.array(lab, np.int32), label_offset005=np.array(lab_off, np.int32),
               lower=np.array(lb, np.float32), upper=np.array(ub, np.float32))

    # v = 1 general boxes (ray segments)
    rng = np.random.default_rng(7)
    cen = rng.uniform(-1, 1, (12, 3)).astype(np.float32)
    vec = (rng.standard_normal((12, 1, 3)) * (2.0 ** -rng.integers(0, 8, (12, 1, 1)))).astype(np.float32)
    glab, glb, gub = [], [], []
    for i in range(cen.shape[0]):
        glab.append(int(func.classify_general_box(params, jnp.array(cen[i]), jnp.array(vec[i]))))
        b = bounds(params, jnp.array(cen[i]), jnp.array(vec[i]))
        glb.append(float(b[0]))
        gub.append(float(b[1]))
    out.update(seg_center=cen, seg_vecs=vec, seg_label=np.array(glab, np.int32),
               seg_lower=np.array(glb, np.float32), seg_upper=np.array(gub, np.float32))

    # rigid transform prepended (mlp.prepend_op + spatial_transformation), first 9 boxes
    th = 0.7
    R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]], np.float32)
    t = np.array([0.2, -0.1, 0.05], np.float32)
    p2 = m["mlp"].prepend_op(params, m["mlp"].spatial_transformation())
    p2["0000.spatial_transformation.R"] = jnp.array(R)
    p2["0000.spatial_transformation.t"] = jnp.array(t)
    tl, tlb, tub = [], [], []
    for i in range(9, 18):
        l, u = jnp.array(lo[i]), jnp.array(hi[i])
        tl.append(int(func.classify_box(p2, l, u)))
        c = 0.5 * (l + u)
        b = bounds(p2, c, jnp.diag(u - c))
        tlb.append(float(b[0]))
        tub.append(float(b[1]))
    pts = rng.uniform(-1, 1, (16, 3)).astype(np.float32)
    out.update(xf_R=R, xf_t=t, xf_label=np.array(tl, np.int32), xf_lower=np.array(tlb, np.float32),
               xf_upper=np.array(tub, np.float32), xf_points=pts,
               xf_values=np.array([float(func(p2, jnp.array(x))) for x in pts], np.float32))
    return out


def case_classify_sdf(name, lipschitz):
    """sdf.WeakSDFImplicitFunction (src/sdf.py): labels of axis-aligned boxes (two offsets) and v=1 / v=2 general boxes."""
    m = _ref_modules()
    jnp = m["jnp"]
    func, params = _load(m, name, "sdf", sdf_lipschitz=lipschitz)
    import zlib
    lo, hi = _boxes(seed=zlib.crc32(f"{name}-sdf".encode()) % 1000)
    lab = [int(func.classify_box(params, jnp.array(lo[i]), jnp.array(hi[i]))) for i in range(lo.shape[0])]
    lab_off = [int(func.classify_box(params, jnp.array(lo[i]), jnp.array(hi[i]), offset=0.05)) for i in range(lo.shape[0])]
    rng = np.random.default_rng(7)
    cen = rng.uniform(-1, 1, (16, 3)).astype(np.float32)
    vec = (rng.standard_normal((16, 2, 3)) * (2.0 ** -rng.integers(2, 9, (16, 1, 1)))).astype(np.float32)
    glab = [int(func.classify_general_box(params, jnp.array(cen[i]), jnp.array(vec[i]))) for i in range(16)]
    glab1 = [int(func.classify_general_box(params, jnp.array(cen[i]), jnp.array(vec[i, :1]))) for i in range(16)]
    return dict(box_lower=lo, box_upper=hi, lipschitz=np.float32(lipschitz), label=np.array(lab, np.int32),
                label_offset005=np.array(lab_off, np.int32), gen_center=cen, gen_vecs=vec, gen_label=np.array(glab, np.int32),
                gen_label_v1=np.array(glab1, np.int32))


def case_classify_slope(name):
    """slope_interval.SlopeIntervalImplicitFunction: labels (two offsets) + may-contain bounds of axis-aligned boxes,
    v=1 / v=2 general boxes, and a prepended rigid transform."""
    m = _ref_modules()
    jnp = m["jnp"]
    import slope_interval as si
    func, params = _load(m, name, "slope_interval")
    import zlib
    lo, hi = _boxes(seed=zlib.crc32(f"{name}-slope".encode()) % 1000)

    def bounds(p, center, vecs):
        out = func.slope_interval_func(p, si.coordinates_in_general_box(center, vecs))
        sl, su = si.slope_bounds(out)
        return si.primal_may_contain_bounds(out, sl, su)

    lab, lab_off, lb, ub = [], [], [], []
    for i in range(lo.shape[0]):
        l, u = jnp.array(lo[i]), jnp.array(hi[i])
        lab.append(int(func.classify_box(params, l, u)))
        lab_off.append(int(func.classify_box(params, l, u, offset=0.05)))
        c = 0.5 * (l + u)
        b = bounds(params, c, jnp.diag(u - c))
        lb.append(float(b[0])); ub.append(float(b[1]))
    rng = np.random.default_rng(7)
    cen = rng.uniform(-1, 1, (12, 3)).astype(np.float32)
    vec = (rng.standard_normal((12, 2, 3)) * (2.0 ** -rng.integers(0, 8, (12, 1, 1)))).astype(np.float32)
    glab, glb, gub, g1lab, g1lb, g1ub = [], [], [], [], [], []
    for i in range(12):
        glab.append(int(func.classify_general_box(params, jnp.array(cen[i]), jnp.array(vec[i]))))
        b = bounds(params, jnp.array(cen[i]), jnp.array(vec[i])); glb.append(float(b[0])); gub.append(float(b[1]))
        g1lab.append(int(func.classify_general_box(params, jnp.array(cen[i]), jnp.array(vec[i, :1]))))
        b = bounds(params, jnp.array(cen[i]), jnp.array(vec[i, :1])); g1lb.append(float(b[0])); g1ub.append(float(b[1]))
    th = 0.7
    R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]], np.float32)
    t = np.array([0.2, -0.1, 0.05], np.float32)
    p2 = m["mlp"].prepend_op(params, m["mlp"].spatial_transformation())
    p2["0000.spatial_transformation.R"] = jnp.array(R)
    p2["0000.spatial_transformation.t"] = jnp.array(t)
    tl, tlb, tub = [], [], []
    for i in range(9, 18):
        l, u = jnp.array(lo[i]), jnp.array(hi[i])
        tl.append(int(func.classify_box(p2, l, u)))
        c = 0.5 * (l + u)
        b = bounds(p2, c, jnp.diag(u - c)); tlb.append(float(b[0])); tub.append(float(b[1]))
    f32 = lambda a: np.array(a, np.float32)
    i32 = lambda a: np.array(a, np.int32)
    return dict(box_lower=lo, box_upper=hi, label=i32(lab), label_offset005=i32(lab_off), lower=f32(lb), upper=f32(ub),
                gen_center=cen, gen_vecs=vec, gen_label=i32(glab), gen_lower=f32(glb), gen_upper=f32(gub),
                gen1_label=i32(g1lab), gen1_lower=f32(g1lb), gen1_upper=f32(g1ub),
                xf_R=R, xf_t=t, xf_label=i32(tl), xf_lower=f32(tlb), xf_upper=f32(tub))


def _pe_params(m):
    """A positional-encoding MLP as src/main_fit_implicit.py:113-115 builds it: pow2_frequency_encode(4, start_pow=-1,
    with_shift) -> sin -> 24 -> 32 -> 32 -> 32 -> 1 relu net, weights from NumPy seed 5 (glorot-normal scale)."""
    mlp = m["mlp"]
    jnp = m["jnp"]
    rng = np.random.default_rng(5)
    sizes = [24, 32, 32, 32, 1]
    spec = [mlp.pow2_frequency_encode(4, start_pow=-1, with_shift=True), mlp.sin()]
    for i in range(len(sizes) - 1):
        A = (rng.standard_normal((sizes[i], sizes[i + 1])) * np.sqrt(2.0 / (sizes[i] + sizes[i + 1]))).astype(np.float32)
        b = (rng.standard_normal(sizes[i + 1]) * 0.05).astype(np.float32)
        spec.append(mlp.dense(sizes[i], sizes[i + 1], A=jnp.array(A), b=jnp.array(b)))
        if i + 2 != len(sizes):
            spec.append(mlp.relu())
    spec.append(mlp.squeeze_last())
    return mlp.build_spec(spec)


def case_pe(mode, n_trunc=8):
    """sin + pow2_frequency_encode (src/mlp.py:296-322, src/affine_layers.py:100-161, src/slope_interval_layers.py:85-126)
    on a positional-encoding MLP: labels + bounds in `mode`, point values; the params are stored in the fixture."""
    m = _ref_modules()
    jnp = m["jnp"]
    params = _pe_params(m)
    path = f"/tmp/niq_pe_mlp_{os.getpid()}.npz"
    m["mlp"].save(path, params)
    func, params = m["imu"].generate_implicit_from_file(path, mode, **_mode_kwargs(mode, n_trunc))
    os.remove(path)
    out = {"params/" + k: np.array(v) for k, v in params.items()}
    lo, hi = _boxes(seed=21, n_per_scale=3, scales=range(1, 11))
    out.update(box_lower=lo, box_upper=hi, n_trunc=n_trunc)
    if mode == "slope_interval":
        import slope_interval as si

        def bounds(center, vecs):
            o = func.slope_interval_func(params, si.coordinates_in_general_box(center, vecs))
            sl, su = si.slope_bounds(o)
            return si.primal_may_contain_bounds(o, sl, su)
    else:
        affine = m["affine"]

        def bounds(center, vecs):
            import dataclasses
            ctx = dataclasses.replace(func.ctx, affine_domain_terms=vecs.shape[0])
            res = func.affine_func(params, affine.coordinates_in_general_box(ctx, center, vecs), {"ctx": ctx})
            return affine.may_contain_bounds(ctx, res)
    lab, lb, ub = [], [], []
    for i in range(lo.shape[0]):
        l, u = jnp.array(lo[i]), jnp.array(hi[i])
        lab.append(int(func.classify_box(params, l, u)))
        c = 0.5 * (l + u)
        b = bounds(c, jnp.diag(u - c))
        lb.append(float(b[0])); ub.append(float(b[1]))
    pts = np.random.default_rng(3).uniform(-1, 1, (32, 3)).astype(np.float32)
    out.update(label=np.array(lab, np.int32), lower=np.array(lb, np.float32), upper=np.array(ub, np.float32), points=pts,
               values=np.array([float(func(params, jnp.array(x))) for x in pts], np.float32))
    return out


def case_render(name, mode, res, frustum=False, n_side=16):
    """render.render_image (src/render.py:94-150): shading='normal' -- image, depth, counts, hit ids; ray or frustum branch."""
    m = _ref_modules()
    jnp, render = m["jnp"], m["render"]
    func, params = _load(m, name, mode)
    eye = jnp.array((2., 1., 2.))
    look, up, left = render.look_at(eye)
    opts = m["queries"].get_default_cast_opts()
    opts["n_side_init"] = n_side
    with np.errstate(all="ignore"):
        img, depth, counts, hit_ids, n_eval, _ = render.render_image(func, params, eye, look, up, left, res, 30., frustum, opts, shading="normal")
    return dict(img=np.array(img, np.float32), depth=np.array(depth, np.float32), counts=np.array(counts, np.int32),
                hit_ids=np.array(hit_ids, np.int32), n_eval=np.int64(n_eval), res=res, n_side=n_side)


def case_frustum(names, mode, res, n_side, n_substeps=1, n_trunc=8, res_y=None):
    """queries.cast_rays_frustum (src/queries.py:178-587): out_t / out_hit_id / out_count as returned, (res_x, res_y)."""
    m = _ref_modules()
    jnp, queries, render = m["jnp"], m["queries"], m["render"]
    funcs, params = [], []
    for nm in names:
        f, p = _load(m, nm, mode, **_mode_kwargs(mode, n_trunc))
        funcs.append(f)
        params.append(p)
    eye = jnp.array((2.0, 1.0, 2.0))
    look, up, left = render.look_at(eye)
    opts = queries.get_default_cast_opts()
    opts["n_side_init"] = n_side
    opts["n_substeps"] = n_substeps
    res_y = res if res_y is None else res_y
    cam = (eye, look, up, left, 30., 30. if res_y == res else 22., res, res_y)          # a second fov for the non-square case
    with np.errstate(all="ignore"):
        t, hit, cnt, n_evals = queries.cast_rays_frustum(tuple(funcs), tuple(params), cam, opts)
    return dict(eye=np.array(eye), look=np.array(look), up=np.array(up), left=np.array(left), res=res, n_side=n_side,
                n_substeps=n_substeps, n_trunc=n_trunc, res_y=res_y, fov_y=np.float32(30. if res_y == res else 22.), out_t=np.array(t, np.float32), out_hit_id=np.array(hit, np.int32),
                out_count=np.array(cnt, np.int32), n_evals=int(n_evals))


def case_points(name):
    m = _ref_modules()
    jnp = m["jnp"]
    func, params = _load(m, name, "affine_fixed")
    rng = np.random.default_rng(3)
    x = rng.uniform(-1, 1, (64, 3)).astype(np.float32)
    f = np.array([float(func(params, jnp.array(p))) for p in x], np.float32)
    return dict(points=x, values=f)


def case_cast_rays(names, mode, res, n_substeps=1, n_trunc=8):
    m = _ref_modules()
    jnp, queries, render = m["jnp"], m["queries"], m["render"]
    funcs, params = [], []
    for nm in names:
        f, p = _load(m, nm, mode, **_mode_kwargs(mode, n_trunc))
        funcs.append(f)
        params.append(p)
    eye = jnp.array((2.0, 1.0, 2.0))
    look, up, left = render.look_at(eye)
    roots, dirs = render.generate_camera_rays(eye, look, up, res=res, fov_deg=30.0)
    opts = queries.get_default_cast_opts()
    opts["n_substeps"] = n_substeps
    t, hit, cnt, n_evals = queries.cast_rays(tuple(funcs), tuple(params), roots, dirs, opts)
    return dict(roots=np.array(roots), dirs=np.array(dirs), eye=np.array(eye), look=np.array(look),
                up=np.array(up), res=res, n_substeps=n_substeps, n_trunc=n_trunc,
                out_t=np.array(t, np.float32), out_hit_id=np.array(hit, np.int32),
                out_count=np.array(cnt, np.int32), n_evals=int(n_evals))


def case_tree(name, mode, n_trunc=8, **kw):
    m = _ref_modules()
    jnp, kd = m["jnp"], m["kd_tree"]
    func, params = _load(m, name, mode, **_mode_kwargs(mode, n_trunc))
    lower = jnp.array((-1.0, -1.0, -1.0))
    upper = jnp.array((1.0, 1.0, 1.0))
    d = kd.construct_uniform_unknown_levelset_tree(func, params, lower, upper, **kw)
    out = {k: np.array(v) for k, v in d.items()}
    out["n_trunc"] = n_trunc
    for k, v in kw.items():
        out["kw_" + k] = np.array(-1 if v is None else v)
    return out


def case_mc(name, mode, depth, n_sub):
    m = _ref_modules()
    jnp, kd = m["jnp"], m["kd_tree"]
    func, params = _load(m, name, mode)
    lower = jnp.array((-1.0, -1.0, -1.0))
    upper = jnp.array((1.0, 1.0, 1.0))
    tri = kd.hierarchical_marching_cubes(func, params, lower, upper, depth, n_subcell_depth=n_sub)
    return dict(tri_pos=np.array(tri, np.float32), depth=depth, n_sub=n_sub)


def case_intersection(mode, n_trunc, transforms, eps=1e-3):
    m = _ref_modules()
    jnp, kd, mlp = m["jnp"], m["kd_tree"], m["mlp"]
    fA, pA = _load(m, "hammer", mode, **_mode_kwargs(mode, n_trunc))
    fB, pB = _load(m, "bunny", mode, **_mode_kwargs(mode, n_trunc))
    pB = mlp.prepend_op(pB, mlp.spatial_transformation())
    lower = jnp.array((-1.0, -1.0, -1.0))
    upper = jnp.array((1.0, 1.0, 1.0))
    found, locs, Rs, ts = [], [], [], []
    for th, t in transforms:
        R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]], np.float32)
        t = np.array(t, np.float32)
        pB["0000.spatial_transformation.R"] = jnp.array(R)
        pB["0000.spatial_transformation.t"] = jnp.array(t)
        f, ia, ib, loc = kd.find_any_intersection((fA, fB), (pA, pB), lower, upper, eps)
        found.append(bool(f))
        locs.append(np.array(loc, np.float32))
        Rs.append(R)
        ts.append(t)
    return dict(R=np.stack(Rs), t=np.stack(ts), found=np.array(found), loc=np.stack(locs), eps=eps, n_trunc=n_trunc)


def case_closest(name, mode, Q, eps, B):
    m = _ref_modules()
    jnp, kd = m["jnp"], m["kd_tree"]
    func, params = _load(m, name, mode)
    lower = jnp.array((-1.0, -1.0, -1.0))
    upper = jnp.array((1.0, 1.0, 1.0))
    q = np.random.default_rng(0).uniform(-1, 1, (Q, 3)).astype(np.float32)
    d, loc = kd.closest_point(func, params, lower, upper, jnp.array(q), eps=eps, batch_process_size=B)
    return dict(query_points=q, dist=np.array(d, np.float32), loc=np.array(loc, np.float32), eps=eps, B=B)


# ------------------------------------------------------------------------------------------------

CASES = {"mlps": (case_mlps, ())}
for _n in SAMPLES:
    CASES[f"points_{_n}"] = (case_points, (_n,))
    for _mode in ("interval", "affine_fixed", "affine_truncate", "affine_all"):
        CASES[f"classify_{_n}_{_mode}"] = (case_classify, (_n, _mode))
CASES["classify_hammer_affine_truncate64"] = (case_classify, ("hammer", "affine_truncate", 64))
for _n in SAMPLES:                                   # the affine_append bounder (n_append = 4)
    CASES[f"classify_{_n}_affine_append"] = (case_classify, (_n, "affine_append", 4))
for _n, _L in (("fox", 1.0), ("bunny", 2.0), ("hammer", 1.5), ("birdcage_occ", 4.0)):   # the sdf bounder
    CASES[f"classify_{_n}_sdf"] = (case_classify_sdf, (_n, _L))
for _n in SAMPLES:                                   # the slope_interval bounder
    CASES[f"classify_{_n}_slope_interval"] = (case_classify_slope, (_n,))
CASES["tree_fox_slope_d12"] = (case_tree, ("fox", "slope_interval"), dict(split_depth=12, with_exterior_nodes=True))
for _mode in ("interval", "affine_fixed", "affine_truncate", "affine_all", "slope_interval"):   # sin + encode ops
    CASES[f"pe_{_mode}"] = (case_pe, (_mode,))
CASES["render_fox_fixed_r10"] = (case_render, ("fox", "affine_fixed", 10))      # the caller of cast_rays
CASES["render_frustum_fox_fixed_r14"] = (case_render, ("fox", "affine_fixed", 14, True, 2))
CASES["frust_fox_fixed_r12_s4"] = (case_frustum, (("fox",), "affine_fixed", 12, 4))
CASES["frust_fox_bunny_interval_r10_s2_sub2"] = (case_frustum, (("fox", "bunny"), "interval", 10, 2, 2))
CASES["frust_hammer_fixed_r9_s3_sub3"] = (case_frustum, (("hammer",), "affine_fixed", 9, 3, 3))
CASES["frust_fox_fixed_r13x9_s3"] = (case_frustum, (("fox",), "affine_fixed", 13, 3), dict(res_y=9))   # res_x != res_y, fov_x != fov_y
CASES["frust_fox_slope_r10_s2"] = (case_frustum, (("fox",), "slope_interval", 10, 2))
CASES["frust_fox_trunc_r8_s2"] = (case_frustum, (("fox",), "affine_truncate", 8, 2))
CASES["tree_fox_sdf_d12"] = (case_tree, ("fox", "sdf", 1.0), dict(split_depth=12, with_interior_nodes=True))
CASES["tree_fox_append_d9"] = (case_tree, ("fox", "affine_append", 4), dict(split_depth=9))
CASES["rays_fox_fixed_r12"] = (case_cast_rays, (("fox",), "affine_fixed", 12))
CASES["rays_fox_interval_r6"] = (case_cast_rays, (("fox",), "interval", 6))
CASES["rays_fox_all_r6"] = (case_cast_rays, (("fox",), "affine_all", 6))
CASES["rays_fox_fixed_r8_sub3"] = (case_cast_rays, (("fox",), "affine_fixed", 8, 3))
CASES["rays_fox_bunny_fixed_r8"] = (case_cast_rays, (("fox", "bunny"), "affine_fixed", 8))
CASES["tree_fox_fixed_d12"] = (case_tree, ("fox", "affine_fixed"), dict(split_depth=12, with_interior_nodes=True, with_exterior_nodes=True))
CASES["tree_bunny_all_d9"] = (case_tree, ("bunny", "affine_all"), dict(split_depth=9, with_interior_nodes=True, with_exterior_nodes=True))
CASES["tree_fox_trunc_d9"] = (case_tree, ("fox", "affine_truncate"), dict(split_depth=9))
CASES["tree_fox_fixed_thresh"] = (case_tree, ("fox", "affine_fixed"), dict(node_terminate_thresh=300, offset=0.02))
CASES["tree_fox_fixed_b128"] = (case_tree, ("fox", "affine_fixed"), dict(split_depth=10, batch_process_size=128))
CASES["mc_fox_d4_s2"] = (case_mc, ("fox", "affine_fixed", 4, 2))
CASES["mc_bunny_d4_s3"] = (case_mc, ("bunny", "affine_fixed", 4, 3))
CASES["isect_fixed"] = (case_intersection, ("affine_fixed", 0, [(0.3, (0.0, 0.1, 0.05)), (0.3, (1.2, 0.1, 0.05)), (0.3, (1.6, 0.1, 0.05)), (1.1, (0.9, -0.4, 0.3))]))
CASES["isect_trunc64"] = (case_intersection, ("affine_truncate", 64, [(0.3, (1.2, 0.1, 0.05)), (0.3, (1.6, 0.1, 0.05))]))
# the stand-in runs vmap lanes in a Python loop, so the closest-point cases are kept to ~10k lanes; the two
# window sizes give DIFFERENT results for query 2 (0.528 vs 0.430): the order dependence of SURVEY.md F6
CASES["closest_fox_B4"] = (case_closest, ("fox", "affine_fixed", 3, 0.4, 4))
CASES["closest_fox_B256"] = (case_closest, ("fox", "affine_fixed", 3, 0.4, 256))


def case_min_distance(name):
    """slope_interval.SlopeIntervalImplicitFunction.min_distance_to_zero / min_distance_to_zero_in_direction
    (src/slope_interval.py:52-163) of the unmodified reference: axis-aligned boxes over 8 scales, rays, and swept boxes with
    one and two source-range vectors."""
    m = _ref_modules()
    jnp = m["jnp"]
    func, params = _load(m, name, "slope_interval")
    import zlib
    rng = np.random.default_rng(zlib.crc32(f"{name}-mindist".encode()) % 1000)
    n = 24
    cen = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    axis = (rng.uniform(0.2, 1.0, (n, 3)) * (2.0 ** -rng.integers(1, 9, (n, 1)))).astype(np.float32)
    prim, dist = [], []
    for i in range(n):
        p_, d_ = func.min_distance_to_zero(params, jnp.array(cen[i]), jnp.array(axis[i]), return_source_value=True)
        prim.append(float(p_)); dist.append(float(d_))
    src = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    bvec = (rng.standard_normal((n, 3)) * (2.0 ** -rng.integers(0, 7, (n, 1)))).astype(np.float32)
    rngv = (rng.standard_normal((n, 2, 3)) * (2.0 ** -rng.integers(3, 9, (n, 1, 1)))).astype(np.float32)
    r_val, r_dist, b1, b2 = [], [], [], []
    for i in range(n):
        v_, d_ = func.min_distance_to_zero_in_direction(params, jnp.array(src[i]), jnp.array(bvec[i]), return_source_value=True)
        r_val.append(float(v_)); r_dist.append(float(d_))
        b1.append([float(x) for x in func.min_distance_to_zero_in_direction(params, jnp.array(src[i]), jnp.array(bvec[i]),
                                                                             source_range=jnp.array(rngv[i, :1]), return_source_value=True)])
        b2.append([float(x) for x in func.min_distance_to_zero_in_direction(params, jnp.array(src[i]), jnp.array(bvec[i]),
                                                                             source_range=jnp.array(rngv[i]), return_source_value=True)])
    return dict(box_center=cen, box_axis_vec=axis, box_primal=np.array(prim, np.float32), box_distance=np.array(dist, np.float32),
                source=src, bound_vec=bvec, source_range=rngv, ray_value=np.array(r_val, np.float32), ray_distance=np.array(r_dist, np.float32),
                swept1=np.array(b1, np.float32), swept2=np.array(b2, np.float32))


for _n in ("fox", "bunny"):                          # missing item of round 1: the distance helpers of slope_interval.py
    CASES[f"mindist_{_n}_slope"] = (case_min_distance, (_n,))


def run(name):
    t0 = time.time()
    entry = CASES[name]
    fn, args = entry[0], entry[1]
    kw = entry[2] if len(entry) > 2 else {}
    sys.stdout = open(os.devnull, "w")          # the reference prints per-level progress
    try:
        out = fn(*args, **kw)
    finally:
        sys.stdout = sys.__stdout__
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    return name, time.time() - t0


def main():
    names = sys.argv[1:] or list(CASES)
    os.makedirs(GOLD, exist_ok=True)
    with mp.Pool(min(8, len(names))) as pool:
        for name, dt in pool.imap_unordered(run, names):
            print(f"{name}: {dt:.1f}s", flush=True)


if __name__ == "__main__":
    main()
