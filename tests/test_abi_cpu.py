"""CPU-only checks of the boundary: the C-ABI library loads and exports every symbol include/niq.h
declares, the Python host layer mirrors the reference's module surface, and -- with no GPU -- the product
fails loudly instead of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, sample_params

LIB = os.path.join(ROOT, "neural-implicit-queries_b200", "libniq.so")


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    return LIB


def header_symbols():
    src = open(os.path.join(ROOT, "include", "niq.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(niq_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built):
    import _niq
    lib = ctypes.CDLL(built)
    syms = header_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/niq.h but not exported by libniq.so"
    assert sorted(_niq.SYMBOLS) == syms, "the ctypes binding and the header disagree"


def test_no_python_symbol_leaks_torch_types(built):
    # the boundary is plain C: no C++-mangled niq_ entry points
    out = os.popen(f"nm -D --defined-only {built}").read()
    exported = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert all(not e.startswith("_Z") or "niq" not in e.split("3niq")[0] for e in exported if e.startswith("niq_"))
    assert "niq_cast_rays" in exported


def test_mc_tables_match_oracle(built):
    import extract_cell
    from niq_oracle import mc_tables
    tri, ev, vc = extract_cell.get_mc_data()          # host-only call, no GPU needed
    otri, oev, ovc = mc_tables.unpack()
    np.testing.assert_array_equal(tri, otri)
    np.testing.assert_array_equal(ev, oev)
    np.testing.assert_array_equal(vc, ovc.astype(bool))


def test_module_surface_matches_reference():
    import affine, bucketing, extract_cell, implicit_function, implicit_mlp_utils, kd_tree, mlp, queries, render  # noqa: E401
    import sdf, slope_interval  # noqa: E401
    assert (implicit_function.SIGN_UNKNOWN, implicit_function.SIGN_POSITIVE, implicit_function.SIGN_NEGATIVE) == (0, 1, 2)
    for mod, names in ((mlp, "load save prepend_op spatial_transformation get_op_data n_ops func_from_spec build_spec "
                             "dense relu elu sin pow2_frequency_encode squeeze_last quick_mlp_spec initialize_params"),
                       (queries, "get_default_cast_opts cast_rays"),
                       (kd_tree, "construct_uniform_unknown_levelset_tree hierarchical_marching_cubes "
                                 "find_any_intersection closest_point sample_surface bulk_properties"),
                       (extract_cell, "get_mc_data extract_triangles_from_subcells"),
                       (bucketing, "get_next_bucket_size fits_in_smaller_bucket compactify_and_rebucket_arrays"),
                       (render, "camera_ray generate_camera_rays look_at"),
                       (implicit_mlp_utils, "generate_implicit_from_file"),
                       (sdf, "WeakSDFImplicitFunction"), (slope_interval, "SlopeIntervalImplicitFunction"),
                       (affine, "AffineContext AffineImplicitFunction")):
        for n in names.split():
            assert hasattr(mod, n), f"{mod.__name__}.{n} missing"
    assert queries.get_default_cast_opts()["n_max_step"] == 512
    with pytest.raises(ValueError):
        affine.AffineContext("bogus")
    with pytest.raises(RuntimeError):
        implicit_mlp_utils.generate_implicit_from_params({}, "nonsense")


def test_param_dict_helpers_follow_reference_grammar():
    import mlp
    p = sample_params("fox")
    assert mlp.n_ops(p) == 18
    name, args = mlp.get_op_data(p, 0)
    assert name == "dense" and args["A"].shape == (3, 32)
    p2 = mlp.prepend_op(p, mlp.spatial_transformation())
    assert mlp.n_ops(p2) == 19 and mlp.get_op_data(p2, 0)[0] == "spatial_transformation"
    assert mlp.get_op_data(p2, 1)[0] == "dense"
    with pytest.raises(ValueError):
        mlp.n_ops({"oops.dense.A": np.zeros(3)})
    spec = mlp.build_spec(mlp.quick_mlp_spec([3, 16, 16, 1], "relu"))
    init = mlp.initialize_params(spec, 0)
    assert init["0000.dense.A"].shape == (3, 16) and init["0004.dense.b"].shape == (1,)
    with pytest.raises(ValueError):
        mlp.quick_mlp_spec([3, 4, 1], "gelu")       # the reference accepts relu / elu only (src/mlp.py:86-90); 'tanh' is ours
    assert "0001.tanh._" in mlp.initialize_params(mlp.build_spec(mlp.quick_mlp_spec([3, 4, 1], "tanh")), 0)


def test_bucketing_matches_oracle():
    import bucketing
    from niq_oracle import rays
    for s in (1, 127, 128, 129, 5000, 2 ** 20 + 1):
        assert bucketing.get_next_bucket_size(s) == rays.get_next_bucket_size(s)
    m = np.array([0, 1, 1, 0, 1], bool)
    a = np.arange(10, dtype=np.float32).reshape(5, 2)
    om, n, oa = bucketing.compactify_and_rebucket_arrays(m, 128, a)
    rm, rn, ra = rays.compactify_and_rebucket(m, 128, a)
    assert n == rn == 3
    np.testing.assert_array_equal(om, rm)
    np.testing.assert_array_equal(oa, ra)


def test_camera_rays_match_oracle():
    import render
    from niq_oracle import rays
    eye = np.array((2., 1., 2.), np.float32)
    look, up, left = render.look_at(eye)
    olook, oup, oleft = rays.look_at(eye)
    np.testing.assert_array_equal(look, olook)
    np.testing.assert_array_equal(up, oup)
    r, d = render.generate_camera_rays(eye, look, up, res=24, fov_deg=30., res_y=10)
    orr, od = rays.generate_camera_rays(eye, olook, oup, res=24, fov_deg=30., res_y=10)
    np.testing.assert_array_equal(r, orr)
    np.testing.assert_array_equal(d, od)


def test_fails_loudly_without_gpu(built):
    """No CPU fallback: without a CUDA device every compute entry raises (NIQ_ECUDA)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import _niq
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        _niq.Context(0)
    import implicit_mlp_utils
    func = implicit_mlp_utils.generate_implicit_from_params(sample_params("fox"), "affine_fixed")
    with pytest.raises(RuntimeError):
        func.classify_box(sample_params("fox"), np.zeros(3, np.float32), np.ones(3, np.float32))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "neural-implicit-queries_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "niq_oracle" not in src and "jaxshim" not in src, f"{f} references the oracle"


def _header_struct_fields(name):
    src = open(os.path.join(ROOT, "include", "niq.h")).read()
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = re.findall(r"([A-Za-z_][A-Za-z0-9_]*)\s*(?:\[\d+\])?\s*(?:,|$)", decl.split(None, 1)[1] if " " in decl else decl)
        fields += names
    return fields


def test_struct_mirrors_do_not_drift(built):
    """include/niq.h, the ctypes mirrors in _niq.py and the reference-side stubs printed in INTEGRATION.md describe the same
    structs (field names, order and -- for ctypes -- sizes): an array of niq_mode_cfg built from a stale stub would be read at
    the wrong stride."""
    import _niq
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for cname, mirror, stub in (("niq_mode_cfg", _niq.ModeCfg, "_Cfg"), ("niq_cast_opts", _niq.CastOpts, "_Opts"),
                                ("niq_op_desc", _niq.OpDesc, "_Op"), ("niq_camera", _niq.Camera, "_Cam")):
        hdr = _header_struct_fields(cname)
        assert [f[0] for f in mirror._fields_] == hdr, (cname, hdr)
        block = re.search(r"class %s\(C\.Structure\):.*?_fields_ = \[(.*?)\]\n" % stub, doc, flags=re.S).group(1)
        assert re.findall(r'\("([a-z_A-Z0-9]+)"', block) == hdr, (stub, hdr)
    assert ctypes.sizeof(_niq.ModeCfg) == 16 and ctypes.sizeof(_niq.CastOpts) == 32
    # every op kind / mode of the header is known to the binding
    src = open(os.path.join(ROOT, "include", "niq.h")).read()
    assert len(re.findall(r"NIQ_OP_[A-Z0-9_]+ = \d", src)) == len(_niq._OP_KINDS)
    assert len(re.findall(r"NIQ_MODE_[A-Z0-9_]+ = \d", src)) == len(_niq.MODE_IDS)


def test_params_digest_follows_content_not_identity():
    """The handle cache key (_niq.params_digest): equal content -> equal key whatever the objects; any changed byte, shape,
    dtype or key name -> another key.  The reference reads `params` afresh on every call (src/main_intersection.py:171-183
    mutates the transform entries in place), so identity may never be the key."""
    import _niq
    p = sample_params("bunny")
    d = _niq.params_digest(p)
    assert _niq.params_digest({k: np.array(v) for k, v in p.items()}) == d
    rng = np.random.default_rng(0)
    big = [k for k in p if np.asarray(p[k]).nbytes > 256]
    small = [k for k in p if 0 < np.asarray(p[k]).nbytes <= 256]
    for k in [big[0], big[-1], small[0]]:
        for _ in range(8):                                   # one flipped mantissa bit anywhere in the array
            q = {kk: np.array(v) for kk, v in p.items()}
            flat = q[k].reshape(-1).view(np.uint32)
            flat[rng.integers(flat.size)] ^= np.uint32(1)
            assert _niq.params_digest(q) != d
    k = big[0]
    q = dict(p); q[k] = np.ascontiguousarray(p[k].T.reshape(p[k].shape)) if p[k].ndim == 2 else p[k][::-1].copy()
    assert _niq.params_digest(q) != d or np.array_equal(q[k], p[k])      # permuted entries (a plain sum would not notice)
    q = dict(p); q[k] = p[k].reshape(-1)
    assert _niq.params_digest(q) != d                                      # same bytes, another shape
    q = dict(p); q[k] = p[k].astype(np.float64)
    assert _niq.params_digest(q) != d
    q = dict(p); q["9999.extra._"] = np.zeros(0, np.float32)
    assert _niq.params_digest(q) != d
