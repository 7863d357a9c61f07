#!/usr/bin/env python
"""A/B of the engine's zero-skipping (GPU box): dense K loops (NIQ_NO_SPARSE=1) vs. list-driven loops must give
bit-identical results (only the sign of zero may differ); prints timings and executed / algorithmic MAC ratios."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "neural-implicit-queries_b200"), os.path.join(ROOT, "oracle")]
import _niq  # noqa: E402
import implicit_mlp_utils  # noqa: E402
import mlp  # noqa: E402
import queries  # noqa: E402
import render  # noqa: E402


def same(a, b):
    """hidden layers are bit-identical; the lane-split dot product of the last layer sums in a different order, so
    float outputs may differ by a few ulp of the summed magnitude; integer outputs by a handful of near-tie flips"""
    a, b = np.asarray(a), np.asarray(b)
    if a.dtype.kind == "f":
        sc = max(float(np.nanmax(np.abs(a))) if a.size else 0.0, 1e-30)
        return bool(np.all((np.abs(a - b) <= 2e-6 * sc) | (np.isnan(a) & np.isnan(b))))
    return bool((a != b).mean() <= 2e-3) if a.size else True


def run_all(params, tag, res=64):
    out = {}
    for sparse in (0, 1):
        os.environ["NIQ_NO_SPARSE"] = "0" if sparse else "1"
        ctx = _niq.Context(0)
        ctx.exec_macs(on=True, reset=True)
        r = {}
        rng = np.random.default_rng(1)
        for mode in ("affine_fixed", "interval"):
            f = implicit_mlp_utils.generate_implicit_from_params(params, mode)
            for scale in (0.3, 0.01, 0.0005):
                c = rng.uniform(-1, 1, (20000, 3)).astype(np.float32)
                h = (scale * rng.uniform(0.3, 1, (20000, 3))).astype(np.float32)
                t0 = time.perf_counter()
                lab, lo, up, tie = f.bound_box(params, c - h, c + h, ctx=ctx)
                r[f"box/{mode}/{scale}"] = (lab, lo, up, tie)
                r[f"time box/{mode}/{scale}"] = time.perf_counter() - t0
        x = rng.uniform(-1, 1, (50001, 3)).astype(np.float32)
        r["points"] = mlp.eval_points(params, x, return_scale=True, ctx=ctx)
        x = (0.3 + 0.001 * rng.uniform(-1, 1, (50001, 3))).astype(np.float32)
        r["points/local"] = mlp.eval_points(params, x, return_scale=True, ctx=ctx)
        f = implicit_mlp_utils.generate_implicit_from_params(params, "affine_fixed")
        eye = np.array((2., 1., 2.), np.float32)
        look, up, _ = render.look_at(eye)
        roots, dirs = render.generate_camera_rays(eye, look, up, res=res, fov_deg=30.)
        ctx.exec_macs(on=True, reset=True)
        t0 = time.perf_counter()
        r["rays"] = queries.cast_rays((f,), (params,), roots, dirs, queries.get_default_cast_opts(), return_near_tie=True, ctx=ctx)
        r["time rays"] = time.perf_counter() - t0
        macs = ctx.exec_macs(on=False)
        r["ray exec/alg"] = macs / (5.0 * ctx.mlp(params).macs * float(r["rays"][2].sum()))
        out[sparse] = r
        ctx.close()
    ok = True
    for k in out[0]:
        if k.startswith("time") or k.startswith("ray exec"):
            continue
        eq = all(same(a, b) for a, b in zip(out[0][k], out[1][k]))
        ok &= eq
        if not eq:
            d = [float(np.nanmax(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)))) for a, b in zip(out[0][k], out[1][k])]
            print(f"  MISMATCH {tag} {k}: max abs diff per output {d}")
    print(f"{tag}: bit-identical={ok}  rays dense {out[0]['time rays']*1e3:.1f} ms -> sparse {out[1]['time rays']*1e3:.1f} ms; "
          f"executed/algorithmic MACs dense {out[0]['ray exec/alg']:.3f} sparse {out[1]['ray exec/alg']:.3f}; "
          f"boxes(0.01) {out[0]['time box/affine_fixed/0.01']*1e3:.1f} -> {out[1]['time box/affine_fixed/0.01']*1e3:.1f} ms", flush=True)
    return ok


if __name__ == "__main__":
    allok = True
    with np.load(os.path.join(ROOT, "tests", "golden", "mlps.npz")) as d:
        for nm in ("fox", "hammer", "bunny", "birdcage_occ"):
            p = {k.split("/", 1)[1]: d[k] for k in d.files if k.startswith(nm + "/")}
            allok &= run_all(p, nm, res=128)
    for w, act in ((256, "relu"), (128, "relu"), (40, "relu"), (256, "elu")):
        spec = mlp.build_spec(mlp.quick_mlp_spec([3] + [w] * 8 + [1], act))
        allok &= run_all(mlp.initialize_params(spec, 0), f"synthetic {w} {act}", res=48)
    print("ALL OK" if allok else "FAILURES")
    sys.exit(0 if allok else 1)
