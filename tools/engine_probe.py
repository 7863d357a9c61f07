#!/usr/bin/env python
"""Engine micro-benchmark (GPU box): times k_classify_fixed / k_eval_points on a device-resident batch for
one or more builds of the library (variants compiled with -DNIQ_VARIANT=n into build/).  Prints TFLOP/s
(algorithmic: 10*M per box, 2*M per point) per build.  Development tool, not part of the product."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "neural-implicit-queries_b200"), os.path.join(ROOT, "oracle")]


def run(lib_path, width, n_boxes, reps=5, act="relu"):
    import importlib
    import _niq
    _niq._lib = None
    _niq.LIB_PATH = lib_path
    _niq._contexts.clear()
    import implicit_mlp_utils
    import mlp
    spec = mlp.build_spec(mlp.quick_mlp_spec([3] + [width] * 8 + [1], act))
    params = mlp.initialize_params(spec, 0)
    ctx = _niq.Context(0)
    m = ctx.mlp(params)
    rng = np.random.default_rng(0)
    c = rng.uniform(-1, 1, (n_boxes, 3)).astype(np.float32)
    h = np.full((n_boxes, 3), 1e-3, np.float32)
    lo, hi = ctx.to_device(c - h), ctx.to_device(c + h)
    lab = ctx.alloc(4 * n_boxes)
    cfg = _niq.ModeCfg(1, 0, 0)
    L = _niq.lib()

    def call():
        _niq.check(L.niq_classify_boxes(ctx.handle, m.handle, C.byref(cfg), C.c_int64(n_boxes), lo, hi, C.c_float(0.0),
                                        lab, None, None, None, C.c_int(_niq.MEM_DEVICE)))
    call()
    best = 1e30
    for _ in range(reps):
        ctx.timer_start()
        call()
        best = min(best, ctx.timer_stop())
    tf = 10 * m.macs * n_boxes / (best * 1e-3) / 1e12
    labs = ctx.download(lab, (n_boxes,), np.int32)
    peak = ctx.fp32_peak_tflops()
    ctx.close()
    return tf, best, peak, int(labs.sum())


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    widths = [int(a.split("=")[1]) for a in sys.argv[1:] if a.startswith("--width=")] or [256, 64, 32]
    libs = args or [os.path.join(ROOT, "neural-implicit-queries_b200", "libniq.so")]
    for width in widths:
        n = 148 * 16 * (64 if width == 256 else 512)
        for lib in libs:
            tf, ms, peak, chk = run(lib, width, n)
            print(f"W={width:3d} {os.path.basename(lib):24s} {tf:7.2f} TFLOP/s  ({ms:8.3f} ms, {n} boxes, ffma peak {peak:.1f}, frac {tf / peak:.3f}, chk {chk})", flush=True)
