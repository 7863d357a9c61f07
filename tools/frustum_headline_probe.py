"""cast_rays_frustum on the HEADLINE network (random-init 3->256x8->1 ReLU, weights streamed through the ring): kernel time,
executed and algorithmic FP32 rate of k_cast_frustum<256>.  Every frustum of this net crawls to the step limit, so the image is
kept small and n_max_step is stated.  usage: python tools/frustum_headline_probe.py [res] [n_max_step]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "neural-implicit-queries_b200"), ROOT]
import _niq  # noqa: E402
import bench  # noqa: E402
import implicit_mlp_utils  # noqa: E402
import queries  # noqa: E402
import render  # noqa: E402

res = int(sys.argv[1]) if len(sys.argv) > 1 else 192
n_max = int(sys.argv[2]) if len(sys.argv) > 2 else 128
ctx = _niq.default_context()
p = bench.synthetic_params()
f = implicit_mlp_utils.generate_implicit_from_params(p, "affine_fixed")
M = sum(int(np.prod(v.shape)) for k, v in p.items() if k.endswith("dense.A"))
eye = np.array((2., 1., 2.), np.float32)
look, up, left = render.look_at(eye)
cam = (eye, look, up, left, 30., 30., res, res)
opts = queries.get_default_cast_opts()
opts["n_max_step"] = n_max
peak = ctx.fp32_peak_tflops()
for rep in range(2):
    ctx.kernel_timing(True); ctx.kernel_ms(0, reset=True); ctx.exec_macs(on=True, reset=True)
    it = []
    t, hit, cnt, n_evals = queries.cast_rays_frustum((f,), (p,), cam, opts, iter_counts=it, ctx=ctx)
    ms, n_launch = ctx.kernel_ms(0, reset=True)
    macs = ctx.exec_macs(on=False, reset=True)
    ctx.kernel_timing(False)
    alive, steps = 16 * 16, 0
    for a, b in it:
        steps += alive
        alive += b - a
    out = {"net": "3->256x8->1 relu (bench.synthetic_params)", "res": res, "n_max_step": n_max, "kernel_ms": ms, "frustum_steps": steps,
           "finished_frusta": sum(a for a, _ in it), "executed_tflops": 2 * macs / ms / 1e9, "algorithmic_tflops": 14 * M * steps / ms / 1e9,
           "measured_ffma_peak_tflops": peak, "frac_executed": 2 * macs / ms / 1e9 / peak, "hits": int((hit > 0).sum()), "n_evals": n_evals}
    print(json.dumps(out), flush=True)
