import sys, time, numpy as np
sys.path[:0] = ["oracle", "neural-implicit-queries_b200", "."]
import bench, _niq, implicit_mlp_utils, kd_tree
ctx = _niq.default_context(0)
lo, hi = np.full(3, -1, np.float32), np.full(3, 1, np.float32)
for name, depths in (("bunny", (12, 15, 21)), ("fox", (12,)), ("birdcage_occ", (12,))):
    p = bench.sample_mlp(name)
    f = implicit_mlp_utils.generate_implicit_from_params(p, "affine_fixed")
    for d in depths:
        best = 1e9
        for _ in range(6):
            ctx.timer_start()
            t = kd_tree.build_tree(f, p, lo, hi, split_depth=d, ctx=ctx)
            ms = ctx.timer_stop()
            n = t.count(0); st = t.stats(); t.close()
            best = min(best, ms)
        print(f"{name} depth {d}: device {best:.3f} ms, leaves {n}, boxes {st['n_evals']}", flush=True)
p = bench.sample_mlp("bunny"); f = implicit_mlp_utils.generate_implicit_from_params(p, "affine_fixed")
for r in (0, 3):
    best = 1e9
    for _ in range(5):
        ctx.timer_start(); t = kd_tree.build_tree_dealt(f, p, lo, hi, 21, 12, r, 8, ctx=ctx); ms = ctx.timer_stop(); t.close(); best = min(best, ms)
    print(f"bunny d21 dealt rank {r}/8: device {best:.3f} ms", flush=True)
params = bench.synthetic_params()
f5 = implicit_mlp_utils.generate_implicit_from_params(params, "affine_fixed")
for d in (10, 14):
    best = 1e9
    for _ in range(5):
        ctx.timer_start(); t = kd_tree.build_tree(f5, params, lo, hi, split_depth=d, ctx=ctx); ms = ctx.timer_stop(); n = t.count(0); t.close(); best = min(best, ms)
    print(f"8x256 depth {d}: device {best:.3f} ms, leaves {n}", flush=True)
