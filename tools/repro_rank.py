"""development: cast the rays bench.py gives to (rank, world) on ONE gpu, with a watchdog"""
import faulthandler, os, sys, time
faulthandler.dump_traceback_later(int(os.environ.get("WD", "60")), exit=True)
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neural-implicit-queries_b200"), os.path.join(ROOT, "oracle")]
import bench, _niq, implicit_mlp_utils, queries, sharding
rank, world, tiles = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
params = bench.synthetic_params()
func = implicit_mlp_utils.generate_implicit_from_params(params, "affine_fixed")
roots, dirs = bench.camera_rays()
ntx, nty = sharding.tile_ids(bench.RES_X, bench.RES_Y, bench.TILE)
stride = max(1, (ntx * nty) // (tiles * world))
mine = sharding.rank_pixels(bench.RES_X, bench.RES_Y, bench.TILE, rank, world, stride)[:tiles * 256]
print("rays", mine.shape, "stride", stride, flush=True)
ctx = _niq.default_context(0)
opts = queries.get_default_cast_opts()
if len(sys.argv) > 4:
    opts["n_max_step"] = int(sys.argv[4])
t0 = time.time()
t, hit, cnt, ne = queries.cast_rays((func,), (params,), roots[mine], dirs[mine], opts, ctx=ctx)
print(f"done in {time.time()-t0:.2f}s hits {(hit>0).sum()} steps {cnt.sum()} t[min,max]=({t.min()},{t.max()}) nan {np.isnan(t).sum()}", flush=True)
