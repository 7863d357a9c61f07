"""torchrun --nproc-per-node N tools/nccl_frustum.py [res]: sharding.cast_rays_frustum_sharded over NCCL (initial tiles dealt
round-robin, all_reduce of the disjoint images + iteration counts) against the single-GPU call on rank 0, with timings."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "neural-implicit-queries_b200"), ROOT]
import implicit_mlp_utils  # noqa: E402
import queries  # noqa: E402
import render  # noqa: E402
import sharding  # noqa: E402

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
res = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
with np.load(os.path.join(ROOT, "tests", "golden", "mlps.npz")) as d:
    p = {k.split("/", 1)[1]: d[k] for k in d.files if k.startswith("fox/")}
f = implicit_mlp_utils.generate_implicit_from_params(p, "affine_fixed")
eye = np.array((2., 1., 2.), np.float32)
look, up, left = render.look_at(eye)
cam = (eye, look, up, left, 30., 30., res, res)
opts = queries.get_default_cast_opts()
for i in range(3):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    st, sh, sc, sn = sharding.cast_rays_frustum_sharded((f,), (p,), cam, opts)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
if dist.get_rank() == 0:
    t0 = time.perf_counter()
    t, hit, cnt, n = queries.cast_rays_frustum((f,), (p,), cam, opts)
    d1 = time.perf_counter() - t0
    same = np.array_equal(st, t) and np.array_equal(sh, hit) and np.array_equal(sc, cnt) and sn == n
    print(f"world {dist.get_world_size()} res {res}: sharded {dt * 1e3:.2f} ms, single {d1 * 1e3:.2f} ms, N_evals {sn} vs {n}, "
          f"hits {int((sh > 0).sum())}, bit-identical to the single-GPU call: {same}", flush=True)
    assert same
dist.barrier()
dist.destroy_process_group()
