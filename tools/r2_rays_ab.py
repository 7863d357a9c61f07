#!/usr/bin/env python
"""Development A/B of engine builds on the headline ray workload (GPU box).
usage: python tools/r2_rays_ab.py <n_tiles | fox> <libA.so | dephase=<cycles>> [...]     ('-' = the in-tree libniq.so)
Every library runs in its own process (NIQ_LIB); prints rays/s, executed TFLOP/s and whether (t, hit, count) equal the
first library's bit for bit."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker_fox(out):
    """BASELINE configs[0]: fox.npz, 512 x 512 rays, affine_fixed (resident 32-wide net); NIQ_DEPHASE from the environment."""
    sys.path.insert(0, ROOT)
    import bench  # noqa: E402
    import torch

    import _niq
    import implicit_mlp_utils
    import queries
    import render
    ctx = _niq.default_context(0)
    p = bench.sample_mlp("fox")
    f = implicit_mlp_utils.generate_implicit_from_params(p, "affine_fixed")
    eye = np.array((2., 1., 2.), np.float32)
    look, up, _ = render.look_at(eye)
    roots, dirs = render.generate_camera_rays(eye, look, up, res=512, fov_deg=30.)
    n = roots.shape[0]
    dev = torch.device("cuda", 0)
    r_d, d_d = torch.from_numpy(roots).to(dev), torch.from_numpy(dirs).to(dev)
    t_d = torch.zeros(n, dtype=torch.float32, device=dev)
    h_d = torch.zeros(n, dtype=torch.int32, device=dev)
    c_d = torch.zeros(n, dtype=torch.int32, device=dev)
    opts = queries.get_default_cast_opts()

    def step():
        queries.cast_rays_device((f,), (p,), n, r_d.data_ptr(), d_d.data_ptr(), t_d.data_ptr(), h_d.data_ptr(), c_d.data_ptr(), opts,
                                 want_n_evals=False, ctx=ctx)
    step(); step()
    ms = []
    for _ in range(5):
        ctx.timer_start()
        step()
        ms.append(ctx.timer_stop())
    np.savez(out, t=t_d.cpu().numpy(), h=h_d.cpu().numpy(), c=c_d.cpu().numpy())
    print(json.dumps({"lib": os.environ.get("NIQ_LIB", "in-tree"), "dephase": os.environ.get("NIQ_DEPHASE", "default"), "workload": "fox 512^2",
                      "ms": round(min(ms), 3), "ray_steps_per_s": round(float(c_d.sum().item()) / min(ms) * 1e3)}), flush=True)


def worker(n_tiles, out):
    sys.path.insert(0, ROOT)
    import bench  # noqa: E402  (sets sys.path for the package)
    import torch

    import _niq
    import implicit_mlp_utils
    import queries

    ctx = _niq.default_context(0)
    params = bench.synthetic_params()
    func = implicit_mlp_utils.generate_implicit_from_params(params, "affine_fixed")
    roots, dirs = bench.camera_rays()
    opts = queries.get_default_cast_opts()
    tiles, ntx = bench.chosen_tiles(n_tiles)
    mine = bench.pixels_of_tiles(tiles, ntx)
    n = int(mine.shape[0])
    dev = torch.device("cuda", 0)
    r_d, d_d = torch.from_numpy(roots[mine]).to(dev), torch.from_numpy(dirs[mine]).to(dev)
    t_d = torch.zeros(n, dtype=torch.float32, device=dev)
    h_d = torch.zeros(n, dtype=torch.int32, device=dev)
    c_d = torch.zeros(n, dtype=torch.int32, device=dev)
    peak = ctx.fp32_peak_tflops()

    def step():
        queries.cast_rays_device((func,), (params,), n, r_d.data_ptr(), d_d.data_ptr(), t_d.data_ptr(), h_d.data_ptr(),
                                 c_d.data_ptr(), opts, want_n_evals=False, ctx=ctx)
    step()
    torch.cuda.synchronize()
    ctx.exec_macs(on=True, reset=True)
    ms = []
    for _ in range(3):
        ctx.timer_start()
        step()
        ms.append(ctx.timer_stop())
    macs = ctx.exec_macs(on=False, reset=True)
    np.savez(out, t=t_d.cpu().numpy(), h=h_d.cpu().numpy(), c=c_d.cpu().numpy())
    tf = 2 * macs / (sum(ms) * 1e-3) / 1e12
    print(json.dumps({"lib": os.environ.get("NIQ_LIB", "in-tree"), "rays": n, "ms": round(min(ms), 2),
                      "rays_per_s": round(n / min(ms) * 1e3, 1), "exec_tflops": round(tf, 2), "peak": round(peak, 2),
                      "frac": round(tf / peak, 4)}), flush=True)


def main():
    if sys.argv[1] == "--worker":
        return worker_fox(sys.argv[3]) if sys.argv[2] == "fox" else worker(int(sys.argv[2]), sys.argv[3])
    n_tiles = sys.argv[1]
    first = None
    for i, lib in enumerate(sys.argv[2:] or ["-"]):
        env = dict(os.environ)
        if lib.startswith("dephase="):            # "dephase=<cycles>": the in-tree library with NIQ_DEPHASE set
            env["NIQ_DEPHASE"] = lib.split("=", 1)[1]
        elif lib != "-":
            env["NIQ_LIB"] = os.path.abspath(lib)
        out = f"/tmp/r2_ab_{i}.npz"
        subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", n_tiles, out], env=env, check=True)
        r = np.load(out)
        if first is None:
            first = r
        print(f"   equal to first: {all(np.array_equal(r[k], first[k]) for k in ('t', 'h', 'c'))}", flush=True)


if __name__ == "__main__":
    main()
