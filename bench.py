#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the range-analysis hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (BASELINE.json configs[4], the one its metric "rays/sec (1080p cast_rays)" is quoted on):
random-init 3->256x8->1 ReLU MLP (glorot-normal A, b~N(0,1e-2^2), NumPy seed 0), 1920x1080 pinhole rays
(eye (2,1,2), look-at origin, fov 30), queries.cast_rays, affine_fixed, default opts.  With this network every
ray takes exactly n_max_step = 512 steps and none hits, so the full image is 1.06e9 ray-steps
= 4.9 EFLOP; a "step" of this bench is therefore a STATED SUB-SAMPLE: every `tile_stride`-th 16x16 pixel
tile of the image, dealt round-robin over the ranks (per-ray work is uniform, so rays/s of the sub-sample is
rays/s of the image).  Per-GPU work is fixed as N grows ("weak").  One JSON line on stdout from rank 0.

value   = rays/s, whole job, ray buffers already resident in HBM, CUDA-event timed on the context stream.
e2e     = rays/s through the public Python API (queries.cast_rays, NumPy host buffers from pinned memory,
          H2D + D2H inside the timed region), + for N > 1 the NCCL all_gather of the results.
roofline= FP32-FMA bound (NOT hbm / tensor: 2.3 GFLOP per 24 B ray; tensor cores would break the 1e-5
          parity bar): algorithmic flops (10*M per ray-step) / kernel time vs the FFMA peak
          measured on this GPU by a register-only FFMA kernel.
cpu_baseline / --impl reference = the CPU oracle (NumPy restatement of the reference; JAX is not installable
          here, so the reference itself cannot run) on the host cores over a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "neural-implicit-queries_b200"), ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

RES_X, RES_Y, TILE = 1920, 1080, 16
LAYERS = [3] + [256] * 8 + [1]
METRIC = "rays/sec (1080p cast_rays, synthetic 8x256 ReLU MLP, affine_fixed)"


def synthetic_params():
    import mlp
    spec = mlp.build_spec(mlp.quick_mlp_spec(LAYERS, "relu"))
    return mlp.initialize_params(spec, 0)


def camera_rays():
    import render
    eye = np.array((2., 1., 2.), np.float32)
    look, up, _ = render.look_at(eye)
    return render.generate_camera_rays(eye, look, up, res=RES_X, fov_deg=30., res_y=RES_Y)


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md recipe), sampled every 200 ms
    in-process through NVML (pynvml) -- a polling nvidia-smi process measurably delays kernel launches."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, device):
        self.device, self.samples, self.stop_flag, self.thread, self.err = device, [], threading.Event(), None, None

    def _resolve_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.device < len(ids) and ids[self.device].isdigit():
                return int(ids[self.device])
        return self.device

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._resolve_index())
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:           # noqa: BLE001
            self.err = f"nvml unavailable: {e}"
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        while not self.stop_flag.is_set():
            try:
                mhz = float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                try:
                    reasons = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:        # noqa: BLE001
                    reasons = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                power = self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.samples.append((mhz, reasons, power))
            except Exception as e:       # noqa: BLE001
                self.err = str(e)
            self.stop_flag.wait(0.2)

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err or "no sampler"]}
        self.stop_flag.set()
        self.thread.join(2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err or "no samples"]}
        mhz = [s[0] for s in self.samples]
        mask = 0
        for s in self.samples:
            mask |= s[1]
        return {"sm_mhz": float(np.median(mhz)), "sm_max_mhz": self.smax,
                "reasons": sorted(k for k, bit in self.REASONS.items() if mask & bit),
                "power_w_max": max(s[2] for s in self.samples), "samples": len(mhz)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle on the host cores (bounded sample)
# ------------------------------------------------------------------------------------------------

def _cpu_chunk(args):
    params, roots, dirs, opts = args
    from threadpoolctl import threadpool_limits
    from niq_oracle import net, rays
    with threadpool_limits(limits=1):
        t, hit, cnt, n_evals = rays.cast_rays((net.AffineContext("affine_fixed"),), (params,), roots, dirs, opts)
    return int(cnt.sum())


def cpu_cast_rays(params, roots, dirs, opts, pool, n_workers):
    """The oracle's cast_rays over `roots`, rays split over `n_workers` processes (1 BLAS thread each)."""
    chunks = [(params, roots[i::n_workers], dirs[i::n_workers], opts) for i in range(n_workers) if roots[i::n_workers].shape[0]]
    t0 = time.perf_counter()
    steps = sum(pool.map(_cpu_chunk, chunks))
    return time.perf_counter() - t0, steps


def cpu_workers():
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return max(1, min(n, 64))


def sample_rays(roots, dirs, n, seed=0):
    """A bounded sample of the workload's rays: whole 16x16 tiles, seeded."""
    import sharding
    ntx, nty = sharding.tile_ids(RES_X, RES_Y, TILE)
    rng = np.random.default_rng(seed)
    tiles = rng.choice(ntx * nty, size=max(1, n // (TILE * TILE)), replace=False)
    idx = []
    for tl in tiles:
        ty, tx = divmod(int(tl), ntx)
        yy, xx = np.meshgrid(np.arange(ty * TILE, min((ty + 1) * TILE, RES_Y)), np.arange(tx * TILE, min((tx + 1) * TILE, RES_X)), indexing="ij")
        idx.append((yy * RES_X + xx).reshape(-1))
    idx = np.concatenate(idx)[:n]
    return roots[idx], dirs[idx]


def run_reference(args):
    """--impl reference: the CPU oracle port (the reference is pure Python on JAX, which is not installable in
    this image, so there is nothing to build into oracle/_ref) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from niq_oracle import net, rays               # oracle only: nothing of the product runs on this arm
    params = net.random_mlp(LAYERS, "relu", seed=0)          # same bits as synthetic_params() (tests/test_bench_cpu.py)
    eye = np.array((2., 1., 2.), np.float32)
    look, up, _ = rays.look_at(eye)
    roots, dirs = rays.generate_camera_rays(eye, look, up, res=RES_X, fov_deg=30., res_y=RES_Y)
    opts = rays.get_default_cast_opts()
    nw = cpu_workers()
    n_sample = int(os.environ.get("NIQ_BENCH_CPU_RAYS", 32 * nw))     # rays per step: a few seconds of CPU work
    r, d = sample_rays(roots, dirs, n_sample)
    n_sample = r.shape[0]
    with mp.get_context("fork").Pool(nw) as pool:
        for _ in range(args.warmup):
            cpu_cast_rays(params, r[:nw], d[:nw], opts, pool, nw)
        total, steps = 0.0, 0
        for _ in range(args.steps):
            dt, st = cpu_cast_rays(params, r, d, opts, pool, nw)
            total += dt
            steps += st
    value = n_sample * args.steps / total
    sample = f"{n_sample} rays (whole 16x16 tiles, seed 0) x all 512 steps per step, rays split over {nw} processes"
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(n_sample, 1, None),
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": nw, "kind": "port", "sample": sample,
                         "ray_steps_per_s": steps / total},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def workload_config(rays_per_step, world, tile_stride):
    return {
        "workload": "BASELINE configs[4]: random-init 3->256x8->1 ReLU MLP (NumPy seed 0), 1920x1080 camera rays, "
                    "queries.cast_rays affine_fixed, default opts (n_max_step 512: every ray runs all 512 steps, no hits)",
        "rays_per_step": int(rays_per_step), "ray_steps_per_ray": 512, "image": [RES_X, RES_Y], "tile": TILE,
        "tile_stride": tile_stride, "sub_sample": "every tile_stride-th 16x16 tile of the image, dealt round-robin to ranks; "
                                                  "the full image is 2,073,600 rays = 4.9 EFLOP",
        "parallelism": f"ray tiles x{world}", "l2": "256 MiB device buffer rewritten between timed steps (untimed)",
    }


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------

def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: this backend has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import _niq
    import implicit_mlp_utils
    import queries
    import sharding

    ctx = _niq.default_context(local)
    params = synthetic_params()
    func = implicit_mlp_utils.generate_implicit_from_params(params, "affine_fixed")
    M = ctx.mlp(params).macs
    flop_per_ray_step = 10 * M                     # 3 affine rows + 2 point rows, 2 flop per MAC
    roots, dirs = camera_rays()
    opts = queries.get_default_cast_opts()

    ntx, nty = sharding.tile_ids(RES_X, RES_Y, TILE)
    tiles_per_rank = args.tiles
    tile_stride = max(1, (ntx * nty) // (tiles_per_rank * world))
    mine = sharding.rank_pixels(RES_X, RES_Y, TILE, rank, world, tile_stride)[:tiles_per_rank * TILE * TILE]
    n = int(mine.shape[0])
    r_h = torch.from_numpy(roots[mine]).pin_memory()
    d_h = torch.from_numpy(dirs[mine]).pin_memory()
    dev = torch.device("cuda", local)
    r_d, d_d = r_h.to(dev), d_h.to(dev)
    t_d = torch.zeros(n, dtype=torch.float32, device=dev)
    h_d = torch.zeros(n, dtype=torch.int32, device=dev)
    c_d = torch.zeros(n, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    cap = tiles_per_rank * TILE * TILE            # shards differ in size (partial border tiles): the gather is padded to cap
    gathered = torch.empty((world, cap, 3), dtype=torch.int32, device=dev) if world > 1 else None
    pack_h = torch.zeros((cap, 3), dtype=torch.int32).pin_memory() if world > 1 else None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        queries.cast_rays_device((func,), (params,), n, r_d.data_ptr(), d_d.data_ptr(), t_d.data_ptr(), h_d.data_ptr(),
                                 c_d.data_ptr(), opts, want_n_evals=False, ctx=ctx)

    def e2e_step():
        out = queries.cast_rays((func,), (params,), r_h.numpy(), d_h.numpy(), opts, ctx=ctx)
        if world > 1:        # the one collective of the path: gather (t, hit, count) = 12 B/ray over NVLink
            pack_h[:n] = torch.from_numpy(np.stack((out[0].view(np.int32), out[1], out[2]), axis=1))
            dist.all_gather_into_tensor(gathered.view(-1, 3), pack_h.to(dev, non_blocking=True))
            torch.cuda.synchronize()
        return out

    # ---- warm-up ----
    for _ in range(max(args.warmup, 3)):
        device_step()
    peak_tflops = ctx.fp32_peak_tflops()

    # ---- timed: K steps, CUDA events on the context stream, L2 rewritten between steps ----
    launches0 = ctx.launch_count()
    ctx.kernel_timing(True)
    ctx.kernel_ms(0, reset=True)
    ctx.exec_macs(on=True, reset=True)       # executed multiply-adds (the kernels skip exactly-zero columns after relu)
    clocks = ClockSampler(local)
    if not os.environ.get("NIQ_BENCH_NO_CLOCKS"):
        clocks.start()
    barrier()
    step_ms = []
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        ctx.timer_start()
        device_step()
        step_ms.append(ctx.timer_stop())
    barrier()
    clk = clocks.stop()
    kernel_ms, kernel_launches = ctx.kernel_ms(0, reset=True)
    exec_macs = ctx.exec_macs(on=False, reset=True)
    ctx.kernel_timing(False)
    gpu_launches = ctx.launch_count() - launches0
    total_ms = float(sum(step_ms))
    ray_steps = int(c_d.sum().item())

    # ---- e2e through the public API (host buffers) ----
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---- second half of BASELINE's metric: kd-tree boxes/s on the same network (configs[4]: depth-14 level-set tree).
    # The tree of this network is full (nothing is pruned: 32,767 box classifications).  N = 1: the whole tree;
    # N > 1: top levels replicated, subtrees dealt round-robin (sharding.tree_sharded) = strong scaling of one tree.
    import kd_tree
    lo3, hi3 = np.full(3, -1, np.float32), np.full(3, 1, np.float32)
    TREE_DEPTH = 14

    def tree_step():
        if world == 1:
            st = {}
            out = kd_tree.construct_uniform_unknown_levelset_tree(func, params, lo3, hi3, split_depth=TREE_DEPTH, stats=st, ctx=ctx)
            return int(out["unknown_node_valid"].sum()), st["n_evals"]
        lo_l, hi_l = sharding.tree_sharded(func, params, lo3, hi3, TREE_DEPTH, ctx=ctx)
        return int(lo_l.shape[0]), None

    tree_step()
    barrier()
    tree_reps = 5
    t0 = time.perf_counter()
    for _ in range(tree_reps):
        n_leaves, n_tree_evals = tree_step()
    barrier()
    tree_s = (time.perf_counter() - t0) / tree_reps
    tree_dev_ms = None
    if world == 1:
        ctx.timer_start()
        tr = kd_tree.build_tree(func, params, lo3, hi3, split_depth=TREE_DEPTH, ctx=ctx)
        tree_dev_ms = ctx.timer_stop()
        tr.close()

    # ---- optional: the WHOLE 1920x1080 image once (this rank's share of it for N > 1), to check the sub-sample's rate ----
    full_image = None
    if args.full_image:
        mine_full = sharding.rank_pixels(RES_X, RES_Y, TILE, rank, world, 1)
        rf = torch.from_numpy(roots[mine_full]).to(dev)
        df = torch.from_numpy(dirs[mine_full]).to(dev)
        nf = int(mine_full.shape[0])
        tf_ = torch.zeros(nf, dtype=torch.float32, device=dev)
        hf = torch.zeros(nf, dtype=torch.int32, device=dev)
        cf = torch.zeros(nf, dtype=torch.int32, device=dev)
        barrier()
        ctx.timer_start()
        queries.cast_rays_device((func,), (params,), nf, rf.data_ptr(), df.data_ptr(), tf_.data_ptr(), hf.data_ptr(), cf.data_ptr(),
                                 opts, want_n_evals=False, ctx=ctx)
        ms_full = ctx.timer_stop()
        barrier()
        full_image = {"rays_this_rank": nf, "ms": ms_full, "rays_per_s_this_rank": nf / (ms_full * 1e-3),
                      "ray_steps": int(cf.sum().item()), "hits": int((hf != 0).sum().item())}
        del rf, df, tf_, hf, cf

    if world > 1:
        red = torch.tensor([total_ms, e2e_s, kernel_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        total_ms, e2e_s, kernel_ms = (float(x) for x in red.tolist())
        rs = torch.tensor([ray_steps, n], dtype=torch.int64, device=dev)
        dist.all_reduce(rs)
        ray_steps_all, n_all = (int(x) for x in rs.tolist())
    else:
        ray_steps_all, n_all = ray_steps, n

    if rank == 0:
        value = n_all * args.steps / (total_ms * 1e-3)
        k_s = kernel_ms / max(kernel_launches, 1) * 1e-3                  # this rank's kernel, average launch duration
        algorithmic = flop_per_ray_step * ray_steps / k_s / 1e12         # reference formulation: every column of every layer
        achieved = 2.0 * exec_macs / max(kernel_launches, 1) / k_s / 1e12   # FMAs the kernel actually issued (device counter)
        out = {
            "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(n_all, world, tile_stride),
            "ray_steps_per_s": ray_steps_all * args.steps / (total_ms * 1e-3),
            "e2e": {"value": n_all * args.steps / e2e_s, "unit": "rays/s", "h2d_bytes_per_step": n_all * 24,
                    "d2h_bytes_per_step": n_all * 12, "timer": "wall clock around queries.cast_rays (+ all_gather for N>1)"},
            "gpu_launches": int(gpu_launches), "step_ms": [round(x, 2) for x in step_ms],
            "clocks": clk,
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved / peak_tflops,
                         "traffic": 2.14e6, "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture of the same kernel at 18 tiles (profiles/r1_final_cast_rays256_ncu_full_summary.txt): 2.14 MB read (the weights once), 0 written; not memory-bound",
                         "achieved_algorithmic": algorithmic, "frac_algorithmic": algorithmic / peak_tflops,
                         "executed_over_algorithmic_flops": achieved / algorithmic,
                         "note": "achieved = EXECUTED FP32 FMA flops (device counter: columns that are exactly zero after a relu layer are skipped, "
                                 "exact since fma(0,w,acc)=acc) / kernel time; achieved_algorithmic = the reference formulation's 10*M flop per ray-step / kernel time (can exceed the peak)",
                         "kernel": "k_cast_rays<256>", "peak_source": "measured on this GPU: register-only FFMA kernel "
                         "(niq_measure_fp32_peak); nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.5",
                         "flop_per_ray_step": flop_per_ray_step, "ray_steps_per_launch": ray_steps,
                         "kernel_ms_per_launch": kernel_ms / max(kernel_launches, 1),
                         "hbm_note": "not HBM-bound: 37 B per ray moved for 2.35 GFLOP"},
        }
        full_tree_boxes = 2 ** (TREE_DEPTH + 1) - 1
        out["tree"] = {"metric": "kd-tree boxes/s (construct_uniform_unknown_levelset_tree, same 8x256 ReLU MLP, affine_fixed, split_depth 14, domain [-1,1]^3)",
                       "value": full_tree_boxes / tree_s, "unit": "boxes/s", "boxes": full_tree_boxes, "leaves": n_leaves,
                       "ms": tree_s * 1e3, "scaling": "strong" if world > 1 else None,
                       "timer": "wall clock through the public API incl. the leaf download (e2e); 15 levels, each a classify launch + scan/split",
                       "device_ms": tree_dev_ms,
                       "note": "full tree of the random-init net (no pruning): latency-bound (levels 0-7 hold <= 128 boxes); bunny depth-21 (1.79 M boxes) is in --extra"}
        # ---- CPU baseline beside it (rank 0, N=1 only): the oracle port on the host cores ----
        if world == 1 and not args.no_cpu:
            import multiprocessing as mp
            nw = cpu_workers()
            n_s = 64 * nw
            r, d = sample_rays(roots, dirs, n_s)
            with mp.get_context("fork").Pool(nw) as pool:
                cpu_cast_rays(params, r[:nw], d[:nw], opts, pool, nw)
                dt, st = cpu_cast_rays(params, r, d, opts, pool, nw)
            out["cpu_baseline"] = {"value": r.shape[0] / dt, "unit": "rays/s", "cores": nw, "kind": "port",
                                   "sample": f"{r.shape[0]} rays (whole 16x16 tiles, seed 0) x all 512 steps, {nw} processes x 1 BLAS thread",
                                   "ray_steps_per_s": st / dt}
        if full_image is not None:
            out["full_image"] = full_image
        if args.extra:
            out["extra"] = extra_metrics(ctx)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def extra_metrics(ctx):
    """Secondary numbers on the reference's sample inputs (BASELINE configs[0..3]); each timed once after a warm-up."""
    import implicit_mlp_utils
    import kd_tree
    import queries
    import render
    with np.load(os.path.join(ROOT, "tests", "golden", "mlps.npz")) as d:
        mlps = {nm: {k.split("/", 1)[1]: d[k] for k in d.files if k.startswith(nm + "/")} for nm in ("fox", "bunny", "hammer", "birdcage_occ")}
    lo, hi = np.full(3, -1, np.float32), np.full(3, 1, np.float32)
    ex = {}

    def timed(fn, reps=3):
        fn()
        best = 1e30
        for _ in range(reps):
            t0 = time.perf_counter()
            r = fn()
            best = min(best, time.perf_counter() - t0)
        return best, r

    p = mlps["fox"]
    f = implicit_mlp_utils.generate_implicit_from_params(p, "affine_fixed")
    eye = np.array((2., 1., 2.), np.float32)
    look, up, _ = render.look_at(eye)
    roots, dirs = render.generate_camera_rays(eye, look, up, res=512, fov_deg=30.)
    dt, r = timed(lambda: queries.cast_rays((f,), (p,), roots, dirs, queries.get_default_cast_opts(), ctx=ctx))
    ex["cfg1_fox_512x512_cast_rays"] = {"rays_per_s": roots.shape[0] / dt, "ray_steps_per_s": int(r[2].sum()) / dt, "ms": dt * 1e3,
                                        "hits": int((r[1] > 0).sum()), "tflops_algorithmic": 10 * 7296 * int(r[2].sum()) / dt / 1e12}
    # the frustum variant of the same image (src/queries.py:178-587): frusta of pixels marched together, then split
    look, up, left = render.look_at(eye)
    for res in (512, 1024):
        cam = (eye, look, up, left, 30., 30., res, res)
        dt, r = timed(lambda: queries.cast_rays_frustum((f,), (p,), cam, queries.get_default_cast_opts(), ctx=ctx))
        ex[f"cfg1_fox_{res}x{res}_cast_rays_frustum"] = {"pixels_per_s": res * res / dt, "ms": dt * 1e3, "hits": int((r[1] > 0).sum()),
                                                        "n_evals_reference_count": int(r[3])}
    p = mlps["bunny"]
    f = implicit_mlp_utils.generate_implicit_from_params(p, "affine_fixed")
    for depth in (12, 21):
        st = {}
        dt, r = timed(lambda: kd_tree.construct_uniform_unknown_levelset_tree(f, p, lo, hi, split_depth=depth, stats=st))
        ex[f"cfg2_bunny_tree_depth{depth}"] = {"boxes_per_s": st["n_evals"] / dt, "boxes": st["n_evals"], "leaves": int(r["unknown_node_valid"].sum()),
                                               "near_tie": st["n_near_tie"], "ms": dt * 1e3}
    dt, tri = timed(lambda: kd_tree.hierarchical_marching_cubes(f, p, lo, hi, 7, n_subcell_depth=3), reps=2)
    ex["cfg2_bunny_hmc_depth7_sub3"] = {"triangles": int(tri.shape[0]), "ms": dt * 1e3, "leaves_per_s": 4096 / dt}
    ctx.mc_points(reset=True)
    dt, tri = timed(lambda: kd_tree.hierarchical_marching_cubes(f, p, lo, hi, 9, n_subcell_depth=3), reps=1)
    ev, lat = ctx.mc_points(reset=True)
    ex["cfg2_bunny_hmc_depth9_sub3"] = {"triangles": int(tri.shape[0]), "ms": dt * 1e3, "triangles_per_s": tri.shape[0] / dt,
                                        "leaves": lat // 729 // 2, "leaves_per_s": lat / 729 / 2 / dt,
                                        "lattice_points_evaluated_over_reference": ev / max(lat, 1),
                                        "note": "points on a face shared by two leaves are evaluated once (values and triangles unchanged)"}

    # config 3: hammer x bunny under seeded rigid transforms, affine_truncate (n_keep 64, 'absolute'), eps 1e-3
    import mlp
    pA = mlps["hammer"]
    pB = mlp.prepend_op(mlps["bunny"], mlp.spatial_transformation())
    kw = dict(affine_n_truncate=64, affine_truncate_policy="absolute")
    fA = implicit_mlp_utils.generate_implicit_from_params(pA, "affine_truncate", **kw)
    fB = implicit_mlp_utils.generate_implicit_from_params(pB, "affine_truncate", **kw)
    rng = np.random.default_rng(0)
    n_q, n_found, n_nodes, n_rounds, t_tot, t_max = 24, 0, 0, 0, 0.0, 0.0
    # warm-up on a disjoint pair (the longest kind of query: ~19 rounds), so that no timed query pays first-use allocations
    pB["0000.spatial_transformation.R"] = np.eye(3, dtype=np.float32)
    pB["0000.spatial_transformation.t"] = np.array((1.45, 0., 0.), np.float32)
    kd_tree.find_any_intersection((fA, fB), (pA, pB), lo, hi, 1e-3, ctx=ctx)
    for i in range(n_q + 1):
        th = rng.uniform(0, 2 * np.pi)
        pB["0000.spatial_transformation.R"] = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]], np.float32)
        pB["0000.spatial_transformation.t"] = rng.uniform(-1.5, 1.5, 3).astype(np.float32)
        st = {}
        t0 = time.perf_counter()
        found = kd_tree.find_any_intersection((fA, fB), (pA, pB), lo, hi, 1e-3, stats=st, ctx=ctx)[0]
        d = time.perf_counter() - t0
        if i == 0:
            continue                      # warm-up
        n_found += bool(found); n_nodes += st["n_nodes"]; n_rounds += st["n_rounds"]; t_tot += d; t_max = max(t_max, d)
    ex["cfg3_hammer_x_bunny_intersection_truncate64"] = {"queries": n_q, "found": n_found, "queries_per_s": n_q / t_tot, "nodes_per_s": n_nodes / t_tot,
                                                        "nodes": n_nodes, "rounds": n_rounds, "mean_ms": 1e3 * t_tot / n_q, "max_ms": 1e3 * t_max}

    # config 4: birdcage_occ closest_point, affine_fixed, eps 1e-3.  The reference's default global LIFO window
    # (batch_process_size 2048) couples the queries and is sequential by construction; the window >= stack
    # regime is the same algorithm run level-synchronously per query (shardable).  Both on stated sub-samples of the 1 M queries.
    p = mlps["birdcage_occ"]
    f = implicit_mlp_utils.generate_implicit_from_params(p, "affine_fixed")
    q_all = np.random.default_rng(0).uniform(-1, 1, (1000000, 3)).astype(np.float32)
    for tag, nq, B in (("window2048", 256, 2048), ("window_ge_stack", 4096, 2 ** 26)):
        st = {}
        kd_tree.closest_point(f, p, lo, hi, q_all[:64], eps=1e-3, batch_process_size=B, ctx=ctx)
        t0 = time.perf_counter()
        dist_q, _ = kd_tree.closest_point(f, p, lo, hi, q_all[:nq], eps=1e-3, batch_process_size=B, stats=st, ctx=ctx)
        d = time.perf_counter() - t0
        ex[f"cfg4_birdcage_closest_point_{tag}"] = {"queries": nq, "batch_process_size": B, "queries_per_s": nq / d, "node_visits_per_s": st["n_visits"] / d,
                                                   "visits_per_query": st["n_visits"] / nq, "rounds": st["n_rounds"], "max_stack": st["max_stack"], "ms": d * 1e3,
                                                   "finite": int(np.isfinite(dist_q).sum())}
    return ex


def main():
    if os.environ.get("NIQ_BENCH_WATCHDOG"):          # development: dump every thread's Python stack if the run stalls
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["NIQ_BENCH_WATCHDOG"]), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tiles", type=int, default=74, help="16x16 ray tiles per GPU per step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--extra", action="store_true", help="also time the sample-input configs (secondary numbers)")
    ap.add_argument("--full-image", action="store_true", help="additionally cast the whole 1920x1080 image once (about a minute on one GPU)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
