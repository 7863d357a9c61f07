#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the range-analysis hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (BASELINE.json configs[4], the one its metric "rays/sec (1080p cast_rays)" is quoted on):
random-init 3->256x8->1 ReLU MLP (glorot-normal A, b~N(0,1e-2^2), NumPy seed 0), 1920x1080 pinhole rays
(eye (2,1,2), look-at origin, fov 30), queries.cast_rays, affine_fixed, default opts.  With this network every
ray takes exactly n_max_step = 512 steps and none hits, so the full image is 1.06e9 ray-steps
= 4.9 EFLOP (a minute per pass on one B200); a "step" of this bench is therefore a STATED, FIXED SUB-SAMPLE of the
image: `--tiles-total` (296) 16x16 pixel tiles spread evenly over it = 75,776 rays x 512 steps, dealt round-robin to
the N ranks -- the total work does not change with N ("strong" scaling).  One JSON line on stdout from rank 0.

value   = rays/s, whole job, ray buffers already resident in HBM, CUDA-event timed on the context stream.
e2e     = rays/s through the public Python API (queries.cast_rays at N = 1, sharding.cast_rays_sharded at N > 1): host
          buffers in and out, H2D + D2H inside the timed region, for N > 1 plus the NCCL all_gather of the device-resident
          (t, hit, count) image (12 B/ray).
roofline= FP32-FMA bound (NOT hbm / tensor: 2.3 GFLOP per 24 B ray; tensor cores would break the 1e-5 parity bar):
          executed and algorithmic flops (10*M per ray-step) / kernel time vs the FFMA peak measured on
          this GPU by a register-only FFMA kernel.
tree    = the second half of BASELINE's metric, kd-tree boxes/s: the config's own depth-14 tree (full, latency-bound, does
          not shard) and bunny split_depth 21 (1.79 M boxes) with the subtrees sharded over the N ranks (strong scaling).
configs = BASELINE configs 1-4 on the reference's sample inputs (N = 1 only), each with its own roofline and a CPU baseline
          timed in the same run.
cpu_baseline / --impl reference = the CPU oracle (NumPy restatement of the reference; JAX is not installable here, so the
          reference itself cannot run) on the host cores over a bounded sample of the same workload.
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

    def __init__(self, device, period=0.2):
        self.device, self.samples, self.stop_flag, self.thread, self.err = device, [], threading.Event(), None, None
        self.period = period

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
            self.stop_flag.wait(self.period)

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
        return {"sm_mhz": float(np.median(mhz)), "sm_min_mhz": float(min(mhz)), "sm_max_mhz": self.smax,
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


TILES_TOTAL = 296          # 4 x 74: at N = 8 every GPU still gets 37 tiles = 4 full waves of 148 CTAs x 16 rays


def chosen_tiles(n_total):
    """The stated sub-sample: n_total tiles spread evenly over the row-major tile grid of the image."""
    import sharding
    ntx, nty = sharding.tile_ids(RES_X, RES_Y, TILE)
    stride = max(1, (ntx * nty) // n_total)
    return np.arange(0, ntx * nty, stride)[:n_total], ntx


def pixels_of_tiles(tiles, ntx):
    ty, tx = np.divmod(np.asarray(tiles, np.int64), ntx)
    oy, ox = np.meshgrid(np.arange(TILE), np.arange(TILE), indexing="ij")
    py = (ty[:, None, None] * TILE + oy[None]).reshape(len(tiles), -1)
    px = (tx[:, None, None] * TILE + ox[None]).reshape(len(tiles), -1)
    ok = (py < RES_Y) & (px < RES_X)
    return (py * RES_X + px)[ok].astype(np.int64)


def workload_config(n_tiles, world):
    tiles, ntx = chosen_tiles(n_tiles)
    return {
        "workload": "BASELINE configs[4]: random-init 3->256x8->1 ReLU MLP (NumPy seed 0), 1920x1080 camera rays, "
                    "queries.cast_rays affine_fixed, default opts (n_max_step 512: every ray runs all 512 steps, no hits)",
        "rays_per_step": int(pixels_of_tiles(tiles, ntx).shape[0]), "ray_steps_per_ray": 512, "image": [RES_X, RES_Y], "tile": TILE,
        "tiles_per_step": int(len(tiles)),
        "sub_sample": "a fixed set of 16x16 tiles spread evenly over the image (every 27th tile of the row-major tile grid), dealt "
                      "round-robin to the ranks; the full image is 2,073,600 rays = 4.9 EFLOP (bench.py --full-image casts it once)",
        "parallelism": f"ray tiles x{world}", "l2": "256 MiB device buffer rewritten between timed steps (untimed)",
    }


def run_reference(args):
    """--impl reference: the CPU oracle port (the reference is pure Python on JAX, which is not installable in
    this image, so there is nothing to build into oracle/_ref) on all host cores.  Same workload / config
    as the GPU arm; each step is a bounded sample of that workload's rays (whole tiles of the same tile set)."""
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
    tiles, ntx = chosen_tiles(args.tiles_total)
    n_sample = int(os.environ.get("NIQ_BENCH_CPU_RAYS", 32 * nw))     # rays per step: a few seconds of CPU work
    idx = pixels_of_tiles(tiles[:max(1, n_sample // (TILE * TILE))], ntx)[:n_sample]
    r, d = roots[idx], dirs[idx]
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
    sample = (f"{n_sample} rays per step (the first whole 16x16 tiles of the GPU arm's tile set) x all 512 steps, rays split over "
              f"{nw} processes x 1 BLAS thread")
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.tiles_total, args.gpus),
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": nw, "kind": "port", "sample": sample,
                         "ray_steps_per_s": steps / total},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------

def dram_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, as extracted from the committed ncu --set full
    capture by tools/ncu_traffic.py (profiles/r2_dram_traffic.json); None when no capture of that kernel is on record."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_dram_traffic.json")) as f:
            rec = json.load(f).get(kernel)
        return rec
    except (OSError, ValueError):
        return None


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
    import kd_tree
    import queries
    import sharding

    ctx = _niq.default_context(local)
    params = synthetic_params()
    func = implicit_mlp_utils.generate_implicit_from_params(params, "affine_fixed")
    M = ctx.mlp(params).macs
    flop_per_ray_step = 10 * M                     # 3 affine rows + 2 point rows, 2 flop per MAC
    roots, dirs = camera_rays()
    opts = queries.get_default_cast_opts()

    # ---- the fixed tile set, dealt round-robin: total work is the same for every N (strong scaling) ----
    tiles, ntx = chosen_tiles(args.tiles_total)
    pix_of = lambda r, w: pixels_of_tiles(tiles[r::w], ntx)
    mine = pix_of(rank, world)
    n = int(mine.shape[0])
    n_all = int(pixels_of_tiles(tiles, ntx).shape[0])
    dev = torch.device("cuda", local)
    r_d, d_d = torch.from_numpy(roots[mine]).to(dev), torch.from_numpy(dirs[mine]).to(dev)
    t_d = torch.zeros(n, dtype=torch.float32, device=dev)
    h_d = torch.zeros(n, dtype=torch.int32, device=dev)
    c_d = torch.zeros(n, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        queries.cast_rays_device((func,), (params,), n, r_d.data_ptr(), d_d.data_ptr(), t_d.data_ptr(), h_d.data_ptr(),
                                 c_d.data_ptr(), opts, want_n_evals=False, ctx=ctx)

    def e2e_step():
        # the call a user makes: host arrays in, host arrays out (N > 1: the image's rays dealt over the ranks, results
        # gathered over NVLink from the device buffers, every rank ends up with the whole result on the host)
        if world == 1:
            return queries.cast_rays((func,), (params,), roots[mine], dirs[mine], opts, ctx=ctx)
        return sharding.cast_rays_sharded((func,), (params,), roots, dirs, opts, RES_X, RES_Y, TILE, pixels_of_rank=pix_of)

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

    # ---- e2e through the public API (host buffers); a few steps are enough for a stable mean ----
    e2e_steps = max(1, min(args.steps, 5))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps

    # ---- kd-tree boxes/s (second half of BASELINE's metric) ----
    lo3, hi3 = np.full(3, -1, np.float32), np.full(3, 1, np.float32)
    tree = {}
    if rank == 0:
        # (a) the config's own tree: depth 14 of the 8x256 net = a FULL tree (nothing pruned, 32,767 boxes), 15 dependent
        # levels whose cost is the latency of one pass through the net: does not shard (replicas only), reported at N = 1 size
        def d14():
            st = {}
            out = kd_tree.construct_uniform_unknown_levelset_tree(func, params, lo3, hi3, split_depth=14, stats=st, ctx=ctx)
            return int(out["unknown_node_valid"].sum()), st["n_evals"]
        d14()
        t0 = time.perf_counter()
        for _ in range(5):
            n_leaves, n_boxes = d14()
        d14_s = (time.perf_counter() - t0) / 5
        ctx.timer_start()
        tr = kd_tree.build_tree(func, params, lo3, hi3, split_depth=14, ctx=ctx)
        d14_dev_ms = ctx.timer_stop()
        tr.close()
        tree["cfg5_depth14"] = {"value": n_boxes / d14_s, "unit": "boxes/s", "boxes": n_boxes, "leaves": n_leaves, "ms": d14_s * 1e3,
                                "device_ms": d14_dev_ms, "boxes_per_s_device": n_boxes / (d14_dev_ms * 1e-3),
                                "tflops_algorithmic_device": 10 * M * n_boxes / (d14_dev_ms * 1e-3) / 1e12,
                                "scaling": "replicas only (15 dependent levels, each one pass through the net: latency-bound, rank 0 alone)",
                                "kernel": "k_tree_persistent<256, TileBox3> (one cooperative launch per build)",
                                "timer": "ms: wall clock through kd_tree.construct_uniform_unknown_levelset_tree incl. the leaf download; device_ms: CUDA events"}
    barrier()
    # (a') the same depth-14 tree with its subtrees sharded (N > 1): one dealt launch per rank (kd_tree.build_tree_dealt), the top
    # 10 levels replicated (<= 1,024 boxes: less than one pass of the grid), levels 10-14 on the rank's own eighth, so that EVERY
    # level is a single pass; leaves all-gathered from the device buffers.  15 dependent levels remain: the floor at any N.
    d14_sh_s, d14_sh_leaves = 0.0, 0
    if world > 1:
        def d14_sharded():
            _, counts = sharding.tree_sharded(func, params, lo3, hi3, 14, top_depth=10, to_host=False, ctx=ctx)
            return int(sum(counts))
        d14_sharded(); d14_sharded()
        barrier()
        t0 = time.perf_counter()
        for _ in range(5):
            d14_sh_leaves = d14_sharded()
        barrier()
        d14_sh_s = (time.perf_counter() - t0) / 5
    # (b) a tree that can scale: bunny.npz (8x64 ELU), split_depth 21 = 1.79 M boxes, 380 K leaves; N > 1: top levels replicated,
    # subtrees dealt round-robin, leaves all-gathered from the device buffers (sharding.tree_sharded)
    bunny = sample_mlp("bunny")
    fb = implicit_mlp_utils.generate_implicit_from_params(bunny, "affine_fixed")
    Mb = ctx.mlp(bunny).macs

    def d21_device():            # until every rank holds all leaves in HBM
        if world == 1:
            tr = kd_tree.build_tree(fb, bunny, lo3, hi3, split_depth=21, ctx=ctx)
            st, nl = tr.stats(), tr.count(0)
            tr.close()
            return nl, st["n_evals"]
        out, counts = sharding.tree_sharded(fb, bunny, lo3, hi3, 21, to_host=False, ctx=ctx)
        return int(sum(counts)), None

    def d21_host():              # the public call: all leaves on the host of every rank
        if world == 1:
            out = kd_tree.construct_uniform_unknown_levelset_tree(fb, bunny, lo3, hi3, split_depth=21, ctx=ctx)
            return int(out["unknown_node_valid"].sum())
        return int(sharding.tree_sharded(fb, bunny, lo3, hi3, 21, ctx=ctx)[0].shape[0])

    d21_device(); d21_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        n_leaves21, n_boxes21 = d21_device()
    barrier()
    d21_s = (time.perf_counter() - t0) / 5
    t0 = time.perf_counter()
    for _ in range(3):
        d21_host()
    barrier()
    d21_host_s = (time.perf_counter() - t0) / 3
    BOXES21 = 1786421 if n_boxes21 is None else n_boxes21      # boxes of the single-device tree (the sharded build classifies the
                                                               # replicated top levels on every rank: not counted twice)
    # where the sharded build's time goes (names the residual of the strong scaling): this rank's own build alone
    # (replicated top + own subtrees, one dealt launch) against the whole call (+ the NCCL gather of the leaves)
    own_s = own_dev_ms = 0.0
    own_levels = own_boxes = 0
    if world > 1:
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            ctx.timer_start()
            tr = sharding._build_own_tree(fb, bunny, lo3, hi3, 21, None, rank, world, {"ctx": ctx})
            own_dev_ms = ctx.timer_stop()
            stt = tr.stats()
            own_levels, own_boxes = stt["n_levels"], stt["n_evals"]
            tr.close()
        own_s = (time.perf_counter() - t0) / 3

    # ---- optional: the WHOLE 1920x1080 image once (this rank's share of it for N > 1) ----
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
            # whole job: the slowest rank's share sets the time; and once through the public sharded call (host rays in, the whole
            # image back on the host of every rank, one NCCL all_gather of the device-resident results)
            mx = torch.tensor([ms_full], dtype=torch.float64, device=dev)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            barrier()
            t0 = time.perf_counter()
            ft, fh, fc, _ = sharding.cast_rays_sharded((func,), (params,), roots, dirs, opts, RES_X, RES_Y, TILE)
            barrier()
            e2e_full = time.perf_counter() - t0
            full_image.update({"ms_max_over_ranks": float(mx.item()), "rays": RES_X * RES_Y,
                               "rays_per_s": RES_X * RES_Y / (float(mx.item()) * 1e-3),
                               "e2e_s": e2e_full, "e2e_rays_per_s": RES_X * RES_Y / e2e_full,
                               "e2e_ray_steps": int(fc.sum()), "e2e_hits": int((fh != 0).sum())})

    if world > 1:
        red = torch.tensor([total_ms, e2e_s, kernel_ms, d21_s, d21_host_s, own_s, own_dev_ms, own_boxes, d14_sh_s], dtype=torch.float64, device=dev)
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        total_ms, e2e_s, kernel_ms, d21_s, d21_host_s, own_s, own_dev_ms, own_boxes_max, d14_sh_s = (float(x) for x in red.tolist())
        rs = torch.tensor([ray_steps, n], dtype=torch.int64, device=dev)
        dist.all_reduce(rs)
        ray_steps_all, n_sum = (int(x) for x in rs.tolist())
        assert n_sum == n_all
    else:
        ray_steps_all = ray_steps

    if rank == 0:
        value = n_all * args.steps / (total_ms * 1e-3)
        k_s = kernel_ms / max(kernel_launches, 1) * 1e-3                  # this rank's kernel, average launch duration
        algorithmic = flop_per_ray_step * ray_steps / k_s / 1e12         # reference formulation: every column of every layer
        achieved = 2.0 * exec_macs / max(kernel_launches, 1) / k_s / 1e12   # FMAs the kernel actually issued (device counter)
        traffic = dram_traffic("k_cast_rays<256>")
        out = {
            "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args.tiles_total, world),
            "ray_steps_per_s": ray_steps_all * args.steps / (total_ms * 1e-3),
            "full_image_extrapolated_s": RES_X * RES_Y / value,
            "e2e": {"value": n_all / e2e_s, "unit": "rays/s", "h2d_bytes_per_step": n_all * 24,
                    "d2h_bytes_per_step": n_all * 12 * (world if world > 1 else 1), "steps": e2e_steps,
                    "timer": "wall clock around queries.cast_rays (N = 1) / sharding.cast_rays_sharded (N > 1: + one NCCL all_gather of "
                             "the device-resident results, every rank reads the whole image back)"},
            "gpu_launches": int(gpu_launches), "step_ms": [round(x, 2) for x in step_ms],
            "clocks": clk,
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved / peak_tflops,
                         "traffic": None if traffic is None else traffic["bytes_per_launch"],
                         "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, read from profiles/r2_dram_traffic.json "
                                         "(tools/ncu_traffic.py over the committed ncu --set full capture); null = no capture on record. "
                                         "The kernel is not memory-bound: the weights are read once per CTA",
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
        tree["bunny_depth21"] = {"value": BOXES21 / d21_s, "unit": "boxes/s", "boxes": BOXES21, "leaves": n_leaves21, "ms": d21_s * 1e3,
                                 "scaling": "strong", "e2e": {"value": BOXES21 / d21_host_s, "unit": "boxes/s", "ms": d21_host_s * 1e3,
                                                              "d2h_bytes": n_leaves21 * 24},
                                 "tflops_algorithmic": 10 * Mb * BOXES21 / d21_s / 1e12,
                                 "parallelism": "single cooperative kernel" if world == 1 else f"top levels replicated, subtrees dealt round-robin x{world}, "
                                                "leaves all-gathered from device buffers (one NCCL all_gather_into_tensor of 24 B/leaf)",
                                 "timer": "wall clock, max over ranks: value = until the leaves are in HBM (of every rank for N > 1); e2e = until every rank "
                                          "holds them on the host (sharding.tree_sharded / kd_tree.construct_uniform_unknown_levelset_tree)"}
        if world > 1:
            tree["cfg5_depth14"]["sharded"] = {
                "value": 32767 / d14_sh_s, "unit": "boxes/s", "ms": d14_sh_s * 1e3, "leaves": d14_sh_leaves, "top_depth": 10, "scaling": "strong",
                "boxes": 32767, "parallelism": f"levels 0-9 replicated, the frontier entering level 10 dealt x{world} inside one persistent launch per rank, "
                                               "leaves all-gathered from device buffers",
                "timer": "wall clock, max over ranks, until the leaves are in HBM of every rank (boxes = the single tree's 32,767)",
                "note": "every level is now one pass of the grid; the 15 dependent passes through the 8 x 256 net are the floor at any N"}
            tree["bunny_depth21"]["residual"] = {
                "own_build_ms": own_s * 1e3, "own_build_device_ms": own_dev_ms, "gather_ms": (d21_s - own_s) * 1e3,
                "levels": own_levels, "boxes_classified_max_rank": int(own_boxes_max), "boxes_single_tree_over_world": BOXES21 / world,
                "note": "own_build = this rank's dealt launch (max over ranks, wall / CUDA events): the 22 dependent levels cost one pass "
                        "through the net each however few boxes a rank holds (the latency floor that does not shrink with N), the top "
                        "levels are replicated, and the slowest rank's share sets the time; gather = count-carrying all_gather of 24 B/leaf "
                        "+ its host-side launch / sync"}
        out["tree"] = {"metric": "kd-tree boxes/s (construct_uniform_unknown_levelset_tree, affine_fixed, domain [-1,1]^3)", **tree}
        # ---- CPU baseline beside it (rank 0, N=1 only): the oracle port on the host cores ----
        if world == 1 and not args.no_cpu:
            import multiprocessing as mp
            nw = cpu_workers()
            idx = pixels_of_tiles(tiles[:max(1, (64 * nw) // (TILE * TILE))], ntx)[:64 * nw]
            r, d = roots[idx], dirs[idx]
            with mp.get_context("fork").Pool(nw) as pool:
                cpu_cast_rays(params, r[:nw], d[:nw], opts, pool, nw)
                dt, st = cpu_cast_rays(params, r, d, opts, pool, nw)
            out["cpu_baseline"] = {"value": r.shape[0] / dt, "unit": "rays/s", "cores": nw, "kind": "port",
                                   "sample": f"{r.shape[0]} rays (the first whole 16x16 tiles of the step's tile set) x all 512 steps, {nw} processes x 1 BLAS thread",
                                   "ray_steps_per_s": st / dt}
        if full_image is not None:
            out["full_image"] = full_image
        if world == 1 and not args.no_configs:
            # the configs are short, latency-bound launches: record the clocks they ran at too -- sparsely (one NVML query per
            # second): at the 200 ms of the headline sampler the persistent closest-point kernel measured 13 % slower
            cc = ClockSampler(local, period=1.0)
            if not os.environ.get("NIQ_BENCH_NO_CLOCKS"):
                cc.start()
            out["configs"] = config_metrics(ctx, peak_tflops, cpu=not args.no_cpu)
            out["configs"]["clocks"] = cc.stop()
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def sample_mlp(name):
    with np.load(os.path.join(ROOT, "tests", "golden", "mlps.npz")) as d:
        return {k.split("/", 1)[1]: d[k] for k in d.files if k.startswith(name + "/")}


def cfg3_transforms(n, seed=0):
    """rotation about z by U[0,2pi), translation U[-1.5,1.5]^3, np.random.default_rng(0)."""
    rng = np.random.default_rng(seed)
    Rs, ts = [], []
    for _ in range(n):
        th = rng.uniform(0, 2 * np.pi)
        Rs.append(np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]], np.float32))
        ts.append(rng.uniform(-1.5, 1.5, 3).astype(np.float32))
    return np.stack(Rs), np.stack(ts)


# ---- CPU legs of the configs block: the oracle on bounded samples, one process each (1 BLAS thread) ----

def _cpu_cfg(task):
    from threadpoolctl import threadpool_limits
    from niq_oracle import mc, net, rays, tree
    lo, hi = np.full(3, -1, np.float32), np.full(3, 1, np.float32)
    kind = task[0]
    with threadpool_limits(limits=1):
        t0 = time.perf_counter()
        if kind == "cfg1":
            _, p, roots, dirs = task
            out = rays.cast_rays((net.AffineContext("affine_fixed"),), (p,), roots, dirs, rays.get_default_cast_opts())
            units = {"rays": int(roots.shape[0]), "ray_steps": int(out[2].sum())}
        elif kind == "cfg2_tree":
            _, p, depth = task
            st = {}
            tree.construct_uniform_unknown_levelset_tree(net.AffineContext("affine_fixed"), p, lo, hi, split_depth=depth, stats=st)
            units = {"boxes": int(st["n_evals"])}
        elif kind == "cfg2_mc":
            _, p, llo, lhi = task
            tri = mc.extract_mesh_from_leaves(p, llo, lhi, 3)
            units = {"leaves": int(llo.shape[0]), "triangles": int(np.asarray(tri).reshape(-1, 9).shape[0])}
        elif kind == "cfg3":
            _, pA, pB, R, t = task
            c = net.AffineContext("affine_truncate", truncate_count=64)
            nodes = 0
            for i in range(R.shape[0]):
                st = {}
                tree.find_any_intersection((c, c), (pA, net.prepend_op(pB, net.spatial_transformation(R[i], t[i]))), lo, hi, 1e-3, stats=st)
                nodes += st["n_nodes"]
            units = {"queries": int(R.shape[0]), "nodes": int(nodes)}
        else:
            _, p, q = task
            st = {}
            tree.closest_point(net.AffineContext("affine_fixed"), p, lo, hi, q, eps=1e-3, batch_process_size=2048, stats=st)
            units = {"queries": int(q.shape[0]), "visits": int(st["n_visits"])}
        return kind, time.perf_counter() - t0, units


def config_metrics(ctx, peak_tflops, cpu=True):
    """BASELINE configs 1-4 on the reference's sample inputs: absolute numbers, fraction of the measured FFMA peak
    (algorithmic flops of SURVEY.md 8(d); executed flops from the device counter where the engine runs), and the CPU oracle
    timed in the same run on a bounded sample (one process per leg, 1 BLAS thread)."""
    import implicit_mlp_utils
    import kd_tree
    import mlp
    import queries
    import render
    mlps = {nm: sample_mlp(nm) for nm in ("fox", "bunny", "hammer", "birdcage_occ")}
    lo, hi = np.full(3, -1, np.float32), np.full(3, 1, np.float32)
    cfg = {}

    def timed(fn, reps=3):
        fn()
        best, dev_best, macs = 1e30, 1e30, 0
        for _ in range(reps):
            ctx.exec_macs(on=True, reset=True)
            ctx.timer_start()
            t0 = time.perf_counter()
            r = fn()
            dt = time.perf_counter() - t0
            dms = ctx.timer_stop()
            macs = ctx.exec_macs(on=False, reset=True)
            if dt < best:
                best, dev_best = dt, dms
        return best, dev_best, macs, r

    def roof(flop_alg, macs, seconds):
        ex = 2.0 * macs / seconds / 1e12 if macs else None
        alg = flop_alg / seconds / 1e12
        return {"bound": "fp32", "unit": "TFLOP/s", "peak": peak_tflops, "achieved_algorithmic": alg, "frac_algorithmic": alg / peak_tflops,
                "achieved": ex, "frac": None if ex is None else ex / peak_tflops}

    # ---- config 1: fox 512x512 cast_rays, affine_fixed ----
    p = mlps["fox"]
    f = implicit_mlp_utils.generate_implicit_from_params(p, "affine_fixed")
    Mf = ctx.mlp(p).macs
    eye = np.array((2., 1., 2.), np.float32)
    look, up, _ = render.look_at(eye)
    roots, dirs = render.generate_camera_rays(eye, look, up, res=512, fov_deg=30.)
    o = queries.get_default_cast_opts()
    dt, dms, macs, r = timed(lambda: queries.cast_rays((f,), (p,), roots, dirs, o, ctx=ctx))
    steps1 = int(r[2].sum())
    cfg["cfg1_fox_512x512_cast_rays"] = {"rays_per_s": roots.shape[0] / dt, "ray_steps_per_s": steps1 / dt, "ms": dt * 1e3, "device_ms": dms,
                                         "hits": int((r[1] > 0).sum()), "ray_steps": steps1, "roofline": roof(10 * Mf * steps1, macs, dms * 1e-3),
                                         "kernel": "k_cast_rays<32>"}
    # ---- config 5's other activation: the same 3->256x8->1 shape with TanH (north_star "ReLU/TanH").  PARITY UNPINNED: the reference has
    # no tanh rule; ours is checked for soundness and against its own oracle (tests/test_tanh.py) ----
    pt = mlp.initialize_params(mlp.build_spec(mlp.quick_mlp_spec(LAYERS, "tanh")), 0)
    ft = implicit_mlp_utils.generate_implicit_from_params(pt, "affine_fixed")
    tl, ntx = chosen_tiles(74)
    pix = pixels_of_tiles(tl, ntx)
    r5, d5 = camera_rays()
    r5, d5 = r5[pix], d5[pix]
    dt, dms, macs, r = timed(lambda: queries.cast_rays((ft,), (pt,), r5, d5, o, ctx=ctx), reps=2)
    steps5 = int(r[2].sum())
    cfg["cfg5_tanh_8x256_cast_rays"] = {"rays_per_s": r5.shape[0] / dt, "ray_steps_per_s": steps5 / dt, "ms": dt * 1e3, "device_ms": dms, "rays": int(r5.shape[0]),
                                        "hits": int((r[1] > 0).sum()), "ray_steps": steps5, "roofline": roof(10 * ctx.mlp(pt).macs * steps5, macs, dms * 1e-3),
                                        "kernel": "k_cast_rays<256>", "parity": "UNPINNED (no tanh rule in the reference; self-written oracle, soundness tests)",
                                        "sample": "74 of the image's 16x16 tiles (every 109th)"}
    # ---- config 2: bunny tree depth 12 / 21, hierarchical marching cubes depth 7 (n_subcell_depth 3) ----
    p = mlps["bunny"]
    f = implicit_mlp_utils.generate_implicit_from_params(p, "affine_fixed")
    Mb = ctx.mlp(p).macs
    for depth in (12, 21):
        st = {}
        dt, dms, macs, r = timed(lambda: kd_tree.construct_uniform_unknown_levelset_tree(f, p, lo, hi, split_depth=depth, stats=st, ctx=ctx))
        cfg[f"cfg2_bunny_tree_depth{depth}"] = {"boxes_per_s": st["n_evals"] / dt, "boxes": st["n_evals"], "leaves": int(r["unknown_node_valid"].sum()),
                                                "near_tie": st["n_near_tie"], "ms": dt * 1e3, "device_ms": dms,
                                                "roofline": roof(10 * Mb * st["n_evals"], macs, dms * 1e-3), "kernel": "k_tree_persistent<64, TileBox3>"}
    ctx.mc_points(reset=True)
    dt, dms, macs, tri = timed(lambda: kd_tree.hierarchical_marching_cubes(f, p, lo, hi, 7, n_subcell_depth=3, ctx=ctx), reps=2)
    ev, lat = ctx.mc_points(reset=True)
    cfg["cfg2_bunny_hmc_depth7_sub3"] = {"triangles": int(tri.shape[0]), "leaves": 4096, "ms": dt * 1e3, "device_ms": dms, "leaves_per_s": 4096 / dt,
                                         "triangles_per_s": tri.shape[0] / dt, "lattice_points_evaluated_over_reference": ev / max(lat, 1),
                                         "roofline": roof(2 * Mb * 729 * 4096 + 10 * Mb * 8191, macs, dms * 1e-3),
                                         "kernel": "k_tree_persistent<64> + k_eval_points<64> + k_mc_count / k_mc_write",
                                         "hbm_pass": "k_mc_write: 36 B per triangle written, lattice values re-read from L2"}
    # ---- config 3: hammer x bunny under 64 seeded rigid transforms, affine_truncate (n_keep 64, 'absolute'), eps 1e-3 ----
    pA = mlps["hammer"]
    pB = mlp.prepend_op(mlps["bunny"], mlp.spatial_transformation())
    kw = dict(affine_n_truncate=64, affine_truncate_policy="absolute")
    fA = implicit_mlp_utils.generate_implicit_from_params(pA, "affine_truncate", **kw)
    fB = implicit_mlp_utils.generate_implicit_from_params(pB, "affine_truncate", **kw)
    R, t = cfg3_transforms(64)

    def one(i, st=None):
        pB["0000.spatial_transformation.R"], pB["0000.spatial_transformation.t"] = R[i], t[i]
        return kd_tree.find_any_intersection((fA, fB), (pA, pB), lo, hi, 1e-3, stats=st, ctx=ctx)[0]
    for i in range(64):               # untimed pass over the whole list: the stream-ordered pool grows to what the largest query needs
        one(i)
    n_found = n_nodes = n_rounds = 0
    lat = []
    for i in range(64):
        st = {}
        t0 = time.perf_counter()
        n_found += bool(one(i, st))
        lat.append(time.perf_counter() - t0)
        n_nodes += st["n_nodes"]; n_rounds += st["n_rounds"]
    tot = float(np.sum(lat))
    FLOP_NODE3 = 2 * 3.80e6 + 14 * 2 * ctx.mlp(pA).macs             # 2 truncate-64 classifies + 14 point evaluations
    cfg["cfg3_hammer_x_bunny_intersection_truncate64"] = {"queries": 64, "found": n_found, "queries_per_s": 64 / tot, "nodes_per_s": n_nodes / tot,
                                                         "nodes": n_nodes, "rounds": n_rounds, "mean_ms": 1e3 * tot / 64, "max_ms": 1e3 * max(lat),
                                                         "median_ms": 1e3 * float(np.median(lat)),
                                                         "roofline": roof(FLOP_NODE3 * n_nodes, 0, tot),
                                                         "kernel": "k_isect_persistent (one cooperative launch per query: the rounds stay on the device)",
                                                         "timer": "wall clock per kd_tree.find_any_intersection call (latency of one query at a time)"}
    # the same 64 queries as ONE call: all of them advance round by round inside one persistent kernel
    stb = {}
    tb, _, _, rb = timed(lambda: kd_tree.find_any_intersection_batch((fA, fB), (pA, pB), lo, hi, 1e-3, R_B=R, t_B=t, stats=stb, ctx=ctx), reps=2)
    nb = int(np.sum(stb["n_nodes"]))
    cfg["cfg3_batched_64_transforms"] = {"queries": 64, "found": int(rb[0].sum()), "queries_per_s": 64 / tb, "nodes_per_s": nb / tb, "nodes": nb,
                                         "ms": 1e3 * tb, "roofline": roof(FLOP_NODE3 * nb, 0, tb),
                                         "kernel": "k_isect_persistent (niq_find_any_intersection_batch: one launch for the whole list)",
                                         "same_verdicts_as_single_queries": bool(int(rb[0].sum()) == n_found)}
    # ---- config 4: birdcage_occ closest_point, affine_fixed, eps 1e-3, Q = 256 of the 1 M seeded queries ----
    p = mlps["birdcage_occ"]
    f = implicit_mlp_utils.generate_implicit_from_params(p, "affine_fixed")
    Mc = ctx.mlp(p).macs
    q_all = np.random.default_rng(0).uniform(-1, 1, (1000000, 3)).astype(np.float32)
    for tag, nq, B in (("window2048", 256, 2048), ("window_ge_stack", 4096, 2 ** 26)):
        st = {}
        dt, dms, macs, _ = timed(lambda: kd_tree.closest_point(f, p, lo, hi, q_all[:nq], eps=1e-3, batch_process_size=B, stats=st, ctx=ctx), reps=1)
        cfg[f"cfg4_birdcage_closest_point_{tag}"] = {"queries": nq, "batch_process_size": B, "queries_per_s": nq / dt, "node_visits_per_s": st["n_visits"] / dt,
                                                    "visits_per_query": st["n_visits"] / nq, "rounds": st["n_rounds"], "max_stack": st["max_stack"], "ms": dt * 1e3,
                                                    "device_ms": dms, "roofline": roof((10 + 14) * Mc * st["n_visits"], macs, dt),
                                                    "kernel": "k_cp_persistent<64> (one cooperative launch for all rounds)" if B <= 2048 else "k_classify_fixed<64> + k_eval_points<64> + k_cp_* (one launch set per round: a few large rounds)"}
    # ---- the CPU oracle beside each config, bounded samples, in parallel processes ----
    if cpu:
        import multiprocessing as mp
        from niq_oracle import net
        sub = (np.arange(512)[::16][:, None] * 512 + np.arange(512)[::16][None, :]).reshape(-1)          # 32 x 32 rays of the 512^2 image
        top = kd_tree.construct_uniform_unknown_levelset_tree(implicit_mlp_utils.generate_implicit_from_params(mlps["bunny"], "affine_fixed"),
                                                              mlps["bunny"], lo, hi, split_depth=12, ctx=ctx)
        v = top["unknown_node_valid"]
        llo, lhi = top["unknown_node_lower"][v][::64], top["unknown_node_upper"][v][::64]                 # 64 of the 4,096 leaves
        tasks = [("cfg1", mlps["fox"], roots[sub], dirs[sub]), ("cfg2_tree", mlps["bunny"], 12), ("cfg2_mc", mlps["bunny"], llo, lhi),
                 ("cfg3", mlps["hammer"], mlps["bunny"], R[:3], t[:3]), ("cfg4", mlps["birdcage_occ"], q_all[:4])]
        with mp.get_context("fork").Pool(min(len(tasks), cpu_workers())) as pool:
            res = {k: (dt, u) for k, dt, u in pool.map(_cpu_cfg, tasks, chunksize=1)}
        def base(kind, key, unit, sample):
            dt, u = res[kind]
            return {"value": u[key] / dt, "unit": unit, "cores": 1, "kind": "port", "sample": sample, "seconds": dt}
        cfg["cfg1_fox_512x512_cast_rays"]["cpu_baseline"] = base("cfg1", "ray_steps", "ray-steps/s", "32 x 32 rays (every 16th pixel of the 512^2 image), oracle cast_rays")
        cfg["cfg2_bunny_tree_depth12"]["cpu_baseline"] = base("cfg2_tree", "boxes", "boxes/s", "the whole depth-12 tree (8,191 boxes), oracle")
        cfg["cfg2_bunny_tree_depth21"]["cpu_baseline"] = dict(cfg["cfg2_bunny_tree_depth12"]["cpu_baseline"], sample="as depth 12 (per-box cost does not depend on depth)")
        cfg["cfg2_bunny_hmc_depth7_sub3"]["cpu_baseline"] = base("cfg2_mc", "leaves", "leaves/s", "64 of the 4,096 leaves (every 64th), oracle extraction with n_subcell_depth 3")
        cfg["cfg3_hammer_x_bunny_intersection_truncate64"]["cpu_baseline"] = base("cfg3", "nodes", "nodes/s", "the first 3 of the 64 transforms, oracle find_any_intersection")
        c4 = base("cfg4", "visits", "node visits/s", "the first 4 queries at batch_process_size 2048, oracle closest_point")
        cfg["cfg4_birdcage_closest_point_window2048"]["cpu_baseline"] = c4
        cfg["cfg4_birdcage_closest_point_window_ge_stack"]["cpu_baseline"] = dict(c4, sample=c4["sample"] + " (per-visit cost is the same)")
    return cfg


def main():
    if os.environ.get("NIQ_BENCH_WATCHDOG"):          # development: dump every thread's Python stack if the run stalls
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["NIQ_BENCH_WATCHDOG"]), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tiles-total", type=int, default=TILES_TOTAL, help="16x16 ray tiles per step, whole job (fixed for every N)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs block (BASELINE configs 1-4)")
    ap.add_argument("--full-image", action="store_true", help="additionally cast the whole 1920x1080 image once (about a minute on one GPU)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
