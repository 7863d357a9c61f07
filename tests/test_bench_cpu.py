"""CPU checks of bench.py's host logic: the synthetic MLP and camera of the product arm and of the
reference (oracle) arm are the same bits; the reference arm prints the contract's JSON line."""
import json
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT


def test_synthetic_workload_same_bits_on_both_arms():
    sys.path.insert(0, ROOT)
    import bench
    from niq_oracle import net, rays
    a = bench.synthetic_params()
    b = net.random_mlp(bench.LAYERS, "relu", seed=0)
    assert sorted(a) == sorted(b)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k])
    r, d = bench.camera_rays()
    eye = np.array((2., 1., 2.), np.float32)
    look, up, _ = rays.look_at(eye)
    orr, od = rays.generate_camera_rays(eye, look, up, res=bench.RES_X, fov_deg=30., res_y=bench.RES_Y)
    np.testing.assert_array_equal(r, orr)
    np.testing.assert_array_equal(d, od)
    assert r.shape == (1920 * 1080, 3)


def test_reference_arm_json_line():
    env = dict(os.environ, NIQ_BENCH_CPU_RAYS="8")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "rays/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_fixed_tile_set_is_the_same_for_every_rank_count():
    """Strong scaling: the step's tile set does not depend on N, the ranks' shares partition it, and both arms describe the
    same workload (`config` identical)."""
    sys.path.insert(0, ROOT)
    import bench
    tiles, ntx = bench.chosen_tiles(bench.TILES_TOTAL)
    assert len(tiles) == bench.TILES_TOTAL == 296 and len(set(tiles.tolist())) == 296
    allpix = bench.pixels_of_tiles(tiles, ntx)
    assert allpix.shape[0] == len(set(allpix.tolist())) and allpix.max() < bench.RES_X * bench.RES_Y
    for world in (1, 2, 4, 8):
        parts = [bench.pixels_of_tiles(tiles[r::world], ntx) for r in range(world)]
        assert sum(p.shape[0] for p in parts) == allpix.shape[0]
        assert sorted(np.concatenate(parts).tolist()) == sorted(allpix.tolist())
    cfg = bench.workload_config(bench.TILES_TOTAL, 1)
    assert cfg["rays_per_step"] == allpix.shape[0] and cfg["tiles_per_step"] == 296


def test_config_cpu_legs_run_on_small_samples():
    """The CPU (oracle) legs of bench.py's configs block, on tiny samples: each returns (kind, seconds, units)."""
    sys.path.insert(0, ROOT)
    import bench
    from conftest import sample_params
    fox = sample_params("fox")
    r, d = bench.camera_rays()
    k, dt, u = bench._cpu_cfg(("cfg1", fox, r[::200000], d[::200000]))
    assert k == "cfg1" and dt > 0 and u["rays"] == r[::200000].shape[0] and u["ray_steps"] >= u["rays"]
    k, dt, u = bench._cpu_cfg(("cfg2_tree", fox, 4))
    assert u["boxes"] >= 5
    k, dt, u = bench._cpu_cfg(("cfg4", fox, np.zeros((2, 3), np.float32)))
    assert u["queries"] == 2 and u["visits"] > 0
    R, t = bench.cfg3_transforms(3)
    assert R.shape == (3, 3, 3) and t.shape == (3, 3) and np.allclose(np.linalg.det(R), 1.0, atol=1e-5)
