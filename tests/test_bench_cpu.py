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
