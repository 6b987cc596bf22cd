"""bench.py's reference arm (CPU only: the oracle on the host cores) prints exactly one JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("configs[1]")


def test_reference_arm_on_another_rank_prints_nothing():
    """Under torchrun (N > 1) rank 0 alone runs the reference arm; the other ranks exit 0 without work or output."""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29599")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    assert p.stdout.strip() == ""


def test_committed_bench_lines_follow_the_contract():
    """The bench lines committed under profiles/ (measured on B200s, this round) carry every key of the contract and are
    internally consistent: value = frames per step / time per step, roofline.achieved = algorithmic bytes / launch time,
    frac = achieved / peak, parity checked in the run."""
    import glob

    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_bench_c*_*gpu.json")))
    assert len(files) >= 6
    for f in files:
        d = json.load(open(f))
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                    "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "parity_in_run"):
            assert key in d, (f, key)
        assert d["unit"] == "frames/s" and d["scaling"] == "weak" and d["dtype"] == "f32" and d["vs_baseline"] is None
        assert d["steps"] == 20 and d["warmup"] >= 3 and d["gpu_launches"] > 0
        frames_per_step = d["config"]["frames_per_step"]
        assert frames_per_step == d["config"]["streams_per_gpu"] * d["n_gpus"]
        assert abs(d["value"] - frames_per_step / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"], f
        r = d["roofline"]
        assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["avg_launch_ms"] * 1e-3) / 1e9) <= 1e-6 * r["achieved"], f
        assert r["avg_launch_ms"] <= d["ms_per_step"]
        e = d["e2e"]
        assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
        assert not set(d["clocks"].get("reasons", [])) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}, f
        p = d["parity_in_run"]
        assert p["ok"] is True and p["streams"] >= 1 and p["frames"] >= 20 and p["max_rad"] <= 1e-3 and p["max_m"] <= 1e-3
        if "dense" not in d["config"]["workload"]:
            assert p["max_rad"] <= 1e-4 and p["max_m"] <= 1e-4, f  # the reference's own candidate modes: the strict bar
        if d["n_gpus"] == 1 and "cpu_baseline" in d:
            assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1
