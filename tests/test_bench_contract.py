"""bench.py contract, CPU side: the reference arm runs without a GPU and prints ONE JSON line with the agreed keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-sample", "4"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "MP/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("megapixels/sec") and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["vs_baseline"] is None
    # the reference arm states the SAME config object as the GPU arm (the driver compares them)
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.build_config("cfg2", 1) and d["scaling"] == "weak"


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_committed_gpu_bench_lines_keep_the_contract():
    """Every committed GPU bench line under profiles/ (written by bench.py on a B200) carries the keys of the contract:
    the base line, `e2e` with its byte counts, `gpu_launches`, `clocks`, `roofline` (bound / achieved / peak / frac /
    traffic) and - at one GPU - `cpu_baseline`; frac = achieved / peak and value = pixels / time are self-consistent."""
    import glob
    paths = sorted(glob.glob(os.path.join(ROOT, "profiles", "r01_bench_cfg*.json")) +
                   glob.glob(os.path.join(ROOT, "profiles", "r02_bench_cfg*.json")))
    assert len(paths) >= 12
    for path in paths:
        d = json.loads(open(path).read().strip().splitlines()[-1])
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
            assert k in d, (path, k)
        r02 = os.path.basename(path).startswith("r02")
        strong = r02 and ("cfg3" in path or "cfg4" in path)          # round 2: cfg3 / cfg4 split ONE batch over the ranks
        assert d["unit"] == "MP/s" and d["scaling"] == ("strong" if strong else "weak") and d["data"] == "synthetic"
        assert d["vs_baseline"] is None
        if r02 and d["n_gpus"] == 1 and "cfg2" in path:
            assert d["sustained"]["seconds"] >= 3.0 and d["sustained"]["clocks"]["sm_mhz"] > 0
        assert d["warmup"] >= 3 and d["gpu_launches"] > 0 and "workload" in d["config"] and "l2" in d["config"]
        assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(d["e2e"])
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
        assert d["e2e"]["value"] != d["value"]                                    # measured separately, not copied
        r = d["roofline"]
        assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert 0.3 < r["frac"] < 1.0
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        h, w = d["config"]["image"]
        mp = d["config"]["images_per_step"] * h * w / 1e6
        assert abs(d["value"] - mp / (d["ms_per_step"] / 1e3)) <= 1e-6 * d["value"]
        if d["n_gpus"] == 1 and ("cpu_baseline" in d or not r02):    # some r02 lines were taken with --no-cpu-baseline
            cb = d["cpu_baseline"]
            assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]
