"""bench.py --impl reference (the CPU arm: the GPU arm's batched workload restated in NumPy on the host cores, plus the
reference-semantics serial port and -- when its checkout is present -- the unmodified reference itself) prints exactly
ONE JSON line with the contract's keys and the SAME `config` as the GPU arm; runs without a GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    env = dict(os.environ, DMFG_BENCH_CPU_EPISODES="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--gpus", "1"], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "population-steps/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("population-steps/sec") and d["value"] > 0 and d["n_gpus"] == 1
    assert d["steps"] == 1 and d["warmup"] == 0 and d["scaling"] == "weak" and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "episodes" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.bench_config(1 << 20, 1)          # the two arms are quoted on one config
    assert cb["reference_semantics_port"]["value"] > 0
    ref = cb["reference_itself"]
    if os.path.isfile("/root/reference/mfg_ac2.py"):
        assert ref["kind"] == "reference" and ref["value"] > 0
    else:
        assert ref is None


def test_reference_arm_other_ranks_print_nothing():
    env = dict(os.environ, DMFG_BENCH_CPU_EPISODES="2", RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--gpus", "2"], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.gpu
def test_gpu_arm_prints_one_json_line_with_the_contract_keys():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--log2-pops", "13", "--steps", "3",
                          "--warmup", "3", "--no-cpu-baseline", "--skip-big-modes"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 3 and d["data"] == "synthetic"
    assert d["gpu_launches"] == 9 and "workload" in d["config"] and "l2" in d["config"]     # counted, 3 per step
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == (1 << 13) * 15 * 4 and e["d2h_bytes_per_step"] == (136 + 2) * 8
    r = d["roofline"]
    assert r["bound"] == "issue" and r["peak"] == 4.0 and r["hbm"]["unit"] == "GB/s"
    assert abs(r["hbm"]["frac"] - r["hbm"]["achieved"] / r["hbm"]["peak"]) < 1e-12
    if r["frac"] is not None:                                    # the instruction count matches the compiled sources
        assert abs(r["frac"] - r["achieved"] / 4.0) < 1e-12 and r["traffic"] is not None
    else:
        assert "STALE" in r["inst_count_source"]
    assert d["clocks"]["samples"] >= 0 and isinstance(d["clocks"]["reasons"], list)
    assert d["modes"]["irl_update"]["value"] > 0
