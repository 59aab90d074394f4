"""The reference arm of bench.py (CPU oracle on the host cores) prints the contract's JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3",
                                   "--nodes", "20000", "--ref-nodes", "20000"], cwd=ROOT, timeout=300).decode().strip().splitlines()[-1]
    d = json.loads(out)
    assert d["impl"] == "reference" and d["metric"].startswith("gossip edge-updates/sec") and d["unit"] == "edge-updates/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 3
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None


def test_other_ranks_exit_quietly_under_torchrun_env():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "3",
                                   "--nodes", "20000", "--ref-nodes", "20000"], cwd=ROOT, env=env, timeout=120).decode().strip()
    assert out == ""


def test_b200_arm_dry_run_against_the_host_build():
    """bench.py cannot run here (no GPU); its Python logic can: tests/bench_dry_run.py stubs the torch.cuda calls and points the
    driver at the host-compiled library, then checks the JSON line against the contract."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "bench_dry_run.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "bench dry run ok" in r.stdout
