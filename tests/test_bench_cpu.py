"""bench.py pieces that run without a GPU: the reference arm (the reference's own CPU `count`
timed through its progress lines), the one-JSON-line contract, the data cache."""
import json
import subprocess
import sys
import time
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_one_json_line(built, tmp_path):
    res = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "tiny", "--steps", "2",
                          "--warmup", "1", "--cache-dir", str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "k-mers/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["gpu_launches"] == 0
    assert d["value"] > 1e5 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and "reads of the workload" in cb["sample"]
    assert d["config"]["workload"] == "tiny"
    # the cache is reused: a second run generates nothing
    res2 = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "tiny", "--steps", "1",
                           "--warmup", "0", "--cache-dir", str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert res2.returncode == 0 and "generated" not in res2.stderr


def test_other_ranks_of_the_reference_arm_do_nothing(tmp_path):
    import os
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    t0 = time.time()
    res = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--cache-dir", str(tmp_path)],
                         capture_output=True, text=True, env=env, timeout=120)
    assert res.returncode == 0 and res.stdout.strip() == "" and time.time() - t0 < 60


def test_gpu_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    res = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--workload", "tiny"], capture_output=True, text=True, timeout=300)
    assert res.returncode != 0 and "no CPU fallback" in (res.stderr + res.stdout)
