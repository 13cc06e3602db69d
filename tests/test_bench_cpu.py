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


def _bench():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", ROOT / "bench.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_chunk_layout_packs_whole_records_at_aligned_offsets():
    """bench.py cuts the HBM-resident reads of the variable-length (HiFi) workloads into chunks of whole records;
    qk_submit_device wants 16-byte-aligned chunks no larger than a slot."""
    import numpy as np
    b = _bench()
    rng = np.random.default_rng(3)
    per = rng.integers(1001, 100000, size=5000).astype(np.uint64)
    cap = 1 << 20
    offsets, c_offs, c_sizes = b.chunk_layout(per, cap)
    assert offsets.size == per.size and len(c_offs) == len(c_sizes) > 100
    assert all(o % 256 == 0 for o in c_offs) and all(0 < s <= cap for s in c_sizes)
    assert sum(c_sizes) == int(per.sum())
    ends = offsets + per
    assert np.all(offsets[1:] >= ends[:-1])                       # records in order, never overlapping
    starts = set(int(o) for o in c_offs)
    gaps = np.flatnonzero(offsets[1:] != ends[:-1]) + 1           # a gap only where a new chunk starts
    assert all(int(offsets[i]) in starts for i in gaps)
    for o, s in zip(c_offs, c_sizes):                             # every chunk is exactly a run of records
        i = int(np.searchsorted(offsets, o))
        assert int(offsets[i]) == o
        j = int(np.searchsorted(ends, o + s))
        assert int(ends[j]) == o + s


def test_hifi_lengths_do_not_depend_on_how_the_stream_is_cut():
    import numpy as np
    b = _bench()
    w = b.SYNTH_WORKLOADS["config4"]
    whole = b.synth_lens(w, 42, 0, 3 << 20)
    assert whole.min() >= 1000 and whole.max() <= 99998 and (whole > 65536).any()
    assert 14000 < np.median(whole) < 16000
    for first, n in ((0, 10), (12345, 999), ((1 << 20) - 7, 50), ((2 << 20) + 3, 1 << 19)):
        assert np.array_equal(b.synth_lens(w, 42, first, n), whole[first:first + n])
    assert b.synth_lens(b.SYNTH_WORKLOADS["config3"], 42, 0, 10) is None
    assert not np.array_equal(b.synth_lens(w, 43, 0, 100), whole[:100])
