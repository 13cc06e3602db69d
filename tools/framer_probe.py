#!/usr/bin/env python
"""Host-side measurements for the end-to-end roofline: memory bandwidth of the box (read, memcpy) and the
multi-threaded framer alone (raw GB/s in, framed GB/s out) by thread count, with and without the staged
non-temporal stores, packed chunks or text.  No GPU needed.  usage: tools/framer_probe.py [reads=2000000]"""
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
from conftest import load_package  # noqa: E402

qk = load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
rng = np.random.default_rng(1)
rec = np.empty((n, 317), dtype=np.uint8)
rec[:, 0] = ord("@"); rec[:, 1] = ord("r"); rec[:, 2:12] = ord("0"); rec[:, 12] = 10
rec[:, 13:163] = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=(n, 150))
rec[:, 163] = 10; rec[:, 164] = ord("+"); rec[:, 165] = 10; rec[:, 166:316] = ord("I"); rec[:, 316] = 10
data = np.ascontiguousarray(rec).reshape(-1)
cpus = len(os.sched_getaffinity(0))
out = {"cpus": cpus, "raw_bytes": int(data.size), "host_memory": {}, "framer": []}
for t in sorted({1, 4, 8, cpus}):
    if t <= cpus:
        r, c = qk.bench_host_memory(256 << 20, t)
        out["host_memory"][t] = {"read_gbs": r, "memcpy_gbs": c}
for packed, nt in (("1", "1"), ("0", "1"), ("0", "0")):       # packed chunks (the default); text; text without the staged non-temporal stores
    os.environ["QK_PACKED"] = packed
    os.environ["QK_FRAMER_NT"] = nt
    for t in sorted({1, 4, 8, 12, cpus}):
        if t <= cpus:
            a, b = qk.bench_framer(data.ctypes.data, data.size, threads=t, repeats=3)
            out["framer"].append({"packed": packed == "1", "nt_stores": nt == "1", "threads": t, "raw_gbs": a, "out_gbs": b})   # out: bytes of text, or positions (0.375 bytes each) when packed
print(json.dumps(out))
