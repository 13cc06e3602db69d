#!/usr/bin/env python
"""Roofline denominators on the box: random-sector gather GB/s (qk_bench_gather) for sector
sizes 32/64/128 B, loads in flight 1..8, table sizes 1..32 GiB; and pinned H2D GB/s.  One JSON object per line on stdout.  (Measurement tooling.)"""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))

if len(sys.argv) > 1 and sys.argv[1] == "child":
    from conftest import load_package
    qk = load_package()
    sizes = [int(x) for x in sys.argv[2].split(",")]
    with qk.Context(n_slots=1, chunk_capacity=64 << 20) as ctx:
        for gib in sizes:
            for gran in (32, 64, 128):
                for mlp in (1, 2, 4, 8):
                    g = ctx.bench_gather(gib << 30, gran=gran, loads_in_flight=mlp, n_gathers=1 << 29)
                    print(json.dumps({"table_GiB": gib, "gran": gran,
                                      "in_flight": mlp, "GBs": round(g, 1), "G_per_s": round(g / gran, 2)}), flush=True)
        print(json.dumps({"h2d_GBs": round(ctx.bench_h2d(64 << 20, 16), 2)}), flush=True)
else:
    sizes = sys.argv[1] if len(sys.argv) > 1 else "1,32"
    subprocess.run([sys.executable, __file__, "child", sizes], check=True)
