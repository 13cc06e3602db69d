#!/usr/bin/env python
"""Run bench.py over the secondary workloads of BASELINE.json:configs (HiFi reads, k sweep,
sparse dictionary, the CPU-runnable config 1) and collect the JSON lines.
usage: tools/sweep.py out.jsonl [workload ...]"""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
out = Path(sys.argv[1])
names = sys.argv[2:] or ["config1", "hifi", "k20", "k25", "k31", "sparse10"]
with open(out, "w") as f:
    for name in names:
        res = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--workload", name, "--steps", "5", "--warmup", "3", "--no-cpu"],
                             capture_output=True, text=True)
        line = res.stdout.strip().splitlines()[-1] if res.stdout.strip() else '{"workload": "%s", "error": %r}' % (name, res.stderr[-400:])
        f.write(line + "\n")
        f.flush()
        print(name, "rc", res.returncode, file=sys.stderr)
