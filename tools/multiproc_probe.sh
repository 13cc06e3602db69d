#!/bin/bash
# Does a count kernel slow down just because another process drives another GPU of the box?
# Two independent single-GPU benches (no torch.distributed, no NCCL) side by side vs one alone.
cd "$(dirname "$0")/.."
python bench.py --kernel-only --steps 20 --warmup 3 2>/dev/null > gpurun_out/probe_alone.json
CUDA_VISIBLE_DEVICES=0 python bench.py --kernel-only --steps 60 --warmup 3 --cache-dir /dev/shm/qk_bench_cache 2>/dev/null > gpurun_out/probe_a.json &
CUDA_VISIBLE_DEVICES=1 python bench.py --kernel-only --steps 60 --warmup 3 --cache-dir /dev/shm/qk_bench_cache 2>/dev/null > gpurun_out/probe_b.json &
wait
python - <<'PY'
import json
for n in ("alone", "a", "b"):
    d = json.load(open(f"gpurun_out/probe_{n}.json"))
    print(n, round(d["value"] / 1e9, 1), "G k-mers/s", round(d["roofline"]["avg_launch_ms"], 4), "ms/launch")
PY
