#!/usr/bin/env python
"""Aggregate pinned H2D bandwidth of the box with all ranks copying at once (torchrun), and the
NUMA picture the box exposes.  The e2e legs of bench.py are bound by this number at N > 1."""
import glob
import json
import os
import subprocess
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
from conftest import load_package  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
qk = load_package()
out = {}
if rank == 0:
    out["numa_nodes"] = sorted(os.path.basename(p) for p in glob.glob("/sys/devices/system/node/node*"))
    out["gpu_numa"] = {}
    for p in glob.glob("/sys/bus/pci/devices/*/numa_node"):
        try:
            if (Path(p).parent / "vendor").read_text().strip() == "0x10de":
                out["gpu_numa"][Path(p).parent.name] = Path(p).read_text().strip()
        except OSError:
            pass
    out["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[-3000:]
    out["cpus"] = len(os.sched_getaffinity(0))
with qk.Context(device=local, n_slots=2, chunk_capacity=64 << 20) as ctx:
    alone = None
    if world > 1:
        for r in range(world):                      # one rank at a time
            dist.barrier()
            if r == rank:
                mine_alone = ctx.bench_h2d(64 << 20, 16)
        dist.barrier()
    together = ctx.bench_h2d(64 << 20, 64)         # everybody at once
    t = torch.tensor([together, mine_alone if world > 1 else together], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        g = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(g, t)
        vals = [x.tolist() for x in g]
    else:
        vals = [t.tolist()]
if rank == 0:
    out["h2d_together_gbs"] = [round(v[0], 1) for v in vals]
    out["h2d_alone_gbs"] = [round(v[1], 1) for v in vals]
    out["h2d_together_sum_gbs"] = round(sum(v[0] for v in vals), 1)
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
