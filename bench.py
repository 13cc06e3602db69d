#!/usr/bin/env python
"""bench.py -- k-mers/s of the `quicKmer2 count` hot path on B200 (and the reference's CPU
path beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

One "step" = one whole counting job over the workload's read set: zero the depth counters,
run codec + probe + increment over every framed chunk, (N > 1: NCCL-reduce the counters to
rank 0).  Workloads follow BASELINE.json:configs:

  config2  64 Mb reference with 200 x 20 kb segmental duplications, k=30 dictionary
           (58 M k-mers, 1 GiB device table), 30x = 12.8 M x 150 bp FASTQ reads -- the
           default, the configuration the metric is quoted on for one GPU
  config1  1 Mb reference, 1 M x 150 bp FASTA reads (the reference's CPU-runnable case)
  tiny     smoke-sized (CI)

Printed keys (one JSON line, rank 0):
  value        k-mers/s, framed chunks already resident in HBM (device-event span, max over ranks)
  e2e          k-mers/s through the public call (raw FASTA/FASTQ bytes in pinned host memory ->
               H2D of whole-line pieces -> device framing + count kernels -> D2H of the uint16 depths)
  e2e_preframed  same but from pre-framed pinned chunks: the H2D-overlap pipeline alone
  roofline     dominant kernel (qk_count_kernel): algorithmic bytes / mean launch time
  cpu_baseline the reference's own `count -t T` (oracle/_ref/quicKmer2) on a bounded sample
               of the same workload, on this box's host cores

--impl reference times that CPU path alone (the reference arm).  Data are synthetic and
seeded (quick-mer2_b200/bin/qk_synth); they are cached under --cache-dir so the two arms,
run back to back on one box, share them.  Nothing here reads /root/reference.
"""
import argparse
import importlib.util
import json
import os
import re
import shutil
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
PKG_DIR = ROOT / "quick-mer2_b200"
SYNTH = PKG_DIR / "bin" / "qk_synth"
REF_BIN = ROOT / "oracle" / "_ref" / "quicKmer2"
PORT_BIN = ROOT / "oracle" / "_build" / "qk_oracle"

WORKLOADS = {
    # name: ref args, dict args, reads args (per rank), cpu sample reads, description
    "config2": dict(
        ref=["--bases", 64000000, "--contigs", 4, "--seed", 2024, "--segdups", 200, "--segdup-len", 20000,
             "--divergence-ppm", 10000, "--nblock", 50000],
        dict=["--k", 30, "--slots", "128M", "--ctrl-block", 100000],
        reads=["--n", 12800000, "--len", 150, "--err-ppm", 2000, "--fastq"], fastq=True, seed=42,
        sample_reads=12800000,   # the CPU leg counts the whole per-GPU read set (~7 s): its fixed sleep(1) must not dominate
        desc="64 Mb ref + 200x20kb segdups, k=30 (58.2M k-mers, 128Mi-slot .qm), 30x = 12.8M x 150bp FASTQ"),
    "config1": dict(
        ref=["--bases", 1000000, "--contigs", 1, "--seed", 1],
        dict=["--k", 30, "--slots", "4M", "--ctrl-block", 10000],
        reads=["--n", 1000000, "--len", 150, "--err-ppm", 2000], fastq=False, seed=42,
        sample_reads=1000000,
        desc="1 Mb ref, k=30 (4Mi-slot .qm), 1M x 150bp FASTA"),
    # config 4 shape: HiFi-like reads (log-normal, median 15 kb, up to the 99,998-base line limit,
    # some past the 65,536 run-counter wrap) against the config-2 dictionary, FASTA as the
    # documented samtools|awk pipe produces it
    "hifi": dict(
        ref=["--bases", 64000000, "--contigs", 4, "--seed", 2024, "--segdups", 200, "--segdup-len", 20000,
             "--divergence-ppm", 10000, "--nblock", 50000],
        dict=["--k", 30, "--slots", "128M", "--ctrl-block", 100000],
        reads=["--n", 128000, "--len", 15000, "--hifi", "--max-len", 99998, "--min-len", 1000, "--err-ppm", 1000],
        fastq=False, seed=42, sample_reads=12800,
        desc="config-2 dictionary, 30x HiFi-like reads (median 15 kb, max 99,998), FASTA"),
    # config 5: k sweep and a sparse (1/10) dictionary that sits near the L2 size
    **{f"k{k}": dict(
        ref=["--bases", 64000000, "--contigs", 4, "--seed", 2024, "--segdups", 200, "--segdup-len", 20000,
             "--divergence-ppm", 10000, "--nblock", 50000],
        dict=["--k", k, "--slots", "128M", "--ctrl-block", 100000],
        reads=["--n", 12800000, "--len", 150, "--err-ppm", 2000, "--fastq"], fastq=True, seed=42,
        sample_reads=1600000, desc=f"config-2 reference and reads, k={k}") for k in (20, 25, 31)},
    "sparse10": dict(
        ref=["--bases", 64000000, "--contigs", 4, "--seed", 2024, "--segdups", 200, "--segdup-len", 20000,
             "--divergence-ppm", 10000, "--nblock", 50000],
        dict=["--k", 30, "--slots", "128M", "--ctrl-block", 100000], sparse=10,
        reads=["--n", 12800000, "--len", 150, "--err-ppm", 2000, "--fastq"], fastq=True, seed=42,
        sample_reads=1600000, desc="config-2 reference and reads, dictionary thinned 1/10 by the reference's `sparse` (5.8M k-mers)"),
    "tiny": dict(
        ref=["--bases", 300000, "--contigs", 2, "--seed", 5],
        dict=["--k", 30, "--ctrl-block", 10000],
        reads=["--n", 100000, "--len", 150, "--err-ppm", 2000, "--fastq"], fastq=True, seed=42,
        sample_reads=100000,
        desc="300 kb ref, k=30, 100k x 150bp FASTQ"),
}


def load_package():
    spec = importlib.util.spec_from_file_location("quickmer2_b200", PKG_DIR / "__init__.py",
                                                  submodule_search_locations=[str(PKG_DIR)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["quickmer2_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------- data ---------
def synth(*args):
    res = subprocess.run([str(SYNTH), *map(str, args)], capture_output=True, text=True)
    if res.returncode:
        raise RuntimeError(f"qk_synth {args}: {res.stderr}")
    return res.stdout


def cache_dir(arg):
    if arg:
        d = Path(arg)
    else:
        base = Path("/dev/shm") if Path("/dev/shm").is_dir() and shutil.disk_usage("/dev/shm").free > 48 << 30 else Path("/tmp")
        d = base / "qk_bench_cache"
    d.mkdir(parents=True, exist_ok=True)
    return d


def ensure(path: Path, make):
    """Create `path` once (atomic rename), also when several ranks race for it."""
    if path.exists():
        return
    lock = path.with_suffix(path.suffix + ".lock")
    try:
        fd = os.open(lock, os.O_CREAT | os.O_EXCL | os.O_WRONLY)
    except FileExistsError:
        while not path.exists():
            time.sleep(0.5)
            try:
                stale = time.time() - lock.stat().st_mtime > 900   # a generator that died: take over
            except FileNotFoundError:
                stale = False
            if stale:
                lock.unlink(missing_ok=True)
            if not lock.exists() and not path.exists():
                return ensure(path, make)
        return
    try:
        os.close(fd)
        t0 = time.time()
        make()
        log(f"generated {path.name} in {time.time() - t0:.1f} s")
    finally:
        lock.unlink(missing_ok=True)


def prepare(name, cdir, rank_seed_offset=0, sample=False):
    """Reference, dictionary and reads of a workload; returns paths."""
    w = WORKLOADS[name]
    d = cdir / name
    d.mkdir(parents=True, exist_ok=True)
    ref = d / "ref.fa"

    def make_dict():
        tmp = d / "ref.tmp.fa"
        synth("ref", "--out", tmp, *w["ref"])
        os.replace(tmp, ref)
        synth("dict", "--ref", ref, "--out", d / "tmpdict", "--threads", min(16, os.cpu_count() or 1), *w["dict"])
        os.replace(d / "tmpdict.qgc", d / "ref.fa.qgc")
        if w.get("sparse"):          # thin with the reference's own `sparse` (Q.c:1306-1483): writes ref.fa.rqm
            os.replace(d / "tmpdict.qm", d / "full.fa.qm")
            os.symlink(ref, d / "full.fa")
            res = subprocess.run([str(REF_BIN), "sparse", str(w["sparse"]), "full.fa"], cwd=d, capture_output=True, text=True)
            if res.returncode or not (d / "full.fa.rqm").exists():
                raise RuntimeError("reference sparse failed: " + res.stdout[-300:])
            (d / "full.fa.qm").unlink()
            (d / "ref.fa.qgc").unlink()   # ordinals changed; the bench does not use it
            os.replace(d / "full.fa.rqm", d / "ref.fa.qm")
        else:
            os.replace(d / "tmpdict.qm", d / "ref.fa.qm")
    ensure(d / "ref.fa.qm", make_dict)
    ext = "fq" if w["fastq"] else "fa"
    seed = w["seed"] + rank_seed_offset
    reads_args = list(w["reads"])
    if sample and w["sample_reads"] < reads_args[reads_args.index("--n") + 1]:
        reads_args[reads_args.index("--n") + 1] = w["sample_reads"]
        reads = d / f"sample_{w['sample_reads']}_s{seed}.{ext}"
    else:
        reads = d / f"reads_s{seed}.{ext}"      # (a sample as large as the workload is the workload's own file)

    def make_reads():
        tmp = reads.with_suffix(".tmp")
        synth("reads", "--ref", ref, "--out", tmp, "--seed", seed, *reads_args)
        os.replace(tmp, reads)
    ensure(reads, make_reads)
    return d, ref, reads


# --------------------------------------------------------------------------- clocks -------
class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "25"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.proc.wait()
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 7] or [r for _, r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i] == "Active" for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons, "samples": len(rows),
                "power_w_max": max(float(r[2]) for r in rows)}


# --------------------------------------------------------------------------- CPU arm ------
def run_reference_count(ref_prefix: Path, reads: Path, threads: int, out_prefix: Path):
    """Run the compiled reference `count -t threads` on a pty (line-buffered stdout), time the
    counting phase from its own progress lines -- "Read 0x.. hash" (Q.c:359, dictionary loaded)
    to "Counting elapse" (Q.c:481) -- and stop it there: the serial chain-walk dump that
    follows (Q.c:490-518) is not part of the metric.  Returns (k-mers, seconds)."""
    import pty
    master, slave = pty.openpty()
    cmd = [str(REF_BIN), "count"] + (["-t", str(threads)] if threads else []) + [str(ref_prefix), str(reads), str(out_prefix)]
    proc = subprocess.Popen(cmd, stdout=slave, stderr=subprocess.DEVNULL, stdin=subprocess.DEVNULL)
    os.close(slave)
    buf, t_loaded, t_done, total = b"", None, None, None
    while True:
        try:
            data = os.read(master, 65536)
        except OSError:
            break
        if not data:
            break
        now = time.perf_counter()
        buf += data
        if t_loaded is None and re.search(rb"Read 0x[0-9A-F]+ hash", buf):
            t_loaded = now
        m = re.search(rb"Counting elapse (\d+) sec, total (\d+) kmers", buf)
        if m:
            t_done, total = now, int(m.group(2))
            break
    if proc.poll() is None:
        proc.kill()          # the exact process started above
    proc.wait()
    os.close(master)
    for ext in (".bin", ".txt"):
        Path(str(out_prefix) + ext).unlink(missing_ok=True)
    if total is None or t_loaded is None:
        raise RuntimeError("reference count did not report: " + buf.decode(errors="replace")[-500:])
    return total, t_done - t_loaded


def run_port_count(ref_prefix: Path, reads: Path, out_prefix: Path):
    t0 = time.perf_counter()
    res = subprocess.run([str(PORT_BIN), "count", str(ref_prefix), str(reads), str(out_prefix)], capture_output=True, text=True)
    dt = time.perf_counter() - t0
    for ext in (".bin", ".txt"):
        Path(str(out_prefix) + ext).unlink(missing_ok=True)
    return json.loads(res.stdout)["total_kmers"], dt


def cpu_threads():
    # the reference has ONE producer thread (Q.c:397-456) feeding -t N consumers; its README
    # (README.md:95) reports gains up to 6 threads.  Use what the box has, up to 16 consumers.
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    return max(1, min(16, n - 1)), n


def cpu_baseline(name, cdir, threads=None):
    w = WORKLOADS[name]
    if not REF_BIN.exists():      # only the single-threaded port is here: keep its sample to ~10-20 s of CPU work
        w = dict(w, sample_reads=max(1000, w["sample_reads"] // 64))
        WORKLOADS[name] = w
    d, ref, sample = prepare(name, cdir, sample=True)
    t, ncpu = cpu_threads()
    t = threads or t
    sleep_s = 0.0
    if REF_BIN.exists():
        total, secs = run_reference_count(ref, sample, t, d / f"cpu_out_{os.getpid()}")
        kind, cores = "reference", t + 1       # N consumers + the producer (main) thread
        # the threaded path ends with a fixed sleep(1) before joining its workers (Q.c:469): not work,
        # so it is taken out of the reference's time (in the reference's favour)
        sleep_s = 1.0 if t > 0 and secs > 2.0 else 0.0
    elif PORT_BIN.exists():
        total, secs = run_port_count(ref, sample, d / f"cpu_out_{os.getpid()}")
        kind, cores = "port", 1
    else:
        raise RuntimeError("neither oracle/_ref/quicKmer2 nor oracle/_build/qk_oracle is built")
    return {"value": total / (secs - sleep_s), "unit": "k-mers/s", "cores": cores, "kind": kind, "host_cpus": ncpu,
            "sample": f"{w['sample_reads']} reads of the workload ({total} k-mers) in {secs:.2f} s, "
                      + (f"quicKmer2 count -t {t}; dictionary load and .bin dump excluded; {sleep_s:.0f} s of that is the reference's fixed "
                         f"sleep(1) (Q.c:469) and is not counted"
                         if kind == "reference" else "oracle port, single thread, whole command"),
            "seconds": secs - sleep_s, "kmers": total}


def reference_arm(args):
    """`--impl reference`: the reference's CPU `count` on the box's host cores; a step is one
    run over the bounded sample.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cdir = cache_dir(args.cache_dir)
    subprocess.run(["make", "-s", "-C", str(PKG_DIR), str(SYNTH)], check=True)
    if not REF_BIN.exists():                     # the compiled reference did not travel: time the port instead
        subprocess.run(["make", "-s", "-C", str(ROOT / "oracle"), "port"], check=True)
    w = WORKLOADS[args.workload]
    times, kmers, base = [], 0, None
    for i in range(args.warmup + args.steps):
        base = cpu_baseline(args.workload, cdir)
        if i >= args.warmup:
            times.append(base["seconds"])
            kmers += base["kmers"]
        log(f"reference step {i}: {base['kmers'] / base['seconds'] / 1e6:.1f} M k-mers/s")
    value = kmers / sum(times)
    base["value"] = value
    base.pop("seconds"), base.pop("kmers")
    print(json.dumps({
        "impl": "reference", "metric": "k-mers/sec for quicKmer2 count", "value": value, "unit": "k-mers/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": args.workload, "desc": w["desc"], "step": f"CPU count over a {w['sample_reads']}-read sample"},
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


# --------------------------------------------------------------------------- GPU arm ------
def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def gpu_arm(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the count path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    qk = load_package()
    import importlib
    qd = importlib.import_module("quickmer2_b200.dist")
    if rank == 0:
        qk.build()
    else:
        while not (qk.LIB_PATH.exists() and qk.SYNTH_PATH.exists()):
            time.sleep(0.5)
        time.sleep(1.0)                              # rank 0's make may still be writing
    if os.environ.get("QK_BENCH_PREALLOC_MB"):       # diagnostic: another CUDA allocation before the context
        _shift = torch.empty(int(os.environ["QK_BENCH_PREALLOC_MB"]) << 20, dtype=torch.uint8, device=f"cuda:{local}")
    # The context goes first: with any other CUDA allocation made before it (a 64 MB torch tensor, or
    # NCCL's buffers at init) the count kernels run 6-7 % slower -- measured, cause not established
    # (profiles/README.md) -- so the process group is initialised after it.
    chunk_cap = args.chunk_mib << 20
    ctx = qk.Context(device=local, n_slots=args.slots, chunk_capacity=chunk_cap)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    cdir = cache_dir(args.cache_dir)
    w = WORKLOADS[args.workload]

    # ---- data: every rank has its own reads shard (seed + rank); rank 0 builds the dictionary
    if rank == 0:
        d, ref, reads = prepare(args.workload, cdir, 0)
    if world > 1:
        dist.barrier()
    d, ref, reads = prepare(args.workload, cdir, rank)

    t0 = time.perf_counter()
    if rank == 0:
        n_kmers = ctx.load_dictionary(ref.with_suffix(".fa.qm"))
    load_s = time.perf_counter() - t0
    bcast_s = 0.0
    if world > 1:
        # replicate the built table: descriptor through the host, image over NCCL/NVLink
        t0 = time.perf_counter()
        n_kmers = qd.replicate_dictionary(qk, ctx, rank, local, dist)
        torch.cuda.synchronize()
        bcast_s = time.perf_counter() - t0
    desc = ctx.table_desc()
    counters_t = qd.counters_tensor(ctx, local)
    counters_ab = [counters_t]
    if world > 1:                                 # second counter buffer: reduce of job s overlaps counting of job s + 1
        ctx.select_counters(1)
        counters_ab.append(qd.counters_tensor(ctx, local))
        ctx.select_counters(0)

    # ---- host side: raw reads in memory, framed chunks (pinned), device-resident copy ----
    raw_np = np.fromfile(reads, dtype=np.uint8)
    raw_pinned = torch.from_numpy(raw_np).pin_memory()       # the e2e leg DMAs straight from here
    t0 = time.perf_counter()
    chunks, fst = qk.frame(raw_np, seekable=True, chunk_capacity=chunk_cap)
    frame_s = time.perf_counter() - t0
    sizes = [len(c) for c in chunks]
    offs, at = [], 0
    for s in sizes:
        offs.append(at)
        at = (at + s + 255) // 256 * 256
    framed = torch.empty(at + 4096, dtype=torch.uint8).pin_memory()
    fnp = framed.numpy()
    for c, o in zip(chunks, offs):
        fnp[o:o + len(c)] = np.frombuffer(c, dtype=np.uint8)
    del chunks
    dev = framed.to(f"cuda:{local}")
    torch.cuda.synchronize()
    dev_base, host_base = dev.data_ptr(), framed.data_ptr()
    n_framed = sum(sizes)

    stream0 = torch.cuda.ExternalStream(ctx.slot_stream(0), device=f"cuda:{local}")

    pending = [None, None]                        # the reduce still in flight on each counter buffer
    job_no = [0]

    def job_device():
        # one stream (slot 0): launches run back to back, so the per-launch event times are not
        # inflated by two kernels sharing the SMs and the kernel's share of the step is meaningful.
        # Nothing here waits on the host: reset, kernels and (N > 1) the reduce are stream-ordered.
        b = job_no[0] & 1 if world > 1 else 0
        if world > 1:
            ctx.select_counters(b)
            if pending[b] is not None:            # slot 0 waits (on the device) for the reduce that last used this buffer
                with torch.cuda.stream(stream0):
                    pending[b].wait()
                pending[b] = None
        ctx.reset_async()
        for o, s in zip(offs, sizes):
            ctx.submit_device(dev_base + o, s, slot=0)

    def finish_device_step():
        b = job_no[0] & 1 if world > 1 else 0
        job_no[0] += 1
        if world > 1 and not os.environ.get("QK_BENCH_NO_REDUCE"):   # (diagnostic knob: what the reduce costs)
            with torch.cuda.stream(stream0):      # NCCL starts after slot 0's kernels; slot 0 does NOT wait for it
                pending[b] = dist.reduce(counters_ab[b], 0, op=dist.ReduceOp.SUM, async_op=True)

    def drain_device():
        for b in (0, 1):
            if pending[b] is not None:
                with torch.cuda.stream(stream0):
                    pending[b].wait()
                pending[b] = None

    def job_preframed():
        ctx.reset()
        for i, (o, s) in enumerate(zip(offs, sizes)):
            ctx.submit_host(host_base + o, s, slot=i % ctx.n_slots)

    def job_raw():
        ctx.reset()
        ctx.count_mem(raw_pinned.data_ptr(), raw_pinned.numel())

    def finish_step():
        """N > 1: combine the per-GPU counters on rank 0 (the one exchange step of the path)."""
        ctx.sync()
        if world > 1:
            dist.reduce(counters_t, 0, op=dist.ReduceOp.SUM)
            torch.cuda.synchronize()

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- value: inputs resident in HBM ---------------------------------------------------
    for _ in range(args.warmup):
        job_device(); finish_device_step()
    drain_device()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ctx.reset()                                   # zero the library's kernel/launch accounting
    t_wall0 = time.time()
    ctx.span_begin()
    for _ in range(args.steps):
        job_device(); finish_device_step()
    drain_device()                                # the last reduce is inside the span too
    span_ms = ctx.span_end()
    barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    tm = ctx.timing()                             # kernel_ms / launches summed over the K timed steps
    tm = {"kernel_ms": tm["kernel_ms"] / args.steps, "h2d_ms": tm["h2d_ms"] / args.steps, "launches": tm["launches"] // args.steps}
    stats = ctx.stats()                           # device totals of the last step (each step zeroes them)
    step_kmers, step_hits = stats["total_kmers"], stats["hits"]
    if world > 1:
        t = torch.tensor([span_ms], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        span_ms = float(t.item())
        t = torch.tensor([step_kmers, step_hits, n_framed], dtype=torch.int64, device=f"cuda:{local}")
        dist.all_reduce(t)
        job_kmers, job_hits, job_bytes = (int(x) for x in t.tolist())
    else:
        job_kmers, job_hits, job_bytes = step_kmers, step_hits, n_framed
    ms_per_step = span_ms / args.steps
    value = job_kmers / (ms_per_step * 1e-3)

    # sanity (outside the timed region): every hit landed on exactly one counter (the buffer of the last job)
    if rank == 0:
        total_counts = int(ctx.counters().astype(np.int64).sum())
        assert total_counts == job_hits or os.environ.get("QK_BENCH_NO_REDUCE"), (total_counts, job_hits)
    ctx.select_counters(0)

    # ---- e2e_preframed and e2e: host buffers, copies inside the timed region --------------
    result_pinned = torch.empty(n_kmers, dtype=torch.int16).pin_memory()
    result_np = result_pinned.numpy().view(np.uint16)

    def timed_host(job, steps, warm):
        for _ in range(warm):
            job(); finish_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            job(); finish_step()
            if rank == 0:
                ctx.finish(result_np)             # D2H of the step's result: uint16 depths in .bin order (pinned)
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt / steps

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    if args.kernel_only:
        pre_s = raw_s = float("nan")
        h2d_ms_step = None
    else:
        pre_s = timed_host(job_preframed, args.steps, 1)
        h2d_ms_step = ctx.timing()["h2d_ms"]
        raw_s = timed_host(job_raw, e2e_steps, 1)

    # ---- e2e from a FILE (page cache): reader threads pread() into the pinned slots ---------
    file_s = None
    if world == 1 and not args.kernel_only:
        def job_file():
            ctx.reset()
            ctx.count_file(reads, threads=args.reader_threads)
        file_s = timed_host(job_file, 2, 1)

    if rank != 0:
        ctx.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel -------------------------------------------------
    launches = int(tm["launches"])
    alg_bytes_step = 32 * step_kmers + 4 * step_hits + n_framed   # bucket sector + counter word + input byte
    avg_launch_ms = tm["kernel_ms"] / max(1, launches)
    achieved = alg_bytes_step / max(1, launches) / (avg_launch_ms * 1e-3) / 1e9
    peak, peak_src = measured_peaks()
    if args.kernel_only:
        gather = h2d = float("nan")
    else:
        gather = ctx.bench_gather(int(desc.table_bytes), gran=32, loads_in_flight=8, n_gathers=1 << 30)
        h2d = ctx.bench_h2d(min(chunk_cap, 64 << 20), repeats=16)
    ext = bool(desc.has_ext) and not os.environ.get("QK_CLASSIC_KERNEL")
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture of
    # `bench.py --kernel-only` on this workload (profiles/r1_count_ext_kernel_ncu_full_summary.csv); null elsewhere
    traffic = 1249798000 + 194004224 if (args.workload == "config2" and ext and args.chunk_mib == 64) else None
    roofline = {
        "bound": "hbm", "kernel": "qk_count_ext_kernel<4,1>" if ext else "qk_count_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "peak_source": peak_src, "traffic": traffic,
        "traffic_source": "ncu --set full, profiles/r1_count_ext_kernel_ncu_full_summary.csv" if traffic else None,
        "algorithmic_bytes_per_launch": alg_bytes_step / max(1, launches),
        "bytes_model": "SURVEY 8(d): 32 B bucket sector per emitted k-mer + 4 B counter per hit + 1 B per input byte"
                       + (" (the extension kernel derives most hits from dictionary order and probes far fewer sectors, so measured DRAM traffic is BELOW this figure)" if ext else ""),
        "probes_avoided_fraction": (stats.get("ext_verified", 0) / max(1, step_kmers)),
        "avg_launch_ms": avg_launch_ms, "launches_per_step": launches,
        "kernel_share_of_step": tm["kernel_ms"] / ms_per_step if world == 1 else None,
        "gather_peak_gbs": gather, "gather_peak_how": f"random 32 B sector loads over a {int(desc.table_bytes) >> 20} MiB table, 8 in flight/thread (qk_bench_gather)",
        "frac_of_gather_peak": (32 * step_kmers / max(1, launches)) / (avg_launch_ms * 1e-3) / 1e9 / gather,
        "h2d_peak_gbs": h2d,
        "frac_of_h2d_peak_preframed": (n_framed / pre_s / 1e9) / h2d,
    }

    base = None
    if world == 1 and not args.no_cpu and not args.kernel_only:
        try:
            base = cpu_baseline(args.workload, cdir)
            base.pop("seconds"), base.pop("kmers")
        except Exception as e:  # the baseline must not sink the GPU number
            base = {"value": None, "unit": "k-mers/s", "cores": 0, "kind": "unavailable", "sample": str(e)}

    bases_step = fst["bases"]
    out = {
        "metric": "k-mers/sec for quicKmer2 count", "value": value, "unit": "k-mers/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": args.workload, "desc": w["desc"], "k": int(desc.k), "dict_kmers": n_kmers,
                   "table_MiB": int(desc.table_bytes) >> 20, "reads_per_gpu": fst["lines"], "kmers_per_step": job_kmers,
                   "hit_fraction": job_hits / max(1, job_kmers), "chunk_MiB": args.chunk_mib,
                   "parallelism": f"reads sharded over {world} GPU(s), dictionary replicated"
                                  + (", NCCL reduce of the counters per step, overlapped with the next step's counting (two counter buffers)" if world > 1 else ""),
                   "l2": "inputs (framed reads + table) exceed the 126 MB L2 every step; no flush needed"
                         if n_framed + int(desc.table_bytes) > (256 << 20) else "working set fits L2: HBM term does not bind"},
        "bases_per_s": bases_step * world / (ms_per_step * 1e-3),
        "e2e": {"value": job_kmers / raw_s, "unit": "k-mers/s", "h2d_bytes_per_step": int(raw_np.size), "d2h_bytes_per_step": 2 * n_kmers + 64,
                "h2d_gbs": raw_np.size / raw_s / 1e9, "frac_of_h2d_peak": (raw_np.size / raw_s / 1e9) / h2d if h2d == h2d else None,
                "bases_per_s": bases_step * world / raw_s,
                "path": "raw FASTA/FASTQ bytes in pinned host memory -> qk_count_raw_mem (cut at line ends, H2D, device framing, count kernels) -> qk_finish (uint16 depths D2H)",
                "steps": e2e_steps, "raw_bytes_per_step": int(raw_np.size)},
        "e2e_preframed": {"value": job_kmers / pre_s, "unit": "k-mers/s", "h2d_gbs": n_framed / pre_s / 1e9,
                          "h2d_ms_per_step": h2d_ms_step,
                          "path": "pre-framed pinned host chunks -> qk_submit (H2D + kernel per chunk) -> qk_finish"},
        "e2e_file": None if file_s is None else {
            "value": job_kmers / file_s, "unit": "k-mers/s", "file_gbs": raw_np.size / file_s / 1e9, "reader_threads": args.reader_threads,
            "path": "reads FILE (page cache) -> qk_count_raw_file_mt (pread into pinned slots, H2D, device framing, count) -> qk_finish"},
        "gpu_launches": launches * args.steps,
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": base,
        "setup": {"dict_load_build_s": load_s, "dict_bcast_s": bcast_s, "host_frame_s": frame_s,
                  "stash_used": int(desc.stash_used), "n_buckets": int(desc.n_buckets)},
    }
    print(json.dumps(out).replace('NaN', 'null'), flush=True)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--cache-dir", default=None)
    ap.add_argument("--chunk-mib", type=int, default=64)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--reader-threads", type=int, default=12)
    ap.add_argument("--slots", type=int, default=12)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--kernel-only", action="store_true",
                    help="device-resident leg only (no e2e, micro-benchmarks or CPU leg): for ncu and quick iteration")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        log("note: the timing rules ask for >= 3 warm-up steps")
    # stdout carries exactly one JSON line: libraries that print there (NCCL's version banner
    # does when NCCL_DEBUG is set) are sent to stderr for the duration
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    real_stdout = os.fdopen(saved, "w")
    sys.stdout = real_stdout
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)
    real_stdout.flush()


if __name__ == "__main__":
    main()
