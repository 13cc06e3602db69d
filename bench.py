#!/usr/bin/env python
"""bench.py -- k-mers/s of the `quicKmer2 count` hot path on B200 (and the reference's CPU
path beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

One "step" = one whole counting job over the workload's read set: zero the depth counters,
run codec + probe + increment over every framed chunk, (N > 1: NCCL-reduce the counters to
rank 0).  Workloads follow BASELINE.json:configs:

  config3  3.1 Gb human-scale reference (16 contigs, ~1/6 of it segmental duplications), k=30,
           2^32-slot QM11 dictionary (48 GiB file, ~2.2 G unique k-mers, 32 GiB device table),
           30x = 620 M x 150 bp reads per GPU -- the default: the configuration the metric is
           quoted on (BASELINE.json:configs[2], 1/2/4/8 B200).  Generated on the GPU
           (tools/qk_synth_gpu.cu); the reads of the HBM-resident leg never exist on the host.
  config4  the same dictionary, 30x HiFi-like reads (median 15 kb, up to the 99,998-base line limit)
  config2  64 Mb reference with 200 x 20 kb segmental duplications, k=30 dictionary
           (58 M k-mers, 1 GiB device table), 30x = 12.8 M x 150 bp FASTQ reads
  config1  1 Mb reference, 1 M x 150 bp FASTA reads (the reference's CPU-runnable case)
  tiny / synth_small   smoke-sized (CI) versions of the file-based and the GPU-generated kind

Printed keys (one JSON line, rank 0):
  value        k-mers/s, framed chunks already resident in HBM (device-event span, max over ranks)
  e2e          k-mers/s through the public call (raw FASTA/FASTQ bytes in pinned host memory ->
               H2D of whole-line pieces -> device framing + count kernels -> D2H of the uint16 depths)
  e2e_preframed  same but from pre-framed pinned chunks: the H2D-overlap pipeline alone
  roofline     dominant kernel: SURVEY 8(d) algorithmic bytes / mean launch time (`frac`), and the same on
               the bytes the kernel actually asks for, from its own probe counters (`frac_issued`)
  parity_ok    outside the timed region: every rank counts its shard once more; the reduced counters
               must equal the sum of the per-rank ones under a position-weighted checksum (linearity)
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


HUMAN = dict(n_bases=3_100_000_000, contigs=16, seed=2026, dup_period=120_000, dup_len=26_000, div_ppm=5000, nblock=50_000)
SMALL = dict(n_bases=40_000_000, contigs=4, seed=2026, dup_period=120_000, dup_len=26_000, div_ppm=5000, nblock=50_000)
# GPU-generated workloads (tools/qk_synth_gpu.cu): genome + dictionary made on the device and written as the
# QM11 files both arms read; reads generated straight into HBM (value) / pinned host memory (e2e) / a
# sample file (CPU arm).  reads = per GPU.
SYNTH_WORKLOADS = {
    "config3": dict(genome=HUMAN, dict_dir="human_k30", k=30, slots=1 << 32, ctrl_block=100_000,
                    reads=620_000_000, read_len=150, err_ppm=2000, fastq=True, hifi=False, seed=42, sample_reads=20_700_000,
                    desc="3.1 Gb synthetic human-scale reference (16 contigs, 22% segmental duplications at 0.5% divergence), k=30, "
                         "2^32-slot .qm; 30x = 620M x 150bp reads per GPU"),
    "config4": dict(genome=HUMAN, dict_dir="human_k30", k=30, slots=1 << 32, ctrl_block=100_000,
                    reads=5_470_000, read_len=15000, err_ppm=1000, fastq=False, hifi=True, seed=42, sample_reads=150_000,
                    desc="config-3 dictionary; 30x HiFi-like reads (log-normal, median 15 kb, clipped to [1 kb, 99,998]), FASTA"),
    "synth_small": dict(genome=SMALL, dict_dir="small_k30", k=30, slots=1 << 26, ctrl_block=100_000,
                        reads=2_000_000, read_len=150, err_ppm=2000, fastq=True, hifi=False, seed=42, sample_reads=200_000,
                        desc="40 Mb GPU-generated reference, k=30, 2M x 150bp reads (CI-sized config 3)"),
    "synth_small_hifi": dict(genome=SMALL, dict_dir="small_k30", k=30, slots=1 << 26, ctrl_block=100_000,
                             reads=20_000, read_len=15000, err_ppm=1000, fastq=False, hifi=True, seed=42, sample_reads=2_000,
                             desc="40 Mb GPU-generated reference, k=30, 20k HiFi-like reads (CI-sized config 4)"),
}


def load_synth_gpu():
    sys.path.insert(0, str(ROOT / "tools"))
    import qk_synth_gpu
    return qk_synth_gpu


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
        # the human-scale dictionary files need ~60 GB (tmpfs counts against RAM: ask for headroom)
        base = Path("/dev/shm") if Path("/dev/shm").is_dir() and shutil.disk_usage("/dev/shm").free > 100 << 30 else Path("/tmp")
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


def synth_genome(w, device=0):
    qs = load_synth_gpu()
    return qs.Genome.create(device=device, **w["genome"])


def synth_lens(w, seed, first, n):
    """Lengths of reads [first, first + n) of stream `seed` (HiFi workloads), else None."""
    if not w["hifi"]:
        return None
    qs = load_synth_gpu()
    # blocks of 2^20 reads so that any sub-range sees the same lengths
    B = 1 << 20
    out = []
    for b in range(first // B, (first + n + B - 1) // B):
        out.append(qs.hifi_lengths(B, seed * 1_000_003 + b, median=w["read_len"]))
    lens = np.concatenate(out) if out else np.zeros(0, np.uint32)
    return lens[first - (first // B) * B:][:n]


def prepare_synth(name, cdir, sample=False, genome=None, device=0):
    """GPU-generated workload: the QM11 dictionary files (shared by the workloads that name the same
    dict_dir) and, if asked, the reads sample file of the CPU arm.  Needs a CUDA device only when
    something is missing from the cache.  Returns (dir, ref prefix, sample path or None)."""
    w = SYNTH_WORKLOADS[name]
    qs = load_synth_gpu()
    d = cdir / w["dict_dir"]
    d.mkdir(parents=True, exist_ok=True)
    ref = d / "ref.fa"
    own = [None]

    def g():
        if genome is not None:
            return genome
        if own[0] is None:
            own[0] = synth_genome(w, device)
        return own[0]

    def make_dict():
        info = g().build_dict(k=w["k"], slots=w["slots"], ctrl_block=w["ctrl_block"])
        log(f"dictionary on the GPU: {info}")
        g().write_dict(d / "tmpdict", threads=min(16, os.cpu_count() or 1))
        g().free_dict()
        (d / "dict_info.json").write_text(json.dumps(info))
        os.replace(d / "tmpdict.qgc", d / "ref.fa.qgc")
        os.replace(d / "tmpdict.qm", d / "ref.fa.qm")
    ensure(d / "ref.fa.qm", make_dict)
    reads = None
    if sample:
        n = w["sample_reads"]
        reads = d / f"sample_{name}_{n}_s{w['seed']}.{'fq' if w['fastq'] else 'fa'}"

        def make_sample():
            tmp = reads.with_suffix(".tmp")
            g().reads_to_file(tmp, w["seed"], 0, n, w["read_len"], w["err_ppm"], qs.FASTQ if w["fastq"] else qs.FASTA,
                              lens=synth_lens(w, w["seed"], 0, n))
            os.replace(tmp, reads)
        ensure(reads, make_sample)
    if own[0] is not None:
        own[0].close()
    return d, ref, reads


def prepare(name, cdir, rank_seed_offset=0, sample=False):
    """Reference, dictionary and reads of a workload; returns paths."""
    if name in SYNTH_WORKLOADS:
        return prepare_synth(name, cdir, sample=sample)
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
    table = SYNTH_WORKLOADS if name in SYNTH_WORKLOADS else WORKLOADS
    w = table[name]
    if not REF_BIN.exists():      # only the single-threaded port is here: keep its sample to ~10-20 s of CPU work
        w = dict(w, sample_reads=max(1000, w["sample_reads"] // 64))
        table[name] = w
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
    subprocess.run(["make", "-s", "-C", str(PKG_DIR), "all"], check=True)      # qk_synth and the GPU data generator (prebuilt files travel)
    if not REF_BIN.exists():                     # the compiled reference did not travel: time the port instead
        subprocess.run(["make", "-s", "-C", str(ROOT / "oracle"), "port"], check=True)
    w = (SYNTH_WORKLOADS if args.workload in SYNTH_WORKLOADS else WORKLOADS)[args.workload]
    times, kmers, base = [], 0, None
    t_start, steps_done, warm_done = time.time(), 0, 0
    for i in range(args.warmup + args.steps):
        # Every step is a fresh run of the reference command, which first loads the dictionary (48 GiB at
        # human scale): once the run has used its time budget, stop after >= 1 warm-up and >= 2 timed
        # steps and report the steps really taken.
        over = time.time() - t_start > args.reference_budget_s
        if i < args.warmup and over and warm_done >= 1:
            continue
        if i >= args.warmup and over and steps_done >= 2:
            break
        base = cpu_baseline(args.workload, cdir)
        if i >= args.warmup:
            times.append(base["seconds"])
            kmers += base["kmers"]
            steps_done += 1
        else:
            warm_done += 1
        log(f"reference step {i}: {base['kmers'] / base['seconds'] / 1e6:.1f} M k-mers/s")
    value = kmers / sum(times)
    base["value"] = value
    base.pop("seconds"), base.pop("kmers")
    print(json.dumps({
        "impl": "reference", "metric": "k-mers/sec for quicKmer2 count", "value": value, "unit": "k-mers/s",
        "n_gpus": args.gpus, "steps": steps_done, "warmup": warm_done, "ms_per_step": 1e3 * sum(times) / len(times),
        "steps_requested": args.steps, "warmup_requested": args.warmup,
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


def profile_traffic(kernel, workload, chunk_mib):
    """DRAM bytes per launch of the dominant kernel from a committed `ncu --set full` capture of this
    very workload (profiles/traffic.json, keyed by kernel | workload | chunk size), else None."""
    p = ROOT / "profiles" / "traffic.json"
    if not p.exists():
        return None, None
    ent = json.loads(p.read_text()).get(f"{kernel}|{workload}|{chunk_mib}")
    return (ent["dram_bytes_per_launch"], ent["source"]) if ent else (None, None)


class FileData:
    """Reads of a file-based workload: raw bytes (pinned), host-framed chunks (pinned) and their copy in HBM."""

    def __init__(self, qk, torch, reads, chunk_cap, local):
        raw_np = np.fromfile(reads, dtype=np.uint8)
        self.raw = torch.from_numpy(raw_np).pin_memory()
        self.raw_ptr = self.raw.data_ptr()
        self.raw_bytes = int(raw_np.size)
        t0 = time.perf_counter()
        chunks, fst = qk.frame(raw_np, seekable=True, chunk_capacity=chunk_cap)
        self.frame_s = time.perf_counter() - t0
        self.sizes = [len(c) for c in chunks]
        self.offs, at = [], 0
        for sz in self.sizes:
            self.offs.append(at)
            at = (at + sz + 255) // 256 * 256
        self.framed = torch.empty(at + 4096, dtype=torch.uint8).pin_memory()
        fnp = self.framed.numpy()
        for c, o in zip(chunks, self.offs):
            fnp[o:o + len(c)] = np.frombuffer(c, dtype=np.uint8)
        del chunks
        self.dev = self.framed.to(f"cuda:{local}")
        torch.cuda.synchronize()
        self.dev_base, self.host_base = self.dev.data_ptr(), self.framed.data_ptr()
        self.n_framed = sum(self.sizes)
        self.lines, self.bases = fst["lines"], fst["bases"]
        # the host legs run over the same reads
        self.h_offs, self.h_sizes, self.h_framed_bytes = self.offs, self.sizes, self.n_framed
        self.h_lines, self.h_bases = self.lines, self.bases
        self.reads_file = reads
        self.note = None


class SynthData:
    """Reads of a GPU-generated workload.  value: `reads` records per GPU generated straight into HBM as
    framed sequence lines (they never exist on the host).  Host legs (e2e, e2e_preframed): a bounded
    prefix of the same stream, as raw FASTQ/FASTA and as framed lines, in pinned host memory."""

    def __init__(self, qs, torch, genome, w, seed, chunk_cap, local, world, reads_scale, e2e_cov_cap):
        fmt_raw = qs.FASTQ if w["fastq"] else qs.FASTA
        n = max(1000, int(w["reads"] * reads_scale))
        L = w["read_len"]
        free_b, _ = C_mem_info(qs, local)
        # ---- value: framed reads in HBM, cut into chunks of whole records at 16-byte-aligned offsets
        lens = synth_lens(w, seed, 0, n)
        per = (lens.astype(np.uint64) + 1) if lens is not None else None
        budget = free_b - (6 << 30)
        need = int(per.sum()) if per is not None else n * (L + 1)
        if need > budget:                                 # (a table larger than planned): fewer reads, said in `note`
            keep = budget / need
            n = int(n * keep) // 16 * 16
            lens = None if lens is None else lens[:n]
            per = None if per is None else per[:n]
            self.note = f"reads per GPU cut to {n} to fit HBM next to the table"
        else:
            self.note = None
        if lens is None:
            rpc = max(16, (chunk_cap // (L + 1)) // 16 * 16)        # reads per chunk; 16 records = a multiple of 16 bytes
            starts = list(range(0, n, rpc))
            self.sizes = [min(rpc, n - a) * (L + 1) for a in starts]
            self.offs = [a * (L + 1) for a in starts]
            total = n * (L + 1)
            self.devbuf = qs.DeviceBuffer(local, total + 4096)
            genome.reads_into(self.devbuf.ptr, seed, 0, n, L, w["err_ppm"], qs.FRAMED)
            self.bases = n * L
        else:
            offsets, self.offs, self.sizes = chunk_layout(per, chunk_cap)
            total = int(offsets[-1] + per[-1]) if n else 0
            self.devbuf = qs.DeviceBuffer(local, total + 4096)
            genome.reads_into(self.devbuf.ptr, seed, 0, n, L, w["err_ppm"], qs.FRAMED, lens, offsets)
            self.bases = int(lens.astype(np.uint64).sum())
        self.dev_base = self.devbuf.ptr
        self.n_framed = sum(self.sizes)
        self.lines = n
        self.frame_s = 0.0

        # ---- host legs: a prefix of the stream in pinned memory, sized to what the host can pin
        avail = host_mem_available()
        rec_raw = qs.record_bytes(L, fmt_raw) if lens is None else None
        mean_raw = rec_raw if lens is None else float(np.mean(per)) * (1 if not w["fastq"] else 2) + 15
        mean_framed = (L + 1) if lens is None else float(np.mean(per))
        # leave room for the reference's CPU run that follows on rank 0 (its table + counters: 10 bytes per slot)
        reserve = (24 << 30) + 10 * w["slots"]
        cap_bytes = max(1 << 28, int((avail - reserve) / max(1, world) * 0.6))
        hn = int(min(n, e2e_cov_cap * n / 30.0, cap_bytes / (mean_raw + mean_framed)))
        hn = max(16, hn // 16 * 16)
        hlens = None if lens is None else lens[:hn]
        raw_total, raw_offsets = qs.layout(hn, L, fmt_raw, hlens)
        self._qs = qs
        self.raw_ptr = qs.lib().qs_pinned_alloc(raw_total + 64)          # (not torch: its pinned allocator never gives memory back)
        if not self.raw_ptr:
            raise MemoryError(f"cannot pin {raw_total} bytes of host memory")
        self.raw_bytes = raw_total
        fill_pinned(qs, genome, self.raw_ptr, seed, hn, L, w["err_ppm"], fmt_raw, hlens, raw_offsets, local)
        if hlens is None:
            rpc = max(16, (chunk_cap // (L + 1)) // 16 * 16)
            starts = list(range(0, hn, rpc))
            self.h_sizes = [min(rpc, hn - a) * (L + 1) for a in starts]
            self.h_offs = [a * (L + 1) for a in starts]
            f_total, f_offsets = hn * (L + 1), None
        else:
            f_offsets, self.h_offs, self.h_sizes = chunk_layout(per[:hn], chunk_cap)
            f_total = int(f_offsets[-1] + per[hn - 1])
        self.host_base = qs.lib().qs_pinned_alloc(f_total + 4096)
        if not self.host_base:
            raise MemoryError(f"cannot pin {f_total} bytes of host memory")
        fill_pinned(qs, genome, self.host_base, seed, hn, L, w["err_ppm"], qs.FRAMED, hlens, f_offsets, local)
        self.h_framed_bytes = sum(self.h_sizes)
        self.h_lines = hn
        self.h_bases = hn * L if hlens is None else int(hlens.astype(np.uint64).sum())
        self.reads_file = None

    def free_host(self):
        for p in (self.raw_ptr, self.host_base):
            if p:
                self._qs.lib().qs_pinned_free(p)
        self.raw_ptr = self.host_base = None


def chunk_layout(per, chunk_cap):
    """Records of `per[i]` bytes packed into chunks of at most chunk_cap bytes, every chunk starting at a
    16-byte-aligned offset: returns (record offsets, chunk offsets, chunk sizes)."""
    n = per.size
    offsets = np.zeros(n, dtype=np.uint64)
    c_offs, c_sizes = [], []
    csum = np.concatenate([[0], np.cumsum(per, dtype=np.uint64)])
    i, at = 0, 0
    while i < n:
        j = int(np.searchsorted(csum, csum[i] + np.uint64(chunk_cap - 16), side="right")) - 1
        j = max(j, i + 1)
        offsets[i:j] = (csum[i:j] - csum[i]) + np.uint64(at)
        size = int(csum[j] - csum[i])
        c_offs.append(at)
        c_sizes.append(size)
        at = (at + size + 255) // 256 * 256
        i = j
    return offsets, c_offs, c_sizes


def fill_pinned(qs, genome, host_ptr, seed, n, L, err_ppm, fmt, lens, offsets, local, piece_bytes=1 << 30):
    """Generate records [0, n) on the device a piece at a time and copy them to pinned host memory at
    host_ptr (+ their offsets)."""
    tmp = qs.DeviceBuffer(local, piece_bytes + (1 << 20))
    try:
        if lens is None:
            rec = qs.record_bytes(L, fmt)
            per_piece = max(1, piece_bytes // rec)
            for a in range(0, n, per_piece):
                m = min(per_piece, n - a)
                genome.reads_into(tmp.ptr, seed, a, m, L, err_ppm, fmt)
                if qs.lib().qs_copy_to_host(host_ptr + a * rec, tmp.ptr, m * rec):
                    raise RuntimeError("D2H failed")
        else:
            rec = (lens.astype(np.uint64) + 1) if fmt == qs.FRAMED else (13 + lens.astype(np.uint64) + 1) if fmt == qs.FASTA \
                else (13 + 2 * (lens.astype(np.uint64) + 1) + 2)
            ends = offsets + rec
            a = 0
            while a < n:                          # groups of records spanning at most one staging piece
                base = int(offsets[a])
                b = max(a + 1, int(np.searchsorted(ends, np.uint64(base + piece_bytes), side="right")))
                span = int(ends[b - 1]) - base    # (alignment gaps between chunks carry garbage: never read)
                genome.reads_into(tmp.ptr, seed, a, b - a, L, err_ppm, fmt, lens[a:b], offsets[a:b] - np.uint64(base))
                if qs.lib().qs_copy_to_host(host_ptr + base, tmp.ptr, span):
                    raise RuntimeError("D2H failed")
                a = b
    finally:
        tmp.free()


def host_mem_available():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) << 10
    except OSError:
        pass
    return 32 << 30


def C_mem_info(qs, device):
    import ctypes
    f, t = ctypes.c_uint64(), ctypes.c_uint64()
    qs.lib().qs_device_mem_info(device, ctypes.byref(f), ctypes.byref(t))
    return f.value, t.value


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
    # The context goes first (see profiles/README.md, allocation order), the process group after it.
    chunk_cap = args.chunk_mib << 20
    ctx = qk.Context(device=local, n_slots=args.slots, chunk_capacity=chunk_cap)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    cdir = cache_dir(args.cache_dir)
    synth = args.workload in SYNTH_WORKLOADS
    w = (SYNTH_WORKLOADS if synth else WORKLOADS)[args.workload]

    # ---- data: every rank has its own reads shard (seed + rank); rank 0 builds the dictionary
    t_setup = time.perf_counter()
    genome = None
    if synth:
        qs = load_synth_gpu()
        genome = synth_genome(w, local)
        if rank == 0:
            d, ref, _ = prepare_synth(args.workload, cdir, sample=not args.no_cpu and not args.kernel_only, genome=genome, device=local)
        if world > 1:
            dist.barrier()
        d, ref = cdir / w["dict_dir"], cdir / w["dict_dir"] / "ref.fa"
        reads = None
    else:
        if rank == 0:
            d, ref, reads = prepare(args.workload, cdir, 0)
        if world > 1:
            dist.barrier()
        d, ref, reads = prepare(args.workload, cdir, rank)
    gen_s = time.perf_counter() - t_setup

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

    # ---- reads: in HBM for `value`; raw + framed in pinned host memory for the e2e legs ----
    t0 = time.perf_counter()
    if synth:
        data = SynthData(qs, torch, genome, w, w["seed"] + rank, chunk_cap, local, world, args.reads_scale, args.e2e_coverage)
        genome.close()
    else:
        data = FileData(qk, torch, reads, chunk_cap, local)
    reads_s = time.perf_counter() - t0
    log(f"rank {rank}: data ready (dictionary files {gen_s:.1f} s, load+build {load_s:.1f} s, broadcast {bcast_s:.1f} s, reads {reads_s:.1f} s); "
        f"{data.lines} reads / {data.n_framed >> 20} MiB framed in HBM in {len(data.sizes)} chunks")

    stream0 = torch.cuda.ExternalStream(ctx.slot_stream(0), device=f"cuda:{local}")

    pending = [None, None]                        # the reduce still in flight on each counter buffer
    job_no = [0]

    def job_device(reduce=True):
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
        base = data.dev_base
        for o, sz in zip(data.offs, data.sizes):
            ctx.submit_device(base + o, sz, slot=0)
        job_no[0] += 1
        if world > 1 and reduce and not os.environ.get("QK_BENCH_NO_REDUCE"):   # (diagnostic knob: what the reduce costs)
            with torch.cuda.stream(stream0):      # NCCL starts after slot 0's kernels; slot 0 does NOT wait for it
                pending[b] = dist.reduce(counters_ab[b], 0, op=dist.ReduceOp.SUM, async_op=True)
        return b

    def drain_device():
        for b in (0, 1):
            if pending[b] is not None:
                with torch.cuda.stream(stream0):
                    pending[b].wait()
                pending[b] = None

    def job_preframed():
        ctx.reset_async()
        for i, (o, sz) in enumerate(zip(data.h_offs, data.h_sizes)):
            ctx.submit_host(data.host_base + o, sz, slot=i % ctx.n_slots)

    cpus = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    framer_threads = args.framer_threads or max(1, cpus // max(1, world))
    # auto: the host frames FASTQ when it has the cores (it halves the bytes on the link); FASTA gains nothing from it
    is_fastq = w["fastq"]
    host_framing = args.framer == "host" or (args.framer == "auto" and framer_threads >= 12 and is_fastq)   # measured: profiles/README.md

    def job_raw_device():
        ctx.reset_async()
        ctx.count_mem(data.raw_ptr, data.raw_bytes)

    shipped = {"bytes": 0}                      # what the host framer handed to the GPU in its last job (text lines, or packed chunks)

    def job_raw_host():
        ctx.reset_async()
        shipped["bytes"] = ctx.count_mem_mt(data.raw_ptr, data.raw_bytes, threads=framer_threads)["sink_bytes"]

    job_raw = job_raw_host if host_framing else job_raw_device

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

    def all_sum(vals):
        if world == 1:
            return [int(v) for v in vals]
        t = torch.tensor(vals, dtype=torch.int64, device=f"cuda:{local}")
        dist.all_reduce(t)
        return [int(x) for x in t.tolist()]

    def all_max(v):
        if world == 1:
            return float(v)
        t = torch.tensor([v], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: inputs resident in HBM ---------------------------------------------------
    for _ in range(args.warmup):
        job_device()
    drain_device()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ctx.reset()                                   # zero the library's kernel/launch accounting
    t_wall0 = time.time()
    ctx.span_begin()
    for _ in range(args.steps):
        job_device()
    drain_device()                                # the last reduce is inside the span too
    span_ms = ctx.span_end()
    barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    tm = ctx.timing()                             # kernel_ms / launches summed over the K timed steps
    tm = {"kernel_ms": tm["kernel_ms"] / args.steps, "h2d_ms": tm["h2d_ms"] / args.steps, "launches": tm["launches"] // args.steps}
    stats = ctx.stats()                           # device totals of the last step (each step zeroes them)
    step_kmers, step_hits = stats["total_kmers"], stats["hits"]
    span_ms = all_max(span_ms)
    job_kmers, job_hits, job_bytes, job_bases = all_sum([step_kmers, step_hits, data.n_framed, data.bases])
    ms_per_step = span_ms / args.steps
    value = job_kmers / (ms_per_step * 1e-3)

    # ---- parity, outside the timed region: the reduced counters are the sum of the per-rank ones ---------
    # One more job without the reduce; every rank takes a position-weighted checksum of its own counters
    # (linear in the counters, arithmetic mod 2^63), the checksums are added over the ranks; then the reduce
    # runs and rank 0 takes the same checksum of the reduced counters.  Also: every hit is on exactly one counter.
    def checksum(t):
        tot, wsum = 0, 0
        n = t.numel()
        piece = 1 << 27
        for a in range(0, n, piece):
            x = t[a:a + piece].to(torch.int64) & 0xFFFFFFFF
            idx = torch.arange(a, a + x.numel(), device=x.device, dtype=torch.int64)
            wgt = ((idx * 2654435761) & 0xFFFFF) | 1
            tot += int(x.sum().item())
            wsum = (wsum + int((x * wgt).sum().item())) & ((1 << 62) - 1)
        return tot, wsum

    b = job_device(reduce=False)
    ctx.sync()
    torch.cuda.synchronize()
    own_tot, own_w = checksum(counters_ab[b])
    own_hits = ctx.stats()["hits"]
    sum_tot, sum_hits = all_sum([own_tot, own_hits])
    (sum_w,) = all_sum([own_w])
    sum_w &= (1 << 62) - 1
    if world > 1:
        dist.reduce(counters_ab[b], 0, op=dist.ReduceOp.SUM)
        torch.cuda.synchronize()
    parity = None
    if rank == 0:
        red_tot, red_w = checksum(counters_ab[b])
        parity = {"ok": bool(red_tot == sum_tot == sum_hits and red_w == sum_w), "counter_sum": red_tot, "hits": sum_hits,
                  "weighted_checksum_reduced": red_w, "weighted_checksum_sum_of_ranks": sum_w,
                  "how": "extra job outside the timed region: sum over ranks of a position-weighted checksum of each rank's own "
                         "counters == the same checksum of the NCCL-reduced counters on rank 0; counter sum == hits"}
        assert parity["ok"] or os.environ.get("QK_BENCH_NO_REDUCE"), parity
    ctx.select_counters(0)

    # ---- e2e_preframed and e2e: host buffers, copies inside the timed region --------------
    result_np, result_ptr = [None, None], [None, None]
    if rank == 0:                                   # (own pinned allocations: given back before the CPU leg)
        import ctypes
        qsl = load_synth_gpu().lib()
        for i in (0, 1):                            # two: the read-back of one step overlaps the counting of the next
            result_ptr[i] = qsl.qs_pinned_alloc(2 * n_kmers + 64)
            if not result_ptr[i]:
                raise MemoryError("cannot pin the result buffer")
            result_np[i] = np.ctypeslib.as_array(ctypes.cast(result_ptr[i], ctypes.POINTER(ctypes.c_uint16)), shape=(n_kmers,))

    def timed_host(job, steps, warm):
        """Host-buffer legs.  Every step: inputs from host memory (H2D inside), count, (N > 1: reduce), and the
        read-back of the step's uint16 depths into pinned memory.  Steps alternate between the two counter
        buffers so that the read-back of step s (qk_finish_async, its own stream) runs while step s + 1 is being
        counted; the timer stops only when the last read-back has landed."""
        def one(step):
            b = step & 1
            ctx.select_counters(b)
            job()
            ctx.sync()
            if world > 1:
                dist.reduce(counters_ab[b], 0, op=dist.ReduceOp.SUM)
                torch.cuda.synchronize()
            marks.append(time.perf_counter())
            if os.environ.get("QK_BENCH_DEBUG"):
                log(f"  {job.__name__} step {step} buffer {b}: {ctx.stats()}")
            if rank == 0:
                ctx.finish_wait()                 # the previous step's depths (normally long there)
                ctx.finish_async(result_np[b])    # this step's: uint16, .bin order, pinned
            marks.append(time.perf_counter())
        marks = []
        for i in range(warm):
            one(i)
        if rank == 0:
            ctx.finish_wait()
        barrier()
        marks = []
        t0 = time.perf_counter()
        for i in range(steps):
            one(warm + i)
        if rank == 0:
            ctx.finish_wait()
        barrier()
        dt = all_max(time.perf_counter() - t0)
        if rank == 0:
            log(f"{job.__name__}: per step [count ms, hand the result to the read-back ms] = "
                + str([round(1e3 * (b - a), 1) for a, b in zip([t0] + marks[:-1], marks)]) + f"; total {1e3 * dt:.1f} ms")
        st = ctx.stats()
        (k, h) = all_sum([st["total_kmers"], st["hits"]])
        ctx.select_counters(0)
        return dt / steps, (k, h)               # k-mers and dictionary hits of the last job: every leg must agree on both

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    h_kmers = None
    if args.kernel_only:
        pre_s = raw_s = float("nan")
        h2d_ms_step = None
    else:
        h2d_before = ctx.timing()["h2d_ms"]
        pre_s, h_seen = timed_host(job_preframed, e2e_steps, 1)
        h_kmers = h_seen[0]
        h2d_ms_step = (ctx.timing()["h2d_ms"] - h2d_before) / (e2e_steps + 1)
        raw_s, h_seen_raw = timed_host(job_raw, e2e_steps, 1)
        assert h_seen_raw == h_seen, (h_seen_raw, h_seen)
        other_s = None
        if world == 1:                              # the other framing policy, for comparison
            other_s, k2 = timed_host(job_raw_device if host_framing else job_raw_host, e2e_steps, 1)
            assert k2 == h_seen, (k2, h_seen)
        main_shipped = shipped["bytes"] if host_framing else 0
        unpacked_s = None
        if host_framing and main_shipped and main_shipped < 0.9 * data.h_framed_bytes:
            # ... and what the same path does when the host ships the sequence lines as text (QK_PACKED=0)
            keep = os.environ.get("QK_PACKED")
            os.environ["QK_PACKED"] = "0"
            try:
                unpacked_s, k3 = timed_host(job_raw_host, e2e_steps, 1)
            finally:
                if keep is None:
                    del os.environ["QK_PACKED"]
                else:
                    os.environ["QK_PACKED"] = keep
            assert k3 == h_seen, (k3, h_seen)
            shipped["bytes"] = main_shipped
    (h_raw_bytes, h_framed_bytes, h_bases, h_shipped) = all_sum([data.raw_bytes, data.h_framed_bytes, data.h_bases, shipped["bytes"]])
    if not h_shipped:
        h_shipped = h_framed_bytes
    packed = h_shipped < 0.9 * h_framed_bytes       # the host framer packs to 0.375 bytes per position when the kernel can read that

    # ---- e2e from a FILE (page cache): reader threads pread() into the pinned slots ---------
    file_s = None
    file_reads = data.reads_file
    if synth and rank == 0 and not args.no_cpu and not args.kernel_only:
        file_reads = prepare_synth(args.workload, cdir, sample=True)[2]      # the CPU arm's sample (already cached)
    if world == 1 and not args.kernel_only and file_reads is not None:
        def job_file():
            ctx.reset_async()
            if host_framing:
                ctx.count_file_mt(file_reads, threads=framer_threads)
            else:
                ctx.count_file(file_reads, threads=args.reader_threads)
        file_s, (file_kmers, _) = timed_host(job_file, 2, 1)
        file_bytes = os.path.getsize(file_reads)

    framer_gbs = None
    if rank == 0 and not args.kernel_only:
        framer_gbs = qk.bench_framer(data.raw_ptr, min(data.raw_bytes, 4 << 30), threads=framer_threads, repeats=2)
    if rank == 0:
        result_np = [None, None]
        for p in result_ptr:
            qsl.qs_pinned_free(p)
    if synth:
        data.free_host()
        data.devbuf.free()                          # the micro-benchmarks below need the HBM
    else:
        data.dev = None
    if rank != 0:
        ctx.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel -------------------------------------------------
    launches = int(tm["launches"])
    alg_bytes_step = 32 * step_kmers + 4 * step_hits + data.n_framed   # bucket sector + counter word + input byte
    # what the kernel really asks of memory, from its own counters: 32 B per bucket probe it issued, 16 B of
    # extension-array words per walk it started, 4 B per hit, 1 B per input byte
    issued_bytes_step = 32 * stats["bucket_probes"] + 16 * stats["walks"] + 4 * step_hits + data.n_framed
    avg_launch_ms = tm["kernel_ms"] / max(1, launches)
    achieved = alg_bytes_step / max(1, launches) / (avg_launch_ms * 1e-3) / 1e9
    achieved_issued = issued_bytes_step / max(1, launches) / (avg_launch_ms * 1e-3) / 1e9
    peak, peak_src = measured_peaks()
    if args.kernel_only:
        gather = h2d = float("nan")
    else:
        torch.cuda.empty_cache()
        gather = ctx.bench_gather(int(desc.table_bytes), gran=32, loads_in_flight=8, n_gathers=1 << 30)
        h2d = ctx.bench_h2d(min(chunk_cap, 64 << 20), repeats=16)
    ext = bool(desc.has_ext) and not os.environ.get("QK_CLASSIC_KERNEL")
    kernel_name = "qk_count_ext32_kernel" if ext else "qk_count_kernel"
    traffic, traffic_src = profile_traffic(kernel_name, args.workload, args.chunk_mib)
    probe_rate = stats["bucket_probes"] / max(1, launches) / (avg_launch_ms * 1e-3)
    roofline = {
        "bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_src,
        "algorithmic_bytes_per_launch": alg_bytes_step / max(1, launches),
        "bytes_model": "SURVEY 8(d): 32 B bucket sector per emitted k-mer + 4 B counter per hit + 1 B per input byte",
        "achieved_issued": achieved_issued, "frac_issued": achieved_issued / peak,
        "issued_bytes_per_launch": issued_bytes_step / max(1, launches),
        "issued_model": "bytes the kernel asks for, from its own counters: 32 B x bucket probes issued + 16 B x dictionary-order walks "
                        "started + 4 B x hits + 1 B x input bytes",
        "bucket_probes_per_kmer": stats["bucket_probes"] / max(1, step_kmers),
        "probes_avoided_fraction": 1.0 - stats["bucket_probes"] / max(1, step_kmers),
        "hits_by_walk_fraction": stats.get("ext_verified", 0) / max(1, step_hits),
        "avg_launch_ms": avg_launch_ms, "launches_per_step": launches,
        "kernel_share_of_step": tm["kernel_ms"] / ms_per_step if world == 1 else None,
        "gather_peak_gbs": gather, "gather_peak_how": f"random 32 B sector loads over a {int(desc.table_bytes) >> 20} MiB table, 8 in flight/thread (qk_bench_gather)",
        "probe_rate_gps": probe_rate / 1e9,
        "frac_of_gather_peak": 32 * probe_rate / 1e9 / gather if gather == gather else None,
        "frac_of_gather_peak_how": "bucket probes issued per second x 32 B / the measured random-sector gather GB/s",
        "h2d_peak_gbs": h2d,
        "frac_of_h2d_peak_preframed": (h_framed_bytes / world / pre_s / 1e9) / h2d if h2d == h2d else None,
    }

    base = None
    if world == 1 and not args.no_cpu and not args.kernel_only:
        try:
            base = cpu_baseline(args.workload, cdir)
            base.pop("seconds"), base.pop("kmers")
        except Exception as e:  # the baseline must not sink the GPU number
            base = {"value": None, "unit": "k-mers/s", "cores": 0, "kind": "unavailable", "sample": str(e)}

    nan = float("nan")
    hk = h_kmers if h_kmers is not None else 0
    out = {
        "metric": "k-mers/sec for quicKmer2 count", "value": value, "unit": "k-mers/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "value_is": "counting of framed reads already resident in HBM (H2D and record framing excluded); e2e includes both",
        "config": {"workload": args.workload, "desc": w["desc"], "k": int(desc.k), "dict_kmers": n_kmers,
                   "table_MiB": int(desc.table_bytes) >> 20, "reads_per_gpu": data.lines, "kmers_per_step": job_kmers,
                   "hit_fraction": job_hits / max(1, job_kmers), "chunk_MiB": args.chunk_mib,
                   "parallelism": f"reads sharded over {world} GPU(s), dictionary replicated"
                                  + (", NCCL reduce of the counters per step, overlapped with the next step's counting (two counter buffers)" if world > 1 else ""),
                   "l2": "inputs (framed reads + table) exceed the 126 MB L2 every step; no flush needed"
                         if data.n_framed + int(desc.table_bytes) > (256 << 20) else "working set fits L2: HBM term does not bind",
                   "note": data.note},
        "bases_per_s": job_bases / (ms_per_step * 1e-3),
        "e2e": {"value": hk / raw_s, "unit": "k-mers/s", "h2d_bytes_per_step": h_shipped if host_framing else h_raw_bytes,
                "d2h_bytes_per_step": 2 * n_kmers + 64,
                "h2d_gbs": (h_shipped if host_framing else h_raw_bytes) / raw_s / 1e9,
                "frac_of_h2d_peak": ((h_shipped if host_framing else h_raw_bytes) / world / raw_s / 1e9) / h2d if h2d == h2d else None,
                "bases_per_s": h_bases / raw_s,
                "framing": (f"host, {framer_threads} threads per GPU" + (", packed chunks (2-bit codes + reset flags, 24 bytes per 64 positions)" if packed else "")) if host_framing else "device",
                "host_raw_gbs": h_raw_bytes / raw_s / 1e9,
                "host_framer_alone_gbs": None if framer_gbs is None else {"raw_in": framer_gbs[0], "framed_out": framer_gbs[1], "threads": framer_threads},
                "text_chunks": None if args.kernel_only or unpacked_s is None else {
                    "framing": "host, sequence lines shipped as text (QK_PACKED=0)", "value": hk / unpacked_s, "h2d_bytes_per_step": h_framed_bytes},
                "other_framing": None if args.kernel_only or other_s is None else {
                    "framing": "device" if host_framing else f"host, {framer_threads} threads", "value": hk / other_s,
                    "h2d_bytes_per_step": h_raw_bytes if host_framing else h_shipped},
                "path": ("raw FASTA/FASTQ bytes in host memory -> qk_count_mem_mt (host threads frame blocks in parallel, sequence lines only"
                         + (", packed," if packed else "") + " into the pinned slots, H2D, count kernels) -> qk_finish (uint16 depths D2H)"
                         if host_framing else
                         "raw FASTA/FASTQ bytes in pinned host memory -> qk_count_raw_mem (cut at line ends, H2D, device framing, count kernels) -> qk_finish (uint16 depths D2H)"),
                "steps": e2e_steps, "raw_bytes_per_step": h_raw_bytes, "reads_per_step": data.h_lines * world, "kmers_per_step": hk,
                "sample": None if not synth else f"the first {data.h_lines} reads of each GPU's stream ({data.h_lines / max(1, data.lines) * 30:.1f}x of its 30x): "
                                                 "what the host can hold pinned; the rate does not depend on the length of the stream"},
        "e2e_preframed": {"value": hk / pre_s, "unit": "k-mers/s", "h2d_gbs": h_framed_bytes / pre_s / 1e9,
                          "h2d_ms_per_step": h2d_ms_step,
                          "path": "pre-framed pinned host chunks -> qk_submit (H2D + kernel per chunk) -> qk_finish"},
        "e2e_file": None if file_s is None else {
            "value": file_kmers / file_s, "unit": "k-mers/s", "file_gbs": file_bytes / file_s / 1e9,
            "threads": framer_threads if host_framing else args.reader_threads,
            "path": ("reads FILE (page cache, mapped) -> qk_count_file_mt (host threads frame blocks of the mapping, sequence lines into the pinned slots, H2D, count) -> qk_finish"
                     if host_framing else
                     "reads FILE (page cache) -> qk_count_raw_file_mt (pread into pinned slots, H2D, device framing, count) -> qk_finish")},
        "gpu_launches": launches * args.steps,
        "parity_ok": parity["ok"], "parity": parity,
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": base,
        "setup": {"dict_files_s": gen_s, "dict_load_build_s": load_s, "dict_bcast_s": bcast_s, "reads_s": reads_s, "host_frame_s": data.frame_s,
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
    ap.add_argument("--workload", default="config3", choices=sorted(WORKLOADS) + sorted(SYNTH_WORKLOADS))
    ap.add_argument("--reads-scale", type=float, default=1.0, help="GPU-generated workloads: fraction of the workload's reads per GPU (quick runs)")
    ap.add_argument("--e2e-coverage", type=float, default=4.0, help="GPU-generated workloads: the host legs run over at most this coverage (of 30x) in pinned memory")
    ap.add_argument("--framer", default="auto", choices=["auto", "host", "device"],
                    help="e2e leg: record framing by host threads (ships sequence lines only) or on the device (ships the raw stream); "
                         "auto = host when there are >= 12 host CPUs per GPU")
    ap.add_argument("--framer-threads", type=int, default=0, help="host framer threads per GPU (0 = host CPUs / GPUs)")
    ap.add_argument("--reference-budget-s", type=float, default=240.0, help="--impl reference: stop starting new steps after this many seconds")
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
