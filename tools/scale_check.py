#!/usr/bin/env python
"""Bit-exact check against the compiled reference at a larger scale than the unit tests:
a 512 Mb (default) synthetic reference, a 2^30-slot QM11 dictionary (12 GiB file, ~0.5 G
k-mers, 8 GiB device table), a few million reads; our CLI and `oracle/_ref/quicKmer2 count`
on the same files, `.bin` and `.txt` compared byte for byte.  Prints one JSON line.
(Test tooling: runs the reference binary, so it lives outside the package.)"""
import argparse
import json
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
SYNTH = ROOT / "quick-mer2_b200" / "bin" / "qk_synth"
CLI = ROOT / "quick-mer2_b200" / "bin" / "quicKmer2_b200"
REF = ROOT / "oracle" / "_ref" / "quicKmer2"


def run(cmd, cwd, **kw):
    t0 = time.time()
    res = subprocess.run([str(c) for c in cmd], cwd=cwd, capture_output=True, text=True, **kw)
    if res.returncode:
        sys.exit(f"FAILED {cmd}\n{res.stdout[-2000:]}\n{res.stderr[-2000:]}")
    return res, time.time() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dir", default="/dev/shm/qk_scale")
    ap.add_argument("--bases", default="512M")
    ap.add_argument("--slots", default="1G")
    ap.add_argument("--reads", default="4M")
    ap.add_argument("--ref-threads", type=int, default=8)
    args = ap.parse_args()
    d = Path(args.dir)
    d.mkdir(parents=True, exist_ok=True)
    out = {"bases": args.bases, "slots": args.slots, "reads": args.reads}
    _, out["gen_ref_s"] = run([SYNTH, "ref", "--out", "ref.fa", "--bases", args.bases, "--contigs", 4, "--seed", 31, "--segdups", 400,
                               "--segdup-len", 20000, "--nblock", 100000], d)
    res, out["gen_dict_s"] = run([SYNTH, "dict", "--ref", "ref.fa", "--k", 30, "--slots", args.slots, "--ctrl-block", 100000,
                                  "--threads", 16], d)
    out["dict"] = json.loads(res.stdout.strip().splitlines()[-1])
    _, out["gen_reads_s"] = run([SYNTH, "reads", "--ref", "ref.fa", "--out", "reads.fq", "--n", args.reads, "--len", 150, "--seed", 9,
                                 "--fastq"], d)
    import os
    res, out["ours_wall_s"] = run([CLI, "count", "-t", 12, "ref.fa", "reads.fq", "ours"], d,
                                  env=dict(os.environ, QK_TIMING="1", QK_READER_THREADS="8"))
    out["ours"] = json.loads(res.stderr.strip().splitlines()[-1])
    out["ours_timing"] = [l for l in res.stderr.splitlines() if l.startswith("[qk]")]
    res, out["reference_wall_s"] = run([REF, "count", "-t", args.ref_threads, "ref.fa", "reads.fq", "theirs"], d)
    out["reference_stdout"] = [l for l in res.stdout.splitlines() if "elapse" in l or "depth" in l]
    out["bin_identical"] = (d / "ours.bin").read_bytes() == (d / "theirs.bin").read_bytes()
    out["txt_identical"] = (d / "ours.txt").read_bytes() == (d / "theirs.txt").read_bytes()
    out["bin_entries"] = (d / "ours.bin").stat().st_size // 2
    print(json.dumps(out))
    sys.exit(0 if out["bin_identical"] and out["txt_identical"] else 1)


if __name__ == "__main__":
    main()
