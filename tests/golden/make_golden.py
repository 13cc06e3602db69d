#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the authoring container (needs /root/reference, compiled to oracle/_ref/quicKmer2 by
`make -C oracle ref`).  Each case directory gets:

    ref.fa                reference FASTA (seeded, qk_synth)
    ctrl.bed              control regions given to `search -c` (when the case has a .qgc)
    ref.fa.qm / .qgc      dictionary written by `quicKmer2 search -e 0`   (reference output)
    reads.fa | reads.fq   reads (seeded, qk_synth, plus hand-made edge cases)
    expect.bin / .txt     written by `quicKmer2 count -t <T>`             (reference output)
    meta.json             commands, k, number of k-mers, "total ... kmers" of the reference

The reference has no tests or golden vectors of its own; these fixtures are
what pins the oracle (tests/test_oracle.py) and, through it, the CUDA path (tests/test_gpu_*).
Cases follow the known-answer matrix of SURVEY.md 4.3 (T1-T15).
"""
import json
import os
import random
import re
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = ROOT / "oracle" / "_ref" / "quicKmer2"
SYNTH = ROOT / "quick-mer2_b200" / "bin" / "qk_synth"


def run(cmd, cwd, **kw):
    res = subprocess.run([str(c) for c in cmd], cwd=cwd, capture_output=True, text=True, **kw)
    if res.returncode != 0:
        sys.exit(f"FAILED {cmd}\n{res.stdout}\n{res.stderr}")
    return res.stdout


def read_fasta_seq(path):
    return "".join(l.strip() for l in open(path) if not l.startswith(">"))


def revcomp(s):
    return s[::-1].translate(str.maketrans("ACGTacgtN", "TGCAtgcaN"))


def edge_reads(seq, rng, k):
    """Hand-made lines for T4-T7, T10, T13 on top of the synthetic reads."""
    out = []
    def take(n):
        n = min(n, len(seq) - 1)
        a = rng.randrange(0, len(seq) - n)
        return seq[a:a + n]

    r = take(150)
    out.append(r[:70] + "N" + r[71:])                    # T4: N resets
    out.append(r[:40] + "NNNN" + r[44:100] + "N")        # N runs, N at line end
    out.append(take(150).lower())                        # T5: lower case
    r = take(150)
    out.append(r[:60].lower() + "n" + r[61:])            # T5: 'n' is NOT a reset (encodes as G)
    out.append(take(150) + "\r")                         # T6: CR counted as a base
    out.append("")                                       # empty line
    out.append(take(k - 1))                              # shorter than k: no k-mer
    out.append(take(k))                                  # exactly k: one k-mer
    out.append("A" * 200)                                # T13: key 0
    out.append("T" * 200)
    out.append(revcomp(take(150)))
    out.append(take(150).replace("A", "R", 1))           # IUPAC byte: encoded through (c>>1)&3
    out.append(take(75) + "\t " + take(75))              # arbitrary bytes are bases
    out.append(take(300)[:149] + "\n" + take(200))       # T10: multi-line record = independent reads
    return out


def write_reads(path, names_seqs, fastq, rng):
    with open(path, "w", newline="") as f:
        for i, s in enumerate(names_seqs):
            if fastq:
                body = s.replace("\n", "")
                # quality strings starting with '@' or '>' are legal FASTQ
                q0 = rng.choice("@>I5#")
                f.write(f"@e{i}\n{body}\n+\n{q0}{'I' * max(0, len(body) - 1)}\n")
            else:
                f.write(f">e{i}\n{s}\n")                  # an embedded newline = multi-line record (T10)


def make_case(name, *, k, bases, contigs=1, seed=1, slots="16K", ctrl=False, fastq=False, n_reads=300, read_len=150,
              threads=0, long_lines=(), wrap=False, err_ppm=2000):
    d = HERE / name
    if d.exists():
        shutil.rmtree(d)
    d.mkdir(parents=True)
    rng = random.Random(seed * 7919 + k)
    run([SYNTH, "ref", "--out", "ref.fa", "--bases", bases, "--contigs", contigs, "--seed", seed, "--nblock",
         40 if bases >= 2000 else 0], d)
    if wrap:  # plant one copy of the period-2 30-mer (AC)^15 so that a read of ACAC... hits it ~50k times
        lines = (d / "ref.fa").read_text().split("\n")
        lines[10] = "G" + "AC" * 15 + "G" + lines[10][32:]
        (d / "ref.fa").write_text("\n".join(lines))
    search = [REF, "search", "-k", k, "-e", 0, "-s", slots]
    if ctrl:
        run([SYNTH, "ctrl", "--ref", "ref.fa", "--out", "ctrl.bed", "--block", 500], d)
        search += ["-c", "ctrl.bed"]
    search_out = run(search + ["ref.fa"], d)
    reads = "reads.fq" if fastq else "reads.fa"
    run([SYNTH, "reads", "--ref", "ref.fa", "--out", "synth." + reads, "--n", n_reads, "--len", read_len, "--seed",
         seed + 41, "--err-ppm", err_ppm] + (["--fastq", "--rand-qual"] if fastq else []), d)
    seq = read_fasta_seq(d / "ref.fa").replace("N", "")
    extra = edge_reads(seq, rng, k)
    for L in long_lines:                                  # T7: uint16 run counter wraps at 65,536
        reps = L // len(seq) + 2
        s = (seq * reps)[:L]
        extra.append(s)
    if wrap:                                              # T12: depth wraps at 65,536, does not saturate
        extra += ["AC" * 49999] * 2                       # ~100k hits on one k-mer => count mod 65,536
    write_reads(d / ("edge." + reads), extra, fastq, rng)
    with open(d / reads, "wb") as out:
        out.write((d / ("synth." + reads)).read_bytes())
        out.write((d / ("edge." + reads)).read_bytes())
    (d / ("synth." + reads)).unlink()
    (d / ("edge." + reads)).unlink()
    cmd = [REF, "count"] + (["-t", threads] if threads else []) + ["ref.fa", reads, "expect"]
    count_out = run(cmd, d)
    m = re.search(r"total (\d+) kmers", count_out)
    meta = {
        "case": name, "k": k, "reads": reads, "fastq": fastq, "threads": threads,
        "search_cmd": " ".join(str(c) for c in search[1:] + ["ref.fa"]),
        "count_cmd": " ".join(str(c) for c in cmd[1:]),
        "total_kmers": int(m.group(1)),
        "n_kmers": (d / "expect.bin").stat().st_size // 2,
        "has_txt": (d / "expect.txt").exists(),
        "reference_stdout": [l for l in count_out.splitlines() if not l.startswith("[Option]")],
    }
    (d / "meta.json").write_text(json.dumps(meta, indent=1) + "\n")
    (d / "ref.fa.bed").unlink(missing_ok=True)           # window file: not an input of count
    size = sum(p.stat().st_size for p in d.iterdir())
    print(f"{name:24s} k={k:2d} n_kmers={meta['n_kmers']:6d} total={meta['total_kmers']:8d} {size / 1024:7.0f} KiB")


def main():
    if not REF.exists():
        sys.exit("oracle/_ref/quicKmer2 missing: run `make -C oracle ref` in the authoring container")
    subprocess.run(["make", "-s", "-C", str(ROOT / "quick-mer2_b200"), str(SYNTH)], check=True)
    # T1/T11/T15: plain FASTA and FASTQ, with a .qgc so that .txt is produced; threaded + unthreaded
    make_case("k30_fasta_t0", k=30, bases=6000, contigs=2, seed=1, ctrl=True)
    make_case("k30_fastq_t3", k=30, bases=6000, contigs=2, seed=2, ctrl=True, fastq=True, threads=3)
    # T2: the 60-bit reverse-complement register is independent of k
    make_case("k12_fasta", k=12, bases=3000, seed=3, n_reads=150)
    make_case("k20_fasta", k=20, bases=3000, seed=4, n_reads=150, ctrl=True)
    make_case("k25_fastq", k=25, bases=3000, seed=5, n_reads=150, fastq=True)
    make_case("k31_fasta", k=31, bases=3000, seed=6, n_reads=150)
    make_case("k3_fasta", k=3, bases=40, seed=9, slots="256", n_reads=20, read_len=30, err_ppm=0)
    # T7: lines of 65,535 / 65,536 / 70,000 / 99,998 bases; T12: counter wrap
    make_case("k30_long_lines", k=30, bases=3000, seed=7, n_reads=20, long_lines=(65535, 65536, 70000, 99998))
    make_case("k30_wrap_t2", k=30, bases=3000, seed=8, n_reads=20, wrap=True, threads=2)


if __name__ == "__main__":
    main()
