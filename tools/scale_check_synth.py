#!/usr/bin/env python
"""Bit-exact check against the compiled reference on a GPU-generated workload of bench.py -- by
default config 3: the 3.1 Gb / 2^32-slot human-scale dictionary (48 GiB .qm) and a 1x prefix of its
30x read stream.  Our command and `oracle/_ref/quicKmer2 count` run on the same files; `.bin`
(4.4 GB at human scale) and `.txt` are compared byte for byte.  Shares bench.py's data cache, so a
bench run in the same gpurun call finds the dictionary already written.  Prints one JSON line.
(Test tooling: runs the reference binary, so it lives outside the package.)"""
import argparse
import json
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
CLI = ROOT / "quick-mer2_b200" / "bin" / "quicKmer2_b200"
REF = ROOT / "oracle" / "_ref" / "quicKmer2"


def same_file(a: Path, b: Path, piece=64 << 20):
    if a.stat().st_size != b.stat().st_size:
        return False
    with open(a, "rb") as fa, open(b, "rb") as fb:
        while True:
            x, y = fa.read(piece), fb.read(piece)
            if x != y:
                return False
            if not x:
                return True


SMOOTH_STUB = """#!/usr/bin/env python3
# stand-in for the reference's smooth_GC_mrsfast.py (LOWESS; needs numpy.float and matplotlib, absent in this image):
# 401 float32 on stdout, a fixed curve in the range the real one is clamped to
import math, struct, sys
vals = [min(3.0, max(1 / 3, 1.0 + 0.8 * math.sin(i / 37.0) + (i % 11) * 0.013)) for i in range(401)]
sys.stdout.buffer.write(struct.pack("<401f", *vals))
"""


def est_leg(d: Path, ref: Path, sample: Path, out: dict, args) -> bool:
    """`est ref sample output.bed` (Q.c:555-685) on the .bin / .txt our count just wrote: <ref>.bed with one window per
    --est-window dictionary entries, the Python smoother replaced by a fixed curve on the PATH, our command and the
    reference's on the same files."""
    n = out["bin_entries"]
    w = args.est_window
    t0 = time.time()
    with open(str(ref) + ".bed", "w") as f:
        for a in range(0, n - w, w * 4096):
            f.write("".join(f"chrS\t{x}\t{x + w}\t{x}\t{x + w}\n" for x in range(a, min(n - w, a + w * 4096), w)))
    stub = d / f"stubbin_{os.getpid()}"
    stub.mkdir(exist_ok=True)
    (stub / "smooth_GC_mrsfast.py").write_text(SMOOTH_STUB)
    (stub / "smooth_GC_mrsfast.py").chmod(0o755)
    env = dict(os.environ, PATH=f"{stub}:{os.environ['PATH']}", QK_TIMING="1")
    e = {"windows": (n - w + w - 1) // w, "window_kmers": w, "bed_s": time.time() - t0}
    ok = True
    outs = {}
    for who, exe in (("ours", CLI), ("reference", REF)):
        if who == "reference" and not REF.exists():
            continue
        bed = Path(f"{sample}.{who}.bed")
        t0 = time.time()
        res = subprocess.run([str(exe), "est", str(ref), str(sample), str(bed)], capture_output=True, text=True, env=env)
        e[f"{who}_wall_s"] = time.time() - t0
        e[f"{who}_stdout"] = res.stdout.splitlines()[-3:]
        if res.returncode:
            e[f"{who}_failed"] = (res.stdout + res.stderr)[-1000:]
            ok = False
            continue
        e[f"{who}_lines"] = sum(1 for _ in open(bed))
        outs[who] = bed
    if len(outs) == 2:
        e["bed_identical"] = same_file(outs["ours"], outs["reference"])
        ok = ok and e["bed_identical"]
    for p in outs.values():
        p.unlink(missing_ok=True)
    (stub / "smooth_GC_mrsfast.py").unlink()
    stub.rmdir()
    out["est"] = e
    return ok


def main():
    import bench
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="config3", choices=sorted(bench.SYNTH_WORKLOADS))
    ap.add_argument("--cache-dir", default=None)
    ap.add_argument("--ref-threads", type=int, default=15)
    ap.add_argument("--gpus", default="0", help="device list for our command (-g)")
    ap.add_argument("--skip-reference", action="store_true", help="only our command (timing); no comparison")
    ap.add_argument("--est", action="store_true", help="then `est` on our .bin/.txt: our command against the reference's, output.bed compared")
    ap.add_argument("--est-window", type=int, default=1000, help="k-mers per window of the generated <ref>.bed")
    args = ap.parse_args()
    cdir = bench.cache_dir(args.cache_dir)
    subprocess.run(["make", "-s", "-C", str(ROOT / "quick-mer2_b200"), "all"], check=True)
    w = bench.SYNTH_WORKLOADS[args.workload]
    t0 = time.time()
    d, ref, sample = bench.prepare_synth(args.workload, cdir, sample=True)
    out = {"workload": args.workload, "desc": w["desc"], "data_s": time.time() - t0, "sample_reads": w["sample_reads"],
           "qm_bytes": (d / "ref.fa.qm").stat().st_size, "reads_bytes": sample.stat().st_size}
    if (d / "dict_info.json").exists():
        out["dict"] = json.loads((d / "dict_info.json").read_text())
    ours, theirs = d / f"ours_{os.getpid()}", d / f"theirs_{os.getpid()}"
    t0 = time.time()
    res = subprocess.run([str(CLI), "count", "-t", "12", "-g", args.gpus, str(ref), str(sample), str(ours)], capture_output=True, text=True,
                         env=dict(os.environ, QK_TIMING="1"))
    out["ours_wall_s"] = time.time() - t0
    if res.returncode:
        sys.exit(f"our command failed\n{res.stdout[-2000:]}\n{res.stderr[-2000:]}")
    out["ours"] = json.loads(res.stderr.strip().splitlines()[-1])
    out["ours_timing"] = [l for l in res.stderr.splitlines() if l.startswith("[qk]")]
    out["ours_stdout"] = [l for l in res.stdout.splitlines() if "elapse" in l or "depth" in l or l.endswith("G kmers")]
    ok = True
    if not args.skip_reference:
        t0 = time.time()
        res = subprocess.run([str(REF), "count", "-t", str(args.ref_threads), str(ref), str(sample), str(theirs)], capture_output=True, text=True)
        out["reference_wall_s"] = time.time() - t0
        if res.returncode:
            sys.exit(f"reference failed\n{res.stdout[-2000:]}")
        out["reference_stdout"] = [l for l in res.stdout.splitlines() if "elapse" in l or "depth" in l or l.endswith("G kmers")]
        out["bin_identical"] = same_file(Path(str(ours) + ".bin"), Path(str(theirs) + ".bin"))
        out["txt_identical"] = same_file(Path(str(ours) + ".txt"), Path(str(theirs) + ".txt"))
        out["stdout_lines_identical"] = [l for l in out["ours_stdout"] if "elapse" not in l] == [l for l in out["reference_stdout"] if "elapse" not in l]
        ok = out["bin_identical"] and out["txt_identical"]
    out["bin_entries"] = Path(str(ours) + ".bin").stat().st_size // 2
    if args.est:
        ok = est_leg(d, ref, ours, out, args) and ok
    for p in (ours, theirs):
        for ext in (".bin", ".txt"):
            Path(str(p) + ext).unlink(missing_ok=True)
    print(json.dumps(out))
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
