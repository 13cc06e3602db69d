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


def main():
    import bench
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="config3", choices=sorted(bench.SYNTH_WORKLOADS))
    ap.add_argument("--cache-dir", default=None)
    ap.add_argument("--ref-threads", type=int, default=15)
    ap.add_argument("--gpus", default="0", help="device list for our command (-g)")
    ap.add_argument("--skip-reference", action="store_true", help="only our command (timing); no comparison")
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
    for p in (ours, theirs):
        for ext in (".bin", ".txt"):
            Path(str(p) + ext).unlink(missing_ok=True)
    print(json.dumps(out))
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
