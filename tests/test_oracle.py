"""Pin the oracle (oracle/qk_oracle.c) against the reference.

The reference ships no tests or golden vectors for `count`.  The fixtures
under tests/golden/ were produced by the UNMODIFIED reference compiled from
/root/reference/QuicKmer.c (tests/golden/make_golden.py); the oracle must reproduce every
one of them byte for byte.  Where the compiled reference is present (oracle/_ref/quicKmer2:
the authoring container, and the GPU box because the binary travels with the snapshot) the
oracle is also checked against live runs on fresh seeded inputs.
"""
import subprocess

import numpy as np
import pytest

import os

from conftest import GOLDEN, ROOT, golden_cases, golden_meta, weird_stream


@pytest.mark.parametrize("case", golden_cases())
def test_oracle_matches_reference_golden(case, oracle, tmp_path):
    meta = golden_meta(case)
    d = GOLDEN / case
    st = oracle.count(d / "ref.fa", d / meta["reads"], tmp_path / "out")
    assert (tmp_path / "out.bin").read_bytes() == (d / "expect.bin").read_bytes()
    assert st["total_kmers"] == meta["total_kmers"]          # Q.c:481 "total %lu kmers"
    assert st["fastq"] == int(meta["fastq"])
    assert st["undefined_lines"] == 0                        # fixtures stay inside defined behaviour
    if meta["has_txt"]:
        assert (tmp_path / "out.txt").read_bytes() == (d / "expect.txt").read_bytes()
    else:
        assert not (tmp_path / "out.txt").exists()


def test_golden_cover_the_known_answer_matrix():
    """k sweep, FASTA+FASTQ, threaded+unthreaded, long lines, counter wrap."""
    metas = [golden_meta(c) for c in golden_cases()]
    assert {m["k"] for m in metas} >= {3, 12, 20, 25, 30, 31}
    assert any(m["fastq"] for m in metas) and any(not m["fastq"] for m in metas)
    assert any(m["threads"] for m in metas) and any(not m["threads"] for m in metas)
    assert any(m["has_txt"] for m in metas)
    wrap = np.fromfile(GOLDEN / "k30_wrap_t2" / "expect.bin", dtype=np.uint16)
    reads = (GOLDEN / "k30_wrap_t2" / "reads.fa").read_text()
    # T12: (AC)^49999 twice = 2 * 49,985 windows of (AC)^15 or (CA)^15 >= 65,536 hits on one
    # dictionary k-mer, so its depth must have wrapped; no entry may read 65,535 (no saturation)
    assert reads.count("AC" * 49999) == 2
    assert wrap.max() < 65535
    long_reads = [len(l) for l in (GOLDEN / "k30_long_lines" / "reads.fa").read_text().split("\n")]
    assert {65535, 65536, 70000, 99998} <= set(long_reads)   # T7/T8


def test_djb_known_values(oracle):
    # Q.c:66-76: h = 5381; h = h*33 + byte, 8 bytes LSB first, u64 wrap
    def djb(key):
        h = 5381
        for b in range(8):
            h = (h * 33 + ((key >> (8 * b)) & 0xFF)) & 0xFFFFFFFFFFFFFFFF
        return h
    for key in (0, 1, 0xFF, 0x0123456789ABCDEF, (1 << 60) - 1, 0xFFFFFFFFFFFFFFFF):
        assert oracle.djb(key) == djb(key)


def py_codec(k, line: bytes):
    """Literal transcription of SURVEY.md Appendix A process(line) -- a third, independent statement."""
    mask = (1 << (2 * k)) - 1 if k < 32 else 0
    fwd = rc = cur = 0
    out = []
    for c in line:
        if c == 0x4E:
            fwd = rc = cur = 0
            continue
        cur = (cur + 1) & 0xFFFF
        L = (c >> 1) & 3
        fwd = ((fwd << 2) | L) & 0xFFFFFFFFFFFFFFFF
        rc = (rc | (((L - 2) & 3) << 60)) >> 2
        if cur >= k:
            key = fwd & mask
            out.append(min(key, rc))
    return out


@pytest.mark.parametrize("k", [3, 12, 20, 25, 30, 31, 32])
def test_codec_restatement(k, oracle):
    rng = np.random.default_rng(k)
    lines = [bytes(rng.choice(np.frombuffer(b"ACGTacgtNn\r-", dtype=np.uint8), size=int(n)))
             for n in (0, 1, k - 1, k, k + 1, 64, 150, 151, 1000)]
    chunk = b"".join(l + b"\n" for l in lines)
    want = [key for l in lines for key in py_codec(k, l)]
    got = oracle.chunk_keys(k, chunk)
    assert got.tolist() == want


def test_run_counter_wraps_at_65536(oracle):
    # T7: after 65,536 non-N bytes cur_chars is 0 and the next k-1 positions emit nothing
    k = 30
    line = b"ACGT" * 17000  # 68,000 bases
    got = oracle.chunk_keys(k, line + b"\n").size
    assert got == (65535 - k + 1) + (68000 - 65536 - k + 1)


def test_oracle_against_live_reference(oracle, ref_binary, synth, tmp_path):
    if ref_binary is None:
        pytest.skip("oracle/_ref/quicKmer2 not built here")
    synth("ref", "--out", tmp_path / "ref.fa", "--bases", 200000, "--contigs", 3, "--seed", 77, "--segdups", 4,
          "--segdup-len", 3000, "--nblock", 500)
    synth("ctrl", "--ref", tmp_path / "ref.fa", "--out", tmp_path / "ctrl.bed", "--block", 5000)
    res = subprocess.run([str(ref_binary), "search", "-k", "30", "-e", "0", "-s", "1M", "-c", "ctrl.bed", "ref.fa"],
                         cwd=tmp_path, capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    for name, extra in (("reads.fq", ["--fastq", "--rand-qual"]), ("reads.fa", ["--lower-ppm", 100000])):
        synth("reads", "--ref", tmp_path / "ref.fa", "--out", tmp_path / name, "--n", 20000, "--len", 150, "--seed", 5,
              *extra)
        res = subprocess.run([str(ref_binary), "count", "-t", "3", "ref.fa", name, "live"], cwd=tmp_path,
                             capture_output=True, text=True)
        assert res.returncode == 0, res.stdout + res.stderr
        st = oracle.count(tmp_path / "ref.fa", tmp_path / name, tmp_path / "port")
        assert (tmp_path / "port.bin").read_bytes() == (tmp_path / "live.bin").read_bytes()
        assert (tmp_path / "port.txt").read_bytes() == (tmp_path / "live.txt").read_bytes()
        assert f"total {st['total_kmers']} kmers" in res.stdout
        assert st["hits"] > 0.5 * st["total_kmers"]


def test_reference_on_a_pipe_loses_the_first_line(oracle, ref_binary, tmp_path):
    """Q.c:396: on FASTA input the reference rewinds with fseek(0), which fails on a pipe, so the
    first line is consumed (README.md:89-90 is the documented pipe use).  Pin that against the
    reference itself: piping X must give what the file 'X minus its first line' gives."""
    if ref_binary is None:
        pytest.skip("oracle/_ref/quicKmer2 not built here")
    d = GOLDEN / "k30_fasta_t0"
    seq = [l for l in (d / "reads.fa").read_text().split("\n") if l and not l.startswith(">")]
    data = ("\n".join(seq[:200]) + "\n").encode()          # headerless: the first line is a read
    for name in ("ref.fa.qm", "ref.fa.qgc"):
        (tmp_path / name).write_bytes((d / name).read_bytes())
    res = subprocess.run([str(ref_binary), "count", "ref.fa", "/dev/stdin", "piped"], cwd=tmp_path, input=data,
                         capture_output=True)
    assert res.returncode == 0, res.stdout
    (tmp_path / "all.fa").write_bytes(data)
    (tmp_path / "tail.fa").write_bytes(data[data.index(b"\n") + 1:])
    oracle.count(tmp_path / "ref.fa", tmp_path / "tail.fa", tmp_path / "want_tail")
    oracle.count(tmp_path / "ref.fa", tmp_path / "all.fa", tmp_path / "want_all")
    piped = (tmp_path / "piped.bin").read_bytes()
    assert piped == (tmp_path / "want_tail.bin").read_bytes()
    assert piped != (tmp_path / "want_all.bin").read_bytes()   # the lost line did carry dictionary k-mers


@pytest.mark.parametrize("fastq_like", [True, False])
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_line_state_machine_against_live_reference(seed, fastq_like, oracle, ref_binary, tmp_path):
    """The GPU framing tests compare with the oracle on streams that are NOT valid FASTA/FASTQ
    ('>' where a read is expected, quality lines starting with '>' or '@', empty lines ...).
    Here the oracle's reading of those streams is pinned to the reference binary itself."""
    if ref_binary is None:
        pytest.skip("oracle/_ref/quicKmer2 not built here")
    d = GOLDEN / "k30_fasta_t0"
    seq = "".join(l.strip() for l in open(d / "ref.fa") if not l.startswith(">")).replace("N", "")
    rng = np.random.default_rng(100 + seed)
    text = weird_stream(rng, seq, fastq_like, 3000)
    (tmp_path / "w.txt").write_text(text, newline="")
    for name in ("ref.fa.qm", "ref.fa.qgc"):
        (tmp_path / name).write_bytes((d / name).read_bytes())
    for threads in ([], ["-t", "2"]):
        res = subprocess.run([str(ref_binary), "count", *threads, "ref.fa", "w.txt", "live"], cwd=tmp_path,
                             capture_output=True, text=True)
        assert res.returncode == 0, res.stdout + res.stderr
        st = oracle.count(tmp_path / "ref.fa", tmp_path / "w.txt", tmp_path / "port")
        assert (tmp_path / "port.bin").read_bytes() == (tmp_path / "live.bin").read_bytes()
        assert (tmp_path / "port.txt").read_bytes() == (tmp_path / "live.txt").read_bytes()
        assert f"total {st['total_kmers']} kmers" in res.stdout
    assert st["fastq"] == int(fastq_like) and st["lines"] > 300


@pytest.mark.parametrize("k", [20, 30])
def test_synthetic_dictionary_equals_search(k, oracle, ref_binary, synth, tmp_path):
    """`qk_synth dict` (used where the reference binary cannot travel) must describe the same
    dictionary as the reference's `search -e 0`: same unique k-mers in the same chain order, hence
    the same .bin for any reads (slot placement may differ -- it is not observable)."""
    if ref_binary is None:
        pytest.skip("oracle/_ref/quicKmer2 not built here")
    a, b = tmp_path / "a", tmp_path / "b"
    a.mkdir(); b.mkdir()
    synth("ref", "--out", a / "ref.fa", "--bases", 300000, "--contigs", 3, "--seed", 5 + k, "--segdups", 6, "--segdup-len", 4000,
          "--nblock", 700)
    (b / "ref.fa").write_bytes((a / "ref.fa").read_bytes())
    res = subprocess.run([str(ref_binary), "search", "-k", str(k), "-e", "0", "-s", "1M", "ref.fa"], cwd=a, capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    synth("dict", "--ref", b / "ref.fa", "--k", k, "--slots", "1M")
    synth("reads", "--ref", a / "ref.fa", "--out", tmp_path / "r.fa", "--n", 30000, "--len", 150, "--seed", 3)
    bin_a, sa = oracle.count_bin(a / "ref.fa.qm", tmp_path / "r.fa")
    bin_b, sb = oracle.count_bin(b / "ref.fa.qm", tmp_path / "r.fa")
    assert bin_a.size == bin_b.size > 200000
    assert np.array_equal(bin_a, bin_b) and sa["hits"] == sb["hits"] > 0


def test_chain_is_all_that_count_reads_of_a_dictionary(oracle, ref_binary, synth, tmp_path):
    """Two dictionaries `search` never writes but `count` accepts: occupied slots that are not on the chain (what `sparse`
    leaves when it does not resize), and empty slots ON the chain (key 0, the poly-A k-mer of an `index` list) -- Find_hash(0)
    "finds" the first empty slot on its path (Q.c:98), so that one collects the poly-A / poly-T k-mers.  The live
    reference and the restatement must agree on both; the CUDA path is held to the same files in
    tests/test_gpu_parity.py::test_occupied_slots_off_the_chain / test_empty_slot_on_the_chain."""
    if ref_binary is None:
        pytest.skip("oracle/_ref/quicKmer2 not built here")
    synth("ref", "--out", tmp_path / "ref.fa", "--bases", 150000, "--contigs", 2, "--seed", 21, "--segdups", 3, "--segdup-len", 2000)
    res = subprocess.run([str(ref_binary), "search", "-k", "30", "-e", "0", "-s", "1M", "ref.fa"], cwd=tmp_path, capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    raw = (tmp_path / "ref.fa.qm").read_bytes()
    H, first = int.from_bytes(raw[8:16], "little"), int.from_bytes(raw[16:24], "little")
    keys = np.frombuffer(raw, dtype="<u8", count=H, offset=24).copy()
    nxt = np.frombuffer(raw, dtype="<u4", count=H, offset=24 + 8 * H).copy()
    slots, c = [], first
    while True:
        slots.append(c)
        c = int(nxt[c])
        if c == first:
            break
    keep = [s for i, s in enumerate(slots) if i % 3 != 1 or i == 0]
    for x, y in zip(keep, keep[1:] + keep[:1]):
        nxt[x] = y
    c = oracle.djb(0) & (H - 1)
    step = -1 if c & (H >> 1) else 1
    while keys[c] != 0:
        c += step
    found, other = c, int(np.flatnonzero(keys == 0)[-1])
    for e, i in ((found, 700), (other, 3)):
        nxt[e] = nxt[keep[i]]
        nxt[keep[i]] = e
    synth("reads", "--ref", tmp_path / "ref.fa", "--out", tmp_path / "r.fa", "--n", 8000, "--len", 150, "--seed", 4)
    with open(tmp_path / "r.fa", "a") as f:
        f.write(">polyA\n" + "A" * 70 + "\n>polyT\n" + "T" * 45 + "\n")
    (tmp_path / "ref.fa.qm").write_bytes(raw[:24] + keys.tobytes() + nxt.tobytes())
    want, st = oracle.count_bin(tmp_path / "ref.fa.qm", tmp_path / "r.fa")
    assert want.size == len(keep) + 2 < len(slots)
    order, c = [], first
    for _ in range(want.size):
        order.append(c)
        c = int(nxt[c])
    assert want[order.index(found)] == 41 + 16 and want[order.index(other)] == 0 and want.sum() > 100000
    res = subprocess.run([str(ref_binary), "count", "ref.fa", "r.fa", "live"], cwd=tmp_path, capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert (tmp_path / "live.bin").read_bytes() == want.tobytes()
    # -t N: the last batch of 4,096 keys is filled up with zeros, and those are looked up too (Q.c:458-466)
    padded, _ = oracle.count_bin(tmp_path / "ref.fa.qm", tmp_path / "r.fa", threads=3)
    pad = 4096 - st["total_kmers"] % 4096
    assert padded[order.index(found)] == 41 + 16 + pad and (padded != want).sum() == 1
    for t in ("1", "3"):
        res = subprocess.run([str(ref_binary), "count", "-t", t, "ref.fa", "r.fa", "live"], cwd=tmp_path, capture_output=True, text=True)
        assert res.returncode == 0, res.stdout + res.stderr
        assert (tmp_path / "live.bin").read_bytes() == padded.tobytes(), t
    st2 = oracle.count(tmp_path / "ref.fa", tmp_path / "r.fa", tmp_path / "port", threads=3)
    assert (tmp_path / "port.bin").read_bytes() == padded.tobytes() and st2["total_kmers"] == st["total_kmers"]


@pytest.mark.parametrize("k,size", [(30, "4M"), (30, "256K"), (20, "256K"), (12, "1M"), (31, "512K")])
def test_search_pass1_restatement_against_live_reference(k, size, oracle, ref_binary, synth, tmp_path):
    """SURVEY 8(f) rank 4, oracle first: `search` pass 1 (hash_from_fasta, Q.c:824-923) restated -- sequences that run
    over their lines, no 16-bit wrap, key 0 left out, occurrences capped at 255, and the in-place re-probing when the
    table doubles -- against the two figures the reference prints after it ("Uniq count U, total T"), with the table
    large enough from the start and with two or three doublings on the way."""
    import re
    if ref_binary is None:
        pytest.skip("oracle/_ref/quicKmer2 not built here")
    synth("ref", "--out", tmp_path / "ref.fa", "--bases", 900000, "--contigs", 4, "--seed", 100 + k, "--segdups", 12, "--segdup-len", 6000,
          "--nblock", 1500)
    with open(tmp_path / "ref.fa", "a") as f:
        f.write(">polyA_and_a_long_header " + "x" * 300 + "\n" + "A" * 150 + "\n" + "ACGT" * 40 + "\n")
    res = subprocess.run([str(ref_binary), "search", "-k", str(k), "-e", "0", "-s", size, "ref.fa"], cwd=tmp_path, capture_output=True, text=True,
                         errors="replace")
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr
    m = re.search(r"Uniq count (\d+), total (\d+)", res.stdout)
    assert m, res.stdout[-2000:]
    slots = int(size[:-1]) << (20 if size.endswith("M") else 10)
    p1 = oracle.search_pass1(tmp_path / "ref.fa", k, slots)
    assert (p1["unique"], p1["distinct"]) == (int(m.group(1)), int(m.group(2)))
    assert (p1["resizes"] > 0) == (size == "256K" or (size, k) == ("512K", 31))
    grown = oracle.search_pass1(tmp_path / "ref.fa", k, 4096)           # eight or nine doublings: the sweep loses no key
    assert (grown["unique"], grown["distinct"]) == (p1["unique"], p1["distinct"]) and grown["resizes"] >= 8
    occupied = p1["keys"] != 0
    assert occupied.sum() == p1["distinct"] and (p1["occ"][~occupied] == 0).all() and p1["occ"][occupied].min() >= 1
    if k >= 20:                                   # duplicated stretches: some k-mers occur more than once
        assert 0.5 * p1["distinct"] < p1["unique"] < p1["distinct"]
    # the dictionary `search -e 0` goes on to write holds exactly the k-mers that occur once
    raw = (tmp_path / "ref.fa.qm").read_bytes()
    H = int.from_bytes(raw[8:16], "little")
    written = np.frombuffer(raw, dtype="<u8", count=H, offset=24)
    assert np.array_equal(np.sort(written[written != 0]), np.sort(p1["keys"][p1["occ"] == 1]))


SMOOTH_STUB = """#!/usr/bin/env python3
# stand-in for the reference's smooth_GC_mrsfast.py (LOWESS; needs numpy.float and matplotlib, absent here):
# 401 float32 on stdout, a deterministic curve with the same range the real one has (clamped to [1/3, 3])
import math, struct, sys
vals = [min(3.0, max(1 / 3, 1.0 + 0.8 * math.sin(i / 37.0) + (i % 11) * 0.013)) for i in range(401)]
sys.stdout.buffer.write(struct.pack("<401f", *vals))
"""


def write_smooth_stub(d):
    """smooth_GC_mrsfast.py on a PATH directory; returns (env for subprocesses, the curve as a file for qk_oracle est)."""
    b = d / "stubbin"
    b.mkdir(exist_ok=True)
    (b / "smooth_GC_mrsfast.py").write_text(SMOOTH_STUB)
    (b / "smooth_GC_mrsfast.py").chmod(0o755)
    curve = subprocess.run([str(b / "smooth_GC_mrsfast.py")], capture_output=True, check=True).stdout
    (d / "curve.f32").write_bytes(curve)
    return dict(os.environ, PATH=f"{b}:{os.environ['PATH']}"), d / "curve.f32"


def test_est_restatement_equals_the_reference(oracle, ref_binary, synth, tmp_path):
    """qko_est (oracle) against the compiled reference's `est` (Q.c:555-685) on the reference's own search /
    count outputs, the Python smoother replaced by a stub that prints a fixed curve: output.bed byte for byte."""
    if ref_binary is None:
        pytest.skip("compiled reference not available")
    synth("ref", "--out", tmp_path / "ref.fa", "--bases", 400000, "--contigs", 2, "--seed", 12, "--segdups", 3, "--segdup-len", 3000,
          "--nblock", 500)
    synth("ctrl", "--ref", tmp_path / "ref.fa", "--out", tmp_path / "ctrl.bed", "--block", 20000)
    run = lambda *a, **kw: subprocess.run([str(ref_binary), *map(str, a)], cwd=tmp_path, capture_output=True, text=True, **kw)
    assert run("search", "-k", 30, "-e", 0, "-s", "1M", "-w", 700, "-c", "ctrl.bed", "ref.fa").returncode == 0
    synth("reads", "--ref", tmp_path / "ref.fa", "--out", tmp_path / "r.fq", "--n", 60000, "--len", 150, "--seed", 4, "--fastq")
    assert run("count", "-t", 3, "ref.fa", "r.fq", "samp").returncode == 0
    env, curve = write_smooth_stub(tmp_path)
    res = run("est", "ref.fa", "samp", "theirs.bed", env=env)
    assert res.returncode == 0 and "Mean sequencing depth" in res.stdout
    port = ROOT / "oracle" / "_build" / "qk_oracle"
    assert subprocess.run([str(port), "est", "ref.fa", "samp", "port.bed", str(curve)], cwd=tmp_path).returncode == 0
    theirs = (tmp_path / "theirs.bed").read_bytes()
    assert theirs == (tmp_path / "port.bed").read_bytes() and theirs.count(b"\n") > 500
