"""Host-side C code (no GPU needed): the C-ABI library loads and exports every symbol the
headers declare, the record framer reproduces the reference's fgets loop (Q.c:393-398,
451-455), the QM11 header reader and the .bin/.txt writers produce the reference's bytes."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden_cases, golden_meta


def declared_symbols():
    names = set()
    for h in (ROOT / "include").glob("*.h"):
        text = re.sub(r"/\*.*?\*/", "", h.read_text(), flags=re.S)
        names |= set(re.findall(r"\b(qk_[a-z0-9_]+)\s*\(", text))
    return names


def test_library_exports_every_declared_symbol(qk):
    lib = qk.lib()
    declared = declared_symbols()
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/*.h but not exported"
    # and the ctypes table of the Python face covers them all
    assert declared <= set(qk.SIGNATURES), declared - set(qk.SIGNATURES)
    out = subprocess.run(["nm", "-D", "--defined-only", str(qk.LIB_PATH)], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (qk_\w+)", out))
    assert declared <= exported


def test_library_is_sm100a_only(qk):
    out = subprocess.run(["cuobjdump", "-lelf", str(qk.LIB_PATH)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_a_device(qk):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert qk.lib().qk_device_count() <= 0
    with pytest.raises(qk.QkError) as e:
        qk.Context()
    assert e.value.code == 1  # QK_ERR_CUDA


def test_product_does_not_reference_the_oracle():
    """The oracle is the checker: nothing under the package may import, link or run it."""
    for p in (ROOT / "quick-mer2_b200").rglob("*"):
        if p.is_file() and p.suffix in {".py", ".c", ".cu", ".cuh", ".h", ""} and "build" not in p.parts \
                and "bin" not in p.parts and p.name != "Makefile":
            assert "oracle" not in p.read_text(errors="ignore").lower(), p
    assert "oracle" not in (ROOT / "quick-mer2_b200" / "Makefile").read_text().lower()
    out = subprocess.run(["ldd", str(ROOT / "quick-mer2_b200" / "libquickmer2_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


# ---------------------------------------------------------------------------- framer -----
@pytest.mark.parametrize("case", golden_cases())
@pytest.mark.parametrize("cap", [100000, 1 << 20])
def test_framer_matches_oracle_on_golden_reads(case, cap, qk, oracle):
    meta = golden_meta(case)
    path = GOLDEN / case / meta["reads"]
    want, ost = oracle.frame_file(path)
    chunks, st = qk.frame(path.read_bytes(), seekable=True, chunk_capacity=cap)
    assert b"".join(chunks) == want
    assert all(c.endswith(b"\n") and len(c) <= cap for c in chunks)
    assert st["lines"] == ost["lines"] and st["bases"] == ost["bases"] and st["fastq"] == ost["fastq"]
    assert st["raw_bytes"] == path.stat().st_size


def frame_all(qk, data, **kw):
    chunks, st = qk.frame(data, **kw)
    return b"".join(chunks), st


def test_framer_fasta_rules(qk):
    data = b">r1 desc\nACGT\nGGCC\n>r2\n\nTTTT\n>\n"
    out, st = frame_all(qk, data)
    # '>' lines skipped, every other line (also the empty one) is an independent read (T10)
    assert out == b"ACGT\nGGCC\n\nTTTT\n"
    assert st["lines"] == 4 and st["bases"] == 12 and st["fastq"] == 0


def test_framer_fastq_rules(qk):
    # quality lines starting with '>' or '@' are skipped by position, not by content (Q.c:451-455)
    data = b"@r1\nACGT\n+\n>III\n@r2\nGGNC\n+r2\n@@@@\n"
    out, st = frame_all(qk, data)
    assert out == b"ACGT\nGGNC\n"
    assert st["fastq"] == 1 and st["lines"] == 2
    # a '>' line where a read is expected is skipped WITHOUT consuming the 3 trailing lines (Q.c:398)
    data = b"@r1\n>odd\nACGT\n+\nIIII\n@r2\nTTTT\n+\nIIII\n"
    out, _ = frame_all(qk, data)
    assert out == b"ACGT\nTTTT\n"


def test_framer_pipe_loses_first_line(qk, oracle, tmp_path):
    data = b"ACGTACGT\n>h\nGGGG\n"
    assert frame_all(qk, data, seekable=True)[0] == b"ACGTACGT\nGGGG\n"
    assert frame_all(qk, data, seekable=False)[0] == b"GGGG\n"          # Q.c:396: fseek fails on a pipe
    fq = b"@h\nACGT\n+\nIIII\n"
    assert frame_all(qk, fq, seekable=False)[0] == b"ACGT\n"            # FASTQ never rewinds
    # through a real pipe with the C entry point
    r, w = os.pipe()
    os.write(w, data)
    os.close(w)
    L = qk.lib()
    fr = L.qk_framer_open_fd(r, 0)
    dst = np.empty(100000, dtype=np.uint8)
    n, nl = C.c_size_t(), C.c_uint32()
    assert L.qk_framer_next(fr, dst.ctypes.data, dst.size, C.byref(n), None, 0, C.byref(nl)) == 1
    assert dst[: n.value].tobytes() == b"GGGG\n" and nl.value == 1
    assert L.qk_framer_next(fr, dst.ctypes.data, dst.size, C.byref(n), None, 0, C.byref(nl)) == 0
    L.qk_framer_close(fr)


def test_framer_edge_inputs(qk):
    assert frame_all(qk, b"")[0] == b""
    assert frame_all(qk, b"\n")[0] == b"\n"
    out, st = frame_all(qk, b">h\nACGT")                 # T9: unterminated last line gets a '\n'
    assert out == b"ACGT\n" and st["unterminated"] == 1
    out, st = frame_all(qk, b"ACGT\r\nGG\r\n")           # T6: '\r' stays in the line (a base for the codec)
    assert out == b"ACGT\r\nGG\r\n" and st["bases"] == 8
    big = b"A" * 99998 + b"\n"                           # T8: the longest line the reference reads whole
    out, st = frame_all(qk, b">h\n" + big + b"CC\n", chunk_capacity=100000)
    assert out == big + b"CC\n" and st["long_lines"] == 0
    chunks, _ = qk.frame(b">h\n" + big + b"CC\n", chunk_capacity=100000)
    assert chunks == [big, b"CC\n"]                      # chunk boundaries fall between lines


def test_framer_offsets(qk):
    data = b">a\nACGT\nGG\n\nTTTTT\n"
    chunks, offsets, st = qk.frame(data, with_offsets=True)
    assert chunks == [b"ACGT\nGG\n\nTTTTT\n"]
    assert offsets[0].tolist() == [0, 5, 8, 9, 15]


def test_framer_streams_a_file_larger_than_its_window(qk, oracle, synth, tmp_path):
    synth("ref", "--out", tmp_path / "ref.fa", "--bases", 300000, "--seed", 3)
    synth("reads", "--ref", tmp_path / "ref.fa", "--out", tmp_path / "r.fq", "--n", 60000, "--len", 150, "--seed", 9,
          "--fastq", "--rand-qual")                       # ~19 MB > the 8 MiB read window
    want, ost = oracle.frame_file(tmp_path / "r.fq")
    L = qk.lib()
    fr = L.qk_framer_open(os.fsencode(str(tmp_path / "r.fq")))
    assert fr
    dst = np.empty(1 << 20, dtype=np.uint8)
    n, nl = C.c_size_t(), C.c_uint32()
    got, lines = [], 0
    while L.qk_framer_next(fr, dst.ctypes.data, dst.size, C.byref(n), None, 0, C.byref(nl)) == 1:
        got.append(dst[: n.value].tobytes())
        lines += nl.value
    st = qk.FramerStats()
    L.qk_framer_get_stats(fr, C.byref(st))
    L.qk_framer_close(fr)
    assert b"".join(got) == want
    assert lines == ost["lines"] == 60000 == st.lines
    assert L.qk_framer_open(b"/nonexistent/reads.fa") is None


# ---------------------------------------------------------------------------- files ------
def test_qm_header(qk):
    hdr = qk.QmHeader()
    assert qk.lib().qk_qm_read_header(os.fsencode(str(GOLDEN / "k30_fasta_t0" / "ref.fa.qm")), C.byref(hdr)) == 0
    assert (hdr.k, hdr.hash_size, hdr.first_idx) == (30, 0x4000, 0x3A61)   # meta.json: reference stdout
    assert qk.lib().qk_qm_read_header(b"/nonexistent.qm", C.byref(hdr)) == 6


@pytest.mark.parametrize("case", [c for c in golden_cases() if golden_meta(c)["has_txt"]])
def test_gc_txt_writer_matches_reference(case, qk, tmp_path):
    """.txt from (.bin, .qgc) with the sums done in numpy: the formatting and the
    mean/variance arithmetic of Q.c:529-538 are the host's."""
    d = GOLDEN / case
    depth = np.fromfile(d / "expect.bin", dtype=np.uint16).astype(np.int64)
    qgc = np.fromfile(d / "ref.fa.qgc", dtype=np.uint16)
    ctrl = (qgc & 0x8000) != 0
    bins = (qgc & 0x1FF).astype(np.int64)
    s = np.bincount(bins[ctrl], weights=depth[ctrl], minlength=401).astype(np.uint64)
    q = np.bincount(bins[ctrl], weights=depth[ctrl] ** 2, minlength=401).astype(np.int64)
    c = np.bincount(bins[ctrl], minlength=401).astype(np.uint64)
    mean = qk.write_gc_txt(tmp_path / "o.txt", s, q, c)
    assert (tmp_path / "o.txt").read_bytes() == (d / "expect.txt").read_bytes()
    line = [l for l in golden_meta(case)["reference_stdout"] if l.startswith("Mean sequencing depth")][0]
    assert line == f"Mean sequencing depth: {mean:.2f}"


def test_bin_writer(qk, tmp_path):
    a = np.arange(70000, dtype=np.uint32).astype(np.uint16)
    assert qk.lib().qk_write_bin(os.fsencode(str(tmp_path / "a.bin")), a.ctypes.data, a.size) == 0
    assert (tmp_path / "a.bin").read_bytes() == a.astype("<u2").tobytes()
    assert qk.lib().qk_write_bin(b"/nonexistent/dir/a.bin", a.ctypes.data, a.size) == 6


def test_cli_usage_and_missing_inputs(qk, tmp_path):
    res = qk.run_cli(["count"])
    assert res.returncode == 1 and "quicKmer2 count [Options] ref.fa sample.fast[a/q] Out_prefix" in res.stdout
    res = qk.run_cli(["count", "-h"])
    assert res.returncode == 1
    res = qk.run_cli(["count", tmp_path / "missing", tmp_path / "r.fa", tmp_path / "o"])
    assert res.returncode == 1 and "open fail" in res.stdout       # the reference segfaults here (Q.c:345)
    res = qk.run_cli(["est"])
    assert res.returncode == 1


# ---------------------------------------------------------------------------- geometry ---
@pytest.mark.parametrize("n", [1, 5, 1000, 5873, 999971, 58217593, 499708597, 1487556863, 2200000000, (1 << 32) - 1])
@pytest.mark.parametrize("k", [3, 30, 31])
def test_table_geometry_invariants(n, k, qk):
    """What the probe relies on: power-of-two buckets, <= 2.1 keys per 4-entry bucket, an entry
    (remainder + ordinal) that leaves bit 63 free, ordinals in 32 bits, padded extension arrays."""
    d = qk.TableDesc()
    assert qk.lib().qk_table_geometry(n, k, C.byref(d)) == 0
    assert d.n_kmers == n and d.k == k
    assert d.n_buckets & (d.n_buckets - 1) == 0 and d.n_buckets >= 64
    assert n / d.n_buckets <= 2.1
    assert d.rem_bits + d.bucket_bits == 60 and (1 << d.bucket_bits) == d.n_buckets
    assert (1 << d.ord_bits) > n and d.ord_bits <= 32            # holds ordinal + 1
    assert d.rem_bits + d.ord_bits <= 63
    assert d.table_bytes == d.n_buckets * 32
    assert d.stash_slots & (d.stash_slots - 1) == 0 and d.stash_bytes == d.stash_slots * 16
    assert d.has_ext == (1 if k == 30 else 2 if 3 <= k < 30 else 3 if k == 31 else 0)
    if d.has_ext:
        assert d.ext_bytes >= ((n + 15) // 16 + 4) * 12 and d.cont_bytes == 0
    else:
        assert d.ext_bytes == 0 and d.cont_bytes == 0


def test_table_geometry_rejects_nonsense(qk):
    d = qk.TableDesc()
    for n, k in ((0, 30), (1 << 32, 30), (10, 0), (10, 33)):
        assert qk.lib().qk_table_geometry(n, k, C.byref(d)) == 2


def test_human_scale_fits_one_gpu(qk):
    """2.2 G k-mers (2^32-slot .qm).  Peak device memory = build transients + table."""
    d = qk.TableDesc()
    n = 2_200_000_000
    qk.lib().qk_table_geometry(n, 30, C.byref(d))
    resident = d.table_bytes + d.stash_bytes + d.ext_bytes + d.cont_bytes + 4 * (n + 1)
    raw = (1 << 32) * 12                      # keys + chain while the chain is ranked
    kbo = 8 * (n + 1)                         # keys by ordinal while the table is filled
    assert max(raw + kbo, kbo + resident) < 170e9


# ---------------------------------------------------------------------------- streams ----
def test_stream_plain_and_gzip(qk, tmp_path):
    """qk_stream_*: plain files come back as they are, gzip files (also concatenated members, as
    bgzip writes them) inflated; the same through a pipe; corrupt or truncated gzip data is an error."""
    import gzip
    raw = (GOLDEN / "k30_fastq_t3" / "reads.fq").read_bytes()
    got, gz = qk.read_stream(GOLDEN / "k30_fastq_t3" / "reads.fq")
    assert got == raw and not gz
    (tmp_path / "one.gz").write_bytes(gzip.compress(raw))
    got, gz = qk.read_stream(tmp_path / "one.gz", piece=1000)          # small pieces: many partial inflates
    assert got == raw and gz
    cut = raw.index(b"\n@", len(raw) // 3) + 1
    (tmp_path / "two.gz").write_bytes(gzip.compress(raw[:cut]) + gzip.compress(b"") + gzip.compress(raw[cut:], 1))
    got, gz = qk.read_stream(tmp_path / "two.gz", piece=1 << 20)
    assert got == raw and gz
    big = raw * 40                                                      # several refills of the 1 MiB input window
    (tmp_path / "big.gz").write_bytes(gzip.compress(big, 1))
    got, _ = qk.read_stream(tmp_path / "big.gz", piece=333333)
    assert got == big
    r, w = os.pipe()                                                    # through a pipe
    if os.fork() == 0:
        os.close(r)
        os.write(w, gzip.compress(raw))
        os._exit(0)
    os.close(w)
    got, gz = qk.read_stream(fd=r, seekable=False)
    assert got == raw and gz
    os.wait()
    (tmp_path / "trunc.gz").write_bytes(gzip.compress(raw)[:-200])
    with pytest.raises(qk.QkError):
        qk.read_stream(tmp_path / "trunc.gz")
    bad = bytearray(gzip.compress(raw)); bad[len(bad) // 2] ^= 0xFF
    (tmp_path / "bad.gz").write_bytes(bytes(bad))
    with pytest.raises(qk.QkError):
        qk.read_stream(tmp_path / "bad.gz")
    (tmp_path / "empty").write_bytes(b"")
    assert qk.read_stream(tmp_path / "empty") == (b"", False)
    with pytest.raises(qk.QkError):
        qk.read_stream(tmp_path / "missing")


def bgzf_compress(data: bytes, block=60000, level=6) -> bytes:
    """BGZF as bgzip / BAM write it: gzip members with a 'BC' extra field holding the block size."""
    import struct
    import zlib
    out = bytearray()
    pieces = [data[i:i + block] for i in range(0, len(data), block)] + [b""]   # + the empty end marker
    for piece in pieces:
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        comp = c.compress(piece) + c.flush()
        bsize = 12 + 6 + len(comp) + 8 - 1
        out += struct.pack("<4BI2BH2BHH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6, ord("B"), ord("C"), 2, bsize)
        out += comp + struct.pack("<II", zlib.crc32(piece), len(piece))
    return bytes(out)


def test_stream_bgzf_is_inflated_in_parallel(qk, tmp_path):
    import gzip
    raw = (GOLDEN / "k30_fastq_t3" / "reads.fq").read_bytes() * 30          # ~2.9 MB: ~50 blocks
    (tmp_path / "r.bgz").write_bytes(bgzf_compress(raw))
    assert gzip.decompress((tmp_path / "r.bgz").read_bytes()) == raw       # it IS valid multi-member gzip
    for piece in (777, 1 << 16, 1 << 22):
        got, gz = qk.read_stream(tmp_path / "r.bgz", piece=piece)
        assert got == raw and gz
    os.environ["QK_NO_BGZF"] = "1"                                         # the serial inflater must agree
    try:
        assert qk.read_stream(tmp_path / "r.bgz")[0] == raw
    finally:
        del os.environ["QK_NO_BGZF"]
    big = os.urandom(1 << 20) * 3 + raw * 20                                # incompressible blocks, > one 16 MiB window of input? no: several batches of output
    (tmp_path / "big.bgz").write_bytes(bgzf_compress(big, block=65280, level=1))
    assert qk.read_stream(tmp_path / "big.bgz", piece=1 << 20)[0] == big
    data = bytearray(bgzf_compress(raw))
    data[len(data) // 2] ^= 0x55                                            # corrupt one block: CRC or inflate must catch it
    (tmp_path / "bad.bgz").write_bytes(bytes(data))
    with pytest.raises(qk.QkError):
        qk.read_stream(tmp_path / "bad.bgz")
    (tmp_path / "cut.bgz").write_bytes(bgzf_compress(raw)[:-5000])          # truncated in the middle of a block
    with pytest.raises(qk.QkError):
        qk.read_stream(tmp_path / "cut.bgz")
    r, w = os.pipe()
    if os.fork() == 0:
        os.close(r)
        os.write(w, bgzf_compress(raw[:200000]))
        os._exit(0)
    os.close(w)
    assert qk.read_stream(fd=r, seekable=False)[0] == raw[:200000]
    os.wait()


# ---------------------------------------------------------------------------- framer fuzz
@pytest.mark.parametrize("fastq_like", [True, False])
@pytest.mark.parametrize("seed", range(4))
def test_host_framer_on_malformed_streams(seed, fastq_like, qk, oracle, tmp_path):
    """The host framer against the oracle's fgets loop on streams that are not valid FASTA/FASTQ
    (the oracle's reading of them is pinned to the reference binary in tests/test_oracle.py)."""
    from conftest import weird_stream
    rng = np.random.default_rng(500 + seed)
    seq = "".join(rng.choice(list("ACGTacgtN"), size=5000))
    text = weird_stream(rng, seq, fastq_like, 2000)
    (tmp_path / "w.txt").write_text(text, newline="")
    want, ost = oracle.frame_file(tmp_path / "w.txt")
    for cap in (100000, 123457, 1 << 20):
        chunks, st = qk.frame(text.encode(), seekable=True, chunk_capacity=cap)
        assert b"".join(chunks) == want
        assert (st["lines"], st["bases"], st["fastq"]) == (ost["lines"], ost["bases"], ost["fastq"])
    # a pipe loses the first line in FASTA mode only (Q.c:395-396)
    piped, pst = frame_all(qk, text.encode(), seekable=False)
    if fastq_like:
        assert piped == want
    else:
        (tmp_path / "t.txt").write_text(text[text.index("\n") + 1:], newline="")
        assert piped == oracle.frame_file(tmp_path / "t.txt")[0]


# ---------------------------------------------------------------------------- multi-threaded framer
@pytest.mark.parametrize("fastq_like", [True, False])
@pytest.mark.parametrize("seed", range(3))
def test_mt_framer_on_malformed_streams(seed, fastq_like, qk, oracle, tmp_path):
    """qk_frame_mem_mt (all host cores, blocks framed in parallel, state handed from block to block)
    produces the oracle's framed stream byte for byte -- whatever the thread count, the block size
    (it follows the chunk capacity), the number of consumers and their back-pressure."""
    from conftest import weird_stream
    rng = np.random.default_rng(900 + seed)
    seq = "".join(rng.choice(list("ACGTacgtN"), size=5000))
    text = weird_stream(rng, seq, fastq_like, 6000)
    (tmp_path / "w.txt").write_text(text, newline="")
    want, ost = oracle.frame_file(tmp_path / "w.txt")
    for threads, n_ctx, n_slots, cap, busy in ((1, 1, 2, 200000, 0), (4, 1, 3, 200000, 3), (7, 3, 2, 300000, 2), (3, 2, 4, 1 << 20, 0)):
        chunks, who, st = qk.frame_mt(text.encode(), seekable=True, threads=threads, n_ctx=n_ctx, n_slots=n_slots, cap=cap, busy_every=busy)
        assert b"".join(chunks) == want, (threads, n_ctx, cap)
        assert (st["lines"], st["bases"], st["fastq"], st["raw_bytes"]) == (ost["lines"], ost["bases"], ost["fastq"], len(text.encode()))
        assert all(len(c) <= cap and c.endswith(b"\n") for c in chunks)
        assert sum(l for _, l in who) == ost["lines"]
        if n_ctx > 1 and len(chunks) >= 2 * n_ctx:
            assert len({c for c, _ in who}) == n_ctx       # every consumer got chunks
    # a pipe loses the first line in FASTA mode only (Q.c:395-396)
    piped = b"".join(qk.frame_mt(text.encode(), seekable=False, threads=3, cap=200000)[0])
    assert piped == b"".join(qk.frame(text.encode(), seekable=False, chunk_capacity=200000)[0])   # (pinned to the oracle above)
    if fastq_like:
        assert piped == want


def test_mt_framer_edges(qk):
    """Empty input, no newline at all, unterminated last line (T9), lines that span several blocks, lines of
    the maximum length, and the scalar scan (QK_NO_AVX512) against the single-threaded framer."""
    rng = np.random.default_rng(77)
    base = lambda n: bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n))
    cases = [b"", b"ACGT", b"\n", b"\n\n\n", b">h\nACGT\nAC", b"@r\n" + base(150) + b"\n+\n" + b"I" * 150,
             b">x\n" + base(99998) + b"\n>y\n" + base(70000) + b"\n" + base(30) + b"\n",
             b"".join(b">r%d\n" % i + base(int(rng.integers(0, 4000))) + b"\n" for i in range(800)),
             b"".join(b"@r%d\n" % i + base(150) + b"\n+\n" + bytes(rng.integers(33, 75, size=150, dtype=np.uint8)) + b"\n" for i in range(3000))]
    for data in cases:
        ref_chunks, rst = qk.frame(data, seekable=True, chunk_capacity=400000)
        for env in ({}, {"QK_NO_AVX512": "1"}):
            os.environ.update(env)
            try:
                for threads in (1, 5):
                    chunks, _, st = qk.frame_mt(data, seekable=True, threads=threads, cap=400000)
                    assert b"".join(chunks) == b"".join(ref_chunks)
                    assert {k: st[k] for k in ("lines", "bases", "unterminated", "long_lines", "fastq")} == \
                           {k: rst[k] for k in ("lines", "bases", "unterminated", "long_lines", "fastq")}
            finally:
                for k in env:
                    del os.environ[k]
    # a line longer than a chunk is an error, not a hang
    with pytest.raises(qk.QkError):
        qk.frame_mt(b">x\n" + base(300000) + b"\n", threads=2, cap=200000)


def canonical_text(framed: bytes) -> bytes:
    """What the count kernels keep of a framed stream: per byte its 2-bit code (Q.c:411) and whether it resets the
    register ('N', '\\n': Q.c:403-404) -- written back as text (A C T G for the codes 0..3), empty lines dropped."""
    b = np.frombuffer(framed, dtype=np.uint8)
    out = np.frombuffer(b"ACTG", dtype=np.uint8)[(b >> 1) & 3].copy()
    out[b == ord("N")] = ord("N")
    out[b == ord("\n")] = ord("\n")
    return b"".join(l + b"\n" for l in out.tobytes().split(b"\n") if l)


def packed_text(qk, chunks) -> bytes:
    out = []
    for c in chunks:
        codes, flags = qk.unpack_chunk(c)
        t = np.frombuffer(b"ACTG", dtype=np.uint8)[codes].copy()
        assert set(np.unique(codes[flags == 1])) <= {1, 3}          # a flag sits on a '\n' (code 1) or an 'N' (code 3)
        t[(flags == 1) & (codes == 3)] = ord("N")
        t[(flags == 1) & (codes == 1)] = ord("\n")
        assert t.size % 64 == 0 and t[-1] == ord("\n")
        out.append(t.tobytes())
    return b"".join(l + b"\n" for l in b"".join(out).split(b"\n") if l)


@pytest.mark.parametrize("fastq_like", [False, True])
def test_mt_framer_packed_chunks(fastq_like, qk, oracle, tmp_path):
    """Packed chunks (24 bytes per 64 positions: 2-bit codes + reset flags) say exactly what the text chunks say to the
    count kernels: same codes, same resets, in the same order; only empty lines are added (every block is filled up
    to a multiple of 64 positions).  AVX-512 and scalar packers, any thread count, malformed streams."""
    from conftest import weird_stream
    rng = np.random.default_rng(4242 + fastq_like)
    seq = "".join(rng.choice(list("ACGTacgtNn"), size=6000))
    text = weird_stream(rng, seq, fastq_like, 9000).encode()
    (tmp_path / "w.txt").write_bytes(text)
    want, ost = oracle.frame_file(tmp_path / "w.txt")
    for env in ({}, {"QK_NO_AVX512": "1"}, {"QK_FRAMER_NT": "0"}):
        os.environ.update(env)
        try:
            for threads, n_ctx, n_slots, cap, busy in ((1, 1, 2, 200000, 0), (5, 2, 3, 300032, 3), (3, 1, 4, 1 << 20, 0)):
                chunks, who, st = qk.frame_mt(text, threads=threads, n_ctx=n_ctx, n_slots=n_slots, cap=cap, busy_every=busy, packed=True)
                assert packed_text(qk, chunks) == canonical_text(want), (env, threads)
                assert (st["lines"], st["bases"], st["fastq"]) == (ost["lines"], ost["bases"], ost["fastq"])
                assert all(len(c) % 24 == 0 and len(c) // 24 * 64 <= cap for c in chunks) and sum(l for _, l in who) == ost["lines"]
        finally:
            for k in env:
                del os.environ[k]
    # edges: nothing, a lone unterminated line, lines of every length around the group size, one of the maximum length
    base = lambda n: bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=n))
    for data in (b"", b"ACGT", b">h\nAC", b"".join(b">r\n" + base(n) + b"\n" for n in range(0, 200)), b">x\n" + base(99998) + b"\n" + base(63) + b"\n"):
        ref_chunks, rst = qk.frame(data, seekable=True, chunk_capacity=400000)
        for threads in (1, 4):
            chunks, _, st = qk.frame_mt(data, threads=threads, cap=400000, packed=True)
            assert packed_text(qk, chunks) == canonical_text(b"".join(ref_chunks))
            assert (st["lines"], st["bases"], st["unterminated"]) == (rst["lines"], rst["bases"], rst["unterminated"])


def pack_model(framed: bytes) -> bytes:
    """The packed-chunk format stated in numpy (include/qk_host.h, struct qk_chunk_sink): the stream filled up with
    '\\n' to a multiple of 64 positions; per 64 positions four little-endian 32-bit words of 2-bit codes (16 positions per
    word, the first in the top pair) and 64 flags ('N' or '\\n'), bit p = position p."""
    text = np.frombuffer(framed + b"\n" * (-len(framed) % 64), dtype=np.uint8)
    codes = ((text >> 1) & 3).astype(np.uint32).reshape(-1, 4, 16)
    words = (codes << (2 * (15 - np.arange(16, dtype=np.uint32)))[None, None, :]).sum(axis=2, dtype=np.uint64).astype("<u4")
    flags = np.packbits(((text == ord("N")) | (text == ord("\n"))).reshape(-1, 64), axis=1, bitorder="little")
    return np.concatenate([words.view(np.uint8).reshape(-1, 16), flags], axis=1).tobytes()


def test_packed_chunk_bytes_are_the_documented_format(qk):
    """One block in, one packed chunk out: byte for byte what the format says (AVX-512 and scalar packers); this is
    the layout qk_fetch16<true> (csrc/qk_count.cu) reads, held to the same model in tests/test_gpu_parity.py."""
    rng = np.random.default_rng(0)
    framed = b"".join(bytes(rng.choice(np.frombuffer(b"ACGTNacgtn\r", dtype=np.uint8), size=int(rng.integers(0, 300)))) + b"\n" for _ in range(500))
    raw = b">h\n" + framed.replace(b"\n", b"\n>h\n")[:-3]
    for env in ({}, {"QK_NO_AVX512": "1"}):
        os.environ.update(env)
        try:
            chunks, _, st = qk.frame_mt(raw, threads=1, cap=4 << 20, packed=True)
            assert len(chunks) == 1 and chunks[0] == pack_model(framed), env
            assert st["sink_bytes"] == len(chunks[0])
        finally:
            for k in env:
                del os.environ[k]


def test_mt_framer_on_arbitrary_bytes(qk):
    """Any byte soup -- NULs, high bytes, '>' and '@' anywhere, CR, runs of newlines -- frames like the single-threaded
    statement of the rules, as text and packed; the packed chunk says of every byte what Q.c:403-411 would."""
    from hypothesis import given, settings, strategies as st_

    alphabet = st_.sampled_from([b"A", b"C", b"G", b"T", b"N", b"n", b"a", b"\n", b"\n", b">", b"@", b"+", b"\r", b"\x00", b"\xff", b"I", b" "])

    @settings(max_examples=60, deadline=None)
    @given(st_.lists(alphabet, min_size=0, max_size=3000), st_.integers(1, 5), st_.booleans())
    def check(parts, threads, seekable):
        data = b"".join(parts) * 40            # long enough for several 16 KiB blocks at this chunk capacity
        ref_chunks, rst = qk.frame(data, seekable=seekable, chunk_capacity=200000)
        want = b"".join(ref_chunks)
        chunks, _, st = qk.frame_mt(data, seekable=seekable, threads=threads, cap=200000)
        assert b"".join(chunks) == want
        assert (st["lines"], st["bases"], st["unterminated"], st["fastq"]) == (rst["lines"], rst["bases"], rst["unterminated"], rst["fastq"])
        pchunks, _, pst = qk.frame_mt(data, seekable=seekable, threads=threads, cap=200000, packed=True)
        assert packed_text(qk, pchunks) == canonical_text(want)
        assert (pst["lines"], pst["bases"]) == (rst["lines"], rst["bases"])

    check()


# ---------------------------------------------------------------------------- BAM input
def make_bam(reads, refs=(("chr1", 1000000),), text="@HD\tVN:1.6\n"):
    """A minimal BAM (SAM spec 4.2), uncompressed bytes: reads = [(name, flag, sequence)]."""
    import struct
    out = bytearray(b"BAM\1" + struct.pack("<I", len(text)) + text.encode() + struct.pack("<I", len(refs)))
    for name, length in refs:
        out += struct.pack("<I", len(name) + 1) + name.encode() + b"\0" + struct.pack("<I", length)
    codes = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
    for i, (name, flag, seq) in enumerate(reads):
        n = len(seq)
        packed = bytearray((n + 1) // 2)
        for j, ch in enumerate(seq):
            packed[j >> 1] |= codes[ch] << (4 if j % 2 == 0 else 0)
        cigar = struct.pack("<I", (n << 4) | 0) if n else b""            # nM
        tags = b"NMC\x00" if i % 3 == 0 else b""
        body = struct.pack("<iiBBHHHIiii", 0, 100 + i, len(name) + 1, 30, 4680, 1 if n else 0, flag, n, -1, -1, 0)
        body += name.encode() + b"\0" + cigar + bytes(packed) + b"\x28" * n + tags
        out += struct.pack("<I", len(body)) + body
    return bytes(out)


def test_bam_is_turned_into_the_text_of_its_reads(qk, tmp_path):
    """tutorial.md:144-146 feeds `samtools view -F 3840 s.cram | awk '{print ">\\n"$10}'` to count.  A BAM file is
    taken directly: same text, no samtools -- secondary / QC-fail / duplicate / supplementary records dropped."""
    rng = np.random.default_rng(3)
    reads = []
    for i in range(4000):
        n = int(rng.choice([0, 1, 2, 75, 150, 151, 2000, 70001])) if i % 50 == 0 else 150
        seq = "".join(rng.choice(list("ACGTN"), size=n, p=[0.245, 0.245, 0.245, 0.245, 0.02])) if n else ""
        if i % 97 == 0 and n:
            seq = seq[: n // 2] + "RYM" + seq[n // 2 + 3:]                # IUPAC codes travel as they are
        flag = int(rng.choice([0, 16, 99, 147, 256, 272, 512, 1024, 2048, 2064, 4]))
        reads.append((f"r{i}", flag, seq))
    raw = make_bam(reads)
    want = "".join(f">\n{s}\n" for _, f, s in reads if not (f & 3840) and s).encode()
    (tmp_path / "a.bam").write_bytes(bgzf_compress(raw, block=20000) + bgzf_compress(b""))     # + the BGZF end marker
    for piece in (777, 1 << 16, 1 << 22):
        got, gz = qk.read_stream(tmp_path / "a.bam", piece=piece)
        assert got == want and gz
    import gzip
    (tmp_path / "plain_gzip.bam").write_bytes(gzip.compress(raw))                                # not BGZF: one gzip member
    assert qk.read_stream(tmp_path / "plain_gzip.bam")[0] == want
    os.environ["QK_BAM_EXCLUDE"] = "0"                                                            # keep every record
    try:
        assert qk.read_stream(tmp_path / "a.bam")[0] == "".join(f">\n{s}\n" for _, f, s in reads if s).encode()
    finally:
        del os.environ["QK_BAM_EXCLUDE"]
    r, w = os.pipe()
    if os.fork() == 0:
        os.close(r)
        data = bgzf_compress(raw, block=30000)
        while data:
            data = data[os.write(w, data[:65536]):]
        os._exit(0)
    os.close(w)
    assert qk.read_stream(fd=r, seekable=False)[0] == want
    os.wait()
    (tmp_path / "cut.bam").write_bytes(bgzf_compress(raw[:-40]))                                  # ends inside a record
    with pytest.raises(qk.QkError):
        qk.read_stream(tmp_path / "cut.bam")
    (tmp_path / "tiny.gz").write_bytes(gzip.compress(b"AC"))                                      # gzip text shorter than the probe
    assert qk.read_stream(tmp_path / "tiny.gz")[0] == b"AC"
    (tmp_path / "text.gz").write_bytes(gzip.compress(b"BAM is not what this is\n"))
    assert qk.read_stream(tmp_path / "text.gz")[0] == b"BAM is not what this is\n"
