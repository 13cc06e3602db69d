"""Parity of the CUDA count path with the reference, through the C ABI
(include/quickmer2_b200.h + include/qk_host.h).  Bit-exact: this is integer work.

Three anchors:
  * the golden .bin/.txt written by the unmodified reference (tests/golden/),
  * the oracle (oracle/qk_oracle.c, itself pinned in tests/test_oracle.py) on seeded inputs,
  * size-independent properties at larger sizes (counter sum = hits, additivity, chunking
    and tiling invariance).
"""
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle_binding
from conftest import GOLDEN, ROOT as ROOT_DIR, golden_cases, golden_meta, weird_stream

pytestmark = pytest.mark.gpu


def gpu_bin(ctx, qm, reads=None, raw=None, seekable=True):
    ctx.load_dictionary(qm)
    st = ctx.count_file(reads) if reads is not None else ctx.count_raw(raw, seekable)
    st.update(ctx.stats())
    return ctx.finish(), st


# ------------------------------------------------------------------ golden fixtures -------
@pytest.mark.parametrize("framer", ["device", "host", "device_small_chunks"])
@pytest.mark.parametrize("case", golden_cases())
def test_golden_bin_and_txt(case, framer, qk, tmp_path):
    """quicKmer2 count ref.fa reads out -> same .bin and .txt bytes as the reference wrote,
    with the record framing done on the device (default) or by the host framer."""
    meta = golden_meta(case)
    d = GOLDEN / case
    kw = dict(host_framer=True) if framer == "host" else {}
    if framer == "device_small_chunks":
        kw = dict(n_slots=3, chunk_capacity=200000)
    st = qk.count(d / "ref.fa", d / meta["reads"], tmp_path / "out", **kw)
    assert (tmp_path / "out.bin").read_bytes() == (d / "expect.bin").read_bytes()
    assert st["total_kmers"] == meta["total_kmers"]
    assert st["n_kmers"] == meta["n_kmers"]
    if meta["has_txt"]:
        assert (tmp_path / "out.txt").read_bytes() == (d / "expect.txt").read_bytes()
    else:
        assert not (tmp_path / "out.txt").exists()
    assert st["launches"] >= 1 and st["kernel_ms"] > 0


@pytest.mark.parametrize("case", ["k30_fasta_t0", "k30_fastq_t3", "k25_fastq"])
def test_cli_is_a_drop_in(case, qk, tmp_path):
    """The C command prints the reference's stdout lines and writes the reference's files."""
    meta = golden_meta(case)
    d = GOLDEN / case
    res = qk.run_cli(["count", "-t", "4", d / "ref.fa", d / meta["reads"], tmp_path / "cli"])
    assert res.returncode == 0, res.stdout + res.stderr
    assert (tmp_path / "cli.bin").read_bytes() == (d / "expect.bin").read_bytes()
    if meta["has_txt"]:
        assert (tmp_path / "cli.txt").read_bytes() == (d / "expect.txt").read_bytes()
    got = [l for l in res.stdout.splitlines() if not l.startswith("[Option]")]
    want = meta["reference_stdout"]
    norm = lambda l: "Counting elapse" if l.startswith("Counting elapse") else l
    if not meta["has_txt"]:
        want = [l if not l.startswith("GC control file") else l.replace(l.split(" ")[3], str(d / "ref.fa.qgc")) for l in want]
    assert [norm(l) for l in got] == [norm(l) for l in want]
    total = [l for l in got if l.startswith("Counting elapse")][0]
    assert total.endswith(f"total {meta['total_kmers']} kmers")


def test_cli_progress_lines(qk, tmp_path):
    """Q.c:446: `Read %liG kmers` once per 2^30 k-mers processed, before `Counting elapse`.  The
    threshold is lowered (QK_PROGRESS_SHIFT) so that a small input crosses it."""
    meta = golden_meta("k30_fastq_t3")
    d = GOLDEN / "k30_fastq_t3"
    res = qk.run_cli(["count", d / "ref.fa", d / meta["reads"], tmp_path / "p"], env=dict(os.environ, QK_PROGRESS_SHIFT="12"))
    assert res.returncode == 0, res.stdout + res.stderr
    lines = res.stdout.splitlines()
    at = [i for i, l in enumerate(lines) if l.startswith("Counting elapse")][0]
    n = meta["total_kmers"] >> 12
    assert n >= 3
    assert lines[at - n:at] == [f"Read {g}G kmers" for g in range(1, n + 1)]
    assert not any(l.startswith("Read ") and l.endswith("G kmers") for l in lines[:at - n] + lines[at:])
    # at the reference's own threshold a small input prints none
    res = qk.run_cli(["count", d / "ref.fa", d / meta["reads"], tmp_path / "p"])
    assert not any(l.endswith("G kmers") for l in res.stdout.splitlines())


def test_cli_reads_from_a_pipe(qk, oracle, tmp_path):
    """README.md:89-90: samtools | awk | quicKmer2 count ref /dev/fd/0 out.  On a pipe the
    reference's fseek(0) fails and the first line is consumed (Q.c:396)."""
    d = GOLDEN / "k30_fasta_t0"
    data = (d / "reads.fa").read_bytes()
    res = subprocess.run([str(qk.CLI_PATH), "count", str(d / "ref.fa"), "/dev/fd/0", str(tmp_path / "p")],
                         input=data, capture_output=True)
    assert res.returncode == 0, res.stdout
    # first line is a '>' header, so losing it changes nothing
    assert (tmp_path / "p.bin").read_bytes() == (d / "expect.bin").read_bytes()
    # headerless stream: the first line IS a read and is lost, exactly as the reference loses it
    # (pinned against the reference binary in tests/test_oracle.py)
    seq = [l for l in (d / "reads.fa").read_text().split("\n") if l and not l.startswith(">")]
    data = ("\n".join(seq[:200]) + "\n").encode()
    res = subprocess.run([str(qk.CLI_PATH), "count", str(d / "ref.fa"), "/dev/stdin", str(tmp_path / "q")],
                         input=data, capture_output=True)
    assert res.returncode == 0, res.stdout
    (tmp_path / "tail.fa").write_bytes(data[data.index(b"\n") + 1:])
    want, _ = oracle.count_bin(d / "ref.fa.qm", tmp_path / "tail.fa")
    assert np.array_equal(np.fromfile(tmp_path / "q.bin", dtype=np.uint16), want)


def test_gzip_input(qk, tmp_path):
    """reads.fq.gz (one member, several members, through a pipe): same .bin / .txt as the plain file."""
    import gzip
    d = GOLDEN / "k30_fastq_t3"
    raw = (d / "reads.fq").read_bytes()
    cut = raw.index(b"\n@", len(raw) // 2) + 1
    (tmp_path / "one.fq.gz").write_bytes(gzip.compress(raw))
    (tmp_path / "two.fq.gz").write_bytes(gzip.compress(raw[:cut]) + gzip.compress(raw[cut:]))
    from test_host import bgzf_compress
    (tmp_path / "blocks.fq.gz").write_bytes(bgzf_compress(raw, block=20000))   # BGZF: inflated by several threads
    for name in ("one.fq.gz", "two.fq.gz", "blocks.fq.gz"):
        res = qk.run_cli(["count", "-t", "2", d / "ref.fa", tmp_path / name, tmp_path / "z"])
        assert res.returncode == 0, res.stdout + res.stderr
        assert (tmp_path / "z.bin").read_bytes() == (d / "expect.bin").read_bytes()
        assert (tmp_path / "z.txt").read_bytes() == (d / "expect.txt").read_bytes()
    res = subprocess.run([str(qk.CLI_PATH), "count", str(d / "ref.fa"), "/dev/stdin", str(tmp_path / "zp")],
                         input=gzip.compress(raw), capture_output=True)
    assert res.returncode == 0, res.stdout
    assert (tmp_path / "zp.bin").read_bytes() == (d / "expect.bin").read_bytes()
    st = qk.count(d / "ref.fa", tmp_path / "one.fq.gz", tmp_path / "py")      # the library call
    assert (tmp_path / "py.bin").read_bytes() == (d / "expect.bin").read_bytes() and st["fastq"] == 1
    (tmp_path / "bad.fq.gz").write_bytes(gzip.compress(raw)[:-100])
    res = qk.run_cli(["count", d / "ref.fa", tmp_path / "bad.fq.gz", tmp_path / "bad"])
    assert res.returncode == 1 and "Counting failed" in res.stdout                # a truncated file is an error, not a short count


@pytest.mark.parametrize("packed", ["text", "packed"])
@pytest.mark.parametrize("case", golden_cases())
def test_golden_with_the_multithreaded_host_framer(case, packed, qk, tmp_path, monkeypatch):
    """qk_count_file_mt / qk_count_mem_mt: host threads frame blocks of the input in parallel and ship only
    the sequence lines -- as text, or packed to 2-bit codes + reset flags (24 bytes per 64 positions, read by
    qk_count_ext32_kernel<.., PACKED>) -- same .bin as the reference wrote, from a file (mapped), from memory,
    with small chunks and few slots."""
    monkeypatch.setenv("QK_PACKED", "1" if packed == "packed" else "0")
    meta = golden_meta(case)
    d = GOLDEN / case
    want = np.fromfile(d / "expect.bin", dtype=np.uint16)
    raw = (d / meta["reads"]).read_bytes()
    buf = np.frombuffer(raw, dtype=np.uint8)
    # (packed is what runs by default: it gets both slot geometries and the command; text, the fallback, one)
    for n_slots, cap, threads in ((2, 200000, 5), (4, 4 << 20, 16))[: 2 if packed == "packed" else 1]:
        with qk.Context(device=0, n_slots=n_slots, chunk_capacity=cap) as ctx:
            ctx.load_dictionary(d / "ref.fa.qm")
            st = ctx.count_file_mt(d / meta["reads"], threads=threads)
            assert np.array_equal(ctx.finish(), want)
            assert ctx.stats()["total_kmers"] == meta["total_kmers"]
            framed = st["bases"] + st["lines"]
            if packed == "packed":      # (every 512 KiB block is filled up to a multiple of 64 positions)
                assert 0.375 * framed <= st["sink_bytes"] <= 0.375 * framed + 24 * (2 + len(raw) // (16 << 10)) and st["sink_bytes"] % 24 == 0
            else:
                assert st["sink_bytes"] == framed
            ctx.reset()
            st2 = ctx.count_mem_mt(buf.ctypes.data, buf.size, seekable=True, threads=threads)
            assert np.array_equal(ctx.finish(), want) and st2 == st
    # both framing policies of the command
    for pol in ("host", "device") if packed == "packed" else ():
        res = qk.run_cli(["count", "-t", "3", d / "ref.fa", d / meta["reads"], tmp_path / pol], env=dict(os.environ, QK_FRAMER=pol))
        assert res.returncode == 0, res.stdout + res.stderr
        assert (tmp_path / f"{pol}.bin").read_bytes() == (d / "expect.bin").read_bytes()


def test_mt_host_framer_feeds_several_gpus(qk, mid_dict, oracle, synth, tmp_path):
    """One queue of framed chunks, several GPUs taking from it; the sum of their counters is the count."""
    n = qk.lib().qk_device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (gpurun --gpus 2)")
    reads = tmp_path / "r.fq"
    synth("reads", "--ref", mid_dict / "ref.fa", "--out", reads, "--n", 300000, "--len", 150, "--seed", 77, "--fastq", "--rand-qual")
    want, ost = oracle.count_bin(mid_dict / "ref.fa.qm", reads)
    ctxs = [qk.Context(device=i, n_slots=3, chunk_capacity=2 << 20) for i in range(min(n, 4))]
    try:
        for c in ctxs:
            c.load_dictionary(mid_dict / "ref.fa.qm")
        st = ctxs[0].count_file_mt(reads, threads=8, peers=ctxs[1:])
        per = [c.counters().astype(np.uint64) for c in ctxs]
        assert all(p.sum() > 0 for p in per)                      # every GPU took chunks
        assert np.array_equal((sum(per) & 0xFFFF).astype(np.uint16), want)
        assert sum(c.stats()["total_kmers"] for c in ctxs) == ost["total_kmers"] and st["lines"] == ost["lines"]
    finally:
        for c in ctxs:
            c.close()


# ------------------------------------------------------------------ chunking / tiling ------
@pytest.mark.parametrize("case", ["k30_fasta_t0", "k30_long_lines", "k12_fasta", "k31_fasta"])
def test_chunk_boundaries_do_not_matter(case, qk, oracle, gpu_ctx):
    meta = golden_meta(case)
    d = GOLDEN / case
    want = np.fromfile(d / "expect.bin", dtype=np.uint16)
    framed, _ = oracle.frame_file(d / meta["reads"])
    lines = framed.split(b"\n")[:-1]
    gpu_ctx.load_dictionary(d / "ref.fa.qm")
    for cap in (100000, 131071, 250000, 1 << 20):
        gpu_ctx.reset()
        chunk, slot = b"", 0
        for l in lines:
            if len(chunk) + len(l) + 1 > cap:
                gpu_ctx.submit_chunk(chunk, slot=slot)
                slot = (slot + 1) % gpu_ctx.n_slots
                chunk = b""
            chunk += l + b"\n"
        gpu_ctx.submit_chunk(chunk, slot=slot)
        assert np.array_equal(gpu_ctx.finish(), want), cap
        assert gpu_ctx.stats()["total_kmers"] == meta["total_kmers"]


@pytest.mark.parametrize("tiles", [1, 2, 3, 7, 64])
def test_cta_span_boundaries_do_not_matter(tiles, qk, tmp_path):
    """Long lines cross tiles and CTA spans; the carried reset position and the 32-base halo
    must make the split invisible.  Separate process: the span length is an
    environment knob read once."""
    d = GOLDEN / "k30_long_lines"
    code = (
        "import sys; sys.path.insert(0, %r); from conftest import load_package; qk = load_package();"
        "qk.count(%r, %r, %r)" % (str(GOLDEN.parent), str(d / "ref.fa"), str(d / "reads.fa"), str(tmp_path / "o"))
    )
    env = dict(os.environ, QK_TILES_PER_CTA=str(tiles))
    res = subprocess.run(["python", "-c", code], env=env, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert (tmp_path / "o.bin").read_bytes() == (d / "expect.bin").read_bytes()


def test_unaligned_tail_and_tiny_chunks(qk, oracle, gpu_ctx):
    d = GOLDEN / "k30_fasta_t0"
    gpu_ctx.load_dictionary(d / "ref.fa.qm")
    framed, _ = oracle.frame_file(d / "reads.fa")
    lines = framed.split(b"\n")[:-1]
    gpu_ctx.reset()
    for i, l in enumerate(lines):                       # one line per chunk, lengths 0..300
        gpu_ctx.submit_chunk(l + b"\n", slot=i % gpu_ctx.n_slots)
    assert np.array_equal(gpu_ctx.finish(), np.fromfile(d / "expect.bin", dtype=np.uint16))
    gpu_ctx.reset()
    gpu_ctx.submit_chunk(b"")                           # empty chunk is legal
    gpu_ctx.submit_chunk(b"\n")
    assert gpu_ctx.stats()["total_kmers"] == 0


# ------------------------------------------------------------------ oracle, seeded inputs --
@pytest.fixture(scope="module")
def mid_dict(synth, tmp_path_factory):
    """2 Mb reference with segmental duplications and an N block; dictionary by qk_synth dict."""
    d = tmp_path_factory.mktemp("mid")
    synth("ref", "--out", d / "ref.fa", "--bases", 2000000, "--contigs", 3, "--seed", 11, "--segdups", 10,
          "--segdup-len", 5000, "--nblock", 3000)
    synth("dict", "--ref", d / "ref.fa", "--k", 30, "--ctrl-block", 20000)
    return d


@pytest.mark.parametrize("kind", ["fasta", "fastq", "lower_crlf", "hifi"])
def test_against_oracle_seeded(kind, mid_dict, qk, oracle, synth, tmp_path):
    args = {
        "fasta": ["--n", 200000, "--len", 150],
        "fastq": ["--n", 150000, "--len", 150, "--fastq", "--rand-qual"],
        "lower_crlf": ["--n", 50000, "--len", 101, "--lower-ppm", 300000, "--crlf"],
        "hifi": ["--n", 300, "--len", 15000, "--hifi", "--max-len", 99998, "--err-ppm", 1000],
    }[kind]
    reads = tmp_path / ("reads.fq" if kind == "fastq" else "reads.fa")
    synth("reads", "--ref", mid_dict / "ref.fa", "--out", reads, "--seed", 1234, *args)
    ost = oracle.count(mid_dict / "ref.fa", reads, tmp_path / "want")
    st = qk.count(mid_dict / "ref.fa", reads, tmp_path / "got")
    assert (tmp_path / "got.bin").read_bytes() == (tmp_path / "want.bin").read_bytes()
    assert (tmp_path / "got.txt").read_bytes() == (tmp_path / "want.txt").read_bytes()
    for key in ("total_kmers", "hits", "lines", "bases"):
        assert st[key] == ost[key], key
    assert st["hits"] > 0.4 * st["total_kmers"]


@pytest.mark.parametrize("kind", ["fastq", "lower_crlf", "hifi"])
def test_packed_chunks_against_oracle(kind, mid_dict, qk, oracle, synth, tmp_path, monkeypatch):
    """Packed chunks from the host framer (2-bit codes + reset flags) give the oracle's depths: many reads, N blocks,
    lower case and CR (bases like any other, Q.c:411), lines beyond the 65,536 run-counter wrap; one chunk or many."""
    args = {
        "fastq": ["--n", 150000, "--len", 150, "--fastq", "--rand-qual"],
        "lower_crlf": ["--n", 50000, "--len", 101, "--lower-ppm", 300000, "--crlf"],
        "hifi": ["--n", 300, "--len", 15000, "--hifi", "--max-len", 99998, "--err-ppm", 1000],
    }[kind]
    reads = tmp_path / ("reads.fq" if kind == "fastq" else "reads.fa")
    synth("reads", "--ref", mid_dict / "ref.fa", "--out", reads, "--seed", 4321, *args)
    want, ost = oracle.count_bin(mid_dict / "ref.fa.qm", reads)
    monkeypatch.setenv("QK_PACKED", "1")
    for n_slots, cap, threads in ((3, 1 << 20, 6), (8, 32 << 20, 3)):
        with qk.Context(device=0, n_slots=n_slots, chunk_capacity=cap) as ctx:
            ctx.load_dictionary(mid_dict / "ref.fa.qm")
            st = ctx.count_file_mt(reads, threads=threads)
            got, gst = ctx.finish(), ctx.stats()
            assert np.array_equal(got, want)
            assert (gst["total_kmers"], gst["hits"], st["lines"], st["bases"]) == (ost["total_kmers"], ost["hits"], ost["lines"], ost["bases"])
            assert st["sink_bytes"] < 0.4 * (st["bases"] + st["lines"])       # it WAS packed


def test_packed_chunk_api(qk, oracle, gpu_ctx, tmp_path):
    """qk_submit_packed by hand: a chunk packed in numpy counts like its text; sizes must be whole groups of 64
    positions; a dictionary without the dictionary-order kernel (k = 32) refuses packed chunks."""
    d = GOLDEN / "k30_fasta_t0"
    meta = golden_meta("k30_fasta_t0")
    want = np.fromfile(d / "expect.bin", dtype=np.uint16)
    framed, _ = oracle.frame_file(d / meta["reads"])
    from test_host import pack_model                                           # the format, stated in numpy
    packed = np.frombuffer(pack_model(framed), dtype=np.uint8)
    n_pos = packed.size // 24 * 64
    assert n_pos - 64 < len(framed) <= n_pos
    gpu_ctx.load_dictionary(d / "ref.fa.qm")
    gpu_ctx.submit_packed(packed, n_pos)
    assert np.array_equal(gpu_ctx.finish(), want) and gpu_ctx.stats()["total_kmers"] == meta["total_kmers"]
    with pytest.raises(qk.QkError):
        gpu_ctx.submit_packed(packed, n_pos - 16)
    keys, nxt, first = oracle_binding.build_qm_arrays(oracle, np.arange(1, 50, dtype=np.uint64) * 104729, 256)
    gpu_ctx.load_dictionary_arrays(32, keys, nxt, first)
    with pytest.raises(qk.QkError) as e:
        gpu_ctx.submit_packed(packed, n_pos)
    assert e.value.code == 4                                                   # QK_ERR_STATE


@pytest.mark.parametrize("k", [3, 5, 12, 20, 25, 29, 30, 31])
def test_k_sweep_against_oracle(k, qk, oracle, synth, tmp_path):
    """T2: the 60-bit reverse-complement register is independent of k."""
    bases = 60 if k < 6 else 300000
    synth("ref", "--out", tmp_path / "ref.fa", "--bases", bases, "--seed", 100 + k)
    synth("dict", "--ref", tmp_path / "ref.fa", "--k", k)
    synth("reads", "--ref", tmp_path / "ref.fa", "--out", tmp_path / "r.fa", "--n", 20000, "--len", min(150, bases - 1),
          "--seed", k)
    want, ost = oracle.count_bin(tmp_path / "ref.fa.qm", tmp_path / "r.fa")
    with qk.Context(n_slots=2, chunk_capacity=1 << 20) as ctx:
        got, st = gpu_bin(ctx, tmp_path / "ref.fa.qm", tmp_path / "r.fa")
    assert np.array_equal(got, want)
    assert st["total_kmers"] == ost["total_kmers"] and st["hits"] == ost["hits"]


def test_duplicate_keys_from_index(qk, oracle, gpu_ctx, tmp_path):
    """T14: `index` can write the same k-mer into two slots (Q.c:209-216); both are on the
    chain, only the one Find_hash reaches first ever receives counts."""
    rng = np.random.default_rng(5)
    uniq = rng.integers(1, 1 << 60, size=400, dtype=np.uint64)
    order = np.concatenate([uniq, uniq[:50], uniq[100:120]])      # 70 duplicated entries
    rng.shuffle(order)
    keys, nxt, first = oracle_binding.build_qm_arrays(oracle, order, 2048)
    oracle_binding.write_qm(tmp_path / "d.qm", 30, keys, nxt, first)

    def decode(key):                                               # key -> a 30-mer whose canonical key it is
        return "".join("ACTG"[(int(key) >> (2 * (29 - i))) & 3] for i in range(30))
    reads = "".join(f">r{i}\n{decode(kk)}\n" for i, kk in enumerate(order.tolist() * 3))
    (tmp_path / "r.fa").write_text(reads)
    want, ost = oracle.count_bin(tmp_path / "d.qm", tmp_path / "r.fa")
    got, st = gpu_bin(gpu_ctx, tmp_path / "d.qm", tmp_path / "r.fa")
    assert got.size == order.size == want.size
    assert np.array_equal(got, want)
    assert (want == 0).sum() >= 70                                  # the shadowed copies never count
    assert gpu_ctx.table_desc().skipped_keys == 70


def test_dictionary_keys_no_read_can_produce(qk, oracle, gpu_ctx, tmp_path):
    """Keys >= 2^60 (possible with `index`, k >= 31) are on the chain but can never match:
    the canonical key is <= the 60-bit reverse-complement register (Q.c:415-420)."""
    order = np.array([5, (1 << 61) + 3, 77, (1 << 62) - 1, 9], dtype=np.uint64)
    keys, nxt, first = oracle_binding.build_qm_arrays(oracle, order, 64)
    oracle_binding.write_qm(tmp_path / "d.qm", 31, keys, nxt, first)
    (tmp_path / "r.fa").write_text(">a\n" + "A" * 28 + "CC" + "A" + "\n")
    want, _ = oracle.count_bin(tmp_path / "d.qm", tmp_path / "r.fa")
    got, _ = gpu_bin(gpu_ctx, tmp_path / "d.qm", tmp_path / "r.fa")
    assert np.array_equal(got, want) and got.size == 5
    assert gpu_ctx.table_desc().skipped_keys == 2


def test_corrupt_chain_is_rejected(qk, oracle, gpu_ctx):
    order = np.arange(1, 200, dtype=np.uint64) * 7919
    keys, nxt, first = oracle_binding.build_qm_arrays(oracle, order, 1024)
    bad = nxt.copy()
    occupied = np.flatnonzero(keys)
    bad[occupied[10]] = occupied[10]                    # a self-loop: the walk never returns to first
    with pytest.raises(qk.QkError) as e:
        gpu_ctx.load_dictionary_arrays(30, keys, bad, first)
    assert e.value.code == 5                            # QK_ERR_FORMAT
    bad = nxt.copy()
    bad[first] = np.flatnonzero(keys == 0)[0]           # chain runs into an empty slot and from there to slot 0, for ever
    bad[0] = 0
    with pytest.raises(qk.QkError) as e:
        gpu_ctx.load_dictionary_arrays(30, keys, bad, first)
    assert e.value.code == 5
    bad = nxt.copy()
    chain = chain_slots(nxt, first)
    bad[chain[20]] = chain[3]                           # a cycle that does not contain first
    with pytest.raises(qk.QkError):
        gpu_ctx.load_dictionary_arrays(30, keys, bad, first)
    assert gpu_ctx.load_dictionary_arrays(30, keys, nxt, first) == order.size   # and the intact one loads


def read_qm(path):
    raw = Path(path).read_bytes()
    H = int.from_bytes(raw[8:16], "little")
    keys = np.frombuffer(raw, dtype="<u8", count=H, offset=24).copy()
    nxt = np.frombuffer(raw, dtype="<u4", count=H, offset=24 + 8 * H).copy()
    return raw[4], keys, nxt, int.from_bytes(raw[16:24], "little")


def chain_slots(nxt, first):
    out, c = [], first
    while True:
        out.append(c)
        c = int(nxt[c])
        if c == first:
            return out


def test_occupied_slots_off_the_chain(mid_dict, qk, oracle, synth, gpu_ctx, tmp_path):
    """`sparse` thins the chain and, when enough k-mers are left, keeps the table as it is (Q.c:1443-1461): the
    dropped keys stay in their slots.  `count` still finds them (Q.c:90-99) but never prints them -- and a key that
    sits behind a dropped copy of itself is never reached."""
    k, keys, nxt, first = read_qm(mid_dict / "ref.fa.qm")
    slots = chain_slots(nxt, first)
    keep = [s for i, s in enumerate(slots) if i % 3 != 1 or i == 0]           # every third entry leaves the chain
    assert keep[0] == first
    for a, b in zip(keep, keep[1:] + keep[:1]):
        nxt[a] = b
    # ... and what next[] holds in the empty slots is nobody's business (`index` and `sparse` malloc it, Q.c:189, 1365)
    rng = np.random.default_rng(5)
    empty = np.flatnonzero(keys == 0)
    nxt[empty] = rng.integers(0, 1 << 32, size=empty.size, dtype=np.uint64).astype(np.uint32)
    nxt[empty[::7]] = empty[::7]                                              # self-loops among them
    oracle_binding.write_qm(tmp_path / "thin.qm", k, keys, nxt, first)
    synth("reads", "--ref", mid_dict / "ref.fa", "--out", tmp_path / "r.fa", "--n", 60000, "--len", 150, "--seed", 8)
    want, ost = oracle.count_bin(tmp_path / "thin.qm", tmp_path / "r.fa")
    got, st = gpu_bin(gpu_ctx, tmp_path / "thin.qm", tmp_path / "r.fa")
    assert want.size == len(keep) < len(slots) and np.count_nonzero(keys) == len(slots)
    assert np.array_equal(got, want) and want.sum() > 0
    assert st["total_kmers"] == ost["total_kmers"]


def test_empty_slot_on_the_chain(mid_dict, qk, oracle, synth, gpu_ctx, tmp_path):
    """An `index` input holding the poly-A k-mer puts key 0 -- an empty slot -- on the chain.  Find_hash(0) stops at
    the first empty slot of its probe path and "finds" it (Q.c:98), so that slot, if it is the one on the chain,
    collects the poly-A/poly-T k-mers; any other empty slot on the chain is an entry that stays 0."""
    k, keys, nxt, first = read_qm(mid_dict / "ref.fa.qm")
    H = keys.size
    c = oracle.djb(0) & (H - 1)
    step = -1 if c & (H >> 1) else 1
    while keys[c] != 0:
        c += step
    found, other = c, int(np.flatnonzero(keys == 0)[-1])
    assert found != other
    slots = chain_slots(nxt, first)
    at = {found: 1000, other: 5}
    for e, i in at.items():                                     # splice e in after the i-th entry
        nxt[e] = nxt[slots[i]]
        nxt[slots[i]] = e
    oracle_binding.write_qm(tmp_path / "a.qm", k, keys, nxt, first)
    synth("reads", "--ref", mid_dict / "ref.fa", "--out", tmp_path / "r.fa", "--n", 20000, "--len", 150, "--seed", 9)
    with open(tmp_path / "r.fa", "a") as f:
        f.write(">polyA\n" + "A" * 70 + "\n>polyT\n" + "T" * 45 + "\n>mixed\n" + "A" * 29 + "C" + "A" * 40 + "\n")
    want, ost = oracle.count_bin(tmp_path / "a.qm", tmp_path / "r.fa")
    got, st = gpu_bin(gpu_ctx, tmp_path / "a.qm", tmp_path / "r.fa")
    assert want.size == len(slots) + 2 and np.array_equal(got, want)
    order = chain_slots(nxt, first)
    assert want[order.index(found)] == 41 + 16 + 11 and want[order.index(other)] == 0
    assert st["total_kmers"] == ost["total_kmers"]
    # -t N: the reference fills its last batch of 4,096 keys up with zeros and looks those up too (Q.c:458-466;
    # pinned against the live reference in tests/test_oracle.py) -- the command and qk.count(threads=) do the same
    padded, _ = oracle.count_bin(tmp_path / "a.qm", tmp_path / "r.fa", threads=3)
    assert padded[order.index(found)] == (68 + 4096 - ost["total_kmers"] % 4096) & 0xFFFF
    (tmp_path / "p.fa.qm").write_bytes((tmp_path / "a.qm").read_bytes())
    qk.count(tmp_path / "p.fa", tmp_path / "r.fa", tmp_path / "lib", threads=3)
    assert (tmp_path / "lib.bin").read_bytes() == padded.tobytes()
    for t, expect in (("3", padded), ("0", want)):
        res = qk.run_cli(["count", "-t", t, tmp_path / "p.fa", tmp_path / "r.fa", tmp_path / "cli"])
        assert res.returncode == 0, res.stdout + res.stderr
        assert (tmp_path / "cli.bin").read_bytes() == expect.tobytes(), t


# ------------------------------------------------------------------ dictionary-order extension
def test_extension_path_is_exercised_and_exact(mid_dict, qk, oracle, synth, tmp_path):
    """k = 30: most hits must come from walking the dictionary order (both strands: odd reads
    are reverse complements), and the result must equal the oracle's and the classic kernel's."""
    synth("reads", "--ref", mid_dict / "ref.fa", "--out", tmp_path / "clean.fa", "--n", 100000, "--len", 150, "--seed", 5,
          "--err-ppm", 0)
    synth("reads", "--ref", mid_dict / "ref.fa", "--out", tmp_path / "noisy.fa", "--n", 100000, "--len", 150, "--seed", 6,
          "--err-ppm", 50000)                               # 5 % substitutions: every read breaks the walk several times
    for name, min_frac in (("clean.fa", 0.85), ("noisy.fa", 0.05)):
        want, ost = oracle.count_bin(mid_dict / "ref.fa.qm", tmp_path / name)
        with qk.Context(n_slots=2, chunk_capacity=4 << 20) as ctx:
            ctx.load_dictionary(mid_dict / "ref.fa.qm")
            assert ctx.table_desc().has_ext == 1
            ctx.count_file(tmp_path / name)
            st = ctx.stats()
            assert np.array_equal(ctx.finish(), want), name
            assert st["total_kmers"] == ost["total_kmers"] and st["hits"] == ost["hits"]
            assert st["ext_verified"] >= min_frac * st["hits"], (name, st)
    # the classic kernel (every position probes) on the same input, in a process of its own
    code = ("import sys; sys.path.insert(0, %r); from conftest import load_package; qk = load_package();"
            "st = qk.count(%r, %r, %r); print(st['hits'])" % (str(GOLDEN.parent), str(mid_dict / "ref.fa"), str(tmp_path / "noisy.fa"),
                                                          str(tmp_path / "classic")))
    for env_extra in ({"QK_CLASSIC_KERNEL": "1"}, {"QK_NO_EXT": "1"}):
        res = subprocess.run(["python", "-c", code], env=dict(os.environ, **env_extra), capture_output=True, text=True)
        assert res.returncode == 0, res.stderr
        assert np.array_equal(np.fromfile(tmp_path / "classic.bin", dtype=np.uint16), want)


def test_extension_on_chimeric_and_repetitive_reads(mid_dict, qk, oracle, tmp_path):
    """Reads that jump between loci and strands mid-read, carry N, or sit in tandem repeats: the
    walk must stop exactly where the dictionary order stops describing the read."""
    rng = np.random.default_rng(3)
    seq = "".join(l.strip() for l in open(mid_dict / "ref.fa") if not l.startswith(">"))
    comp = str.maketrans("ACGTN", "TGCAN")
    reads = []
    for i in range(20000):
        parts = []
        for _ in range(int(rng.integers(1, 4))):
            a = int(rng.integers(0, len(seq) - 200))
            piece = seq[a:a + int(rng.integers(20, 120))]
            if rng.integers(0, 2):
                piece = piece.translate(comp)[::-1]
            parts.append(piece)
        r = "".join(parts)
        if i % 7 == 0:
            r = r[:40] + "N" + r[41:]
        if i % 11 == 0:
            unit = seq[a:a + int(rng.integers(1, 9))]
            r = r[:50] + unit * 12 + r[50:]
        reads.append(r)
    (tmp_path / "c.fa").write_text("".join(f">c{i}\n{r}\n" for i, r in enumerate(reads)))
    want, ost = oracle.count_bin(mid_dict / "ref.fa.qm", tmp_path / "c.fa")
    with qk.Context(n_slots=2, chunk_capacity=1 << 20) as ctx:
        ctx.load_dictionary(mid_dict / "ref.fa.qm")
        ctx.count_file(tmp_path / "c.fa")
        st = ctx.stats()
        assert np.array_equal(ctx.finish(), want)
        assert st["hits"] == ost["hits"] and st["ext_verified"] > 0


# ------------------------------------------------------------------ device framing ---------
@pytest.mark.parametrize("fastq_like", [True, False])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_device_framing_state_machine(seed, fastq_like, mid_dict, qk, oracle, tmp_path):
    rng = np.random.default_rng(seed)
    seq = "".join(l.strip() for l in open(mid_dict / "ref.fa") if not l.startswith(">"))[:400000].replace("N", "")
    text = weird_stream(rng, seq, fastq_like, 6000)
    (tmp_path / "w.txt").write_text(text, newline="")
    want, ost = oracle.count_bin(mid_dict / "ref.fa.qm", tmp_path / "w.txt")
    assert ost["fastq"] == int(fastq_like) and ost["lines"] > 500
    for cap in (200000, 1 << 20):
        with qk.Context(n_slots=3, chunk_capacity=cap) as ctx:
            ctx.load_dictionary(mid_dict / "ref.fa.qm")
            st = ctx.count_file(tmp_path / "w.txt"); st.update(ctx.stats())
            assert np.array_equal(ctx.finish(), want)
            for key in ("lines", "bases", "fastq", "total_kmers", "hits"):
                assert st[key] == ost[key], (key, cap)
            # the same bytes from memory (pageable -> staged through the pinned slots)
            ctx.reset()
            st2 = ctx.count_raw(text.encode()); st2.update(ctx.stats())
            assert np.array_equal(ctx.finish(), want) and st2["lines"] == ost["lines"]


def test_device_framing_pipe_and_unterminated(mid_dict, qk, oracle, gpu_ctx, tmp_path):
    seq = "".join(l.strip() for l in open(mid_dict / "ref.fa") if not l.startswith(">"))[:100000].replace("N", "")
    body = "".join(f">r{i}\n{seq[i * 100:i * 100 + 150]}\n" for i in range(500))
    gpu_ctx.load_dictionary(mid_dict / "ref.fa.qm")

    def run(data, seekable):
        gpu_ctx.reset()
        st = gpu_ctx.count_raw(data.encode(), seekable=seekable); st.update(gpu_ctx.stats())
        return gpu_ctx.finish(), st

    def want(data):
        (tmp_path / "x.fa").write_text(data, newline="")
        return oracle.count_bin(mid_dict / "ref.fa.qm", tmp_path / "x.fa")

    first = seq[7000:7150] + "\n"
    got, st = run(first + body, True)                       # seekable FASTA: the first line is a read
    w, ost = want(first + body)
    assert np.array_equal(got, w) and st["lines"] == ost["lines"] == 501
    got, st = run(first + body, False)                      # pipe: the first line is lost (Q.c:396)
    w, ost = want(body)
    assert np.array_equal(got, w) and st["lines"] == ost["lines"] == 500
    got, st = run(body + seq[9000:9150], True)              # T9: unterminated last line gets its '\n'
    w, ost = want(body + seq[9000:9150] + "\n")
    assert np.array_equal(got, w) and st["lines"] == 501 and st["unterminated"] == 1
    got, st = run("", True)
    assert st["lines"] == 0 and st["total_kmers"] == 0 and not got.any()
    got, st = run("@only a header\n", True)
    assert st["lines"] == 0 and st["fastq"] == 1


def test_pinned_source_is_dmaed_directly(mid_dict, qk, oracle, synth, tmp_path):
    import torch
    synth("reads", "--ref", mid_dict / "ref.fa", "--out", tmp_path / "r.fq", "--n", 100000, "--len", 150, "--seed", 8,
          "--fastq", "--rand-qual")
    want, ost = oracle.count_bin(mid_dict / "ref.fa.qm", tmp_path / "r.fq")
    raw = torch.from_numpy(np.fromfile(tmp_path / "r.fq", dtype=np.uint8)).pin_memory()
    assert qk.lib().qk_host_is_pinned(raw.data_ptr()) == 1
    with qk.Context(n_slots=4, chunk_capacity=4 << 20) as ctx:
        ctx.load_dictionary(mid_dict / "ref.fa.qm")
        st = ctx.count_mem(raw.data_ptr(), raw.numel()); st.update(ctx.stats())
        assert np.array_equal(ctx.finish(), want)
        assert st["lines"] == ost["lines"] == 100000 and st["total_kmers"] == ost["total_kmers"]


# ------------------------------------------------------------------ sharded file -----------
@pytest.mark.parametrize("world", [2, 3, 5])
@pytest.mark.parametrize("kind", ["fastq", "fasta", "fastq_out_of_phase"])
def test_sharded_file_equals_whole_file(kind, world, mid_dict, qk, oracle, synth, tmp_path):
    """One reads file cut into `world` line-aligned shards, each counted by its own context
    (the ranks of quick-mer2_b200/dist.py, emulated one after another on this GPU), counters
    added: must equal the reference's whole-file result, for FASTQ through the guessed line
    state and -- when the guess is wrong -- through the verify-and-recount loop."""
    from conftest import load_dist
    qd = load_dist()
    reads = tmp_path / ("r.fa" if kind == "fasta" else "r.fq")
    synth("reads", "--ref", mid_dict / "ref.fa", "--out", reads, "--n", 30000, "--len", 150, "--seed", 77,
          *(["--fastq", "--rand-qual"] if kind != "fasta" else []))
    if kind == "fastq_out_of_phase":                      # '>' where a read is expected shifts the phase (Q.c:398)
        data = reads.read_bytes()
        cut = data.index(b"\n@", len(data) // 10) + 1
        reads.write_bytes(data[:cut] + b"@odd\n>not a read\n+\nIIII\n" + data[cut:])
    want, ost = oracle.count_bin(mid_dict / "ref.fa.qm", reads)
    plans = [qd.shard_plan(qk, reads, r, world) for r in range(world)]
    ctxs = [qk.Context(n_slots=2, chunk_capacity=1 << 20) for _ in range(world)]
    try:
        finals, stats = [], []
        for ctx, p in zip(ctxs, plans):
            ctx.load_dictionary(mid_dict / "ref.fa.qm")
            st, fin = ctx.count_range(reads, p["begin"], p["end"], p["fastq"], p["state"])
            finals.append(fin); stats.append(st)
        recounts = 0
        while True:
            bad = qd.first_wrong_guess([p["state"] for p in plans], finals, [p["end"] - p["begin"] for p in plans])
            if bad is None:
                break
            recounts += 1
            assert recounts <= world
            plans[bad]["state"] = finals[bad - 1]
            ctxs[bad].reset()
            stats[bad], finals[bad] = ctxs[bad].count_range(reads, plans[bad]["begin"], plans[bad]["end"], True, plans[bad]["state"])
        total = sum(c.counters().astype(np.int64) for c in ctxs)
        assert np.array_equal(qd.wrap16(total), want)
        assert sum(s["lines"] for s in stats) == ost["lines"]
        assert sum(c.stats()["total_kmers"] for c in ctxs) == ost["total_kmers"]
        assert (recounts > 0) == (kind == "fastq_out_of_phase")
    finally:
        for c in ctxs:
            c.close()


def _n_gpus(qk):
    return qk.lib().qk_device_count()


@pytest.mark.parametrize("framer", ["host", "device"])
@pytest.mark.parametrize("kind", ["fastq", "fasta", "fastq_out_of_phase", "golden"])
def test_cli_on_several_gpus(kind, framer, mid_dict, qk, oracle, synth, tmp_path):
    """`quicKmer2_b200 count -g 0,1[,2,3]`: one process, NCCL broadcast of the dictionary, one
    shard per GPU, NCCL reduce of the counters -- same .bin / .txt as one GPU and the reference."""
    n = _n_gpus(qk)
    if n < 2:
        pytest.skip("needs at least 2 GPUs (gpurun --gpus 2)")
    gpus = ",".join(str(i) for i in range(min(n, 4)))
    if kind == "golden":
        d = GOLDEN / "k30_fastq_t3"
        ref, reads = d / "ref.fa", d / "reads.fq"
        want_bin, want_txt = (d / "expect.bin").read_bytes(), (d / "expect.txt").read_bytes()
    else:
        ref = mid_dict / "ref.fa"
        reads = tmp_path / ("r.fa" if kind == "fasta" else "r.fq")
        synth("reads", "--ref", ref, "--out", reads, "--n", 200000, "--len", 150, "--seed", 21,
              *(["--fastq", "--rand-qual"] if kind != "fasta" else []))
        if kind == "fastq_out_of_phase":
            data = reads.read_bytes()
            cut = data.index(b"\n@", len(data) // 10) + 1
            reads.write_bytes(data[:cut] + b"@odd\n>not a read\n+\nIIII\n" + data[cut:])
        oracle.count(ref, reads, tmp_path / "want")
        want_bin, want_txt = (tmp_path / "want.bin").read_bytes(), (tmp_path / "want.txt").read_bytes()
    # framer = host: all host cores frame, chunks go to whichever GPU is free; device: one raw shard per GPU
    res = qk.run_cli(["count", "-t", "4", "-g", gpus, ref, reads, tmp_path / "multi"], env=dict(os.environ, QK_FRAMER=framer))
    assert res.returncode == 0, res.stdout + res.stderr
    assert (tmp_path / "multi.bin").read_bytes() == want_bin
    assert (tmp_path / "multi.txt").read_bytes() == want_txt
    import json
    info = json.loads(res.stderr.strip().splitlines()[-1])
    assert info["gpus"] == min(n, 4)
    res1 = qk.run_cli(["count", "-g", "0", ref, reads, tmp_path / "single"])
    assert res1.returncode == 0
    assert json.loads(res1.stderr.strip().splitlines()[-1])["total_kmers"] == info["total_kmers"]
    assert [l for l in res.stdout.splitlines() if "total" in l][0].split("total")[1] == \
           [l for l in res1.stdout.splitlines() if "total" in l][0].split("total")[1]


# ------------------------------------------------------------------ properties at size -----
def test_properties_at_size(qk, synth, tmp_path):
    """16 Mb dictionary, 2 M reads (~240 M k-mers): too slow for the oracle in a unit test, so
    check what must hold at any size."""
    synth("ref", "--out", tmp_path / "ref.fa", "--bases", 16000000, "--contigs", 4, "--seed", 2024, "--segdups", 50,
          "--segdup-len", 20000, "--nblock", 50000)
    synth("dict", "--ref", tmp_path / "ref.fa", "--k", 30)
    synth("reads", "--ref", tmp_path / "ref.fa", "--out", tmp_path / "a.fa", "--n", 1000000, "--len", 150, "--seed", 1)
    synth("reads", "--ref", tmp_path / "ref.fa", "--out", tmp_path / "b.fq", "--n", 1000000, "--len", 150, "--seed", 2,
          "--fastq")
    with qk.Context(n_slots=4, chunk_capacity=16 << 20) as ctx:
        n = ctx.load_dictionary(tmp_path / "ref.fa.qm")
        p, n2 = ctx.counters_device_ptr()
        assert n2 == n
        sa = ctx.count_file(tmp_path / "a.fa"); sa.update(ctx.stats())
        a32 = ctx.counters().astype(np.int64)
        assert int(a32.sum()) == sa["hits"]                          # every hit lands on exactly one counter
        assert 0.99 * 121e6 < sa["total_kmers"] <= 1000000 * 121     # 150 bp reads, k=30; a few reads cross the N block
        a = ctx.finish()
        assert np.array_equal(a, (a32 & 0xFFFF).astype(np.uint16))
        ctx.count_file(tmp_path / "b.fq")                            # additivity: counters accumulate
        ab = ctx.finish().astype(np.int64)
        ctx.reset()
        sb = ctx.count_file(tmp_path / "b.fq"); sb.update(ctx.stats())
        b = ctx.finish().astype(np.int64)
        assert np.array_equal(ab, a.astype(np.int64) + b)
        assert sb["fastq"] == 1 and 0.99 * 121e6 < sb["total_kmers"] <= 1000000 * 121
        # depth ~ 2 * 150 bp * 1 M / 16 Mb ~ 18x on unique sequence
        assert 10 < ab[ab > 0].mean() < 25
    with qk.Context(n_slots=2, chunk_capacity=1 << 20) as ctx2:      # other chunking, same answer
        ctx2.load_dictionary(tmp_path / "ref.fa.qm")
        ctx2.count_file(tmp_path / "b.fq")
        assert np.array_equal(ctx2.finish().astype(np.int64), b)


def test_two_counter_buffers_and_async_reset(qk, oracle, gpu_ctx):
    """Back-to-back jobs: buffer 1 takes a job while buffer 0 keeps the previous result."""
    d = GOLDEN / "k30_fasta_t0"
    want = np.fromfile(d / "expect.bin", dtype=np.uint16)
    gpu_ctx.load_dictionary(d / "ref.fa.qm")
    gpu_ctx.count_file(d / "reads.fa")
    gpu_ctx.select_counters(1)
    assert not gpu_ctx.finish().any()                       # a fresh buffer
    gpu_ctx.count_file(d / "reads.fa")
    gpu_ctx.count_file(d / "reads.fa")
    assert np.array_equal(gpu_ctx.finish().astype(np.int64), 2 * want.astype(np.int64))
    gpu_ctx.reset_async()                                   # stream-ordered: no host sync needed before the next job
    gpu_ctx.count_file(d / "reads.fa")
    assert np.array_equal(gpu_ctx.finish(), want)
    gpu_ctx.select_counters(0)
    assert np.array_equal(gpu_ctx.finish(), want)           # untouched by the jobs on buffer 1
    assert gpu_ctx.slot_stream(0) != 0


def test_bin_streamed_from_device(qk, gpu_ctx, tmp_path):
    d = GOLDEN / "k30_fastq_t3"
    gpu_ctx.load_dictionary(d / "ref.fa.qm")
    gpu_ctx.count_file(d / "reads.fq")
    gpu_ctx.write_bin(tmp_path / "s.bin")
    assert (tmp_path / "s.bin").read_bytes() == (d / "expect.bin").read_bytes()
    with pytest.raises(qk.QkError):
        gpu_ctx.write_bin("/nonexistent/dir/s.bin")


def test_depth_wraps_like_uint16(qk, oracle, gpu_ctx, tmp_path):
    """T12: one 30-mer seen 70,000 times reads 70,000 mod 65,536 = 4,464 (no saturation)."""
    kmer = "ACGTTGCATGCCGATAGGCTAACGTTAGCC"
    order = np.array(oracle.chunk_keys(30, kmer.encode() + b"\n"), dtype=np.uint64)
    keys, nxt, first = oracle_binding.build_qm_arrays(oracle, order, 16)
    oracle_binding.write_qm(tmp_path / "d.qm", 30, keys, nxt, first)
    (tmp_path / "r.fa").write_text((kmer + "\n") * 70000)
    got, st = gpu_bin(gpu_ctx, tmp_path / "d.qm", tmp_path / "r.fa")
    assert got.tolist() == [4464] and st["hits"] == 70000


def test_roofline_microbenchmarks_run(qk, gpu_ctx):
    g = gpu_ctx.bench_gather(1 << 30, gran=32, loads_in_flight=4, n_gathers=1 << 26)
    h = gpu_ctx.bench_h2d(8 << 20, repeats=8)
    assert 50 < g < 8000 and 1 < h < 200


# ------------------------------------------------------------------ dictionaries from the reference's own tools (T14)
def _ref_tool(ref_binary, args, cwd):
    res = subprocess.run([str(ref_binary), *map(str, args)], cwd=cwd, capture_output=True, text=True)
    assert res.returncode == 0, res.stdout[-2000:]
    return res.stdout


def _count_both(qk, ref_binary, d, prefix, reads, threads=2):
    """Our command and the reference's on the same files; returns (ours .bin, theirs .bin, ours .txt or None, theirs)."""
    res = qk.run_cli(["count", "-t", threads, prefix, reads, "ours"], cwd=d)
    assert res.returncode == 0, res.stdout + res.stderr
    _ref_tool(ref_binary, ["count", "-t", threads, prefix, reads, "theirs"], d)
    txt = lambda p: (d / p).read_bytes() if (d / p).exists() else None
    return (d / "ours.bin").read_bytes(), (d / "theirs.bin").read_bytes(), txt("ours.txt"), txt("theirs.txt")


@pytest.mark.parametrize("k", [30, 31, 32, 20])
def test_dictionary_written_by_reference_index(k, qk, ref_binary, synth, tmp_path):
    """`quicKmer2 index` (Q.c:127-254): k-mers from column 4 of a BED file, duplicates and reverse-complement
    duplicates included (both copies get a slot and a place on the chain, Q.c:209-216).  k = 32 is the
    reference's degenerate case (mask 1 << 64, Q.c:419): every read k-mer becomes key 0 and the .bin is zeros."""
    if ref_binary is None:
        pytest.skip("compiled reference not available")
    synth("ref", "--out", tmp_path / "g.fa", "--bases", 30000, "--contigs", 1, "--seed", 3)
    seq = "".join(l for l in (tmp_path / "g.fa").read_text().split("\n") if not l.startswith(">"))
    comp = str.maketrans("ACGT", "TGCA")
    rng = np.random.default_rng(k)
    starts = rng.permutation(len(seq) - k)[:4000]
    kmers = [seq[a:a + k] for a in starts]
    kmers += kmers[:60] + [m.translate(comp)[::-1] for m in kmers[100:140]]      # duplicates on both strands
    with open(tmp_path / "kmers.bed", "w") as f:
        for i, m in enumerate(kmers):
            f.write(f"chr1\t{i}\t{i + k}\t{m}\n")
    _ref_tool(ref_binary, ["index", "-k", k, "-s", "16K", "kmers.bed", "idx.fa.qm"], tmp_path)
    synth("reads", "--ref", tmp_path / "g.fa", "--out", tmp_path / "r.fq", "--n", 3000, "--len", 150, "--seed", 8, "--fastq")
    ours, theirs, _, _ = _count_both(qk, ref_binary, tmp_path, "idx.fa", "r.fq")
    assert ours == theirs and len(ours) == 2 * len(kmers)
    got = np.frombuffer(ours, dtype=np.uint16)
    assert (got.sum() > 0) == (k != 32)


@pytest.mark.parametrize("tool", ["search_e1", "search_e2", "sparse"])
def test_dictionary_written_by_reference_search_and_sparse(tool, qk, ref_binary, synth, tmp_path):
    """`search -e 1 / -e 2` (edit-distance filter, Q.c:1190-1231) and `sparse` (thinned .rqm, Q.c:1306-1483):
    the dictionaries the reference's own tools write, counted by both commands."""
    if ref_binary is None:
        pytest.skip("compiled reference not available")
    synth("ref", "--out", tmp_path / "ref.fa", "--bases", 60000, "--contigs", 2, "--seed", 12, "--segdups", 3, "--segdup-len", 3000,
          "--divergence-ppm", 20000, "--nblock", 500)
    synth("ctrl", "--ref", tmp_path / "ref.fa", "--out", tmp_path / "ctrl.bed", "--block", 5000)
    e = {"search_e1": 1, "search_e2": 2, "sparse": 0}[tool]
    _ref_tool(ref_binary, ["search", "-k", 30, "-e", e, "-t", 4, "-s", "256K", "-c", "ctrl.bed", "ref.fa"], tmp_path)
    prefix = "ref.fa"
    if tool == "sparse":
        _ref_tool(ref_binary, ["sparse", "-c", "ctrl.bed", 7, "ref.fa"], tmp_path)
        (tmp_path / "thin.fa.qm").write_bytes((tmp_path / "ref.fa.rqm").read_bytes())   # count opens <prefix>.qm only (Q.c:335-337)
        for ext in (".qgc",):
            if (tmp_path / ("ref.fa" + ext)).exists():
                (tmp_path / ("thin.fa" + ext)).write_bytes((tmp_path / ("ref.fa" + ext)).read_bytes())
        prefix = "thin.fa"
    synth("reads", "--ref", tmp_path / "ref.fa", "--out", tmp_path / "r.fq", "--n", 20000, "--len", 150, "--seed", 4, "--fastq")
    ours, theirs, otxt, ttxt = _count_both(qk, ref_binary, tmp_path, prefix, "r.fq", threads=3)
    assert ours == theirs and len(ours) > 1000
    assert otxt == ttxt
    assert np.frombuffer(ours, dtype=np.uint16).sum() > 0


def test_finish_async_overlaps_the_next_sample(qk, tmp_path):
    """qk_finish_async: the depths of one sample travel to (page-locked) host memory while the next sample is
    counted into the other counter buffer; both results are what the reference wrote."""
    import ctypes
    import sys
    from conftest import ROOT
    sys.path.insert(0, str(ROOT / "tools"))
    import qk_synth_gpu as qs
    d = GOLDEN / "k30_fastq_t3"
    want = np.fromfile(d / "expect.bin", dtype=np.uint16)
    raw = (d / "reads.fq").read_bytes()
    half = raw[: raw.index(b"\n@", len(raw) // 2) + 1]          # a second, different sample: the first half of the reads
    (tmp_path / "half.fq").write_bytes(half)
    with qk.Context(device=0, n_slots=3, chunk_capacity=1 << 20) as ctx:
        n = ctx.load_dictionary(d / "ref.fa.qm")
        ctx.count_file(tmp_path / "half.fq")
        want_half = ctx.finish().copy()
        ptrs = [qs.lib().qs_pinned_alloc(2 * n + 64) for _ in range(2)]
        outs = [np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint16)), shape=(n,)) for p in ptrs]
        try:
            for rep in range(3):
                ctx.select_counters(0)
                ctx.reset_async()
                ctx.count_file(d / "reads.fq")
                ctx.finish_async(outs[0])
                ctx.select_counters(1)
                ctx.reset_async()
                ctx.count_file(tmp_path / "half.fq")          # runs while the first result is on its way
                ctx.finish_wait()
                assert np.array_equal(outs[0], want)
                ctx.finish_async(outs[1])
                ctx.finish_wait()
                assert np.array_equal(outs[1], want_half)
            with pytest.raises(qk.QkError):
                ctx.finish_async(np.empty(n, dtype=np.uint16))  # pageable memory is refused
        finally:
            ctx.finish_wait()
            ctx.select_counters(0)
            for p in ptrs:
                qs.lib().qs_pinned_free(p)


def test_bam_input(qk, oracle, tmp_path):
    """A BAM of reads (file and pipe): same .bin / .txt as the FASTA that `samtools view -F 3840 | awk` would
    make of it (tutorial.md:144-146)."""
    from test_host import bgzf_compress, make_bam
    d = GOLDEN / "k30_fasta_t0"
    seqs = [l for l in (d / "reads.fa").read_text().split("\n") if l and not l.startswith(">")]
    seqs = [s for s in seqs if set(s) <= set("ACGTN")]          # BAM stores upper-case IUPAC codes only
    assert len(seqs) > 100
    (tmp_path / "same.fa").write_text("".join(f">\n{s}\n" for s in seqs))
    oracle.count(d / "ref.fa", tmp_path / "same.fa", tmp_path / "want")
    reads = [(f"q{i}", 16 if i % 2 else 0, s) for i, s in enumerate(seqs)]
    reads.insert(5, ("dup", 1024, seqs[0]))                       # dropped by -F 3840
    reads.insert(9, ("sec", 256, seqs[1]))
    (tmp_path / "r.bam").write_bytes(bgzf_compress(make_bam(reads), block=40000) + bgzf_compress(b""))
    res = qk.run_cli(["count", "-t", "3", d / "ref.fa", tmp_path / "r.bam", tmp_path / "b"])
    assert res.returncode == 0, res.stdout + res.stderr
    assert (tmp_path / "b.bin").read_bytes() == (tmp_path / "want.bin").read_bytes()
    assert (tmp_path / "b.txt").read_bytes() == (tmp_path / "want.txt").read_bytes()
    assert np.fromfile(tmp_path / "b.bin", dtype=np.uint16).sum() > 0
    res = subprocess.run([str(qk.CLI_PATH), "count", str(d / "ref.fa"), "/dev/stdin", str(tmp_path / "p")],
                         input=(tmp_path / "r.bam").read_bytes(), capture_output=True)
    assert res.returncode == 0, res.stdout
    assert (tmp_path / "p.bin").read_bytes() == (tmp_path / "want.bin").read_bytes()


# ------------------------------------------------------------------ est: window depths on the device
@pytest.mark.parametrize("windows", ["contiguous", "list_ends_early", "gaps_overlaps_and_beyond"])
def test_est_window_reduction(windows, qk, oracle, ref_binary, synth, tmp_path):
    """`quicKmer2_b200 est ref sample out.bed` (Q.c:555-685, window reduction on the device) writes the bytes
    the oracle's restatement writes -- which is pinned to the compiled reference in tests/test_oracle.py -- on a
    dictionary of 1.4 M k-mers (three 1 MiB blocks of .qgc), incl. the lines the reference repeats after the
    last window of the list, windows that overlap, leave gaps or lie beyond the data.  The Python smoother is a stub."""
    from test_oracle import write_smooth_stub
    synth("ref", "--out", tmp_path / "ref.fa", "--bases", 1500000, "--contigs", 3, "--seed", 21, "--segdups", 8, "--segdup-len", 4000,
          "--nblock", 2000)
    synth("dict", "--ref", tmp_path / "ref.fa", "--k", 30, "--ctrl-block", 25000)
    synth("reads", "--ref", tmp_path / "ref.fa", "--out", tmp_path / "r.fq", "--n", 150000, "--len", 150, "--seed", 4, "--fastq")
    res = qk.run_cli(["count", tmp_path / "ref.fa", tmp_path / "r.fq", tmp_path / "samp"])
    assert res.returncode == 0, res.stdout + res.stderr
    n = (tmp_path / "samp.bin").stat().st_size // 2
    assert n > 2 * 524288
    rng = np.random.default_rng(len(windows))
    rows, at = [], 0
    if windows == "contiguous":
        while at + 700 <= n:
            rows.append((at, at + 700)); at += 700
    elif windows == "list_ends_early":                       # the list stops in the first block: two more blocks follow
        while at + 900 <= 300000:
            rows.append((at, at + 900)); at += 900
    else:
        while at < n + 5000:                                 # overlaps, gaps, empty and reversed ranges, past the end
            size = int(rng.integers(1, 3000))
            left = max(0, at + int(rng.integers(-500, 800)))
            rows.append((left, left + size))
            at = left + size
    with open(tmp_path / "ref.fa.bed", "w") as f:
        for i, (a, b) in enumerate(rows):
            f.write(f"chr{1 + i % 3}\t{a * 2}\t{b * 2}\t{a}\t{b}\n")
    env, curve = write_smooth_stub(tmp_path)
    res = qk.run_cli(["est", tmp_path / "ref.fa", tmp_path / "samp", tmp_path / "ours.bed"], env=env)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "Mean sequencing depth" in res.stdout
    port = ROOT_DIR / "oracle" / "_build" / "qk_oracle"
    assert subprocess.run([str(port), "est", str(tmp_path / "ref.fa"), str(tmp_path / "samp"), str(tmp_path / "want.bed"), str(curve)]).returncode == 0
    ours, want = (tmp_path / "ours.bed").read_bytes(), (tmp_path / "want.bed").read_bytes()
    assert ours == want and ours.count(b"\n") >= 100
    if windows == "list_ends_early":
        assert ours.count(b"\n") == len(rows) + 2            # the last window, printed once more per remaining block
    if ref_binary is not None:                               # and the reference itself, where it travelled
        r = subprocess.run([str(ref_binary), "est", str(tmp_path / "ref.fa"), str(tmp_path / "samp"), str(tmp_path / "theirs.bed")],
                           env=env, capture_output=True, text=True)
        assert r.returncode == 0 and (tmp_path / "theirs.bed").read_bytes() == ours
        assert [l for l in r.stdout.splitlines() if "depth" in l] == [l for l in res.stdout.splitlines() if "depth" in l]
