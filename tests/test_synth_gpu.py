"""The GPU data generator (tools/qk_synth_gpu.cu) that makes the human-scale bench / scale-check
inputs, pinned at small scale:

  * its dictionary = the one `qk_synth dict` writes for the same FASTA (that one is pinned to the
    reference's `search -e 0` in tests/test_oracle.py): same keys in the same chain order, same
    .qgc, every key where Find_hash (Q.c:90-99) finds it;
  * its reads are valid FASTQ / FASTA / framed lines, identical across the three formats, and
    the product counts them exactly like the oracle does.
"""
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, str(ROOT / "tools"))

pytestmark = pytest.mark.gpu


def chain_keys(keys, nxt, first):
    out, c = [], int(first)
    while True:
        out.append(int(keys[c]))
        c = int(nxt[c])
        if c == first:
            return np.asarray(out, dtype=np.uint64)


@pytest.mark.parametrize("k", [30, 25, 12])
def test_gpu_dictionary_equals_qk_synth_dict(k, synth, oracle, tmp_path):
    import qk_synth_gpu as qs
    ref = tmp_path / "ref.fa"
    synth("ref", "--out", ref, "--bases", 300000, "--contigs", 3, "--seed", 11, "--segdups", 6, "--segdup-len", 5000,
          "--nblock", 3000)
    synth("dict", "--ref", ref, "--k", k, "--ctrl-block", 7000)
    raw = (tmp_path / "ref.fa.qm").read_bytes()
    H = int.from_bytes(raw[8:16], "little")
    first = int.from_bytes(raw[16:24], "little")
    hk = np.frombuffer(raw, dtype="<u8", count=H, offset=24)
    hn = np.frombuffer(raw, dtype="<u4", count=H, offset=24 + 8 * H)
    want = chain_keys(hk, hn, first)
    want_gc = np.fromfile(tmp_path / "ref.fa.qgc", dtype=np.uint16)

    g = qs.Genome.from_fasta(ref)
    try:
        info = g.build_dict(k=k, slots=H, ctrl_block=7000)
        assert info["n_unique"] == want.size and info["slots"] == H
        keys, nxt, qgc = g.download_dict()
        got = chain_keys(keys, nxt, info["first"])
        assert np.array_equal(got, want)
        assert np.array_equal(qgc, want_gc)
        assert np.count_nonzero(keys) == want.size
        # Find_hash reaches every key: home slot by DJB, walk toward the middle, no empty slot on the way
        for s in np.flatnonzero(keys)[:: max(1, want.size // 20000)]:
            key = int(keys[s])
            c = oracle.djb(key) & (H - 1)
            step = -1 if c & (H >> 1) else 1
            while int(keys[c]) != key:
                assert keys[c] != 0
                c += step
            assert c == s
        # the file it writes is that dictionary
        g.write_dict(tmp_path / "gpu", threads=3)
        w = (tmp_path / "gpu.qm").read_bytes()
        assert w[:8] == raw[:8] and int.from_bytes(w[8:16], "little") == H and int.from_bytes(w[16:24], "little") == info["first"]
        assert w[24:24 + 8 * H] == keys.tobytes() and w[24 + 8 * H:] == nxt.tobytes()
        assert (tmp_path / "gpu.qgc").read_bytes() == want_gc.tobytes()
    finally:
        g.close()


def test_gpu_genome_and_reads(qk, oracle, tmp_path):
    import qk_synth_gpu as qs
    g = qs.Genome.create(400000, contigs=3, seed=5, dup_period=40000, dup_len=8000, div_ppm=5000, nblock=2500)
    try:
        seq = g.download(0, 400000)
        assert set(seq) <= set(b"ACGTN") and seq.count(b"N") == 2500
        starts = g.contigs()
        assert list(starts) == [0, 133333, 266666, 400000]
        # the duplicated stretches make k-mers non-unique
        info = g.build_dict(k=30, ctrl_block=9000)
        assert 0.6 * 400000 < info["n_unique"] < 0.95 * 400000
        g.write_dict(tmp_path / "ref.fa")

        n = 5000
        fq = g.reads_bytes(seed=3, first=100, n=n, fmt=qs.FASTQ)
        fa = g.reads_bytes(seed=3, first=100, n=n, fmt=qs.FASTA)
        fr = g.reads_bytes(seed=3, first=100, n=n, fmt=qs.FRAMED)
        lq, la, lf = fq.split(b"\n"), fa.split(b"\n"), fr.split(b"\n")
        assert len(lq) == 4 * n + 1 and len(la) == 2 * n + 1 and len(lf) == n + 1
        assert lq[0] == b"@r0000000100" and la[0] == b">r0000000100" and lq[2] == b"+" and lq[3] == b"I" * 150
        assert lq[1::4] == la[1::2] == lf[:-1]
        assert all(len(l) == 150 for l in lf[:-1])
        # a read is a substring of the genome (or its reverse complement) up to a few substitutions
        comp = bytes.maketrans(b"ACGT", b"TGCA")
        found = 0
        for i, r in enumerate(lf[:200]):
            cand = r if (100 + i) % 2 == 0 else r.translate(comp)[::-1]
            found += any(seq.find(cand[a:a + 40]) >= 0 for a in (0, 55, 110))
        assert found >= 195
        # same reads whatever the piece boundaries
        assert g.reads_bytes(seed=3, first=100, n=1000, fmt=qs.FASTQ) + g.reads_bytes(seed=3, first=1100, n=4000, fmt=qs.FASTQ) == fq

        (tmp_path / "r.fq").write_bytes(fq)
        want, ost = oracle.count_bin(tmp_path / "ref.fa.qm", tmp_path / "r.fq")
        with qk.Context(device=0, n_slots=2, chunk_capacity=4 << 20) as ctx:
            ctx.load_dictionary(tmp_path / "ref.fa.qm")
            ctx.count_file(tmp_path / "r.fq")
            st = ctx.stats()
            assert np.array_equal(ctx.finish(), want)
            # (reads that overlap the N block emit fewer than 121 k-mers)
            assert 0.98 * n * 121 < st["total_kmers"] == ost["total_kmers"] <= n * 121 and st["hits"] == ost["hits"] > 0.5 * n * 121

        # HiFi-like records: caller-given lengths, some beyond the 65,536 run-counter wrap
        lens = np.array([1000, 70000, 99998, 15000, 133333], dtype=np.uint32)
        hf = g.reads_bytes(seed=4, first=0, n=5, fmt=qs.FASTA, lens=lens)
        lines = hf.split(b"\n")
        assert [len(l) for l in lines[1::2]] == list(lens) and lines[0] == b">r0000000000"
        (tmp_path / "h.fa").write_bytes(hf[: hf.rindex(b">")])      # drop the 133,333-base line (beyond the reference's buffer)
        want, _ = oracle.count_bin(tmp_path / "ref.fa.qm", tmp_path / "h.fa")
        with qk.Context(device=0, n_slots=2, chunk_capacity=4 << 20) as ctx:
            ctx.load_dictionary(tmp_path / "ref.fa.qm")
            ctx.count_file(tmp_path / "h.fa")
            assert np.array_equal(ctx.finish(), want)
    finally:
        g.close()
