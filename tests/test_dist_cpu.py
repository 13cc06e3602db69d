"""Multi-rank host logic on CPU: world_size 2 and 3 over gloo (no GPU).

Covers what the N > 1 path does outside the kernels (quick-mer2_b200/dist.py): line-aligned
sharding of one reads file, the FASTQ line-state guess with its verify-and-recount loop, the
table-descriptor broadcast, and the counter reduction with the reference's 16-bit wrap.  The
per-shard "device" is replaced by a small model of the reference's line state machine
(Q.c:397-398, 451-455) plus the oracle's codec -- test infrastructure only.
"""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))


def model_count(data: bytes, fastq: bool, state: int, k: int, oracle):
    """Reference line state machine over `data` (whole lines) from `state`; returns
    (read lines, emitted k-mers, final state)."""
    lines = data.split(b"\n")[:-1]
    kept = []
    for l in lines:
        if state > 0:
            state = (state + 1) & 3
        elif l[:1] == b">":
            pass
        else:
            kept.append(l)
            if fastq:
                state = 1
    kmers = int(oracle.chunk_keys(k, b"".join(x + b"\n" for x in kept)).size) if kept else 0
    return len(kept), kmers, state


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def worker(rank, world, port, path, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from conftest import load_dist, load_package
        import oracle_binding
        qk = load_package()
        qd = load_dist()
        oracle = oracle_binding.Oracle()
        plan = qd.shard_plan(qk, path, rank, world)
        passes = []

        def count_fn(b, e, fastq, state):
            with open(path, "rb") as f:
                f.seek(b)
                data = f.read(e - b)
            lines, kmers, final = model_count(data, fastq, state, 30, oracle)
            passes.append(state)
            return {"lines": lines, "kmers": kmers}, final

        def all_gather(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out

        st = qd.count_sharded(plan, rank, world, count_fn, lambda: None, all_gather)
        # "reduce the counters": every rank holds int32 counters; rank 0 gets the sum, then wraps
        counters = torch.full((8,), 40000 + rank, dtype=torch.int32)
        dist.reduce(counters, 0, op=dist.ReduceOp.SUM)
        # table descriptor broadcast (bytes of struct qk_table_desc)
        desc = qk.TableDesc()
        if rank == 0:
            desc.n_kmers, desc.n_buckets, desc.k, desc.rem_bits = 123456789012, 1 << 30, 30, 30
        raw = torch.frombuffer(bytearray(bytes(desc)), dtype=torch.uint8).clone()
        dist.broadcast(raw, 0)
        got = qk.TableDesc.from_buffer_copy(raw.numpy().tobytes())
        torch.save({"plan": plan, "stats": st, "passes": passes, "sum": counters.numpy().copy(),
                    "desc": (int(got.n_kmers), int(got.n_buckets), int(got.k), int(got.rem_bits))},
                   Path(out_dir) / f"rank{rank}.pt")
    finally:
        dist.destroy_process_group()


def run_world(world, path, tmp_path):
    mp.spawn(worker, args=(world, free_port(), str(path), str(tmp_path)), nprocs=world, join=True)
    return [torch.load(tmp_path / f"rank{r}.pt", weights_only=False) for r in range(world)]


def fastq_text(n, rng, seq):
    out = []
    for i in range(n):
        a = int(rng.integers(0, len(seq) - 200))
        L = int(rng.integers(40, 160))
        q0 = rng.choice(list("@>I+5"))
        out.append(f"@r{i}\n{seq[a:a + L]}\n+\n{q0}{'I' * (L - 1)}\n")
    return "".join(out)


@pytest.fixture(scope="module")
def seq():
    rng = np.random.default_rng(7)
    return "".join(rng.choice(list("ACGT"), size=20000))


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_fastq_equals_whole_file(world, seq, built, oracle, tmp_path):
    rng = np.random.default_rng(world)
    text = fastq_text(3000, rng, seq)
    path = tmp_path / "r.fq"
    path.write_text(text, newline="")
    res = run_world(world, path, tmp_path)
    whole = model_count(text.encode(), True, 3, 30, oracle)
    assert sum(r["stats"]["lines"] for r in res) == whole[0] == 3000
    assert sum(r["stats"]["kmers"] for r in res) == whole[1]
    # shards tile the file, are line aligned, and every guess was right (valid FASTQ): one round
    assert res[0]["plan"]["begin"] == 0 and res[-1]["plan"]["end"] == len(text)
    for a, b in zip(res, res[1:]):
        assert a["plan"]["end"] == b["plan"]["begin"]
        assert text[b["plan"]["begin"] - 1] == "\n"
    assert all(r["stats"]["rounds"] == 1 and len(r["passes"]) == 1 for r in res)
    assert all(r["plan"]["guessed"] for r in res[1:]) and not res[0]["plan"]["guessed"]
    # reduce + 16-bit wrap; descriptor arrives intact
    from conftest import load_dist
    qd = load_dist()
    total = sum(40000 + r for r in range(world))
    assert qd.wrap16(res[0]["sum"]).tolist() == [total & 0xFFFF] * 8 and total > 65535
    assert all(r["desc"] == (123456789012, 1 << 30, 30, 30) for r in res)


def test_wrong_guess_is_detected_and_recounted(seq, built, oracle, tmp_path):
    """A '>' line where a read is expected does not consume the next three lines (Q.c:398), so
    everything after it is out of phase with what a record-shaped guess assumes."""
    rng = np.random.default_rng(11)
    text = fastq_text(400, rng, seq) + "@odd\n>not a read\n+\nIIII\n" + fastq_text(2000, rng, seq)
    path = tmp_path / "odd.fq"
    path.write_text(text, newline="")
    res = run_world(2, path, tmp_path)
    whole = model_count(text.encode(), True, 3, 30, oracle)
    assert sum(r["stats"]["lines"] for r in res) == whole[0]
    assert sum(r["stats"]["kmers"] for r in res) == whole[1]
    assert res[1]["stats"]["rounds"] == 2 and len(res[1]["passes"]) == 2      # rank 1 recounted
    assert res[1]["passes"][0] != res[1]["passes"][1]
    assert len(res[0]["passes"]) == 1


def test_sharded_fasta_needs_no_state(seq, built, oracle, tmp_path):
    text = "".join(f">r{i}\n{seq[i * 7:i * 7 + 120]}\n" for i in range(2000))
    path = tmp_path / "r.fa"
    path.write_text(text, newline="")
    res = run_world(2, path, tmp_path)
    whole = model_count(text.encode(), False, 0, 30, oracle)
    assert sum(r["stats"]["lines"] for r in res) == whole[0] == 2000
    assert sum(r["stats"]["kmers"] for r in res) == whole[1]
    assert not any(r["plan"]["guessed"] for r in res)


def test_state_guess_and_bounds_edge_cases(qk, tmp_path):
    import ctypes as C
    L = qk.lib()
    s = C.c_uint32()

    def guess(text):
        b = np.frombuffer(text.encode(), dtype=np.uint8)
        rc = L.qk_fastq_state_guess(b.ctypes.data, b.size, C.byref(s))
        return rc, s.value
    rec = "@h\nACGT\n+\nIIII\n"
    assert guess(rec * 3) == (0, 3)                       # at a header: one more line to discard
    assert guess("ACGT\n+\nIIII\n" + rec * 3) == (0, 0)   # at a read
    assert guess("+\nIIII\n" + rec * 3) == (0, 1)
    assert guess("@III\n" + rec * 3) == (0, 2)            # a quality line that starts with '@'
    assert guess("no newline at all")[0] == 5
    p = tmp_path / "x.fa"
    p.write_bytes(b"AAAA\n" * 10)
    b, e = C.c_uint64(), C.c_uint64()
    for world in (1, 2, 3, 7, 64):
        cuts = []
        for r in range(world):
            assert L.qk_shard_bounds(os.fsencode(str(p)), r, world, C.byref(b), C.byref(e)) == 0
            cuts.append((b.value, e.value))
            assert b.value % 5 == 0 and e.value % 5 == 0
        assert cuts[0][0] == 0 and cuts[-1][1] == 50
        assert all(x[1] == y[0] for x, y in zip(cuts, cuts[1:]))
    assert L.qk_shard_bounds(b"/nonexistent", 0, 2, C.byref(b), C.byref(e)) == 6
