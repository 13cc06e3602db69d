"""Multi-GPU plumbing for the count path: one process per GPU, ``torch.distributed`` as the
transport (NCCL on GPUs, gloo in the CPU tests).  SURVEY.md 8(e): reads are independent units
and counting is an integer sum, so the read stream is SHARDED, the dictionary REPLICATED and
the per-GPU counters ADDED -- one collective on the data path.

The reference has no counterpart (single process, Q.c:304-545); what must hold is that the
result is the one `quicKmer2 count` writes for the whole file, for any number of shards.

Sharding one FASTA/FASTQ file: shard r is the line-aligned byte range ``qk_shard_bounds``
gives.  The reference's framing loop (Q.c:393-398, 451-455) is a state machine over lines, so
a shard needs the state its predecessor ends in.  FASTA: always 0.  FASTQ: every rank GUESSES
it from the first lines of its shard (``qk_fastq_state_guess``), all shards are counted at
once, then the assumed states are compared with the predecessors' final states; a rank whose
guess was wrong (only possible on malformed FASTQ) zeroes its counters and recounts.  The
result is exact either way; the guess only buys parallelism.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np


class DevMem:
    """A raw device allocation seen through ``__cuda_array_interface__`` (for torch.as_tensor)."""

    def __init__(self, ptr: int, n: int, typestr: str = "|u1"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}


def shard_plan(qk, reads_path, rank: int, world: int) -> dict:
    """Byte range, format and assumed incoming line state of shard `rank` (host only)."""
    L = qk.lib()
    b, e = C.c_uint64(), C.c_uint64()
    rc = L.qk_shard_bounds(os.fsencode(str(reads_path)), rank, world, C.byref(b), C.byref(e))
    if rc:
        raise qk.QkError(rc, f"cannot shard {reads_path} (not a regular file?)")
    with open(reads_path, "rb") as f:
        first = f.read(1)
        fastq = first == b"@"                              # Q.c:395
        if rank == 0:
            state, guessed = (3 if fastq else 0), False    # the first line of a FASTQ is consumed (Q.c:393-395)
        elif not fastq:
            state, guessed = 0, False
        else:
            f.seek(b.value)
            window = np.frombuffer(f.read(1 << 20), dtype=np.uint8)
            s = C.c_uint32()
            L.qk_fastq_state_guess(window.ctypes.data, window.size, C.byref(s))
            state, guessed = s.value, True
    return {"begin": b.value, "end": e.value, "fastq": fastq, "state": state, "guessed": guessed}


def first_wrong_guess(assumed: list[int], final: list[int], sizes: list[int]):
    """The first rank whose assumed incoming state differs from what its predecessors hand on,
    or None.  Shards after it cannot be judged until it has been recounted.  Rank 0 starts the
    stream, so its state is known; empty shards pass the state through."""
    carry = assumed[0]
    for r, (a, f, n) in enumerate(zip(assumed, final, sizes)):
        if n == 0:
            continue
        if r > 0 and a != carry:
            return r
        carry = f
    return None


def count_sharded(plan: dict, rank: int, world: int, count_fn, reset_fn, all_gather_fn) -> dict:
    """Count this rank's shard; repeat while some rank's assumed line state proves wrong.

    count_fn(begin, end, fastq, state) -> (stats, final_state); reset_fn() zeroes this rank's
    counters; all_gather_fn(obj) -> list of every rank's obj.  Returns the stats of the last
    (correct) pass plus ``rounds``.
    """
    state = plan["state"]
    size = plan["end"] - plan["begin"]
    stats, final = count_fn(plan["begin"], plan["end"], plan["fastq"], state)
    rounds = 1
    while True:
        table = all_gather_fn((state, final, size))
        bad = first_wrong_guess([t[0] for t in table], [t[1] for t in table], [t[2] for t in table])
        if bad is None:
            stats["rounds"] = rounds
            return stats
        if rounds > world:
            raise RuntimeError("shard line states did not converge")
        if rank == bad:                                     # recount with the true incoming state
            carry = table[0][0]
            for s_, f_, n_ in table[:rank]:
                if n_ > 0:
                    carry = f_
            state = carry
            reset_fn()
            stats, final = count_fn(plan["begin"], plan["end"], plan["fastq"], state)
        rounds += 1


# ---- device side (needs torch + a CUDA context per rank) -------------------------------------
def replicate_dictionary(qk, ctx, rank: int, device: int, dist) -> int:
    """Rank 0 holds a built table; every other rank adopts its geometry and receives the image
    (table, stash, extension arrays) by broadcast over NCCL.  Returns n_kmers."""
    import torch

    desc = ctx.table_desc() if rank == 0 else qk.TableDesc()
    raw = torch.frombuffer(bytearray(bytes(desc)), dtype=torch.uint8).to(f"cuda:{device}")
    dist.broadcast(raw, 0)
    if rank != 0:
        desc = qk.TableDesc.from_buffer_copy(raw.cpu().numpy().tobytes())
        ctx.adopt(desc)
    for ptr, nbytes in ctx.table_images():              # table, stash, extension arrays
        dist.broadcast(torch.as_tensor(DevMem(ptr, nbytes), device=f"cuda:{device}"), 0)
    torch.cuda.synchronize()
    return int(desc.n_kmers)


def counters_tensor(ctx, device: int):
    """The u32 counters of a context as an int32 torch tensor (int32 sum wraps like u32)."""
    import torch

    ptr, n = ctx.counters_device_ptr()
    return torch.as_tensor(DevMem(ptr, n, "<i4"), device=f"cuda:{device}")


def wrap16(counters: np.ndarray) -> np.ndarray:
    """u32/i32 counters -> the reference's uint16 depth (wraps mod 65,536, Q.c:23,291; T12)."""
    return (counters.astype(np.int64) & 0xFFFF).astype(np.uint16)
