"""The line state machine of qk_frame.cu as function composition, restated in Python (CPU only).

The device framer relies on three algebraic facts: (1) every line is a function on the four
line states, (2) composing these functions is associative, so any split of the stream into
segments (threads, warps, CTAs, chunks) can be scanned, and (3) replaying a segment from its
true incoming state gives the same keep/drop decision per line as the reference's sequential
loop (Q.c:397-398, 451-455).  The element encoding below is the one of the kernel: map[s] in
bits 2s+1:2s, keep[s] in bit 8+s, bit 12 = "segment has a line".
"""
import numpy as np
import pytest

IDENTITY = 0xE4


def fe_line(first_byte, fastq):
    hdr = first_byte == ord(">")
    m0 = 1 if (fastq and not hdr) else 0
    return (m0 | (2 << 2) | (3 << 4) | (0 << 6)) | ((0 if hdr else 1) << 8) | (1 << 12)


def fe_map(e, s):
    return (e >> (2 * s)) & 3


def fe_compose(a, b):
    r = 0
    for s in range(4):
        mid = fe_map(a, s)
        r |= fe_map(b, mid) << (2 * s)
        keep = (b >> (8 + mid)) & 1 if (b >> 12) & 1 else (a >> (8 + s)) & 1
        r |= keep << (8 + s)
    return r | ((a | b) & (1 << 12))


def fe_apply(e, sk):
    s = sk & 3
    keep = (e >> (8 + s)) & 1 if (e >> 12) & 1 else (sk >> 2) & 1
    return fe_map(e, s) | (keep << 2)


def sequential(lines, fastq, state):
    """The reference loop: keep flags per line and the final state."""
    keep = []
    for l in lines:
        if state > 0:
            keep.append(0)
            state = (state + 1) & 3
        elif l[:1] == b">":
            keep.append(0)
        else:
            keep.append(1)
            if fastq:
                state = 1
    return keep, state


def random_lines(rng, n):
    firsts = [b">", b"@", b"+", b"A", b"C", b"", b"N", b"I"]
    return [firsts[int(rng.integers(0, len(firsts)))] + b"x" * int(rng.integers(0, 3)) for _ in range(n)]


@pytest.mark.parametrize("fastq", [False, True])
def test_scan_over_any_segmentation_equals_the_sequential_loop(fastq):
    rng = np.random.default_rng(11 + fastq)
    for _ in range(300):
        lines = random_lines(rng, int(rng.integers(1, 40)))
        state0 = int(rng.integers(0, 4))
        want_keep, want_state = sequential(lines, fastq, state0)
        elems = [fe_line(l[0] if l else ord("\n"), fastq) for l in lines]
        # random segmentation: compose inside segments first, then chain the segment totals
        cuts = sorted(set(rng.integers(0, len(lines) + 1, int(rng.integers(0, 6))).tolist() + [0, len(lines)]))
        sk, got_keep = state0, []
        for a, b in zip(cuts, cuts[1:]):
            total = IDENTITY
            for e in elems[a:b]:
                total = fe_compose(total, e)
            inner = sk                                     # replay the segment from its incoming state
            for e in elems[a:b]:
                inner = fe_apply(e, inner)
                got_keep.append((inner >> 2) & 1)
            sk_after = fe_apply(total, sk)
            assert (sk_after & 3) == (inner & 3)           # the composed total carries the state across
            if b > a:
                assert (sk_after >> 2) & 1 == got_keep[-1]  # ... and the keep flag of the open line
            sk = sk_after
        assert got_keep == want_keep and (sk & 3) == want_state


def test_composition_is_associative():
    rng = np.random.default_rng(5)
    pool = [IDENTITY] + [fe_line(b, f) for b in (ord(">"), ord("@"), ord("A"), ord("\n")) for f in (0, 1)]
    for _ in range(2000):
        a, b, c = (pool[int(rng.integers(0, len(pool)))] for _ in range(3))
        # close the pool under composition now and then
        if rng.integers(0, 4) == 0:
            pool.append(fe_compose(a, b))
        assert fe_compose(fe_compose(a, b), c) == fe_compose(a, fe_compose(b, c))
        assert fe_compose(IDENTITY, a) == a and fe_compose(a, IDENTITY) == a
