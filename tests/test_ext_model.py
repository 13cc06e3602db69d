"""The dictionary-order walk of qk_count_ext32_kernel, restated in Python for ANY k and checked
against a plain key -> ordinal lookup on adversarial inputs (CPU only).

The CUDA kernel only exists for k = 30, where coincidences are astronomically rare.  The
exactness argument is about keys, not about k, so the same construction is
run here with k = 3..9 on low-complexity, palindrome-rich and repeat-rich sequences, where
"accidental" continuations, reverse-complement palindromes and k-mers that continue in both
orientations occur all the time.  If the walk ever assigned an ordinal that the dictionary
lookup does not, the identities the kernel relies on would be wrong.

Mirrors quick-mer2_b200/csrc/qk_dict.cu (qk_orient_insert_kernel) and qk_count.cu
(qk_count_ext32_kernel): block-wise orientation with look-ahead at run starts, cont / last /
first / strand per ordinal, one anchor per lane of 32 positions (the second half of 16 carries on
from where the first half's walk ended), steps checked in order.
"""
import numpy as np
import pytest

RUN = 16
BLOCK = 8          # ordinals per orientation block (256 on the device): small, to hit the block seams


def rc(x, k):
    """reverse complement in the reference's encoding (A=0 C=1 T=2 G=3, complement = ^2)"""
    out = 0
    for _ in range(k):
        out = (out << 2) | ((x & 3) ^ 2)
        x >>= 2
    return out


def canonical_stream(codes, k):
    """(key, is_fwd) for every position that ends a k-mer (Q.c:399-420 with a k-base rc register)."""
    M = (1 << (2 * k)) - 1
    fwd = rcv = 0
    out = []
    for i, c in enumerate(codes):
        fwd = ((fwd << 2) | c) & M
        rcv = (rcv >> 2) | ((c ^ 2) << (2 * (k - 1)))
        if i >= k - 1:
            out.append((min(fwd, rcv), fwd <= rcv, fwd))
        else:
            out.append(None)
    return out


def build_dictionary(ref_codes, k):
    """Unique canonical k-mers of the reference in reference order (what `search -e 0` keeps)."""
    stream = [s for s in canonical_stream(ref_codes, k) if s is not None]
    counts = {}
    for key, _, _ in stream:
        counts[key] = counts.get(key, 0) + 1
    ordered = [key for key, _, _ in stream if counts[key] == 1 and key != 0]
    return ordered


def orient(keys, k):
    """qk_orient_insert_kernel: walking orientation F, cont / last / first / strand per ordinal."""
    n = len(keys)
    Mlow = (1 << (2 * (k - 1))) - 1
    F = [0] * n
    cont = [0] * n
    for begin in range(0, n, BLOCK):
        prev = None
        for o in range(begin, min(n, begin + BLOCK)):
            K, Kr = keys[o], rc(keys[o], k)
            f, c = K, 0
            if prev is not None:
                want = prev & Mlow
                if (K >> 2) == want:
                    f, c = K, 1
                elif (Kr >> 2) == want:
                    f, c = Kr, 1
            if not c and o + 1 < n:
                nK, nKr = keys[o + 1], rc(keys[o + 1], k)
                if (nK >> 2) == (K & Mlow) or (nKr >> 2) == (K & Mlow):
                    f = K
                elif (nK >> 2) == (Kr & Mlow) or (nKr >> 2) == (Kr & Mlow):
                    f = Kr
            F[o], cont[o] = f, c
            prev = f
    last = [f & 3 for f in F]
    first = [f >> (2 * (k - 1)) for f in F]
    strand = [int(F[o] == keys[o]) for o in range(n)]
    return cont, last, first, strand


def walk_counts(read_codes, keys, k, lookup, cont, last, first, strand):
    """Ordinal per emitting position the way qk_count_ext32_kernel derives it; also how many came from
    the walk.  A lane owns 32 positions = two halves of 16: the first half probes an anchor (its first
    emitting position) and walks; the second half carries on from the ordinal position 15 was settled
    with -- no probe -- and only probes an anchor of its own when position 15 was not settled."""
    stream = canonical_stream(read_codes, k)
    n = len(keys)
    got, walked, probes = {}, 0, 0
    L = len(read_codes)
    for base in range(0, L, 2 * RUN):
        end = None                                   # (ordinal, plus) of position base + 15, when settled
        for half in (0, 1):
            hb = base + RUN * half
            run = [p for p in range(hb, min(hb + RUN, L)) if stream[p] is not None]
            if not run:
                continue
            settled, probed_anchor = {}, None
            if half == 1 and end is not None:
                ja, (oa, plus) = hb - 1, end
            else:
                ja = run[0]
                key, is_fwd, _ = stream[ja]
                oa = lookup.get(key)
                probes += 1
                probed_anchor = ja
                if oa is not None:
                    plus = bool(strand[oa]) == is_fwd
                    settled[ja] = oa
            if oa is not None:
                o = oa
                for p in range(ja + 1, min(hb + RUN, L)):
                    b = read_codes[p]
                    if plus:
                        if o + 1 >= n or not cont[o + 1] or last[o + 1] != b:
                            break
                        o += 1
                    else:
                        if o - 1 < 0 or not cont[o] or first[o - 1] != (b ^ 2):
                            break
                        o -= 1
                    if stream[p] is not None:
                        settled[p] = o
                        walked += 1
            for p in run:
                if p in settled:
                    got[p] = settled[p]
                elif p != probed_anchor:
                    o = lookup.get(stream[p][0])     # pooled probe
                    probes += 1
                    if o is not None:
                        got[p] = o
            if half == 0:
                last_p = hb + RUN - 1
                end = (settled[last_p], plus) if last_p in settled else None
    return got, walked


def brute(read_codes, k, lookup):
    return {p: lookup[s[0]] for p, s in enumerate(canonical_stream(read_codes, k)) if s is not None and s[0] in lookup}


def make_reference(rng, n, flavour):
    if flavour == "uniform":
        return rng.integers(0, 4, n).tolist()
    if flavour == "at_rich":                                # low complexity: lots of accidental overlaps
        return rng.choice(4, n, p=[0.45, 0.05, 0.45, 0.05]).tolist()
    if flavour == "tandem":                                 # short tandem repeats with point mutations
        unit = rng.integers(0, 4, int(rng.integers(2, 7))).tolist()
        ref = (unit * (n // len(unit) + 1))[:n]
        for i in rng.integers(0, n, n // 15):
            ref[i] = int(rng.integers(0, 4))
        return ref
    if flavour == "palindromic":                            # a stretch followed by its reverse complement, repeatedly
        ref = []
        while len(ref) < n:
            s = rng.integers(0, 4, int(rng.integers(5, 40))).tolist()
            ref += s + [c ^ 2 for c in reversed(s)]
        return ref[:n]
    raise ValueError(flavour)


@pytest.mark.parametrize("flavour", ["uniform", "at_rich", "tandem", "palindromic"])
@pytest.mark.parametrize("k", [3, 4, 5, 6, 9])
def test_walk_equals_lookup(k, flavour):
    rng = np.random.default_rng(1000 * k + len(flavour))
    total_walked = total = 0
    n_ref = max(k + 12, min(1500, 4 ** k // 3))               # short enough that unique k-mers exist at small k
    for trial in range(12):
        ref = make_reference(rng, n_ref, flavour)
        keys = build_dictionary(ref, k)
        if len(keys) < 4:
            continue
        lookup = {key: o for o, key in enumerate(keys)}
        cont, last, first, strand = orient(keys, k)
        for _ in range(40):
            span = min(80, len(ref))
            a = int(rng.integers(0, len(ref) - span + 1))
            read = list(ref[a:a + int(rng.integers(k, span + 1))])
            if rng.integers(0, 2):
                read = [c ^ 2 for c in reversed(read)]       # reverse strand
            for i in rng.integers(0, len(read), int(rng.integers(0, 3))):
                read[i] = int(rng.integers(0, 4))            # sequencing errors
            if rng.integers(0, 4) == 0:                      # chimeric: jump to another locus mid-read
                b = int(rng.integers(0, max(1, len(ref) - 30)))
                read = read[:len(read) // 2] + list(ref[b:b + 30])
            got, walked = walk_counts(read, keys, k, lookup, cont, last, first, strand)
            assert got == brute(read, k, lookup)
            total_walked += walked
            total += len(got)
    if total == 0:
        pytest.skip("no unique k-mers in these references (tiny k, repetitive sequence)")
    assert total_walked > 0                                  # the walk really was exercised


def test_orientation_is_consistent():
    """cont[o] means exactly: F_o is F_{o-1} shifted by one base; strand says which orientation F_o is."""
    rng = np.random.default_rng(7)
    k = 7
    ref = make_reference(rng, 3000, "uniform")
    keys = build_dictionary(ref, k)
    cont, last, first, strand = orient(keys, k)
    Mlow = (1 << (2 * (k - 1))) - 1
    F = [keys[o] if strand[o] else rc(keys[o], k) for o in range(len(keys))]
    for o in range(1, len(keys)):
        if cont[o]:
            assert (F[o] >> 2) == (F[o - 1] & Mlow) and last[o] == (F[o] & 3) and first[o - 1] == F[o - 1] >> (2 * (k - 1))
        assert o % BLOCK != 0 or cont[o] == 0
    assert sum(cont) > 0.5 * len(keys)                       # a random reference is mostly walkable


# ------------------------------------------------------------------------------------------------
# k != 30.  The reference's key is min(fwd_k, rc_W) with a reverse-complement register of W = 30 bases
# WHATEVER k is (Q.c:415-420).  For k < W the key is the forward k-mer unless the newest bases are all
# T; for k = W + 1 it is the W-base reverse complement unless the oldest base is A.  Both are chains in
# dictionary order (forward: K' = ((K << 2) | b) & mask; reverse: K' = (K >> 2) | comp(b) << 2(W-1)), so the
# same walk applies -- in one direction only, and only over positions whose key TYPE the read itself
# settles.  Modelled here with W = 6 so that coincidences are common.
W = 6


def mixed_stream(codes, k):
    """(key, fwd_k, rc_W, run length) per emitting position: the reference's codec with a W-base rc register."""
    M = (1 << (2 * k)) - 1
    fwd = rcv = 0
    out = []
    for i, c in enumerate(codes):
        fwd = ((fwd << 2) | c) & ((1 << 64) - 1)
        rcv = (rcv | ((c ^ 2) << (2 * W))) >> 2
        if i + 1 >= k:
            f = fwd & M
            out.append((min(f, rcv), f, rcv))
        else:
            out.append(None)
    return out


def build_mixed_dictionary(ref_codes, k):
    stream = [s for s in mixed_stream(ref_codes, k) if s is not None]
    counts = {}
    for key, _, _ in stream:
        counts[key] = counts.get(key, 0) + 1
    return [key for key, _, _ in stream if counts[key] == 1 and key != 0]


def mixed_ext(keys, k):
    """cont / last2 (forward chain, k < W) or cont / top2 (reverse chain, k = W + 1) per ordinal."""
    n = len(keys)
    M = (1 << (2 * k)) - 1
    cont, sym = [0] * n, [0] * n
    for begin in range(0, n, BLOCK):
        for o in range(begin, min(n, begin + BLOCK)):
            if k < W:
                sym[o] = keys[o] & 3
                cont[o] = int(o > begin and (keys[o] >> 2) == (keys[o - 1] & (M >> 2)))
            else:
                sym[o] = keys[o] >> (2 * (W - 1))
                cont[o] = int(o > begin and (keys[o] & ((1 << (2 * (W - 1))) - 1)) == (keys[o - 1] >> 2))
    return cont, sym


def type_certain(read_codes, p, k):
    """What the kernel can tell from the read alone, conservatively: the key at p is the forward k-mer
    (k < W: one of the newest min(k, W - k) bases is not T) / the W-base reverse complement (k = W + 1:
    the oldest of the k bases is not A)."""
    if k < W:
        w = min(k, W - k)
        return any(read_codes[p - i] != 2 for i in range(w))
    return read_codes[p - W] != 0


def mixed_walk_counts(read_codes, keys, k, lookup, cont, sym):
    stream = mixed_stream(read_codes, k)
    n = len(keys)
    got, walked = {}, 0
    L = len(read_codes)
    for base in range(0, L, 2 * RUN):
        end = None
        for half in (0, 1):
            hb = base + RUN * half
            run = [p for p in range(hb, min(hb + RUN, L)) if stream[p] is not None]
            if not run:
                continue
            settled, probed_anchor, oa = {}, None, None
            if half == 1 and end is not None:
                ja, oa = hb - 1, end
            else:
                ja = run[0]
                key, f, r = stream[ja]
                o = lookup.get(key)
                probed_anchor = ja
                if o is not None:
                    settled[ja] = o
                    if (key == f) if k < W else (key == r):      # the anchor's key is of the walkable type (exact compare)
                        oa = o
            if oa is not None:
                o = oa
                for p in range(ja + 1, min(hb + RUN, L)):
                    b = read_codes[p]
                    want = b if k < W else b ^ 2
                    if stream[p] is None or not type_certain(read_codes, p, k) or o + 1 >= n or not cont[o + 1] or sym[o + 1] != want:
                        break
                    o += 1
                    settled[p] = o
                    walked += 1
            for p in run:
                if p in settled:
                    got[p] = settled[p]
                elif p != probed_anchor:
                    o = lookup.get(stream[p][0])
                    if o is not None:
                        got[p] = o
            if half == 0:
                last_p = hb + RUN - 1
                # the second half may carry on only from a position whose key type is known to be the walkable one
                ok = last_p in settled and (last_p != probed_anchor or oa is not None)
                end = settled[last_p] if ok else None
    return got, walked


@pytest.mark.parametrize("flavour", ["uniform", "at_rich", "tandem", "palindromic"])
@pytest.mark.parametrize("k", [2, 3, 4, 5])
def test_mixed_key_walk_equals_lookup(k, flavour):
    rng = np.random.default_rng(77 * k + len(flavour))
    total_walked = total = 0
    for trial in range(12):
        ref = make_reference(rng, 600, flavour)
        keys = build_mixed_dictionary(ref, k)
        if len(keys) < 4:
            continue
        lookup = {key: o for o, key in enumerate(keys)}
        cont, sym = mixed_ext(keys, k)
        for _ in range(40):
            span = min(90, len(ref))
            a = int(rng.integers(0, len(ref) - span + 1))
            read = list(ref[a:a + int(rng.integers(k + 1, span + 1))])
            if rng.integers(0, 3) == 0:
                read = [c ^ 2 for c in reversed(read)]
            for i in rng.integers(0, len(read), int(rng.integers(0, 3))):
                read[i] = int(rng.integers(0, 4))
            got, walked = mixed_walk_counts(read, keys, k, lookup, cont, sym)
            want = {p: lookup[s[0]] for p, s in enumerate(mixed_stream(read, k)) if s is not None and s[0] in lookup}
            assert got == want
            total_walked += walked
            total += len(got)
    if total == 0:
        pytest.skip("no unique k-mers in these references")
    assert total_walked > 0


# ------------------------------------------------------------------------------------------------
# k = W + 1 (the reference's k = 31).  The key is the W-mer of the newest W bases: FORWARD iff the base W positions
# back is A and forward <= reverse complement, else reverse-complemented.  So the dictionary is a sequence of W-mers in
# one of two forms, the orientation pass of the canonical case applies as it is, and a walk step needs one thing more:
# the read's key at that position must have the form the dictionary stored for that ordinal.
def walk_counts_kp1(read_codes, keys, lookup, cont, last, first, strand):
    k = W + 1
    stream = mixed_stream(read_codes, k)
    MW = (1 << (2 * W)) - 1
    n = len(keys)
    L = len(read_codes)

    def form_is_fwd(p):                      # what the kernel computes from the read: A back there and fwd <= rc
        _, f, r = stream[p]
        return f <= r                        # f has its top base in bits 2W+1:2W: nonzero unless that base is A

    got, walked = {}, 0
    for base in range(0, L, 2 * RUN):
        end = None
        for half in (0, 1):
            hb = base + RUN * half
            run = [p for p in range(hb, min(hb + RUN, L)) if stream[p] is not None]
            if not run:
                continue
            settled, probed_anchor, oa = {}, None, None
            if half == 1 and end is not None:
                ja, (oa, plus) = hb - 1, end
            else:
                ja = run[0]
                probed_anchor = ja
                oa = lookup.get(stream[ja][0])
                if oa is not None:
                    settled[ja] = oa
                    plus = bool(strand[oa]) == form_is_fwd(ja)
            if oa is not None:
                o = oa
                for p in range(ja + 1, min(hb + RUN, L)):
                    b = read_codes[p]
                    if stream[p] is None:
                        break
                    if plus:
                        if o + 1 >= n or not cont[o + 1] or last[o + 1] != b or bool(strand[o + 1]) != form_is_fwd(p):
                            break
                        o += 1
                    else:
                        if o - 1 < 0 or not cont[o] or first[o - 1] != (b ^ 2) or bool(strand[o - 1]) == form_is_fwd(p):
                            break
                        o -= 1
                    settled[p] = o
                    walked += 1
            for p in run:
                if p in settled:
                    got[p] = settled[p]
                elif p != probed_anchor:
                    o = lookup.get(stream[p][0])
                    if o is not None:
                        got[p] = o
            if half == 0:
                last_p = hb + RUN - 1
                end = (settled[last_p], plus) if last_p in settled else None
    return got, walked


@pytest.mark.parametrize("flavour", ["uniform", "at_rich", "tandem", "palindromic"])
def test_k31_style_walk_equals_lookup(flavour):
    k = W + 1
    rng = np.random.default_rng(4242 + len(flavour))
    total_walked = total = 0
    for trial in range(14):
        ref = make_reference(rng, 700, flavour)
        keys = build_mixed_dictionary(ref, k)
        if len(keys) < 4:
            continue
        lookup = {key: o for o, key in enumerate(keys)}
        cont, last, first, strand = orient(keys, W)        # the keys are W-mers: the canonical case's orientation pass
        for _ in range(40):
            span = min(90, len(ref))
            a = int(rng.integers(0, len(ref) - span + 1))
            read = list(ref[a:a + int(rng.integers(k + 1, span + 1))])
            if rng.integers(0, 2):
                read = [c ^ 2 for c in reversed(read)]
            for i in rng.integers(0, len(read), int(rng.integers(0, 3))):
                read[i] = int(rng.integers(0, 4))
            got, walked = walk_counts_kp1(read, keys, lookup, cont, last, first, strand)
            want = {p: lookup[s[0]] for p, s in enumerate(mixed_stream(read, k)) if s is not None and s[0] in lookup}
            assert got == want
            total_walked += walked
            total += len(got)
    if total == 0:
        pytest.skip("no unique k-mers in these references")
    assert total_walked > 0                                  # the walk really was exercised
