// qk_count.cu -- the count kernel (codec + probe + depth increment, fused) and the result
// kernels (16-bit narrowing, GC control curve).
//
// Replaces the reference's hot loops:
//   loop A  Q.c:397-456  per-byte codec on the main thread
//   loop B  Q.c:256-296  Find_hash + atomic depth increment in the worker threads
//   loop C  Q.c:498-518  chain-order dump (here the counters are already in chain order)
//
// Position-parallel form of the codec.  A chunk is a run of
// sequence lines, each ending in '\n'.  For byte position p let last(p) be the position of
// the last reset byte ('N' or '\n') at or before p (-1 = chunk start) and r = p - last(p).
// Position p emits iff r > 0 and (r mod 65536) >= k            (Q.c:402,410,418; T7)
//   fwd = 2-bit codes of bytes p-31..p, newest in bits 1:0, masked to 2k bits  (Q.c:412,419)
//   rc  = complemented codes of the last min(r,30) bytes, newest in bits 59:58 (Q.c:414-416)
//   key = min(fwd, rc)                                                         (Q.c:420)
// so one thread per position needs only a 31-byte left halo and r.
//
// Mapping: a CTA owns a contiguous span of 4 KiB tiles and walks it tile by tile, carrying
// last() across tiles (long reads span many tiles); only the span's first tile needs a
// backward search.  Per tile the 256 threads load 16 bytes each (coalesced 128-bit loads,
// next tile prefetched into registers), pack them to 2 bits/base and a reset bitmask in
// shared memory; then lane = position: 32 consecutive positions per warp, so the two code
// words and the mask word a warp reads are shared-memory broadcasts and the depth
// increments of a warp land on consecutive ordinals (one or two 128-byte lines per RED).
// Four positions per thread are in flight at once: four independent 32-byte bucket loads
// (LDG.E.256, no L1 allocation) are issued before any is consumed.
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "qk_common.cuh"

#define QK_WORDS (QK_TILE / 32)       // 32-base code words / 32-byte mask words per tile
#define QK_POS_PER_THREAD (QK_TILE / QK_THREADS)
#define QK_UNROLL 4
#define QK_NONE INT_MIN

int qk_ring_push(qk_ctx *ctx, qk_slot *sl, int kind, qk_timing_pair **out);

struct qk_count_args {
    const uint8_t *bytes;
    uint32_t n_bytes;
    uint32_t n_tiles;
    uint32_t tiles_per_cta;
    uint32_t *counters;
    unsigned long long *stats;
    qk_table_view tv;
};

// ---- byte -> 2-bit code / reset flag packing --------------------------------------------
// four bytes -> 8 bits, first byte in the top pair: ((c>>1)&3 per byte, Q.c:411)
__device__ __forceinline__ uint32_t qk_pack4(uint32_t w)
{
    return (((w >> 1) & 0x03030303u) * 0x40100401u) >> 24;
}
// four bytes -> 4 flags, bit j set iff byte j is 'N' (Q.c:404) or '\n' (Q.c:403)
__device__ __forceinline__ uint32_t qk_reset4(uint32_t w)
{
    uint32_t m = __vcmpeq4(w, 0x4E4E4E4Eu) | __vcmpeq4(w, 0x0A0A0A0Au);
    return ((m & 0x01010101u) * 0x01020408u) >> 24;
}
__device__ __forceinline__ uint32_t qk_codes16(const uint4 &v)
{
    return (qk_pack4(v.x) << 24) | (qk_pack4(v.y) << 16) | (qk_pack4(v.z) << 8) | qk_pack4(v.w);
}
__device__ __forceinline__ uint32_t qk_resets16(const uint4 &v)
{
    return qk_reset4(v.x) | (qk_reset4(v.y) << 4) | (qk_reset4(v.z) << 8) | (qk_reset4(v.w) << 12);
}

// 16 bytes at chunk offset pos (16-byte aligned); bytes at or beyond n read as '\n'.  The
// 128-bit load may touch up to 15 bytes past n: chunk buffers are readable up to the next
// 16-byte boundary (slot buffers are tile-padded; cudaMalloc granularity covers the rest).
__device__ __forceinline__ uint4 qk_load16(const uint8_t *__restrict__ bytes, uint32_t pos, uint32_t n)
{
    uint4 v = make_uint4(0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au);
    if (pos < n) {
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "l"(bytes + pos));
        const uint32_t rem = n - pos;
        if (rem < 16) { // chunk tail: force the bytes past n to '\n'
            uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int keep = (int)rem - 4 * i;
                const uint32_t m = keep >= 4 ? 0xFFFFFFFFu : keep <= 0 ? 0u : ((1u << (8 * keep)) - 1);
                w[i] = (w[i] & m) | (0x0A0A0A0Au & ~m);
            }
            v = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
    return v;
}

// What a lane holds of 16 positions between its load and its use.  ASCII chunks: the 16 bytes, turned into codes and
// reset flags when they are needed (the conversion then overlaps the next load).  PACKED chunks (host/qk_framer_mt.c:
// per 64 positions four 32-bit code words and 64 flags, 24 bytes) hold the two already.
template <bool PACKED> struct qk_in16;
template <> struct qk_in16<false> {
    uint4 v;
    __device__ __forceinline__ uint32_t codes() const { return qk_codes16(v); }
    __device__ __forceinline__ uint32_t resets() const { return qk_resets16(v); }
};
template <> struct qk_in16<true> {
    uint32_t c, r;
    __device__ __forceinline__ uint32_t codes() const { return c; }
    __device__ __forceinline__ uint32_t resets() const { return r; }
};
// the 16 positions at chunk position pos (a multiple of 16); positions at or beyond n read as '\n'
template <bool PACKED> __device__ __forceinline__ qk_in16<PACKED> qk_fetch16(const uint8_t *__restrict__ bytes, uint32_t pos, uint32_t n);
template <> __device__ __forceinline__ qk_in16<false> qk_fetch16<false>(const uint8_t *__restrict__ bytes, uint32_t pos, uint32_t n)
{
    qk_in16<false> x;
    x.v = qk_load16(bytes, pos, n);
    return x;
}
template <> __device__ __forceinline__ qk_in16<true> qk_fetch16<true>(const uint8_t *__restrict__ bytes, uint32_t pos, uint32_t n)
{
    qk_in16<true> x;
    x.c = 0;
    x.r = 0xFFFFu;
    if (pos < n) {   // (n is a multiple of 64: a group is there whole or not at all)
        const uint8_t *g = bytes + (size_t)(pos >> 6) * 24;
        const uint32_t w = (pos >> 4) & 3u;
        unsigned short r16;
        asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(x.c) : "l"(g + 4 * w));
        asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(r16) : "l"(g + 16 + 2 * w));
        x.r = r16;
    }
    return x;
}

// the same load with an L2 fetch-size hint of 64 B: a missing sector then costs 64 B of DRAM
// traffic instead of the default 128 B (measured, profiles/r1_gather_probe_ncu.txt)
__device__ __forceinline__ qk_bucket qk_ld_bucket64(const qk_bucket *p)
{
    qk_bucket v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(v.e[0]), "=l"(v.e[1]), "=l"(v.e[2]), "=l"(v.e[3])
                 : "l"(p));
    return v;
}

__device__ __forceinline__ qk_bucket qk_ld_bucket(const qk_bucket *p)
{
    qk_bucket v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(v.e[0]), "=l"(v.e[1]), "=l"(v.e[2]), "=l"(v.e[3])
                 : "l"(p));
    return v;
}

// reverse the order of the 2-bit groups of a 64-bit word
__device__ __forceinline__ uint64_t qk_rev_pairs(uint64_t x)
{
    uint64_t z = __brevll(x);
    return ((z & 0x5555555555555555ull) << 1) | ((z >> 1) & 0x5555555555555555ull);
}

// stash probe: linear over 16-byte entries; returns ordinal + 1 or 0
// (the strand bit of the entry comes back in bit 0 of *strand)
__device__ __noinline__ uint32_t qk_stash_find(const qk_table_view &tv, uint64_t key, uint32_t *strand = nullptr)
{
    uint64_t s = qk_mix_stash(key) & tv.stash_mask;
    for (;;) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(tv.stash + s));
        const uint64_t sk = ((uint64_t)v.y << 32) | v.x;
        if (sk == (key | QK_STASH_TAKEN)) {
            if (strand) *strand = v.w & 1u;
            return v.z;
        }
        if (sk == 0) return 0;
        s = (s + 1) & tv.stash_mask;
    }
}

__global__ void __launch_bounds__(QK_THREADS, 3) qk_count_kernel(const qk_count_args a)
{
    __shared__ uint64_t s_codes[2][QK_WORDS + 1]; // [0] = halo: the 32 bases before the tile
    __shared__ uint64_t s_rc[2][QK_WORDS + 1];    // the same bases complemented, in reverse order
    __shared__ uint32_t s_mask[2][QK_WORDS];      // reset flags, bit j of word w = byte 32w+j
    __shared__ int s_last[2][QK_WORDS];           // last reset before word w (chunk position)
    __shared__ int s_red[QK_THREADS / 32];
    __shared__ int s_carry;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile0 = blockIdx.x * a.tiles_per_cta;
    if (tile0 >= a.n_tiles) return;
    const uint32_t tile_end = min(tile0 + a.tiles_per_cta, a.n_tiles);
    const uint8_t *__restrict__ bytes = a.bytes;
    const uint32_t n = a.n_bytes;

    // ---- span start: last reset before the span, and the 32-base halo -----------------------
    {
        int found = QK_NONE;
        uint32_t pos = tile0 * QK_TILE;
        while (pos > 0 && found == QK_NONE) {
            pos -= QK_TILE;
            const uint32_t at = pos + tid * 16;
            const uint32_t m = qk_resets16(qk_load16(bytes, at, n));
            int own = m ? (int)(at + 31 - __clz(m)) : QK_NONE;
            for (int o = 16; o; o >>= 1) own = max(own, __shfl_xor_sync(0xffffffffu, own, o));
            if (lane == 0) s_red[warp] = own;
            __syncthreads();
            found = s_red[0];
#pragma unroll
            for (int w = 1; w < QK_THREADS / 32; ++w) found = max(found, s_red[w]);
            __syncthreads();
        }
        if (tid == 0) {
            s_carry = (found == QK_NONE) ? -1 : found;
            uint64_t halo = 0;
            const uint32_t base = tile0 * QK_TILE;
            if (base >= 32)
                halo = ((uint64_t)qk_codes16(qk_load16(bytes, base - 32, n)) << 32) | qk_codes16(qk_load16(bytes, base - 16, n));
            s_codes[(tile0 & 1) ^ 1][QK_WORDS] = halo; // where the "previous tile" leaves its last word
            s_rc[(tile0 & 1) ^ 1][QK_WORDS] = qk_rev_pairs(halo) ^ 0xAAAAAAAAAAAAAAAAull;
        }
    }

    const qk_table_view tv = a.tv;
    const uint32_t k = tv.k;
    const uint64_t rem_mask = ((uint64_t)1 << tv.rem_bits) - 1;
    const uint32_t ord_mask = tv.ord_bits >= 32 ? 0xFFFFFFFFu : (1u << tv.ord_bits) - 1; // ord_bits <= 32
    uint32_t n_emit = 0, n_hit = 0;

    uint4 cur = qk_load16(bytes, tile0 * QK_TILE + tid * 16, n);
    for (uint32_t tile = tile0; tile < tile_end; ++tile) {
        const uint32_t buf = tile & 1;
        const uint32_t base = tile * QK_TILE;
        __syncthreads(); // previous tile's readers are done with buf; s_carry / halo visible
        {
            // forward codes: first base of the word in the top pair; reverse-complement stream:
            // complemented codes ((L-2)&3 = L^2, Q.c:414), first base of the word in the BOTTOM
            // pair -- so both strands of the window ending at p are one funnel shift away
            const uint32_t c = qk_codes16(cur);
            uint32_t z = __brev(c);
            z = ((z & 0x55555555u) << 1) | ((z >> 1) & 0x55555555u);
            reinterpret_cast<uint32_t *>(s_codes[buf])[2 + (tid ^ 1)] = c;
            reinterpret_cast<uint32_t *>(s_rc[buf])[2 + tid] = z ^ 0xAAAAAAAAu;
        }
        reinterpret_cast<uint16_t *>(s_mask[buf])[tid] = (uint16_t)qk_resets16(cur);
        if (tid == 0) {
            s_codes[buf][0] = s_codes[buf ^ 1][QK_WORDS];
            s_rc[buf][0] = s_rc[buf ^ 1][QK_WORDS];
        }
        if (tile + 1 < tile_end) cur = qk_load16(bytes, base + QK_TILE + tid * 16, n); // prefetch
        __syncthreads();
        if (warp == 0) { // exclusive max-scan of the per-word last reset
            int carry = s_carry;
#pragma unroll
            for (int i = 0; i < QK_WORDS / 32; ++i) {
                const uint32_t w = i * 32 + lane;
                const uint32_t m = s_mask[buf][w];
                int incl = m ? (int)(base + w * 32 + 31 - __clz(m)) : QK_NONE;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    int up = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl = max(incl, up);
                }
                int excl = __shfl_up_sync(0xffffffffu, incl, 1);
                if (lane == 0) excl = QK_NONE;
                s_last[buf][w] = max(carry, excl);
                carry = max(carry, __shfl_sync(0xffffffffu, incl, 31));
            }
            if (lane == 0) s_carry = carry;
        }
        __syncthreads();

#pragma unroll 1
        for (int it = 0; it < QK_POS_PER_THREAD / QK_UNROLL; ++it) {
            uint64_t key[QK_UNROLL], q[QK_UNROLL];
            const qk_bucket *bp[QK_UNROLL];
            bool valid[QK_UNROLL];
            uint32_t run[QK_UNROLL];
            qk_bucket bk[QK_UNROLL];
            bool any = false;
#pragma unroll
            for (int u = 0; u < QK_UNROLL; ++u) {
                const uint32_t w = (it * QK_UNROLL + u) * (QK_THREADS / 32) + warp; // word holding p
                const uint32_t p = base + w * 32 + lane;
                const uint32_t m = s_mask[buf][w] & (0xFFFFFFFFu >> (31 - lane));
                const int last = m ? (int)(base + w * 32 + 31 - __clz(m)) : s_last[buf][w];
                run[u] = (uint32_t)((int)p - last);
                valid[u] = p < n && run[u] != 0 && (run[u] & 0xFFFFu) >= k;
                any = any || valid[u];
            }
            // header / quality lines blanked by the device framer, runs of N: nothing to look up
            if (!__any_sync(0xffffffffu, any)) continue;
#pragma unroll
            for (int u = 0; u < QK_UNROLL; ++u) {
                const uint32_t w = (it * QK_UNROLL + u) * (QK_THREADS / 32) + warp;
                const uint32_t r = run[u];
                const uint64_t A = s_codes[buf][w], B = s_codes[buf][w + 1];
                const uint32_t sh = 2 * (31 - lane);
                const uint64_t x = (B >> sh) | ((A << 1) << (63 - sh)); // 32 bases ending at p
                const uint64_t fwd = x & tv.kmask;
                const uint64_t RA = s_rc[buf][w], RB = s_rc[buf][w + 1];
                // 32 complemented bases ending at p, newest in the top pair; >> 4 leaves the 30
                // newest with the newest at bits 59:58 (the reference's 60-bit register)
                uint64_t rc = ((RB << sh) | ((RA >> 1) >> (63 - sh))) >> 4;
                if (k < 30) rc &= ~(((uint64_t)1 << (60 - 2 * min(r, 30u))) - 1); // zero fill below the run
                key[u] = min(fwd, rc);
                const uint64_t h = qk_mix60(key[u]);
                bp[u] = tv.buckets + (h >> tv.rem_bits);
                q[u] = (h & rem_mask) << tv.ord_bits;
            }
#pragma unroll
            for (int u = 0; u < QK_UNROLL; ++u) {
                bk[u].e[0] = bk[u].e[1] = bk[u].e[2] = bk[u].e[3] = 0;
                if (valid[u]) bk[u] = qk_ld_bucket(bp[u]);
            }
#pragma unroll
            for (int u = 0; u < QK_UNROLL; ++u) {
                if (!valid[u]) continue;
                ++n_emit;
                // entry == (rem << ord_bits) | ord1: high words equal (bit 63 aside) and the low
                // words differ only inside the ordinal field, which is then ord1 != 0
                uint32_t ord1 = 0;
                const uint32_t qhi = (uint32_t)(q[u] >> 32), qlo = (uint32_t)q[u];
#pragma unroll
                for (int e = 0; e < QK_BUCKET_ENTRIES; ++e) {
                    const uint32_t dhi = ((uint32_t)(bk[u].e[e] >> 32) ^ qhi) & 0x3FFFFFFFu;
                    const uint32_t dlo = (uint32_t)bk[u].e[e] ^ qlo;
                    if (dhi == 0 && dlo - 1 < ord_mask) ord1 = dlo;
                }
                // a key of this bucket is in the stash only if the bucket says so (bit 62 of its last entry)
                if (ord1 == 0 && (bk[u].e[QK_BUCKET_ENTRIES - 1] & QK_ENTRY_OVERFLOW)) ord1 = qk_stash_find(tv, key[u]);
                if (ord1) {
                    ++n_hit;
                    atomicAdd(a.counters + (ord1 - 1), 1u);
                }
            }
        }
    }

    for (int o = 16; o; o >>= 1) {
        n_emit += __shfl_xor_sync(0xffffffffu, n_emit, o);
        n_hit += __shfl_xor_sync(0xffffffffu, n_hit, o);
    }
    if (lane == 0) {
        atomicAdd(a.stats + 0, (unsigned long long)n_emit);
        atomicAdd(a.stats + 1, (unsigned long long)n_hit);
        atomicAdd(a.stats + 4, (unsigned long long)n_emit); // one bucket probe per emitted k-mer
    }
}


// =============================================================================================
// qk_count_ext_kernel -- the same result with far fewer table probes (k = 30).
//
// The probe is the expensive part: one random 32-byte sector = one DRAM row activation, and
// HBM3e sustains ~42 G of those per second whatever the load instruction (profiles/).  But
// the k-mers of a read are not independent: consecutive positions are consecutive dictionary
// ordinals wherever the read matches the reference.  The build pass (qk_orient_insert_kernel)
// left, per ordinal o, the last/first base of the dictionary k-mer F_o in a walking orientation
// and a bit cont[o] = "F_o is F_{o-1} shifted by one base".  So here a thread owns QK_RUN = 16
// consecutive positions, probes the table for the FIRST emitting one only (the anchor), and
// derives the neighbours from it:
//     anchor k-mer f, key = K_o.  Same strand as F_o (strand bit == (key == fwd)):
//         f shifted by read base b equals F_{o+1}  iff  cont[o+1] and b == last[o+1]
//     opposite strand:
//         f shifted by b equals rc(F_{o-1})        iff  cont[o] and comp(b) == first[o-1]
// i.e. position anchor+i has ordinal o+i (o-i) as long as every step so far held -- an exact
// statement about keys, not a heuristic.  All 15 steps are checked at once with bit-parallel
// compares of 30-bit fields.  Positions after the first failed step (sequencing error, N,
// end of a unique stretch) are probed individually as before; an anchor that misses leaves
// its whole run to individual probes.  Per thread: 1 probe + the failures instead of 16.
#define QK_RUN 16

__device__ __forceinline__ uint32_t qk_rev16pairs(uint32_t c)
{
    uint32_t z = __brev(c);
    return ((z & 0x55555555u) << 1) | ((z >> 1) & 0x55555555u);
}
// The extension array holds, per group of 16 ordinals, three adjacent words: last bases (2 bits per
// ordinal), first bases (2 bits per ordinal), continuation bits (low 16 bits).  A walk reads the fields
// of 15 consecutive ordinals: two groups = 24 contiguous bytes, one L2 line, where three separate arrays
// cost two or three random DRAM sectors per walk.
// 15 two-bit fields starting at ordinal q (ordinal q in bits 1:0); which = 0: last base, 1: first base
__device__ __forceinline__ uint32_t qk_extract30(const uint32_t *__restrict__ ext, uint64_t q, uint32_t which)
{
    const uint32_t *g = ext + (q >> 4) * QK_EXT_GROUP_WORDS + which;
    const uint32_t lo = __ldg(g), hi = __ldg(g + QK_EXT_GROUP_WORDS);
    return __funnelshift_r(lo, hi, 2 * (uint32_t)(q & 15)) & 0x3FFFFFFFu;
}
// 15 continuation bits starting at ordinal q
__device__ __forceinline__ uint32_t qk_extract15(const uint32_t *__restrict__ ext, uint64_t q)
{
    const uint32_t *g = ext + (q >> 4) * QK_EXT_GROUP_WORDS + 2;
    const uint32_t both = (__ldg(g) & 0xFFFFu) | (__ldg(g + QK_EXT_GROUP_WORDS) << 16);
    return (both >> (uint32_t)(q & 15)) & 0x7FFFu;
}

struct qk_probe {
    const qk_bucket *bp;
    uint64_t key, q;
};
__device__ __forceinline__ qk_probe qk_probe_prepare(const qk_table_view &tv, uint64_t key)
{
    qk_probe p;
    const uint64_t h = qk_mix60(key);
    p.key = key;
    p.bp = tv.buckets + (h >> tv.rem_bits);
    p.q = (h & (((uint64_t)1 << tv.rem_bits) - 1)) << tv.ord_bits;
    return p;
}
// ordinal + 1 of the probed key (0 = absent); *strand = strand bit of its entry
__device__ __forceinline__ uint32_t qk_probe_resolve(const qk_table_view &tv, const qk_probe &p, const qk_bucket &bk,
                                                     uint32_t ord_mask, uint32_t *strand)
{
    uint32_t ord1 = 0;
    const uint32_t qhi = (uint32_t)(p.q >> 32), qlo = (uint32_t)p.q;
#pragma unroll
    for (int e = 0; e < QK_BUCKET_ENTRIES; ++e) {
        const uint32_t ehi = (uint32_t)(bk.e[e] >> 32);
        const uint32_t dlo = (uint32_t)bk.e[e] ^ qlo;
        if (((ehi ^ qhi) & 0x3FFFFFFFu) == 0 && dlo - 1 < ord_mask) {
            ord1 = dlo;
            *strand = ehi >> 31;
        }
    }
    // a key of this bucket is in the stash only if the bucket says so (bit 62 of its last entry)
    if (ord1 == 0 && (bk.e[QK_BUCKET_ENTRIES - 1] & QK_ENTRY_OVERFLOW)) ord1 = qk_stash_find(tv, p.key, strand);
    return ord1;
}

// Work decomposition: a WARP owns a span of 512-byte sub-tiles and walks it on its own -- no
// CTA-wide barrier anywhere, so a warp waiting for DRAM never holds up its neighbours (with
// CTA-wide tiles ncu showed as many cycles stalled on the barrier as on memory).  Lane l owns
// bytes [16 l, 16 l + 16) of the sub-tile.  Probes that the walk cannot avoid are pooled in a
// per-warp queue and issued 64 at a time with every lane busy, and the depth increments are made
// position-parallel from a per-warp ordinal array, so that a warp's REDs fall on consecutive
// counters (one or two 128-byte lines per instruction).
#ifndef QK_POOL_UNROLL
#define QK_POOL_UNROLL 2 // pooled probes per lane per round
#endif
#define QK_SUB 512
#define QK_SUB_WORDS (QK_SUB / 32)
#define QK_WARPS (QK_THREADS / 32)

struct qk_warp_smem {
    uint64_t codes[QK_SUB_WORDS + 1]; // [0] = halo: the 32 bases before the sub-tile
    uint32_t mask[QK_SUB_WORDS + 1];  // reset flags; [0] = halo word
    uint32_t pad;
    uint32_t ord[QK_SUB];             // ordinal + 1 per position of the sub-tile (0 = no hit)
    uint16_t queue[QK_SUB];           // positions that need a probe of their own
};

template <int MINB, bool L64>
__global__ void __launch_bounds__(QK_THREADS, MINB) qk_count_ext_kernel(const qk_count_args a)
{
    __shared__ __align__(16) qk_warp_smem s_all[QK_WARPS];
    const uint32_t lane = threadIdx.x & 31;
    qk_warp_smem &sm = s_all[threadIdx.x >> 5];
    const uint32_t FULL = 0xffffffffu;

    const uint32_t n = a.n_bytes;
    const uint32_t n_subs = (n + QK_SUB - 1) / QK_SUB;
    const uint32_t subs_per_warp = a.tiles_per_cta * (QK_TILE / QK_SUB) / QK_WARPS; // = tiles_per_cta: 8 sub-tiles per tile, 8 warps
    const uint32_t sub0 = (blockIdx.x * QK_WARPS + (threadIdx.x >> 5)) * subs_per_warp;
    if (sub0 >= n_subs) return;
    const uint32_t sub_end = min(sub0 + subs_per_warp, n_subs);
    const uint8_t *__restrict__ bytes = a.bytes;

    // ---- span start: last reset before the span (for the 16-bit run counter), halo -------------
    int carry_last;
    uint64_t halo_c = 0;
    uint32_t halo_m = 0xFFFFFFFFu; // before the chunk: as good as resets
    {
        int found = QK_NONE;
        uint32_t pos = sub0 * QK_SUB;
        while (pos > 0 && found == QK_NONE) {
            pos -= QK_SUB;
            const uint32_t at = pos + lane * 16;
            const uint32_t m = qk_resets16(qk_load16(bytes, at, n));
            found = __reduce_max_sync(FULL, m ? (int)(at + 31 - __clz(m)) : QK_NONE);
        }
        carry_last = (found == QK_NONE) ? -1 : found;
        const uint32_t base = sub0 * QK_SUB;
        if (base >= 32) {
            const uint4 h0 = qk_load16(bytes, base - 32, n), h1 = qk_load16(bytes, base - 16, n);
            halo_c = ((uint64_t)qk_codes16(h0) << 32) | qk_codes16(h1);
            halo_m = qk_resets16(h0) | (qk_resets16(h1) << 16);
        }
    }

    const qk_table_view tv = a.tv;
    const uint32_t ord_mask = tv.ord_bits >= 32 ? 0xFFFFFFFFu : (1u << tv.ord_bits) - 1;
    uint32_t n_emit = 0, n_hit = 0, n_ext = 0, n_probe = 0, n_walk = 0;
    const uint32_t w = lane >> 1, half = lane & 1; // the word and the half of it this lane owns

    auto key_at = [&](uint32_t idx, bool *is_fwd) -> uint64_t { // canonical 30-mer ending at position idx of the sub-tile
        const uint64_t A = sm.codes[idx >> 5], B = sm.codes[(idx >> 5) + 1];
        const uint32_t sh = 2 * (31 - (idx & 31));
        const uint64_t x = ((B >> sh) | ((A << 1) << (63 - sh))) & QK_M60;
        const uint64_t rc = (qk_rev_pairs(x) >> 4) ^ 0x0AAAAAAAAAAAAAAAull;
        *is_fwd = x <= rc;
        return min(x, rc);
    };

    uint4 cur = qk_load16(bytes, sub0 * QK_SUB + lane * 16, n);
    for (uint32_t sub = sub0; sub < sub_end; ++sub) {
        const uint32_t base = sub * QK_SUB;
        __syncwarp(); // everybody is done reading the previous sub-tile
        const uint32_t my_codes = qk_codes16(cur); // first base in the top pair
        const uint32_t my_resets = qk_resets16(cur);
        reinterpret_cast<uint32_t *>(sm.codes)[2 + (lane ^ 1)] = my_codes;
        reinterpret_cast<uint16_t *>(sm.mask)[2 + lane] = (uint16_t)my_resets;
        if (lane == 0) {
            sm.codes[0] = halo_c;
            sm.mask[0] = halo_m;
        }
        if (sub + 1 < sub_end) cur = qk_load16(bytes, base + QK_SUB + lane * 16, n); // prefetch
        __syncwarp();

        // ---- which of my 16 positions end a 30-mer: no reset among the 30 bytes ending there ----
        const uint64_t M64 = ((uint64_t)sm.mask[w + 1] << 32) | sm.mask[w];
        uint64_t S = M64 | (M64 << 1);
        S |= S << 2; S |= S << 4; S |= S << 8; S |= S << 14; // bit p: a reset in [p-29, p]
        uint32_t emit = ~(uint32_t)(S >> (32 + 16 * half)) & 0xFFFFu;
        const uint32_t p0 = base + 16 * lane;
        if (emit && p0 + 15 - (uint32_t)carry_last >= 65536u) {
            // uint16 cur_chars (Q.c:402): a position whose run length mod 65,536 is below k emits
            // nothing.  Only lines longer than 65 k get here: find the exact last reset before p0.
            int last0 = carry_last;
            for (int ww = (int)w; ww >= 0 && last0 == carry_last; --ww) {
                uint32_t m = sm.mask[ww + 1];
                if ((uint32_t)ww == w) m &= half ? 0xFFFFu : 0u;
                if (m) last0 = (int)(base + ww * 32 + 31 - __clz(m));
            }
            const uint32_t r0 = (uint32_t)((int)p0 - last0);
            for (uint32_t j = 0; j < QK_RUN; ++j)
                if (((r0 + j) & 0xFFFFu) < 30u) emit &= ~(1u << j);
        }
        {   // last reset seen so far, for the next sub-tile
            const int own = my_resets ? (int)(p0 + 31 - __clz(my_resets)) : QK_NONE;
            carry_last = max(carry_last, __reduce_max_sync(FULL, own));
        }
        n_emit += __popc(emit);

        // ---- anchor: my first emitting position; walk the dictionary order from it ---------------
        const uint32_t ja = emit ? __ffs(emit) - 1 : 0;
        bool a_fwd = false;
        qk_probe ap;
        qk_bucket abk;
        abk.e[0] = abk.e[1] = abk.e[2] = abk.e[3] = 0;
        if (emit) {
            ap = qk_probe_prepare(tv, key_at(16 * lane + ja, &a_fwd));
            abk = L64 ? qk_ld_bucket64(ap.bp) : qk_ld_bucket(ap.bp);
        }
        uint32_t a_strand = 0, a_ord1 = 0, verified = 0;
        bool plus = true;
        if (emit) a_ord1 = qk_probe_resolve(tv, ap, abk, ord_mask, &a_strand);
        n_probe += emit != 0;
        const uint64_t oa = a_ord1 ? a_ord1 - 1 : 0;
        if (a_ord1) {
            const uint32_t nsteps = 15 - ja;
            plus = (a_strand != 0) == a_fwd;
            if (nsteps && (plus || oa >= 15)) {
                ++n_walk;
                // read bases of the steps, step i (position ja + i) in bits 2i-1:2i-2
                uint32_t R = qk_rev16pairs(my_codes) >> (2 * (ja + 1));
                uint32_t D, Cb;
                if (plus) {
                    D = qk_extract30(tv.ext, oa + 1, 0);
                    Cb = qk_extract15(tv.ext, oa + 1);
                } else {
                    D = qk_rev16pairs(qk_extract30(tv.ext, oa - 15, 1)) >> 2;
                    Cb = __brev(qk_extract15(tv.ext, oa - 14)) >> 17;
                    R ^= 0xAAAAAAAAu;
                }
                const uint32_t X = D ^ R;
                const uint32_t mism = (X | (X >> 1)) & 0x15555555u;
                uint32_t len = mism ? (uint32_t)(__ffs(mism) - 1) >> 1 : 15u;
                const uint32_t brk = ~Cb & 0x7FFFu;
                if (brk) len = min(len, (uint32_t)__ffs(brk) - 1);
                const uint32_t rs = (my_resets & 0xFFFFu) >> (ja + 1);
                if (rs) len = min(len, (uint32_t)__ffs(rs) - 1);
                len = min(len, nsteps);
                verified = ((1u << len) - 1) << (ja + 1);
            }
        }
        const uint32_t ve = verified & emit;
        n_ext += __popc(ve);
        {   // ordinal + 1 of my 16 positions as far as known now
            uint32_t o[QK_RUN];
#pragma unroll
            for (uint32_t j = 0; j < QK_RUN; ++j) {
                const uint32_t step = j - ja;
                const uint32_t walked = (uint32_t)(plus ? oa + step : oa - step) + 1;
                o[j] = (ve >> j) & 1u ? walked : (j == ja ? a_ord1 : 0u);
            }
            uint4 *dst = reinterpret_cast<uint4 *>(sm.ord + 16 * lane);
#pragma unroll
            for (int v = 0; v < 4; ++v) dst[v] = make_uint4(o[4 * v], o[4 * v + 1], o[4 * v + 2], o[4 * v + 3]);
        }

        // ---- positions the walk could not settle: pool them over the warp ------------------------
        uint32_t todo = emit & ~verified;
        if (emit) todo &= ~(1u << ja);
        const uint32_t cnt = __popc(todo);
        n_probe += cnt;
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(FULL, incl, o);
            if (lane >= (uint32_t)o) incl += up;
        }
        const uint32_t total = __shfl_sync(FULL, incl, 31);
        {
            uint32_t off = incl - cnt;
            while (todo) {
                sm.queue[off++] = (uint16_t)(16 * lane + __ffs(todo) - 1);
                todo &= todo - 1;
            }
        }
        __syncwarp();
        for (uint32_t q0 = 0; q0 < total; q0 += 32 * QK_POOL_UNROLL) {
            qk_probe pr[QK_POOL_UNROLL];
            qk_bucket bk[QK_POOL_UNROLL];
            uint32_t idx[QK_POOL_UNROLL];
            bool on[QK_POOL_UNROLL];
#pragma unroll
            for (int u = 0; u < QK_POOL_UNROLL; ++u) {
                const uint32_t e = q0 + 32 * u + lane;
                on[u] = e < total;
                idx[u] = on[u] ? sm.queue[e] : 0;
                bool f;
                pr[u] = qk_probe_prepare(tv, key_at(idx[u], &f));
                bk[u].e[0] = bk[u].e[1] = bk[u].e[2] = bk[u].e[3] = 0;
            }
#pragma unroll
            for (int u = 0; u < QK_POOL_UNROLL; ++u)
                if (on[u]) bk[u] = L64 ? qk_ld_bucket64(pr[u].bp) : qk_ld_bucket(pr[u].bp);
#pragma unroll
            for (int u = 0; u < QK_POOL_UNROLL; ++u) {
                if (!on[u]) continue;
                uint32_t st;
                sm.ord[idx[u]] = qk_probe_resolve(tv, pr[u], bk[u], ord_mask, &st);
            }
        }
        __syncwarp();

        // ---- depth increments, position-parallel: consecutive lanes, consecutive counters --------
#pragma unroll 4
        for (uint32_t it = 0; it < QK_SUB / 32; ++it) {
            const uint32_t o1 = sm.ord[it * 32 + lane];
            if (o1) {
                ++n_hit;
                atomicAdd(a.counters + (o1 - 1), 1u);
            }
        }
        halo_c = sm.codes[QK_SUB_WORDS];
        halo_m = sm.mask[QK_SUB_WORDS];
    }

    for (int o = 16; o; o >>= 1) {
        n_emit += __shfl_xor_sync(FULL, n_emit, o);
        n_hit += __shfl_xor_sync(FULL, n_hit, o);
        n_ext += __shfl_xor_sync(FULL, n_ext, o);
        n_probe += __shfl_xor_sync(FULL, n_probe, o);
        n_walk += __shfl_xor_sync(FULL, n_walk, o);
    }
    if (lane == 0) {
        atomicAdd(a.stats + 0, (unsigned long long)n_emit);
        atomicAdd(a.stats + 1, (unsigned long long)n_hit);
        atomicAdd(a.stats + 2, (unsigned long long)n_ext);
        atomicAdd(a.stats + 4, (unsigned long long)n_probe);
        atomicAdd(a.stats + 6, (unsigned long long)n_walk);
    }
}

// =============================================================================================
// qk_count_ext32_kernel -- the walk carried over 32 positions per lane.
//
// What bounds the kernel at human scale is the number of random L2 -> DRAM requests (ncu: 37 G
// requests/s = the rate of the random-sector micro-benchmark over a 32 GiB table), so the way to go
// faster is to issue fewer: a lane owns 32 consecutive positions (two halves of 16); the second
// half starts from the ordinal its first half's walk ENDED on -- no probe -- and walks 16 more
// steps; only when the first half's walk did not reach its last position does the second half
// probe an anchor of its own.  Where a read follows the dictionary that is one probe and two
// extension-array reads per 32 k-mers instead of per 16.  Everything else is as in
// qk_count_ext_kernel: warp-private 1 KiB sub-tiles, pooled probes for what the walks leave
// open, position-parallel depth increments.
#define QK_SUB2 1024
#define QK_SUB2_WORDS (QK_SUB2 / 32)

struct qk_warp_smem2 {
    uint64_t codes[QK_SUB2_WORDS + 1]; // [0] = halo: the 32 bases before the sub-tile
    uint32_t mask[QK_SUB2_WORDS + 1];  // reset flags; [0] = halo word
    uint32_t pad;
    uint32_t ord[QK_SUB2];             // ordinal + 1 per position (0 = no hit), rows of 32 swizzled by 16-byte column
    uint16_t queue[QK_SUB2];           // positions that need a probe of their own
};

// where position idx of the sub-tile lives in ord[]: row = owning lane, 16-byte columns XOR-swizzled by the
// row so that the lanes' uint4 stores (row stride 128 B) and the position-parallel reads are both conflict-free
__device__ __forceinline__ uint32_t qk_ord_slot(uint32_t idx)
{
    const uint32_t row = idx >> 5, j = idx & 31;
    return row * 32 + ((((j >> 2) ^ (row & 7)) << 2) | (j & 3));
}

// 16 two-bit fields starting at ordinal q (ordinal q in bits 1:0); which = 0: last base, 1: first base
__device__ __forceinline__ uint32_t qk_extract32(const uint32_t *__restrict__ ext, uint64_t q, uint32_t which)
{
    const uint32_t *g = ext + (q >> 4) * QK_EXT_GROUP_WORDS + which;
    return __funnelshift_r(__ldg(g), __ldg(g + QK_EXT_GROUP_WORDS), 2 * (uint32_t)(q & 15));
}
// 16 continuation bits starting at ordinal q
__device__ __forceinline__ uint32_t qk_extract16(const uint32_t *__restrict__ ext, uint64_t q)
{
    const uint32_t *g = ext + (q >> 4) * QK_EXT_GROUP_WORDS + 2;
    const uint32_t both = (__ldg(g) & 0xFFFFu) | (__ldg(g + QK_EXT_GROUP_WORDS) << 16);
    return (both >> (uint32_t)(q & 15)) & 0xFFFFu;
}
// 16 strand bits (F_o == K_o) starting at ordinal q: the high halves of the same words
__device__ __forceinline__ uint32_t qk_extract16s(const uint32_t *__restrict__ ext, uint64_t q)
{
    const uint32_t *g = ext + (q >> 4) * QK_EXT_GROUP_WORDS + 2;
    const uint32_t both = (__ldg(g) >> 16) | (__ldg(g + QK_EXT_GROUP_WORDS) & 0xFFFF0000u);
    return (both >> (uint32_t)(q & 15)) & 0xFFFFu;
}

// Walk from position ja (-1 = the position just before this half) with ordinal oa over the 16 positions of a
// half: c16 = their codes (first base in the top pair), resets16 = their reset flags.  Returns the positions
// (bits 0..15) whose k-mer is PROVEN to be dictionary ordinal oa +- (j - ja).
//   MODE 0 (k = 30, canonical 30-mers): on strand `plus` ordinals go up and the read base is compared with the last
//          base of the next dictionary k-mer; on the other strand they go down and its complement is compared with
//          the first base of the previous one.
//   MODE 1 (k < 30, keys = forward k-mers): ordinals go up, read base against the last base of the next key.
//   MODE 2 (k = 31): the keys are 30-mers too -- the newest 30 bases -- stored forward iff the 31st base back is A and
//          forward <= reverse complement, else reverse-complemented (Q.c:415-420 with a 60-bit register).  The walk
//          is MODE 0's over the 30-mers, in both directions; a step also needs the read's key to have the FORM the
//          dictionary stored for that ordinal: form16 = per position, "the read's key is the forward 30-mer".
// In mode 1 the walk may only step onto a position whose key is certainly the forward k-mer -- unc16 flags (even bit
// of each pair, first position in the top pair) the positions where the read itself cannot settle that; they end
// the walk and are probed like any other open position.
template <int MODE>
__device__ __forceinline__ uint32_t qk_walk16(const qk_table_view &tv, uint64_t oa, bool plus, int ja, uint32_t c16, uint32_t resets16,
                                              uint32_t unc16)
{
    const uint32_t nsteps = (uint32_t)(15 - ja);          // 0..16
    if (nsteps == 0 || !(plus || oa >= 16)) return 0;
    const uint32_t sh = (uint32_t)(ja + 1);               // 0..15
    uint32_t R = qk_rev16pairs(c16) >> (2 * sh);          // read base of step i (position ja + i) in bits 2i-1 : 2i-2
    uint32_t D, Cb, St = 0;
    if (MODE == 1 || plus) {
        D = qk_extract32(tv.ext, oa + 1, 0);
        Cb = qk_extract16(tv.ext, oa + 1);
        if (MODE == 2) St = qk_extract16s(tv.ext, oa + 1);            // step i: strand bit of ordinal oa + i in bit i - 1
    } else {
        D = qk_rev16pairs(qk_extract32(tv.ext, oa - 16, 1));
        Cb = __brev(qk_extract16(tv.ext, oa - 15)) >> 16;
        if (MODE == 2) St = ~(__brev(qk_extract16s(tv.ext, oa - 16)) >> 16); // ordinal oa - i; on this strand the forms swap
        R ^= 0xAAAAAAAAu;
    }
    const uint32_t X = D ^ R;
    const uint32_t mism = (X | (X >> 1)) & 0x55555555u;
    uint32_t len = mism ? (uint32_t)(__ffs(mism) - 1) >> 1 : 16u;
    const uint32_t brk = ~Cb & 0xFFFFu;
    if (brk) len = min(len, (uint32_t)__ffs(brk) - 1);
    const uint32_t rs = (resets16 & 0xFFFFu) >> sh;
    if (rs) len = min(len, (uint32_t)__ffs(rs) - 1);
    if (MODE == 1) {
        const uint32_t U = (qk_rev16pairs(unc16) >> (2 * sh)) & 0x55555555u;
        if (U) len = min(len, (uint32_t)(__ffs(U) - 1) >> 1);
    }
    if (MODE == 2) {   // unc16 here: bit j = the read's key at position j is the FORWARD 30-mer; it must be the stored form
        const uint32_t bad = ((unc16 >> sh) ^ St) & 0xFFFFu;
        if (bad) len = min(len, (uint32_t)__ffs(bad) - 1);
    }
    len = min(len, nsteps);
    return ((1u << len) - 1) << sh;
}

template <int MINB, bool L64, int MODE, bool PACKED>
__global__ void __launch_bounds__(QK_THREADS, MINB) qk_count_ext32_kernel(const qk_count_args a)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    qk_warp_smem2 *s_all = reinterpret_cast<qk_warp_smem2 *>(s_raw);
    const uint32_t lane = threadIdx.x & 31;
    qk_warp_smem2 &sm = s_all[threadIdx.x >> 5];
    const uint32_t FULL = 0xffffffffu;

    const uint32_t n = a.n_bytes;
    const uint32_t n_subs = (n + QK_SUB2 - 1) / QK_SUB2;
    const uint32_t subs_per_warp = max(1u, a.tiles_per_cta * (QK_TILE / QK_SUB2) / QK_WARPS);
    const uint32_t sub0 = (blockIdx.x * QK_WARPS + (threadIdx.x >> 5)) * subs_per_warp;
    if (sub0 >= n_subs) return;
    const uint32_t sub_end = min(sub0 + subs_per_warp, n_subs);
    const uint8_t *__restrict__ bytes = a.bytes;

    // ---- span start: last reset before the span (for the 16-bit run counter), halo -------------
    int carry_last;
    uint64_t halo_c = 0;
    uint32_t halo_m = 0xFFFFFFFFu; // before the chunk: as good as resets
    {
        int found = QK_NONE;
        uint32_t pos = sub0 * QK_SUB2;
        while (pos > 0 && found == QK_NONE) {
            pos -= 512;
            const uint32_t at = pos + lane * 16;
            const uint32_t m = qk_fetch16<PACKED>(bytes, at, n).resets();
            found = __reduce_max_sync(FULL, m ? (int)(at + 31 - __clz(m)) : QK_NONE);
        }
        carry_last = (found == QK_NONE) ? -1 : found;
        const uint32_t base = sub0 * QK_SUB2;
        if (base >= 32) {
            const qk_in16<PACKED> h0 = qk_fetch16<PACKED>(bytes, base - 32, n), h1 = qk_fetch16<PACKED>(bytes, base - 16, n);
            halo_c = ((uint64_t)h0.codes() << 32) | h1.codes();
            halo_m = h0.resets() | (h1.resets() << 16);
        }
    }

    const qk_table_view tv = a.tv;
    const uint32_t ord_mask = tv.ord_bits >= 32 ? 0xFFFFFFFFu : (1u << tv.ord_bits) - 1;
    uint32_t n_emit = 0, n_hit = 0, n_ext = 0, n_probe = 0, n_walk = 0;

    const uint32_t k = tv.k;
    // key of the k-mer ending at position idx of the sub-tile (Q.c:412-420); *is_fwd = it is the forward k-mer
    auto key_at = [&](uint32_t idx, bool *is_fwd) -> uint64_t {
        const uint64_t A = sm.codes[idx >> 5], B = sm.codes[(idx >> 5) + 1];
        const uint32_t sh = 2 * (31 - (idx & 31));
        const uint64_t x32 = (B >> sh) | ((A << 1) << (63 - sh));   // the 32 bases ending at idx
        uint64_t rc = (qk_rev_pairs(x32 & QK_M60) >> 4) ^ 0x0AAAAAAAAAAAAAAAull;
        if (MODE == 0) {
            const uint64_t x = x32 & QK_M60;
            *is_fwd = x <= rc;
            return min(x, rc);
        }
        if (MODE == 1) {   // k < 30: the register holds min(run length, 30) bases, zero-filled below (Q.c:414-416)
            const uint64_t M = ((uint64_t)sm.mask[(idx >> 5) + 1] << 32) | sm.mask[idx >> 5];
            const uint32_t win = (uint32_t)(M >> ((idx & 31) + 3)) & 0x3FFFFFFFu;   // reset flags of the 30 bytes ending at idx
            if (win) rc &= ~(((uint64_t)1 << (2 * (32 - __clz(win)) )) - 1);        // run = 29 - top flag: keep the top `run` bases
        }
        const uint64_t f = x32 & tv.kmask;
        *is_fwd = f <= rc;
        return min(f, rc);
    };

    // the warp loads 1 KiB as two fully coalesced 512-byte rows; ownership (32 contiguous positions per lane)
    // is taken from shared memory afterwards
    qk_in16<PACKED> cur0 = qk_fetch16<PACKED>(bytes, sub0 * QK_SUB2 + lane * 16, n), cur1 = qk_fetch16<PACKED>(bytes, sub0 * QK_SUB2 + 512 + lane * 16, n);
    for (uint32_t sub = sub0; sub < sub_end; ++sub) {
        const uint32_t base = sub * QK_SUB2;
        __syncwarp(); // everybody is done reading the previous sub-tile
        {
            uint32_t *c32 = reinterpret_cast<uint32_t *>(sm.codes);
            uint16_t *m16 = reinterpret_cast<uint16_t *>(sm.mask);
            const uint32_t r0 = cur0.resets(), r1 = cur1.resets();
            c32[2 + (lane ^ 1)] = cur0.codes();                // first base of a 32-base word in its top pair
            c32[2 + 32 + (lane ^ 1)] = cur1.codes();
            m16[2 + lane] = (uint16_t)r0;
            m16[2 + 32 + lane] = (uint16_t)r1;
            if (lane == 0) {
                sm.codes[0] = halo_c;
                sm.mask[0] = halo_m;
            }
            // last reset seen so far, for the next sub-tile
            const int own0 = r0 ? (int)(base + 16 * lane + 31 - __clz(r0)) : QK_NONE;
            const int own1 = r1 ? (int)(base + 512 + 16 * lane + 31 - __clz(r1)) : QK_NONE;
            const int seen = __reduce_max_sync(FULL, max(own0, own1));
            if (sub + 1 < sub_end) {
                cur0 = qk_fetch16<PACKED>(bytes, base + QK_SUB2 + lane * 16, n); // prefetch
                cur1 = qk_fetch16<PACKED>(bytes, base + QK_SUB2 + 512 + lane * 16, n);
            }
            __syncwarp();
            // ---- which of my 32 positions end a 30-mer: no reset among the 30 bytes ending there ----
            const uint32_t my_mask = sm.mask[lane + 1];
            const uint64_t M64 = ((uint64_t)my_mask << 32) | sm.mask[lane];
            uint64_t S = M64 | (M64 << 1);
            if (MODE == 0) { S |= S << 2; S |= S << 4; S |= S << 8; S |= S << 14; } // bit p: a reset in [p-29, p]
            else
                for (uint32_t have = 2; have < k;) {                                 // bit p: a reset in [p-k+1, p]
                    const uint32_t step = min(have, k - have);
                    S |= S << step;
                    have += step;
                }
            uint32_t emit = ~(uint32_t)(S >> 32);
            const uint32_t p0 = base + 32 * lane;
            if (emit && p0 + 31 - (uint32_t)carry_last >= 65536u) {
                // uint16 cur_chars (Q.c:402): a position whose run length mod 65,536 is below k emits nothing.
                // Only lines longer than 65 k get here: find the exact last reset before p0.
                int last0 = carry_last;
                for (int ww = (int)lane - 1; ww >= 0 && last0 == carry_last; --ww) {
                    const uint32_t m = sm.mask[ww + 1];
                    if (m) last0 = (int)(base + ww * 32 + 31 - __clz(m));
                }
                const uint32_t run0 = (uint32_t)((int)p0 - last0);
                for (uint32_t j = 0; j < 32; ++j)   // (a reset inside my own word restarts the run: no wrap there)
                    if (!(my_mask & ((2u << j) - 1)) && ((run0 + j) & 0xFFFFu) < k) emit &= ~(1u << j);
            }
            carry_last = max(carry_last, seen);
            n_emit += __popc(emit);

            const uint64_t W = sm.codes[lane + 1];      // my 32 bases
            uint64_t unc = 0;                           // modes 1, 2: positions whose key type the read does not settle
            if (MODE == 1) {   // the key is the forward k-mer unless the newest min(k, 30 - k) bases are all T (code 2)
                const uint64_t Wp = sm.codes[lane];
                uint64_t lo = (W >> 1) & ~W & 0x5555555555555555ull, hi = (Wp >> 1) & ~Wp & 0x5555555555555555ull;
                const uint32_t need = min(k, 30u - k);
                for (uint32_t have = 1; have < need;) {   // AND over a window of `need` bases; older bases sit at higher bits
                    const uint32_t step = min(have, need - have), sft = 2 * step;
                    lo &= (lo >> sft) | (hi << (64 - sft));
                    hi &= hi >> sft;
                    have += step;
                }
                unc = lo;
            }
            uint32_t form32 = 0;      // mode 2: bit j = the key at my position j is the FORWARD 30-mer (else its reverse complement)
            if (MODE == 2) {          // that needs an A 30 positions back AND forward <= reverse complement (Q.c:420)
                const uint64_t Wp = sm.codes[lane];
                const uint64_t lo = ~(W | (W >> 1)) & 0x5555555555555555ull, hi = ~(Wp | (Wp >> 1)) & 0x5555555555555555ull;
                uint64_t isA = (lo >> 60) | (hi << 4);    // pair-indexed, first position in the top pair: the base 30 back is A
                uint32_t cand = 0;                        // the same, position-indexed
#pragma unroll
                for (int j = 0; j < 32; ++j) cand |= (uint32_t)((isA >> (2 * (31 - j))) & 1u) << j;
                cand &= emit;
                while (cand) {
                    const uint32_t j = __ffs(cand) - 1;
                    cand &= cand - 1;
                    bool f;
                    key_at(32 * lane + j, &f);
                    form32 |= (uint32_t)f << j;
                }
            }
            uint32_t verified = 0, anchors = 0;
            // ---- first half: anchor = my first emitting position; walk the dictionary order from it -------
            uint32_t e0 = emit & 0xFFFFu, e1 = emit >> 16;
            uint64_t end_ord = 0;       // ordinal of position 15 when the first half's walk (or anchor) settled it
            bool end_known = false, end_plus = false;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const uint32_t eh = half ? e1 : e0;
                const uint32_t c16 = half ? (uint32_t)W : (uint32_t)(W >> 32);
                const uint32_t u16 = MODE == 2 ? (form32 >> (16 * half)) & 0xFFFFu : half ? (uint32_t)unc : (uint32_t)(unc >> 32);
                const uint32_t r16 = (my_mask >> (16 * half)) & 0xFFFFu;
                int ja = -1;
                uint32_t a_ord1 = 0;
                uint64_t oa = 0;
                bool plus = true, walkable = true;
                if (half == 1 && end_known) {               // carry on from where the first half ended: no probe
                    oa = end_ord;
                    plus = end_plus;
                    a_ord1 = 1;                             // (only "known" matters below)
                } else if (eh) {
                    ja = __ffs(eh) - 1;
                    bool a_fwd = false;
                    uint32_t a_strand = 0;
                    const qk_probe ap = qk_probe_prepare(tv, key_at(32 * lane + 16 * half + ja, &a_fwd));
                    const qk_bucket abk = L64 ? qk_ld_bucket64(ap.bp) : qk_ld_bucket(ap.bp);
                    a_ord1 = qk_probe_resolve(tv, ap, abk, ord_mask, &a_strand);
                    ++n_probe;
                    anchors |= 1u << (16 * half + ja);
                    oa = a_ord1 ? a_ord1 - 1 : 0;
                    if (MODE != 1) plus = (a_strand != 0) == a_fwd;   // the read runs along F_o iff its key has F_o's form
                    else walkable = a_fwd;                            // mode 1: the anchor's key is the forward k-mer (exact compare)
                }
                uint32_t ve = 0;
                if (a_ord1 && eh && walkable) {
                    ve = qk_walk16<MODE>(tv, oa, plus, ja, c16, r16, u16) & eh;
                    n_walk += ja < 15;
                }
                n_ext += __popc(ve);
                verified |= ve << (16 * half);
                {   // ordinal + 1 of the 16 positions of this half as far as known now
                    uint32_t o[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int step = j - ja;
                        const uint32_t walked = (uint32_t)(plus ? oa + step : oa - step) + 1;
                        o[j] = (ve >> j) & 1u ? walked : (j == ja ? a_ord1 : 0u);
                    }
#pragma unroll
                    for (int v = 0; v < 4; ++v)
                        *reinterpret_cast<uint4 *>(sm.ord + lane * 32 + (((4 * half + v) ^ (lane & 7)) << 2)) =
                            make_uint4(o[4 * v], o[4 * v + 1], o[4 * v + 2], o[4 * v + 3]);
                    if (half == 0) {
                        end_known = (e0 >> 15) & 1u && o[15] != 0 && (walkable || ja != 15);
                        end_plus = plus;
                        end_ord = (uint64_t)o[15] - 1;
                    }
                }
            }

            // ---- positions the walks could not settle: pool them over the warp ------------------------
            uint32_t todo = emit & ~verified & ~anchors;
            const uint32_t cnt = __popc(todo);
            n_probe += cnt;
            uint32_t incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t up = __shfl_up_sync(FULL, incl, o);
                if (lane >= (uint32_t)o) incl += up;
            }
            const uint32_t total = __shfl_sync(FULL, incl, 31);
            {
                uint32_t off = incl - cnt;
                while (todo) {
                    sm.queue[off++] = (uint16_t)(32 * lane + __ffs(todo) - 1);
                    todo &= todo - 1;
                }
            }
            __syncwarp();
            for (uint32_t q0 = 0; q0 < total; q0 += 32 * QK_POOL_UNROLL) {
                qk_probe pr[QK_POOL_UNROLL];
                qk_bucket bk[QK_POOL_UNROLL];
                uint32_t idx[QK_POOL_UNROLL];
                bool on[QK_POOL_UNROLL];
#pragma unroll
                for (int u = 0; u < QK_POOL_UNROLL; ++u) {
                    const uint32_t e = q0 + 32 * u + lane;
                    on[u] = e < total;
                    idx[u] = on[u] ? sm.queue[e] : 0;
                    bool f;
                    pr[u] = qk_probe_prepare(tv, key_at(idx[u], &f));
                    bk[u].e[0] = bk[u].e[1] = bk[u].e[2] = bk[u].e[3] = 0;
                }
#pragma unroll
                for (int u = 0; u < QK_POOL_UNROLL; ++u)
                    if (on[u]) bk[u] = L64 ? qk_ld_bucket64(pr[u].bp) : qk_ld_bucket(pr[u].bp);
#pragma unroll
                for (int u = 0; u < QK_POOL_UNROLL; ++u) {
                    if (!on[u]) continue;
                    uint32_t st;
                    sm.ord[qk_ord_slot(idx[u])] = qk_probe_resolve(tv, pr[u], bk[u], ord_mask, &st);
                }
            }
            __syncwarp();

            // ---- depth increments, position-parallel: consecutive lanes, consecutive counters --------
#pragma unroll 4
            for (uint32_t it = 0; it < QK_SUB2 / 32; ++it) {
                const uint32_t o1 = sm.ord[it * 32 + ((((lane >> 2) ^ (it & 7)) << 2) | (lane & 3))];
                if (o1) {
                    ++n_hit;
                    atomicAdd(a.counters + (o1 - 1), 1u);
                }
            }
            halo_c = sm.codes[QK_SUB2_WORDS];
            halo_m = sm.mask[QK_SUB2_WORDS];
        }
    }

    for (int o = 16; o; o >>= 1) {
        n_emit += __shfl_xor_sync(FULL, n_emit, o);
        n_hit += __shfl_xor_sync(FULL, n_hit, o);
        n_ext += __shfl_xor_sync(FULL, n_ext, o);
        n_probe += __shfl_xor_sync(FULL, n_probe, o);
        n_walk += __shfl_xor_sync(FULL, n_walk, o);
    }
    if (lane == 0) {
        atomicAdd(a.stats + 0, (unsigned long long)n_emit);
        atomicAdd(a.stats + 1, (unsigned long long)n_hit);
        atomicAdd(a.stats + 2, (unsigned long long)n_ext);
        atomicAdd(a.stats + 4, (unsigned long long)n_probe);
        atomicAdd(a.stats + 6, (unsigned long long)n_walk);
    }
}

static int qk_table_view_of(qk_ctx *ctx, qk_table_view *tv)
{
    if (ctx->dict_state != 2) return qk_fail(ctx, QK_ERR_STATE, "no dictionary built on this context");
    const qk_table_desc *d = &ctx->desc;
    tv->buckets = ctx->buckets;
    tv->stash = ctx->stash;
    tv->stash_mask = d->stash_slots - 1;
    tv->kmask = (d->k >= 32) ? 0 : (((uint64_t)1 << (2 * d->k)) - 1); // Q.c:419 on x86-64
    tv->k = d->k;
    tv->rem_bits = d->rem_bits;
    tv->ord_bits = d->ord_bits;
    tv->has_stash = d->stash_used != 0;
    tv->ext = d->has_ext ? ctx->ext : NULL;
    tv->n_kmers = d->n_kmers;
    return QK_OK;
}

static int qk_launch_count_as(qk_ctx *ctx, qk_slot *sl, const uint8_t *dev_bytes, size_t n_bytes, bool packed);
int qk_launch_count(qk_ctx *ctx, qk_slot *sl, const uint8_t *dev_bytes, size_t n_bytes)
{
    return qk_launch_count_as(ctx, sl, dev_bytes, n_bytes, false);
}

// n_bytes = positions of the chunk; packed: dev_bytes holds them as 24 bytes per 64 (qk_fetch16<true>)
static int qk_launch_count_as(qk_ctx *ctx, qk_slot *sl, const uint8_t *dev_bytes, size_t n_bytes, bool packed)
{
    qk_count_args a;
    int rc = qk_table_view_of(ctx, &a.tv);
    if (rc) return rc;
    a.bytes = dev_bytes;
    a.n_bytes = (uint32_t)n_bytes;
    a.n_tiles = (uint32_t)((n_bytes + QK_TILE - 1) / QK_TILE);
    static int tiles_env = -1;
    if (tiles_env < 0) {
        const char *e = getenv("QK_TILES_PER_CTA");
        tiles_env = e ? atoi(e) : 0;
    }
    // classic kernel: 3 CTAs/SM, 4 waves; extension kernel: 4 CTAs/SM, 2 waves (longer warp spans
    // amortise the span-start search; measured +3 %)
    uint32_t target_ctas = (uint32_t)ctx->sm_count * (a.tv.ext ? 8 : 12);
    uint32_t tpc = (a.n_tiles + target_ctas - 1) / target_ctas;
    if (tpc < 4) tpc = 4;
    tpc = (tpc + 1) & ~1u;       // whole 1 KiB sub-tiles per warp (8 warps per CTA, 4 KiB tiles)
    if (tiles_env > 0) tpc = ((uint32_t)tiles_env + 1) & ~1u;
    a.tiles_per_cta = tpc;
    a.counters = ctx->counters;
    a.stats = ctx->stats;
    const uint32_t grid = (a.n_tiles + tpc - 1) / tpc;
    qk_timing_pair *tp;
    rc = qk_ring_push(ctx, sl, 1, &tp);
    if (rc) return rc;
    QK_CUDA(ctx, cudaEventRecord(tp->a, sl->stream));
    static int classic = -1; // QK_CLASSIC_KERNEL=1: probe every position even when the extension arrays exist
    if (classic < 0) classic = getenv("QK_CLASSIC_KERNEL") != NULL;
    static int plain_loads = -1; // QK_EXT_PLAIN_LOADS=1: bucket loads without the .L2::64B hint (A/B knob, -2 %)
    if (plain_loads < 0) plain_loads = getenv("QK_EXT_PLAIN_LOADS") != NULL;
    static int run16 = -1;       // QK_EXT_RUN16=1: the 16-positions-per-lane walk (A/B knob)
    if (run16 < 0) run16 = getenv("QK_EXT_RUN16") != NULL;
    if (packed && !(a.tv.ext && ctx->desc.has_ext))
        return qk_fail(ctx, QK_ERR_STATE, "packed chunks are read by the dictionary-order kernel only (3 <= k <= 31)");
    if (packed || (a.tv.ext && !classic && (!run16 || ctx->desc.has_ext != 1))) {
        const size_t smem = QK_WARPS * sizeof(qk_warp_smem2);
        static int attr_set[64];
        if (ctx->device < 64 && !attr_set[ctx->device]) {
            QK_CUDA(ctx, cudaFuncSetAttribute(qk_count_ext32_kernel<4, true, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            QK_CUDA(ctx, cudaFuncSetAttribute(qk_count_ext32_kernel<4, false, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            QK_CUDA(ctx, cudaFuncSetAttribute(qk_count_ext32_kernel<4, true, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            QK_CUDA(ctx, cudaFuncSetAttribute(qk_count_ext32_kernel<4, true, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            QK_CUDA(ctx, cudaFuncSetAttribute(qk_count_ext32_kernel<4, true, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            QK_CUDA(ctx, cudaFuncSetAttribute(qk_count_ext32_kernel<4, true, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            QK_CUDA(ctx, cudaFuncSetAttribute(qk_count_ext32_kernel<4, true, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set[ctx->device] = 1;
        }
        if (packed) {
            if (ctx->desc.has_ext == 2) qk_count_ext32_kernel<4, true, 1, true><<<grid, QK_THREADS, smem, sl->stream>>>(a);
            else if (ctx->desc.has_ext == 3) qk_count_ext32_kernel<4, true, 2, true><<<grid, QK_THREADS, smem, sl->stream>>>(a);
            else qk_count_ext32_kernel<4, true, 0, true><<<grid, QK_THREADS, smem, sl->stream>>>(a);
        } else if (ctx->desc.has_ext == 2) qk_count_ext32_kernel<4, true, 1, false><<<grid, QK_THREADS, smem, sl->stream>>>(a);
        else if (ctx->desc.has_ext == 3) qk_count_ext32_kernel<4, true, 2, false><<<grid, QK_THREADS, smem, sl->stream>>>(a);
        else if (plain_loads) qk_count_ext32_kernel<4, false, 0, false><<<grid, QK_THREADS, smem, sl->stream>>>(a);
        else qk_count_ext32_kernel<4, true, 0, false><<<grid, QK_THREADS, smem, sl->stream>>>(a);
    } else if (a.tv.ext && ctx->desc.has_ext == 1 && !classic) {
        // 4 CTAs/SM at 64 registers: 5 and 6 CTAs/SM spill and measured 2-5 % slower (profiles/README.md)
        if (plain_loads) qk_count_ext_kernel<4, false><<<grid, QK_THREADS, 0, sl->stream>>>(a);
        else qk_count_ext_kernel<4, true><<<grid, QK_THREADS, 0, sl->stream>>>(a);
    } else qk_count_kernel<<<grid, QK_THREADS, 0, sl->stream>>>(a);
    QK_CUDA(ctx, cudaGetLastError());
    QK_CUDA(ctx, cudaEventRecord(tp->b, sl->stream));
    ctx->launches++;
    return QK_OK;
}

extern "C" int qk_submit(qk_ctx *ctx, uint32_t slot, const uint8_t *bytes, size_t n_bytes, const uint32_t *line_off,
                         uint32_t n_lines)
{
    (void)line_off;
    if (!ctx || slot >= ctx->n_slots || (!bytes && n_bytes)) return QK_ERR_ARG;
    if (n_bytes > ctx->chunk_capacity) return qk_fail(ctx, QK_ERR_ARG, "chunk of %zu bytes exceeds the slot capacity", n_bytes);
    if (ctx->dict_state != 2) return qk_fail(ctx, QK_ERR_STATE, "no dictionary built on this context");
    ctx->lines += n_lines;
    if (n_bytes == 0) return QK_OK;
    QK_CUDA(ctx, cudaSetDevice(ctx->device));
    qk_slot *sl = &ctx->slots[slot];
    qk_timing_pair *tp;
    int rc = qk_ring_push(ctx, sl, 0, &tp);
    if (rc) return rc;
    QK_CUDA(ctx, cudaEventRecord(tp->a, sl->stream));
    QK_CUDA(ctx, cudaMemcpyAsync(sl->dev, bytes, n_bytes, cudaMemcpyHostToDevice, sl->stream));
    QK_CUDA(ctx, cudaEventRecord(tp->b, sl->stream));
    QK_CUDA(ctx, cudaEventRecord(sl->h2d_done, sl->stream));
    return qk_launch_count(ctx, sl, sl->dev, n_bytes);
}

extern "C" int qk_submit_device(qk_ctx *ctx, uint32_t slot, const uint8_t *dev_bytes, size_t n_bytes)
{
    if (!ctx || slot >= ctx->n_slots || !dev_bytes) return QK_ERR_ARG;
    if (((uintptr_t)dev_bytes & 15) != 0) return qk_fail(ctx, QK_ERR_ARG, "device chunk must be 16-byte aligned");
    if (n_bytes >= ((size_t)1 << 31)) return qk_fail(ctx, QK_ERR_ARG, "device chunk must be < 2^31 bytes");
    if (n_bytes == 0) return QK_OK;
    QK_CUDA(ctx, cudaSetDevice(ctx->device));
    return qk_launch_count(ctx, &ctx->slots[slot], dev_bytes, n_bytes);
}

// A packed chunk (host/qk_framer_mt.c: per 64 positions of the framed stream four 32-bit words of 2-bit codes and 64
// reset flags, 24 bytes): 0.375 bytes per position over the link instead of 1, and nothing for the kernel to convert.
extern "C" int qk_submit_packed(qk_ctx *ctx, uint32_t slot, const uint8_t *packed, size_t n_positions, uint32_t n_lines)
{
    if (!ctx || slot >= ctx->n_slots || (!packed && n_positions) || (n_positions & 63)) return QK_ERR_ARG;
    if (n_positions > ctx->chunk_capacity) return qk_fail(ctx, QK_ERR_ARG, "chunk of %zu positions exceeds the slot capacity", n_positions);
    if (ctx->dict_state != 2) return qk_fail(ctx, QK_ERR_STATE, "no dictionary built on this context");
    if (!ctx->desc.has_ext) return qk_fail(ctx, QK_ERR_STATE, "packed chunks are read by the dictionary-order kernel only (3 <= k <= 31)");
    ctx->lines += n_lines;
    if (n_positions == 0) return QK_OK;
    QK_CUDA(ctx, cudaSetDevice(ctx->device));
    qk_slot *sl = &ctx->slots[slot];
    qk_timing_pair *tp;
    int rc = qk_ring_push(ctx, sl, 0, &tp);
    if (rc) return rc;
    QK_CUDA(ctx, cudaEventRecord(tp->a, sl->stream));
    QK_CUDA(ctx, cudaMemcpyAsync(sl->dev, packed, n_positions / 64 * 24, cudaMemcpyHostToDevice, sl->stream));
    QK_CUDA(ctx, cudaEventRecord(tp->b, sl->stream));
    QK_CUDA(ctx, cudaEventRecord(sl->h2d_done, sl->stream));
    return qk_launch_count_as(ctx, sl, sl->dev, n_positions, true);
}

extern "C" int qk_counters_device_ptr(const qk_ctx *ctx, uint32_t **counters, uint64_t *n_kmers)
{
    if (!ctx) return QK_ERR_ARG;
    if (ctx->dict_state != 2) return QK_ERR_STATE;
    if (counters) *counters = ctx->counters;
    if (n_kmers) *n_kmers = ctx->desc.n_kmers;
    return QK_OK;
}

extern "C" int qk_reset_counters(qk_ctx *ctx)
{
    if (!ctx) return QK_ERR_ARG;
    if (ctx->dict_state != 2) return QK_ERR_STATE;
    int rc = qk_sync(ctx);
    if (rc) return rc;
    QK_CUDA(ctx, cudaMemset(ctx->counters, 0, (ctx->desc.n_kmers + 1) * sizeof(uint32_t)));
    QK_CUDA(ctx, cudaMemset(ctx->stats, 0, QK_STATS_WORDS * sizeof(unsigned long long)));
    QK_CUDA(ctx, cudaMemset(ctx->frame_stream + 1, 0, 3 * sizeof(unsigned long long))); // totals, not the line state
    QK_CUDA(ctx, cudaDeviceSynchronize()); // the slot streams do not order against stream 0
    ctx->lines = 0;
    ctx->kernel_ms = ctx->h2d_ms = 0;
    ctx->launches = 0;
    return QK_OK;
}

// n more occurrences of one canonical key, as if the reads had held them.  What the command needs it for:
// with -t N the reference pads its last, partly filled batch of 4,096 keys with zeros and its workers
// look those up like any key (Q.c:458-466, 284-291); Find_hash(0) "finds" the first empty slot on its
// path (Q.c:98), so an empty slot that is ON the chain -- key 0 of an `index` list -- receives the padding.
__global__ void qk_add_depth_kernel(const qk_table_view tv, uint32_t *counters, uint64_t key, uint32_t n)
{
    const qk_probe p = qk_probe_prepare(tv, key);
    const qk_bucket bk = qk_ld_bucket(p.bp);
    const uint32_t ord_mask = tv.ord_bits >= 32 ? 0xFFFFFFFFu : (1u << tv.ord_bits) - 1;
    uint32_t strand;
    const uint32_t ord1 = qk_probe_resolve(tv, p, bk, ord_mask, &strand);
    if (ord1) atomicAdd(counters + (ord1 - 1), n);
}

extern "C" int qk_add_depth(qk_ctx *ctx, uint64_t key, uint32_t n)
{
    if (!ctx) return QK_ERR_ARG;
    qk_table_view tv;
    int rc = qk_table_view_of(ctx, &tv);
    if (rc) return rc;
    if (key >> QK_KEY_BITS) return QK_OK; // no read produces it: never in the table
    rc = qk_sync(ctx);
    if (rc) return rc;
    QK_CUDA(ctx, cudaSetDevice(ctx->device));
    qk_add_depth_kernel<<<1, 1, 0, ctx->slots[0].stream>>>(tv, ctx->counters, key, n);
    QK_CUDA(ctx, cudaGetLastError());
    QK_CUDA(ctx, cudaStreamSynchronize(ctx->slots[0].stream));
    return QK_OK;
}

// Two counter buffers, so that the reduce / download of one job overlaps the counting of the
// next (samples run back to back against one dictionary).  Selecting a buffer affects the
// launches, resets and downloads issued AFTER the call; work already enqueued keeps its buffer.
extern "C" int qk_counters_select(qk_ctx *ctx, uint32_t which)
{
    if (!ctx || which > 1) return QK_ERR_ARG;
    if (ctx->dict_state != 2) return QK_ERR_STATE;
    if (!ctx->counters_buf[which]) {
        QK_CUDA(ctx, cudaSetDevice(ctx->device));
        QK_CUDA(ctx, cudaMalloc((void **)&ctx->counters_buf[which], (ctx->desc.n_kmers + 1) * sizeof(uint32_t)));
        QK_CUDA(ctx, cudaMemset(ctx->counters_buf[which], 0, (ctx->desc.n_kmers + 1) * sizeof(uint32_t)));
        QK_CUDA(ctx, cudaDeviceSynchronize());
    }
    ctx->counters = ctx->counters_buf[which];
    return QK_OK;
}

// Stream-ordered reset for back-to-back jobs: zeroes the counters and the device totals on
// slot 0's stream after joining every other slot stream into it, and makes the other slots
// wait for it -- no host synchronisation.  (Host-side timing accumulators keep running.)
extern "C" int qk_reset_counters_async(qk_ctx *ctx)
{
    if (!ctx) return QK_ERR_ARG;
    if (ctx->dict_state != 2) return QK_ERR_STATE;
    QK_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s0 = ctx->slots[0].stream;
    for (uint32_t s = 1; s < ctx->n_slots; ++s) {
        QK_CUDA(ctx, cudaEventRecord(ctx->span_join, ctx->slots[s].stream));
        QK_CUDA(ctx, cudaStreamWaitEvent(s0, ctx->span_join, 0));
    }
    QK_CUDA(ctx, cudaMemsetAsync(ctx->counters, 0, (ctx->desc.n_kmers + 1) * sizeof(uint32_t), s0));
    QK_CUDA(ctx, cudaMemsetAsync(ctx->stats, 0, QK_STATS_WORDS * sizeof(unsigned long long), s0));
    QK_CUDA(ctx, cudaMemsetAsync(ctx->frame_stream + 1, 0, 3 * sizeof(unsigned long long), s0));
    QK_CUDA(ctx, cudaEventRecord(ctx->span_join, s0));
    for (uint32_t s = 1; s < ctx->n_slots; ++s) QK_CUDA(ctx, cudaStreamWaitEvent(ctx->slots[s].stream, ctx->span_join, 0));
    ctx->lines = 0;
    return QK_OK;
}

extern "C" int qk_counters_download(qk_ctx *ctx, uint64_t offset, uint32_t *out, uint64_t count)
{
    if (!ctx || !out) return QK_ERR_ARG;
    if (ctx->dict_state != 2) return qk_fail(ctx, QK_ERR_STATE, "no dictionary built on this context");
    if (offset + count > ctx->desc.n_kmers) return qk_fail(ctx, QK_ERR_ARG, "counter range outside the dictionary");
    int rc = qk_sync(ctx);
    if (rc) return rc;
    QK_CUDA(ctx, cudaSetDevice(ctx->device));
    QK_CUDA(ctx, cudaMemcpy(out, ctx->counters + offset, count * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return QK_OK;
}

// ---- results ------------------------------------------------------------------------------
// uint32 counter -> the reference's uint16 depth: wraps mod 65,536, never saturates (T12)
__global__ void qk_narrow_kernel(const uint32_t *__restrict__ counters, uint16_t *__restrict__ out, uint64_t n)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = (uint16_t)(counters[i] & 0xFFFFu);
}

// D2H of the depths, piece by piece: narrow on the device, copy into one of two pinned staging
// buffers, hand the piece to the consumer while the next one is in flight (a direct pageable
// cudaMemcpy runs at a few GB/s).  qk_finish with a pinned destination skips the staging.
#define QK_FINISH_PIECE ((uint64_t)16 << 20) // entries per staged piece (32 MiB)
extern "C" int qk_finish_pieces(qk_ctx *ctx, qk_piece_fn consume, void *user)
{
    if (!ctx || !consume) return QK_ERR_ARG;
    if (ctx->dict_state != 2) return qk_fail(ctx, QK_ERR_STATE, "no dictionary built on this context");
    const uint64_t n_kmers = ctx->desc.n_kmers;
    int rc = qk_sync(ctx);
    if (rc) return rc;
    QK_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->slots[0].stream;
    if (!ctx->narrow_dev) QK_CUDA(ctx, cudaMalloc((void **)&ctx->narrow_dev, 2 * QK_FINISH_PIECE * sizeof(uint16_t)));
    if (!ctx->narrow_host)
        QK_CUDA(ctx, cudaHostAlloc((void **)&ctx->narrow_host, 2 * QK_FINISH_PIECE * sizeof(uint16_t), cudaHostAllocDefault));
    cudaEvent_t done[2];
    QK_CUDA(ctx, cudaEventCreateWithFlags(&done[0], cudaEventDisableTiming));
    QK_CUDA(ctx, cudaEventCreateWithFlags(&done[1], cudaEventDisableTiming));
    cudaError_t e = cudaSuccess;
    uint64_t prev_at = 0, prev_m = 0;
    int b = 0, crc = 0;
    for (uint64_t at = 0; at < n_kmers && e == cudaSuccess && !crc; at += QK_FINISH_PIECE, b ^= 1) {
        const uint64_t m = n_kmers - at < QK_FINISH_PIECE ? n_kmers - at : QK_FINISH_PIECE;
        uint16_t *dev = ctx->narrow_dev + (uint64_t)b * QK_FINISH_PIECE;
        qk_narrow_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(ctx->counters + at, dev, m);
        e = cudaMemcpyAsync(ctx->narrow_host + (uint64_t)b * QK_FINISH_PIECE, dev, m * sizeof(uint16_t), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaEventRecord(done[b], st);
        if (prev_m && e == cudaSuccess) { // consume the other buffer while this piece is in flight
            e = cudaEventSynchronize(done[b ^ 1]);
            if (e == cudaSuccess) crc = consume(user, ctx->narrow_host + (uint64_t)(b ^ 1) * QK_FINISH_PIECE, prev_at, prev_m);
        }
        prev_at = at;
        prev_m = m;
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (prev_m && e == cudaSuccess && !crc) crc = consume(user, ctx->narrow_host + (uint64_t)(b ^ 1) * QK_FINISH_PIECE, prev_at, prev_m);
    cudaEventDestroy(done[0]);
    cudaEventDestroy(done[1]);
    if (e != cudaSuccess) return qk_cuda_fail(ctx, e, "D2H of counts");
    return crc;
}

static int qk_finish_copy(void *user, const uint16_t *piece, uint64_t offset, uint64_t count)
{
    memcpy((uint16_t *)user + offset, piece, count * sizeof(uint16_t));
    return QK_OK;
}

extern "C" int qk_finish(qk_ctx *ctx, uint16_t *counts_out, uint64_t n_kmers)
{
    if (!ctx || !counts_out) return QK_ERR_ARG;
    if (ctx->dict_state != 2) return qk_fail(ctx, QK_ERR_STATE, "no dictionary built on this context");
    if (n_kmers != ctx->desc.n_kmers) return qk_fail(ctx, QK_ERR_ARG, "n_kmers does not match the dictionary");
    cudaPointerAttributes attr;
    const bool pinned = cudaPointerGetAttributes(&attr, counts_out) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (!pinned) return qk_finish_pieces(ctx, qk_finish_copy, counts_out);
    int rc = qk_sync(ctx);
    if (rc) return rc;
    QK_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->slots[0].stream;
    if (!ctx->narrow_dev) QK_CUDA(ctx, cudaMalloc((void **)&ctx->narrow_dev, 2 * QK_FINISH_PIECE * sizeof(uint16_t)));
    int b = 0;
    for (uint64_t at = 0; at < n_kmers; at += QK_FINISH_PIECE, b ^= 1) { // stream order keeps the two device buffers safe
        const uint64_t m = n_kmers - at < QK_FINISH_PIECE ? n_kmers - at : QK_FINISH_PIECE;
        uint16_t *dev = ctx->narrow_dev + (uint64_t)b * QK_FINISH_PIECE;
        qk_narrow_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(ctx->counters + at, dev, m);
        QK_CUDA(ctx, cudaMemcpyAsync(counts_out + at, dev, m * sizeof(uint16_t), cudaMemcpyDeviceToHost, st));
    }
    QK_CUDA(ctx, cudaStreamSynchronize(st));
    return QK_OK;
}

// The same without blocking the caller: the download of the counter buffer selected NOW is enqueued on a
// stream of its own, after everything enqueued on the slot streams so far, so that the next job -- counting
// into the OTHER counter buffer (qk_counters_select) -- overlaps it.  counts_out must be page-locked.
extern "C" int qk_finish_async(qk_ctx *ctx, uint16_t *counts_out, uint64_t n_kmers)
{
    if (!ctx || !counts_out) return QK_ERR_ARG;
    if (ctx->dict_state != 2) return qk_fail(ctx, QK_ERR_STATE, "no dictionary built on this context");
    if (n_kmers != ctx->desc.n_kmers) return qk_fail(ctx, QK_ERR_ARG, "n_kmers does not match the dictionary");
    if (!qk_host_is_pinned(counts_out)) return qk_fail(ctx, QK_ERR_ARG, "qk_finish_async needs page-locked host memory");
    int rc = qk_finish_wait(ctx);
    if (rc) return rc;
    QK_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->finish_stream;
    for (uint32_t s = 0; s < ctx->n_slots; ++s) {
        QK_CUDA(ctx, cudaEventRecord(ctx->span_join, ctx->slots[s].stream));
        QK_CUDA(ctx, cudaStreamWaitEvent(st, ctx->span_join, 0));
    }
    if (!ctx->narrow_dev) QK_CUDA(ctx, cudaMalloc((void **)&ctx->narrow_dev, 2 * QK_FINISH_PIECE * sizeof(uint16_t)));
    const uint32_t *src = ctx->counters;
    int b = 0;
    for (uint64_t at = 0; at < n_kmers; at += QK_FINISH_PIECE, b ^= 1) { // stream order keeps the two device buffers safe
        const uint64_t m = n_kmers - at < QK_FINISH_PIECE ? n_kmers - at : QK_FINISH_PIECE;
        uint16_t *dev = ctx->narrow_dev + (uint64_t)b * QK_FINISH_PIECE;
        qk_narrow_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(src + at, dev, m);
        QK_CUDA(ctx, cudaMemcpyAsync(counts_out + at, dev, m * sizeof(uint16_t), cudaMemcpyDeviceToHost, st));
    }
    QK_CUDA(ctx, cudaEventRecord(ctx->finish_done, st));
    ctx->finish_pending = 1;
    return QK_OK;
}

extern "C" int qk_finish_wait(qk_ctx *ctx)
{
    if (!ctx) return QK_ERR_ARG;
    if (!ctx->finish_pending) return QK_OK;
    QK_CUDA(ctx, cudaEventSynchronize(ctx->finish_done));
    ctx->finish_pending = 0;
    return QK_OK;
}

// GC control curve (Q.c:501-508).  Per CTA: 32-bit shared histograms over <= 16384 entries
// (depth <= 65535 so no partial sum overflows), flushed with 64-bit global atomics.  The
// reference squares in `int` (Q.c:507), i.e. the product wraps to negative above 46,340:
// sumsq = sum(u) - 2^32 * #(u >= 2^31) with u the unsigned product.
#define QK_GC_PER_CTA 16384
__global__ void __launch_bounds__(256) qk_gc_kernel(const uint32_t *__restrict__ counters, const uint16_t *__restrict__ qgc,
                                                    uint64_t n, unsigned long long *sum, long long *sumsq,
                                                    unsigned long long *count, unsigned long long *big)
{
    __shared__ uint32_t h_cnt[QK_GC_BINS], h_sum[QK_GC_BINS], h_lo[QK_GC_BINS], h_hi[QK_GC_BINS], h_neg[QK_GC_BINS];
    for (int i = threadIdx.x; i < QK_GC_BINS; i += blockDim.x) h_cnt[i] = h_sum[i] = h_lo[i] = h_hi[i] = h_neg[i] = 0;
    __syncthreads();
    const uint64_t begin = (uint64_t)blockIdx.x * QK_GC_PER_CTA;
    const uint64_t end = begin + QK_GC_PER_CTA < n ? begin + QK_GC_PER_CTA : n;
    for (uint64_t i = begin + threadIdx.x; i < end; i += blockDim.x) {
        const uint32_t g = qgc[i];
        if (!(g & 0x8000u)) continue;
        const uint32_t bin = g & 0x1FFu;
        if (bin >= QK_GC_BINS) {                  // the reference indexes past its arrays here (Q.c:504-507): dropped, and counted
            if (big) atomicAdd(big, 1ull);
            continue;
        }
        const uint32_t d = counters[i] & 0xFFFFu;
        const uint32_t u = d * d;
        atomicAdd(&h_cnt[bin], 1u);
        atomicAdd(&h_sum[bin], d);
        atomicAdd(&h_lo[bin], u & 0xFFFFu);
        atomicAdd(&h_hi[bin], u >> 16);
        if (u >> 31) atomicAdd(&h_neg[bin], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < QK_GC_BINS; i += blockDim.x) {
        if (!h_cnt[i]) continue;
        atomicAdd(&count[i], (unsigned long long)h_cnt[i]);
        atomicAdd(&sum[i], (unsigned long long)h_sum[i]);
        long long sq = (long long)h_lo[i] + ((long long)h_hi[i] << 16) - ((long long)h_neg[i] << 32);
        atomicAdd(reinterpret_cast<unsigned long long *>(&sumsq[i]), (unsigned long long)sq);
    }
}

extern "C" int qk_gc_curve(qk_ctx *ctx, const uint16_t *qgc, uint64_t n_kmers, uint64_t sum[QK_GC_BINS],
                           int64_t sumsq[QK_GC_BINS], uint64_t count[QK_GC_BINS])
{
    if (!ctx || !qgc || !sum || !sumsq || !count) return QK_ERR_ARG;
    if (ctx->dict_state != 2) return qk_fail(ctx, QK_ERR_STATE, "no dictionary built on this context");
    if (n_kmers != ctx->desc.n_kmers) return qk_fail(ctx, QK_ERR_ARG, "n_kmers does not match the dictionary");
    int rc = qk_sync(ctx);
    if (rc) return rc;
    QK_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t piece = (uint64_t)64 << 20;
    uint16_t *dq = NULL;
    unsigned long long *acc = NULL;
    QK_CUDA(ctx, cudaMalloc((void **)&dq, (n_kmers < piece ? n_kmers : piece) * sizeof(uint16_t)));
    cudaError_t e = cudaMalloc((void **)&acc, 3 * QK_GC_BINS * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(acc, 0, 3 * QK_GC_BINS * sizeof(unsigned long long));
    for (uint64_t at = 0; e == cudaSuccess && at < n_kmers; at += piece) {
        const uint64_t m = n_kmers - at < piece ? n_kmers - at : piece;
        e = cudaMemcpy(dq, qgc + at, m * sizeof(uint16_t), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) break;
        qk_gc_kernel<<<(unsigned)((m + QK_GC_PER_CTA - 1) / QK_GC_PER_CTA), 256>>>(
            ctx->counters + at, dq, m, acc, reinterpret_cast<long long *>(acc + QK_GC_BINS), acc + 2 * QK_GC_BINS, nullptr);
        e = cudaGetLastError();
    }
    unsigned long long host[3 * QK_GC_BINS];
    if (e == cudaSuccess) e = cudaMemcpy(host, acc, sizeof host, cudaMemcpyDeviceToHost);
    cudaFree(dq);
    cudaFree(acc);
    if (e != cudaSuccess) return qk_cuda_fail(ctx, e, "GC control curve");
    for (int i = 0; i < QK_GC_BINS; ++i) {
        sum[i] = host[i];
        sumsq[i] = (int64_t)host[QK_GC_BINS + i];
        count[i] = host[2 * QK_GC_BINS + i];
    }
    return QK_OK;
}

// The curve from .qgc pieces that sit in the slots' pinned buffers (the host's reader threads put them there):
// H2D + histogram kernel per piece on the slot's stream, so the 4.5 GB .qgc of a human-scale dictionary goes through
// at ingest speed instead of through one fread and a pageable copy.
extern "C" int qk_gc_begin(qk_ctx *ctx)
{
    if (!ctx) return QK_ERR_ARG;
    if (ctx->dict_state != 2) return qk_fail(ctx, QK_ERR_STATE, "no dictionary built on this context");
    int rc = qk_sync(ctx);
    if (rc) return rc;
    QK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->gc_acc) QK_CUDA(ctx, cudaMalloc((void **)&ctx->gc_acc, (3 * QK_GC_BINS + 1) * sizeof(unsigned long long)));
    QK_CUDA(ctx, cudaMemsetAsync(ctx->gc_acc, 0, (3 * QK_GC_BINS + 1) * sizeof(unsigned long long), ctx->slots[0].stream));
    QK_CUDA(ctx, cudaStreamSynchronize(ctx->slots[0].stream));
    return QK_OK;
}

extern "C" int qk_gc_from_slot(qk_ctx *ctx, uint32_t slot, uint64_t ordinal_offset, uint64_t count)
{
    if (!ctx || slot >= ctx->n_slots || !ctx->gc_acc) return QK_ERR_ARG;
    if (ordinal_offset + count > ctx->desc.n_kmers || count * sizeof(uint16_t) > ctx->chunk_capacity)
        return qk_fail(ctx, QK_ERR_ARG, ".qgc piece outside the dictionary or larger than a slot");
    if (count == 0) return QK_OK;
    QK_CUDA(ctx, cudaSetDevice(ctx->device));
    qk_slot *sl = &ctx->slots[slot];
    QK_CUDA(ctx, cudaMemcpyAsync(sl->dev, sl->host, count * sizeof(uint16_t), cudaMemcpyHostToDevice, sl->stream));
    QK_CUDA(ctx, cudaEventRecord(sl->h2d_done, sl->stream));
    unsigned long long *acc = ctx->gc_acc;
    qk_gc_kernel<<<(unsigned)((count + QK_GC_PER_CTA - 1) / QK_GC_PER_CTA), 256, 0, sl->stream>>>(
        ctx->counters + ordinal_offset, reinterpret_cast<const uint16_t *>(sl->dev), count, acc,
        reinterpret_cast<long long *>(acc + QK_GC_BINS), acc + 2 * QK_GC_BINS, acc + 3 * QK_GC_BINS);
    QK_CUDA(ctx, cudaGetLastError());
    return QK_OK;
}

extern "C" int qk_gc_end(qk_ctx *ctx, uint64_t sum[QK_GC_BINS], int64_t sumsq[QK_GC_BINS], uint64_t count[QK_GC_BINS],
                         uint64_t *bins_out_of_range)
{
    if (!ctx || !sum || !sumsq || !count || !ctx->gc_acc) return QK_ERR_ARG;
    int rc = qk_sync(ctx);
    if (rc) return rc;
    unsigned long long host[3 * QK_GC_BINS + 1];
    QK_CUDA(ctx, cudaMemcpy(host, ctx->gc_acc, sizeof host, cudaMemcpyDeviceToHost));
    for (int i = 0; i < QK_GC_BINS; ++i) {
        sum[i] = host[i];
        sumsq[i] = (int64_t)host[QK_GC_BINS + i];
        count[i] = host[2 * QK_GC_BINS + i];
    }
    if (bins_out_of_range) *bins_out_of_range = host[3 * QK_GC_BINS];
    return QK_OK;
}

