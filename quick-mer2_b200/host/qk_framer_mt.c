/*
 * qk_framer_mt.c -- the record framer of Q.c:393-398, 451-455 on all host cores, feeding any
 * number of GPUs from one shared queue of framed chunks.
 *
 * The reference frames on its single producer thread (fgets, one line at a time) and is bound by
 * it.  Here the input (a byte range in memory: a caller's buffer or an mmap'd file) is cut into
 * fixed-size BLOCKS; worker threads claim blocks in order and, per block,
 *   1. find the lines that START in the block (AVX-512 newline scan, 64 bytes per compare);
 *   2. run the reference's line state machine over them for each of its four entry states
 *      (s = lines still to be discarded; '>' lines skipped; FASTQ: three lines dropped after a read
 *      -- the same machine csrc/qk_frame.cu runs on the device), giving output size and exit
 *      state per entry state;
 *   3. wait for the block before them to be RESOLVED (its exit state and output position known --
 *      an O(1) hand-over, the only sequential step), resolve themselves, and
 *   4. copy their sequence lines -- only those -- into the pinned buffer of the chunk slot they
 *      were given.  FASTQ therefore crosses PCIe at ~1.25 bytes per k-mer instead of ~2.6.
 * A chunk is a slot of ANY of the contexts: the resolver opens the next chunk on whichever GPU
 * has a free slot first, so GPUs with a faster host link simply take more chunks (on an 8-GPU box
 * four of the GPUs get 1.5x the H2D bandwidth of the others when all copy at once).  A chunk is
 * submitted (qk_submit: H2D + count kernel on the slot's stream) by whichever thread finishes
 * its last copy.  Counting is an integer sum, so the result does not depend on which GPU counted
 * which chunk.
 */
#define _GNU_SOURCE
#define _FILE_OFFSET_BITS 64
#include <fcntl.h>
#include <immintrin.h>
#include <pthread.h>
#include <sched.h>
#include <stdatomic.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include "qk_host_internal.h"

#define QK_MT_MAX_CTX 16
#define QK_MT_MAX_THREADS 128
#define QK_PACKED_DEFAULT 1          /* qk_count_mem_mt / qk_count_file_mt ship packed chunks unless QK_PACKED=0 */

enum { SLOT_FREE = 0, SLOT_FILLING = 1, SLOT_SUBMITTING = 2, SLOT_SUBMITTED = 3 };

typedef struct {
    _Atomic int state;          /* SLOT_* */
    _Atomic int pending;        /* blocks still copying into the chunk */
    _Atomic int closed;         /* no further block will be assigned   */
    uint64_t seq;               /* ordinal of the chunk in the framed stream */
    size_t fill;                /* bytes assigned (resolver only)      */
    uint32_t lines;             /* sequence lines assigned             */
} mt_chunk;

typedef struct {
    const uint8_t *data;
    size_t n;
    int fastq;
    size_t block;
    uint64_t n_blocks;
    const qk_chunk_sink *sink;
    uint32_t n_ctx, n_slots;
    size_t cap;
    uint64_t n_chunks;                   /* chunks opened so far (resolver only) */
    pthread_mutex_t submit_mu[QK_MT_MAX_CTX];
    mt_chunk chunk[QK_MT_MAX_CTX][QK_HOST_MAX_SLOTS];
    uint32_t next_slot[QK_MT_MAX_CTX];   /* ring position per context (resolver only) */
    _Atomic uint64_t next_block;
    _Atomic uint64_t resolved;           /* blocks [0, resolved) have their place */
    /* the chain, owned by whoever resolves block `resolved` */
    uint32_t state;
    int cur_ctx, cur_slot, rr;           /* open chunk (-1 = none), round-robin start */
    uint64_t lines, bases, long_lines, unterminated;
    uint64_t sink_bytes;                 /* bytes assigned to chunks so far (resolver only) */
    _Atomic int err;
    int avx512;
    uint16_t tab[2][256];                /* the line state machine from all four entry states at once */
    size_t stage_cap;                    /* bytes of a worker's staging buffer */
    int nt;                              /* staged non-temporal stores (QK_FRAMER_NT=0 turns them off) */
    int populate;                        /* the input is a file mapping: populate each block's pages in one call */
    int packed;                          /* the sink takes packed chunks: 24 bytes per 64 positions (see qk_pack_positions) */
} mt_job;

/* ---- line scan + state machine ------------------------------------------------------------ */
/* one line through the reference's state machine (csrc/qk_frame.cu has the same table) */
static inline int line_step(uint32_t *s, int is_header, int fastq)
{
    if (*s) { *s = (*s + 1) & 3u; return 0; }
    if (is_header) return 0;               /* Q.c:398 */
    if (fastq) *s = 1;                     /* Q.c:451-455: three more lines go */
    return 1;
}

/* The machine run from all four entry states at once: P packs the four current states (2 bits each,
 * trajectory s0 in bits 2 s0 + 1 : 2 s0); tab[is_header][P] = next P | keep mask << 8 (bit s0 of the
 * mask: the trajectory that entered in s0 keeps this line). */
typedef struct {
    const uint16_t (*tab)[256];
    uint32_t P;
    size_t out_bytes[4];
    uint32_t out_lines[4];
    uint64_t *starts;      /* starts[k] = first byte of line k; starts[n_lines] = one past the last line */
    uint8_t *keep;         /* keep mask of line k */
    size_t n_lines;
} mt_sim;

static void build_table(uint16_t tab[2][256], int fastq)
{
    for (int hdr = 0; hdr < 2; ++hdr)
        for (uint32_t P = 0; P < 256; ++P) {
            uint32_t np = 0, k = 0;
            for (uint32_t s0 = 0; s0 < 4; ++s0) {
                uint32_t st = (P >> (2 * s0)) & 3u;
                k |= (uint32_t)line_step(&st, hdr, fastq) << s0;
                np |= st << (2 * s0);
            }
            tab[hdr][P] = (uint16_t)(np | (k << 8));
        }
}

/* the line [ls, le) is complete (le = one past its '\n') */
static inline __attribute__((always_inline)) void sim_line(mt_sim *S, uint8_t first_byte, size_t ls, size_t le)
{
    const uint32_t e = S->tab[first_byte == '>'][S->P];
    const uint32_t k = e >> 8;
    const size_t len = le - ls;
    S->P = e & 0xFFu;
    S->keep[S->n_lines] = (uint8_t)k;
    S->starts[++S->n_lines] = le;
    S->out_bytes[0] += len & ((size_t)0 - (k & 1u));
    S->out_bytes[1] += len & ((size_t)0 - ((k >> 1) & 1u));
    S->out_bytes[2] += len & ((size_t)0 - ((k >> 2) & 1u));
    S->out_bytes[3] += len & ((size_t)0 - ((k >> 3) & 1u));
    S->out_lines[0] += k & 1u;
    S->out_lines[1] += (k >> 1) & 1u;
    S->out_lines[2] += (k >> 2) & 1u;
    S->out_lines[3] += (k >> 3) & 1u;
}

/* Lines from `from` (a line start) on: every '\n' at q in [from, n) completes a line, up to and including the
 * first q >= stop_at.  Returns 1 if the scan stopped at such a q, 0 if it ran into the end of the data (then
 * S->starts[S->n_lines] is where the unfinished rest begins).
 * Two passes so that neither has an unpredictable branch: (A) newline positions of whole 64-byte windows are
 * appended to starts[] two at a time whether or not the window has that many (the cursor advances by the real
 * count; a third and later newline in one window is the rare, predicted-not-taken case); (B) the state machine
 * runs over the finished lines (table look-ups and masked adds). */
__attribute__((target("avx512f,avx512bw,bmi,bmi2,popcnt"))) static int scan_avx512(const uint8_t *d, size_t from, size_t n, size_t stop_at,
                                                                                   mt_sim *S)
{
    const __m512i nl = _mm512_set1_epi8('\n');
    uint64_t *out = S->starts + 1;           /* out[k] = start of line k + 1 = one past the k-th newline */
    size_t p = from;
    int ended = 0;
    /* (A) whole windows strictly before the stop position; then window by window until a newline >= stop_at */
    const size_t bulk_end = stop_at > 64 ? stop_at - 64 : 0;
    while (p + 64 <= n && p < bulk_end) {
        uint64_t m = _mm512_cmpeq_epi8_mask(_mm512_loadu_si512((const void *)(d + p)), nl);
        const unsigned cnt = (unsigned)__builtin_popcountll(m);
        out[0] = p + (size_t)_tzcnt_u64(m) + 1;
        m = _blsr_u64(m);
        out[1] = p + (size_t)_tzcnt_u64(m) + 1;
        if (__builtin_expect(cnt > 2, 0)) {
            m = _blsr_u64(m);
            for (unsigned k = 2; k < cnt; ++k, m = _blsr_u64(m)) out[k] = p + (size_t)_tzcnt_u64(m) + 1;
        }
        out += cnt;
        p += 64;
    }
    while (p < n && !ended) {
        __mmask64 m;
        if (p + 64 <= n) m = _mm512_cmpeq_epi8_mask(_mm512_loadu_si512((const void *)(d + p)), nl);
        else {
            const __mmask64 live = (~0ull) >> (64 - (n - p));
            m = _mm512_cmpeq_epi8_mask(_mm512_maskz_loadu_epi8(live, (const void *)(d + p)), nl) & live;
        }
        while (m) {
            const size_t q = p + (size_t)__builtin_ctzll(m);
            m &= m - 1;
            *out++ = q + 1;
            if (q >= stop_at) { ended = 1; break; }
        }
        p += 64;
    }
    /* (B) -- on locals: the keep bytes may alias anything, so nothing of S is touched inside the loop */
    const size_t lines = (size_t)(out - (S->starts + 1));
    const uint64_t *__restrict__ st = S->starts;
    uint8_t *__restrict__ keep = S->keep;
    const uint16_t(*__restrict__ tab)[256] = S->tab;
    uint32_t P = S->P, l0 = 0, l1 = 0, l2 = 0, l3 = 0;
    size_t b0 = 0, b1 = 0, b2 = 0, b3 = 0;
    size_t ls = st[0];
    for (size_t k = 0; k < lines; ++k) {
        const size_t le = st[k + 1], len = le - ls;
        const uint32_t e = tab[d[ls] == '>'][P];
        const uint32_t kk = e >> 8;
        P = e & 0xFFu;
        keep[k] = (uint8_t)kk;
        b0 += len & ((size_t)0 - (kk & 1u));
        b1 += len & ((size_t)0 - ((kk >> 1) & 1u));
        b2 += len & ((size_t)0 - ((kk >> 2) & 1u));
        b3 += len & ((size_t)0 - ((kk >> 3) & 1u));
        l0 += kk & 1u;
        l1 += (kk >> 1) & 1u;
        l2 += (kk >> 2) & 1u;
        l3 += (kk >> 3) & 1u;
        ls = le;
    }
    S->P = P;
    S->out_bytes[0] += b0; S->out_bytes[1] += b1; S->out_bytes[2] += b2; S->out_bytes[3] += b3;
    S->out_lines[0] += l0; S->out_lines[1] += l1; S->out_lines[2] += l2; S->out_lines[3] += l3;
    S->n_lines = lines;
    return ended;
}

static int scan_plain(const uint8_t *d, size_t from, size_t n, size_t stop_at, mt_sim *S)
{
    size_t ls = from;
    while (ls < n) {
        const uint8_t *e = memchr(d + ls, '\n', n - ls);
        if (!e) break;
        const size_t q = (size_t)(e - d);
        sim_line(S, d[ls], ls, q + 1);
        ls = q + 1;
        if (q >= stop_at) return 1;
    }
    return 0;
}

/* the kept lines of trajectory s_in, back to back at o; returns the number of lines longer than the reference's buffer.
 * The lines are gathered in `stage` (thread-local, cache-resident, co-aligned with o modulo 64) and go to the pinned
 * chunk buffer with non-temporal 64-byte stores: the destination is written once and read by the DMA engine only, so
 * the read-for-ownership a normal store costs (as much DRAM traffic again as the write itself) is saved. */
__attribute__((target("avx512f,avx512bw"))) static inline size_t gather_avx512(uint8_t *w, const uint8_t *d, size_t n, const mt_sim *S,
                                                                                 uint32_t s_in, uint64_t *long_lines)
{
    uint64_t longl = 0;
    uint8_t *const w0 = w;
    for (size_t k = 0; k < S->n_lines; ++k) {
        if (!((S->keep[k] >> s_in) & 1u)) continue;
        const size_t ls = S->starts[k], le = S->starts[k + 1], len = le - ls;
        const uint8_t *src = d + ls;
        if (le > n) {                                            /* unterminated last line: we terminate it (T9) */
            memcpy(w, src, n - ls);
            w[n - ls] = '\n';
        } else if (len >= 64 && len <= 4096) {                   /* whole vectors, the last one flush with the end */
            size_t i = 0;
            for (; i + 64 <= len; i += 64) _mm512_storeu_si512((void *)(w + i), _mm512_loadu_si512((const void *)(src + i)));
            if (i < len) _mm512_storeu_si512((void *)(w + len - 64), _mm512_loadu_si512((const void *)(src + len - 64)));
        } else if (len < 64) {
            const __mmask64 mk = ((uint64_t)1 << len) - 1;
            _mm512_mask_storeu_epi8((void *)w, mk, _mm512_maskz_loadu_epi8(mk, (const void *)src));
        } else memcpy(w, src, len);
        w += len;
        longl += len > QK_MAX_LINE_BYTES;
    }
    *long_lines = longl;
    return (size_t)(w - w0);
}

/* `total` staged bytes at r (co-aligned with o modulo 64) -> destination */
__attribute__((target("avx512f,avx512bw"))) static inline void stream_out(uint8_t *o, const uint8_t *r, size_t total)
{
    size_t at = 0;
    const size_t head = (64 - ((uintptr_t)o & 63)) & 63;
    if (head && total >= head) { memcpy(o, r, head); at = head; }
    for (; at + 64 <= total; at += 64) _mm512_stream_si512((void *)(o + at), _mm512_load_si512((const void *)(r + at)));
    if (at < total) memcpy(o + at, r + at, total - at);
    _mm_sfence();                                                /* before the chunk is handed to the copy engine */
}

__attribute__((target("avx512f,avx512bw"))) static uint64_t copy_avx512(uint8_t *o, const uint8_t *d, size_t n, const mt_sim *S, uint32_t s_in,
                                                                        uint8_t *stage)
{
    uint64_t longl = 0;
    uint8_t *w = stage + ((uintptr_t)o & 63);
    const size_t total = gather_avx512(w, d, n, S, s_in, &longl);
    stream_out(o, w, total);
    return longl;
}

static uint64_t copy_plain(uint8_t *o, const uint8_t *d, size_t n, const mt_sim *S, uint32_t s_in)
{
    uint64_t longl = 0;
    for (size_t k = 0; k < S->n_lines; ++k) {
        if (!((S->keep[k] >> s_in) & 1u)) continue;
        const size_t ls = S->starts[k], le = S->starts[k + 1], len = le - ls;
        if (le > n) {
            memcpy(o, d + ls, n - ls);
            o[n - ls] = '\n';
        } else memcpy(o, d + ls, len);
        o += len;
        longl += len > QK_MAX_LINE_BYTES;
    }
    return longl;
}

/* ---- packed chunks ----------------------------------------------------------------------------
 * What the count kernels make of a framed chunk first is 2 bits per position ((c >> 1) & 3, Q.c:411) and a flag
 * per position (c is 'N' or '\n': the k-mer register starts over, Q.c:403-404).  A PACKED chunk is exactly that,
 * made here: per 64 positions, four little-endian 32-bit code words (16 positions each, the first in the top
 * pair -- qk_codes16 in csrc/qk_count.cu) and the 64 flags (bit p = position p), 24 bytes for 64 -- 0.375 bytes
 * per position over PCIe and out of the host's DRAM twice (written here, read by the copy engine) instead of 1.
 * Every block's output starts on a multiple of 64 positions and is filled up with '\n' positions (empty lines:
 * no k-mers), so that blocks pack independently. */
static void pack_plain(uint8_t *o, const uint8_t *ascii, size_t n_pos)     /* n_pos % 64 == 0 */
{
    for (size_t g = 0; g < n_pos / 64; ++g, o += 24, ascii += 64) {
        uint64_t flags = 0;
        for (int w = 0; w < 4; ++w) {
            uint32_t codes = 0;
            for (int i = 0; i < 16; ++i) {
                const uint8_t c = ascii[16 * w + i];
                codes |= (uint32_t)((c >> 1) & 3u) << (2 * (15 - i));
                flags |= (uint64_t)(c == 'N' || c == '\n') << (16 * w + i);
            }
            memcpy(o + 4 * w, &codes, 4);
        }
        memcpy(o + 16, &flags, 8);
    }
}

__attribute__((target("avx512f,avx512bw,avx512vl"))) static void pack_avx512(uint8_t *o, const uint8_t *ascii, size_t n_pos)
{
    const __m512i three = _mm512_set1_epi8(3), weights = _mm512_set1_epi32(0x01041040), ones = _mm512_set1_epi16(1);
    const __m512i nl = _mm512_set1_epi8('\n'), en = _mm512_set1_epi8('N');
    const __m128i flip = _mm_set_epi8(12, 13, 14, 15, 8, 9, 10, 11, 4, 5, 6, 7, 0, 1, 2, 3);
    for (size_t g = 0; g < n_pos / 64; ++g, o += 24, ascii += 64) {
        const __m512i v = _mm512_loadu_si512((const void *)ascii);
        const __m512i c = _mm512_and_si512(_mm512_srli_epi16(v, 1), three);
        /* four codes -> one byte, the first in the top pair: bytes (c0, c1, c2, c3) . (64, 16, 4, 1) */
        const __m512i quads = _mm512_madd_epi16(_mm512_maddubs_epi16(c, weights), ones);
        const __m128i bytes = _mm_shuffle_epi8(_mm512_cvtepi32_epi8(quads), flip);   /* a code word is little-endian: its first quad last */
        const uint64_t flags = _mm512_cmpeq_epi8_mask(v, nl) | _mm512_cmpeq_epi8_mask(v, en);
        _mm_storeu_si128((__m128i *)o, bytes);
        memcpy(o + 16, &flags, 8);
    }
}

/* the kept lines of trajectory s_in as a piece of a packed chunk at o (pad_pos positions, a multiple of 64): gathered
 * in `stage` as copy_avx512 does, filled up with '\n', packed into `pstage` (co-aligned with o modulo 64), streamed out */
__attribute__((target("avx512f,avx512bw,avx512vl"))) static uint64_t copy_packed_avx512(uint8_t *o, const uint8_t *d, size_t n, const mt_sim *S,
                                                                                       uint32_t s_in, size_t pad_pos, uint8_t *stage,
                                                                                       uint8_t *pstage, int nt)
{
    uint64_t longl = 0;
    const size_t out = gather_avx512(stage, d, n, S, s_in, &longl);
    memset(stage + out, '\n', pad_pos - out);
    uint8_t *r = pstage + ((uintptr_t)o & 63);
    pack_avx512(r, stage, pad_pos);
    if (nt) stream_out(o, r, pad_pos / 64 * 24);
    else memcpy(o, r, pad_pos / 64 * 24);
    return longl;
}

/* ---- chunks --------------------------------------------------------------------------------- */
static void try_submit(mt_job *j, int c, int s)
{
    mt_chunk *ch = &j->chunk[c][s];
    if (!atomic_load(&ch->closed) || atomic_load(&ch->pending) != 0) return;
    int expect = SLOT_FILLING;
    if (!atomic_compare_exchange_strong(&ch->state, &expect, SLOT_SUBMITTING)) return; /* somebody else does */
    pthread_mutex_lock(&j->submit_mu[c]);
    int rc = ch->fill ? j->sink->submit(j->sink->user, (uint32_t)c, (uint32_t)s, ch->seq, ch->fill, ch->lines) : QK_OK;
    pthread_mutex_unlock(&j->submit_mu[c]);
    /* only now may the resolver ask ready()/wait() about this buffer: they speak of the submission just made */
    atomic_store(&ch->state, ch->fill ? SLOT_SUBMITTED : SLOT_FREE);
    if (rc) { int z = 0; atomic_compare_exchange_strong(&j->err, &z, rc); }
}

/* resolver only: close the open chunk */
static void close_chunk(mt_job *j)
{
    if (j->cur_ctx < 0) return;
    mt_chunk *ch = &j->chunk[j->cur_ctx][j->cur_slot];
    atomic_store(&ch->closed, 1);
    try_submit(j, j->cur_ctx, j->cur_slot);
    j->cur_ctx = j->cur_slot = -1;
}

/* resolver only: open a chunk on the first context (round robin) whose next slot is free; if none
 * is, wait -- blocking on the one GPU there is, polling when there are several so that whichever
 * GPU frees a slot first gets the chunk */
static int open_chunk(mt_job *j)
{
    for (int pass = 0;; ++pass) {
        if (atomic_load(&j->err)) return atomic_load(&j->err);
        int took = -1;
        for (uint32_t t = 0; t < j->n_ctx && took < 0; ++t) {
            const int c = (int)((j->rr + t) % j->n_ctx);
            const int s = (int)j->next_slot[c];
            const int st = atomic_load(&j->chunk[c][s].state);
            if (st == SLOT_FILLING || st == SLOT_SUBMITTING) continue; /* its last block is still being copied / handed over */
            if (st == SLOT_SUBMITTED) {
                int ready = j->sink->ready(j->sink->user, (uint32_t)c, (uint32_t)s);
                if (ready < 0) return -ready;
                if (!ready && j->n_ctx == 1 && pass > 0) {      /* nothing else to wait for */
                    int rc = j->sink->wait(j->sink->user, (uint32_t)c, (uint32_t)s);
                    if (rc) return rc;
                    ready = 1;
                }
                if (!ready) continue;
            }
            took = c;
        }
        if (took >= 0) {
            const int c = took, s = (int)j->next_slot[c];
            mt_chunk *ch = &j->chunk[c][s];
            ch->fill = 0;
            ch->lines = 0;
            ch->seq = j->n_chunks++;
            atomic_store(&ch->pending, 0);
            atomic_store(&ch->closed, 0);
            atomic_store(&ch->state, SLOT_FILLING);
            j->next_slot[c] = (uint32_t)((s + 1) % (int)j->n_slots);
            j->cur_ctx = c;
            j->cur_slot = s;
            j->rr = (c + 1) % (int)j->n_ctx;
            return QK_OK;
        }
        if (pass < 16) sched_yield();
        else usleep(20);
    }
}

/* ---- worker ----------------------------------------------------------------------------------- */
typedef struct { mt_job *j; uint64_t *starts; uint8_t *keep, *stage, *pstage; size_t starts_cap; } mt_worker;

static void *worker(void *arg)
{
    mt_worker *w = arg;
    mt_job *j = w->j;
    const uint8_t *d = j->data;
    const size_t n = j->n;
    for (;;) {
        const uint64_t i = atomic_fetch_add(&j->next_block, 1);
        if (i >= j->n_blocks) return NULL;
        const size_t a = (size_t)(i * j->block), b = a + j->block < n ? a + j->block : n;
#ifdef MADV_POPULATE_READ
        {   /* a mapped file: one call maps the block's pages instead of a fault per 4 KiB (harmless on ordinary memory) */
            const uintptr_t pg = (uintptr_t)sysconf(_SC_PAGESIZE) - 1, lo = ((uintptr_t)(d + a)) & ~pg, hi = ((uintptr_t)(d + b)) & ~pg;
            if (j->populate && hi > lo) madvise((void *)lo, hi - lo, MADV_POPULATE_READ);
        }
#endif
        /* 1 + 2. the lines that start in [a, b) -- the first one begins at 0 for block 0, else after the first
         * '\n' at or after a - 1 -- and the state machine over them from each entry state */
        size_t first = 0;
        if (i > 0) {
            const uint8_t *e = memchr(d + a - 1, '\n', n - (a - 1));
            first = e ? (size_t)(e - d) + 1 : n;
        }
        mt_sim S;
        memset(&S, 0, sizeof S);
        S.tab = j->tab;
        S.P = 0xE4u;                                                /* trajectory s0 starts in state s0 */
        S.starts = w->starts;
        S.keep = w->keep;
        int unterminated = 0;
        if (first < b && !atomic_load(&j->err)) {
            S.starts[0] = first;
            const int ended = j->avx512 ? scan_avx512(d, first, n, b - 1, &S) : scan_plain(d, first, n, b - 1, &S);
            if (!ended && S.starts[S.n_lines] < n) {                /* the input ends inside a line (T9): as if a '\n' sat at n */
                sim_line(&S, d[S.starts[S.n_lines]], S.starts[S.n_lines], n + 1);
                unterminated = 1;
            }
        }
        const size_t n_lines = S.n_lines;
        const size_t *out_bytes = S.out_bytes;
        const uint32_t *out_lines = S.out_lines;
        uint32_t exit_state[4], last_kept[4];
        for (uint32_t s0 = 0; s0 < 4; ++s0) {
            exit_state[s0] = (S.P >> (2 * s0)) & 3u;
            last_kept[s0] = n_lines ? (S.keep[n_lines - 1] >> s0) & 1u : 0u;
        }
        /* 3. resolve in order */
        for (int spin = 0; atomic_load_explicit(&j->resolved, memory_order_acquire) != i; ++spin) {
            if (spin < 2000) _mm_pause();
            else sched_yield();
        }
        int rc = atomic_load(&j->err);
        const uint32_t s_in = j->state;
        const size_t out = out_bytes[s_in];
        /* what the block takes of its chunk: its bytes, or -- packed -- whole groups of 64 positions */
        const size_t take = j->packed ? (out + 63) & ~(size_t)63 : out;
        uint8_t *dst = NULL;
        int c = -1, s = -1;
        if (!rc && out) {
            if (take > j->cap) rc = QK_ERR_ARG;                     /* a line longer than a chunk */
            if (!rc && j->cur_ctx >= 0 && j->chunk[j->cur_ctx][j->cur_slot].fill + take > j->cap) close_chunk(j);
            if (!rc && j->cur_ctx < 0) rc = open_chunk(j);
            if (!rc) {
                c = j->cur_ctx;
                s = j->cur_slot;
                mt_chunk *ch = &j->chunk[c][s];
                dst = j->sink->buffer(j->sink->user, (uint32_t)c, (uint32_t)s) + (j->packed ? ch->fill / 64 * 24 : ch->fill);
                ch->fill += take;
                ch->lines += out_lines[s_in];
                atomic_fetch_add(&ch->pending, 1);
            }
        }
        if (rc) { int z = 0; atomic_compare_exchange_strong(&j->err, &z, rc); }
        else {
            j->state = exit_state[s_in];
            j->lines += out_lines[s_in];
            j->bases += out - out_lines[s_in];
            j->sink_bytes += j->packed ? take / 64 * 24 : take;
            j->unterminated += (uint64_t)(unterminated && last_kept[s_in]); /* counted when it is a read, as qk_framer_next does */
        }
        if (i + 1 == j->n_blocks) close_chunk(j);                   /* (its own copy below still holds it open through `pending`) */
        atomic_store_explicit(&j->resolved, i + 1, memory_order_release);
        /* 4. copy the sequence lines */
        if (dst) {
            uint64_t longl;
            if (!j->packed) longl = j->avx512 && j->nt && out <= j->stage_cap ? copy_avx512(dst, d, n, &S, s_in, w->stage) : copy_plain(dst, d, n, &S, s_in);
            else if (j->avx512 && take <= j->stage_cap) longl = copy_packed_avx512(dst, d, n, &S, s_in, take, w->stage, w->pstage, j->nt);
            else {                                                  /* no AVX-512, or a block with a line of megabytes */
                uint8_t *tmp = malloc(take);
                if (tmp) {
                    longl = copy_plain(tmp, d, n, &S, s_in);
                    memset(tmp + out, '\n', take - out);
                    pack_plain(dst, tmp, take);
                    free(tmp);
                } else { int z = 0; longl = 0; atomic_compare_exchange_strong(&j->err, &z, QK_ERR_NOMEM); }
            }
            if (longl) { pthread_mutex_lock(&j->submit_mu[0]); j->long_lines += longl; pthread_mutex_unlock(&j->submit_mu[0]); }
            atomic_fetch_sub(&j->chunk[c][s].pending, 1);
            try_submit(j, c, s);
        }
    }
}

static __thread int qk_mt_input_is_mapping;   /* set by qk_count_file_mt around its call */

static uint32_t default_threads(uint32_t n_ctx)
{
    const char *e = getenv("QK_FRAMER_THREADS");
    if (e && atoi(e) > 0) return (uint32_t)atoi(e);
    long cpus = sysconf(_SC_NPROCESSORS_ONLN);
    (void)n_ctx;
    if (cpus < 1) cpus = 1;
    return (uint32_t)(cpus > QK_MT_MAX_THREADS ? QK_MT_MAX_THREADS : cpus);
}

int qk_frame_mem_mt(const qk_chunk_sink *sink, const uint8_t *data, size_t n, int seekable, uint32_t threads, qk_framer_stats *st)
{
    if (!sink || !sink->buffer || !sink->ready || !sink->wait || !sink->submit || sink->n_ctx < 1 || sink->n_ctx > QK_MT_MAX_CTX ||
        sink->n_slots < 1 || sink->n_slots > QK_HOST_MAX_SLOTS || sink->cap < 200000 || (!data && n))
        return QK_ERR_ARG;
    mt_job *j = calloc(1, sizeof *j);
    if (!j) return QK_ERR_NOMEM;
    int rc = QK_OK;
    j->sink = sink;
    j->n_ctx = sink->n_ctx;
    j->n_slots = sink->n_slots;
    j->packed = sink->packed != 0;
    j->cap = j->packed ? sink->cap & ~(size_t)63 : sink->cap;        /* packed: positions, in whole groups of 64 */
    for (uint32_t c = 0; c < j->n_ctx; ++c) pthread_mutex_init(&j->submit_mu[c], NULL);
    j->data = data;
    j->n = n;
    j->fastq = n && data[0] == '@';                                  /* Q.c:395 */
    j->state = n && (j->fastq || !seekable) ? 3u : 0u;               /* first line consumed: FASTQ always, FASTA on a pipe (Q.c:396) */
    /* blocks: large enough to amortise the hand-over, small enough that a block's output (<= block + one
     * line) fits a chunk several times over */
    j->block = (size_t)512 << 10;        /* block + its kept lines (staged) + the line table stay in a 2 MB L2 */
    while (j->block > ((size_t)16 << 10) && j->block + 100000 > j->cap / 2) j->block >>= 1;
    j->n_blocks = (n + j->block - 1) / j->block;
    j->cur_ctx = j->cur_slot = -1;
    j->avx512 = __builtin_cpu_supports("avx512bw") && !getenv("QK_NO_AVX512");
    { const char *e = getenv("QK_FRAMER_NT"); j->nt = e ? atoi(e) != 0 : 1; }
    j->populate = qk_mt_input_is_mapping;
    build_table(j->tab, j->fastq);
    j->stage_cap = j->block + ((size_t)128 << 10);
    if (!threads) threads = default_threads(j->n_ctx);
    if (threads > QK_MT_MAX_THREADS) threads = QK_MT_MAX_THREADS;
    if (threads > j->n_blocks) threads = j->n_blocks ? (uint32_t)j->n_blocks : 1;
    /* a block can hold one line start per byte, and its last line's end: block + 2 entries */
    const size_t starts_cap = j->block + 3;
    mt_worker wk[QK_MT_MAX_THREADS];
    pthread_t th[QK_MT_MAX_THREADS];
    uint32_t started = 0, allocated = 0;
    for (uint32_t t = 0; t < threads && !rc; ++t) {
        wk[t].j = j;
        wk[t].starts_cap = starts_cap;
        wk[t].starts = malloc(starts_cap * sizeof(uint64_t));
        wk[t].keep = malloc(starts_cap);
        wk[t].stage = j->avx512 ? aligned_alloc(64, (j->stage_cap + 128 + 63) / 64 * 64) : NULL;
        wk[t].pstage = j->avx512 && j->packed ? aligned_alloc(64, (j->stage_cap / 64 * 24 + 128 + 63) / 64 * 64) : NULL;
        if (!wk[t].starts || !wk[t].keep || (j->avx512 && !wk[t].stage) || (j->avx512 && j->packed && !wk[t].pstage)) {
            free(wk[t].starts); free(wk[t].keep); free(wk[t].stage); free(wk[t].pstage);
            rc = QK_ERR_NOMEM;
            break;
        }
        ++allocated;
    }
    if (rc) { int z = 0; atomic_compare_exchange_strong(&j->err, &z, rc); }
    for (uint32_t t = 0; !rc && t + 1 < threads; ++t) {              /* the calling thread is the last worker */
        if (pthread_create(&th[t], NULL, worker, &wk[t]) != 0) break; /* fewer workers, same result */
        ++started;
    }
    if (!rc && j->n_blocks) worker(&wk[threads - 1]);
    for (uint32_t t = 0; t < started; ++t) pthread_join(th[t], NULL);
    for (uint32_t t = 0; t < allocated; ++t) { free(wk[t].starts); free(wk[t].keep); free(wk[t].stage); free(wk[t].pstage); }
    if (!rc) rc = atomic_load(&j->err);
    for (uint32_t c = 0; c < j->n_ctx; ++c) pthread_mutex_destroy(&j->submit_mu[c]);
    if (st) {
        memset(st, 0, sizeof *st);
        st->lines = j->lines;
        st->bases = j->bases;
        st->raw_bytes = n;
        st->long_lines = j->long_lines;
        st->unterminated = j->unterminated;
        st->fastq = j->fastq;
        st->sink_bytes = j->sink_bytes;
    }
    free(j);
    return rc;
}

/* ---- the sink that counts: chunk slots of one context per GPU ------------------------------------ */
typedef struct { qk_ctx *ctx[QK_MT_MAX_CTX]; int packed; } ctx_sink;
static uint8_t *cs_buffer(void *u, uint32_t c, uint32_t s) { return qk_slot_host_buffer(((ctx_sink *)u)->ctx[c], s); }
static int cs_ready(void *u, uint32_t c, uint32_t s) { return qk_slot_ready(((ctx_sink *)u)->ctx[c], s); }
static int cs_wait(void *u, uint32_t c, uint32_t s) { return qk_wait_slot(((ctx_sink *)u)->ctx[c], s); }
static int cs_submit(void *u, uint32_t c, uint32_t s, uint64_t seq, size_t n_bytes, uint32_t n_lines)
{
    (void)seq;                              /* counting is an integer sum: chunk order does not matter */
    qk_ctx *ctx = ((ctx_sink *)u)->ctx[c];
    if (((ctx_sink *)u)->packed) return qk_submit_packed(ctx, s, qk_slot_host_buffer(ctx, s), n_bytes, n_lines);
    return qk_submit(ctx, s, qk_slot_host_buffer(ctx, s), n_bytes, NULL, n_lines); /* "sem_post", Q.c:431-432 */
}

int qk_count_mem_mt(qk_ctx *const *ctxs, uint32_t n_ctx, const uint8_t *data, size_t n, int seekable, uint32_t threads,
                    qk_framer_stats *st)
{
    if (!ctxs || n_ctx < 1 || n_ctx > QK_MT_MAX_CTX || (!data && n)) return QK_ERR_ARG;
    ctx_sink cs;
    qk_chunk_sink sink = {&cs, n_ctx, 0, 0, cs_buffer, cs_ready, cs_wait, cs_submit, 0};
    /* Packed chunks (2-bit codes + reset flags, 0.375 bytes per position: see qk_chunk_sink) when every context counts
     * with the dictionary-order kernel, the only one that reads them (3 <= k <= 31).  QK_PACKED=0 ships the text. */
    const char *pk = getenv("QK_PACKED");
    cs.packed = pk ? atoi(pk) != 0 : QK_PACKED_DEFAULT;
    for (uint32_t c = 0; c < n_ctx; ++c) {
        uint32_t ns = 0;
        size_t cap = 0;
        int rc = qk_ctx_info(ctxs[c], &ns, &cap);
        if (rc) return rc;
        if (c == 0) { sink.n_slots = ns; sink.cap = cap; }
        else if (ns != sink.n_slots || cap != sink.cap) return QK_ERR_ARG;   /* same slot geometry everywhere */
        cs.ctx[c] = ctxs[c];
        qk_table_desc desc;
        if (qk_dict_describe(ctxs[c], &desc) != QK_OK || !desc.has_ext) cs.packed = 0;
    }
    sink.packed = cs.packed;
    int rc = qk_frame_mem_mt(&sink, data, n, seekable, threads, st);
    for (uint32_t c = 0; c < n_ctx; ++c) {
        int r2 = qk_sync(ctxs[c]);                                    /* drain + join, Q.c:458-479 */
        if (!rc) rc = r2;
    }
    return rc;
}

/* ---- measurement: the framer alone, chunks discarded ---------------------------------------------- */
typedef struct { uint8_t *buf[QK_HOST_MAX_SLOTS]; _Atomic uint64_t bytes; } null_sink;
static uint8_t *ns_buffer(void *u, uint32_t c, uint32_t s) { (void)c; return ((null_sink *)u)->buf[s]; }
static int ns_ready(void *u, uint32_t c, uint32_t s) { (void)u; (void)c; (void)s; return 1; }
static int ns_wait(void *u, uint32_t c, uint32_t s) { (void)u; (void)c; (void)s; return 0; }
static int ns_submit(void *u, uint32_t c, uint32_t s, uint64_t seq, size_t n_bytes, uint32_t n_lines)
{
    (void)c; (void)s; (void)seq; (void)n_lines;
    atomic_fetch_add(&((null_sink *)u)->bytes, n_bytes);
    return QK_OK;
}

int qk_bench_framer(const uint8_t *data, size_t n, uint32_t threads, int repeats, double *raw_gbs, double *framed_gbs)
{
    if (!data || !n || repeats < 1) return QK_ERR_ARG;
    null_sink ns;
    memset(&ns, 0, sizeof ns);
    const uint32_t n_slots = 8;
    const size_t cap = (size_t)32 << 20;
    for (uint32_t s = 0; s < n_slots; ++s) {
        ns.buf[s] = malloc(cap);
        if (!ns.buf[s]) { for (uint32_t t = 0; t < s; ++t) free(ns.buf[t]); return QK_ERR_NOMEM; }
        memset(ns.buf[s], 0, cap);
    }
    const char *pk = getenv("QK_PACKED");
    qk_chunk_sink sink = {&ns, 1, n_slots, cap, ns_buffer, ns_ready, ns_wait, ns_submit, pk ? atoi(pk) != 0 : QK_PACKED_DEFAULT};
    int rc = qk_frame_mem_mt(&sink, data, n, 1, threads, NULL);    /* warm-up */
    atomic_store(&ns.bytes, 0);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int r = 0; !rc && r < repeats; ++r) rc = qk_frame_mem_mt(&sink, data, n, 1, threads, NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const double dt = (t1.tv_sec - t0.tv_sec) + (t1.tv_nsec - t0.tv_nsec) * 1e-9;
    if (raw_gbs) *raw_gbs = (double)n * repeats / dt / 1e9;
    if (framed_gbs) *framed_gbs = (double)atomic_load(&ns.bytes) / dt / 1e9;
    for (uint32_t s = 0; s < n_slots; ++s) free(ns.buf[s]);
    return rc;
}

/* Host memory bandwidth with `threads` threads, for scale: GB/s of a pure read (sum) and of memcpy (bytes copied). */
typedef struct { uint8_t *a, *b; size_t n; int mode; uint64_t sink; } hm_job;
static void *hm_worker(void *arg)
{
    hm_job *h = arg;
    if (h->mode == 0) {
        const uint64_t *p = (const uint64_t *)h->a;
        uint64_t acc = 0;
        for (size_t i = 0; i < h->n / 8; ++i) acc += p[i];
        h->sink = acc;
    } else memcpy(h->b, h->a, h->n);
    return NULL;
}
int qk_bench_host_memory(size_t bytes_per_thread, uint32_t threads, double *read_gbs, double *copy_gbs)
{
    if (threads < 1 || threads > QK_MT_MAX_THREADS || bytes_per_thread < 4096) return QK_ERR_ARG;
    hm_job job[QK_MT_MAX_THREADS];
    pthread_t th[QK_MT_MAX_THREADS];
    memset(job, 0, sizeof job);
    int rc = QK_OK;
    for (uint32_t t = 0; t < threads; ++t) {
        job[t].n = bytes_per_thread / 8 * 8;
        job[t].a = malloc(job[t].n);
        job[t].b = malloc(job[t].n);
        if (!job[t].a || !job[t].b) { rc = QK_ERR_NOMEM; break; }
        memset(job[t].a, 1, job[t].n);
        memset(job[t].b, 2, job[t].n);
    }
    for (int mode = 0; !rc && mode < 2; ++mode) {
        double best = 0;
        for (int rep = 0; rep < 3; ++rep) {
            struct timespec t0, t1;
            clock_gettime(CLOCK_MONOTONIC, &t0);
            for (uint32_t t = 0; t < threads; ++t) { job[t].mode = mode; pthread_create(&th[t], NULL, hm_worker, &job[t]); }
            for (uint32_t t = 0; t < threads; ++t) pthread_join(th[t], NULL);
            clock_gettime(CLOCK_MONOTONIC, &t1);
            const double dt = (t1.tv_sec - t0.tv_sec) + (t1.tv_nsec - t0.tv_nsec) * 1e-9;
            const double g = (double)job[0].n * threads / dt / 1e9;
            if (g > best) best = g;
        }
        if (mode == 0 && read_gbs) *read_gbs = best;
        if (mode == 1 && copy_gbs) *copy_gbs = best;
    }
    for (uint32_t t = 0; t < threads; ++t) { free(job[t].a); free(job[t].b); }
    return rc;
}

/* A regular file: mapped, not read -- the page cache is the input buffer, the workers touch it
 * once and copy only the sequence lines.  Pipes and gzip go through the sequential stream path. */
int qk_count_file_mt(qk_ctx *const *ctxs, uint32_t n_ctx, const char *reads_path, uint32_t threads, qk_framer_stats *st)
{
    if (!ctxs || n_ctx < 1 || !reads_path) return QK_ERR_ARG;
    int fd = open(reads_path, O_RDONLY);
    if (fd < 0) return QK_ERR_IO;
    struct stat sb;
    uint8_t magic[2] = {0, 0};
    const int regular = fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode) && sb.st_size > 0;
    const int gz = regular && pread(fd, magic, 2, 0) == 2 && magic[0] == 0x1f && magic[1] == 0x8b;
    if (!regular || gz) {
        int seekable = lseek(fd, 0, SEEK_CUR) != (off_t)-1;
        int rc = qk_count_raw_fd(ctxs[0], fd, seekable, st);         /* one GPU: the stream is sequential */
        close(fd);
        return rc;
    }
    void *map = mmap(NULL, (size_t)sb.st_size, PROT_READ, MAP_SHARED, fd, 0);
    close(fd);
    if (map == MAP_FAILED) return QK_ERR_IO;
    madvise(map, (size_t)sb.st_size, MADV_SEQUENTIAL);
    qk_mt_input_is_mapping = getenv("QK_POPULATE") != NULL;   /* measured slower than plain faults on tmpfs (23.7 vs 37 GB/s): off */
    int rc = qk_count_mem_mt(ctxs, n_ctx, map, (size_t)sb.st_size, 1, threads, st);
    qk_mt_input_is_mapping = 0;
    munmap(map, (size_t)sb.st_size);
    return rc;
}
