/*
 * qk_oracle.c -- CPU restatement of QuicK-mer2's `count` path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker the CUDA path is compared
 * against.  Nothing under quick-mer2_b200/ may include, link, or execute it; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it.
 *
 * Parity pin: the reference ships no golden vectors for `count`, so
 * this restatement is pinned against outputs of the reference itself, compiled
 * unmodified from /root/reference/QuicKmer.c into oracle/_ref/quicKmer2 by
 * oracle/Makefile.  tests/golden/ holds .bin/.txt fixtures produced by that binary
 * (tests/golden/make_golden.py); tests/test_oracle.py checks this file against them
 * and, when oracle/_ref/quicKmer2 is present, against live runs of it.
 *
 * Every function cites the lines of QuicKmer.c ("Q.c") it restates.  It is a
 * restatement, not a copy: serial, single-threaded (the reference's result depends on
 * -t in one way only, the zero padding of its last worker batch: qko_fifo_padding), with
 * explicit handling of the two inputs on which
 * the reference has undefined behaviour (flagged in qko_stats.undefined_lines).
 */
#define _FILE_OFFSET_BITS 64
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define QKO_LINE_CAP 100000 /* Q.c:388 `char line[100000]`; fgets reads <= 99,999 bytes */
#define QKO_GC_BINS 401     /* Q.c:495-497 */

typedef struct {
    uint8_t k;        /* Q.c:346  byte 4 of the .qm header */
    uint64_t n_slots; /* Q.c:349  Hash_size, bytes 8..15   */
    uint64_t first;   /* Q.c:351  first_idx, bytes 16..23  */
    uint64_t *keys;   /* Q.c:359  Hash_size x u64          */
    uint32_t *next;   /* Q.c:483  Hash_size x u32 chain    */
} qko_dict;

typedef struct {
    uint64_t total_kmers;     /* Q.c:445 process_kmers                      */
    uint64_t hits;            /* emitted keys found in the dictionary        */
    uint64_t lines;           /* sequence lines processed                    */
    uint64_t bases;           /* bytes of sequence lines excluding '\n'      */
    uint64_t undefined_lines; /* lines the reference would run off (T8, T9)  */
    int fastq;                /* Q.c:395                                     */
} qko_stats;

/* ---- a-3: Q.c:66-76 --------------------------------------------------------------- */
uint64_t qko_djb(uint64_t key)
{
    uint64_t h = 5381;
    for (int b = 0; b < 8; ++b) {
        h = h * 33u + (key & 0xFFu);
        key >>= 8;
    }
    return h;
}

/* ---- a-4: Q.c:90-99.  Returns 1 and the slot when `key` sits on its probe path. ----
 * Home slot = djb & (H-1); walk +1 from the lower half, -1 from the upper half, until an
 * empty slot or the key.  Note that key 0 "matches" the first empty slot (Q.c:98). */
int qko_find(const qko_dict *d, uint64_t key, uint64_t *slot_out)
{
    uint64_t s = qko_djb(key) & (d->n_slots - 1);
    int64_t step = (s & (d->n_slots >> 1)) ? -1 : 1;
    while (d->keys[s] != 0 && d->keys[s] != key)
        s = (uint64_t)((int64_t)s + step);
    *slot_out = s;
    return d->keys[s] == key;
}

/* ---- a-7: Q.c:345-359, 483 --------------------------------------------------------- */
int qko_dict_load(const char *qm_path, qko_dict *d)
{
    memset(d, 0, sizeof *d);
    FILE *f = fopen(qm_path, "rb");
    if (!f) return 1;
    uint8_t hdr[24];
    if (fread(hdr, 1, 24, f) != 24) { fclose(f); return 2; }
    d->k = hdr[4];
    memcpy(&d->n_slots, hdr + 8, 8);
    memcpy(&d->first, hdr + 16, 8);
    if (d->n_slots == 0 || (d->n_slots & (d->n_slots - 1))) { fclose(f); return 3; }
    d->keys = malloc(d->n_slots * sizeof(uint64_t));
    d->next = malloc(d->n_slots * sizeof(uint32_t));
    if (!d->keys || !d->next) { fclose(f); return 4; }
    if (fread(d->keys, 8, d->n_slots, f) != d->n_slots) { fclose(f); return 5; }
    if (fread(d->next, 4, d->n_slots, f) != d->n_slots) { fclose(f); return 6; }
    fclose(f);
    return 0;
}

void qko_dict_free(qko_dict *d)
{
    free(d->keys);
    free(d->next);
    memset(d, 0, sizeof *d);
}

/* Chain length = number of .bin entries (Q.c:498-516: do { } while (c != first)). */
uint64_t qko_chain_length(const qko_dict *d)
{
    uint64_t n = 0;
    uint32_t c = (uint32_t)d->first;
    do { ++n; c = d->next[c]; } while (c != (uint32_t)d->first);
    return n;
}

/* ---- a-2: Q.c:399-420, one sequence line ------------------------------------------
 * `line` holds `len` bytes, none of which is '\n' (the terminator is implied at len).
 * If keys_out != NULL every emitted canonical key is appended (up to cap) -- used by the
 * codec unit tests; if depth != NULL hits are tallied by hash slot as Q.c:443 does. */
uint64_t qko_count_line(const qko_dict *d, uint8_t k, const uint8_t *line, size_t len,
                        uint16_t *depth, uint64_t *keys_out, size_t cap, uint64_t *hits)
{
    /* Q.c:419: ((uint64_t)1 << (k<<1)) - 1; x86-64 masks the shift count to 6 bits, so
     * k=32 gives (1<<0)-1 = 0. */
    const uint64_t mask = ((uint64_t)1 << ((2u * k) & 63u)) - 1;
    uint64_t fwd = 0, rc = 0, emitted = 0;
    uint16_t run = 0; /* Q.c:402 uint16_t cur_chars: wraps at 65,536 (T7) */
    for (size_t i = 0; i < len; ++i) {
        uint8_t c = line[i];
        if (c == 'N') { /* Q.c:404-408: only upper-case N resets (T5) */
            fwd = 0; rc = 0; run = 0;
            continue;
        }
        ++run;
        uint64_t code = (c >> 1) & 3u;                 /* Q.c:411 */
        fwd = (fwd << 2) | code;                       /* Q.c:412-413 */
        rc |= (uint64_t)((code - 2u) & 3u) << 60;      /* Q.c:414-415 */
        rc >>= 2;                                      /* Q.c:416: 60-bit register */
        if (run >= k) {                                /* Q.c:418 */
            uint64_t key = fwd & mask;                 /* Q.c:419 */
            if (key > rc) key = rc;                    /* Q.c:420 */
            if (keys_out && emitted < cap) keys_out[emitted] = key;
            ++emitted;
            if (depth) {
                uint64_t slot;
                if (qko_find(d, key, &slot)) {         /* Q.c:442-443 */
                    depth[slot]++;
                    if (hits && key != 0) ++*hits;
                }
            }
        }
    }
    return emitted;
}

/* Keys of a packed chunk: sequence lines separated by '\n' (the layout the device
 * consumes).  Returns the number of emitted keys; fills keys_out up to cap. */
uint64_t qko_chunk_keys(uint8_t k, const uint8_t *bytes, size_t n, uint64_t *keys_out, size_t cap)
{
    uint64_t total = 0;
    size_t start = 0;
    for (size_t i = 0; i <= n; ++i) {
        if (i == n || bytes[i] == '\n') {
            if (i > start || i < n) {
                size_t room = total < cap ? cap - (size_t)total : 0;
                total += qko_count_line(NULL, k, bytes + start, i - start, NULL,
                                        keys_out ? keys_out + (total < cap ? total : cap) : NULL,
                                        room, NULL);
            }
            start = i + 1;
        }
    }
    return total;
}

/* ---- a-1: Q.c:393-398, 451-455.  fgets(line, 100000, f) semantics ------------------
 * Reads up to 99,999 bytes, stopping after a '\n'.  Returns the number of bytes stored
 * (0 at EOF).  *has_nl says whether the stored bytes end in '\n'.  NUL bytes are kept:
 * the reference's scan loop (Q.c:403) only stops at '\n'. */
static size_t qko_getline(FILE *f, uint8_t *line, int *has_nl)
{
    size_t n = 0;
    int c;
    *has_nl = 0;
    while (n < QKO_LINE_CAP - 1 && (c = getc_unlocked(f)) != EOF) {
        line[n++] = (uint8_t)c;
        if (c == '\n') { *has_nl = 1; break; }
    }
    return n;
}

/* Stream a FASTA/FASTQ file; tally depth by hash slot.  Q.c:393-456. */
int qko_count_stream(const qko_dict *d, FILE *f, uint16_t *depth, qko_stats *st)
{
    static uint8_t line[QKO_LINE_CAP];
    int nl;
    memset(st, 0, sizeof *st);
    size_t n = qko_getline(f, line, &nl);       /* Q.c:393 */
    if (n && line[0] == '@') st->fastq = 1;     /* Q.c:395 */
    else if (fseeko(f, 0, SEEK_SET) != 0) { /* Q.c:396: on a pipe the first line is lost */ }
    while ((n = qko_getline(f, line, &nl)) > 0) {   /* Q.c:397 */
        if (line[0] == '>') continue;               /* Q.c:398 */
        size_t len = nl ? n - 1 : n;
        if (!nl) st->undefined_lines++;  /* T8/T9: reference scans past the buffer here */
        st->total_kmers += qko_count_line(d, d->k, line, len, depth, NULL, 0, &st->hits);
        st->lines++;
        st->bases += len;
        if (st->fastq) {                            /* Q.c:451-455 */
            qko_getline(f, line, &nl);
            qko_getline(f, line, &nl);
            qko_getline(f, line, &nl);
        }
    }
    return 0;
}

/* The sequence lines the loop of Q.c:397-456 would process, concatenated, each followed by
 * '\n' -- the framing alone, for checking the host framer.  Returns bytes needed. */
size_t qko_frame_stream(FILE *f, uint8_t *out, size_t cap, qko_stats *st)
{
    static uint8_t line[QKO_LINE_CAP];
    int nl;
    size_t total = 0;
    memset(st, 0, sizeof *st);
    size_t n = qko_getline(f, line, &nl);
    if (n && line[0] == '@') st->fastq = 1;
    else if (fseeko(f, 0, SEEK_SET) != 0) { /* pipe: first line lost (Q.c:396) */ }
    while ((n = qko_getline(f, line, &nl)) > 0) {
        if (line[0] == '>') continue;
        size_t len = nl ? n - 1 : n;
        if (!nl) st->undefined_lines++;
        if (total + len + 1 <= cap) {
            memcpy(out + total, line, len);
            out[total + len] = '\n';
        }
        total += len + 1;
        st->lines++;
        st->bases += len;
        if (st->fastq) {
            qko_getline(f, line, &nl);
            qko_getline(f, line, &nl);
            qko_getline(f, line, &nl);
        }
    }
    return total;
}

size_t qko_frame_file(const char *path, uint8_t *out, size_t cap, qko_stats *st)
{
    FILE *f = fopen(path, "rb");
    if (!f) return (size_t)-1;
    size_t n = qko_frame_stream(f, out, cap, st);
    fclose(f);
    return n;
}

/* Count a file against a loaded dictionary; depth is indexed by hash slot (Q.c:443). */
int qko_count_file(const qko_dict *d, const char *path, uint16_t *depth, qko_stats *st)
{
    FILE *f = fopen(path, "rb");
    if (!f) return 1;
    qko_count_stream(d, f, depth, st);
    fclose(f);
    return 0;
}

/* ---- a-8: Q.c:490-518.  Depth by hash slot -> depth in chain (reference) order ---- */
uint64_t qko_chain_gather(const qko_dict *d, const uint16_t *depth, uint16_t *out, uint64_t cap)
{
    uint64_t n = 0;
    uint32_t c = (uint32_t)d->first;            /* Q.c:494: first_idx truncated to u32 */
    do {
        if (n < cap) out[n] = depth[c];
        ++n;
        c = d->next[c];
    } while (c != (uint32_t)d->first);
    return n;
}

/* ---- a-9: Q.c:495-509, 522-542.  GC control curve from ordered depths + .qgc ------ */
typedef struct {
    double curve[QKO_GC_BINS];   /* sum of depth, later mean  */
    double sq[QKO_GC_BINS];      /* sum of depth^2, later var */
    uint32_t count[QKO_GC_BINS];
    double mean_depth;           /* Q.c:539-540 */
} qko_gc;

void qko_gc_accumulate(const uint16_t *ordered, const uint16_t *qgc, uint64_t n, qko_gc *g)
{
    memset(g, 0, sizeof *g);
    for (uint64_t i = 0; i < n; ++i) {
        if (qgc[i] & 0x8000u) {                               /* Q.c:504 */
            unsigned bin = qgc[i] & 0x1FFu;
            if (bin >= QKO_GC_BINS) continue; /* reference would write out of bounds */
            int dd = (int)((uint32_t)ordered[i] * (uint32_t)ordered[i]); /* Q.c:507 int product */
            g->curve[bin] += ordered[i];                      /* Q.c:505 */
            g->count[bin] += 1;                               /* Q.c:506 */
            g->sq[bin] += dd;
        }
    }
}

/* Finalise and print the 401 lines of <out>.txt exactly as Q.c:529-538 formats them. */
int qko_gc_write(qko_gc *g, const char *txt_path)
{
    FILE *f = fopen(txt_path, "w");
    if (!f) return 1;
    double total_depth = 0;
    uint64_t total_count = 0;
    for (int i = 0; i < QKO_GC_BINS; ++i) {
        total_count += g->count[i];
        total_depth += g->curve[i];
        if (g->count[i]) {
            g->curve[i] /= g->count[i];
            volatile double m2 = g->curve[i] * g->curve[i]; /* no FMA contraction */
            g->sq[i] = g->sq[i] / g->count[i] - m2;
        }
        fprintf(f, "%.2f\t%f\t%i\t%f\n", i / 4.0, g->curve[i], g->count[i], g->sq[i]);
    }
    g->mean_depth = total_depth / total_count;
    fclose(f);
    return 0;
}

/* ---- Q.c:458-466, 284-291: `count -t N`, N > 0.  K-mers travel to the workers in batches of FIFO_size = 4,096
 * (Q.c:12); the last batch is filled up with zeros -- a whole batch of them when the total is a multiple of
 * 4,096 -- and the workers look every entry up.  Find_hash(0) "finds" the first empty slot on key 0's path, so
 * that slot's depth grows by the padding.  Returns the padding. */
uint64_t qko_fifo_padding(const qko_dict *d, uint64_t total_kmers, uint16_t *depth)
{
    uint64_t pad = 4096 - total_kmers % 4096, slot;
    qko_find(d, 0, &slot);
    depth[slot] = (uint16_t)(depth[slot] + pad);
    return pad;
}

/* ---- whole command: Q.c:304-545 ---------------------------------------------------- */
int qko_count_t(const char *ref_prefix, const char *reads_path, const char *out_prefix, unsigned threads, qko_stats *st);
int qko_count(const char *ref_prefix, const char *reads_path, const char *out_prefix, qko_stats *st)
{
    return qko_count_t(ref_prefix, reads_path, out_prefix, 0, st);
}

int qko_count_t(const char *ref_prefix, const char *reads_path, const char *out_prefix, unsigned threads, qko_stats *st)
{
    char path[4096];
    qko_dict d;
    snprintf(path, sizeof path, "%s.qm", ref_prefix);
    int rc = qko_dict_load(path, &d);
    if (rc) return 10 + rc;
    FILE *reads = fopen(reads_path, "rb");
    if (!reads) { qko_dict_free(&d); return 2; }
    uint16_t *depth = calloc(d.n_slots, sizeof(uint16_t));
    if (!depth) return 3;
    qko_count_stream(&d, reads, depth, st);
    fclose(reads);
    if (threads & 0xFF) qko_fifo_padding(&d, st->total_kmers, depth);   /* uint8_t thread_count, Q.c:306 */

    uint64_t n = qko_chain_length(&d);
    uint16_t *ordered = malloc(n * sizeof(uint16_t));
    if (!ordered) return 3;
    qko_chain_gather(&d, depth, ordered, n);
    snprintf(path, sizeof path, "%s.bin", out_prefix);
    FILE *bin = fopen(path, "wb");
    if (!bin) return 4;
    fwrite(ordered, 2, n, bin);
    fclose(bin);

    snprintf(path, sizeof path, "%s.qgc", ref_prefix);
    FILE *qgc = fopen(path, "rb");
    if (qgc) {                                              /* Q.c:486-488, 522 */
        uint16_t *g = calloc(n, sizeof(uint16_t));
        size_t got = fread(g, 2, n, qgc);
        (void)got; /* short .qgc: the reference keeps stale buffer contents; we use zeros */
        fclose(qgc);
        qko_gc gc;
        qko_gc_accumulate(ordered, g, n, &gc);
        snprintf(path, sizeof path, "%s.txt", out_prefix);
        qko_gc_write(&gc, path);
        free(g);
    }
    free(ordered);
    free(depth);
    qko_dict_free(&d);
    return 0;
}

/* ---- `search` pass 1: Q.c:824-923 (hash_from_fasta) with the table growth of Q.c:738-822 ------------------------
 * SURVEY 8(f) rank 4, the oracle half: which canonical k-mers the reference genome holds and how often.  Differences
 * from `count`'s codec (qko_count_line) that a device version has to reproduce:
 *   - the FASTA is read 199 bytes at a time (fgets(buf, 200)); the register is NOT reset at a line end, only at a
 *     line that starts with '>' and at 'N' -- a sequence continues over its lines (and a header line longer than
 *     199 bytes continues as sequence); reading stops at a piece that starts with a NUL byte;
 *   - the run counter saturates at k (no 16-bit wrap); the key 0 (poly-A / poly-T) is never stored;
 *   - occurrences saturate at 255;
 *   - when more than 0.8 of the slots are taken -- checked after every piece -- the table doubles and is re-probed IN
 *     PLACE (upper half of the old range downwards, then the lower half upwards).  Which slot a key ends up in
 *     depends on that order, so it is restated as it is; the totals the reference prints came out the same for
 *     every starting size tried (tests/test_oracle.py), i.e. the sweep loses no key.
 * Returns 0 and fills *out (keys/occ are malloc'd, hash_size slots each). */
typedef struct {
    uint64_t hash_size;   /* final Hash_size                                     */
    uint64_t distinct;    /* `count`: keys entered (Q.c:877), re-entries included */
    uint64_t unique;      /* slots with occurrence 1 (Q.c:916-921)               */
    uint64_t resizes;
    uint64_t *keys;
    uint8_t *occ;
} qko_pass1;

static uint64_t qko_p1_find(const uint64_t *keys, uint64_t n_slots, uint64_t key)   /* Q.c:90-99 on a bare array */
{
    uint64_t s = qko_djb(key) & (n_slots - 1);
    const int64_t step = (s & (n_slots >> 1)) ? -1 : 1;
    while (keys[s] != 0 && keys[s] != key) s = (uint64_t)((int64_t)s + step);
    return s;
}

static void qko_p1_rehome(uint64_t *keys, uint8_t *occ, uint64_t n_slots, uint64_t at)   /* one step of Q.c:760-790 */
{
    if (!keys[at]) return;
    const uint64_t to = qko_p1_find(keys, n_slots, keys[at]);
    if (to == at) return;
    keys[to] = keys[at]; keys[at] = 0;
    occ[to] = occ[at];   occ[at] = 0;
}

int qko_search_pass1(const char *fasta_path, uint8_t k, uint64_t hash_size, qko_pass1 *out)
{
    memset(out, 0, sizeof *out);
    FILE *f = fopen(fasta_path, "r");
    if (!f || k < 1 || k > 32 || hash_size < 2 || (hash_size & (hash_size - 1))) { if (f) fclose(f); return 1; }
    uint64_t *keys = calloc(hash_size, sizeof *keys);
    uint8_t *occ = calloc(hash_size, 1);
    if (!keys || !occ) { fclose(f); return 2; }
    const uint64_t mask = k < 32 ? (((uint64_t)1 << (2 * k)) - 1) : 0;       /* Q.c:860 on x86-64: 1 << 64 == 1 << 0 */
    char buf[200];
    uint8_t charge = 0;
    uint64_t fwd = 0, rc = 0, distinct = 0, resizes = 0;
    while (fgets(buf, 200, f) && buf[0]) {                                   /* Q.c:835 */
        if (buf[0] == '>') { charge = 0; fwd = rc = 0; continue; }           /* Q.c:838-845 */
        for (const char *p = buf; *p && *p != '\n'; ++p) {
            if (*p == 'N') { charge = 0; fwd = rc = 0; continue; }           /* Q.c:848-854 */
            const uint64_t code = ((uint8_t)*p >> 1) & 3;                    /* Q.c:855-861 */
            fwd = (fwd << 2) | code;
            rc = (rc | (((code - 2) & 3) << 60)) >> 2;
            uint64_t key = fwd & mask;
            if (key > rc) key = rc;
            if (charge < k) ++charge;
            if (!key || charge != k) continue;                               /* Q.c:864 */
            const uint64_t s = qko_p1_find(keys, hash_size, key);            /* Q.c:866-875 */
            if (!keys[s]) { keys[s] = key; ++distinct; }
            if (occ[s] < 255) ++occ[s];                                      /* Q.c:888 */
        }
        if ((double)distinct > 0.8 * (double)hash_size) {                    /* Q.c:891-895 */
            const uint64_t old = hash_size, grown = hash_size << 1;
            uint64_t *k2 = realloc(keys, grown * sizeof *keys);
            uint8_t *o2 = k2 ? realloc(occ, grown) : NULL;
            if (!k2 || !o2) { free(k2 ? k2 : keys); free(occ); fclose(f); return 2; }
            keys = k2; occ = o2;
            memset(keys + old, 0, (grown - old) * sizeof *keys);
            memset(occ + old, 0, grown - old);     /* (the reference leaves realloc's bytes; glibc hands back zero pages
                                                    *  for tables of this size -- slots of empty keys, never read before
                                                    *  they are written, Q.c:766) */
            hash_size = grown;
            for (uint64_t i = old - 1; i >= (old >> 1); --i) qko_p1_rehome(keys, occ, hash_size, i);   /* Q.c:758-772 */
            for (uint64_t i = 0; i < (old >> 1); ++i) qko_p1_rehome(keys, occ, hash_size, i);         /* Q.c:773-788 */
            ++resizes;
        }
    }
    fclose(f);
    uint64_t unique = 0;
    for (uint64_t i = 0; i < hash_size; ++i) unique += occ[i] == 1;          /* Q.c:916-921 */
    out->hash_size = hash_size;
    out->distinct = distinct;
    out->unique = unique;
    out->resizes = resizes;
    out->keys = keys;
    out->occ = occ;
    return 0;
}

void qko_pass1_free(qko_pass1 *p) { free(p->keys); free(p->occ); memset(p, 0, sizeof *p); }

/* ---- est: window depths (Q.c:555-685), given the correction curve ----------------------------------
 * The reference gets its 401-float correction curve from `popen("smooth_GC_mrsfast.py <sample>.txt")`
 * (Q.c:642-650: LOWESS in Python); everything else of main_estimate is restated here with the curve as
 * an input: mean depth from <sample>.txt (Q.c:626-638), then one pass over <ref>.qgc / <sample>.bin in
 * blocks of 1 MiB BYTES, accumulating correction[gc & 0x1FF] * depth (a float product added to a double)
 * into the current <ref>.bed window [left, right) and printing a window when the first k-mer index
 * >= right comes by (Q.c:660-682) -- including what that loop does after the last window: it leaves the
 * inner loop only, so every further block prints that window once more, its value divided again. */
#define QKO_EST_BLOCK (1024 * 1024)
int qko_est(const char *ref_prefix, const char *sample_prefix, const char *out_path, const float *correction /* 401 */)
{
    char path[4096];
    snprintf(path, sizeof path, "%s.qgc", ref_prefix);
    FILE *fg = fopen(path, "rb");
    snprintf(path, sizeof path, "%s.bed", ref_prefix);
    FILE *fw = fopen(path, "r");
    snprintf(path, sizeof path, "%s.bin", sample_prefix);
    FILE *fd = fopen(path, "rb");
    snprintf(path, sizeof path, "%s.txt", sample_prefix);
    FILE *ft = fopen(path, "r");
    FILE *fo = fopen(out_path, "w");
    if (!fg || !fw || !fd || !ft || !fo) return 1;
    char word[255];
    double total_depth = 0;
    uint64_t total_count = 0;
    float percent, depth_f;
    uint32_t n;
    while (fscanf(ft, "%f\t%f\t%i\t%254s\n", &percent, &depth_f, &n, word) == 4) { /* Q.c:633-636 */
        total_depth += depth_f * n;                                                    /* float * uint32 -> float */
        total_count += n;
    }
    total_depth /= total_count;
    fclose(ft);
    static uint16_t gc[QKO_EST_BLOCK], dep[QKO_EST_BLOCK];
    double cur = 0.0;
    uint64_t idx = 0;
    char chrom[64], wb[64], we[64];
    uint32_t left = 0, right = 0;
    int have = fscanf(fw, "%63s\t%63s\t%63s\t%u\t%u\n", chrom, wb, we, &left, &right) == 5;   /* Q.c:657 */
    (void)have;
    uint32_t got;
    while ((got = (uint32_t)fread(gc, 1, QKO_EST_BLOCK, fg)) != 0) {
        if (fread(dep, 1, got, fd) != got) { /* short .bin: the reference uses what is in the buffer */ }
        got >>= 1;
        for (uint32_t i = 0; i < got; ++i, ++idx) {
            if (idx >= right) {                                                            /* Q.c:667-675 */
                cur /= right - left;
                cur /= total_depth / 2;
                fprintf(fo, "%s\t%s\t%s\t%f\n", chrom, wb, we, cur);
                if (fscanf(fw, "%63s\t%63s\t%63s\t%u\t%u\n", chrom, wb, we, &left, &right) != 5) break;
                cur = 0.0;
            }
            if (idx < right && idx >= left) cur += correction[gc[i] & 0x1FF] * dep[i];    /* Q.c:677-679 */
        }
    }
    fclose(fo); fclose(fg); fclose(fw); fclose(fd);
    return 0;
}

#ifdef QKO_MAIN
int main(int argc, char **argv)
{
    if (argc == 6 && !strcmp(argv[1], "est")) {        /* qk_oracle est ref_prefix sample_prefix out.bed curve.f32 */
        float corr[512] = {0};
        FILE *f = fopen(argv[5], "rb");
        if (!f || fread(corr, 4, 401, f) != 401) { fprintf(stderr, "qk_oracle: cannot read the 401-float curve\n"); return 1; }
        fclose(f);
        return qko_est(argv[2], argv[3], argv[4], corr);
    }
    if (argc < 5 || strcmp(argv[1], "count")) {
        fprintf(stderr, "usage: qk_oracle count ref_prefix reads out_prefix | qk_oracle est ref_prefix sample_prefix out.bed curve.f32\n");
        return 1;
    }
    qko_stats st;
    int rc = qko_count(argv[argc - 3], argv[argc - 2], argv[argc - 1], &st);
    if (rc) { fprintf(stderr, "qk_oracle: error %d\n", rc); return 1; }
    printf("{\"total_kmers\": %llu, \"hits\": %llu, \"lines\": %llu, \"bases\": %llu, "
           "\"undefined_lines\": %llu, \"fastq\": %d}\n",
           (unsigned long long)st.total_kmers, (unsigned long long)st.hits,
           (unsigned long long)st.lines, (unsigned long long)st.bases,
           (unsigned long long)st.undefined_lines, st.fastq);
    return 0;
}
#endif
