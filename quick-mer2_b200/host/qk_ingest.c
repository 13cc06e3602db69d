/*
 * qk_ingest.c -- getting bytes to the device: the .qm dictionary (Q.c:345-359, 483) and the reads
 * (Q.c:393-456) through the slots' pinned buffers, with reader threads where the input is a
 * regular file; sharding one reads file over several GPUs.  The device frames the reads
 * (csrc/qk_frame.cu); the host only cuts the stream at line ends.  See include/qk_host.h.
 */
#define _GNU_SOURCE
#define _FILE_OFFSET_BITS 64
#include <errno.h>
#include <fcntl.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include "qk_host_internal.h"

/* One array of the .qm (keys: 8-byte elements at file offset 24; chain: 4-byte elements after
 * the keys) -> device, through the slots' pinned buffers: reader threads pread() pieces in
 * parallel, this thread enqueues the H2D copies in order. */
static int qm_upload_array(qk_ctx *ctx, int fd, uint64_t file_off, uint64_t n_elems, int kind, uint32_t threads);

int qk_qm_load(qk_ctx *ctx, const char *qm_path, qk_qm_header *hdr_out, uint64_t *n_kmers_out)
{
    qk_qm_header hdr;
    int rc = qk_qm_read_header(qm_path, &hdr);
    if (rc) return rc;
    if (hdr_out) *hdr_out = hdr;
    rc = qk_dict_begin(ctx, hdr.k, hdr.hash_size, hdr.first_idx);
    if (rc) return rc;
    int fd = open(qm_path, O_RDONLY);
    if (fd < 0) return QK_ERR_IO;
    struct stat sb;
    if (fstat(fd, &sb) != 0 || (uint64_t)sb.st_size < 24 + hdr.hash_size * 12) { close(fd); return QK_ERR_IO; } /* short file */
    const uint32_t threads = qk_reader_threads_default();
    const int verbose = getenv("QK_TIMING") != NULL;
    struct timespec t0, t1, t2;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    rc = qm_upload_array(ctx, fd, 24, hdr.hash_size, 0, threads);                           /* keys, Q.c:359 */
    if (!rc) rc = qm_upload_array(ctx, fd, 24 + hdr.hash_size * 8, hdr.hash_size, 1, threads); /* chain, Q.c:483 */
    close(fd);
    if (rc) return rc;
    if (verbose) { qk_sync(ctx); clock_gettime(CLOCK_MONOTONIC, &t1); }
    rc = qk_dict_build(ctx, n_kmers_out);
    if (verbose) {
        clock_gettime(CLOCK_MONOTONIC, &t2);
        fprintf(stderr, "[qk] .qm upload %.3f s (%u readers), table build %.3f s\n",
                (t1.tv_sec - t0.tv_sec) + (t1.tv_nsec - t0.tv_nsec) * 1e-9, threads,
                (t2.tv_sec - t1.tv_sec) + (t2.tv_nsec - t1.tv_nsec) * 1e-9);
    }
    return rc;
}


/* ---- raw streams: the device frames (qk_frame.cu); the host only cuts at line ends ------ */
static void raw_mode(uint8_t first_byte, int seekable, int *fastq, int *skip_first)
{
    *fastq = first_byte == '@';             /* Q.c:395 */
    *skip_first = *fastq || !seekable;      /* FASTQ: the first line is consumed; pipe: fseek fails (Q.c:396) */
}

static int raw_finish(qk_ctx *ctx, qk_framer_stats *st, uint64_t raw_bytes, uint64_t unterminated, int fastq)
{
    int rc = qk_sync(ctx);
    if (rc || !st) return rc;
    memset(st, 0, sizeof *st);
    uint64_t lines = 0, bases = 0;
    rc = qk_raw_stats(ctx, &lines, &bases, NULL);
    st->lines = lines;
    st->bases = bases;
    st->raw_bytes = raw_bytes;
    st->unterminated = unterminated;
    st->fastq = fastq;
    return rc;
}

int qk_count_raw_mem(qk_ctx *ctx, const uint8_t *data, size_t n, int seekable, qk_framer_stats *st)
{
    uint32_t n_slots = 0;
    size_t cap = 0;
    int rc = qk_ctx_info(ctx, &n_slots, &cap);
    if (rc) return rc;
    if (!data && n) return QK_ERR_ARG;
    int fastq = 0, skip_first = 0;
    if (n) raw_mode(data[0], seekable, &fastq, &skip_first);
    rc = qk_raw_begin(ctx, fastq, skip_first);
    if (rc) return rc;
    const int pinned = n && qk_host_is_pinned(data) && qk_host_is_pinned(data + n - 1);
    uint64_t unterminated = 0;
    uint32_t slot = 0;
    for (size_t pos = 0; pos < n; slot = (slot + 1) % n_slots) {
        const size_t end = n - pos > cap ? pos + cap : n;
        const uint8_t *nl = memrchr(data + pos, '\n', end - pos);
        const size_t take = nl ? (size_t)(nl - (data + pos)) + 1 : 0;
        if (take && pinned) {               /* true DMA straight from the caller's buffer */
            rc = qk_submit_raw(ctx, slot, data + pos, take);
            pos += take;
        } else {
            if (!take && (end < n || end - pos >= cap)) return QK_ERR_ARG; /* a line longer than a chunk */
            rc = qk_wait_slot(ctx, slot);
            if (rc) return rc;
            uint8_t *host = qk_slot_host_buffer(ctx, slot);
            if (take) {
                memcpy(host, data + pos, take);
                rc = qk_submit_raw(ctx, slot, host, take);
                pos += take;
            } else {                        /* T9: last line without '\n' -- we terminate it */
                memcpy(host, data + pos, end - pos);
                host[end - pos] = '\n';
                rc = qk_submit_raw(ctx, slot, host, end - pos + 1);
                pos = end;
                unterminated++;
            }
        }
        if (rc) return rc;
    }
    return raw_finish(ctx, st, n, unterminated, fastq);
}


int qk_count_raw_fd(qk_ctx *ctx, int fd, int seekable, qk_framer_stats *st)
{
    int dupfd = dup(fd);                     /* the stream owns its descriptor; ours stays with the caller */
    if (dupfd < 0) return QK_ERR_IO;
    qk_stream *s = qk_stream_open_fd(dupfd, seekable);
    if (!s) { close(dupfd); return QK_ERR_IO; }
    int rc = qk_count_raw_stream(ctx, s, st);
    qk_stream_close(s);
    return rc;
}

int qk_count_raw_stream(qk_ctx *ctx, qk_stream *in, qk_framer_stats *st)
{
    if (!in) return QK_ERR_ARG;
    const int seekable = qk_stream_seekable(in);
    uint32_t n_slots = 0;
    size_t cap = 0;
    int rc = qk_ctx_info(ctx, &n_slots, &cap);
    if (rc) return rc;
    uint8_t *tail = malloc(cap);            /* the partial last line of the previous piece */
    if (!tail) return QK_ERR_NOMEM;
    size_t tail_len = 0;
    uint64_t raw_bytes = 0, unterminated = 0;
    int fastq = 0, started = 0, eof = 0;
    uint32_t slot = 0;
    while (!eof || tail_len) {
        rc = qk_wait_slot(ctx, slot);       /* "find an idle worker", Q.c:433-437 */
        if (rc) break;
        uint8_t *host = qk_slot_host_buffer(ctx, slot);
        memcpy(host, tail, tail_len);
        size_t have = tail_len;
        tail_len = 0;
        if (!eof && have < cap) {            /* plain or inflated bytes, as many as fit */
            ssize_t got = qk_stream_read(in, host + have, cap - have);
            if (got < 0) { rc = QK_ERR_IO; break; }
            if ((size_t)got < cap - have) eof = 1;
            have += (size_t)got;
        }
        if (rc || have == 0) break;
        if (!started) {
            int skip_first;
            raw_mode(host[0], seekable, &fastq, &skip_first);
            rc = qk_raw_begin(ctx, fastq, skip_first);
            if (rc) break;
            started = 1;
        }
        const uint8_t *nl = memrchr(host, '\n', have);
        size_t take = nl ? (size_t)(nl - host) + 1 : 0;
        if (eof && take < have) {           /* T9: unterminated last line */
            if (have >= cap) { rc = QK_ERR_ARG; break; }
            host[have] = '\n';
            raw_bytes += have;
            take = have + 1;
            have = take;
            unterminated++;
        } else {
            if (!take) { rc = QK_ERR_ARG; break; } /* a line longer than a chunk */
            raw_bytes += take;
        }
        tail_len = have - take;
        memcpy(tail, host + take, tail_len);
        rc = qk_submit_raw(ctx, slot, host, take); /* "sem_post", Q.c:431-432 */
        if (rc) break;
        slot = (slot + 1) % n_slots;
    }
    free(tail);
    if (rc) return rc;
    if (!started) {
        rc = qk_raw_begin(ctx, 0, 0);
        if (rc) return rc;
    }
    return raw_finish(ctx, st, raw_bytes, unterminated, fastq);
}


/* ---- sharding one reads file over several GPUs ------------------------------------------- */
/* first line start at or after `at` (a line starts at 0 and after every '\n') */
static int64_t line_start_at_or_after(int fd, uint64_t at, uint64_t size)
{
    if (at == 0) return 0;
    uint8_t buf[65536];
    uint64_t pos = at - 1;              /* if byte at-1 is '\n', `at` itself is a line start */
    while (pos < size) {
        ssize_t got = pread(fd, buf, sizeof buf, (off_t)pos);
        if (got < 0 && errno == EINTR) continue;
        if (got <= 0) return got < 0 ? -1 : (int64_t)size;
        const uint8_t *nl = memchr(buf, '\n', (size_t)got);
        if (nl) return (int64_t)(pos + (uint64_t)(nl - buf) + 1);
        pos += (uint64_t)got;
    }
    return (int64_t)size;
}

int qk_shard_bounds(const char *reads_path, uint32_t rank, uint32_t world, uint64_t *begin, uint64_t *end)
{
    if (!reads_path || !begin || !end || world == 0 || rank >= world) return QK_ERR_ARG;
    int fd = open(reads_path, O_RDONLY);
    if (fd < 0) return QK_ERR_IO;
    struct stat sb;
    if (fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode)) { close(fd); return QK_ERR_IO; }
    const uint64_t size = (uint64_t)sb.st_size;
    int64_t b = line_start_at_or_after(fd, size / world * rank, size);
    int64_t e = rank + 1 == world ? (int64_t)size : line_start_at_or_after(fd, size / world * (rank + 1), size);
    close(fd);
    if (b < 0 || e < 0) return QK_ERR_IO;
    *begin = (uint64_t)b;
    *end = (uint64_t)e;
    return QK_OK;
}

/* Guess the line state of a FASTQ stream at a line start from the next few lines: find a line
 * i starting with '@' whose line i+2 starts with '+' and whose lines i+1 and i+3 have equal
 * length -- a record header, examined in state 3 -- so the state at line 0 is (3 - i) mod 4.
 * A guess only: callers verify it against the true state handed on by the previous shard. */
int qk_fastq_state_guess(const uint8_t *window, size_t n, uint32_t *line_state)
{
    if (!window || !line_state) return QK_ERR_ARG;
    size_t start[12], len[12];
    int nl = 0;
    size_t pos = 0;
    while (nl < 12 && pos < n) {
        const uint8_t *e = memchr(window + pos, '\n', n - pos);
        if (!e) break;
        start[nl] = pos;
        len[nl] = (size_t)(e - (window + pos));
        pos += len[nl] + 1;
        ++nl;
    }
    for (int i = 0; i + 3 < nl && i < 8; ++i)
        if (window[start[i]] == '@' && len[i + 2] >= 1 && window[start[i + 2]] == '+' && len[i + 1] == len[i + 3]) {
            *line_state = (uint32_t)((3 - i) & 3);
            return QK_OK;
        }
    *line_state = 0;
    return QK_ERR_FORMAT;
}

static int count_range_mt(qk_ctx *ctx, int fd, uint64_t begin, uint64_t end, uint32_t threads, uint64_t *unterminated);
#define QK_HEAD ((size_t)128 << 10) /* >= the longest line the reference reads (100,000 bytes) */

int qk_count_raw_range(qk_ctx *ctx, const char *reads_path, uint64_t begin, uint64_t end, int fastq, uint32_t line_state,
                       qk_framer_stats *st, uint32_t *final_state)
{
    return qk_count_raw_range_mt(ctx, reads_path, begin, end, fastq, line_state, 0, st, final_state);
}

int qk_count_raw_range_mt(qk_ctx *ctx, const char *reads_path, uint64_t begin, uint64_t end, int fastq, uint32_t line_state,
                          uint32_t threads, qk_framer_stats *st, uint32_t *final_state)
{
    uint32_t n_slots = 0;
    size_t cap = 0;
    int rc = qk_ctx_info(ctx, &n_slots, &cap);
    if (rc) return rc;
    int fd = open(reads_path, O_RDONLY);
    if (fd < 0) return QK_ERR_IO;
    rc = qk_raw_begin_state(ctx, fastq, line_state);
    uint64_t unterminated = 0, pos = begin;
    uint32_t slot = 0;
    if (!rc && cap >= 4 * QK_HEAD && end > begin) {      /* reader threads fill the pinned slots in parallel */
        rc = count_range_mt(ctx, fd, begin, end, threads ? threads : qk_reader_threads_default(), &unterminated);
        pos = end;
    }
    while (!rc && pos < end) {
        rc = qk_wait_slot(ctx, slot);
        if (rc) break;
        uint8_t *host = qk_slot_host_buffer(ctx, slot);
        size_t want = end - pos > cap ? cap : (size_t)(end - pos), have = 0;
        while (have < want) {
            ssize_t got = pread(fd, host + have, want - have, (off_t)(pos + have));
            if (got < 0 && errno == EINTR) continue;
            if (got <= 0) break;
            have += (size_t)got;
        }
        if (have == 0) { rc = QK_ERR_IO; break; }
        const uint8_t *nl = memrchr(host, '\n', have);
        size_t take = nl ? (size_t)(nl - host) + 1 : 0;
        if (pos + have >= end && take < have) {           /* unterminated last line of the file */
            if (have >= cap) { rc = QK_ERR_ARG; break; }
            host[have] = '\n';
            rc = qk_submit_raw(ctx, slot, host, have + 1);
            unterminated++;
            pos += have;
        } else {
            if (!take) { rc = QK_ERR_ARG; break; }
            rc = qk_submit_raw(ctx, slot, host, take);
            pos += take;
        }
        slot = (slot + 1) % n_slots;
    }
    close(fd);
    if (rc) return rc;
    if (final_state) {
        rc = qk_raw_state(ctx, final_state);
        if (rc) return rc;
    }
    return raw_finish(ctx, st, end - begin, unterminated, fastq);
}


/* ---- parallel ingest of a regular file -----------------------------------------------------
 * The reference has ONE producer thread (Q.c:397-456) and is bound by it.  Here the producer's
 * only per-byte work is getting the bytes into pinned memory, and that is what is
 * parallelised: reader threads pread() fixed-size pieces of the range straight into the slots'
 * pinned buffers (at offset QK_HEAD), the submitting thread takes the pieces in order, puts
 * the partial last line of the previous piece in front (that is what the QK_HEAD bytes of
 * headroom are for), cuts at the last '\n' and enqueues H2D + framing + counting. */
typedef struct {
    qk_ctx *ctx;
    int fd;
    uint64_t begin, end;
    size_t body;                 /* file bytes per piece */
    uint32_t n_slots;
    uint64_t n_pieces;
    pthread_mutex_t mu;
    pthread_cond_t cv;
    uint64_t next_piece;         /* next piece a reader may claim            */
    uint64_t submitted;          /* pieces the submitting thread is done with */
    uint64_t filled[QK_HOST_MAX_SLOTS]; /* piece index + 1 sitting in each slot, 0 = none */
    size_t filled_len[QK_HOST_MAX_SLOTS];
    size_t head;                 /* bytes of headroom in front of each piece */
    int err;
} qk_ingest;

static void *ingest_reader(void *arg)
{
    qk_ingest *g = arg;
    for (;;) {
        pthread_mutex_lock(&g->mu);
        const uint64_t i = g->next_piece;
        if (i >= g->n_pieces || g->err) { pthread_mutex_unlock(&g->mu); return NULL; }
        g->next_piece++;
        while (!g->err && i >= g->submitted + g->n_slots) pthread_cond_wait(&g->cv, &g->mu); /* slot still holds piece i - n_slots */
        pthread_mutex_unlock(&g->mu);
        const uint32_t slot = (uint32_t)(i % g->n_slots);
        int rc = qk_wait_slot(g->ctx, slot);     /* its last H2D has left the pinned buffer */
        uint8_t *host = qk_slot_host_buffer(g->ctx, slot) + g->head;
        const uint64_t at = g->begin + i * g->body;
        const size_t want = g->end - at > g->body ? g->body : (size_t)(g->end - at);
        size_t have = 0;
        while (!rc && have < want) {
            ssize_t got = pread(g->fd, host + have, want - have, (off_t)(at + have));
            if (got < 0 && errno == EINTR) continue;
            if (got <= 0) { rc = QK_ERR_IO; break; }
            have += (size_t)got;
        }
        pthread_mutex_lock(&g->mu);
        if (rc) g->err = rc;
        g->filled[slot] = i + 1;
        g->filled_len[slot] = have;
        pthread_cond_broadcast(&g->cv);
        pthread_mutex_unlock(&g->mu);
    }
}

static int count_range_mt(qk_ctx *ctx, int fd, uint64_t begin, uint64_t end, uint32_t threads, uint64_t *unterminated)
{
    qk_ingest g;
    memset(&g, 0, sizeof g);
    size_t cap = 0;
    int rc = qk_ctx_info(ctx, &g.n_slots, &cap);
    if (rc) return rc;
    if (cap < 4 * QK_HEAD || g.n_slots > QK_HOST_MAX_SLOTS) return QK_ERR_ARG;
    g.ctx = ctx;
    g.fd = fd;
    g.begin = begin;
    g.end = end;
    g.head = QK_HEAD;
    g.body = cap - QK_HEAD - 1;
    g.n_pieces = (end - begin + g.body - 1) / g.body;
    pthread_mutex_init(&g.mu, NULL);
    pthread_cond_init(&g.cv, NULL);
    if (threads > g.n_slots) threads = g.n_slots;
    if (threads < 1) threads = 1;
    pthread_t th[QK_HOST_MAX_SLOTS];
    uint32_t started = 0;
    for (; started < threads; ++started)
        if (pthread_create(&th[started], NULL, ingest_reader, &g) != 0) break;
    if (started == 0) rc = QK_ERR_NOMEM;
    uint8_t *tail = malloc(QK_HEAD);
    size_t tail_len = 0;
    if (!tail) rc = QK_ERR_NOMEM;
    for (uint64_t i = 0; !rc && i < g.n_pieces; ++i) {
        const uint32_t slot = (uint32_t)(i % g.n_slots);
        pthread_mutex_lock(&g.mu);
        while (!g.err && g.filled[slot] != i + 1) pthread_cond_wait(&g.cv, &g.mu);
        rc = g.err;
        size_t have = g.filled_len[slot];
        pthread_mutex_unlock(&g.mu);
        if (rc) break;
        uint8_t *body = qk_slot_host_buffer(ctx, slot) + QK_HEAD;
        uint8_t *from = body - tail_len;
        memcpy(from, tail, tail_len);
        size_t total = tail_len + have;
        const uint8_t *nl = memrchr(from, '\n', total);
        size_t take = nl ? (size_t)(nl - from) + 1 : 0;
        if (i + 1 == g.n_pieces && take < total) {      /* unterminated last line of the range */
            from[total] = '\n';                          /* body is one byte short of the buffer end */
            take = ++total;
            ++*unterminated;
        }
        tail_len = total - take;
        if (tail_len > QK_HEAD) { rc = QK_ERR_ARG; break; } /* a line longer than 128 KiB */
        memcpy(tail, from + take, tail_len);
        if (take) rc = qk_submit_raw(ctx, slot, from, take);
        pthread_mutex_lock(&g.mu);
        g.submitted = i + 1;
        pthread_cond_broadcast(&g.cv);
        pthread_mutex_unlock(&g.mu);
    }
    pthread_mutex_lock(&g.mu);
    if (rc && !g.err) g.err = rc;                        /* stop the readers */
    pthread_cond_broadcast(&g.cv);
    pthread_mutex_unlock(&g.mu);
    for (uint32_t t = 0; t < started; ++t) pthread_join(th[t], NULL);
    free(tail);
    pthread_mutex_destroy(&g.mu);
    pthread_cond_destroy(&g.cv);
    return rc ? rc : g.err;
}

/* A file range of fixed-size elements through the slots' pinned buffers: reader threads pread() pieces in parallel,
 * this thread hands them over IN ORDER: handle(ctx, slot, element offset, element count, user). */
int qk_ingest_elements(qk_ctx *ctx, int fd, uint64_t file_off, uint64_t n_elems, size_t esz, uint32_t threads,
                       qk_piece_handler handle, void *user)
{
    qk_ingest g;
    memset(&g, 0, sizeof g);
    size_t cap = 0;
    int rc = qk_ctx_info(ctx, &g.n_slots, &cap);
    if (rc) return rc;
    g.ctx = ctx;
    g.fd = fd;
    g.begin = file_off;
    g.end = file_off + n_elems * esz;
    g.head = 0;
    g.body = cap / 8 * 8;                     /* whole elements per piece */
    g.n_pieces = (g.end - g.begin + g.body - 1) / g.body;
    pthread_mutex_init(&g.mu, NULL);
    pthread_cond_init(&g.cv, NULL);
    if (threads > g.n_slots) threads = g.n_slots;
    if (threads < 1) threads = 1;
    pthread_t th[QK_HOST_MAX_SLOTS];
    uint32_t started = 0;
    for (; started < threads; ++started)
        if (pthread_create(&th[started], NULL, ingest_reader, &g) != 0) break;
    if (started == 0) rc = QK_ERR_NOMEM;
    for (uint64_t i = 0; !rc && i < g.n_pieces; ++i) {
        const uint32_t slot = (uint32_t)(i % g.n_slots);
        pthread_mutex_lock(&g.mu);
        while (!g.err && g.filled[slot] != i + 1) pthread_cond_wait(&g.cv, &g.mu);
        rc = g.err;
        const size_t have = g.filled_len[slot];
        pthread_mutex_unlock(&g.mu);
        if (rc) break;
        rc = handle(ctx, slot, i * (g.body / esz), have / esz, user);
        pthread_mutex_lock(&g.mu);
        g.submitted = i + 1;
        pthread_cond_broadcast(&g.cv);
        pthread_mutex_unlock(&g.mu);
    }
    pthread_mutex_lock(&g.mu);
    if (rc && !g.err) g.err = rc;
    pthread_cond_broadcast(&g.cv);
    pthread_mutex_unlock(&g.mu);
    for (uint32_t t = 0; t < started; ++t) pthread_join(th[t], NULL);
    pthread_mutex_destroy(&g.mu);
    pthread_cond_destroy(&g.cv);
    return rc ? rc : g.err;
}

static int qm_piece(qk_ctx *ctx, uint32_t slot, uint64_t elem_offset, uint64_t count, void *user)
{
    return qk_dict_upload_from_slot(ctx, slot, *(int *)user, elem_offset, count);
}

static int qm_upload_array(qk_ctx *ctx, int fd, uint64_t file_off, uint64_t n_elems, int kind, uint32_t threads)
{
    return qk_ingest_elements(ctx, fd, file_off, n_elems, kind ? 4 : 8, threads, qm_piece, &kind);
}

static int gc_piece(qk_ctx *ctx, uint32_t slot, uint64_t elem_offset, uint64_t count, void *user)
{
    (void)user;
    return qk_gc_from_slot(ctx, slot, elem_offset, count);
}

/* Q.c:484-488, 495-509: the .qgc flags of every dictionary k-mer against the final depths.  The file is streamed
 * to the device like the dictionary was; entries it lacks are taken as non-control (*entries_read tells). */
int qk_gc_curve_file(qk_ctx *ctx, const char *qgc_path, uint64_t n_kmers, uint64_t sum[QK_GC_BINS], int64_t sumsq[QK_GC_BINS],
                     uint64_t count[QK_GC_BINS], uint64_t *entries_read, uint64_t *bins_out_of_range)
{
    if (!ctx || !qgc_path) return QK_ERR_ARG;
    int fd = open(qgc_path, O_RDONLY);
    if (fd < 0) return QK_ERR_IO;
    struct stat sb;
    if (fstat(fd, &sb) != 0) { close(fd); return QK_ERR_IO; }
    uint64_t n = (uint64_t)sb.st_size / 2;
    if (n > n_kmers) n = n_kmers;
    if (entries_read) *entries_read = n;
    int rc = qk_gc_begin(ctx);
    if (!rc && n) rc = qk_ingest_elements(ctx, fd, 0, n, 2, qk_reader_threads_default(), gc_piece, NULL);
    close(fd);
    if (rc) return rc;
    return qk_gc_end(ctx, sum, sumsq, count, bins_out_of_range);
}

static int file_is_gzip(const char *path)
{
    uint8_t magic[2] = {0, 0};
    int fd = open(path, O_RDONLY);
    if (fd < 0) return 0;
    ssize_t got = pread(fd, magic, 2, 0);
    close(fd);
    return got == 2 && magic[0] == 0x1f && magic[1] == 0x8b;
}

uint32_t qk_reader_threads_default(void)
{
    const char *e = getenv("QK_READER_THREADS");
    if (e && atoi(e) > 0) return (uint32_t)atoi(e);
    long cpus = sysconf(_SC_NPROCESSORS_ONLN);       /* measured: ~3 GB/s of pread per thread, flat beyond ~16 */
    if (cpus < 1) cpus = 1;
    return (uint32_t)(cpus > 8 ? 8 : cpus);          /* callers cap it at the slot count */
}

int qk_count_raw_file_mt(qk_ctx *ctx, const char *reads_path, uint32_t threads, qk_framer_stats *st)
{
    int fd = open(reads_path, O_RDONLY);
    if (fd < 0) return QK_ERR_IO;
    struct stat sb;
    size_t cap = 0;
    qk_ctx_info(ctx, NULL, &cap);
    if (fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode) || sb.st_size == 0 || cap < 4 * QK_HEAD || file_is_gzip(reads_path)) {
        /* pipe, device, empty file, gzip, or slots too small for the headroom: sequential path */
        int seekable = lseek(fd, 0, SEEK_CUR) != (off_t)-1;
        int rc = qk_count_raw_fd(ctx, fd, seekable, st);
        close(fd);
        return rc;
    }
    uint8_t first = 0;
    int fastq = 0, skip_first = 0;
    int rc = pread(fd, &first, 1, 0) == 1 ? QK_OK : QK_ERR_IO;
    if (!rc) {
        raw_mode(first, 1, &fastq, &skip_first);
        rc = qk_raw_begin(ctx, fastq, skip_first);
    }
    uint64_t unterminated = 0;
    if (!rc) rc = count_range_mt(ctx, fd, 0, (uint64_t)sb.st_size, threads ? threads : qk_reader_threads_default(), &unterminated);
    close(fd);
    if (rc) return rc;
    return raw_finish(ctx, st, (uint64_t)sb.st_size, unterminated, fastq);
}

int qk_count_raw_file(qk_ctx *ctx, const char *reads_path, qk_framer_stats *st)
{
    return qk_count_raw_file_mt(ctx, reads_path, 0, st);
}


/* ---- one reads file over the GPUs of a qk_multi ---------------------------------------------
 * Shard r = the line-aligned r-th part of the file (qk_shard_bounds), counted by context r with
 * its own reader threads.  FASTA shards start in line state 0; FASTQ shards start in the state
 * guessed from their first lines, all at once, and afterwards every assumed state is checked
 * against the state its predecessor really ended in -- a shard whose guess was wrong (malformed
 * FASTQ only) is zeroed and recounted from the true state, then the check repeats. */
typedef struct {
    qk_ctx *ctx;
    const char *path;
    uint64_t begin, end;
    int fastq;
    uint32_t state, final_state, threads;
    qk_framer_stats st;
    int rc, todo;
} shard_job;

static void *shard_worker(void *arg)
{
    shard_job *j = arg;
    j->rc = qk_count_raw_range_mt(j->ctx, j->path, j->begin, j->end, j->fastq, j->state, j->threads, &j->st, &j->final_state);
    return NULL;
}

int qk_count_file_multi(qk_multi *m, const char *reads_path, uint32_t threads_per_gpu, qk_framer_stats *st)
{
    const uint32_t n = qk_multi_size(m);
    if (!m || n == 0 || !reads_path) return QK_ERR_ARG;
    struct stat sb;
    if (n == 1 || stat(reads_path, &sb) != 0 || !S_ISREG(sb.st_mode) || sb.st_size == 0 || file_is_gzip(reads_path))
        return qk_count_raw_file_mt(qk_multi_ctx(m, 0), reads_path, threads_per_gpu, st); /* pipes, gzip: one GPU */
    int fd = open(reads_path, O_RDONLY);
    if (fd < 0) return QK_ERR_IO;
    uint8_t first = 0;
    if (pread(fd, &first, 1, 0) != 1) { close(fd); return QK_ERR_IO; }
    const int fastq = first == '@';
    shard_job job[QK_HOST_MAX_SLOTS];
    memset(job, 0, sizeof job);
    uint8_t *window = malloc(1 << 20);
    int rc = window ? QK_OK : QK_ERR_NOMEM;
    for (uint32_t r = 0; !rc && r < n; ++r) {
        shard_job *j = &job[r];
        j->ctx = qk_multi_ctx(m, r);
        j->path = reads_path;
        j->fastq = fastq;
        j->threads = threads_per_gpu;
        j->todo = 1;
        rc = qk_shard_bounds(reads_path, r, n, &j->begin, &j->end);
        if (rc) break;
        if (r == 0) j->state = fastq ? 3 : 0;            /* the first line of a FASTQ is consumed (Q.c:393-395) */
        else if (fastq) {
            ssize_t got = pread(fd, window, 1 << 20, (off_t)j->begin);
            qk_fastq_state_guess(window, got > 0 ? (size_t)got : 0, &j->state);
        }
    }
    free(window);
    close(fd);
    for (uint32_t round = 0; !rc && round <= n; ++round) {
        pthread_t th[QK_HOST_MAX_SLOTS];
        for (uint32_t r = 0; r < n; ++r)
            if (job[r].todo && pthread_create(&th[r], NULL, shard_worker, &job[r]) != 0) { job[r].rc = QK_ERR_NOMEM; job[r].todo = 2; }
        for (uint32_t r = 0; r < n; ++r) {
            if (job[r].todo == 1) pthread_join(th[r], NULL);
            if (job[r].todo && job[r].rc) rc = job[r].rc;
            job[r].todo = 0;
        }
        if (rc) break;
        uint32_t carry = job[0].state, bad = n;
        for (uint32_t r = 0; r < n; ++r) {               /* the first shard whose assumed state was wrong */
            if (job[r].end == job[r].begin) continue;
            if (r > 0 && job[r].state != carry) { bad = r; break; }
            carry = job[r].final_state;
        }
        if (bad == n) break;
        job[bad].state = carry;
        job[bad].todo = 1;
        rc = qk_reset_counters(job[bad].ctx);
    }
    if (rc || !st) return rc;
    memset(st, 0, sizeof *st);
    st->fastq = fastq;
    for (uint32_t r = 0; r < n; ++r) {
        st->lines += job[r].st.lines;
        st->bases += job[r].st.bases;
        st->raw_bytes += job[r].st.raw_bytes;
        st->unterminated += job[r].st.unterminated;
    }
    return QK_OK;
}

