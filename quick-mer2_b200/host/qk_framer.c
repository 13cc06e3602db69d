/*
 * qk_framer.c -- the reference's fgets framing loop (Q.c:393-398, 451-455) on the host, and the
 * driver that feeds framed chunks to the device.  The default path frames on the device
 * (csrc/qk_frame.cu); this one is the alternative (QK_HOST_FRAMER=1) and the CPU-testable
 * statement of the rules.  See include/qk_host.h.
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

/* ------------------------------------------------------------------ framer ----------- */

struct qk_framer {
    int fd;             /* -1 for in-memory input */
    int seekable;
    int own_buf;
    uint8_t *buf;
    size_t cap, pos, have;
    int eof;
    int started;        /* first line seen */
    int skip;           /* FASTQ: lines still to discard after a read (Q.c:451-455) */
    qk_framer_stats st;
};

static qk_framer *framer_new(void)
{
    qk_framer *f = calloc(1, sizeof *f);
    if (f) f->fd = -1;
    return f;
}

qk_framer *qk_framer_open_fd(int fd, int seekable)
{
    qk_framer *f = framer_new();
    if (!f) return NULL;
    f->fd = fd;
    f->seekable = seekable;
    f->cap = (size_t)8 << 20;
    f->buf = malloc(f->cap);
    f->own_buf = 1;
    if (!f->buf) { free(f); return NULL; }
    return f;
}

qk_framer *qk_framer_open(const char *path)
{
    int fd = open(path, O_RDONLY);
    if (fd < 0) return NULL;
    /* Q.c:396 fseek(0): works on regular files, fails silently on pipes */
    int seekable = lseek(fd, 0, SEEK_CUR) != (off_t)-1;
    qk_framer *f = qk_framer_open_fd(fd, seekable);
    if (!f) close(fd);
    return f;
}

qk_framer *qk_framer_open_mem(const uint8_t *data, size_t n, int seekable)
{
    qk_framer *f = framer_new();
    if (!f) return NULL;
    f->buf = (uint8_t *)data;
    f->cap = f->have = n;
    f->eof = 1;
    f->seekable = seekable;
    return f;
}

void qk_framer_close(qk_framer *f)
{
    if (!f) return;
    if (f->fd >= 0) close(f->fd);
    if (f->own_buf) free(f->buf);
    free(f);
}

void qk_framer_get_stats(const qk_framer *f, qk_framer_stats *st)
{
    if (f && st) *st = f->st;
}

/* slide the unread tail to the front and read more; returns bytes added (0 at EOF) */
static long framer_refill(qk_framer *f)
{
    if (f->eof || f->fd < 0) { f->eof = 1; return 0; }
    if (f->pos) {
        memmove(f->buf, f->buf + f->pos, f->have - f->pos);
        f->have -= f->pos;
        f->pos = 0;
    }
    if (f->have == f->cap) { /* one line larger than the window: grow */
        uint8_t *nb = realloc(f->buf, f->cap * 2);
        if (!nb) return -1;
        f->buf = nb;
        f->cap *= 2;
    }
    for (;;) {
        ssize_t got = read(f->fd, f->buf + f->have, f->cap - f->have);
        if (got < 0 && errno == EINTR) continue;
        if (got < 0) return -1;
        if (got == 0) f->eof = 1;
        f->have += (size_t)got;
        return (long)got;
    }
}

int qk_framer_next(qk_framer *f, uint8_t *dst, size_t cap, size_t *n_bytes, uint32_t *line_off, uint32_t off_cap,
                   uint32_t *n_lines)
{
    if (!f || !dst || !n_bytes || cap < 100000) return -QK_ERR_ARG;
    size_t out = 0;
    uint32_t nl = 0;
    if (line_off) {
        if (off_cap < 2) return -QK_ERR_ARG;
        line_off[0] = 0;
    }
    for (;;) {
        size_t searched = 0;
        const uint8_t *line = f->buf + f->pos;
        const uint8_t *end = NULL;
        size_t avail = f->have - f->pos;
        if (avail) end = memchr(line, '\n', avail);
        int unterminated = 0;
        size_t len;
        (void)searched;
        if (!end) {
            if (!f->eof) {
                long got = framer_refill(f);
                if (got < 0) return -QK_ERR_IO;
                continue;
            }
            if (avail == 0) break;          /* end of input */
            unterminated = 1;               /* T9: reference is undefined; we terminate the line */
            len = avail;
        } else {
            len = (size_t)(end - line) + 1; /* includes the '\n' */
        }
        if (!f->started) {                  /* Q.c:393-396 */
            f->started = 1;
            if (line[0] == '@') { f->st.fastq = 1; goto consume; }
            if (!f->seekable) goto consume; /* fseek on a pipe fails: first line is lost */
        }
        if (f->skip) { f->skip--; goto consume; }
        if (line[0] == '>') goto consume;   /* Q.c:398 */
        {
            size_t need = len + (size_t)unterminated;
            if (out + need > cap || (line_off && nl + 2 > off_cap)) {
                if (out == 0) return -QK_ERR_ARG; /* a single line larger than the chunk */
                break;                            /* chunk full: leave the line for the next call */
            }
            memcpy(dst + out, line, len);
            if (unterminated) { dst[out + len] = '\n'; f->st.unterminated++; }
            out += need;
            ++nl;
            if (line_off) line_off[nl] = (uint32_t)out;
            f->st.lines++;
            f->st.bases += need - 1;
            if (need > QK_MAX_LINE_BYTES) f->st.long_lines++;
            if (f->st.fastq) f->skip = 3;   /* Q.c:451-455 */
        }
    consume:
        f->pos += len;
        f->st.raw_bytes += len;
    }
    *n_bytes = out;
    if (n_lines) *n_lines = nl;
    return out ? 1 : 0;
}


/* ------------------------------------------------------------------ driver ----------- */
int qk_count_framer(qk_ctx *ctx, qk_framer *f, qk_framer_stats *st)
{
    uint32_t n_slots = 0;
    size_t cap = 0;
    int rc = qk_ctx_info(ctx, &n_slots, &cap);
    if (rc) return rc;
    uint32_t slot = 0;
    for (;;) {
        rc = qk_wait_slot(ctx, slot);       /* "find an idle worker", Q.c:433-437 */
        if (rc) return rc;
        uint8_t *dst = qk_slot_host_buffer(ctx, slot);
        size_t n = 0;
        uint32_t nl = 0;
        int r = qk_framer_next(f, dst, cap, &n, NULL, 0, &nl);
        if (r < 0) return -r;
        if (r == 0) break;
        rc = qk_submit(ctx, slot, dst, n, NULL, nl); /* "sem_post", Q.c:431-432 */
        if (rc) return rc;
        slot = (slot + 1) % n_slots;
    }
    rc = qk_sync(ctx);                       /* drain + join, Q.c:458-479 */
    if (st) qk_framer_get_stats(f, st);
    return rc;
}

int qk_count_file(qk_ctx *ctx, const char *reads_path, qk_framer_stats *st)
{
    qk_framer *f = qk_framer_open(reads_path);
    if (!f) return QK_ERR_IO;
    int rc = qk_count_framer(ctx, f, st);
    qk_framer_close(f);
    return rc;
}

