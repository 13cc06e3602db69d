/*
 * qk_files.c -- the small file formats of `count`: QM11 header (Q.c:345-351), .bin (Q.c:498-518)
 * and .txt (Q.c:522-542) writers.  See include/qk_host.h.
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

/* ------------------------------------------------------------------ QM11 reader ------ */
int qk_qm_read_header(const char *qm_path, qk_qm_header *hdr)
{
    if (!qm_path || !hdr) return QK_ERR_ARG;
    FILE *f = fopen(qm_path, "rb");
    if (!f) return QK_ERR_IO;
    uint8_t raw[24];
    size_t got = fread(raw, 1, sizeof raw, f);
    fclose(f);
    if (got != sizeof raw) return QK_ERR_IO;
    hdr->k = raw[4];                        /* Q.c:345-346 */
    memcpy(&hdr->hash_size, raw + 8, 8);    /* Q.c:348-349 */
    memcpy(&hdr->first_idx, raw + 16, 8);   /* Q.c:350-351 */
    return QK_OK;
}


/* ------------------------------------------------------------------ writers ---------- */
int qk_write_bin(const char *path, const uint16_t *counts, uint64_t n)
{
    FILE *f = fopen(path, "wb");
    if (!f) return QK_ERR_IO;
    size_t w = fwrite(counts, sizeof(uint16_t), n, f); /* Q.c:512,517 */
    int rc = (w == n) ? QK_OK : QK_ERR_IO;
    if (fclose(f) != 0) rc = QK_ERR_IO;
    return rc;
}

/* A piece of the .bin (Q.c:510-513 flushes per 1 Mi entries; here 16 Mi) written at its offset by several
 * threads: one thread's write() into the page cache runs at 2-3 GB/s, which made the dump the longest
 * part of the command at human scale (a 4.7 GB .bin). */
#define QK_BIN_WRITERS 4
typedef struct { int fd; const uint8_t *src; size_t n; off_t at; int rc; } bin_job;
static void *bin_writer(void *arg)
{
    bin_job *j = arg;
    size_t done = 0;
    while (done < j->n) {
        ssize_t w = pwrite(j->fd, j->src + done, j->n - done, j->at + (off_t)done);
        if (w < 0 && errno == EINTR) continue;
        if (w <= 0) { j->rc = QK_ERR_IO; return NULL; }
        done += (size_t)w;
    }
    return NULL;
}

static int write_piece(void *user, const uint16_t *piece, uint64_t offset, uint64_t count)
{
    const int fd = *(int *)user;
    const size_t bytes = (size_t)count * sizeof(uint16_t);
    bin_job job[QK_BIN_WRITERS];
    pthread_t th[QK_BIN_WRITERS];
    int n = bytes >= ((size_t)4 << 20) ? QK_BIN_WRITERS : 1, started = 0, rc = QK_OK;
    const size_t per = (bytes / (size_t)n + 4095) & ~(size_t)4095;
    for (int t = 0; t < n; ++t) {
        const size_t a = (size_t)t * per, z = t + 1 == n ? bytes : (a + per < bytes ? a + per : bytes);
        job[t] = (bin_job){fd, (const uint8_t *)piece + a, a < z ? z - a : 0, (off_t)(offset * sizeof(uint16_t) + a), QK_OK};
        if (t + 1 == n || pthread_create(&th[t], NULL, bin_writer, &job[t]) != 0) bin_writer(&job[t]); /* the last share here */
        else ++started;
    }
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    for (int t = 0; t < n; ++t)
        if (job[t].rc) rc = job[t].rc;
    return rc;
}

int qk_write_bin_from_device(qk_ctx *ctx, const char *path)
{
    int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) return QK_ERR_IO;
    int rc = qk_finish_pieces(ctx, write_piece, &fd);
    if (close(fd) != 0 && !rc) rc = QK_ERR_IO;
    return rc;
}

int qk_write_gc_txt(const char *path, const uint64_t sum[QK_GC_BINS], const int64_t sumsq[QK_GC_BINS],
                    const uint64_t count[QK_GC_BINS], double *mean_depth)
{
    FILE *f = fopen(path, "w");
    if (!f) return QK_ERR_IO;
    double total_depth = 0;
    uint64_t total_count = 0;
    for (int i = 0; i < QK_GC_BINS; ++i) {          /* Q.c:529-538 */
        double curve = (double)sum[i], sd = (double)sumsq[i];
        uint32_t c32 = (uint32_t)count[i];            /* uint32_t Control_count, Q.c:497 */
        total_count += c32;
        total_depth += curve;
        if (c32) {
            curve /= c32;
            volatile double m2 = curve * curve;       /* keep the product rounded: no FMA */
            sd = sd / c32 - m2;
        }
        fprintf(f, "%.2f\t%f\t%i\t%f\n", i / 4.0, curve, (int)c32, sd);
    }
    if (mean_depth) *mean_depth = total_depth / (double)total_count; /* Q.c:539 */
    return fclose(f) == 0 ? QK_OK : QK_ERR_IO;
}

