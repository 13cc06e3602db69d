/*
 * qk_command.c -- `quicKmer2 count` itself: main_count, Q.c:304-545.  Same arguments, same files,
 * same stdout lines.  See include/qk_host.h.
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

/* ------------------------------------------------------------------ command ---------- */
static void help_count(void)
{
    puts("\nquicKmer2 count [Options] ref.fa sample.fast[a/q] Out_prefix\n\nOptions:");
    puts("-h\t\tShow this help information");
    puts("-t [num]\tNumber of host threads framing / reading the input into pinned memory (counting runs on the GPU)");
    puts("-g [list]\tCUDA device index, or a comma-separated list to shard the reads over several GPUs (default 0)");
}

static double now_sec(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + ts.tv_nsec * 1e-9;
}

int qk_count_main(int argc, char **argv)
{
    int devices[QK_HOST_MAX_SLOTS] = {0};
    uint32_t n_dev = 1;
    unsigned threads = 0;
    if (argc < 2) { help_count(); return 1; }          /* Q.c:309-312 */
    int opt;
    optind = 1;
    while ((opt = getopt(argc, argv, "ht:g:")) != -1) { /* Q.c:314-333 */
        switch (opt) {
        case 'h': help_count(); return 1;
        case 't':
            threads = (uint8_t)atoi(optarg);            /* uint8_t thread_count, Q.c:306 */
            printf("[Option] Set %u threads\n", threads);
            break;
        case 'g':                                       /* one device or a comma-separated list */
            n_dev = 0;
            for (const char *p = optarg; *p && n_dev < QK_HOST_MAX_SLOTS;) {
                devices[n_dev++] = atoi(p);
                p = strchr(p, ',');
                if (!p) break;
                ++p;
            }
            if (n_dev == 0) { puts("Option error, check help"); help_count(); return 1; }
            break;
        case '?': puts("Option error, check help"); help_count(); return 1;
        default: return 1;
        }
    }
    if (argc < 4) { help_count(); return 1; }
    const char *ref_prefix = argv[argc - 3], *reads = argv[argc - 2], *out_prefix = argv[argc - 1]; /* Q.c:335-342 */
    char path[65536];
    snprintf(path, sizeof path, "%s.qm", ref_prefix);
    qk_qm_header hdr;
    if (qk_qm_read_header(path, &hdr) != QK_OK) {
        printf("Dictionary %s open fail\n", path);
        return 1;
    }
    const int host_framer = getenv("QK_HOST_FRAMER") != NULL; /* default: the device frames the raw stream */
    qk_framer *fr = NULL;
    int pipe_fd = -1;                                   /* >= 0: the input is not seekable (README.md:89-90) */
    if (host_framer) fr = qk_framer_open(reads);
    else {
        int probe = open(reads, O_RDONLY);
        if (probe < 0) { puts("Input open fail"); return 1; } /* Q.c:339-341 (the reference goes on and crashes) */
        if (lseek(probe, 0, SEEK_CUR) != (off_t)-1) close(probe); /* regular file: the drivers reopen it */
        else pipe_fd = probe;                           /* a pipe can be opened only once: keep it */
    }
    if (host_framer && !fr) { puts("Input open fail"); return 1; }
    printf("Hash Size: 0x%lX\nFirst location: 0x%lX\n", (unsigned long)hdr.hash_size, (unsigned long)hdr.first_idx);

    double t0 = now_sec();
    qk_multi *m = NULL;
    int rc = qk_multi_create(&m, devices, n_dev, 8, (size_t)32 << 20);
    if (rc) {
        printf("GPU context failed: %s\n", m ? qk_multi_last_error(m) : "no CUDA device");
        qk_multi_destroy(m);
        return 1;
    }
    qk_ctx *ctx = qk_multi_ctx(m, 0);
    uint64_t n_kmers = 0;
    if (getenv("QK_TIMING")) fprintf(stderr, "[qk] contexts (pinned + device slots) %.3f s\n", now_sec() - t0);
    rc = qk_qm_load(ctx, path, NULL, &n_kmers);
    if (rc) {
        printf("Dictionary load failed: %s\n", rc == QK_ERR_IO ? "short read" : qk_last_error(ctx));
        if (rc == QK_ERR_NOMEM) puts("Memory allocation failed"); /* Q.c:355,362 */
        qk_multi_destroy(m);
        return 1;
    }
    rc = qk_multi_replicate(m);                          /* ncclBroadcast of the table to the other GPUs */
    if (rc) { printf("Dictionary broadcast failed: %s\n", qk_multi_last_error(m)); qk_multi_destroy(m); return 1; }
    printf("Read 0x%lX hash\n", (unsigned long)hdr.hash_size);            /* Q.c:359 */
    double t1 = now_sec();
    time_t start_time, end_time;
    time(&start_time);                                                     /* Q.c:387 */
    qk_framer_stats st;
    if (host_framer) {
        rc = qk_count_framer(ctx, fr, &st);
        qk_framer_close(fr);
    } else if (pipe_fd >= 0) {
        rc = qk_count_raw_fd(ctx, pipe_fd, 0, &st);      /* a pipe feeds one GPU */
        close(pipe_fd);
    } else {
        /* Who frames?  With enough host cores per GPU the host does (all cores, sequence lines only over
         * the link, chunks to whichever GPU is free: host/qk_framer_mt.c); otherwise the raw stream is
         * shipped and the device frames it (csrc/qk_frame.cu).  QK_FRAMER=host|device overrides.
         * -t N: framer threads in all / reader threads per GPU (0 = default). */
        const char *pol = getenv("QK_FRAMER");
        const long cpus = sysconf(_SC_NPROCESSORS_ONLN);
        /* measured: 12+ framer threads beat the link-bound device path on FASTQ (half the bytes to ship);
         * FASTA is nearly all sequence already, shipping it raw costs the host nothing */
        int by_host = pol ? !strcmp(pol, "host") : cpus / (long)n_dev >= 12;
        if (!pol && by_host) {
            int fd0 = open(reads, O_RDONLY);
            char c0 = 0;
            if (fd0 < 0 || read(fd0, &c0, 1) != 1 || c0 != '@') by_host = 0;
            if (fd0 >= 0) close(fd0);
        }
        if (by_host) {
            qk_ctx *ctxs[QK_HOST_MAX_SLOTS];
            for (uint32_t i = 0; i < n_dev; ++i) ctxs[i] = qk_multi_ctx(m, i);
            rc = qk_count_file_mt(ctxs, n_dev, reads, threads, &st);
        } else rc = qk_count_file_multi(m, reads, threads, &st);
    }
    uint64_t total = 0, hits = 0;
    for (uint32_t i = 0; !rc && i < n_dev; ++i) {
        uint64_t t = 0, h = 0;
        rc = qk_stats(qk_multi_ctx(m, i), &t, &h, NULL);
        total += t;
        hits += h;
    }
    /* Q.c:458-466: with -t N the reference fills up its last batch of 4,096 keys with zeros, and its workers look
     * them up like any other key (Q.c:284-291) -- seen in the .bin iff the empty slot Find_hash(0) stops at is on the
     * chain (a dictionary made by `index` from a list that holds the poly-A k-mer). */
    if (!rc && threads) rc = qk_add_depth(ctx, 0, (uint32_t)(4096 - total % 4096));
    if (!rc) rc = qk_multi_reduce(m);                    /* ncclReduce of the counters into GPU 0 */
    if (rc) {
        printf("Counting failed: %s / %s\n", qk_last_error(ctx), qk_multi_last_error(m));
        qk_multi_destroy(m);
        return 1;
    }
    time(&end_time);
    double t2 = now_sec();
    {   /* Q.c:446: one line per 2^30 k-mers processed.  The device counts them all in seconds, so the
         * lines are written once the total is known: the same lines, in the same place in the output.
         * (QK_PROGRESS_SHIFT lowers the 30 for tests.) */
        const char *e = getenv("QK_PROGRESS_SHIFT");
        const int shift = e && atoi(e) > 0 && atoi(e) < 63 ? atoi(e) : 30;
        for (uint64_t g = 1; g <= (total >> shift); ++g) printf("Read %liG kmers\n", (long)g);
    }
    printf("Counting elapse %u sec, total %lu kmers\n", (unsigned)(end_time - start_time), (unsigned long)total); /* Q.c:481 */
    printf("Pileup finish\nRead chain file %lu entries\n", (unsigned long)hdr.hash_size);                         /* Q.c:483 */

    snprintf(path, sizeof path, "%s.bin", out_prefix);                     /* Q.c:498-518, written as the pieces arrive */
    rc = qk_write_bin_from_device(ctx, path);
    if (rc) {
        printf("Cannot write %s: %s\n", path, rc == QK_ERR_IO ? "I/O error" : qk_last_error(ctx));
        qk_multi_destroy(m);
        return 1;
    }
    double t_bin = now_sec();

    snprintf(path, sizeof path, "%s.qgc", ref_prefix);                     /* Q.c:484-488 */
    uint64_t sum[QK_GC_BINS], cnt[QK_GC_BINS], qgc_entries = 0, big_bins = 0;
    int64_t sq[QK_GC_BINS];
    /* the .qgc streams through the pinned slots to the device, like the dictionary did (Q.c:495-509) */
    rc = qk_gc_curve_file(ctx, path, n_kmers, sum, sq, cnt, &qgc_entries, &big_bins);
    if (rc == QK_ERR_IO) printf("GC control file %s absent. Continue without GC correction!\n", path);
    else {
        if (rc) { printf("GC curve failed: %s\n", qk_last_error(ctx)); qk_multi_destroy(m); return 1; }
        /* Inputs on which the reference itself is undefined (it reuses stale buffer contents for a short
         * .qgc and indexes past its 401 bins, Q.c:499-508): handled deterministically here, and said aloud. */
        if (qgc_entries < n_kmers)
            fprintf(stderr, "quicKmer2_b200: %s holds %llu entries, the dictionary %llu: the missing ones are taken as non-control\n",
                    path, (unsigned long long)qgc_entries, (unsigned long long)n_kmers);
        if (big_bins)
            fprintf(stderr, "quicKmer2_b200: %s has %llu entries with a GC bin above 400: ignored in the curve\n", path,
                    (unsigned long long)big_bins);
        double mean = 0;
        snprintf(path, sizeof path, "%s.txt", out_prefix);                 /* Q.c:523-525 */
        if (qk_write_gc_txt(path, sum, sq, cnt, &mean)) { printf("Cannot write %s\n", path); qk_multi_destroy(m); return 1; }
        printf("Mean sequencing depth: %.2f\n", mean);                     /* Q.c:540 */
    }
    double t3 = now_sec();
    if (getenv("QK_TIMING")) fprintf(stderr, "[qk] .bin written in %.3f s, GC curve + .txt in %.3f s\n", t_bin - t2, t3 - t_bin);
    double kms = 0, hms = 0;
    uint64_t launches = 0;
    qk_timing(ctx, &kms, &hms, &launches);
    qk_multi_destroy(m);
    puts("Exit quicK-mer2 count");                                          /* Q.c:543 */
    fprintf(stderr,
            "{\"total_kmers\": %llu, \"hits\": %llu, \"lines\": %llu, \"bases\": %llu, \"fastq\": %d, "
            "\"n_kmers\": %llu, \"gpus\": %u, \"load_s\": %.3f, \"count_s\": %.3f, \"dump_s\": %.3f, \"kernel_ms\": %.3f, "
            "\"h2d_ms\": %.3f, \"launches\": %llu, \"threads_option\": %u}\n",
            (unsigned long long)total, (unsigned long long)hits, (unsigned long long)st.lines,
            (unsigned long long)st.bases, st.fastq, (unsigned long long)n_kmers, n_dev, t1 - t0, t2 - t1, t3 - t2, kms, hms,
            (unsigned long long)launches, threads);
    return 0;
}

