/*
 * qk_est_host.c -- `quicKmer2 est ref.fa sample_prefix output.bed` (main_estimate, Q.c:555-685) with the window
 * reduction on the device.  Same arguments, same input files (<ref>.qgc, <ref>.bed, <sample>.bin, <sample>.txt),
 * same stdout lines, the same `smooth_GC_mrsfast.py <sample>.txt` on the PATH for the LOWESS curve (401 float32 on
 * its stdout, Q.c:642-650) and the same bytes in output.bed -- including what the reference's loop does after the
 * last window (it leaves only the inner loop, so every further 1 MiB block of the .qgc prints that window once
 * more, divided again).  See include/qk_host.h.
 */
#define _GNU_SOURCE
#define _FILE_OFFSET_BITS 64
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include "qk_host_internal.h"

#define EST_BLOCK_ENTRIES ((uint64_t)(1024 * 1024) / 2)   /* Q.c:11,660: fread(gc_control, 1, buffer_size, ...) = 1 MiB of bytes */

static void help_est(void)                                /* Q.c:547-553 */
{
    puts("quicKmer2 est ref.fa sample_prefix output.bed");
    puts("\tref.fa\t\tPrefix to genome reference. Program requires .qgc and .bed definition");
    puts("\tsample_prefix\tPrefix to sample.bin");
    puts("\toutput.bed\tOutput bedfile for copy number");
    puts("\nNo options available\n");
}

typedef struct { char chrom[64], begin[64], end[64]; uint32_t left, right; } est_window;

static int est_piece(qk_ctx *ctx, uint32_t slot, uint64_t elem_offset, uint64_t count, void *user)
{
    return qk_est_upload_from_slot(ctx, slot, *(int *)user, elem_offset, count);
}

/* The reduction alone, for callers that have the curve: windows -> the reference's value per printed line.
 * Returns the number of lines (printed windows + the repeats after the last one) through *n_lines; values[i]
 * belongs to window line_window[i]. */
int qk_est_reduce(qk_ctx *ctx, const char *qgc_path, const char *bin_path, const float correction[QK_GC_BINS], double mean_depth,
                  const uint32_t *left, const uint32_t *right, uint64_t n_windows, double **values_out, uint64_t **line_window_out,
                  uint64_t *n_lines)
{
    if (!ctx || !qgc_path || !bin_path || !correction || !values_out || !line_window_out || !n_lines) return QK_ERR_ARG;
    *values_out = NULL;
    *line_window_out = NULL;
    *n_lines = 0;
    int fg = open(qgc_path, O_RDONLY), fb = open(bin_path, O_RDONLY);
    struct stat sg, sb;
    if (fg < 0 || fb < 0 || fstat(fg, &sg) != 0 || fstat(fb, &sb) != 0) {
        if (fg >= 0) close(fg);
        if (fb >= 0) close(fb);
        return QK_ERR_IO;
    }
    const uint64_t n = (uint64_t)sg.st_size / 2;              /* the loop runs over the .qgc (Q.c:660) */
    int rc = (uint64_t)sb.st_size / 2 < n ? QK_ERR_IO : QK_OK; /* a .bin shorter than the .qgc: the reference reads stale buffers */
    uint64_t *lo = NULL, *hi = NULL, *win = NULL;
    double *sums = NULL, *vals = NULL;
    uint64_t printed = 0, P = 0;
    if (!rc && n && n_windows) {
        lo = malloc(n_windows * 8);
        hi = malloc(n_windows * 8);
        sums = malloc(n_windows * 8);
        if (!lo || !hi || !sums) rc = QK_ERR_NOMEM;
    }
    if (!rc && n && n_windows) {
        /* which k-mers each window really accumulates, and whether it is printed (Q.c:664-680): window w becomes
         * current at index q (0 for the first; the index its predecessor was printed at otherwise), accumulates
         * max(left, q) .. right-1, and is printed at the first index >= right that comes after it became current */
        uint64_t q = 0;
        for (uint64_t w = 0; w < n_windows; ++w) {
            const uint64_t at = w == 0 ? right[0] : (q + 1 > right[w] ? q + 1 : right[w]);
            if (at > n - 1) break;                             /* the files end first: not printed */
            lo[w] = left[w] > q ? left[w] : q;
            hi[w] = right[w];
            if (lo[w] > hi[w]) lo[w] = hi[w];
            q = P = at;
            ++printed;
        }
    }
    if (!rc && printed) {
        rc = qk_est_begin(ctx, n);
        int kind = 1;
        if (!rc) rc = qk_ingest_elements(ctx, fg, 0, n, 2, qk_reader_threads_default(), est_piece, &kind);
        kind = 0;
        if (!rc) rc = qk_ingest_elements(ctx, fb, 0, n, 2, qk_reader_threads_default(), est_piece, &kind);
        if (!rc) rc = qk_est_windows(ctx, correction, lo, hi, printed, sums);
        qk_est_end(ctx);
    }
    close(fg);
    close(fb);
    if (!rc && printed) {
        /* after the last window of the list, one more line per remaining block (see the header) */
        uint64_t repeats = 0;
        if (printed == n_windows) {
            const uint64_t blocks = ((uint64_t)sg.st_size + 1024 * 1024 - 1) / (1024 * 1024);
            repeats = blocks - 1 - P / EST_BLOCK_ENTRIES;
        }
        vals = malloc((printed + repeats) * 8);
        win = malloc((printed + repeats) * 8);
        if (!vals || !win) rc = QK_ERR_NOMEM;
        else {
            for (uint64_t w = 0; w < printed; ++w) {
                double v = sums[w];
                v /= right[w] - left[w];                       /* Q.c:668: uint32 difference */
                v /= mean_depth / 2;                           /* Q.c:669 */
                vals[w] = v;
                win[w] = w;
            }
            for (uint64_t r = 0; r < repeats; ++r) {
                double v = vals[printed + r - 1];
                v /= right[n_windows - 1] - left[n_windows - 1];
                v /= mean_depth / 2;
                vals[printed + r] = v;
                win[printed + r] = n_windows - 1;
            }
            *n_lines = printed + repeats;
        }
    }
    free(lo); free(hi); free(sums);
    if (rc) { free(vals); free(win); return rc; }
    *values_out = vals;
    *line_window_out = win;
    return QK_OK;
}

int qk_est_main(int argc, char **argv)
{
    char path[65536];
    if (argc < 4) { help_est(); return 1; }                   /* (the reference checks argc < 2 and then reads argv[argc-3]) */
    const char *ref = argv[argc - 3], *sample = argv[argc - 2], *out = argv[argc - 1];
    snprintf(path, sizeof path, "%s.qgc", ref);
    char qgc_path[65536], bin_path[65536];
    snprintf(qgc_path, sizeof qgc_path, "%s.qgc", ref);
    if (access(qgc_path, R_OK) != 0) { puts("GC control file missing."); help_est(); return 1; }
    snprintf(path, sizeof path, "%s.bed", ref);
    FILE *wf = fopen(path, "r");
    if (!wf) { puts("Window segmentation file missing."); help_est(); return 1; }
    snprintf(bin_path, sizeof bin_path, "%s.bin", sample);
    if (access(bin_path, R_OK) != 0) { puts("Sample file missing (Use prefix)."); help_est(); fclose(wf); return 1; }
    snprintf(path, sizeof path, "%s.txt", sample);
    FILE *gc = fopen(path, "r");
    if (!gc) {
        /* Q.c:597-624 would rebuild the curve here, but shadows its FILE* and then fclose()s NULL: the
         * reference dies at this point.  `count` always writes the .txt when the .qgc exists. */
        puts("Depth control not found. Regenerating...");
        printf("%s is missing: run count again (it writes the .txt whenever %s exists)\n", path, qgc_path);
        fclose(wf);
        return 1;
    }
    double total_depth = 0;                                   /* Q.c:626-638 */
    uint64_t total_count = 0;
    float percent, depth;
    uint32_t cur_count;
    char word[255];
    while (fscanf(gc, "%f\t%f\t%i\t%254s\n", &percent, &depth, &cur_count, word) == 4) {
        total_depth += depth * cur_count;                     /* float * uint32_t, as there */
        total_count += cur_count;
    }
    fclose(gc);
    total_depth /= total_count;
    printf("Mean sequencing depth: %.2f\n", total_depth);
    fflush(stdout);
    char cmd[65536 + 64];                                     /* Q.c:642-650: the LOWESS curve from the Python helper */
    snprintf(cmd, sizeof cmd, "smooth_GC_mrsfast.py %s.txt", sample);
    FILE *pipe_in = popen(cmd, "r");
    float correction[QK_GC_BINS];
    memset(correction, 0, sizeof correction);
    size_t got = pipe_in ? fread(correction, 4, QK_GC_BINS, pipe_in) : 0;
    if (pipe_in) pclose(pipe_in);
    if (got != QK_GC_BINS) fprintf(stderr, "quicKmer2_b200: smooth_GC_mrsfast.py gave %zu of 401 values (is it on the PATH?)\n", got);

    size_t cap = 1024, nw = 0;
    est_window *w = malloc(cap * sizeof *w);
    while (w) {                                               /* the whole window list (Q.c:657,673 read it line by line) */
        if (nw == cap) {
            est_window *g = realloc(w, (cap *= 2) * sizeof *w);
            if (!g) { free(w); w = NULL; break; }
            w = g;
        }
        if (fscanf(wf, "%63s\t%63s\t%63s\t%u\t%u\n", w[nw].chrom, w[nw].begin, w[nw].end, &w[nw].left, &w[nw].right) != 5) break;
        ++nw;
    }
    fclose(wf);
    if (!w) { puts("Memory allocation failed"); return 1; }
    uint32_t *left = malloc((nw ? nw : 1) * 4), *right = malloc((nw ? nw : 1) * 4);
    if (!left || !right) { puts("Memory allocation failed"); return 1; }
    for (size_t i = 0; i < nw; ++i) { left[i] = w[i].left; right[i] = w[i].right; }

    FILE *of = fopen(out, "w");
    if (!of) { printf("Cannot write %s\n", out); return 1; }
    qk_ctx *ctx = NULL;
    int rc = qk_ctx_create(&ctx, 0, 8, (size_t)32 << 20);
    double *vals = NULL;
    uint64_t *line_window = NULL, n_lines = 0;
    if (!rc) rc = qk_est_reduce(ctx, qgc_path, bin_path, correction, total_depth, left, right, nw, &vals, &line_window, &n_lines);
    if (rc) printf("Window reduction failed: %s\n", ctx ? qk_last_error(ctx) : "no CUDA device");
    for (uint64_t i = 0; !rc && i < n_lines; ++i) {
        const est_window *x = &w[line_window[i]];
        fprintf(of, "%s\t%s\t%s\t%f\n", x->chrom, x->begin, x->end, vals[i]);   /* Q.c:670 */
    }
    fclose(of);
    qk_ctx_destroy(ctx);
    free(vals); free(line_window); free(left); free(right); free(w);
    return rc ? 1 : 0;
}
