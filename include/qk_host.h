/*
 * qk_host.h -- host side of the `count` path, in C like the reference: the QM11 reader,
 * the FASTA/FASTQ record framer, the .bin/.txt writers and the `count` command itself.
 * Everything here runs on the CPU and feeds / drains the device through
 * quickmer2_b200.h; none of it computes k-mers, hashes or counts.
 */
#ifndef QK_HOST_H
#define QK_HOST_H

#include <stddef.h>
#include <stdint.h>
#include <sys/types.h>

#include "quickmer2_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define QK_ERR_IO 6 /* file cannot be opened / short read */
#define QK_HOST_MAX_SLOTS 16

/* ---- QM11 dictionary file: written at Q.c:1284-1299, read at Q.c:345-359 and 483 ---- */
typedef struct qk_qm_header {
    uint8_t k;          /* byte 4 */
    uint64_t hash_size; /* bytes 8..15, power of two */
    uint64_t first_idx; /* bytes 16..23 */
} qk_qm_header;

int qk_qm_read_header(const char *qm_path, qk_qm_header *hdr);
/* Stream <qm_path> to the device in pinned pieces and build the table.  Unlike the
 * reference (NULL dereference at Q.c:345) a missing file is an error, not a crash. */
int qk_qm_load(qk_ctx *ctx, const char *qm_path, qk_qm_header *hdr_out, uint64_t *n_kmers_out);

/* ---- record framer: Q.c:393-398 and 451-455 -------------------------------------------
 * Splits a FASTA / 4-line FASTQ byte stream into sequence lines exactly as the
 * reference's fgets loop does:
 *   - the first line decides the format: '@' => FASTQ (that line is consumed); otherwise
 *     FASTA and the stream is rewound -- which silently fails on a pipe, so on a
 *     non-seekable input the first line is lost (Q.c:396);
 *   - a line starting with '>' is skipped, in either format (Q.c:398);
 *   - every other line is one read, '\n' included; multi-line FASTA records are
 *     independent reads;
 *   - in FASTQ mode the three lines after a read are skipped (Q.c:451-455).
 * Outside the reference's defined behaviour, handled deterministically and counted in the
 * stats: a last line without '\n' gets one appended (T9), a line longer than 99,999
 * bytes is passed through whole (T8).
 */
typedef struct qk_framer qk_framer;
typedef struct qk_framer_stats {
    uint64_t lines;         /* sequence lines emitted            */
    uint64_t bases;         /* their bytes, excluding the '\n'   */
    uint64_t raw_bytes;     /* bytes consumed from the input     */
    uint64_t long_lines;    /* > QK_MAX_LINE_BYTES (T8)          */
    uint64_t unterminated;  /* final line without '\n' (T9)      */
    int fastq;
    uint64_t sink_bytes;    /* qk_frame_mem_mt / qk_count_{mem,file}_mt: bytes handed to the chunk consumer, i.e. shipped to
                             * the GPUs -- the sequence lines, or 0.375 bytes per position when the chunks are packed */
} qk_framer_stats;

qk_framer *qk_framer_open(const char *path);        /* NULL if the file cannot be opened */
qk_framer *qk_framer_open_fd(int fd, int seekable); /* takes ownership of fd             */
/* In-memory input (tests, benchmarks): frames `n` bytes at `data`; seekable says which
 * first-line rule applies. The memory must outlive the framer. */
qk_framer *qk_framer_open_mem(const uint8_t *data, size_t n, int seekable);
/* Fill `dst` (capacity `cap` >= 100000) with whole sequence lines.  If line_off != NULL
 * it receives the start offset of every line plus the end offset (at most off_cap
 * entries; filling stops before it would overflow).  Returns 1 if *n_bytes > 0 was
 * produced, 0 at end of input, negative QK_ERR_* on error. */
int qk_framer_next(qk_framer *f, uint8_t *dst, size_t cap, size_t *n_bytes, uint32_t *line_off, uint32_t off_cap,
                   uint32_t *n_lines);
void qk_framer_get_stats(const qk_framer *f, qk_framer_stats *st);
void qk_framer_close(qk_framer *f);

/* ---- byte streams: plain or gzip, regular file or pipe -----------------------------------
 * Everything that reads a reads stream sequentially goes through these: the gzip magic is
 * recognised on files and pipes alike and the data inflated on the fly (concatenated members
 * included); a BAM container is recognised after inflation and turned into the text of its reads.  The reference itself reads plain text only (its documented route for compressed
 * input is a pipe, README.md:89-90). */
typedef struct qk_stream qk_stream;
qk_stream *qk_stream_open(const char *path);        /* NULL if it cannot be opened */
qk_stream *qk_stream_open_fd(int fd, int seekable); /* takes ownership of fd       */
/* Up to `cap` bytes; short only at the end of the stream; 0 = end, -1 = I/O or format error. */
ssize_t qk_stream_read(qk_stream *s, uint8_t *dst, size_t cap);
int qk_stream_is_gzip(const qk_stream *s);
/* 1 once the first read has found a BAM container (BGZF that inflates to "BAM\1"): qk_stream_read then delivers
 * the FASTA text `samtools view -F 3840 | awk '{print ">\n"$10}'` would (README.md:89-90) -- no samtools, no pipe. */
int qk_stream_is_bam(const qk_stream *s);
int qk_stream_seekable(const qk_stream *s);
void qk_stream_close(qk_stream *s);

/* ---- writers: Q.c:498-518 (.bin) and Q.c:522-542 (.txt) -------------------------------- */
int qk_write_bin(const char *path, const uint16_t *counts, uint64_t n);
/* Device -> .bin without a host copy of the whole array (qk_finish_pieces + fwrite). */
int qk_write_bin_from_device(qk_ctx *ctx, const char *path);
/* The GC control curve (Q.c:495-509) straight from the .qgc FILE: streamed through the pinned slots to the device
 * by the reader threads (4.5 GB at human scale).  *entries_read < n_kmers = the file is short (the rest count as
 * non-control); *bins_out_of_range = entries whose GC bin is above 400 (ignored). */
int qk_gc_curve_file(qk_ctx *ctx, const char *qgc_path, uint64_t n_kmers, uint64_t sum[QK_GC_BINS], int64_t sumsq[QK_GC_BINS],
                     uint64_t count[QK_GC_BINS], uint64_t *entries_read, uint64_t *bins_out_of_range);
/* 401 lines "%.2f\t%f\t%i\t%f\n" of bin/4, mean, count, variance; *mean_depth receives the
 * figure printed as "Mean sequencing depth" (Q.c:539-540). */
int qk_write_gc_txt(const char *path, const uint64_t sum[QK_GC_BINS], const int64_t sumsq[QK_GC_BINS],
                    const uint64_t count[QK_GC_BINS], double *mean_depth);

/* ---- streaming driver: file -> framer -> pinned slots -> device -------------------------
 * Counts every read of `reads_path` into the context's counters (does not reset them).
 * Returns the framer statistics. */
int qk_count_file(qk_ctx *ctx, const char *reads_path, qk_framer_stats *st);
int qk_count_framer(qk_ctx *ctx, qk_framer *f, qk_framer_stats *st);

/* Same result with the framing done on the device (qk_raw_begin / qk_submit_raw): the host
 * only cuts the stream at line ends.  `data` may be pinned (then chunks are DMA'd straight
 * from it) or pageable (then they pass through the slots' pinned buffers). */
int qk_count_raw_mem(qk_ctx *ctx, const uint8_t *data, size_t n, int seekable, qk_framer_stats *st);
int qk_count_raw_fd(qk_ctx *ctx, int fd, int seekable, qk_framer_stats *st); /* does not close fd */
int qk_count_raw_stream(qk_ctx *ctx, qk_stream *in, qk_framer_stats *st);    /* plain or gzip */
int qk_count_raw_file(qk_ctx *ctx, const char *reads_path, qk_framer_stats *st);
/* Regular files: `threads` readers (0 = QK_READER_THREADS or 4, at most the slot count)
 * pread() pieces into the pinned slot buffers in parallel; lines longer than 128 KiB are an
 * error on this path.  Pipes fall back to qk_count_raw_fd. */
int qk_count_raw_file_mt(qk_ctx *ctx, const char *reads_path, uint32_t threads, qk_framer_stats *st);

/* ---- all host cores frame, any GPU counts (host/qk_framer_mt.c) ---------------------------------
 * The framing rules above applied by `threads` workers (0 = QK_FRAMER_THREADS or every online
 * CPU) to blocks of the input in parallel; only the sequence lines are copied into the pinned
 * slot buffers, so FASTQ crosses the host link at ~1.25 instead of ~2.6 bytes per k-mer -- and at
 * ~0.47 when they go as PACKED chunks (2-bit codes + reset flags, see qk_chunk_sink.packed and
 * qk_submit_packed), which they do when every context counts with the dictionary-order kernel
 * (3 <= k <= 31) unless QK_PACKED=0 is set.  The framed chunks go to whichever of the n_ctx contexts (same slot count and capacity; one per
 * GPU; same dictionary) has a slot free first.  Same result, byte for byte, as qk_count_framer /
 * qk_count_raw_mem.  qk_count_file_mt maps a regular file (the page cache is the input buffer);
 * pipes and gzip files are counted by ctxs[0] through the sequential stream path. */
int qk_count_mem_mt(qk_ctx *const *ctxs, uint32_t n_ctx, const uint8_t *data, size_t n, int seekable, uint32_t threads,
                    qk_framer_stats *st);
/* The framer alone, with the consumer of the framed chunks as a parameter: n_ctx consumers with
 * n_slots buffers of `cap` bytes each.  The framer fills buffer(c, s) -- after ready(c, s)
 * returned 1 or wait(c, s) returned 0 for a buffer it has submitted before -- and hands it over
 * with submit(c, s, seq, n_bytes, n_lines), seq = ordinal of the chunk in the framed stream; submit
 * is called from worker threads, one call at a time per consumer c. */
typedef struct qk_chunk_sink {
    void *user;
    uint32_t n_ctx, n_slots;
    size_t cap;
    uint8_t *(*buffer)(void *user, uint32_t c, uint32_t s);
    int (*ready)(void *user, uint32_t c, uint32_t s);
    int (*wait)(void *user, uint32_t c, uint32_t s);
    int (*submit)(void *user, uint32_t c, uint32_t s, uint64_t seq, size_t n_bytes, uint32_t n_lines);
    /* != 0: the consumer takes PACKED chunks -- per 64 positions of the framed stream 24 bytes: four little-endian 32-bit
     * words of 2-bit codes ((c >> 1) & 3, Q.c:411; 16 positions per word, the first in the top pair) and 64 flags
     * (bit p: position p is 'N' or '\n', Q.c:403-404), which is all the count kernels keep of a byte.  `cap` and
     * submit's n_bytes then count POSITIONS (multiples of 64; a block's lines are followed by '\n' positions up to
     * the next multiple), the buffer holds n_bytes / 64 * 24 bytes. */
    int packed;
} qk_chunk_sink;
int qk_frame_mem_mt(const qk_chunk_sink *sink, const uint8_t *data, size_t n, int seekable, uint32_t threads, qk_framer_stats *st);
/* Measurement: the framer alone over `data`, chunks discarded; GB/s of raw input consumed and of
 * framed output produced (the host-side term of the end-to-end roofline). */
int qk_bench_framer(const uint8_t *data, size_t n, uint32_t threads, int repeats, double *raw_gbs, double *framed_gbs);
/* For scale: what the host's memory gives `threads` threads -- GB/s of a pure read and of memcpy (bytes copied). */
int qk_bench_host_memory(size_t bytes_per_thread, uint32_t threads, double *read_gbs, double *copy_gbs);
int qk_count_file_mt(qk_ctx *const *ctxs, uint32_t n_ctx, const char *reads_path, uint32_t threads, qk_framer_stats *st);

/* ---- one reads file, several GPUs ---------------------------------------------------------
 * qk_shard_bounds: the line-aligned byte range [begin, end) of shard `rank` of `world`
 * (regular files only).  qk_count_raw_range counts that range on one context, starting the
 * reference's line state machine in `line_state` (see qk_raw_begin_state) and returning the
 * state after the last line.  FASTA shards always start in state 0; for FASTQ the state of a
 * shard is whatever the previous shard ends in: qk_fastq_state_guess proposes it from the
 * first lines of the shard so that all shards can run at once, and the caller compares each
 * shard's assumed state with its predecessor's final state afterwards (quick-mer2_b200/dist.py). */
int qk_shard_bounds(const char *reads_path, uint32_t rank, uint32_t world, uint64_t *begin, uint64_t *end);
int qk_fastq_state_guess(const uint8_t *window, size_t n, uint32_t *line_state);
int qk_count_raw_range(qk_ctx *ctx, const char *reads_path, uint64_t begin, uint64_t end, int fastq, uint32_t line_state,
                       qk_framer_stats *st, uint32_t *final_state);
int qk_count_raw_range_mt(qk_ctx *ctx, const char *reads_path, uint64_t begin, uint64_t end, int fastq, uint32_t line_state,
                          uint32_t threads, qk_framer_stats *st, uint32_t *final_state);
/* The whole file over the GPUs of a qk_multi, one shard and `threads_per_gpu` readers per GPU,
 * with the guess / verify / recount loop described above done here in C.  A pipe is counted by
 * context 0 alone.  Call qk_multi_reduce afterwards. */
int qk_count_file_multi(qk_multi *m, const char *reads_path, uint32_t threads_per_gpu, qk_framer_stats *st);

/* ---- est: main_estimate, Q.c:555-685, with the window reduction on the device -----------------------
 * quicKmer2 est ref.fa sample_prefix output.bed -- same files, same stdout, same bytes in output.bed; the LOWESS
 * curve still comes from `smooth_GC_mrsfast.py <sample>.txt` on the PATH (Q.c:642-650).  qk_est_reduce is the
 * reduction alone for a caller that has the curve: one value per line the reference would print (values_out and
 * line_window_out are malloc'd, n_lines entries each; line i belongs to window line_window_out[i]). */
int qk_est_main(int argc, char **argv);
int qk_est_reduce(qk_ctx *ctx, const char *qgc_path, const char *bin_path, const float correction[QK_GC_BINS], double mean_depth,
                  const uint32_t *left, const uint32_t *right, uint64_t n_windows, double **values_out, uint64_t **line_window_out,
                  uint64_t *n_lines);

/* ---- the command: main_count, Q.c:304-545 -----------------------------------------------
 * quicKmer2 count [-h] [-t N] [-g device[,device...]] ref_prefix reads out_prefix
 * Same positional-from-the-end convention, same stdout lines, same files. */
int qk_count_main(int argc, char **argv);

#ifdef __cplusplus
}
#endif
#endif /* QK_HOST_H */
