/*
 * quickmer2_b200.h -- C ABI of the B200 device side of `quicKmer2 count`.
 *
 * The reference (QuicKmer.c, "Q.c") is a monolith with no FFI; the seams it does have are
 * listed in SURVEY.md 8(b).  This header is the boundary a maintainer of the reference
 * would bind to: each entry point names the reference lines it replaces.  Plain C types
 * only.  All functions return 0 on success or a QK_ERR_* code; qk_last_error() returns a
 * human-readable message for the most recent failure on that context.
 *
 * There is NO CPU fallback behind any of these calls: without a CUDA device
 * qk_ctx_create() fails with QK_ERR_CUDA.
 *
 * Threading: one submitting thread per context (the reference has exactly one producer,
 * Q.c:397-456); qk_wait_slot and qk_slot_host_buffer may be called from other threads (the
 * host's reader threads fill slot buffers that way).  One context drives one GPU.

 */
#ifndef QUICKMER2_B200_H
#define QUICKMER2_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QK_OK 0
#define QK_ERR_CUDA 1   /* CUDA runtime failure (incl. no device)                  */
#define QK_ERR_ARG 2    /* bad argument                                            */
#define QK_ERR_NOMEM 3  /* host or device allocation failed (Q.c:354-366 return 1) */
#define QK_ERR_STATE 4  /* call out of order                                       */
#define QK_ERR_FORMAT 5 /* the chain does not come back to first_idx                */

#define QK_GC_BINS 401          /* Q.c:495-497 */
#define QK_MAX_LINE_BYTES 99999 /* Q.c:388,397: fgets(line, 100000) incl. the '\n' */

typedef struct qk_ctx qk_ctx;

/* Version / build string of the library (static storage). */
const char *qk_version(void);

/* Number of CUDA devices visible, or a negative QK_ERR_* code. */
int qk_device_count(void);

/*
 * Create a context on CUDA device `device` with `n_slots` (1..16) chunk slots of
 * `chunk_capacity` bytes each.  Every slot owns a pinned host buffer, a device buffer, a
 * stream and two events -- the replacement for the reference's per-worker double FIFO
 * (struct FIFO_arg_struc, Q.c:34-41; thread pool set-up Q.c:368-384).
 * chunk_capacity must be >= 2 * 100000 and < 2^31.
 */
int qk_ctx_create(qk_ctx **out, int device, uint32_t n_slots, size_t chunk_capacity);
void qk_ctx_destroy(qk_ctx *ctx);
const char *qk_last_error(const qk_ctx *ctx);
/* Slot count and (tile-rounded) slot capacity of a context. */
int qk_ctx_info(const qk_ctx *ctx, uint32_t *n_slots, size_t *chunk_capacity);

/* ------------------------------------------------------------------ dictionary ------
 * Replaces the .qm load (Q.c:345-359 keys, Q.c:483 chain) and, at build time, the
 * reference's serial chain walk (Q.c:490-516): the chain is list-ranked on the device so
 * that every dictionary k-mer gets its ordinal (its index in the .bin file), and a new
 * bucketised table key -> ordinal is built.  Observable semantics kept: membership in the
 * set of non-zero keys, and for duplicate keys (written by `index`, Q.c:209-216) only the
 * slot Find_hash (Q.c:90-99) reaches first ever receives counts.
 */
int qk_dict_begin(qk_ctx *ctx, uint8_t k, uint64_t hash_size, uint64_t first_idx);
/* Copy `count` keys / chain entries starting at hash slot `slot_offset`.  `keys`/`next`
 * are host pointers (pinned or pageable); the call returns when they may be reused. */
int qk_dict_upload_keys(qk_ctx *ctx, uint64_t slot_offset, const uint64_t *keys, uint64_t count);
int qk_dict_upload_chain(qk_ctx *ctx, uint64_t slot_offset, const uint32_t *next, uint64_t count);
/* The same from the pinned buffer of `slot` (qk_slot_host_buffer), asynchronously on the
 * slot's stream; kind 0 = keys (u64), 1 = chain (u32); qk_wait_slot(slot) tells when the buffer
 * may be refilled.  Lets several reader threads fill slots while copies are in flight. */
int qk_dict_upload_from_slot(qk_ctx *ctx, uint32_t slot, int kind, uint64_t elem_offset, uint64_t count);
/* Rank the chain, build the table, release the raw arrays, zero the counters. */
int qk_dict_build(qk_ctx *ctx, uint64_t *n_kmers_out);

/* Geometry of a built table -- what a peer GPU needs to hold a replica. */
typedef struct qk_table_desc {
    uint64_t n_kmers;        /* chain length = number of .bin entries            */
    uint64_t n_buckets;      /* power of two                                     */
    uint64_t stash_slots;    /* power of two                                     */
    uint64_t stash_used;
    uint64_t table_bytes;    /* n_buckets * 32                                   */
    uint64_t stash_bytes;    /* stash_slots * 16                                 */
    uint32_t k;
    uint32_t bucket_bits;    /* log2(n_buckets)                                  */
    uint32_t ord_bits;       /* bits of the (ordinal + 1) field of an entry      */
    uint32_t rem_bits;       /* 60 - bucket_bits                                 */
    uint64_t skipped_keys;   /* dictionary keys no read can produce (>= 2^60) or
                                shadowed duplicates; their ordinals stay 0       */
    uint64_t ext_bytes;      /* bytes of the dictionary-order extension array: 12 bytes per 16 ordinals */
    uint64_t cont_bytes;     /* 0 (the continuation bits live in the same array)  */
    uint32_t has_ext;        /* dictionary-order extension array: 0 none, 1 canonical 30-mers (k = 30),
                                2 forward k-mers (k < 30), 3 30-base reverse complements (k = 31) */
    uint32_t reserved;
} qk_table_desc;

int qk_dict_describe(const qk_ctx *ctx, qk_table_desc *desc);
/* The geometry qk_dict_build will choose for n_kmers chain entries (no device needed). */
int qk_table_geometry(uint64_t n_kmers, uint32_t k, qk_table_desc *desc);
/* Allocate an (uninitialised) replica with the geometry of `desc` on this context; the
 * host then fills it, e.g. with an NCCL broadcast into the pointers below. */
int qk_dict_adopt(qk_ctx *ctx, const qk_table_desc *desc);
/* Device pointers of the table image (for NCCL broadcast / peer copies). */
int qk_dict_device_ptrs(const qk_ctx *ctx, void **table, void **stash);
/* ... and of the extension array (NULL when has_ext == 0) through *last: per 16 ordinals three words --
 * last bases and first bases of the dictionary k-mers in their walking orientation (2 bits each) and
 * continuation bits (ext_bytes in all).  *first and *cont come back NULL. */
int qk_dict_ext_ptrs(const qk_ctx *ctx, void **last, void **first, void **cont);

/* ------------------------------------------------------------------ counting --------
 * A chunk is a run of SEQUENCE LINES ONLY, each terminated by '\n', each at most
 * QK_MAX_LINE_BYTES bytes including the '\n' (the framing of Q.c:393-398,451-455 is the
 * host's job, see qk_host.h).  The device applies the codec of Q.c:399-420 to every line
 * -- 'N' resets, (c>>1)&3 encoding of every other byte, 64-bit forward and 60-bit
 * reverse-complement registers, 16-bit run counter with wrap -- looks every emitted key
 * up and adds 1 to the counter of its ordinal (Q.c:256-296, 440-444).
 */
/* Pinned host buffer of slot `slot` (chunk_capacity bytes). */
uint8_t *qk_slot_host_buffer(qk_ctx *ctx, uint32_t slot);
/* Enqueue a chunk: async H2D of `n_bytes` from `bytes` (any host pointer; the slot's own
 * pinned buffer gives a true async copy) followed by the count kernel, on the slot's
 * stream.  `line_off` (n_lines + 1 offsets, may be NULL) is not needed by the device and
 * only cross-checked in debug builds; n_lines feeds the statistics.  Returns at once. */
int qk_submit(qk_ctx *ctx, uint32_t slot, const uint8_t *bytes, size_t n_bytes,
              const uint32_t *line_off, uint32_t n_lines);
/* Same, for a chunk already resident in device memory (16-byte aligned). */
int qk_submit_device(qk_ctx *ctx, uint32_t slot, const uint8_t *dev_bytes, size_t n_bytes);
/* A PACKED chunk from the slot's pinned buffer (or any host memory): per 64 positions of a framed stream 24 bytes --
 * four little-endian 32-bit words of 2-bit codes ((c >> 1) & 3, Q.c:411; 16 positions per word, the first in the
 * top pair) and 64 reset flags (bit p: position p is 'N' or '\n', Q.c:403-404).  That is all the count kernels
 * keep of a byte, so the result is the one qk_submit gives for the text; the link carries 0.375 bytes per position.
 * n_positions is a multiple of 64 (fill up with '\n' positions) and at most the slot capacity.  Needs the
 * dictionary-order kernel (table_desc.has_ext != 0, i.e. 3 <= k <= 31), QK_ERR_STATE otherwise.  The host framer
 * makes such chunks (host/qk_framer_mt.c). */
int qk_submit_packed(qk_ctx *ctx, uint32_t slot, const uint8_t *packed, size_t n_positions, uint32_t n_lines);
/* Raw streams: the record framing of Q.c:393-398,451-455 done ON THE DEVICE.  qk_raw_begin
 * starts a stream: `fastq` = its first byte is '@' (Q.c:395); `skip_first_line` = the first line
 * is consumed without being examined (always for FASTQ, and for FASTA on a pipe where the
 * reference's fseek fails, Q.c:396).  qk_submit_raw then takes consecutive pieces of the
 * stream, each made of WHOLE lines (last byte '\n') of any kind -- headers, reads, '+',
 * qualities; H2D copy, framing passes and the count kernel are enqueued on the slot's stream
 * and the call returns at once.  Pieces must be submitted in stream order; the line-phase
 * state is carried from piece to piece on the device. */
int qk_raw_begin(qk_ctx *ctx, int fastq, int skip_first_line);
/* The same with an explicit line state: the number (0..3) of lines still to be discarded
 * before a line is examined again -- for a shard that starts in the middle of a stream.
 * qk_raw_state returns the state after everything submitted so far (syncs). */
int qk_raw_begin_state(qk_ctx *ctx, int fastq, uint32_t line_state);
int qk_raw_state(qk_ctx *ctx, uint32_t *line_state);
int qk_submit_raw(qk_ctx *ctx, uint32_t slot, const uint8_t *bytes, size_t n_bytes);
/* Totals of the raw stream so far (syncs): sequence lines, their bases, all lines seen. */
int qk_raw_stats(qk_ctx *ctx, uint64_t *read_lines, uint64_t *bases, uint64_t *raw_lines);
/* 1 if `p` is page-locked host memory known to CUDA (async copies from it are true DMA). */
int qk_host_is_pinned(const void *p);
/* The cudaStream_t of a slot (as void *), so that a caller can order its own device work --
 * e.g. an NCCL reduce of the counters -- against the slot's work without a host sync. */
void *qk_slot_stream(qk_ctx *ctx, uint32_t slot);
/* Block until the host buffer last submitted on `slot` may be overwritten. */
int qk_wait_slot(qk_ctx *ctx, uint32_t slot);
/* The same without blocking: 1 = may be overwritten, 0 = its copy is still in flight, < 0 = -QK_ERR_*. */
int qk_slot_ready(qk_ctx *ctx, uint32_t slot);
/* Block until every enqueued chunk has been counted. */
int qk_sync(qk_ctx *ctx);

/* Running totals (call after qk_sync): emitted k-mers = the reference's process_kmers
 * ("total %lu kmers", Q.c:445,481), dictionary hits, and sequence lines seen. */
int qk_stats(qk_ctx *ctx, uint64_t *total_kmers, uint64_t *hits, uint64_t *lines);
/* How many of the hits were derived from a neighbouring k-mer through the dictionary-order
 * extension arrays instead of a table probe (0 when the dictionary has none, k != 30). */
int qk_stats_ext(qk_ctx *ctx, uint64_t *verified_by_extension);
/* What the count kernels actually asked of memory (for the roofline): 32-byte bucket sectors
 * loaded from the table (one per emitted k-mer without the extension arrays; anchors + positions
 * the walk could not settle with them), and dictionary-order walks started (each reads four words
 * of the extension arrays). */
int qk_stats_probes(qk_ctx *ctx, uint64_t *bucket_probes, uint64_t *walks);

/* Device pointer of the per-ordinal uint32 counters (n_kmers entries), for an NCCL
 * reduce across GPUs; the low 16 bits are the reference's uint16 depth (Q.c:23,291). */
int qk_counters_device_ptr(const qk_ctx *ctx, uint32_t **counters, uint64_t *n_kmers);
int qk_reset_counters(qk_ctx *ctx);
/* Two counter buffers (0 = default, 1 = allocated on first use): later launches, resets,
 * qk_counters_device_ptr, qk_finish ... use the selected one, so the reduce / download of one
 * job can overlap the counting of the next. */
int qk_counters_select(qk_ctx *ctx, uint32_t which);
/* The same, stream-ordered (no host synchronisation): for jobs run back to back. */
int qk_reset_counters_async(qk_ctx *ctx);
/* n more occurrences of the canonical key `key`, as if the reads had held them (nothing happens if
 * the dictionary does not hold it).  Replaces Q.c:458-466 + 284-291: with -t N the reference pads its
 * last batch of 4,096 keys with zeros and looks them up like any key, which is observable when the
 * empty slot Find_hash(0) stops at is on the chain (Q.c:98) -- the command calls this with key 0 and
 * 4,096 - total_kmers % 4,096 when it is given -t N, N > 0.  Syncs. */
int qk_add_depth(qk_ctx *ctx, uint64_t key, uint32_t n);
/* Copy `count` raw uint32 counters starting at ordinal `offset` to host memory (syncs). */
int qk_counters_download(qk_ctx *ctx, uint64_t offset, uint32_t *out, uint64_t count);

/* ------------------------------------------------------------------ results ---------
 * Replaces the dump loop Q.c:498-518: counts_out[i] = depth of the i-th k-mer of the
 * chain, already in .bin order, wrapped to 16 bits exactly as uint16_t Kmer_depth does.
 */
int qk_finish(qk_ctx *ctx, uint16_t *counts_out, uint64_t n_kmers);
/* The same without blocking: the download of the counter buffer selected at the time of the call is
 * enqueued after everything submitted so far; the caller then selects the other counter buffer
 * (qk_counters_select) and counts the next sample while this one's depths travel.  counts_out must be
 * page-locked; it is complete when qk_finish_wait returns.  The buffer being downloaded must not be reset
 * or counted into before that. */
int qk_finish_async(qk_ctx *ctx, uint16_t *counts_out, uint64_t n_kmers);
int qk_finish_wait(qk_ctx *ctx);
/* The same in pieces, in order, without a full-size host array: `consume` gets `count` depths
 * starting at .bin index `offset` (pinned memory, valid until it returns; non-zero aborts) while
 * the next piece is still on its way -- e.g. to write the .bin as it arrives (Q.c:510-513
 * flushes per 1 Mi entries). */
typedef int (*qk_piece_fn)(void *user, const uint16_t *piece, uint64_t offset, uint64_t count);
int qk_finish_pieces(qk_ctx *ctx, qk_piece_fn consume, void *user);
/* GC control curve sums of Q.c:501-508 from the final depths and the .qgc flags
 * (host array of n_kmers uint16): sum[b] = sum of depth, sumsq[b] = sum of the int
 * product depth*depth, count[b] = entries, for control k-mers of GC bin b. */
int qk_gc_curve(qk_ctx *ctx, const uint16_t *qgc, uint64_t n_kmers, uint64_t sum[QK_GC_BINS],
                int64_t sumsq[QK_GC_BINS], uint64_t count[QK_GC_BINS]);

/* The same from .qgc pieces in the slots' pinned buffers (qk_slot_host_buffer; count entries for the ordinals
 * starting at ordinal_offset), H2D + histogram per piece on the slot's stream: what qk_gc_curve_file (qk_host.h)
 * drives with its reader threads.  qk_wait_slot tells when a buffer may be refilled. */
int qk_gc_begin(qk_ctx *ctx);
int qk_gc_from_slot(qk_ctx *ctx, uint32_t slot, uint64_t ordinal_offset, uint64_t count);
int qk_gc_end(qk_ctx *ctx, uint64_t sum[QK_GC_BINS], int64_t sumsq[QK_GC_BINS], uint64_t count[QK_GC_BINS],
              uint64_t *bins_out_of_range);

/* ------------------------------------------------------------------ est: window depths -
 * Q.c:660-682 on the device.  qk_est_begin allocates two arrays of n_entries uint16 (a sample's depths,
 * the .qgc flags); qk_est_upload_from_slot fills them from the pinned slot buffers (kind 0 = depths,
 * 1 = flags); qk_est_windows returns, per window w, the sum over k-mers lo[w] .. hi[w]-1 of
 * correction[flags & 0x1FF] * depth -- float product, double accumulation, ordinal order: the reference's
 * arithmetic, so the results print to the same bytes.  No dictionary is needed on the context. */
int qk_est_begin(qk_ctx *ctx, uint64_t n_entries);
int qk_est_upload_from_slot(qk_ctx *ctx, uint32_t slot, int kind, uint64_t elem_offset, uint64_t count);
int qk_est_windows(qk_ctx *ctx, const float correction[QK_GC_BINS], const uint64_t *lo, const uint64_t *hi,
                   uint64_t n_windows, double *sums_out);
int qk_est_end(qk_ctx *ctx);

/* ------------------------------------------------------------------ several GPUs -----
 * One process, one context per device: the dictionary built on context 0 is
 * replicated with ncclBroadcast, every context counts its share of the reads, the u32 counters
 * are added into context 0 with ncclReduce; qk_finish / qk_gc_curve then run on context 0.
 * NCCL (libnccl.so.2) is bound at run time and only when n > 1.  For one process PER GPU
 * (torchrun) use qk_dict_describe / adopt / device_ptrs with the launcher's own collectives
 * instead (quick-mer2_b200/dist.py). */
typedef struct qk_multi qk_multi;
int qk_multi_create(qk_multi **out, const int *devices, uint32_t n, uint32_t n_slots, size_t chunk_capacity);
void qk_multi_destroy(qk_multi *m);
const char *qk_multi_last_error(const qk_multi *m);
uint32_t qk_multi_size(const qk_multi *m);
qk_ctx *qk_multi_ctx(qk_multi *m, uint32_t i);
int qk_multi_replicate(qk_multi *m); /* dictionary of context 0 -> all */
int qk_multi_reduce(qk_multi *m);    /* counters of all -> context 0 (syncs) */

/* ------------------------------------------------------------------ measurement -----
 * Device time (ms) spent in count kernels / H2D copies on all slots since the last
 * qk_reset_counters, from CUDA events recorded on the slots' streams, and the number of
 * count-kernel launches. */
int qk_timing(qk_ctx *ctx, double *kernel_ms, double *h2d_ms, uint64_t *launches);
/* Device-clock span over ALL slot streams: qk_span_begin records a start event ordered after
 * everything enqueued so far; qk_span_end records an end event ordered after everything
 * enqueued on every slot stream since, waits for it and returns the elapsed device time. */
int qk_span_begin(qk_ctx *ctx);
int qk_span_end(qk_ctx *ctx, double *elapsed_ms);
/* Micro-benchmarks for the roofline denominators: random `gran`-byte
 * (32, 64 or 128) gathers over a `table_bytes` region with `loads_in_flight` independent loads
 * per thread; and pinned host -> device copy.  Both return GB/s. */
int qk_bench_gather(qk_ctx *ctx, uint64_t table_bytes, uint32_t gran, uint32_t loads_in_flight,
                    uint64_t n_gathers, double *gbs);
int qk_bench_h2d(qk_ctx *ctx, size_t bytes, int repeats, double *gbs);

#ifdef __cplusplus
}
#endif
#endif /* QUICKMER2_B200_H */
