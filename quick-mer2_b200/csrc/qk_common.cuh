// qk_common.cuh -- shared device/host definitions: context layout, key mixer, entry layout.
//
// Table layout in HBM (DESIGN.md "data layout"):
//   bucket  = 4 x uint64 entries = one 32-byte sector, fetched with one LDG.E.256
//   entry   = (remainder << ord_bits) | (ordinal + 1), 0 = empty; rem_bits + ord_bits <= 62 and
//              ord_bits <= 32, so the probe compares the high words with one 32-bit op and the low
//              words with another; bit 63 = strand of the k-mer's walking orientation; bit 62 of the
//              LAST entry of a bucket = "a key of this bucket went to the stash" (a miss in a full
//              bucket needs the stash only then)
//   key     -> h = mix60(key) (a bijection on [0, 2^60)); bucket = top bucket_bits of h,
//              remainder = low rem_bits = 60 - bucket_bits of h (quotienting: the bucket
//              index is implied, so remainder + ordinal fit one 64-bit word)
//   stash   = open-addressed {key | 2^63, ordinal+1} 16-byte entries for the keys whose home
//              bucket was full at build time; consulted only when a bucket is full
// Every key a read can produce is < 2^60 because the reference's reverse-complement
// register is 60 bits wide for every k (Q.c:415-416,420), so 60-bit keys lose nothing.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/quickmer2_b200.h"

#define QK_TILE 4096          // bytes (= positions) per tile
#define QK_THREADS 256        // threads per CTA of the count kernel
#define QK_MAX_SLOTS 16
#define QK_TIMING_RING 64
#define QK_KEY_BITS 60
#define QK_BUCKET_ENTRIES 4
#define QK_FRAME_MAX_CTAS 2048   // CTAs of the device framing passes (per chunk)
#define QK_STATS_WORDS 8
#define QK_ENTRY_OVERFLOW 0x4000000000000000ull // bit 62 of a bucket's last entry
#define QK_EXT_GROUP_WORDS 3

struct __align__(32) qk_bucket { unsigned long long e[QK_BUCKET_ENTRIES]; };
struct __align__(16) qk_stash_entry { unsigned long long key; uint32_t ord1; uint32_t pad; };
#define QK_STASH_TAKEN 0x8000000000000000ull   // set in a used stash entry's key (keys are < 2^60, and 0 is a key)

// Everything the count kernel needs, passed by value.
struct qk_table_view {
    const qk_bucket *buckets;
    const qk_stash_entry *stash;
    uint64_t stash_mask;      // stash_slots - 1
    uint64_t kmask;           // (1 << 2k) - 1, 0 for k = 32 (Q.c:419 on x86-64)
    uint32_t k;
    uint32_t rem_bits;        // 60 - bucket_bits
    uint32_t ord_bits;
    uint32_t has_stash;       // stash_used != 0
    const uint32_t *ext;      // dictionary-order extension array (12 bytes per 16 ordinals), NULL when has_ext == 0 (k < 3, k = 32)
    uint64_t n_kmers;
};

struct qk_timing_pair { cudaEvent_t a, b; int kind; /* 0 = h2d, 1 = kernel */ };

struct qk_slot {
    uint8_t *host;            // pinned
    uint8_t *dev;
    cudaStream_t stream;
    cudaEvent_t h2d_done;
    cudaEvent_t frame_done;   // this slot's chunk has handed the framing state on
    qk_timing_pair ring[QK_TIMING_RING];
    uint32_t ring_head, ring_count;
};

struct qk_ctx {
    int device;
    int sm_count;
    uint32_t n_slots;
    size_t chunk_capacity;
    qk_slot slots[QK_MAX_SLOTS];
    char err[512];

    // raw QM11 arrays, resident only between qk_dict_begin and qk_dict_build
    uint64_t *raw_keys;
    uint32_t *raw_next;
    uint64_t hash_size, first_idx;
    uint8_t k;
    int dict_state;           // 0 none, 1 uploading, 2 built/adopted

    qk_bucket *buckets;
    qk_stash_entry *stash;
    uint32_t *ext;            // dictionary-order extension array (3 <= k <= 31), else NULL: per 16 ordinals three words --
                              // last base (2 bits each), first base (2 bits each), continuation bits (low 16)
    qk_table_desc desc;

    uint32_t *counters;       // n_kmers x u32, indexed by ordinal: the buffer jobs currently count into
    uint32_t *counters_buf[2]; // [0] always allocated with the table; [1] on first qk_counters_select(1)
    unsigned long long *stats; // device: [0] emitted k-mers, [1] hits, [2] hits derived by the dictionary-order walk, [3] micro-benchmark sink,
                               // [4] bucket probes issued, [5] stash probes, [6] walks started (anchors that hit)
    uint64_t lines;

    double kernel_ms, h2d_ms;
    uint64_t launches;
    cudaEvent_t span_a, span_b, span_join;
    uint16_t *narrow_dev, *narrow_host; // qk_finish staging (device / pinned), allocated on first use
    cudaStream_t finish_stream;         // qk_finish_async: the result download runs here, beside the slot streams
    cudaEvent_t finish_done;
    int finish_pending;
    uint16_t *est_depth, *est_qgc;      // qk_est_begin .. qk_est_end: a sample's depths and the GC flags, by ordinal
    uint64_t est_n;
    unsigned long long *gc_acc;         // qk_gc_begin .. qk_gc_end: 3 x 401 sums + 1 count of out-of-range bins

    // device-side record framing (qk_frame.cu)
    unsigned long long *frame_stream;   // device: [0] FSM state, [1] read lines, [2] bases, [3] raw lines
    uint32_t *frame_elems;              // device: n_slots x QK_FRAME_MAX_CTAS per-CTA words
    int raw_fastq, raw_active, raw_prev_slot;
    uint64_t frame_launches;
};

int qk_fail(qk_ctx *ctx, int code, const char *fmt, ...);
int qk_cuda_fail(qk_ctx *ctx, cudaError_t e, const char *what);
#define QK_CUDA(ctx, call)                                            \
    do {                                                              \
        cudaError_t e__ = (call);                                     \
        if (e__ != cudaSuccess) return qk_cuda_fail(ctx, e__, #call); \
    } while (0)

// ---- 60-bit bijective mixer ---------------------------------------------------------
// xorshift and odd multiplication are both invertible mod 2^60, so distinct keys map to
// distinct (bucket, remainder) pairs and the remainder identifies the key in its bucket.
// One round is enough here: the bucket index is the TOP bits of the product, which depend on
// every bit of the input; the remainder (low bits) only has to be a bijection, not mixed.
// (Build-time check: the stash fill matches the Poisson expectation, tests/test_gpu_parity.py.)
#define QK_M60 0x0FFFFFFFFFFFFFFFull
__host__ __device__ __forceinline__ uint64_t qk_mix60(uint64_t x)
{
    x ^= x >> 31;
    return (x * 0x9E3779B97F4A7C15ull) & QK_M60;
}
// second, independent hash for the stash
__host__ __device__ __forceinline__ uint64_t qk_mix_stash(uint64_t x)
{
    x ^= x >> 33;
    x *= 0xFF51AFD7ED558CCDull;
    x ^= x >> 29;
    x *= 0xC4CEB9FE1A85EC53ull;
    x ^= x >> 32;
    return x;
}

// The reference's slot hash (Q.c:66-76), needed at build time only: which of several
// duplicate keys Find_hash would reach first.
__host__ __device__ __forceinline__ uint64_t qk_djb(uint64_t key)
{
    uint64_t h = 5381;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        h = h * 33u + (key & 0xFFu);
        key >>= 8;
    }
    return h;
}
