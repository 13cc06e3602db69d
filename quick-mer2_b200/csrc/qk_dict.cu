// qk_dict.cu -- device dictionary: QM11 arrays -> (ordinal per chain entry) -> bucketised
// table key -> ordinal.
//
// Replaces, on the device:
//   * the key load of Q.c:353-359 and the chain load of Q.c:483,490
//   * the serial chain walk of Q.c:498-516 (c = next[c] once per k-mer): here the chain is
//     LIST-RANKED in parallel so that counters can be indexed by ordinal (= .bin index)
//   * Find_hash's table (Q.c:90-99) -- slot placement is not observable in the .bin, so
//     the table is rebuilt with a strong mixer, 32-byte buckets and quotiented entries
//
// List ranking (DESIGN.md "dictionary build"):
//   1. every slot s with s % stride == 0, plus first_idx (the head), is a splitter; one
//      thread per splitter walks next[] to the following splitter and records
//      (successor, segment length)                                          [walk kernel]
//   2. pointer jumping (Wyllie) over the short splitter list gives each splitter its
//      distance to the end of the chain, hence its ordinal, and marks the splitters the
//      head reaches                                                         [jump kernel]
//   3. every marked splitter re-walks its segment handing out consecutive ordinals
//                                                                        [scatter kernel]
// What the reference's `count` reads is the cycle that starts at first_idx, whatever else
// the arrays hold (Q.c:498-516), so that is what is ranked.  Occupied slots that are not on
// the chain (`sparse` leaves them behind when it thins without resizing, Q.c:1443-1461) take
// part in Find_hash's probing but never get an ordinal; an EMPTY slot on the chain (an
// `index` input containing the poly-A k-mer, key 0) gets one, and receives the counts of
// key-0 k-mers iff it is the slot Find_hash(0) stops at (Q.c:98).  Only a chain that does
// not come back to first_idx is rejected (QK_ERR_FORMAT).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "qk_common.cuh"

// A walk ends at the next slot that is a multiple of the stride; chain order is
// unrelated to slot order, so segment lengths are geometric with mean = stride (<= 512) and
// 2^20 steps is unreachable for a valid dictionary placed by hashing.
#define QK_WALK_CAP (1u << 20)

#include <time.h>
static double qk_now(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + ts.tv_nsec * 1e-9;
}

enum { QK_FLAG_WALK_CAP = 1, QK_FLAG_STASH_FULL = 4 };

struct qk_build_info {
    unsigned long long occupied;
    unsigned long long skipped;
    unsigned long long stash_used;
    unsigned int flags;
    unsigned int pad;
};

extern "C" int qk_dict_begin(qk_ctx *ctx, uint8_t k, uint64_t hash_size, uint64_t first_idx)
{
    if (!ctx) return QK_ERR_ARG;
    if (k < 1 || k > 32) return qk_fail(ctx, QK_ERR_ARG, "k=%u outside 1..32", (unsigned)k);
    if (hash_size < 2 || (hash_size & (hash_size - 1)) || hash_size > ((uint64_t)1 << 32))
        return qk_fail(ctx, QK_ERR_FORMAT, "hash size 0x%llx is not a power of two <= 2^32 (Q.c:20 chain is u32)",
                       (unsigned long long)hash_size);
    if (first_idx >= hash_size) return qk_fail(ctx, QK_ERR_FORMAT, "first index outside the table");
    QK_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaFree(ctx->raw_keys);
    cudaFree(ctx->raw_next);
    cudaFree(ctx->buckets);
    cudaFree(ctx->stash);
    cudaFree(ctx->counters_buf[0]);
    cudaFree(ctx->counters_buf[1]);
    ctx->raw_keys = NULL; ctx->raw_next = NULL; ctx->buckets = NULL; ctx->stash = NULL; ctx->counters = NULL;
    ctx->counters_buf[0] = ctx->counters_buf[1] = NULL;
    ctx->dict_state = 0;
    QK_CUDA(ctx, cudaMalloc((void **)&ctx->raw_keys, hash_size * sizeof(uint64_t)));
    QK_CUDA(ctx, cudaMalloc((void **)&ctx->raw_next, hash_size * sizeof(uint32_t)));
    ctx->k = k;
    ctx->hash_size = hash_size;
    ctx->first_idx = first_idx;
    ctx->dict_state = 1;
    return QK_OK;
}

extern "C" int qk_dict_upload_keys(qk_ctx *ctx, uint64_t slot_offset, const uint64_t *keys, uint64_t count)
{
    if (!ctx || !keys) return QK_ERR_ARG;
    if (ctx->dict_state != 1) return qk_fail(ctx, QK_ERR_STATE, "qk_dict_begin not called");
    if (slot_offset + count > ctx->hash_size) return qk_fail(ctx, QK_ERR_ARG, "key range outside the table");
    QK_CUDA(ctx, cudaMemcpy(ctx->raw_keys + slot_offset, keys, count * sizeof(uint64_t), cudaMemcpyHostToDevice));
    return QK_OK;
}

extern "C" int qk_dict_upload_chain(qk_ctx *ctx, uint64_t slot_offset, const uint32_t *next, uint64_t count)
{
    if (!ctx || !next) return QK_ERR_ARG;
    if (ctx->dict_state != 1) return qk_fail(ctx, QK_ERR_STATE, "qk_dict_begin not called");
    if (slot_offset + count > ctx->hash_size) return qk_fail(ctx, QK_ERR_ARG, "chain range outside the table");
    QK_CUDA(ctx, cudaMemcpy(ctx->raw_next + slot_offset, next, count * sizeof(uint32_t), cudaMemcpyHostToDevice));
    return QK_OK;
}

// Asynchronous variant: the data already sits in the pinned buffer of `slot`; the copy is
// enqueued on the slot's stream and qk_wait_slot(slot) tells when the buffer may be refilled.
// kind 0 = keys (count x u64 at element offset `elem_offset`), 1 = chain (u32).
extern "C" int qk_dict_upload_from_slot(qk_ctx *ctx, uint32_t slot, int kind, uint64_t elem_offset, uint64_t count)
{
    if (!ctx || slot >= ctx->n_slots) return QK_ERR_ARG;
    if (ctx->dict_state != 1) return qk_fail(ctx, QK_ERR_STATE, "qk_dict_begin not called");
    const size_t esz = kind ? sizeof(uint32_t) : sizeof(uint64_t);
    if (elem_offset + count > ctx->hash_size || count * esz > ctx->chunk_capacity)
        return qk_fail(ctx, QK_ERR_ARG, "upload range outside the table or larger than a slot");
    QK_CUDA(ctx, cudaSetDevice(ctx->device));
    qk_slot *sl = &ctx->slots[slot];
    void *dst = kind ? (void *)(ctx->raw_next + elem_offset) : (void *)(ctx->raw_keys + elem_offset);
    QK_CUDA(ctx, cudaMemcpyAsync(dst, sl->host, count * esz, cudaMemcpyHostToDevice, sl->stream));
    QK_CUDA(ctx, cudaEventRecord(sl->h2d_done, sl->stream));
    return QK_OK;
}

// ---- kernels ---------------------------------------------------------------------------

// Node ids: 0..n_split-1 = slot id*stride (occupied or not: an empty slot can be on the chain), n_split = head
// (first_idx), n_split+1 = END.
__device__ __forceinline__ bool qk_node_slot(uint64_t id, uint64_t n_split, uint32_t stride_log2, uint64_t first, uint64_t *slot)
{
    if (id == n_split) { *slot = first; return true; }
    uint64_t s = id << stride_log2;
    *slot = s;
    return s != first;
}

__global__ void qk_walk_kernel(const uint32_t *__restrict__ next, uint64_t hash_size, uint64_t n_split,
                               uint32_t stride_log2, uint64_t first, uint32_t *succ, unsigned long long *dist,
                               uint32_t *seg_len, qk_build_info *info)
{
    uint64_t id = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id > n_split + 1) return;
    const uint32_t END = (uint32_t)(n_split + 1);
    uint64_t slot;
    if (id == n_split + 1 || !qk_node_slot(id, n_split, stride_log2, first, &slot)) {
        succ[id] = END; dist[id] = 0; seg_len[id] = 0;
        return;
    }
    const uint64_t smask = ((uint64_t)1 << stride_log2) - 1;
    uint64_t c = slot;
    uint32_t len = 0;
    do {
        c = next[c];
        ++len;
        if (c >= hash_size) {       // a pointer out of the table (off the chain next[] may hold anything): this walk never ends
            c = 0;
            len = QK_WALK_CAP;
        }
    } while (c != first && (c & smask) != 0 && len < QK_WALK_CAP);
    (void)info;
    succ[id] = (c == first) ? END : (uint32_t)(c >> stride_log2);
    dist[id] = len;
    seg_len[id] = len;      // == QK_WALK_CAP: the walk did not end (fatal only if the head reaches this node)
}

// One round of pointer jumping: dist = distance (in chain entries) to END.
// ... and of reachability from the head: succ_in is succ^(2^j) in round j, so marking
// succ_in[i] for every marked i doubles the marked prefix of the chain each round (a node
// marked during the round may or may not pass it on in the same round; the nodes it would
// reach are reached from already-marked ones in the next).
__global__ void qk_jump_kernel(const uint32_t *__restrict__ succ_in, const unsigned long long *__restrict__ dist_in,
                               uint32_t *succ_out, unsigned long long *dist_out, uint64_t n_nodes, unsigned char *mark)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    uint32_t s = succ_in[i];
    dist_out[i] = dist_in[i] + dist_in[s];
    succ_out[i] = succ_in[s];
    if (mark[i]) mark[s] = 1;
}

// a walk that hit the cap matters only on the chain
__global__ void qk_check_caps_kernel(const uint32_t *__restrict__ seg_len, const unsigned char *__restrict__ mark, uint64_t n_nodes,
                                     qk_build_info *info)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_nodes && mark[i] && seg_len[i] >= QK_WALK_CAP) atomicOr(&info->flags, QK_FLAG_WALK_CAP);
}

// Does Find_hash (Q.c:90-99) reach `slot` when asked for its own key?  False only for the
// later copies of a duplicated key (dictionaries written by `index`).
__device__ __forceinline__ bool qk_is_primary(const uint64_t *__restrict__ keys, uint64_t hash_size, uint64_t slot,
                                              uint64_t key)
{
    uint64_t s = qk_djb(key) & (hash_size - 1);
    const long long step = (s & (hash_size >> 1)) ? -1 : 1;
    for (;;) {
        if (s == slot) return true;
        uint64_t v = keys[s];
        if (v == key) return false; // an earlier copy shadows this slot
        if (v == 0) return false;   // unreachable from its home: never found by the reference
        s = (uint64_t)((long long)s + step);
        if (s >= hash_size) return false;
    }
}

struct qk_build_params {
    qk_bucket *buckets;
    qk_stash_entry *stash;
    uint64_t stash_mask;
    uint64_t stash_limit;
    uint32_t rem_bits, ord_bits;
};

__device__ __forceinline__ void qk_table_insert(const qk_build_params &bp, uint64_t key, uint64_t ord1, uint32_t strand,
                                                qk_build_info *info)
{
    const uint64_t h = qk_mix60(key);
    const uint64_t bucket = h >> bp.rem_bits;
    const uint64_t rem = h & (((uint64_t)1 << bp.rem_bits) - 1);
    const unsigned long long entry = ((unsigned long long)strand << 63) | (rem << bp.ord_bits) | ord1;
    unsigned long long *e = bp.buckets[bucket].e;
#pragma unroll
    for (int i = 0; i < QK_BUCKET_ENTRIES; ++i) {
        if (e[i] == 0 && atomicCAS(&e[i], 0ull, entry) == 0ull) return;   // (entries are never 0 again once set)
    }
    // home bucket full: stash, and say so in the bucket (its last entry is taken and final)
    atomicOr(&e[QK_BUCKET_ENTRIES - 1], QK_ENTRY_OVERFLOW);
    unsigned long long used = atomicAdd(&info->stash_used, 1ull);
    if (used >= bp.stash_limit) { atomicOr(&info->flags, QK_FLAG_STASH_FULL); return; }
    uint64_t s = qk_mix_stash(key) & bp.stash_mask;
    for (;;) {
        unsigned long long old = atomicCAS(&bp.stash[s].key, 0ull, (unsigned long long)key | QK_STASH_TAKEN);
        if (old == 0ull) { bp.stash[s].ord1 = (uint32_t)ord1; bp.stash[s].pad = strand; return; }
        s = (s + 1) & bp.stash_mask;
    }
}

// Segment walk #2: every splitter hands out the ordinals of its segment and leaves the key of
// ordinal o in kbo[o].  Bit 63 flags the ordinals that must never match: keys >= 2^60 (no read
// can produce them) and the later copies of a duplicated key (Find_hash never reaches them).
#define QK_KBO_SKIP 0x8000000000000000ull
__global__ void qk_scatter_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ next, uint64_t hash_size,
                                  uint64_t n_split, uint32_t stride_log2, uint64_t first,
                                  const unsigned long long *__restrict__ dist, const uint32_t *__restrict__ seg_len,
                                  const unsigned char *__restrict__ mark, unsigned long long total,
                                  unsigned long long *__restrict__ kbo, qk_build_info *info)
{
    uint64_t id = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id > n_split || !mark[id]) return;      // only the walkers the head reaches are on the chain
    uint32_t len = seg_len[id];
    if (len == 0) return;
    uint64_t c = (id == n_split) ? first : (id << stride_log2);
    uint64_t ord = total - dist[id];
    for (uint32_t i = 0; i < len; ++i, ++ord) {
        uint64_t key = keys[c];
        uint64_t nx = next[c];
        // (key 0 = an empty slot on the chain: it gets an ordinal like any other, and is the
        //  one Find_hash(0) reaches iff it is the first empty slot on key 0's probe path --
        //  exactly what qk_is_primary checks)
        if ((key >> QK_KEY_BITS) != 0 || !qk_is_primary(keys, hash_size, c, key)) {
            atomicAdd(&info->skipped, 1ull);
            key |= QK_KBO_SKIP;
        }
        kbo[ord] = key;
        c = nx;
    }
}

// reverse complement of a 30-mer in the reference's encoding (A=0 C=1 T=2 G=3, complement = ^2)
__device__ __forceinline__ uint64_t qk_rc30(uint64_t x)
{
    uint64_t z = __brevll(x);
    z = ((z & 0x5555555555555555ull) << 1) | ((z >> 1) & 0x5555555555555555ull);
    return (z >> 4) ^ 0x0AAAAAAAAAAAAAAAull;
}

// Orientation pass + table insert, in ordinal order.  The dictionary in chain order is a
// sequence of overlapping k-mers (consecutive reference positions) broken wherever a k-mer
// was not unique.  For k = 30 (canonical key = min(fwd, rc), both true 30-mers) choose for
// every ordinal o an orientation F_o in {K_o, rc(K_o)} and record
//     cont[o]  = 1 iff F_o is F_{o-1} shifted by one base  (F_o >> 2 == F_{o-1} & (2^58-1))
//     last[o]  = last base of F_o,  first[o] = first base of F_o,  strand[o] = (F_o == K_o)
// so that the count kernel can walk from one verified k-mer of a read to its neighbours without
// probing the table (qk_count_ext_kernel).  These are statements about the KEYS alone, so the
// walk is exact whatever produced the dictionary.  One thread per QK_EXT_BLOCK ordinals; the
// first ordinal of a block never continues (the price of not chaining the blocks).
#define QK_EXT_BLOCK 256
__global__ void qk_orient_insert_kernel(const unsigned long long *__restrict__ kbo, uint64_t n, int with_ext, uint64_t kmask,
                                        qk_build_params bp, uint32_t *__restrict__ ext, qk_build_info *info)
{
    const uint64_t blk = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t begin = blk * QK_EXT_BLOCK;
    if (begin >= n) return;
    const uint64_t end = begin + QK_EXT_BLOCK < n ? begin + QK_EXT_BLOCK : n;
    const uint64_t M58 = ((uint64_t)1 << 58) - 1;
    bool have_prev = false;
    uint64_t prevF = 0;
    uint32_t wl = 0, wf = 0, wc = 0, ws = 0;
    for (uint64_t o = begin; o < end; ++o) {
        const uint64_t raw = kbo[o];
        const uint64_t K = raw & ~QK_KBO_SKIP;
        uint32_t cont = 0, strand = 1;
        uint64_t F = K;
        if (raw & QK_KBO_SKIP) {
            have_prev = false;
        } else {
            if (with_ext == 2) {          // forward chain: F_o = K_o; continues iff K_o = ((K_{o-1} << 2) | base) & mask
                cont = have_prev && (K >> 2) == (prevF & (kmask >> 2));
                prevF = K;
                have_prev = true;
            } else if (with_ext) {        // 1: canonical 30-mers; 3: k = 31, whose keys are 30-mers too -- the newest 30 bases,
                                          // forward when the 31st base back is A and that is the smaller form, else reverse complement
                const uint64_t Kr = qk_rc30(K);
                if (have_prev) {
                    const uint64_t want = prevF & M58;
                    if ((K >> 2) == want) { F = K; cont = 1; }
                    else if ((Kr >> 2) == want) { F = Kr; cont = 1; }
                }
                if (!cont && o + 1 < n) { // a run starts here: face the way the next k-mer continues
                    const uint64_t nraw = kbo[o + 1];
                    if (!(nraw & QK_KBO_SKIP)) {
                        const uint64_t nK = nraw, nKr = qk_rc30(nraw);
                        if ((nK >> 2) == (K & M58) || (nKr >> 2) == (K & M58)) F = K;
                        else if ((nK >> 2) == (Kr & M58) || (nKr >> 2) == (Kr & M58)) F = Kr;
                    }
                }
                strand = F == K;
                prevF = F;
                have_prev = true;
            }
            qk_table_insert(bp, K, o + 1, strand, info);
        }
        if (with_ext) {
            const uint32_t i = (uint32_t)(o - begin);
            wl |= (uint32_t)(F & 3) << (2 * (i & 15));
            wf |= (uint32_t)((F >> 58) & 3) << (2 * (i & 15));
            wc |= cont << (i & 15);
            ws |= ((raw & QK_KBO_SKIP) ? 0u : strand) << (i & 15);
            if ((i & 15) == 15 || o + 1 == end) {                // one group of 16 ordinals: three adjacent words
                uint32_t *g = ext + (o >> 4) * QK_EXT_GROUP_WORDS;
                g[0] = wl;
                g[1] = wf;
                g[2] = wc | (ws << 16);                          // high half: F_o == K_o (read by the k = 31 walk)
                wl = wf = wc = ws = 0;
            }
        }
    }
}

// ---- host side -------------------------------------------------------------------------
static uint32_t qk_bits_for(uint64_t v) // smallest b with v < 2^b
{
    uint32_t b = 0;
    while (b < 64 && (v >> b) != 0) ++b;
    return b;
}

// Geometry for n chain entries: buckets = smallest power of two with <= 2.1 keys per
// 4-entry bucket on average (<= 4 % of the keys overflow into the stash); the stash is sized
// from the Poisson overflow expectation.
static void qk_geometry(uint64_t n, uint32_t k, uint64_t stash_slots_min, qk_table_desc *d)
{
    memset(d, 0, sizeof *d);
    d->n_kmers = n;
    d->k = k;
    uint64_t nb = 1;
    while ((double)n / (double)nb > 2.1) nb <<= 1;
    if (nb < 64) nb = 64;
    d->n_buckets = nb;
    d->bucket_bits = qk_bits_for(nb - 1);
    d->rem_bits = QK_KEY_BITS - d->bucket_bits;
    d->ord_bits = qk_bits_for(n); // holds ordinal + 1 <= n
    if (d->ord_bits == 0) d->ord_bits = 1;
    while (d->rem_bits + d->ord_bits > 62) { // keep bits 63 and 62 of an entry spare: small dictionaries get more buckets
        nb <<= 1;
        d->n_buckets = nb;
        d->bucket_bits = qk_bits_for(nb - 1);
        d->rem_bits = QK_KEY_BITS - d->bucket_bits;
    }
    // expected fraction of keys that find their 4-entry bucket full, keys/bucket ~ Poisson(lambda)
    const double lambda = (double)n / (double)nb;
    double p = exp(-lambda), over = 0;
    for (int x = 1; x <= 64; ++x) {
        p *= lambda / x;
        if (x > QK_BUCKET_ENTRIES) over += (x - QK_BUCKET_ENTRIES) * p;
    }
    uint64_t want = (uint64_t)(over / lambda * (double)n * 1.25) + 1024;
    if (want < stash_slots_min) want = stash_slots_min;
    uint64_t ss = 1024;
    while (ss < 2 * want) ss <<= 1;
    d->stash_slots = ss;
    d->table_bytes = nb * sizeof(qk_bucket);
    d->stash_bytes = ss * sizeof(qk_stash_entry);
    // dictionary-order extension arrays (k = 30 only: for other k the reference's canonical key
    // mixes a k-mer with a 30-base reverse complement, Q.c:415-420, and is not a walkable k-mer)
    // which dictionary-order chain the keys form (qk_count.cu, qk_count_ext32_kernel): 1 = canonical 30-mers (k = 30);
    // 2 = forward k-mers (k < 30: the key is the forward k-mer unless the newest bases are all T); 3 = 30-mers in
    // the form the k = 31 key rule gives them (forward iff the 31st base back is A and forward <= reverse
    // complement); 0 = none (k = 32: every read key is 0, Q.c:419)
    d->has_ext = getenv("QK_NO_EXT") != NULL ? 0 : k == 30 ? 1 : (k >= 3 && k < 30) ? 2 : k == 31 ? 3 : 0;
    d->ext_bytes = d->has_ext ? ((n + 15) / 16 + 4) * QK_EXT_GROUP_WORDS * sizeof(uint32_t) : 0; // 5 bits per ordinal in groups of 16, padded
    d->cont_bytes = 0;                                                        // (the continuation bits live in the same array)
}

// Geometry a dictionary of n k-mers will get, without a device: lets a host size the job
// (table_bytes + stash_bytes + ext_bytes + 4 (n + 1) bytes of counters).
extern "C" int qk_table_geometry(uint64_t n_kmers, uint32_t k, qk_table_desc *desc)
{
    if (!desc || n_kmers == 0 || n_kmers >= ((uint64_t)1 << 32) || k < 1 || k > 32) return QK_ERR_ARG;
    qk_geometry(n_kmers, k, 0, desc);
    return QK_OK;
}

static int qk_alloc_table(qk_ctx *ctx, const qk_table_desc *d)
{
    cudaFree(ctx->buckets);
    cudaFree(ctx->stash);
    cudaFree(ctx->counters_buf[0]);
    cudaFree(ctx->counters_buf[1]);
    ctx->counters_buf[0] = ctx->counters_buf[1] = NULL;
    cudaFree(ctx->ext);
    ctx->buckets = NULL; ctx->stash = NULL; ctx->counters = NULL;
    ctx->ext = NULL;
    QK_CUDA(ctx, cudaMalloc((void **)&ctx->buckets, d->table_bytes));
    QK_CUDA(ctx, cudaMalloc((void **)&ctx->stash, d->stash_bytes));
    if (d->has_ext) {
        QK_CUDA(ctx, cudaMalloc((void **)&ctx->ext, d->ext_bytes));
        QK_CUDA(ctx, cudaMemset(ctx->ext, 0, d->ext_bytes));
    }
    QK_CUDA(ctx, cudaMalloc((void **)&ctx->counters_buf[0], (d->n_kmers + 1) * sizeof(uint32_t)));
    ctx->counters = ctx->counters_buf[0];
    QK_CUDA(ctx, cudaMemset(ctx->counters, 0, (d->n_kmers + 1) * sizeof(uint32_t)));
    QK_CUDA(ctx, cudaMemset(ctx->stats, 0, QK_STATS_WORDS * sizeof(unsigned long long))); // a new dictionary starts a new count
    QK_CUDA(ctx, cudaMemset(ctx->frame_stream, 0, 4 * sizeof(unsigned long long)));
    QK_CUDA(ctx, cudaDeviceSynchronize()); // the slot streams do not order against stream 0
    ctx->lines = 0;
    ctx->kernel_ms = ctx->h2d_ms = 0;
    ctx->launches = 0;
    return QK_OK;
}

extern "C" int qk_dict_build(qk_ctx *ctx, uint64_t *n_kmers_out)
{
    if (!ctx) return QK_ERR_ARG;
    if (ctx->dict_state != 1) return qk_fail(ctx, QK_ERR_STATE, "qk_dict_begin/upload not called");
    QK_CUDA(ctx, cudaSetDevice(ctx->device));
    for (uint32_t s = 0; s < ctx->n_slots; ++s) QK_CUDA(ctx, cudaStreamSynchronize(ctx->slots[s].stream)); // async uploads
    const uint64_t H = ctx->hash_size, first = ctx->first_idx;
    // stride: about 2^19 walkers or more, segments of 16..512 slots
    uint32_t stride_log2 = 4;
    while (stride_log2 < 9 && (H >> (stride_log2 + 1)) >= ((uint64_t)1 << 19)) ++stride_log2;
    if (((uint64_t)1 << stride_log2) > H) stride_log2 = qk_bits_for(H - 1);
    const uint64_t n_split = H >> stride_log2;
    const uint64_t n_nodes = n_split + 2;

    const int verbose = getenv("QK_TIMING") != NULL;
    double tm0 = qk_now(), tm1 = tm0, tm2 = tm0, tm3 = tm0;
    qk_build_info *info = NULL;
    uint32_t *succ[2] = {NULL, NULL}, *seg_len = NULL;
    unsigned char *mark = NULL;
    uint32_t head_succ = 0;
    unsigned long long *dist[2] = {NULL, NULL};
    int rc = QK_OK;
#define QK_TRY(call)                                                               \
    do {                                                                           \
        cudaError_t e__ = (call);                                                  \
        if (e__ != cudaSuccess) { rc = qk_cuda_fail(ctx, e__, #call); goto done; } \
    } while (0)
    qk_build_info hinfo;
    qk_table_desc d;
    unsigned long long total = 0;
    int cur = 0;
    uint64_t stash_min = 0, skipped = 0;
    unsigned long long *kbo = NULL;

    QK_TRY(cudaMalloc((void **)&info, sizeof(qk_build_info)));
    QK_TRY(cudaMemset(info, 0, sizeof(qk_build_info)));
    for (int i = 0; i < 2; ++i) {
        QK_TRY(cudaMalloc((void **)&succ[i], n_nodes * sizeof(uint32_t)));
        QK_TRY(cudaMalloc((void **)&dist[i], n_nodes * sizeof(unsigned long long)));
    }
    QK_TRY(cudaMalloc((void **)&seg_len, n_nodes * sizeof(uint32_t)));
    QK_TRY(cudaMalloc((void **)&mark, n_nodes));
    QK_TRY(cudaMemset(mark, 0, n_nodes));
    QK_TRY(cudaMemset(mark + n_split, 1, 1));    // the head

    qk_walk_kernel<<<(unsigned)((n_nodes + 127) / 128), 128>>>(ctx->raw_next, H, n_split, stride_log2, first,
                                                               succ[0], dist[0], seg_len, info);
    QK_TRY(cudaGetLastError());
    for (uint64_t span = 1; span < n_nodes; span <<= 1) {
        qk_jump_kernel<<<(unsigned)((n_nodes + 255) / 256), 256>>>(succ[cur], dist[cur], succ[cur ^ 1], dist[cur ^ 1],
                                                                   n_nodes, mark);
        cur ^= 1;
    }
    qk_check_caps_kernel<<<(unsigned)((n_nodes + 255) / 256), 256>>>(seg_len, mark, n_nodes, info);
    QK_TRY(cudaGetLastError());
    QK_TRY(cudaMemcpy(&total, dist[cur] + n_split, sizeof total, cudaMemcpyDeviceToHost));
    QK_TRY(cudaMemcpy(&hinfo, info, sizeof hinfo, cudaMemcpyDeviceToHost));
    if (hinfo.flags & QK_FLAG_WALK_CAP) {
        rc = qk_fail(ctx, QK_ERR_FORMAT, "chain walk did not reach a splitter within %u steps: corrupt chain", QK_WALK_CAP);
        goto done;
    }
    QK_TRY(cudaMemcpy(&head_succ, succ[cur] + n_split, sizeof head_succ, cudaMemcpyDeviceToHost));
    // (after >= n_nodes jumps the head either points at END or sits on a cycle that never
    //  comes back to first_idx; a path that does come back is a simple cycle, so it cannot
    //  be longer than the table)
    if (head_succ != (uint32_t)(n_split + 1) || total == 0 || total > H) {
        rc = qk_fail(ctx, QK_ERR_FORMAT, "the chain from first_idx does not come back to it: not a QM11 chain");
        goto done;
    }

    tm1 = qk_now(); // chain ranked
    // keys in ordinal order; after this the raw QM11 arrays are no longer needed
    QK_TRY(cudaMalloc((void **)&kbo, (total + 1) * sizeof(unsigned long long)));
    QK_TRY(cudaMemset(info, 0, sizeof(qk_build_info)));
    qk_scatter_kernel<<<(unsigned)((n_split + 1 + 127) / 128), 128>>>(ctx->raw_keys, ctx->raw_next, H, n_split, stride_log2, first,
                                                                      dist[cur], seg_len, mark, total, kbo, info);
    QK_TRY(cudaGetLastError());
    QK_TRY(cudaMemcpy(&hinfo, info, sizeof hinfo, cudaMemcpyDeviceToHost));
    skipped = hinfo.skipped;
    tm2 = qk_now(); // keys scattered by ordinal
    cudaFree(ctx->raw_keys);
    cudaFree(ctx->raw_next);
    ctx->raw_keys = NULL;
    ctx->raw_next = NULL;

    for (int attempt = 0; attempt < 4; ++attempt) {
        qk_geometry(total, ctx->k, stash_min, &d);
        if (d.rem_bits + d.ord_bits > 62 || d.ord_bits > 32) {
            rc = qk_fail(ctx, QK_ERR_FORMAT, "entry needs %u bits", d.rem_bits + d.ord_bits);
            goto done;
        }
        rc = qk_alloc_table(ctx, &d);
        if (rc) goto done;
        QK_TRY(cudaMemset(ctx->buckets, 0, d.table_bytes));
        QK_TRY(cudaMemset(ctx->stash, 0, d.stash_bytes));
        QK_TRY(cudaMemset(info, 0, sizeof(qk_build_info)));
        qk_build_params bp;
        bp.buckets = ctx->buckets;
        bp.stash = ctx->stash;
        bp.stash_mask = d.stash_slots - 1;
        bp.stash_limit = d.stash_slots / 2;
        bp.rem_bits = d.rem_bits;
        bp.ord_bits = d.ord_bits;
        const uint64_t n_blocks = (total + QK_EXT_BLOCK - 1) / QK_EXT_BLOCK;
        qk_orient_insert_kernel<<<(unsigned)((n_blocks + 63) / 64), 64>>>(kbo, total, (int)d.has_ext,
                                                                          ctx->k >= 32 ? 0 : (((uint64_t)1 << (2 * ctx->k)) - 1), bp, ctx->ext, info);
        QK_TRY(cudaGetLastError());
        QK_TRY(cudaMemcpy(&hinfo, info, sizeof hinfo, cudaMemcpyDeviceToHost));
        if (!(hinfo.flags & QK_FLAG_STASH_FULL)) break;
        stash_min = hinfo.stash_used + 1024; // retry with a stash that holds what was needed
        if (attempt == 3) { rc = qk_fail(ctx, QK_ERR_NOMEM, "stash overflow after 4 attempts"); goto done; }
    }
    tm3 = qk_now(); // table built
    if (verbose)
        fprintf(stderr, "[qk] build: rank %.3f s, scatter %.3f s, alloc+orient+insert %.3f s (%llu k-mers, %llu buckets)\n", tm1 - tm0,
                tm2 - tm1, tm3 - tm2, total, (unsigned long long)d.n_buckets);
    d.stash_used = hinfo.stash_used;
    d.skipped_keys = skipped;
    ctx->desc = d;
    ctx->dict_state = 2;
    if (n_kmers_out) *n_kmers_out = total;

done:
    cudaFree(kbo);
    cudaFree(info);
    cudaFree(succ[0]); cudaFree(succ[1]);
    cudaFree(dist[0]); cudaFree(dist[1]);
    cudaFree(seg_len);
    cudaFree(mark);
    cudaFree(ctx->raw_keys);
    cudaFree(ctx->raw_next);
    ctx->raw_keys = NULL;
    ctx->raw_next = NULL;
    if (rc) ctx->dict_state = 0;
    return rc;
#undef QK_TRY
}

extern "C" int qk_dict_describe(const qk_ctx *ctx, qk_table_desc *desc)
{
    if (!ctx || !desc) return QK_ERR_ARG;
    if (ctx->dict_state != 2) return QK_ERR_STATE;
    *desc = ctx->desc;
    return QK_OK;
}

extern "C" int qk_dict_adopt(qk_ctx *ctx, const qk_table_desc *desc)
{
    if (!ctx || !desc) return QK_ERR_ARG;
    if (desc->n_buckets == 0 || (desc->n_buckets & (desc->n_buckets - 1)) || desc->stash_slots == 0 ||
        (desc->stash_slots & (desc->stash_slots - 1)) || desc->table_bytes != desc->n_buckets * sizeof(qk_bucket) ||
        desc->stash_bytes != desc->stash_slots * sizeof(qk_stash_entry) || desc->rem_bits + desc->ord_bits > 62 ||
        desc->ord_bits > 32 || (desc->has_ext && ((desc->has_ext == 1) != (desc->k == 30) || desc->has_ext > 3 || desc->k > 31 || desc->ext_bytes < ((desc->n_kmers + 15) / 16 + 4) * QK_EXT_GROUP_WORDS * 4)) ||
        desc->rem_bits + desc->bucket_bits != QK_KEY_BITS || desc->k < 1 || desc->k > 32)
        return qk_fail(ctx, QK_ERR_ARG, "inconsistent table descriptor");
    QK_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = qk_alloc_table(ctx, desc);
    if (rc) return rc;
    ctx->desc = *desc;
    ctx->k = (uint8_t)desc->k;
    ctx->dict_state = 2;
    return QK_OK;
}

extern "C" int qk_dict_ext_ptrs(const qk_ctx *ctx, void **last, void **first, void **cont)
{
    if (!ctx) return QK_ERR_ARG;
    if (ctx->dict_state != 2) return QK_ERR_STATE;
    if (last) *last = ctx->ext;      /* one array since round 2: 12 bytes per 16 ordinals (last, first, continuation) */
    if (first) *first = NULL;
    if (cont) *cont = NULL;
    return QK_OK;
}

extern "C" int qk_dict_device_ptrs(const qk_ctx *ctx, void **table, void **stash)
{
    if (!ctx) return QK_ERR_ARG;
    if (ctx->dict_state != 2) return QK_ERR_STATE;
    if (table) *table = ctx->buckets;
    if (stash) *stash = ctx->stash;
    return QK_OK;
}
