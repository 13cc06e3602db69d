// qk_synth_gpu.cu -- seeded synthetic inputs for the count path, generated ON THE GPU so that the
// human-scale configurations of BASELINE.json (3.1 Gb reference, 2^32-slot QM11 dictionary,
// ~600 M reads) can be made in seconds on the box that runs them.  TEST / BENCH DATA GENERATOR:
// nothing in the product (libquickmer2_b200.so, the command) links or loads this.
//
//   genome   a pure function of (seed, position): iid uniform ACGT; every `dup_period` bases the
//            last `dup_len` bases of the period are a copy of a random earlier stretch with
//            `div_ppm` substitutions (segmental duplications -> non-unique k-mers); optional
//            N block; n_contigs equal contigs.  Or uploaded from the host (tests).
//   dict     the QM11 dictionary `quicKmer2 search -e 0` would write for that genome: the
//            canonical k-mers that occur exactly once (Q.c:845-864 codec, Q.c:1217-1231 filter),
//            chained in reference order (Q.c:1048-1052), placed with the reference's probe rule
//            (DJB home slot, walk toward the middle of the table, Q.c:66-99) so that the
//            reference's own `count` reads it.  Same key list and chain order as `qk_synth dict`
//            (host/qk_synth.c, pinned to `search -e 0` by tests/test_oracle.py); slot placement
//            differs only in the order colliding keys were inserted, which `count` cannot see.
//            .qgc: GC count of the 400-base window around the k-mer + control flag for
//            alternating blocks, the definition of qk_synth.c.
//   reads    reference substrings at uniform starts, odd reads reverse-complemented, per-base
//            substitutions; fixed length (150 bp) or caller-given lengths (HiFi); written as
//            framed sequence lines, FASTA or 4-line FASTQ with fixed-width headers, straight
//            into device memory.  Read i is a pure function of (seed, i).
#include <cuda_runtime.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <cub/cub.cuh>

#define QS_OK 0
#define QS_ERR_CUDA 1
#define QS_ERR_ARG 2
#define QS_ERR_NOMEM 3
#define QS_ERR_IO 6
#define QS_ERR_FULL 7

#define QS_MAX_CONTIGS 256
#define QS_RUN 256 // positions per thread in the dictionary passes
#define QS_M60 0x0FFFFFFFFFFFFFFFull
#define QS_MULTI 0x8000000000000000ull

struct qs_contigs {
    uint64_t start[QS_MAX_CONTIGS + 1];
    uint32_t n;
};

struct qs_genome {
    int device;
    uint8_t *seq; // device, n bytes of ACGTN
    uint64_t n;
    qs_contigs contigs;
    // dictionary (device), valid after qs_dict_build
    uint64_t *keys;
    uint32_t *next;
    uint16_t *qgc;
    uint64_t slots, n_unique, first;
    uint32_t k;
    char err[256];
};

static int qs_fail(qs_genome *g, int code, const char *what, cudaError_t e)
{
    if (g) snprintf(g->err, sizeof g->err, "%s: %s", what, e == cudaSuccess ? "" : cudaGetErrorString(e));
    return code;
}
#define QS_CUDA(g, call)                                                                                         \
    do {                                                                                                         \
        cudaError_t e__ = (call);                                                                                \
        if (e__ != cudaSuccess) return qs_fail(g, e__ == cudaErrorMemoryAllocation ? QS_ERR_NOMEM : QS_ERR_CUDA, #call, e__); \
    } while (0)

__host__ __device__ __forceinline__ uint64_t qs_mix(uint64_t x)
{
    x ^= x >> 33; x *= 0xFF51AFD7ED558CCDull; x ^= x >> 33; x *= 0xC4CEB9FE1A85EC53ull; x ^= x >> 33;
    return x;
}
__host__ __device__ __forceinline__ uint64_t qs_hash(uint64_t seed, uint64_t a, uint64_t b)
{
    return qs_mix(qs_mix(seed + 0x9E3779B97F4A7C15ull * (a + 1)) ^ (0xD1B54A32D192ED03ull * (b + 1)));
}
__host__ __device__ __forceinline__ uint64_t qs_below(uint64_t h, uint64_t n)
{
#ifdef __CUDA_ARCH__
    return __umul64hi(h, n);
#else
    return (uint64_t)(((unsigned __int128)h * n) >> 64);
#endif
}

// ------------------------------------------------------------------------- genome --------
struct qs_gen_params {
    uint64_t n, seed;
    uint64_t dup_period, dup_len, nblock_at, nblock_len;
    uint32_t div_threshold; // substitution iff 24 bits of hash < this
};

__device__ __forceinline__ uint8_t qs_base_iid(uint64_t seed, uint64_t p)
{
    // 32 bases per hash
    const uint64_t h = qs_hash(seed, p >> 5, 0);
    return "ACGT"[(h >> (2 * (p & 31))) & 3];
}

__global__ void qs_genome_kernel(uint8_t *__restrict__ seq, qs_gen_params gp)
{
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < gp.n; p += (uint64_t)gridDim.x * blockDim.x) {
        uint8_t c;
        if (p >= gp.nblock_at && p < gp.nblock_at + gp.nblock_len) c = 'N';
        else {
            c = qs_base_iid(gp.seed, p);
            if (gp.dup_period) {
                const uint64_t blk = p / gp.dup_period, off = p % gp.dup_period;
                if (blk > 0 && off >= gp.dup_period - gp.dup_len) {
                    // copy of [src, src + dup_len) of the iid sequence, src anywhere before this period
                    const uint64_t src = qs_below(qs_hash(gp.seed, blk, 1), blk * gp.dup_period - gp.dup_len);
                    const uint64_t q = src + (off - (gp.dup_period - gp.dup_len));
                    c = qs_base_iid(gp.seed, q);
                    const uint64_t h = qs_hash(gp.seed, p, 2);
                    if ((uint32_t)(h & 0xFFFFFF) < gp.div_threshold) {
                        const uint32_t code = (c >> 1) & 3; // A0 C1 T2 G3
                        c = "ACTG"[(code + 1 + (uint32_t)((h >> 24) % 3)) & 3];
                    }
                }
            }
        }
        seq[p] = c;
    }
}

extern "C" const char *qs_last_error(const qs_genome *g) { return g ? g->err : "no genome"; }

static int qs_genome_new(qs_genome **out, int device, uint64_t n, uint32_t n_contigs)
{
    if (!out || n == 0 || n_contigs == 0 || n_contigs > QS_MAX_CONTIGS) return QS_ERR_ARG;
    qs_genome *g = (qs_genome *)calloc(1, sizeof *g);
    if (!g) return QS_ERR_NOMEM;
    *out = g;
    g->device = device;
    g->n = n;
    QS_CUDA(g, cudaSetDevice(device));
    QS_CUDA(g, cudaMalloc((void **)&g->seq, n + 64));
    QS_CUDA(g, cudaMemset(g->seq + n, 'N', 64));
    return QS_OK;
}

extern "C" int qs_genome_create(qs_genome **out, int device, uint64_t n_bases, uint32_t n_contigs, uint64_t seed,
                                uint64_t dup_period, uint64_t dup_len, uint32_t div_ppm, uint64_t nblock)
{
    int rc = qs_genome_new(out, device, n_bases, n_contigs);
    if (rc) return rc;
    qs_genome *g = *out;
    if (dup_period && (dup_len == 0 || dup_len * 2 > dup_period)) return qs_fail(g, QS_ERR_ARG, "dup_len must be <= dup_period / 2", cudaSuccess);
    const uint64_t per = n_bases / n_contigs;
    g->contigs.n = n_contigs;
    for (uint32_t c = 0; c < n_contigs; ++c) g->contigs.start[c] = c * per;
    g->contigs.start[n_contigs] = n_bases;
    qs_gen_params gp;
    gp.n = n_bases;
    gp.seed = seed;
    gp.dup_period = dup_period;
    gp.dup_len = dup_len;
    gp.nblock_len = nblock < n_bases / 2 ? nblock : 0;
    gp.nblock_at = n_bases / 3;
    gp.div_threshold = (uint32_t)((double)div_ppm * 1e-6 * 16777216.0 + 0.5);
    qs_genome_kernel<<<148 * 16, 256>>>(g->seq, gp);
    QS_CUDA(g, cudaGetLastError());
    QS_CUDA(g, cudaDeviceSynchronize());
    return QS_OK;
}

extern "C" int qs_genome_from_host(qs_genome **out, int device, const uint8_t *seq, uint64_t n, const uint64_t *starts,
                                   uint32_t n_contigs)
{
    if (!seq || !starts) return QS_ERR_ARG;
    int rc = qs_genome_new(out, device, n, n_contigs);
    if (rc) return rc;
    qs_genome *g = *out;
    g->contigs.n = n_contigs;
    for (uint32_t c = 0; c <= n_contigs; ++c) g->contigs.start[c] = starts[c];
    QS_CUDA(g, cudaMemcpy(g->seq, seq, n, cudaMemcpyHostToDevice));
    return QS_OK;
}

extern "C" int qs_genome_download(qs_genome *g, uint64_t offset, uint8_t *out, uint64_t count)
{
    if (!g || !out || offset + count > g->n) return QS_ERR_ARG;
    QS_CUDA(g, cudaSetDevice(g->device));
    QS_CUDA(g, cudaMemcpy(out, g->seq + offset, count, cudaMemcpyDeviceToHost));
    return QS_OK;
}

extern "C" int qs_genome_contigs(const qs_genome *g, uint64_t *starts, uint32_t cap, uint32_t *n_contigs)
{
    if (!g || !n_contigs) return QS_ERR_ARG;
    *n_contigs = g->contigs.n;
    if (starts)
        for (uint32_t c = 0; c <= g->contigs.n && c < cap; ++c) starts[c] = g->contigs.start[c];
    return QS_OK;
}

extern "C" const uint8_t *qs_genome_device_ptr(const qs_genome *g) { return g ? g->seq : NULL; }

static void qs_dict_release(qs_genome *g)
{
    cudaFree(g->keys);
    cudaFree(g->next);
    cudaFree(g->qgc);
    g->keys = NULL; g->next = NULL; g->qgc = NULL;
}

extern "C" int qs_dict_free(qs_genome *g)
{
    if (!g) return QS_ERR_ARG;
    cudaSetDevice(g->device);
    qs_dict_release(g);
    return QS_OK;
}

extern "C" void qs_genome_destroy(qs_genome *g)
{
    if (!g) return;
    cudaSetDevice(g->device);
    qs_dict_release(g);
    cudaFree(g->seq);
    cudaGetLastError();
    free(g);
}

// ------------------------------------------------------------------------- dictionary ----
// Canonical keys of positions [begin, end) of the genome, as `search` rolls them (Q.c:845-864):
// state reset at a contig start and at 'N'; emits once k bases are charged and the key is not 0.
// f(p, key) is called for every emitting position p (p = index of the k-mer's LAST base).
template <typename F>
__device__ __forceinline__ void qs_for_each_kmer(const uint8_t *__restrict__ seq, const qs_contigs &ct, uint64_t begin,
                                                 uint64_t end, uint32_t k, uint64_t kmask, F f)
{
    // contig of `begin`
    uint32_t lo = 0, hi = ct.n;
    while (hi - lo > 1) {
        const uint32_t m = (lo + hi) >> 1;
        if (ct.start[m] <= begin) lo = m; else hi = m;
    }
    uint64_t cstart = ct.start[lo], cend = ct.start[lo + 1];
    uint64_t p = begin - cstart < 32 ? cstart : begin - 32; // 32 bases of history are all the state there is
    uint64_t fwd = 0, rc = 0;
    uint32_t charge = 0;
    for (; p < end; ++p) {
        if (p == cend) { // next contig
            ++lo;
            cstart = cend;
            cend = ct.start[lo + 1];
            fwd = rc = 0;
            charge = 0;
        }
        const uint8_t c = seq[p];
        if (c == 'N') { fwd = rc = 0; charge = 0; continue; }
        const uint64_t code = (c >> 1) & 3;
        fwd = (fwd << 2) | code;
        rc = (rc | (((code - 2) & 3) << 60)) >> 2;
        if (charge < k) ++charge;
        if (p < begin || charge < k) continue;
        uint64_t km = fwd & kmask;
        if (km > rc) km = rc;
        if (km != 0) f(p, km);
    }
}

// pass 1: occurrence table.  word = key (< 2^60) | QS_MULTI once the key has been seen twice.
__global__ void qs_occ_kernel(const uint8_t *__restrict__ seq, qs_contigs ct, uint64_t n, uint32_t k, uint64_t kmask,
                              unsigned long long *__restrict__ table, uint64_t tmask, unsigned int *full)
{
    const uint64_t run = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t begin = run * QS_RUN;
    if (begin >= n) return;
    const uint64_t end = begin + QS_RUN < n ? begin + QS_RUN : n;
    qs_for_each_kmer(seq, ct, begin, end, k, kmask, [&](uint64_t, uint64_t key) {
        uint64_t h = qs_mix(key) & tmask;
        for (uint64_t tries = 0; tries <= tmask; ++tries) {
            unsigned long long cur = table[h];
            if (cur == 0) {
                cur = atomicCAS(&table[h], 0ull, (unsigned long long)key);
                if (cur == 0) return;
            }
            if ((cur & QS_M60) == key) {
                if (!(cur & QS_MULTI)) atomicOr(&table[h], QS_MULTI);
                return;
            }
            h = (h + 1) & tmask;
        }
        atomicExch(full, 1u);
    });
}

__device__ __forceinline__ bool qs_is_unique(const unsigned long long *__restrict__ table, uint64_t tmask, uint64_t key)
{
    uint64_t h = qs_mix(key) & tmask;
    for (;;) {
        const unsigned long long cur = table[h];
        if ((cur & QS_M60) == key) return !(cur & QS_MULTI);
        if (cur == 0) return false; // cannot happen
        h = (h + 1) & tmask;
    }
}

// pass 2a: number of unique k-mers ending in each run
__global__ void qs_unique_count_kernel(const uint8_t *__restrict__ seq, qs_contigs ct, uint64_t n, uint32_t k, uint64_t kmask,
                                       const unsigned long long *__restrict__ table, uint64_t tmask,
                                       unsigned long long *__restrict__ run_count, uint32_t *__restrict__ uniq_bits)
{
    const uint64_t run = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t begin = run * QS_RUN;
    if (begin >= n) return;
    const uint64_t end = begin + QS_RUN < n ? begin + QS_RUN : n;
    uint32_t bits[QS_RUN / 32];
#pragma unroll
    for (int i = 0; i < QS_RUN / 32; ++i) bits[i] = 0;
    uint32_t cnt = 0;
    qs_for_each_kmer(seq, ct, begin, end, k, kmask, [&](uint64_t p, uint64_t key) {
        if (qs_is_unique(table, tmask, key)) {
            const uint32_t i = (uint32_t)(p - begin);
#pragma unroll
            for (int w = 0; w < QS_RUN / 32; ++w)
                if ((i >> 5) == (uint32_t)w) bits[w] |= 1u << (i & 31);
            ++cnt;
        }
    });
    run_count[run] = cnt;
#pragma unroll
    for (int i = 0; i < QS_RUN / 32; ++i) uniq_bits[run * (QS_RUN / 32) + i] = bits[i];
}

__host__ __device__ __forceinline__ uint64_t qs_djb(uint64_t key) // Q.c:66-76
{
    uint64_t h = 5381;
    for (int b = 0; b < 8; ++b, key >>= 8) h = h * 33u + (key & 0xFFu);
    return h;
}

// pass 2b: place the unique k-mers (Q.c:90-99 probe rule) and record slot + .qgc word by ordinal
__global__ void qs_place_kernel(const uint8_t *__restrict__ seq, qs_contigs ct, uint64_t n, uint32_t k, uint64_t kmask,
                                const uint32_t *__restrict__ uniq_bits, const unsigned long long *__restrict__ run_base,
                                unsigned long long *__restrict__ keys, uint64_t slots, uint32_t *__restrict__ slot_by_ord,
                                uint16_t *__restrict__ qgc, uint64_t ctrl_block, unsigned int *full)
{
    const uint64_t run = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t begin = run * QS_RUN;
    if (begin >= n) return;
    const uint64_t end = begin + QS_RUN < n ? begin + QS_RUN : n;
    bool any = false;
#pragma unroll
    for (int i = 0; i < QS_RUN / 32; ++i) any = any || uniq_bits[run * (QS_RUN / 32) + i] != 0;
    if (!any) return;
    uint64_t ord = run_base[run];
    const int64_t half_lead = (400 - (int64_t)k) / 2, half_trail = (400 + (int64_t)k) / 2;
    // GC count of the window (r - half_trail, r + half_lead] of the contig, r = contig-relative end position
    uint32_t gc = 0;
    uint64_t gc_for = UINT64_MAX, gc_cstart = 0, gc_cend = 0; // position the running window belongs to
    qs_for_each_kmer(seq, ct, begin, end, k, kmask, [&](uint64_t p, uint64_t key) {
        if (!((uniq_bits[run * (QS_RUN / 32) + ((p - begin) >> 5)] >> ((p - begin) & 31)) & 1u)) return;
        if (gc_for == UINT64_MAX || p < gc_cstart || p >= gc_cend || p - gc_for > 64) {
            uint32_t lo = 0, hi = ct.n;
            while (hi - lo > 1) {
                const uint32_t m = (lo + hi) >> 1;
                if (ct.start[m] <= p) lo = m; else hi = m;
            }
            gc_cstart = ct.start[lo];
            gc_cend = ct.start[lo + 1];
            int64_t a = (int64_t)(p - gc_cstart) - half_trail + 1, b = (int64_t)(p - gc_cstart) + half_lead;
            if (a < 0) a = 0;
            if (b >= (int64_t)(gc_cend - gc_cstart)) b = (int64_t)(gc_cend - gc_cstart) - 1;
            gc = 0;
            for (int64_t i = a; i <= b; ++i) {
                const uint8_t c = seq[gc_cstart + i];
                gc += (c == 'G' || c == 'C');
            }
        } else {
            for (uint64_t q = gc_for + 1; q <= p; ++q) { // slide the window from q - 1 to q
                const int64_t lead = (int64_t)(q - gc_cstart) + half_lead, trail = (int64_t)(q - gc_cstart) - half_trail;
                if (lead < (int64_t)(gc_cend - gc_cstart)) { const uint8_t c = seq[gc_cstart + lead]; gc += (c == 'G' || c == 'C'); }
                if (trail >= 0) { const uint8_t c = seq[gc_cstart + trail]; gc -= (c == 'G' || c == 'C'); }
            }
        }
        gc_for = p;
        uint16_t v = (uint16_t)(gc > 400 ? 400 : gc);
        if (ctrl_block && (((p - gc_cstart) / ctrl_block) & 1)) v |= 0x8000;
        uint64_t s = qs_djb(key) & (slots - 1);
        const long long step = (s & (slots >> 1)) ? -1 : 1;
        for (;;) {
            if (s >= slots) { atomicExch(full, 1u); break; }
            if (keys[s] == 0 && atomicCAS(&keys[s], 0ull, (unsigned long long)key) == 0ull) break;
            s = (uint64_t)((long long)s + step);
        }
        slot_by_ord[ord] = (uint32_t)s;
        if (qgc) qgc[ord] = v;
        ++ord;
    });
}

__global__ void qs_chain_kernel(const uint32_t *__restrict__ slot_by_ord, uint64_t n_unique, uint32_t *__restrict__ next)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_unique; i += (uint64_t)gridDim.x * blockDim.x)
        next[slot_by_ord[i]] = slot_by_ord[i + 1 == n_unique ? 0 : i + 1];
}

struct qs_dict_info {
    uint64_t slots, n_unique, first, n_distinct_slots_t1;
    uint32_t k, pad;
    double occ_s, select_s, place_s;
};

static double qs_now(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + ts.tv_nsec * 1e-9;
}

extern "C" int qs_dict_build(qs_genome *g, uint32_t k, uint64_t slots, uint64_t ctrl_block, int with_qgc, qs_dict_info *info)
{
    if (!g || k < 1 || k > 31) return QS_ERR_ARG;
    QS_CUDA(g, cudaSetDevice(g->device));
    qs_dict_release(g);
    const uint64_t n = g->n;
    const uint64_t kmask = ((uint64_t)1 << (2 * k)) - 1;
    const uint64_t n_runs = (n + QS_RUN - 1) / QS_RUN;
    const unsigned grid = (unsigned)((n_runs + 127) / 128);
    uint64_t tsize = 1;
    while (tsize < n + n / 3 + 16) tsize <<= 1; // <= 75 % full even if every k-mer is distinct
    unsigned long long *t1 = NULL, *run_count = NULL, *run_base = NULL;
    uint32_t *uniq_bits = NULL, *slot_by_ord = NULL;
    unsigned int *full = NULL;
    void *scan_tmp = NULL;
    size_t scan_bytes = 0;
    int rc = QS_OK;
    unsigned int hfull = 0;
    unsigned long long last_base = 0, last_count = 0;
    uint64_t n_unique = 0;
    double t0 = qs_now(), t1s, t2s, t3s;
#define QS_TRY(call)                                                                                              \
    do {                                                                                                          \
        cudaError_t e__ = (call);                                                                                 \
        if (e__ != cudaSuccess) { rc = qs_fail(g, e__ == cudaErrorMemoryAllocation ? QS_ERR_NOMEM : QS_ERR_CUDA, #call, e__); goto done; } \
    } while (0)
    QS_TRY(cudaMalloc((void **)&full, sizeof *full));
    QS_TRY(cudaMemset(full, 0, sizeof *full));
    QS_TRY(cudaMalloc((void **)&t1, tsize * 8));
    QS_TRY(cudaMemset(t1, 0, tsize * 8));
    QS_TRY(cudaMalloc((void **)&run_count, (n_runs + 1) * 8));
    QS_TRY(cudaMalloc((void **)&run_base, (n_runs + 1) * 8));
    QS_TRY(cudaMalloc((void **)&uniq_bits, n_runs * (QS_RUN / 32) * 4));
    qs_occ_kernel<<<grid, 128>>>(g->seq, g->contigs, n, k, kmask, t1, tsize - 1, full);
    QS_TRY(cudaGetLastError());
    QS_TRY(cudaDeviceSynchronize());
    t1s = qs_now();
    qs_unique_count_kernel<<<grid, 128>>>(g->seq, g->contigs, n, k, kmask, t1, tsize - 1, run_count, uniq_bits);
    QS_TRY(cudaGetLastError());
    QS_TRY(cub::DeviceScan::ExclusiveSum(NULL, scan_bytes, run_count, run_base, n_runs));
    QS_TRY(cudaMalloc(&scan_tmp, scan_bytes));
    QS_TRY(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, run_count, run_base, n_runs));
    QS_TRY(cudaMemcpy(&last_base, run_base + n_runs - 1, 8, cudaMemcpyDeviceToHost));
    QS_TRY(cudaMemcpy(&last_count, run_count + n_runs - 1, 8, cudaMemcpyDeviceToHost));
    QS_TRY(cudaMemcpy(&hfull, full, sizeof hfull, cudaMemcpyDeviceToHost));
    if (hfull) { rc = qs_fail(g, QS_ERR_FULL, "occurrence table full", cudaSuccess); goto done; }
    n_unique = last_base + last_count;
    cudaFree(t1); t1 = NULL;
    cudaFree(run_count); run_count = NULL;
    cudaFree(scan_tmp); scan_tmp = NULL;
    t2s = qs_now();
    if (!slots) { slots = 1; while (slots < 2 * n_unique + 2) slots <<= 1; }
    else { uint64_t p2 = 1; while (p2 < slots) p2 <<= 1; slots = p2; }
    if (slots > ((uint64_t)1 << 32) || n_unique == 0 || n_unique > slots / 10 * 8) {
        snprintf(g->err, sizeof g->err, "%llu unique k-mers do not fit %llu slots", (unsigned long long)n_unique,
                 (unsigned long long)slots);
        rc = QS_ERR_FULL;
        goto done;
    }
    QS_TRY(cudaMalloc((void **)&g->keys, slots * 8));
    QS_TRY(cudaMemset(g->keys, 0, slots * 8));
    QS_TRY(cudaMalloc((void **)&slot_by_ord, (n_unique + 1) * 4));
    if (with_qgc) QS_TRY(cudaMalloc((void **)&g->qgc, (n_unique + 1) * 2));
    qs_place_kernel<<<grid, 128>>>(g->seq, g->contigs, n, k, kmask, uniq_bits, run_base, (unsigned long long *)g->keys, slots,
                                   slot_by_ord, g->qgc, ctrl_block, full);
    QS_TRY(cudaGetLastError());
    QS_TRY(cudaMemcpy(&hfull, full, sizeof hfull, cudaMemcpyDeviceToHost));
    if (hfull) { rc = qs_fail(g, QS_ERR_FULL, "probe walked off the table", cudaSuccess); goto done; }
    cudaFree(uniq_bits); uniq_bits = NULL;
    cudaFree(run_base); run_base = NULL;
    QS_TRY(cudaMalloc((void **)&g->next, slots * 4));
    QS_TRY(cudaMemset(g->next, 0, slots * 4));
    qs_chain_kernel<<<148 * 16, 256>>>(slot_by_ord, n_unique, g->next);
    QS_TRY(cudaGetLastError());
    {
        uint32_t first32 = 0;
        QS_TRY(cudaMemcpy(&first32, slot_by_ord, 4, cudaMemcpyDeviceToHost));
        g->first = first32;
    }
    t3s = qs_now();
    g->slots = slots;
    g->n_unique = n_unique;
    g->k = k;
    if (info) {
        memset(info, 0, sizeof *info);
        info->slots = slots;
        info->n_unique = n_unique;
        info->first = g->first;
        info->n_distinct_slots_t1 = tsize;
        info->k = k;
        info->occ_s = t1s - t0;
        info->select_s = t2s - t1s;
        info->place_s = t3s - t2s;
    }
done:
    cudaFree(full);
    cudaFree(t1);
    cudaFree(run_count);
    cudaFree(run_base);
    cudaFree(uniq_bits);
    cudaFree(slot_by_ord);
    cudaFree(scan_tmp);
    if (rc) qs_dict_release(g);
    return rc;
#undef QS_TRY
}

// keys[slots], next[slots], qgc[n_unique] to host arrays (any may be NULL)
extern "C" int qs_dict_download(qs_genome *g, uint64_t *keys, uint32_t *next, uint16_t *qgc)
{
    if (!g || !g->keys) return QS_ERR_ARG;
    QS_CUDA(g, cudaSetDevice(g->device));
    if (keys) QS_CUDA(g, cudaMemcpy(keys, g->keys, g->slots * 8, cudaMemcpyDeviceToHost));
    if (next) QS_CUDA(g, cudaMemcpy(next, g->next, g->slots * 4, cudaMemcpyDeviceToHost));
    if (qgc && g->qgc) QS_CUDA(g, cudaMemcpy(qgc, g->qgc, g->n_unique * 2, cudaMemcpyDeviceToHost));
    return QS_OK;
}

// device array -> file at `file_off`, through two pinned staging buffers; pwrite by several threads
static int qs_write_array(qs_genome *g, int fd, uint64_t file_off, const void *dev, uint64_t bytes, uint8_t *stage[2],
                          size_t stage_bytes, cudaStream_t st, int threads)
{
    cudaEvent_t ev[2];
    QS_CUDA(g, cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    QS_CUDA(g, cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
    int rc = QS_OK, b = 0;
    uint64_t prev_at = 0, prev_m = 0;
    for (uint64_t at = 0; !rc; at += stage_bytes, b ^= 1) {
        const uint64_t m = at < bytes ? (bytes - at < stage_bytes ? bytes - at : stage_bytes) : 0;
        if (m) {
            cudaError_t e = cudaMemcpyAsync(stage[b], (const uint8_t *)dev + at, m, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaEventRecord(ev[b], st);
            if (e != cudaSuccess) { rc = qs_fail(g, QS_ERR_CUDA, "D2H", e); break; }
        }
        if (prev_m) {
            cudaError_t e = cudaEventSynchronize(ev[b ^ 1]);
            if (e != cudaSuccess) { rc = qs_fail(g, QS_ERR_CUDA, "D2H", e); break; }
            const uint8_t *src = stage[b ^ 1];
            const uint64_t per = (prev_m + threads - 1) / threads;
            int bad = 0;
#pragma omp parallel for num_threads(threads) reduction(| : bad)
            for (int t = 0; t < threads; ++t) {
                uint64_t a = (uint64_t)t * per, z = a + per < prev_m ? a + per : prev_m;
                while (a < z) {
                    ssize_t w = pwrite(fd, src + a, z - a, (off_t)(file_off + prev_at + a));
                    if (w <= 0) { bad = 1; break; }
                    a += (uint64_t)w;
                }
            }
            if (bad) rc = qs_fail(g, QS_ERR_IO, "pwrite failed", cudaSuccess);
        }
        prev_at = at;
        prev_m = m;
        if (!m) break;
    }
    cudaEventDestroy(ev[0]);
    cudaEventDestroy(ev[1]);
    return rc;
}

// <prefix>.qm (QM11: header, keys, chain; Q.c:1284-1299) and, if built, <prefix>.qgc
extern "C" int qs_dict_write(qs_genome *g, const char *prefix, int threads)
{
    if (!g || !g->keys || !g->next || !prefix) return QS_ERR_ARG;
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    QS_CUDA(g, cudaSetDevice(g->device));
    const size_t stage_bytes = (size_t)256 << 20;
    uint8_t *stage[2] = {NULL, NULL};
    cudaStream_t st;
    QS_CUDA(g, cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    QS_CUDA(g, cudaHostAlloc((void **)&stage[0], stage_bytes, cudaHostAllocDefault));
    QS_CUDA(g, cudaHostAlloc((void **)&stage[1], stage_bytes, cudaHostAllocDefault));
    char path[4096];
    snprintf(path, sizeof path, "%s.qm", prefix);
    int fd = open(path, O_CREAT | O_TRUNC | O_WRONLY, 0644);
    int rc = fd < 0 ? qs_fail(g, QS_ERR_IO, path, cudaSuccess) : QS_OK;
    if (!rc) {
        uint8_t hdr[24] = {'Q', 'M', '1', '1', (uint8_t)g->k, 0, 100, 100};
        memcpy(hdr + 8, &g->slots, 8);
        memcpy(hdr + 16, &g->first, 8);
        if (pwrite(fd, hdr, 24, 0) != 24) rc = qs_fail(g, QS_ERR_IO, "header", cudaSuccess);
        if (!rc) rc = qs_write_array(g, fd, 24, g->keys, g->slots * 8, stage, stage_bytes, st, threads);
        if (!rc) rc = qs_write_array(g, fd, 24 + g->slots * 8, g->next, g->slots * 4, stage, stage_bytes, st, threads);
        close(fd);
    }
    if (!rc && g->qgc) {
        snprintf(path, sizeof path, "%s.qgc", prefix);
        fd = open(path, O_CREAT | O_TRUNC | O_WRONLY, 0644);
        if (fd < 0) rc = qs_fail(g, QS_ERR_IO, path, cudaSuccess);
        else {
            rc = qs_write_array(g, fd, 0, g->qgc, g->n_unique * 2, stage, stage_bytes, st, threads);
            close(fd);
        }
    }
    cudaFreeHost(stage[0]);
    cudaFreeHost(stage[1]);
    cudaStreamDestroy(st);
    return rc;
}

// ------------------------------------------------------------------------- reads ---------
// format: 0 = framed (sequence line only), 1 = FASTA (">r%010llu\n" + sequence line),
//         2 = FASTQ ("@r%010llu\n" + sequence + "\n+\n" + quality 'I'... + "\n")
struct qs_reads_params {
    uint64_t seed, first_read, n_reads;
    uint32_t len;           // fixed length (when lens == NULL)
    uint32_t err_threshold; // substitution iff 14 bits < this
    int format;
    uint64_t genome_n;
};

__host__ __device__ __forceinline__ uint64_t qs_record_bytes(uint32_t len, int format)
{
    return format == 0 ? (uint64_t)len + 1 : format == 1 ? 13 + (uint64_t)len + 1 : 13 + 2 * ((uint64_t)len + 1) + 2;
}

// where read i starts in the genome and how long it really is (clipped to its contig)
__device__ __forceinline__ void qs_read_place(const qs_contigs &ct, uint64_t genome_n, uint64_t seed, uint64_t i, uint32_t want,
                                              uint64_t *at, uint32_t *len)
{
    uint64_t pos = qs_below(qs_hash(seed, i, 0), genome_n);
    uint32_t lo = 0, hi = ct.n;
    while (hi - lo > 1) {
        const uint32_t m = (lo + hi) >> 1;
        if (ct.start[m] <= pos) lo = m; else hi = m;
    }
    const uint64_t clen = ct.start[lo + 1] - ct.start[lo];
    uint32_t L = want;
    if (clen < L) L = (uint32_t)clen;
    *at = ct.start[lo] + qs_below(qs_hash(seed, i, 1), clen - L + 1);
    *len = L;
}

__device__ __forceinline__ uint8_t qs_read_base(const uint8_t *__restrict__ seq, uint64_t seed, uint64_t i, uint64_t at, uint32_t L,
                                                uint32_t j, uint32_t err_threshold)
{
    uint8_t c;
    if (i & 1) { // reverse complement
        c = seq[at + L - 1 - j];
        c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c;
    } else c = seq[at + j];
    if (err_threshold && c != 'N') {
        const uint32_t u = (uint32_t)(qs_hash(seed, i, 16 + (j >> 2)) >> (16 * (j & 3))) & 0xFFFFu;
        if ((u >> 2) < err_threshold) {
            const uint32_t code = (c >> 1) & 3;
            c = "ACTG"[(code + 1 + (uint32_t)(qs_hash(seed ^ 0x5EEDull, i, j) % 3)) & 3];
        }
    }
    return c;
}

__device__ __forceinline__ uint8_t qs_record_byte(const uint8_t *__restrict__ seq, const qs_reads_params &rp, uint64_t i, uint64_t at,
                                                  uint32_t L, uint32_t pad_len, uint64_t b)
{
    // pad_len = nominal length of the record (reads clipped by a short contig are padded with 'N')
    const int format = rp.format;
    if (format != 0) {
        if (b < 13) {
            if (b == 0) return format == 1 ? '>' : '@';
            if (b == 1) return 'r';
            if (b == 12) return '\n';
            uint64_t v = i;
            for (uint32_t d = 11; d > b; --d) v /= 10;
            return (uint8_t)('0' + v % 10);
        }
        b -= 13;
    }
    if (b < pad_len) return b < L ? qs_read_base(seq, rp.seed, i, at, L, (uint32_t)b, rp.err_threshold) : 'N';
    if (b == pad_len) return '\n';
    b -= (uint64_t)pad_len + 1;
    if (b == 0) return '+';
    if (b == 1) return '\n';
    b -= 2;
    return b < pad_len ? 'I' : '\n';
}

// The bytes of one record, written by `nthr` cooperating threads (thread t of them): the bases four at a
// time (one hash decides the substitutions of four bases), the rest byte by byte.
__device__ __forceinline__ void qs_write_record(const uint8_t *__restrict__ seq, const qs_reads_params &rp, uint64_t i, uint64_t at,
                                                uint32_t L, uint32_t pad_len, uint8_t *__restrict__ dst, uint32_t t, uint32_t nthr)
{
    const int format = rp.format;
    const uint32_t hdr = format ? 13u : 0u;
    for (uint32_t b = t; b < hdr; b += nthr) dst[b] = qs_record_byte(seq, rp, i, at, L, pad_len, b);
    uint8_t *sq = dst + hdr;
    const bool rc = (i & 1) != 0;
    for (uint32_t g = t; 4 * g < pad_len; g += nthr) {
        const uint64_t h = rp.err_threshold ? qs_hash(rp.seed, i, 16 + g) : ~0ull;
#pragma unroll
        for (uint32_t u4 = 0; u4 < 4; ++u4) {
            const uint32_t j = 4 * g + u4;
            if (j >= pad_len) break;
            uint8_t c = 'N';
            if (j < L) {
                if (rc) {
                    c = seq[at + L - 1 - j];
                    c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c;
                } else c = seq[at + j];
                const uint32_t u = (uint32_t)(h >> (16 * u4)) & 0xFFFFu;
                if (rp.err_threshold && c != 'N' && (u >> 2) < rp.err_threshold) {
                    const uint32_t code = (c >> 1) & 3;
                    c = "ACTG"[(code + 1 + (uint32_t)(qs_hash(rp.seed ^ 0x5EEDull, i, j) % 3)) & 3];
                }
            }
            sq[j] = c;
        }
    }
    if (t == 0) sq[pad_len] = '\n';
    if (format == 2) {
        uint8_t *q = sq + pad_len + 1;
        for (uint32_t b = t; b < pad_len + 3; b += nthr) q[b] = b == 0 ? '+' : b == 1 ? '\n' : b == pad_len + 2 ? '\n' : 'I';
    }
}

// fixed-length reads: a warp per read
__global__ void qs_reads_fixed_kernel(const uint8_t *__restrict__ seq, qs_contigs ct, qs_reads_params rp, uint8_t *__restrict__ out)
{
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t rec = qs_record_bytes(rp.len, rp.format);
    for (uint64_t r = warp; r < rp.n_reads; r += n_warps) {
        const uint64_t i = rp.first_read + r;
        uint64_t at = 0;
        uint32_t L = 0;
        if (lane == 0) qs_read_place(ct, rp.genome_n, rp.seed, i, rp.len, &at, &L);
        at = __shfl_sync(0xffffffffu, at, 0);
        L = __shfl_sync(0xffffffffu, L, 0);
        qs_write_record(seq, rp, i, at, L, rp.len, out + r * rec, lane, 32);
    }
}

// caller-given lengths (HiFi): a CTA per read; offsets[r] = where record r starts in `out`
__global__ void qs_reads_var_kernel(const uint8_t *__restrict__ seq, qs_contigs ct, qs_reads_params rp, const uint32_t *__restrict__ lens,
                                    const uint64_t *__restrict__ offsets, uint8_t *__restrict__ out)
{
    for (uint64_t r = blockIdx.x; r < rp.n_reads; r += gridDim.x) {
        const uint64_t i = rp.first_read + r;
        uint64_t at;
        uint32_t L;
        qs_read_place(ct, rp.genome_n, rp.seed, i, lens[r], &at, &L);
        qs_write_record(seq, rp, i, at, L, lens[r], out + offsets[r], threadIdx.x, blockDim.x);
    }
}

extern "C" uint64_t qs_reads_record_bytes(uint32_t len, int format) { return qs_record_bytes(len, format); }

// Reads [first_read, first_read + n_reads) of stream `seed` into device memory at dev_out.
// lens == NULL: all `len` long, records back to back (n_reads * qs_reads_record_bytes(len, format) bytes).
// lens != NULL (host array of n_reads): record r starts at offsets[r] (host array, bytes).
extern "C" int qs_reads_generate(qs_genome *g, uint64_t seed, uint64_t first_read, uint64_t n_reads, uint32_t len, uint32_t err_ppm,
                                 int format, const uint32_t *lens, const uint64_t *offsets, uint8_t *dev_out)
{
    if (!g || !dev_out || format < 0 || format > 2 || (lens && !offsets)) return QS_ERR_ARG;
    if (n_reads == 0) return QS_OK;
    QS_CUDA(g, cudaSetDevice(g->device));
    qs_reads_params rp;
    rp.seed = seed;
    rp.first_read = first_read;
    rp.n_reads = n_reads;
    rp.len = len;
    rp.err_threshold = (uint32_t)((double)err_ppm * 1e-6 * 16384.0 + 0.5);
    rp.format = format;
    rp.genome_n = g->n;
    if (!lens) {
        qs_reads_fixed_kernel<<<148 * 16, 256>>>(g->seq, g->contigs, rp, dev_out);
    } else {
        uint32_t *dl = NULL;
        uint64_t *doff = NULL;
        QS_CUDA(g, cudaMalloc((void **)&dl, n_reads * 4));
        QS_CUDA(g, cudaMalloc((void **)&doff, n_reads * 8));
        QS_CUDA(g, cudaMemcpy(dl, lens, n_reads * 4, cudaMemcpyHostToDevice));
        QS_CUDA(g, cudaMemcpy(doff, offsets, n_reads * 8, cudaMemcpyHostToDevice));
        qs_reads_var_kernel<<<148 * 8, 256>>>(g->seq, g->contigs, rp, dl, doff, dev_out);
        cudaDeviceSynchronize();
        cudaFree(dl);
        cudaFree(doff);
    }
    QS_CUDA(g, cudaGetLastError());
    QS_CUDA(g, cudaDeviceSynchronize());
    return QS_OK;
}

// plain device memory for the callers (ctypes has no cudaMalloc): bytes, returns NULL on failure
extern "C" void *qs_device_alloc(int device, size_t bytes)
{
    void *p = NULL;
    if (cudaSetDevice(device) != cudaSuccess || cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return NULL; }
    return p;
}
extern "C" void qs_device_free(void *p) { cudaFree(p); }
extern "C" void *qs_pinned_alloc(size_t bytes)
{
    void *p = NULL;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return NULL; }
    return p;
}
extern "C" void qs_pinned_free(void *p) { cudaFreeHost(p); }
extern "C" int qs_copy_to_host(void *host, const void *dev, size_t bytes)
{
    return cudaMemcpy(host, dev, bytes, cudaMemcpyDeviceToHost) == cudaSuccess ? QS_OK : QS_ERR_CUDA;
}
extern "C" int qs_device_mem_info(int device, uint64_t *free_bytes, uint64_t *total_bytes)
{
    size_t f = 0, t = 0;
    if (cudaSetDevice(device) != cudaSuccess || cudaMemGetInfo(&f, &t) != cudaSuccess) return QS_ERR_CUDA;
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return QS_OK;
}
