// qk_ctx.cu -- context, chunk slots (pinned host + device buffers, streams, events),
// error reporting, timing ring, roofline micro-benchmarks.
//
// The slots replace the reference's per-worker double FIFO and semaphores
// (struct FIFO_arg_struc Q.c:34-41, thread pool Q.c:368-384, hand-off Q.c:421-438):
// "post" becomes a stream-ordered H2D copy + kernel launch, "idle worker" becomes an
// event that says the pinned buffer may be refilled.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "qk_common.cuh"

// Reader and framer threads report into the same context: the message is formatted aside and
// copied in under a lock, so that two failures at once leave one whole message, not a blend.
static std::mutex g_err_lock;

int qk_fail(qk_ctx *ctx, int code, const char *fmt, ...)
{
    if (ctx) {
        char msg[sizeof ctx->err];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(msg, sizeof msg, fmt, ap);
        va_end(ap);
        std::lock_guard<std::mutex> hold(g_err_lock);
        memcpy(ctx->err, msg, sizeof msg);
    }
    return code;
}

int qk_cuda_fail(qk_ctx *ctx, cudaError_t e, const char *what)
{
    return qk_fail(ctx, e == cudaErrorMemoryAllocation ? QK_ERR_NOMEM : QK_ERR_CUDA, "%s: %s", what,
                   cudaGetErrorString(e));
}

extern "C" const char *qk_version(void) { return "quickmer2_b200 0.1 (sm_100a)"; }

extern "C" int qk_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return -QK_ERR_CUDA; }
    return n;
}

extern "C" const char *qk_last_error(const qk_ctx *ctx) { return ctx ? ctx->err : "no context"; }

extern "C" int qk_ctx_create(qk_ctx **out, int device, uint32_t n_slots, size_t chunk_capacity)
{
    if (!out) return QK_ERR_ARG;
    *out = NULL;
    if (n_slots < 1 || n_slots > QK_MAX_SLOTS) return QK_ERR_ARG;
    if (chunk_capacity < 2 * 100000 || chunk_capacity >= ((size_t)1 << 31)) return QK_ERR_ARG;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0 || device < 0 || device >= n) {
        cudaGetLastError();
        fprintf(stderr, "quickmer2_b200: no usable CUDA device %d (%s); there is no CPU fallback\n", device,
                e != cudaSuccess ? cudaGetErrorString(e) : "device index out of range");
        return QK_ERR_CUDA;
    }
    qk_ctx *ctx = (qk_ctx *)calloc(1, sizeof(qk_ctx));
    if (!ctx) return QK_ERR_NOMEM;
    ctx->device = device;
    ctx->n_slots = n_slots;
    // round the capacity up to whole tiles so vector loads of the last tile stay in bounds
    ctx->chunk_capacity = (chunk_capacity + QK_TILE - 1) / QK_TILE * QK_TILE;
    *out = ctx; // returned even on failure below so the caller can read the message
    QK_CUDA(ctx, cudaSetDevice(device));
    QK_CUDA(ctx, cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device));
    for (uint32_t s = 0; s < n_slots; ++s) {
        qk_slot *sl = &ctx->slots[s];
        QK_CUDA(ctx, cudaHostAlloc((void **)&sl->host, ctx->chunk_capacity, cudaHostAllocDefault));
        QK_CUDA(ctx, cudaMalloc((void **)&sl->dev, ctx->chunk_capacity));
        QK_CUDA(ctx, cudaStreamCreateWithFlags(&sl->stream, cudaStreamNonBlocking));
        QK_CUDA(ctx, cudaEventCreateWithFlags(&sl->h2d_done, cudaEventDisableTiming));
        QK_CUDA(ctx, cudaEventRecord(sl->h2d_done, sl->stream));
        QK_CUDA(ctx, cudaEventCreateWithFlags(&sl->frame_done, cudaEventDisableTiming));
        for (int i = 0; i < QK_TIMING_RING; ++i) {
            QK_CUDA(ctx, cudaEventCreate(&sl->ring[i].a));
            QK_CUDA(ctx, cudaEventCreate(&sl->ring[i].b));
        }
    }
    QK_CUDA(ctx, cudaEventCreate(&ctx->span_a));
    QK_CUDA(ctx, cudaEventCreate(&ctx->span_b));
    QK_CUDA(ctx, cudaEventCreateWithFlags(&ctx->span_join, cudaEventDisableTiming));
    QK_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->finish_stream, cudaStreamNonBlocking));
    QK_CUDA(ctx, cudaEventCreateWithFlags(&ctx->finish_done, cudaEventDisableTiming));
    QK_CUDA(ctx, cudaMalloc((void **)&ctx->stats, QK_STATS_WORDS * sizeof(unsigned long long)));
    QK_CUDA(ctx, cudaMemset(ctx->stats, 0, QK_STATS_WORDS * sizeof(unsigned long long)));
    QK_CUDA(ctx, cudaMalloc((void **)&ctx->frame_stream, QK_STATS_WORDS * sizeof(unsigned long long)));
    QK_CUDA(ctx, cudaMemset(ctx->frame_stream, 0, QK_STATS_WORDS * sizeof(unsigned long long)));
    QK_CUDA(ctx, cudaMalloc((void **)&ctx->frame_elems, (size_t)n_slots * QK_FRAME_MAX_CTAS * sizeof(uint32_t)));
    ctx->raw_prev_slot = -1;
    return QK_OK;
}

extern "C" void qk_ctx_destroy(qk_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (uint32_t s = 0; s < ctx->n_slots; ++s) {
        qk_slot *sl = &ctx->slots[s];
        if (sl->host) cudaFreeHost(sl->host);
        if (sl->dev) cudaFree(sl->dev);
        if (sl->stream) cudaStreamDestroy(sl->stream);
        if (sl->h2d_done) cudaEventDestroy(sl->h2d_done);
        if (sl->frame_done) cudaEventDestroy(sl->frame_done);
        for (int i = 0; i < QK_TIMING_RING; ++i) {
            if (sl->ring[i].a) cudaEventDestroy(sl->ring[i].a);
            if (sl->ring[i].b) cudaEventDestroy(sl->ring[i].b);
        }
    }
    if (ctx->span_a) cudaEventDestroy(ctx->span_a);
    if (ctx->span_b) cudaEventDestroy(ctx->span_b);
    if (ctx->span_join) cudaEventDestroy(ctx->span_join);
    if (ctx->finish_done) cudaEventDestroy(ctx->finish_done);
    if (ctx->finish_stream) cudaStreamDestroy(ctx->finish_stream);
    cudaFree(ctx->raw_keys);
    cudaFree(ctx->raw_next);
    cudaFree(ctx->buckets);
    cudaFree(ctx->stash);
    cudaFree(ctx->ext);
    cudaFree(ctx->counters_buf[0]);
    cudaFree(ctx->counters_buf[1]);
    cudaFree(ctx->stats);
    cudaFree(ctx->frame_stream);
    cudaFree(ctx->frame_elems);
    cudaFree(ctx->narrow_dev);
    cudaFree(ctx->gc_acc);
    cudaFree(ctx->est_depth);
    cudaFree(ctx->est_qgc);
    if (ctx->narrow_host) cudaFreeHost(ctx->narrow_host);
    cudaGetLastError();
    free(ctx);
}

extern "C" int qk_ctx_info(const qk_ctx *ctx, uint32_t *n_slots, size_t *chunk_capacity)
{
    if (!ctx) return QK_ERR_ARG;
    if (n_slots) *n_slots = ctx->n_slots;
    if (chunk_capacity) *chunk_capacity = ctx->chunk_capacity;
    return QK_OK;
}

extern "C" uint8_t *qk_slot_host_buffer(qk_ctx *ctx, uint32_t slot)
{
    if (!ctx || slot >= ctx->n_slots) return NULL;
    return ctx->slots[slot].host;
}

// ---- timing ring: event pairs recorded around copies and kernels, harvested lazily ----
static int qk_ring_pop(qk_ctx *ctx, qk_slot *sl)
{
    qk_timing_pair *p = &sl->ring[sl->ring_head];
    QK_CUDA(ctx, cudaEventSynchronize(p->b));
    float ms = 0;
    QK_CUDA(ctx, cudaEventElapsedTime(&ms, p->a, p->b));
    if (p->kind) ctx->kernel_ms += ms; else ctx->h2d_ms += ms;
    sl->ring_head = (sl->ring_head + 1) % QK_TIMING_RING;
    sl->ring_count--;
    return QK_OK;
}

int qk_ring_push(qk_ctx *ctx, qk_slot *sl, int kind, qk_timing_pair **out)
{
    if (sl->ring_count == QK_TIMING_RING) {
        int rc = qk_ring_pop(ctx, sl);
        if (rc) return rc;
    }
    qk_timing_pair *p = &sl->ring[(sl->ring_head + sl->ring_count) % QK_TIMING_RING];
    p->kind = kind;
    sl->ring_count++;
    *out = p;
    return QK_OK;
}

int qk_ring_drain(qk_ctx *ctx)
{
    for (uint32_t s = 0; s < ctx->n_slots; ++s)
        while (ctx->slots[s].ring_count) {
            int rc = qk_ring_pop(ctx, &ctx->slots[s]);
            if (rc) return rc;
        }
    return QK_OK;
}

extern "C" void *qk_slot_stream(qk_ctx *ctx, uint32_t slot)
{
    if (!ctx || slot >= ctx->n_slots) return NULL;
    return (void *)ctx->slots[slot].stream;
}

extern "C" int qk_wait_slot(qk_ctx *ctx, uint32_t slot)
{
    if (!ctx || slot >= ctx->n_slots) return QK_ERR_ARG;
    // reader / framer threads call this: they have no current device of their own
    QK_CUDA(ctx, cudaSetDevice(ctx->device));
    QK_CUDA(ctx, cudaEventSynchronize(ctx->slots[slot].h2d_done));
    return QK_OK;
}

extern "C" int qk_slot_ready(qk_ctx *ctx, uint32_t slot)
{
    if (!ctx || slot >= ctx->n_slots) return -QK_ERR_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) { cudaGetLastError(); return -QK_ERR_CUDA; }
    const cudaError_t e = cudaEventQuery(ctx->slots[slot].h2d_done);
    if (e == cudaSuccess) return 1;
    if (e == cudaErrorNotReady) { cudaGetLastError(); return 0; }
    return -qk_cuda_fail(ctx, e, "cudaEventQuery");
}

extern "C" int qk_sync(qk_ctx *ctx)
{
    if (!ctx) return QK_ERR_ARG;
    for (uint32_t s = 0; s < ctx->n_slots; ++s) QK_CUDA(ctx, cudaStreamSynchronize(ctx->slots[s].stream));
    return qk_ring_drain(ctx);
}

extern "C" int qk_stats(qk_ctx *ctx, uint64_t *total_kmers, uint64_t *hits, uint64_t *lines)
{
    if (!ctx) return QK_ERR_ARG;
    int rc = qk_sync(ctx);
    if (rc) return rc;
    unsigned long long h[4];
    QK_CUDA(ctx, cudaMemcpy(h, ctx->stats, sizeof h, cudaMemcpyDeviceToHost));
    if (total_kmers) *total_kmers = h[0];
    if (hits) *hits = h[1];
    if (lines) { // framed chunks are counted by the host, raw pieces by the device framer
        unsigned long long f[4];
        QK_CUDA(ctx, cudaMemcpy(f, ctx->frame_stream, sizeof f, cudaMemcpyDeviceToHost));
        *lines = ctx->lines + f[1];
    }
    return QK_OK;
}

extern "C" int qk_stats_ext(qk_ctx *ctx, uint64_t *verified_by_extension)
{
    if (!ctx || !verified_by_extension) return QK_ERR_ARG;
    int rc = qk_sync(ctx);
    if (rc) return rc;
    unsigned long long h[4];
    QK_CUDA(ctx, cudaMemcpy(h, ctx->stats, sizeof h, cudaMemcpyDeviceToHost));
    *verified_by_extension = h[2];
    return QK_OK;
}

extern "C" int qk_stats_probes(qk_ctx *ctx, uint64_t *bucket_probes, uint64_t *walks)
{
    if (!ctx) return QK_ERR_ARG;
    int rc = qk_sync(ctx);
    if (rc) return rc;
    unsigned long long h[QK_STATS_WORDS];
    QK_CUDA(ctx, cudaMemcpy(h, ctx->stats, sizeof h, cudaMemcpyDeviceToHost));
    if (bucket_probes) *bucket_probes = h[4];
    if (walks) *walks = h[6];
    return QK_OK;
}

extern "C" int qk_timing(qk_ctx *ctx, double *kernel_ms, double *h2d_ms, uint64_t *launches)
{
    if (!ctx) return QK_ERR_ARG;
    int rc = qk_sync(ctx);
    if (rc) return rc;
    if (kernel_ms) *kernel_ms = ctx->kernel_ms;
    if (h2d_ms) *h2d_ms = ctx->h2d_ms;
    if (launches) *launches = ctx->launches;
    return QK_OK;
}

// Join every slot stream into slot 0's stream, then record `ev` there.
static int qk_span_mark(qk_ctx *ctx, cudaEvent_t ev)
{
    QK_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s0 = ctx->slots[0].stream;
    for (uint32_t s = 1; s < ctx->n_slots; ++s) {
        QK_CUDA(ctx, cudaEventRecord(ctx->span_join, ctx->slots[s].stream));
        QK_CUDA(ctx, cudaStreamWaitEvent(s0, ctx->span_join, 0));
    }
    QK_CUDA(ctx, cudaEventRecord(ev, s0));
    return QK_OK;
}

extern "C" int qk_span_begin(qk_ctx *ctx)
{
    if (!ctx) return QK_ERR_ARG;
    int rc = qk_span_mark(ctx, ctx->span_a);
    if (rc) return rc;
    // later work on the other slot streams must not start before the mark either
    for (uint32_t s = 1; s < ctx->n_slots; ++s) QK_CUDA(ctx, cudaStreamWaitEvent(ctx->slots[s].stream, ctx->span_a, 0));
    return QK_OK;
}

extern "C" int qk_span_end(qk_ctx *ctx, double *elapsed_ms)
{
    if (!ctx || !elapsed_ms) return QK_ERR_ARG;
    int rc = qk_span_mark(ctx, ctx->span_b);
    if (rc) return rc;
    QK_CUDA(ctx, cudaEventSynchronize(ctx->span_b));
    float ms = 0;
    QK_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->span_a, ctx->span_b));
    *elapsed_ms = ms;
    return QK_OK;
}

// ---- roofline micro-benchmarks --------------------------------------------------------
// Random `GRAN`-byte gathers: each thread draws addresses from a counter-based generator
// (no index array in memory, so the only DRAM traffic is the gathered sectors) and keeps
// MLP independent loads in flight.
__device__ __forceinline__ qk_bucket qk_ld256_stream(const qk_bucket *p)
{
    qk_bucket v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(v.e[0]), "=l"(v.e[1]), "=l"(v.e[2]), "=l"(v.e[3])
                 : "l"(p));
    return v;
}

template <int GRAN, int MLP>
__global__ void __launch_bounds__(256) qk_gather_kernel(const qk_bucket *__restrict__ table, uint64_t n_units,
                                                        uint64_t per_thread, unsigned long long *sink)
{
    constexpr int G = GRAN / 32;
    uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t x = tid * 0x9E3779B97F4A7C15ull + 12345;
    unsigned long long acc = 0;
    for (uint64_t it = 0; it < per_thread; it += MLP) {
        qk_bucket v[MLP][G];
#pragma unroll
        for (int m = 0; m < MLP; ++m) {
            x = x * 6364136223846793005ull + 1442695040888963407ull;
            uint64_t unit = __umul64hi(x ^ (x >> 29), n_units);
#pragma unroll
            for (int g = 0; g < G; ++g) v[m][g] = qk_ld256_stream(table + unit * G + g);
        }
        // ptxas is free to reorder the (non-volatile in PTX) loads and, left alone, issues two,
        // consumes them, issues two more...  The warp barrier pins all MLP loads before any use.
        __syncwarp();
#pragma unroll
        for (int m = MLP - 1; m >= 0; --m)
#pragma unroll
            for (int g = G - 1; g >= 0; --g) // the multiply keeps the chain from being reassociated
                acc = (acc ^ v[m][g].e[0] ^ v[m][g].e[1] ^ v[m][g].e[2] ^ v[m][g].e[3]) * 0x9E3779B97F4A7C15ull;
    }
    if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}

template <int GRAN>
static int qk_gather_dispatch(int mlp, dim3 grid, cudaStream_t st, const qk_bucket *t, uint64_t units, uint64_t per,
                              unsigned long long *sink)
{
    switch (mlp) {
    case 1: qk_gather_kernel<GRAN, 1><<<grid, 256, 0, st>>>(t, units, per, sink); return 1;
    case 2: qk_gather_kernel<GRAN, 2><<<grid, 256, 0, st>>>(t, units, per, sink); return 2;
    case 4: qk_gather_kernel<GRAN, 4><<<grid, 256, 0, st>>>(t, units, per, sink); return 4;
    default: qk_gather_kernel<GRAN, 8><<<grid, 256, 0, st>>>(t, units, per, sink); return 8;
    }
}

extern "C" int qk_bench_gather(qk_ctx *ctx, uint64_t table_bytes, uint32_t gran, uint32_t loads_in_flight,
                               uint64_t n_gathers, double *gbs)
{
    if (!ctx || !gbs || (gran != 32 && gran != 64 && gran != 128) || table_bytes < (1u << 20)) return QK_ERR_ARG;
    QK_CUDA(ctx, cudaSetDevice(ctx->device));
    qk_bucket *table = NULL;
    QK_CUDA(ctx, cudaMalloc((void **)&table, table_bytes));
    QK_CUDA(ctx, cudaMemset(table, 1, table_bytes));
    cudaStream_t st = ctx->slots[0].stream;
    uint64_t units = table_bytes / gran;
    int blocks = ctx->sm_count * 8;
    uint64_t threads = (uint64_t)blocks * 256;
    int mlp = loads_in_flight >= 8 ? 8 : loads_in_flight >= 4 ? 4 : loads_in_flight >= 2 ? 2 : 1;
    uint64_t per = (n_gathers / threads + mlp - 1) / mlp * mlp;
    if (per < (uint64_t)mlp) per = mlp;
    cudaEvent_t a, b;
    QK_CUDA(ctx, cudaEventCreate(&a));
    QK_CUDA(ctx, cudaEventCreate(&b));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) { // first repetition is the warm-up
        QK_CUDA(ctx, cudaEventRecord(a, st));
        if (gran == 32) qk_gather_dispatch<32>(mlp, dim3(blocks), st, table, units, per, ctx->stats + 3);
        else if (gran == 64) qk_gather_dispatch<64>(mlp, dim3(blocks), st, table, units, per, ctx->stats + 3);
        else qk_gather_dispatch<128>(mlp, dim3(blocks), st, table, units, per, ctx->stats + 3);
        QK_CUDA(ctx, cudaEventRecord(b, st));
        QK_CUDA(ctx, cudaEventSynchronize(b));
        QK_CUDA(ctx, cudaGetLastError());
        float ms;
        QK_CUDA(ctx, cudaEventElapsedTime(&ms, a, b));
        if (rep > 0 && ms < best) best = ms;
    }
    *gbs = (double)per * threads * gran / (best * 1e-3) / 1e9;
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(table);
    return QK_OK;
}

extern "C" int qk_bench_h2d(qk_ctx *ctx, size_t bytes, int repeats, double *gbs)
{
    if (!ctx || !gbs || bytes == 0 || bytes > ctx->chunk_capacity || repeats < 1) return QK_ERR_ARG;
    QK_CUDA(ctx, cudaSetDevice(ctx->device));
    qk_slot *sl = &ctx->slots[0];
    cudaEvent_t a, b;
    QK_CUDA(ctx, cudaEventCreate(&a));
    QK_CUDA(ctx, cudaEventCreate(&b));
    QK_CUDA(ctx, cudaMemcpyAsync(sl->dev, sl->host, bytes, cudaMemcpyHostToDevice, sl->stream)); // warm-up
    QK_CUDA(ctx, cudaEventRecord(a, sl->stream));
    for (int r = 0; r < repeats; ++r)
        QK_CUDA(ctx, cudaMemcpyAsync(sl->dev, sl->host, bytes, cudaMemcpyHostToDevice, sl->stream));
    QK_CUDA(ctx, cudaEventRecord(b, sl->stream));
    QK_CUDA(ctx, cudaEventSynchronize(b));
    float ms;
    QK_CUDA(ctx, cudaEventElapsedTime(&ms, a, b));
    *gbs = (double)bytes * repeats / (ms * 1e-3) / 1e9;
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    return QK_OK;
}
