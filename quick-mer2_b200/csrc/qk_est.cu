// qk_est.cu -- the window reduction of `quicKmer2 est` (Q.c:660-682) on the device.
//
// est turns a sample's depths into copy number per window: for every window [left, right) of k-mer
// ordinals (<ref>.bed) it adds correction[gc(i)] * depth(i) -- a FLOAT product added to a DOUBLE, in
// ordinal order -- over the window's k-mers, then divides by the window length and by half the mean
// depth.  The reference does that in one serial pass over <ref>.qgc and <sample>.bin.  Here both files are
// streamed to the device through the pinned slots (as the dictionary is) and one thread per window adds
// its k-mers up in the reference's order and precision, so the printed "%f" values are the same bytes.
// SURVEY.md 8(f) rank 4.
#include "qk_common.cuh"

extern "C" int qk_est_begin(qk_ctx *ctx, uint64_t n_entries)
{
    if (!ctx || n_entries == 0) return QK_ERR_ARG;
    int rc = qk_sync(ctx);
    if (rc) return rc;
    QK_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaFree(ctx->est_depth);
    cudaFree(ctx->est_qgc);
    ctx->est_depth = ctx->est_qgc = NULL;
    ctx->est_n = 0;
    QK_CUDA(ctx, cudaMalloc((void **)&ctx->est_depth, n_entries * sizeof(uint16_t)));
    QK_CUDA(ctx, cudaMalloc((void **)&ctx->est_qgc, n_entries * sizeof(uint16_t)));
    ctx->est_n = n_entries;
    return QK_OK;
}

// kind 0 = depths (<sample>.bin), 1 = GC flags (<ref>.qgc): `count` entries from the pinned buffer of `slot`
extern "C" int qk_est_upload_from_slot(qk_ctx *ctx, uint32_t slot, int kind, uint64_t elem_offset, uint64_t count)
{
    if (!ctx || slot >= ctx->n_slots || !ctx->est_depth) return QK_ERR_ARG;
    if (elem_offset + count > ctx->est_n || count * sizeof(uint16_t) > ctx->chunk_capacity)
        return qk_fail(ctx, QK_ERR_ARG, "est piece outside the arrays or larger than a slot");
    QK_CUDA(ctx, cudaSetDevice(ctx->device));
    qk_slot *sl = &ctx->slots[slot];
    uint16_t *dst = (kind ? ctx->est_qgc : ctx->est_depth) + elem_offset;
    QK_CUDA(ctx, cudaMemcpyAsync(dst, sl->host, count * sizeof(uint16_t), cudaMemcpyHostToDevice, sl->stream));
    QK_CUDA(ctx, cudaEventRecord(sl->h2d_done, sl->stream));
    return QK_OK;
}

// Q.c:677-679 for k-mers lo[w] .. hi[w]-1 of window w, in order.  GC bins above 400 index past the reference's
// 401-float array (undefined there): they contribute nothing here.
__global__ void qk_est_kernel(const uint16_t *__restrict__ depth, const uint16_t *__restrict__ qgc, const float *__restrict__ corr,
                              const unsigned long long *__restrict__ lo, const unsigned long long *__restrict__ hi,
                              uint64_t n_windows, double *__restrict__ sums)
{
    __shared__ float s_corr[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) s_corr[i] = i < QK_GC_BINS ? corr[i] : 0.0f;
    __syncthreads();
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_windows) return;
    double acc = 0.0;
    for (unsigned long long i = lo[w]; i < hi[w]; ++i) {
        const float term = __fmul_rn(s_corr[qgc[i] & 0x1FFu], (float)depth[i]);   // float * (int -> float), not fused
        acc = __dadd_rn(acc, (double)term);
    }
    sums[w] = acc;
}

extern "C" int qk_est_windows(qk_ctx *ctx, const float correction[QK_GC_BINS], const uint64_t *lo, const uint64_t *hi,
                              uint64_t n_windows, double *sums_out)
{
    if (!ctx || !correction || !lo || !hi || !sums_out || !ctx->est_depth) return QK_ERR_ARG;
    if (n_windows == 0) return QK_OK;
    for (uint64_t w = 0; w < n_windows; ++w)
        if (hi[w] > ctx->est_n || (lo[w] > hi[w])) return qk_fail(ctx, QK_ERR_ARG, "window %llu outside the arrays", (unsigned long long)w);
    int rc = qk_sync(ctx);                      // the uploads
    if (rc) return rc;
    QK_CUDA(ctx, cudaSetDevice(ctx->device));
    float *d_corr = NULL;
    unsigned long long *d_lo = NULL, *d_hi = NULL;
    double *d_sums = NULL;
    cudaError_t e = cudaMalloc((void **)&d_corr, QK_GC_BINS * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_lo, n_windows * 8);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_hi, n_windows * 8);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_sums, n_windows * 8);
    if (e == cudaSuccess) e = cudaMemcpy(d_corr, correction, QK_GC_BINS * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_lo, lo, n_windows * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_hi, hi, n_windows * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        qk_est_kernel<<<(unsigned)((n_windows + 127) / 128), 128>>>(ctx->est_depth, ctx->est_qgc, d_corr, d_lo, d_hi, n_windows, d_sums);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(sums_out, d_sums, n_windows * 8, cudaMemcpyDeviceToHost);
    cudaFree(d_corr); cudaFree(d_lo); cudaFree(d_hi); cudaFree(d_sums);
    if (e != cudaSuccess) return qk_cuda_fail(ctx, e, "est windows");
    return QK_OK;
}

extern "C" int qk_est_end(qk_ctx *ctx)
{
    if (!ctx) return QK_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaFree(ctx->est_depth);
    cudaFree(ctx->est_qgc);
    ctx->est_depth = ctx->est_qgc = NULL;
    ctx->est_n = 0;
    return QK_OK;
}
