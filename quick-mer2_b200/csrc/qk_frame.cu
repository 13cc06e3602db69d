// qk_frame.cu -- record framing on the device: raw FASTA / FASTQ bytes in, "sequence lines
// only" out, in place.
//
// Replaces the reference's fgets loop (Q.c:393-398, 451-455) for chunks that are shipped to
// the GPU as they come off the file: header, '+' and quality lines are turned into runs of
// '\n' (empty lines emit no k-mer, so the count kernel needs no change), sequence lines are
// left alone.  The host then does no per-byte work at all on the hot path -- it only cuts the
// stream at line ends -- which matters because a host framer tops out at a few GB/s per core
// while the H2D link moves 55 GB/s.
//
// The reference's loop is a finite-state machine over LINES.  State s = number of lines
// still to be discarded (0 = the next line is examined):
//     s > 0                  : the line is discarded, s <- (s + 1) mod 4   [3 -> 0]
//     s = 0, line[0] == '>'  : skipped, s stays 0                          (Q.c:398)
//     s = 0, otherwise       : the line is a READ; FASTQ: s <- 1 (Q.c:451-455); FASTA: s stays 0
// The first line of the stream: '@' selects FASTQ and is consumed (start in s = 3); a FASTA
// pipe loses its first line the same way (fseek fails, Q.c:396); seekable FASTA starts in 0.
// Each line is therefore a function {0..3} -> {0..3} x {keep, drop}; composition of such
// functions is associative, so the state before every line comes out of a parallel scan
// (a thread composes the lines that start in its 64 bytes, then one CTA-wide scan per 16 KiB):
//   pass 1  qk_frame_reduce : every CTA composes the lines that START in its span
//   pass 2  qk_frame_carry  : one thread chains the CTA totals from the stream state
//   pass 3  qk_frame_apply  : every CTA replays its span from its true incoming state and
//                             blanks the bytes of discarded lines
// The stream state (FSM state, line and base totals) lives in device memory and is handed
// from chunk to chunk in stream order, so the host never needs to know the phase.
#include "qk_common.cuh"

int qk_ring_push(qk_ctx *ctx, qk_slot *sl, int kind, qk_timing_pair **out);
int qk_launch_count(qk_ctx *ctx, qk_slot *sl, const uint8_t *dev_bytes, size_t n_bytes);

#define QK_FRAME_THREADS 256
#define QK_FRAME_PIECES 4                                     // 16-byte pieces per thread per tile
#define QK_FRAME_TILE (QK_FRAME_THREADS * 16 * QK_FRAME_PIECES) // 16 KiB: one CTA-wide scan per tile

// Transition function packed in 13 bits: map[s] in bits 2s+1:2s, keep[s] in bit 8+s (keep flag
// of the LAST line that starts in the segment, entered in state s), bit 12 = segment has a line.
#define QK_FE_IDENTITY 0xE4u
__device__ __forceinline__ uint32_t qk_fe_line(uint32_t first_byte, uint32_t fastq)
{
    const uint32_t hdr = first_byte == '>';
    const uint32_t m0 = (fastq && !hdr) ? 1u : 0u;           // state 0 -> 1 after a FASTQ read
    return (m0 | (2u << 2) | (3u << 4) | (0u << 6)) | ((hdr ? 0u : 1u) << 8) | (1u << 12);
}
__device__ __forceinline__ uint32_t qk_fe_map(uint32_t e, uint32_t s) { return (e >> (2 * s)) & 3u; }
// a then b
__device__ __forceinline__ uint32_t qk_fe_compose(uint32_t a, uint32_t b)
{
    uint32_t r = 0;
#pragma unroll
    for (uint32_t s = 0; s < 4; ++s) {
        const uint32_t mid = qk_fe_map(a, s);
        r |= qk_fe_map(b, mid) << (2 * s);
        const uint32_t keep = (b >> 12) & 1u ? (b >> (8 + mid)) & 1u : (a >> (8 + s)) & 1u;
        r |= keep << (8 + s);
    }
    return r | ((a | b) & (1u << 12));
}
// concrete (state, keep-of-open-line) packed as state | keep << 2, advanced by element e
__device__ __forceinline__ uint32_t qk_fe_apply(uint32_t e, uint32_t sk)
{
    const uint32_t s = sk & 3u;
    const uint32_t keep = (e >> 12) & 1u ? (e >> (8 + s)) & 1u : (sk >> 2) & 1u;
    return qk_fe_map(e, s) | (keep << 2);
}

__device__ __forceinline__ uint4 qk_frame_load16(const uint8_t *__restrict__ bytes, uint32_t pos, uint32_t n)
{
    uint4 v = make_uint4(0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au);
    if (pos < n) {
        v = *reinterpret_cast<const uint4 *>(bytes + pos);
        const uint32_t rem = n - pos;
        if (rem < 16) {
            uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int keep = (int)rem - 4 * i;
                const uint32_t m = keep >= 4 ? 0xFFFFFFFFu : keep <= 0 ? 0u : ((1u << (8 * keep)) - 1);
                w[i] = (w[i] & m) | (0x0A0A0A0Au & ~m);
            }
            v = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
    return v;
}
__device__ __forceinline__ uint32_t qk_eq4(uint32_t w, uint32_t pat)
{
    return ((__vcmpeq4(w, pat) & 0x01010101u) * 0x01020408u) >> 24;
}
__device__ __forceinline__ uint32_t qk_eq16(const uint4 &v, uint32_t pat)
{
    return qk_eq4(v.x, pat) | (qk_eq4(v.y, pat) << 4) | (qk_eq4(v.z, pat) << 8) | (qk_eq4(v.w, pat) << 12);
}
__device__ __forceinline__ uint32_t qk_byte_of(const uint4 &v, uint32_t j)
{
    const uint32_t w = j < 4 ? v.x : j < 8 ? v.y : j < 12 ? v.z : v.w;
    return (w >> (8 * (j & 3))) & 0xFFu;
}

struct qk_frame_args {
    uint8_t *bytes;
    uint32_t n_bytes;
    uint32_t n_tiles;       // QK_FRAME_TILE-byte tiles
    uint32_t tiles_per_cta;
    uint32_t fastq;
    uint32_t *cta_elem;     // per-CTA composed element (pass 1), then incoming state | keep << 2 (pass 2)
    unsigned long long *stream; // [0] FSM state, [1] read lines, [2] bases, [3] raw lines
};

// Element of one thread's 16 bytes + bookkeeping shared by passes 1 and 3.
struct qk_frame_piece {
    uint4 v;
    uint32_t nl;     // bit j: byte j is '\n'
    uint32_t starts; // bit j: byte j is the first byte of a line
    uint32_t elem;
};

__device__ __forceinline__ qk_frame_piece qk_frame_piece_of(const qk_frame_args &a, const uint4 &v, uint32_t pos,
                                                            uint32_t prev_is_nl)
{
    qk_frame_piece p;
    p.v = v;
    p.nl = qk_eq16(p.v, 0x0A0A0A0Au);
    p.starts = ((p.nl << 1) | prev_is_nl) & 0xFFFFu;
    if (pos >= a.n_bytes) p.starts = 0;
    else if (a.n_bytes - pos < 16) p.starts &= (1u << (a.n_bytes - pos)) - 1; // no line starts past the end
    uint32_t e = QK_FE_IDENTITY, m = p.starts;
    while (m) {
        const uint32_t j = __ffs(m) - 1;
        m &= m - 1;
        e = qk_fe_compose(e, qk_fe_line(qk_byte_of(p.v, j), a.fastq));
    }
    p.elem = e;
    return p;
}

// Inclusive scan (composition, in thread order) of `e` over the CTA; returns the exclusive
// prefix for this thread and the CTA total through *total.
__device__ __forceinline__ uint32_t qk_frame_block_scan(uint32_t e, uint32_t *s_warp, uint32_t *total)
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = e;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl = qk_fe_compose(up, incl);
    }
    uint32_t excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = QK_FE_IDENTITY;
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t before = QK_FE_IDENTITY, all = QK_FE_IDENTITY;
#pragma unroll
    for (uint32_t w = 0; w < QK_FRAME_THREADS / 32; ++w) {
        const uint32_t we = s_warp[w];
        if (w < warp) before = qk_fe_compose(before, we);
        all = qk_fe_compose(all, we);
    }
    __syncthreads();
    *total = all;
    return qk_fe_compose(before, excl);
}

template <bool APPLY>
__global__ void __launch_bounds__(QK_FRAME_THREADS) qk_frame_kernel(const qk_frame_args a)
{
    __shared__ uint32_t s_warp[QK_FRAME_THREADS / 32];
    __shared__ uint8_t s_last[QK_FRAME_THREADS];
    const uint32_t tid = threadIdx.x;
    const uint32_t tile0 = blockIdx.x * a.tiles_per_cta;
    if (tile0 >= a.n_tiles) {
        if (!APPLY && tid == 0) a.cta_elem[blockIdx.x] = QK_FE_IDENTITY | (1u << 13);
        return;
    }
    const uint32_t tile_end = min(tile0 + a.tiles_per_cta, a.n_tiles);
    // Is the byte before the span a '\n'?  (The chunk starts at a line start.)  Pass 1 reads it
    // from the data and leaves it in bit 13 of the CTA word; pass 3 must not look at the data
    // again, because the neighbouring CTA may already have blanked that byte.
    uint32_t carry_nl = 1;
    const uint32_t word = APPLY ? a.cta_elem[blockIdx.x] : 0;
    if (APPLY) carry_nl = (word >> 13) & 1u;
    else if (tile0 > 0) carry_nl = a.bytes[tile0 * QK_FRAME_TILE - 1] == '\n';
    const uint32_t first_carry_nl = carry_nl;
    uint32_t run = QK_FE_IDENTITY;                // pass 1: composition so far
    uint32_t sk = word & 7u;                      // pass 3: concrete state | keep << 2
    uint32_t n_reads = 0, n_bases = 0, n_lines = 0;

    for (uint32_t tile = tile0; tile < tile_end; ++tile) {
        const uint32_t pos0 = tile * QK_FRAME_TILE + tid * (16 * QK_FRAME_PIECES);
        uint4 v[QK_FRAME_PIECES];
#pragma unroll
        for (int i = 0; i < QK_FRAME_PIECES; ++i) v[i] = qk_frame_load16(a.bytes, pos0 + 16 * i, a.n_bytes);
        // the byte before my first piece is the last byte of the previous thread's last piece
        s_last[tid] = (uint8_t)(v[QK_FRAME_PIECES - 1].w >> 24);
        __syncthreads();
        uint32_t prev_nl = tid ? (s_last[tid - 1] == '\n') : carry_nl;
        const uint32_t tile_last_nl = s_last[QK_FRAME_THREADS - 1] == '\n';
        qk_frame_piece p[QK_FRAME_PIECES];
        uint32_t elem = QK_FE_IDENTITY;
#pragma unroll
        for (int i = 0; i < QK_FRAME_PIECES; ++i) {
            p[i] = qk_frame_piece_of(a, v[i], pos0 + 16 * i, prev_nl);
            prev_nl = (p[i].nl >> 15) & 1u;
            if (p[i].starts) elem = qk_fe_compose(elem, p[i].elem);
        }
        uint32_t total;
        const uint32_t before = qk_frame_block_scan(elem, s_warp, &total); // syncs: s_last is free again
        if (!APPLY) {
            run = qk_fe_compose(run, total);
        } else {
            uint32_t cur = qk_fe_apply(before, sk);   // state | keep << 2 entering this thread's bytes
#pragma unroll
            for (int i = 0; i < QK_FRAME_PIECES; ++i) {
                const uint32_t pos = pos0 + 16 * i;
                // keep mask of the 16 bytes: segments between line starts
                uint32_t keepmask = 0, m = p[i].starts, from = 0;
                while (m) {
                    const uint32_t j = __ffs(m) - 1;
                    m &= m - 1;
                    if ((cur >> 2) & 1u) keepmask |= ((1u << j) - 1) & ~((1u << from) - 1);
                    const uint32_t e = qk_fe_line(qk_byte_of(p[i].v, j), a.fastq);
                    cur = qk_fe_apply(e, cur);
                    n_reads += (cur >> 2) & 1u;
                    ++n_lines;
                    from = j;
                }
                if ((cur >> 2) & 1u) keepmask |= 0xFFFFu & ~((1u << from) - 1);
                if (pos < a.n_bytes) {
                    uint32_t valid = a.n_bytes - pos >= 16 ? 0xFFFFu : (1u << (a.n_bytes - pos)) - 1;
                    n_bases += __popc(keepmask & ~p[i].nl & valid);
                    uint32_t w[4] = {p[i].v.x, p[i].v.y, p[i].v.z, p[i].v.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t bits = (keepmask >> (4 * q)) & 0xFu;
                        const uint32_t sel = (((bits * 0x00204081u) & 0x01010101u) * 0xFFu);
                        w[q] = (w[q] & sel) | (0x0A0A0A0Au & ~sel);
                    }
                    if (keepmask != 0xFFFFu) *reinterpret_cast<uint4 *>(a.bytes + pos) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            sk = qk_fe_apply(total, sk);
        }
        carry_nl = tile_last_nl;
    }
    if (!APPLY) {
        if (tid == 0) a.cta_elem[blockIdx.x] = run | (first_carry_nl << 13);
    } else {
        for (int o = 16; o; o >>= 1) {
            n_reads += __shfl_xor_sync(0xffffffffu, n_reads, o);
            n_bases += __shfl_xor_sync(0xffffffffu, n_bases, o);
            n_lines += __shfl_xor_sync(0xffffffffu, n_lines, o);
        }
        if ((tid & 31) == 0) {
            if (n_reads) atomicAdd(a.stream + 1, (unsigned long long)n_reads);
            if (n_bases) atomicAdd(a.stream + 2, (unsigned long long)n_bases);
            if (n_lines) atomicAdd(a.stream + 3, (unsigned long long)n_lines);
        }
    }
}

// pass 2: chain the CTA totals from the stream state; leave each CTA's incoming state behind.
// One CTA: everybody stages the <= QK_FRAME_MAX_CTAS words in shared memory, one thread chains
// them there (a dependent global load per word cost 54 ns each), everybody writes back.
__global__ void __launch_bounds__(256) qk_frame_carry(uint32_t *cta_elem, uint32_t n_ctas, unsigned long long *stream)
{
    __shared__ uint32_t s_e[QK_FRAME_MAX_CTAS];
    for (uint32_t b = threadIdx.x; b < n_ctas; b += blockDim.x) s_e[b] = cta_elem[b];
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t sk = (uint32_t)stream[0] & 3u;
        for (uint32_t b = 0; b < n_ctas; ++b) {
            const uint32_t e = s_e[b];
            s_e[b] = sk | (e & (1u << 13));
            sk = qk_fe_apply(e & 0x1FFFu, sk);
        }
        stream[0] = sk & 3u;
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < n_ctas; b += blockDim.x) cta_elem[b] = s_e[b];
}

extern "C" int qk_raw_begin(qk_ctx *ctx, int fastq, int skip_first_line)
{
    return qk_raw_begin_state(ctx, fastq, skip_first_line ? 3u : 0u);
}

extern "C" int qk_raw_state(qk_ctx *ctx, uint32_t *line_state)
{
    if (!ctx || !line_state) return QK_ERR_ARG;
    int rc = qk_sync(ctx);
    if (rc) return rc;
    unsigned long long s = 0;
    QK_CUDA(ctx, cudaMemcpy(&s, ctx->frame_stream, sizeof s, cudaMemcpyDeviceToHost));
    *line_state = (uint32_t)s & 3u;
    return QK_OK;
}

extern "C" int qk_raw_begin_state(qk_ctx *ctx, int fastq, uint32_t line_state)
{
    if (!ctx || line_state > 3) return QK_ERR_ARG;
    int rc = qk_sync(ctx);
    if (rc) return rc;
    QK_CUDA(ctx, cudaSetDevice(ctx->device));
    // On a slot stream and waited for there: a plain cudaMemcpy from pageable memory returns once the bytes are
    // STAGED, and the slot streams (non-blocking) do not order against the default stream -- with another DMA in
    // flight (qk_finish_async) the state landed ~10 ms into the job and re-phased the framer mid-stream.
    unsigned long long init[4] = {line_state, 0, 0, 0};
    QK_CUDA(ctx, cudaMemcpyAsync(ctx->frame_stream, init, sizeof init, cudaMemcpyHostToDevice, ctx->slots[0].stream));
    QK_CUDA(ctx, cudaStreamSynchronize(ctx->slots[0].stream));
    ctx->raw_fastq = fastq ? 1 : 0;
    ctx->raw_active = 1;
    ctx->raw_prev_slot = -1;
    return QK_OK;
}

extern "C" int qk_submit_raw(qk_ctx *ctx, uint32_t slot, const uint8_t *bytes, size_t n_bytes)
{
    if (!ctx || slot >= ctx->n_slots || (!bytes && n_bytes)) return QK_ERR_ARG;
    if (!ctx->raw_active) return qk_fail(ctx, QK_ERR_STATE, "qk_raw_begin not called");
    if (n_bytes > ctx->chunk_capacity) return qk_fail(ctx, QK_ERR_ARG, "chunk of %zu bytes exceeds the slot capacity", n_bytes);
    if (ctx->dict_state != 2) return qk_fail(ctx, QK_ERR_STATE, "no dictionary built on this context");
    if (n_bytes == 0) return QK_OK;
    if (bytes[n_bytes - 1] != '\n') return qk_fail(ctx, QK_ERR_ARG, "a raw chunk must end at a line end");
    QK_CUDA(ctx, cudaSetDevice(ctx->device));
    qk_slot *sl = &ctx->slots[slot];
    qk_timing_pair *tp;
    int rc = qk_ring_push(ctx, sl, 0, &tp);
    if (rc) return rc;
    QK_CUDA(ctx, cudaEventRecord(tp->a, sl->stream));
    QK_CUDA(ctx, cudaMemcpyAsync(sl->dev, bytes, n_bytes, cudaMemcpyHostToDevice, sl->stream));
    QK_CUDA(ctx, cudaEventRecord(tp->b, sl->stream));
    QK_CUDA(ctx, cudaEventRecord(sl->h2d_done, sl->stream));

    qk_frame_args a;
    a.bytes = sl->dev;
    a.n_bytes = (uint32_t)n_bytes;
    a.n_tiles = (uint32_t)((n_bytes + QK_FRAME_TILE - 1) / QK_FRAME_TILE);
    const uint32_t max_ctas = QK_FRAME_MAX_CTAS < (uint32_t)ctx->sm_count * 8 ? QK_FRAME_MAX_CTAS : (uint32_t)ctx->sm_count * 8;
    a.tiles_per_cta = (a.n_tiles + max_ctas - 1) / max_ctas;
    if (a.tiles_per_cta < 1) a.tiles_per_cta = 1;
    const uint32_t n_ctas = (a.n_tiles + a.tiles_per_cta - 1) / a.tiles_per_cta;
    a.fastq = (uint32_t)ctx->raw_fastq;
    a.cta_elem = ctx->frame_elems + (size_t)slot * QK_FRAME_MAX_CTAS;
    a.stream = ctx->frame_stream;
    // the FSM state is handed from chunk to chunk: this chunk's carry pass must run after the
    // previous chunk's, whichever slot stream that was on
    if (ctx->raw_prev_slot >= 0 && (uint32_t)ctx->raw_prev_slot != slot)
        QK_CUDA(ctx, cudaStreamWaitEvent(sl->stream, ctx->slots[ctx->raw_prev_slot].frame_done, 0));
    qk_frame_kernel<false><<<n_ctas, QK_FRAME_THREADS, 0, sl->stream>>>(a);
    qk_frame_carry<<<1, 256, 0, sl->stream>>>(a.cta_elem, n_ctas, a.stream);
    QK_CUDA(ctx, cudaEventRecord(sl->frame_done, sl->stream));
    qk_frame_kernel<true><<<n_ctas, QK_FRAME_THREADS, 0, sl->stream>>>(a);
    QK_CUDA(ctx, cudaGetLastError());
    ctx->raw_prev_slot = (int)slot;
    ctx->frame_launches += 3;
    return qk_launch_count(ctx, sl, sl->dev, n_bytes);
}

extern "C" int qk_raw_stats(qk_ctx *ctx, uint64_t *read_lines, uint64_t *bases, uint64_t *raw_lines)
{
    if (!ctx) return QK_ERR_ARG;
    int rc = qk_sync(ctx);
    if (rc) return rc;
    unsigned long long h[4];
    QK_CUDA(ctx, cudaMemcpy(h, ctx->frame_stream, sizeof h, cudaMemcpyDeviceToHost));
    if (read_lines) *read_lines = h[1];
    if (bases) *bases = h[2];
    if (raw_lines) *raw_lines = h[3];
    return QK_OK;
}

extern "C" int qk_host_is_pinned(const void *p)
{
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return 0; }
    return attr.type == cudaMemoryTypeHost;
}
