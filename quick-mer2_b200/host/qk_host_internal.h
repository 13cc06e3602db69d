/* qk_host_internal.h -- shared between the host .c files only (not installed). */
#ifndef QK_HOST_INTERNAL_H
#define QK_HOST_INTERNAL_H

#include "../../include/qk_host.h"

/* Threads that read input into pinned memory / inflate BGZF blocks: QK_READER_THREADS, else
 * min(8, online CPUs).  Callers cap it at what they can use. */
uint32_t qk_reader_threads_default(void);

/* A file range of fixed-size elements through the slots' pinned buffers: reader threads pread() pieces in parallel,
 * the caller's thread hands them over IN ORDER: handle(ctx, slot, element offset, element count, user). */
typedef int (*qk_piece_handler)(qk_ctx *ctx, uint32_t slot, uint64_t elem_offset, uint64_t count, void *user);
int qk_ingest_elements(qk_ctx *ctx, int fd, uint64_t file_off, uint64_t n_elems, size_t esz, uint32_t threads,
                       qk_piece_handler handle, void *user);

#endif
