/* qk_host_internal.h -- shared between the host .c files only (not installed). */
#ifndef QK_HOST_INTERNAL_H
#define QK_HOST_INTERNAL_H

#include "../../include/qk_host.h"

/* Threads that read input into pinned memory / inflate BGZF blocks: QK_READER_THREADS, else
 * min(8, online CPUs).  Callers cap it at what they can use. */
uint32_t qk_reader_threads_default(void);

#endif
