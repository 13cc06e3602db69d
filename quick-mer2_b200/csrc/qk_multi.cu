// qk_multi.cu -- several GPUs in one process: one qk_ctx per device, the dictionary replicated
// with ncclBroadcast, the per-GPU counters combined with ncclReduce.
//
// The reference has no counterpart (one process, shared memory, Q.c:304-545).  What is kept:
// the result is the one `quicKmer2 count` produces, whatever the number of GPUs -- reads are
// independent units and counting is an integer sum, so sharding the reads and adding u32
// counters is exact (the 16-bit wrap of Q.c:23 is applied once, at the end).
//
// NCCL is bound at run time (dlopen of libnccl.so.2): the single-GPU path, the tests and any
// host without NCCL keep working, and inside a PyTorch process the library PyTorch already
// loaded is the one that gets used.  One NCCL communicator per device, all owned by this
// process (ncclCommInitAll); collectives are issued inside a group on each context's slot-0
// stream.
#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "qk_common.cuh"

#define QK_MULTI_MAX 16

struct qk_nccl_api {
    void *handle;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*GroupStart)(void);
    ncclResult_t (*GroupEnd)(void);
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t);
    const char *(*GetErrorString)(ncclResult_t);
};

struct qk_multi {
    uint32_t n;
    qk_ctx *ctx[QK_MULTI_MAX];
    ncclComm_t comm[QK_MULTI_MAX];
    qk_nccl_api api;
    char err[512];
};

static int qk_multi_fail(qk_multi *m, int code, const char *what, const char *detail)
{
    if (m) snprintf(m->err, sizeof m->err, "%s: %s", what, detail ? detail : "");
    return code;
}

static int qk_nccl_load(qk_multi *m)
{
    qk_nccl_api *a = &m->api;
    const char *names[] = {"libnccl.so.2", "libnccl.so", NULL};
    for (int i = 0; names[i] && !a->handle; ++i) a->handle = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!a->handle) return qk_multi_fail(m, QK_ERR_CUDA, "NCCL not found (multi-GPU needs libnccl.so.2)", dlerror());
#define QK_SYM(field, name)                                                      \
    do {                                                                         \
        *(void **)(&a->field) = dlsym(a->handle, name);                          \
        if (!a->field) return qk_multi_fail(m, QK_ERR_CUDA, "NCCL symbol missing", name); \
    } while (0)
    QK_SYM(CommInitAll, "ncclCommInitAll");
    QK_SYM(CommDestroy, "ncclCommDestroy");
    QK_SYM(GroupStart, "ncclGroupStart");
    QK_SYM(GroupEnd, "ncclGroupEnd");
    QK_SYM(Broadcast, "ncclBroadcast");
    QK_SYM(Reduce, "ncclReduce");
    QK_SYM(GetErrorString, "ncclGetErrorString");
#undef QK_SYM
    return QK_OK;
}

#define QK_NCCL(m, call)                                                                   \
    do {                                                                                   \
        ncclResult_t r__ = (call);                                                         \
        if (r__ != ncclSuccess) return qk_multi_fail(m, QK_ERR_CUDA, #call, (m)->api.GetErrorString(r__)); \
    } while (0)

extern "C" const char *qk_multi_last_error(const qk_multi *m) { return m ? m->err : "no multi-GPU context"; }

extern "C" void qk_multi_destroy(qk_multi *m)
{
    if (!m) return;
    for (uint32_t i = 0; i < m->n; ++i) {
        if (m->comm[i] && m->api.CommDestroy) m->api.CommDestroy(m->comm[i]);
        qk_ctx_destroy(m->ctx[i]);
    }
    free(m);
}

extern "C" int qk_multi_create(qk_multi **out, const int *devices, uint32_t n, uint32_t n_slots, size_t chunk_capacity)
{
    if (!out || !devices || n < 1 || n > QK_MULTI_MAX) return QK_ERR_ARG;
    *out = NULL;
    qk_multi *m = (qk_multi *)calloc(1, sizeof(qk_multi));
    if (!m) return QK_ERR_NOMEM;
    *out = m; // returned even on failure so that the caller can read the message
    for (uint32_t i = 0; i < n; ++i)
        for (uint32_t j = 0; j < i; ++j)
            if (devices[i] == devices[j]) return qk_multi_fail(m, QK_ERR_ARG, "device listed twice", "");
    for (uint32_t i = 0; i < n; ++i) {
        int rc = qk_ctx_create(&m->ctx[i], devices[i], n_slots, chunk_capacity);
        m->n = i + 1;
        if (rc) return qk_multi_fail(m, rc, "context", m->ctx[i] ? qk_last_error(m->ctx[i]) : "no CUDA device");
    }
    if (n > 1) {
        int rc = qk_nccl_load(m);
        if (rc) return rc;
        QK_NCCL(m, m->api.CommInitAll(m->comm, (int)n, devices));
    }
    return QK_OK;
}

extern "C" uint32_t qk_multi_size(const qk_multi *m) { return m ? m->n : 0; }
extern "C" qk_ctx *qk_multi_ctx(qk_multi *m, uint32_t i) { return (m && i < m->n) ? m->ctx[i] : NULL; }

// One array from context 0 to all the others.
static int qk_multi_bcast(qk_multi *m, void *const *ptrs, size_t bytes)
{
    if (bytes == 0) return QK_OK;
    QK_NCCL(m, m->api.GroupStart());
    for (uint32_t i = 0; i < m->n; ++i) {
        cudaSetDevice(m->ctx[i]->device);
        ncclResult_t r = m->api.Broadcast(ptrs[0], ptrs[i], bytes, ncclUint8, 0, m->comm[i], m->ctx[i]->slots[0].stream);
        if (r != ncclSuccess) {
            m->api.GroupEnd();
            return qk_multi_fail(m, QK_ERR_CUDA, "ncclBroadcast", m->api.GetErrorString(r));
        }
    }
    QK_NCCL(m, m->api.GroupEnd());
    return QK_OK;
}

// The dictionary built on context 0 -> every other context (table, stash, extension arrays).
extern "C" int qk_multi_replicate(qk_multi *m)
{
    if (!m) return QK_ERR_ARG;
    qk_table_desc d;
    int rc = qk_dict_describe(m->ctx[0], &d);
    if (rc) return qk_multi_fail(m, rc, "no dictionary on context 0", "");
    for (uint32_t i = 1; i < m->n; ++i) {
        rc = qk_dict_adopt(m->ctx[i], &d);
        if (rc) return qk_multi_fail(m, rc, "adopt", qk_last_error(m->ctx[i]));
    }
    if (m->n == 1) return QK_OK;
    void *p[5][QK_MULTI_MAX];
    for (uint32_t i = 0; i < m->n; ++i) {
        qk_dict_device_ptrs(m->ctx[i], &p[0][i], &p[1][i]);
        qk_dict_ext_ptrs(m->ctx[i], &p[2][i], &p[3][i], &p[4][i]);
    }
    const size_t bytes[5] = {d.table_bytes, d.stash_bytes, d.has_ext ? d.ext_bytes : 0, 0, 0}; // (one extension array)
    for (int a = 0; a < 5; ++a) {
        if (!bytes[a]) continue;
        rc = qk_multi_bcast(m, p[a], bytes[a]);
        if (rc) return rc;
    }
    for (uint32_t i = 0; i < m->n; ++i) {
        rc = qk_sync(m->ctx[i]);
        if (rc) return qk_multi_fail(m, rc, "sync", qk_last_error(m->ctx[i]));
    }
    return QK_OK;
}

// Sum of every context's u32 counters into context 0 (which then holds the whole job).
extern "C" int qk_multi_reduce(qk_multi *m)
{
    if (!m) return QK_ERR_ARG;
    if (m->n == 1) return qk_sync(m->ctx[0]);
    uint32_t *c[QK_MULTI_MAX];
    uint64_t n_kmers = 0;
    for (uint32_t i = 0; i < m->n; ++i) {
        int rc = qk_sync(m->ctx[i]); // every slot stream of the context, not only slot 0's
        if (rc) return qk_multi_fail(m, rc, "sync", qk_last_error(m->ctx[i]));
        rc = qk_counters_device_ptr(m->ctx[i], &c[i], &n_kmers);
        if (rc) return qk_multi_fail(m, rc, "counters", "");
    }
    QK_NCCL(m, m->api.GroupStart());
    for (uint32_t i = 0; i < m->n; ++i) {
        cudaSetDevice(m->ctx[i]->device);
        ncclResult_t r = m->api.Reduce(c[i], c[0], n_kmers, ncclUint32, ncclSum, 0, m->comm[i], m->ctx[i]->slots[0].stream);
        if (r != ncclSuccess) {
            m->api.GroupEnd();
            return qk_multi_fail(m, QK_ERR_CUDA, "ncclReduce", m->api.GetErrorString(r));
        }
    }
    QK_NCCL(m, m->api.GroupEnd());
    for (uint32_t i = 0; i < m->n; ++i) {
        int rc = qk_sync(m->ctx[i]);
        if (rc) return qk_multi_fail(m, rc, "sync", qk_last_error(m->ctx[i]));
    }
    return QK_OK;
}
