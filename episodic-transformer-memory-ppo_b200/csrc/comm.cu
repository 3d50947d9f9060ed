// Multi-GPU gradient exchange behind the C ABI: one NCCL communicator per rank, one in-place sum
// all-reduce of the flat fp32 gradient arena per optimiser step, enqueued on the caller's stream.
//
// The reference is single-process (no collective anywhere); this is the exchange step the
// data-parallel sharding of trainer.py:145-323 over workers needs.  NCCL is bound at run time with dlopen so that
// libtrxlppo.so has no link-time dependency on it: single-GPU users never load it, and under PyTorch the process-wide
// libnccl.so.2 that torch already mapped is reused (same soname), so both share one NCCL build.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/trxl_ppo.h"
#include "common.cuh"

namespace {

// the slice of nccl.h this file needs (ABI-stable since NCCL 2.0)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
constexpr int kNcclFloat32 = 7, kNcclFloat64 = 8, kNcclSum = 0;

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
};

NcclApi g_nccl;

int load_nccl() {
    if (g_nccl.handle) return TRXL_OK;
    const char* override_path = getenv("TRXL_NCCL_LIB");
    const char* names[] = {override_path, "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        if (!n || !n[0]) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        trxl_set_error("NCCL not found (dlopen libnccl.so.2 failed: %s); set TRXL_NCCL_LIB to its path", dlerror());
        return TRXL_ERR_CUDA;
    }
    NcclApi a;
    a.handle = h;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(h, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
    a.AllReduce = (decltype(a.AllReduce))dlsym(h, "ncclAllReduce");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
    a.GetVersion = (decltype(a.GetVersion))dlsym(h, "ncclGetVersion");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce || !a.GetErrorString) {
        trxl_set_error("libnccl is missing a required symbol");
        return TRXL_ERR_CUDA;
    }
    g_nccl = a;
    return TRXL_OK;
}

int nccl_check(ncclResult_t r, const char* what) {
    if (r == 0) return TRXL_OK;
    trxl_set_error("%s failed: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    return TRXL_ERR_CUDA;
}

struct Comm {
    ncclComm_t comm;
    int rank, world;
    long long calls;
};

}  // namespace

extern "C" {

int trxl_comm_unique_id(void* id_out) {
    TRXL_CHECK_ARG(id_out != nullptr, "trxl_comm_unique_id: NULL output");
    TRXL_PROPAGATE(load_nccl());
    ncclUniqueId id;
    TRXL_PROPAGATE(nccl_check(g_nccl.GetUniqueId(&id), "ncclGetUniqueId"));
    memcpy(id_out, id.internal, TRXL_COMM_ID_BYTES);
    return TRXL_OK;
}

int trxl_comm_create(const void* id_bytes, int rank, int world_size, void** comm_out) {
    TRXL_CHECK_ARG(id_bytes && comm_out, "trxl_comm_create: NULL argument");
    TRXL_CHECK_ARG(world_size >= 1 && rank >= 0 && rank < world_size, "trxl_comm_create: bad rank %d of %d", rank, world_size);
    TRXL_PROPAGATE(load_nccl());
    ncclUniqueId id;
    memcpy(id.internal, id_bytes, TRXL_COMM_ID_BYTES);
    ncclComm_t c = nullptr;
    TRXL_PROPAGATE(nccl_check(g_nccl.CommInitRank(&c, world_size, id, rank), "ncclCommInitRank"));
    Comm* out = new Comm{c, rank, world_size, 0};
    *comm_out = out;
    return TRXL_OK;
}

int trxl_comm_destroy(void* comm) {
    if (!comm) return TRXL_OK;
    Comm* c = static_cast<Comm*>(comm);
    int rc = TRXL_OK;
    if (g_nccl.CommDestroy) rc = nccl_check(g_nccl.CommDestroy(c->comm), "ncclCommDestroy");
    delete c;
    return rc;
}

int64_t trxl_comm_calls(void* comm) { return comm ? static_cast<Comm*>(comm)->calls : 0; }

int trxl_comm_nccl_version(void) {
    if (load_nccl() != TRXL_OK || !g_nccl.GetVersion) return -1;
    int v = 0;
    return g_nccl.GetVersion(&v) == 0 ? v : -1;
}

int trxl_allreduce_grads(void* comm, float* buf, int64_t count, void* stream) {
    TRXL_CHECK_ARG(comm && buf && count >= 0, "trxl_allreduce_grads: bad argument");
    Comm* c = static_cast<Comm*>(comm);
    ++c->calls;
    return nccl_check(g_nccl.AllReduce(buf, buf, (size_t)count, kNcclFloat32, kNcclSum, c->comm, (cudaStream_t)stream), "ncclAllReduce(f32)");
}

int trxl_allreduce_f64(void* comm, double* buf, int64_t count, void* stream) {
    TRXL_CHECK_ARG(comm && buf && count >= 0, "trxl_allreduce_f64: bad argument");
    Comm* c = static_cast<Comm*>(comm);
    ++c->calls;
    return nccl_check(g_nccl.AllReduce(buf, buf, (size_t)count, kNcclFloat64, kNcclSum, c->comm, (cudaStream_t)stream), "ncclAllReduce(f64)");
}

}  // extern "C"
