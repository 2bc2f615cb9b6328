// NCCL plumbing for the single per-iteration allreduce of the k x (d+1) partials.
// Replaces MPICommunication.Allreduce(MPI.IN_PLACE, t, MPI.SUM) as issued 2k times per Lloyd
// iteration by heat/core/_operations.py:505-510 (heat/core/communication.py:1089-1110), which
// stages CUDA tensors through the host.  Here: one ncclAllReduce on the kernel's stream, device
// resident, over NVLink/NVSwitch.  libnccl is resolved with dlopen so that the library loads (and
// its symbols can be checked) on machines without NCCL or a GPU.
#include <dlfcn.h>

#include "hk_common.cuh"

namespace hk {
namespace {

struct NcclUniqueId {
    char internal[128];
};
typedef void* NcclComm;
typedef int (*fn_GetUniqueId)(NcclUniqueId*);
typedef int (*fn_CommInitRank)(NcclComm*, int, NcclUniqueId, int);
typedef int (*fn_CommDestroy)(NcclComm);
typedef int (*fn_AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef const char* (*fn_GetErrorString)(int);

struct NcclApi {
    void* lib = nullptr;
    fn_GetUniqueId GetUniqueId = nullptr;
    fn_CommInitRank CommInitRank = nullptr;
    fn_CommDestroy CommDestroy = nullptr;
    fn_AllReduce AllReduce = nullptr;
    fn_GetErrorString GetErrorString = nullptr;
    bool tried = false;
};
NcclApi g_nccl;

int load_nccl() {
    if (g_nccl.lib) return 0;
    if (g_nccl.tried) {
        set_error("libnccl.so.2 could not be loaded");
        return -3;
    }
    g_nccl.tried = true;
    // torch (already imported by the host side) has its bundled libnccl.so.2 mapped; RTLD_NOLOAD first
    // so that this library and torch.distributed share one NCCL instance.
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) {
        set_error("dlopen(libnccl.so.2) failed: %s", dlerror());
        return -3;
    }
    g_nccl.lib = lib;
    g_nccl.GetUniqueId = (fn_GetUniqueId)dlsym(lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (fn_CommInitRank)dlsym(lib, "ncclCommInitRank");
    g_nccl.CommDestroy = (fn_CommDestroy)dlsym(lib, "ncclCommDestroy");
    g_nccl.AllReduce = (fn_AllReduce)dlsym(lib, "ncclAllReduce");
    g_nccl.GetErrorString = (fn_GetErrorString)dlsym(lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce) {
        set_error("libnccl is missing a required symbol");
        g_nccl.lib = nullptr;
        return -3;
    }
    return 0;
}

#define HK_NCCL(expr)                                                                          \
    do {                                                                                       \
        int _r = (expr);                                                                       \
        if (_r != 0) {                                                                         \
            set_error("%s failed: %s", #expr,                                                  \
                      g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "nccl error");       \
            return 2000 + _r;                                                                  \
        }                                                                                      \
    } while (0)

}  // namespace

int comm_unique_id(void* id128) {
    int rc = load_nccl();
    if (rc) return rc;
    HK_NCCL(g_nccl.GetUniqueId(reinterpret_cast<NcclUniqueId*>(id128)));
    return 0;
}

int comm_init(Handle* h, int nranks, int rank, const void* id128) {
    HK_ARG(nranks >= 1 && rank >= 0 && rank < nranks, "hk_comm_init: bad rank %d of %d", rank, nranks);
    h->nranks = nranks;
    h->rank = rank;
    if (nranks == 1) return 0;
    int rc = load_nccl();
    if (rc) return rc;
    HK_CUDA(cudaSetDevice(h->device));
    NcclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    NcclComm c = nullptr;
    HK_NCCL(g_nccl.CommInitRank(&c, nranks, id, rank));
    h->nccl_comm = c;
    return 0;
}

int comm_destroy(Handle* h) {
    if (h->nccl_comm && g_nccl.CommDestroy) {
        g_nccl.CommDestroy((NcclComm)h->nccl_comm);
        h->nccl_comm = nullptr;
    }
    h->nranks = 1;
    h->rank = 0;
    return 0;
}

int comm_allreduce_f64(Handle* h, double* buf, int64_t count, cudaStream_t stream) {
    if (h->nranks <= 1) return 0;  // mirrors the np == 1 short-circuit (communication.py:1064-1065)
    HK_ARG(h->nccl_comm != nullptr, "hk_allreduce_f64: communicator not initialised");
    // ncclFloat64 = 8, ncclSum = 0
    HK_NCCL(g_nccl.AllReduce(buf, buf, (size_t)count, 8, 0, (NcclComm)h->nccl_comm, stream));
    return 0;
}

}  // namespace hk
