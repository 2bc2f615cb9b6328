// Communicator of the Lloyd path: (a) the peer-memory mailbox the fused finish kernel exchanges the k x (d+1)
// partials through (hk_finalize.cu; cudaIpc-mapped device memory of every rank of the box, NVLink/NVSwitch stores,
// no collective call on the critical path) and (b) NCCL plumbing for the generic hk_allreduce_f64 (functional value,
// unfused mixed-dtype path) and as the fallback when peer mapping is unavailable.
// Replaces MPICommunication.Allreduce(MPI.IN_PLACE, t, MPI.SUM) as issued 2k times per Lloyd
// iteration by heat/core/_operations.py:505-510 (heat/core/communication.py:1089-1110), which
// stages CUDA tensors through the host.  Here: one ncclAllReduce on the kernel's stream, device
// resident, over NVLink/NVSwitch.  libnccl is resolved with dlopen so that the library loads (and
// its symbols can be checked) on machines without NCCL or a GPU.
#include <dlfcn.h>
#include <string.h>

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

int comm_destroy(Handle* h);

int comm_unique_id(void* id128) {
    int rc = load_nccl();
    if (rc) return rc;
    HK_NCCL(g_nccl.GetUniqueId(reinterpret_cast<NcclUniqueId*>(id128)));
    return 0;
}

int comm_init(Handle* h, int nranks, int rank, const void* id128) {
    HK_ARG(nranks >= 1 && rank >= 0 && rank < nranks, "hk_comm_init: bad rank %d of %d", rank, nranks);
    comm_destroy(h);  // a re-initialisation must not leak the previous communicator / mailbox
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

static void peer_release(Handle* h) {
    for (int r = 0; r < HK_MAX_RANKS; ++r) {
        if (h->peer_opened[r]) cudaIpcCloseMemHandle(h->peer_opened[r]);
        h->peer_opened[r] = nullptr;
        h->peer_mbox[r] = nullptr;
        h->peer_flags[r] = nullptr;
    }
    if (h->peer_base) cudaFree(h->peer_base);
    h->peer_base = nullptr;
    h->peer_cap = 0;
    h->peer_ready = false;
}

static size_t peer_mbox_bytes(int nranks, size_t cap) { return (size_t)2 * nranks * cap * sizeof(double); }

// allocate this rank's mailbox (cap doubles per [parity][source rank] slot) and export its IPC handle (64 bytes)
int comm_peer_export(Handle* h, int64_t cap, void* handle64) {
    HK_ARG(h->nranks > 1 && h->nranks <= HK_MAX_RANKS, "hk_comm_peer_export: needs 2..%d ranks (have %d)", HK_MAX_RANKS,
           h->nranks);
    HK_ARG(cap >= 1 && handle64 != nullptr, "hk_comm_peer_export: bad argument");
    HK_CUDA(cudaSetDevice(h->device));
    peer_release(h);
    const size_t mb = peer_mbox_bytes(h->nranks, (size_t)cap);
    const size_t total = mb + (size_t)2 * h->nranks * sizeof(uint32_t) + 256;
    HK_CUDA(cudaMalloc(&h->peer_base, total));
    HK_CUDA(cudaMemset(h->peer_base, 0, total));
    HK_CUDA(cudaDeviceSynchronize());
    h->peer_cap = (size_t)cap;
    cudaIpcMemHandle_t ih;
    HK_CUDA(cudaIpcGetMemHandle(&ih, h->peer_base));
    static_assert(sizeof(ih) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &ih, 64);
    return 0;
}

// map every peer's mailbox: handles = nranks x 64 bytes in rank order (as gathered by the host)
int comm_peer_import(Handle* h, const void* handles) {
    HK_ARG(h->peer_base != nullptr && handles != nullptr, "hk_comm_peer_import: export first");
    HK_CUDA(cudaSetDevice(h->device));
    const size_t mb = peer_mbox_bytes(h->nranks, h->peer_cap);
    for (int r = 0; r < h->nranks; ++r) {
        char* base = nullptr;
        if (r == h->rank) {
            base = reinterpret_cast<char*>(h->peer_base);
        } else {
            cudaIpcMemHandle_t ih;
            memcpy(&ih, reinterpret_cast<const char*>(handles) + (size_t)r * 64, 64);
            void* ptr = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&ptr, ih, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                set_error("cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e));
                cudaGetLastError();
                for (int q = 0; q < h->nranks; ++q) {
                    if (h->peer_opened[q]) cudaIpcCloseMemHandle(h->peer_opened[q]);
                    h->peer_opened[q] = nullptr;
                }
                return 1000 + (int)e;
            }
            h->peer_opened[r] = ptr;
            base = reinterpret_cast<char*>(ptr);
        }
        h->peer_mbox[r] = reinterpret_cast<double*>(base);
        h->peer_flags[r] = reinterpret_cast<uint32_t*>(base + mb);
    }
    h->peer_ready = true;
    return 0;
}

int comm_destroy(Handle* h) {
    peer_release(h);
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
