// Shared declarations for libhkmeans.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <utility>
#include <vector>

#include "../../include/hkmeans.h"

namespace hk {

// ---- error plumbing ----------------------------------------------------------------------------
void set_error(const char* fmt, ...);
#define HK_CUDA(expr)                                                                  \
    do {                                                                               \
        cudaError_t _e = (expr);                                                       \
        if (_e != cudaSuccess) {                                                       \
            hk::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                          __FILE__, __LINE__);                                         \
            return 1000 + (int)_e;                                                     \
        }                                                                              \
    } while (0)
#define HK_ARG(cond, ...)                  \
    do {                                   \
        if (!(cond)) {                     \
            hk::set_error(__VA_ARGS__);    \
            return -1;                     \
        }                                  \
    } while (0)

constexpr int HK_MAX_RANKS = 16;

// ---- handle ------------------------------------------------------------------------------------
struct Handle {
    int device = 0;
    int num_sms = 0;
    int smem_optin = 0;       // max dynamic smem per block (opt-in)
    // per-CTA partial sums  [grid][k*(d+1)] doubles (grown on demand)
    double* part = nullptr;
    size_t part_bytes = 0;
    // scratch for the fused step: reduced partials [k*(d+1)] doubles
    double* red = nullptr;
    size_t red_bytes = 0;
    // tensor-core path scratch (TMA descriptors etc.)
    void* tc_scratch = nullptr;
    size_t tc_scratch_bytes = 0;
    // cold-path counters of the tensor-core passes since the last hk_stats_read (device, 6 x u64)
    unsigned long long* stats = nullptr;
    // communicator
    void* nccl_comm = nullptr;
    int nranks = 1;
    int rank = 0;
    // peer-memory exchange of the fused finish (hk_finalize.cu): this rank's mailbox [2][nranks][peer_cap] doubles +
    // flags [2][nranks] u32 in ONE allocation that is IPC-mapped by every peer; peer_mbox[r] / peer_flags[r] are the
    // addresses of rank r's copy in THIS process (own entry = local pointers).  peer_ready: all peers are mapped.
    void* peer_base = nullptr;
    size_t peer_cap = 0;
    double* peer_mbox[HK_MAX_RANKS] = {};
    uint32_t* peer_flags[HK_MAX_RANKS] = {};
    void* peer_opened[HK_MAX_RANKS] = {};
    bool peer_ready = false;
    // ticket + executed-exchange counter of the finish kernel (device, 2 x u32)
    unsigned int* fin_sync = nullptr;
    // CUDA graph of the last hk_lloyd_run call shape
    cudaGraphExec_t run_graph = nullptr, run_graph1 = nullptr;
    cudaStream_t cap_stream = nullptr;
    bool no_graph = false;
    int64_t graph_builds = 0, graph_launches = 0;
    std::vector<uint64_t> run_graph_key;
    int64_t launches = 0;
    std::string variant;
    // optional event timing of the dominant kernel
    bool profile = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
};

// RAII-less helpers: bracket the dominant kernel with events when profiling is on
inline void prof_begin(Handle* h, cudaStream_t st) {
    if (!h->profile) return;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, st);
    h->prof_events.emplace_back(a, b);
}
inline void prof_end(Handle* h, cudaStream_t st) {
    if (!h->profile || h->prof_events.empty()) return;
    cudaEventRecord(h->prof_events.back().second, st);
}

int ensure_part(Handle* h, size_t bytes);
int ensure_red(Handle* h, size_t bytes);
int ensure_stats(Handle* h);

// ---- kernel launchers (each returns an hk error code) -------------------------------------------
// filled by the tensor-core pass when the caller asks it to leave the cross-CTA reduction to the finish kernel
struct SlotInfo {
    const double* fsum = nullptr;
    const double* fcnt = nullptr;
    int nslots = 0, slot_stride = 0, nblocks = 0;
};


struct LloydArgs {
    const void* X;
    int64_t n;
    int d;
    int64_t ldx;
    int dtype;
    const void* C;
    int k;
    void* labels;
    int label_kind;
    double* partials;  // k*(d+1) doubles out (nullptr for assign-only)
    double* fv_out;    // optional: sum over rows of min d^2
    const int32_t* state;  // optional: state[0] != 0 -> the pass is skipped (fit already converged)
    void* row_ws;          // optional caller-owned per-matrix workspace (hk_row_ws_bytes), zeroed when X changes
    int64_t row_ws_bytes;
    SlotInfo* slots;       // optional: the pass may leave its per-CTA slots unreduced and describe them here
    int path;
    cudaStream_t stream;
};

int launch_lloyd_simt(Handle* h, const LloydArgs& a);
int launch_lloyd_tc(Handle* h, const LloydArgs& a);   // fp32 tensor-core path; -2 if shape unsupported
bool tc_supported(const Handle* h, const LloydArgs& a);
int launch_lloyd_row128(Handle* h, const LloydArgs& a);  // exact FMA, rows of exactly 128 bytes
bool row128_supported(const Handle* h, const LloydArgs& a);
int launch_lloyd_dmma(Handle* h, const LloydArgs& a);  // fp64, d = 16, k <= 16: FP64 tensor cores, register accumulators
bool dmma_supported(const Handle* h, const LloydArgs& a);
int launch_lloyd_bigk(Handle* h, const LloydArgs& a);  // fp32, k too large for the fused kernel: multi-pass
bool bigk_supported(const Handle* h, const LloydArgs& a);

// one launch for slot reduce -> cross-GPU sum over peer memory -> finalize (hk_finalize.cu)
struct FinishParams {
    // this rank's contribution: per-CTA slots of the tensor-core pass (fsum != nullptr) or a reduced vector
    const double* fsum;
    const double* fcnt;
    int nslots, slot_stride, nblocks;
    const double* partials_in;
    int k, d;
    double* red;           // [k*(d+1)] this rank's reduced partials
    double* partials_out;  // optional: the globally reduced partials
    // exchange (nranks == 1: none)
    int nranks, rank;
    double* mbox[HK_MAX_RANKS];
    uint32_t* flags[HK_MAX_RANKS];
    size_t cap;
    unsigned int* ticket;
    unsigned int* epoch;
    // finalize
    const void* C_in;
    void* C_out;
    void* C_prev;
    int use_tol;
    double tol_cmp;
    void* shift2_out;
    int32_t* state;
};
int launch_finish(Handle* h, FinishParams& p, int dtype, cudaStream_t stream);

int launch_finalize(Handle* h, const double* partials, const void* C_in, void* C_out, void* C_prev,
                    int k, int d, int dtype, int use_tol, double tol_cmp, void* shift2_out,
                    int32_t* state, cudaStream_t stream);

// hk_kcluster.cu — L1 assignment, radix selection of per-cluster medians, nearest rows, top-k + vote
int launch_assign_l1(Handle* h, const void* X, int64_t n, int d, int64_t ldx, int dtype, const void* C, int k,
                     void* labels, int label_kind, double* fv, cudaStream_t st);
int launch_row_keep(Handle* h, const void* X, int64_t n, int d, int64_t ldx, int dtype, uint8_t* keep, cudaStream_t st);
int launch_select_hist(Handle* h, const void* X, int64_t n, int d, int64_t ldx, int dtype, const int64_t* labels,
                       const uint8_t* keep, int k, const uint64_t* prefix, int pass, unsigned long long* hist,
                       cudaStream_t st);
int launch_select_step(Handle* h, const unsigned long long* hist, int64_t* remaining, uint64_t* prefix, int entries,
                       cudaStream_t st);
int launch_select_value(Handle* h, const uint64_t* prefix, const double* frac, int k, int d, int dtype, void* out,
                        cudaStream_t st);
int launch_kmex_update(Handle* h, const double* partials, const void* medians, const int64_t* counts, void* C, int k,
                       int d, int dtype, double atol, double rtol, int* flag, cudaStream_t st);
int launch_nearest_rows_l1(Handle* h, const void* X, int64_t n, int d, int64_t ldx, int dtype, const void* P, int k,
                           int64_t row_base, double* out_d, int64_t* out_i, cudaStream_t st);
int launch_topk_rows(Handle* h, const void* D, int64_t m, int64_t n, int64_t ldd, int dtype, int kk, void* vals,
                     int64_t* idx, cudaStream_t st);
int launch_knn_vote(Handle* h, const int64_t* idx, int64_t m, int kk, const void* Y, int64_t n, int nc, int64_t ldy,
                    int dtype, int64_t* classes, cudaStream_t st);

// body: 0 = sum (x-y)^2, 1 = quadratic expansion, 2 = sum |x-y|; post: 0 none, 1 sqrt, 2 exp(-v / gden)
int launch_cdist(Handle* h, const void* X, int64_t m, int f, int64_t ldx, const void* Y, int64_t n,
                 int64_t ldy, void* out, int64_t ldo, int dtype, int body, int post, cudaStream_t stream,
                 double gden = 1.0);

int comm_allreduce_f64(Handle* h, double* buf, int64_t count, cudaStream_t stream);

// ---- small device helpers ----------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    // try_wait suspends the thread in hardware for up to the hinted time before it reports "not yet", so the
    // loop below is re-entered rarely instead of burning issue slots (same hint CUTLASS uses)
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(0x989680u)
        : "memory");
}
// 1-D bulk async copy global -> shared, completion signalled on an mbarrier (TMA engine, UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

}  // namespace hk
