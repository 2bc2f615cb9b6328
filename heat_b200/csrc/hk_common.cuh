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
    // per-tile |x|^2 bound cache of the tensor-core path, keyed by the matrix it was computed from
    float* xb = nullptr;
    size_t xb_bytes = 0;
    const void* xb_X = nullptr;
    int64_t xb_n = 0;
    int xb_d = 0;
    int64_t xb_ld = 0;
    // communicator
    void* nccl_comm = nullptr;
    int nranks = 1;
    int rank = 0;
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

// ---- kernel launchers (each returns an hk error code) -------------------------------------------
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
    int path;
    cudaStream_t stream;
};

int launch_lloyd_simt(Handle* h, const LloydArgs& a);
int launch_lloyd_tc(Handle* h, const LloydArgs& a);   // fp32 tensor-core path; -2 if shape unsupported
bool tc_supported(const Handle* h, const LloydArgs& a);
int launch_lloyd_row128(Handle* h, const LloydArgs& a);  // exact FMA, rows of exactly 128 bytes
bool row128_supported(const Handle* h, const LloydArgs& a);
int launch_lloyd_bigk(Handle* h, const LloydArgs& a);  // fp32, k too large for the fused kernel: multi-pass
bool bigk_supported(const Handle* h, const LloydArgs& a);

int launch_finalize(Handle* h, const double* partials, const void* C_in, void* C_out, void* C_prev,
                    int k, int d, int dtype, int use_tol, double tol_cmp, void* shift2_out,
                    int32_t* state, cudaStream_t stream);

int launch_cdist(Handle* h, const void* X, int64_t m, int f, int64_t ldx, const void* Y, int64_t n,
                 int64_t ldy, void* out, int64_t ldo, int dtype, int quad, int sqrt_flag,
                 cudaStream_t stream);

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
