// TMA tensor maps + tcgen05/TMEM PTX wrappers (sm_100a).  No libcuda link: the driver entry point for
// cuTensorMapEncodeTiled is fetched through the runtime.
#pragma once
#include <cuda.h>

#include "hk_common.cuh"

namespace hk {

// 2-D row-major fp32/fp64 matrix [rows][cols] (row pitch ld elements) with a [box_cols x box_rows] box and
// the given swizzle.  Out-of-bounds box rows/cols are zero-filled by the hardware.
int make_tensor_map_2d(CUtensorMap* map, const void* base, int elem_bytes, uint64_t rows, uint64_t cols,
                       uint64_t ld, uint32_t box_cols, uint32_t box_rows, int swizzle_bytes);

// ---- device-side wrappers ----------------------------------------------------------------------------
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;  // L2 cache hint: streaming data, read once
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(hint)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// one lane of a fully converged warp (elect.sync): code around it stays warp-uniform for the compiler
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.b32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// TMEM allocation: one warp, power-of-two column count >= 32; base address lands in shared memory
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], TF32 inputs (fp32 containers, low 13 mantissa bits ignored), fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
          "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
          "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

// single-column TMEM accesses of the accumulator warps (lane = this thread's TMEM lane).  No "memory" clobber:
// they touch nothing the compiler can see, and ordinary shared-memory loads may be scheduled across them.
__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr));
    return __uint_as_float(v);
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, float v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(__float_as_uint(v)));
}
__device__ __forceinline__ void tmem_st8_zero(uint32_t taddr) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(0u));
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;"); }
// the loaded registers are operands of the wait, so their consumers cannot be scheduled above it
__device__ __forceinline__ void tmem_wait_ld4(float& a, float& b, float& c, float& d) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(a), "+f"(b), "+f"(c), "+f"(d));
}
__device__ __forceinline__ void tmem_wait_ld1(float& a) { asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(a)); }

// shared-memory matrix descriptor: K-major operand, 128-byte swizzle, 8-row groups 1024 B apart
// (cute::UMMA::SmemDescriptor bit layout: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version [46,48),
//  layout type [61,64) with SWIZZLE_128B = 2)
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulator, TF32 A and B, both K-major
__host__ __device__ inline uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// byte offset of element (row r, feature f) inside a K-blocked, 128B-swizzled fp32 tile of `rows` rows:
// K-block kb = f/32 holds [rows][32 floats]; inside a row the 16-byte chunk index is XORed with (r & 7)
__device__ __forceinline__ uint32_t sw128_off(int rows, int r, int f) {
    const int kb = f >> 5, fi = f & 31;
    return (uint32_t)(kb * rows * 128 + r * 128 + ((((fi >> 2) ^ (r & 7)) << 4) | ((fi & 3) << 2)));
}

// ---- 32-bit shared-window accessors (keep the hot loops free of 64-bit address arithmetic) -------------
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
// (volatile asm statements keep their relative order; the private-accumulator traffic needs no more than that)
__device__ __forceinline__ void sts_f4_nc(uint32_t a, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ int lds_s32(uint32_t a) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_s32(uint32_t a, int v) {
    asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
#ifdef HK_WATCHDOG
    // debugging aid (-DHK_WATCHDOG): a wait that lasts longer than ~2 s reports who is stuck on what and kills the kernel
    // instead of hanging the GPU
    const long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.b32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity), "r"(0x989680u)
            : "memory");
        if (ok) return;
        if (clock64() - t0 > 4000000000ll) {
            printf("hk watchdog: block %d warp %d lane %d stuck on mbarrier smem+0x%x parity %u\n", (int)blockIdx.x,
                   (int)(threadIdx.x >> 5), (int)(threadIdx.x & 31), bar, parity);
            __trap();
        }
    }
#elif defined(HK_MBAR_SPIN)
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
#else
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity), "r"(0x989680u)
        : "memory");
#endif
}
// one lane waits, then the warp is released (keeps 31 lanes out of the wait loop)
__device__ __forceinline__ void warp_wait(uint32_t bar, uint32_t parity, int lane) {
    if (lane == 0) mbar_wait_a(bar, parity);
    __syncwarp();
}

__device__ __forceinline__ double2 lds_d2(uint32_t a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_d2_nc(uint32_t a, double2 v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.x), "d"(v.y));
}

// ---- register-array scans shared by the epilogues (32 TMEM columns at a time) ---------------------------
// (a0 - t, a1 - t) with one packed FP32x2 add (sm_100 FADD2)
__device__ __forceinline__ void sub2(uint32_t a0, uint32_t a1, uint64_t negthr2, uint32_t& r0, uint32_t& r1) {
    uint64_t in, out;
    asm("mov.b64 %0, {%1, %2};" : "=l"(in) : "r"(a0), "r"(a1));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(out) : "l"(in), "l"(negthr2));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(r0), "=r"(r1) : "l"(out));
}
// sign-bit mask of (s_j < thr) for 32 accumulator columns: bit j <-> column j (4 independent shift chains)
__device__ __forceinline__ unsigned below_mask32(const uint32_t* a, float thr) {
    const uint32_t nt = __float_as_uint(-thr);
    uint64_t negthr2;
    asm("mov.b64 %0, {%1, %1};" : "=l"(negthr2) : "r"(nt));
    unsigned m0 = 0, m1 = 0, m2 = 0, m3 = 0;
#pragma unroll
    for (int j = 6; j >= 0; j -= 2) {
        uint32_t t0, t1;
        sub2(a[j], a[j + 1], negthr2, t0, t1);
        m0 = __funnelshift_l(t1, m0, 1);
        m0 = __funnelshift_l(t0, m0, 1);
        sub2(a[8 + j], a[9 + j], negthr2, t0, t1);
        m1 = __funnelshift_l(t1, m1, 1);
        m1 = __funnelshift_l(t0, m1, 1);
        sub2(a[16 + j], a[17 + j], negthr2, t0, t1);
        m2 = __funnelshift_l(t1, m2, 1);
        m2 = __funnelshift_l(t0, m2, 1);
        sub2(a[24 + j], a[25 + j], negthr2, t0, t1);
        m3 = __funnelshift_l(t1, m3, 1);
        m3 = __funnelshift_l(t0, m3, 1);
    }
    return m0 | (m1 << 8) | (m2 << 16) | (m3 << 24);
}
__device__ __forceinline__ float min32(const uint32_t* a) {
    float m0 = __uint_as_float(a[0]), m1 = __uint_as_float(a[1]), m2 = __uint_as_float(a[2]),
          m3 = __uint_as_float(a[3]);
#pragma unroll
    for (int j = 4; j < 32; j += 4) {
        m0 = fminf(m0, __uint_as_float(a[j]));
        m1 = fminf(m1, __uint_as_float(a[j + 1]));
        m2 = fminf(m2, __uint_as_float(a[j + 2]));
        m3 = fminf(m3, __uint_as_float(a[j + 3]));
    }
    return fminf(fminf(m0, m1), fminf(m2, m3));
}

// K-major SWIZZLE_128B descriptor whose 8-row groups all alias the same 1 KB (stride byte offset 0)
__device__ __forceinline__ uint64_t umma_desc_k_sw128_bcast(uint32_t smem_addr) {
    uint64_t dsc = 0;
    dsc |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    dsc |= (uint64_t)1 << 16;
    dsc |= (uint64_t)1 << 46;
    dsc |= (uint64_t)2 << 61;
    return dsc;
}

}  // namespace hk
