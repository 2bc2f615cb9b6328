// Tensor-core pairwise distances (fp32, quadratic expansion) for sm_100a — BASELINE config 2
// (X 1M x 64 vs Y 4096 x 64): the output write (4*m*n bytes) is the roofline, the contraction 2*m*n*f is three
// times what the FP32 pipes can do in that time, so x.y^T runs on tcgen05 as 3xTF32:
//     x.y ~= xh.yh + xl.yh + xh.yl,   xh = x with the low 13 mantissa bits cleared (what kind::tf32 reads from the
//     raw fp32 tile), xl = x - xh (exact), error ~1e-6 * |x||y| (the reference's sgemm: ~5e-7 * |x||y|)
// A small pre-pass writes xl / yl and the row norms; the main kernel is a warp-specialised GEMM with a fused
// epilogue  out = sqrt(clamp(|x|^2 + |y|^2 - 2 x.y, 0, inf))  (heat/spatial/distance.py:44, 59-64):
//   warp 0      : TMA producer — A tiles (raw X and xl, 128 rows x f) once per row tile, B slots (raw Y / yl,
//                 128 rows x 32 features = 16 KB) through a ring
//   warp 1      : tcgen05.mma issuer (M=128, N=128, K=8), 4 accumulator buffers of 128 columns in TMEM
//   warps 4-11  : epilogue (thread == output row == TMEM lane): tcgen05.ld 32 columns (the next slice's load in
//                 flight) -> distances -> this warp's swizzled 32 x 32 staging tile (double-buffered) -> this warp's
//                 own TMA store (cp.async.bulk.tensor, clipped at the matrix edge); no block-level barrier
// Replaces cdist -> _dist -> _euclidian_fast on the local blocks (heat/spatial/distance.py:32-64, 409-414).
#include <math.h>

#include "hk_tma.cuh"

namespace hk {
namespace {

constexpr int TM = 128;   // output rows per tile (UMMA M)
constexpr int TN = 128;   // output columns per accumulator (UMMA N)
constexpr int NBUF = 4;   // TMEM accumulator buffers (4 x 128 columns)
constexpr int NSLOT = 6;  // B ring slots of 16 KB (barrier layout; the ring itself uses p.nslot <= NSLOT of them)
constexpr int EPI_WARPS = 8;
constexpr int NTHREADS = (4 + EPI_WARPS) * 32;

struct CdParams {
    int64_t m, n;
    int f, nkb;
    int num_row_tiles, num_chunks;
    const float* xn;
    const float* yn;
    int post;      // 0 = squared distances, 1 = sqrt, 2 = 2^(d2 * gscale) (Gaussian kernel)
    float gscale;  // -log2(e) / (2 sigma^2)
    int nslot;  // B ring slots in use (2..NSLOT)
    int nstg;   // output staging buffers per epilogue warp (1 or 2)
    // ARGMIN variant (large-k Lloyd pass): no distance matrix is written; per row the first-index argmin over all n
    // columns goes to labels[], rows whose runner-up is within window * (|x|^2 + *cmax2) are appended to queue[]
    int32_t* labels;
    int32_t* queue;
    int* qcount;
    const float* cmax2;
    float window;
    int64_t row_base;  // global index of row 0 (labels / queue entries are global)
    const int32_t* state;  // optional: state[0] != 0 -> the pass is skipped (fit already converged)
    uint32_t o_A, o_B, o_stage, o_bars;
};

// xl = x - trunc_tf32(x), row norms (ascending feature order inside each lane, then a fixed shuffle tree)
__global__ void split_lo_kernel(const float* __restrict__ A, int64_t rows, int f, int64_t ld, float* __restrict__ lo,
                                float* __restrict__ nrm) {
    const int lpr = f >> 2;  // lanes per row (8, 16, 24, 32)
    const int rpw = 32 / lpr > 0 ? 32 / lpr : 1;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int sub = lane / lpr, li = lane - sub * lpr;
    const int64_t r = warp * rpw + sub;
    float s = 0.f;
    if (sub < rpw && r < rows) {
        const float4 v = *reinterpret_cast<const float4*>(A + r * ld + (li << 2));
        float4 l;
        l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
        l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
        l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
        l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
        *reinterpret_cast<float4*>(lo + r * f + (li << 2)) = l;
        s = fmaf(v.x, v.x, s);
        s = fmaf(v.y, v.y, s);
        s = fmaf(v.z, v.z, s);
        s = fmaf(v.w, v.w, s);
    }
    // reduce over the lanes of a row (lpr is 8, 16 or 32 when it divides 32; 24 lanes: one row per warp)
    if (lpr == 24) {
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    } else {
        for (int o = lpr >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    }
    if (sub < rpw && r < rows && li == 0) nrm[r] = s;
}

// one MUFU instead of the ~8-instruction IEEE sequence: the 3xTF32 product already carries ~1e-6 relative error,
// the epilogue is issue-bound, and the result stays within the parity tolerance of the exact path
__device__ __forceinline__ float sqrt_fast(float v) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

// packed fp32 pairs (add/fma.rn.f32x2: one issue slot for two results, each rounded like the scalar instruction)
__device__ __forceinline__ uint64_t pack2(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
// (|x|^2 + |y|^2) - 2 * dot for two adjacent columns
__device__ __forceinline__ void dist2_pair(float& o0, float& o1, uint32_t d0, uint32_t d1, uint64_t xn2, float y0, float y1) {
    uint64_t dv, s, r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(dv) : "r"(d0), "r"(d1));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(s) : "l"(xn2), "l"(pack2(y0, y1)));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(dv), "l"(pack2(-2.f, -2.f)), "l"(s));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(o0), "=f"(o1) : "l"(r));
}
// max(v, 0) that keeps NaN (v < 0 ? 0 : v) in one instruction
__device__ __forceinline__ float clamp0_nan(float v) {
    float r;
    asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(r) : "f"(v));
    return r;
}

// tcgen05.wait::ld with the loaded registers as operands: their consumers cannot be scheduled above the wait, so the
// next slice's tcgen05.ld may stay in flight while this one is processed
__device__ __forceinline__ void tmem_wait_ld32(uint32_t (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                   "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]),
                   "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]),
                   "+r"(v[30]), "+r"(v[31])
                 :
                 : "memory");
}

__device__ __forceinline__ float exp2_fast(float v) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

template <bool ARGMIN>
__global__ void __launch_bounds__(NTHREADS, 1)
    cdist_tc_kernel(const __grid_constant__ CUtensorMap xh_map, const __grid_constant__ CUtensorMap xl_map,
                    const __grid_constant__ CUtensorMap yh_map, const __grid_constant__ CUtensorMap yl_map,
                    const __grid_constant__ CUtensorMap out_map, const CdParams p) {
    extern __shared__ unsigned char smem_raw[];
    if (ARGMIN && p.state != nullptr && p.state[0] != 0) return;  // uniform across the grid
    unsigned char* smem =
        reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const int nkb = p.nkb;
    const uint32_t a_A = sbase + p.o_A;        // [2 parts][nkb][128 rows x 128 B]
    const uint32_t a_B = sbase + p.o_B;        // [NSLOT][128 rows x 128 B]
    const uint32_t a_stage = sbase + p.o_stage;  // [8 warps][nstg][32 rows x 128 B] output staging, 128B swizzle
    // barriers: a_full | a_empty | b_full[NSLOT] | b_empty[NSLOT] | t_full[NBUF] | t_empty[NBUF]
    const uint32_t b_afull = sbase + p.o_bars;
    const uint32_t b_aempty = b_afull + 8;
    const uint32_t b_bfull = b_afull + 16;
    const uint32_t b_bempty = b_bfull + NSLOT * 8;
    const uint32_t b_tfull = b_bempty + NSLOT * 8;
    const uint32_t b_tempty = b_tfull + NBUF * 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + p.o_bars + 256);

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform for the compiler (uniform-register role code)
    const int lane = tid & 31;
    const uint32_t kb_bytes = TM * 128;
    const uint32_t a_bytes = (uint32_t)(2 * nkb) * kb_bytes;

    if (tid == 0) {
        uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.o_bars);
        mbar_init(bars + 0, 1);
        mbar_init(bars + 1, 1);
        for (int i = 0; i < NSLOT; ++i) {
            mbar_init(bars + 2 + i, 1);
            mbar_init(bars + 2 + NSLOT + i, 1);
        }
        for (int i = 0; i < NBUF; ++i) {
            mbar_init(bars + 2 + 2 * NSLOT + i, 1);
            mbar_init(bars + 2 + 2 * NSLOT + NBUF + i, 4);
        }
        mbar_fence_init();
        tma_prefetch_desc(&xh_map);
        tma_prefetch_desc(&xl_map);
        tma_prefetch_desc(&yh_map);
        tma_prefetch_desc(&yl_map);
        tma_prefetch_desc(&out_map);
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        // convergent warp, one elected lane issues (addresses and descriptors stay in uniform registers)
        {
            uint32_t aph = 0;
            int slot = 0;
            uint32_t sph = 0;
            for (int rt = blockIdx.x; rt < p.num_row_tiles; rt += gridDim.x) {
                mbar_wait_a(b_aempty, aph ^ 1);
                if (elect_one()) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b_afull), "r"(a_bytes)
                             : "memory");
                for (int kb = 0; kb < nkb; ++kb) {
                    asm volatile(
                        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
                        " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(a_A + kb * kb_bytes),
                        "l"(&xh_map), "r"(b_afull), "r"(kb * 32), "r"(rt * TM), "l"(kEvictFirst)
                        : "memory");
                    asm volatile(
                        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
                        " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(a_A + (nkb + kb) * kb_bytes),
                        "l"(&xl_map), "r"(b_afull), "r"(kb * 32), "r"(rt * TM), "l"(kEvictFirst)
                        : "memory");
                }
                }
                __syncwarp();
                aph ^= 1;
                for (int c = 0; c < p.num_chunks; ++c) {
                    for (int kb = 0; kb < nkb; ++kb) {
#pragma unroll
                        for (int part = 0; part < 2; ++part) {
                            mbar_wait_a(b_bempty + slot * 8, sph ^ 1);
                            if (elect_one()) {
                            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b_bfull + slot * 8),
                                         "r"(kb_bytes)
                                         : "memory");
                            asm volatile(
                                "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
                                " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(a_B + slot * kb_bytes),
                                "l"(part == 0 ? &yh_map : &yl_map), "r"(b_bfull + slot * 8), "r"(kb * 32), "r"(c * TN),
                                "l"(kEvictLast)
                                : "memory");
                            }
                            __syncwarp();
                            if (++slot == p.nslot) {
                                slot = 0;
                                sph ^= 1;
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // convergent warp: waits by all lanes, tcgen05 instructions by one elected lane
        {
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
            const uint32_t idesc = umma_idesc_tf32(TM, TN);
            uint32_t aph = 0;
            int slot = 0;
            uint32_t sph = 0;
            int buf = 0;
            uint32_t tph = 0;
            for (int rt = blockIdx.x; rt < p.num_row_tiles; rt += gridDim.x) {
                mbar_wait_a(b_afull, aph);
                aph ^= 1;
                for (int c = 0; c < p.num_chunks; ++c) {
                    mbar_wait_a(b_tempty + buf * 8, tph ^ 1);
                    tc_fence_after();
                    const uint32_t dcol = tmem_u + (uint32_t)(buf * TN);
                    for (int kb = 0; kb < nkb; ++kb) {
                        // slot 0: raw Y block (read as yh) against xh and xl; slot 1: yl block against xh
                        mbar_wait_a(b_bfull + slot * 8, sph);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t bb = a_B + slot * kb_bytes;
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                const uint64_t bd = umma_desc_k_sw128(bb + ks * 32);
                                umma_tf32(dcol, umma_desc_k_sw128(a_A + kb * kb_bytes + ks * 32), bd, idesc,
                                          (kb | ks) != 0 ? 1u : 0u);
                                umma_tf32(dcol, umma_desc_k_sw128(a_A + (nkb + kb) * kb_bytes + ks * 32), bd, idesc, 1u);
                            }
                            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                             b_bempty + slot * 8)
                                         : "memory");
                        }
                        __syncwarp();
                        if (++slot == p.nslot) {
                            slot = 0;
                            sph ^= 1;
                        }
                        mbar_wait_a(b_bfull + slot * 8, sph);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t bb = a_B + slot * kb_bytes;
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                umma_tf32(dcol, umma_desc_k_sw128(a_A + kb * kb_bytes + ks * 32),
                                          umma_desc_k_sw128(bb + ks * 32), idesc, 1u);
                            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                             b_bempty + slot * 8)
                                         : "memory");
                        }
                        __syncwarp();
                        if (++slot == p.nslot) {
                            slot = 0;
                            sph ^= 1;
                        }
                    }
                    if (elect_one())
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                         b_tfull + buf * 8)
                                     : "memory");
                    __syncwarp();
                    if (++buf == NBUF) {
                        buf = 0;
                        tph ^= 1;
                    }
                }
                // the A tiles may be overwritten once every MMA of this row tile has completed
                if (elect_one())
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b_aempty)
                                 : "memory");
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue warps =================
        const int we = warp - 4;
        const int q = we & 3;    // TMEM lane quarter == warp % 4
        const int grp = we >> 2;  // two groups of four warps alternate over the column chunks
        const int row = q * 32 + lane;
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t wstage = a_stage + (uint32_t)(we * p.nstg) * 4096u;  // this warp's nstg staging buffers of 4 KB
        int sb = 0;
        int lc = 0;  // chunk counter of this CTA (all row tiles)
        for (int rt = blockIdx.x; rt < p.num_row_tiles; rt += gridDim.x) {
            const int64_t grow = (int64_t)rt * TM + row;
            const float xn = grow < p.m ? __ldg(p.xn + grow) : 0.f;
            // ARGMIN: running minimum over this group's chunks (same scheme as the fused Lloyd epilogue)
            float m_best = INFINITY, m_second = INFINITY;
            unsigned mk_best = 0;
            int c_best = 0;
            const float thr = ARGMIN ? p.window * (xn + __ldg(p.cmax2)) : 0.f;
            for (int c = 0; c < p.num_chunks; ++c, ++lc) {
                if ((lc & 1) != grp) continue;
                const int buf = lc & (NBUF - 1);
                const uint32_t tph = (uint32_t)((lc / NBUF) & 1);
                warp_wait(b_tfull + buf * 8, tph, lane);
                tc_fence_after();
                if constexpr (!ARGMIN) {
                    const uint64_t xn2 = pack2(xn, xn);
                    // Per warp: 32 rows x 32 columns (4 KB) per slice through this warp's own staging buffers and its
                    // own TMA stores; the next slice's tcgen05.ld is in flight while this one is converted.
                    uint32_t a[2][32];
                    tmem_ld32(tlane + (uint32_t)(buf * TN), a[0]);
#pragma unroll
                    for (int sl = 0; sl < TN / 32; ++sl) {
                        uint32_t(&v)[32] = a[sl & 1];
                        tmem_wait_ld32(v);
                        if (sl + 1 < TN / 32) {
                            tmem_ld32(tlane + (uint32_t)(buf * TN + (sl + 1) * 32), a[(sl + 1) & 1]);
                        } else {
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_a(b_tempty + buf * 8);  // accumulator drained by this warp
                        }
                        const int col0 = c * TN + sl * 32;
                        // the TMA store that last read this staging buffer must have finished reading it
                        if (lane == 0) {
                            if (p.nstg == 2)
                                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                            else
                                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        }
                        __syncwarp();
                        const uint32_t sbuf = wstage + (uint32_t)sb * 4096u;
                        const uint32_t srow = sbuf + (uint32_t)lane * 128u;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float4 yv = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (col0 + j < p.n) yv = __ldg(reinterpret_cast<const float4*>(p.yn + col0 + j));
                            float4 o;
                            dist2_pair(o.x, o.y, v[j + 0], v[j + 1], xn2, yv.x, yv.y);
                            dist2_pair(o.z, o.w, v[j + 2], v[j + 3], xn2, yv.z, yv.w);
                            o.x = clamp0_nan(o.x);
                            o.y = clamp0_nan(o.y);
                            o.z = clamp0_nan(o.z);
                            o.w = clamp0_nan(o.w);
                            if (p.post == 1) {
                                o.x = sqrt_fast(o.x);
                                o.y = sqrt_fast(o.y);
                                o.z = sqrt_fast(o.z);
                                o.w = sqrt_fast(o.w);
                            } else if (p.post == 2) {
                                o.x = exp2_fast(o.x * p.gscale);
                                o.y = exp2_fast(o.y * p.gscale);
                                o.z = exp2_fast(o.z * p.gscale);
                                o.w = exp2_fast(o.w * p.gscale);
                            }
                            // staging tile: 32 rows x 128 B, 16-byte chunk index XOR (row & 7) (matches the store map)
                            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(
                                             srow + (uint32_t)((((j >> 2) ^ (lane & 7))) << 4)),
                                         "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w)
                                         : "memory");
                        }
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                                             &out_map),
                                         "r"(sbuf), "r"(col0), "r"(rt * TM + q * 32)
                                         : "memory");
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                        sb = (sb + 1 == p.nstg) ? 0 : sb + 1;
                    }
                } else {
#pragma unroll 1
                for (int sl = 0; sl < TN / 32; ++sl) {
                    uint32_t a[32];
                    tmem_ld32(tlane + (uint32_t)(buf * TN + sl * 32), a);
                    tmem_wait_ld();
                    if (sl == TN / 32 - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_a(b_tempty + buf * 8);  // accumulator drained by this warp
                    }
                    const int col0 = c * TN + sl * 32;
                    // d^2 of 32 columns in place, then chunk minimum + mask of the columns within thr of it
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 yv = make_float4(INFINITY, INFINITY, INFINITY, INFINITY);  // columns >= n never win
                        if (col0 + j < p.n) yv = __ldg(reinterpret_cast<const float4*>(p.yn + col0 + j));
                        float o0 = fmaf(-2.f, __uint_as_float(a[j + 0]), xn + yv.x);
                        float o1 = fmaf(-2.f, __uint_as_float(a[j + 1]), xn + yv.y);
                        float o2 = fmaf(-2.f, __uint_as_float(a[j + 2]), xn + yv.z);
                        float o3 = fmaf(-2.f, __uint_as_float(a[j + 3]), xn + yv.w);
                        a[j + 0] = __float_as_uint(o0 < 0.f ? 0.f : o0);
                        a[j + 1] = __float_as_uint(o1 < 0.f ? 0.f : o1);
                        a[j + 2] = __float_as_uint(o2 < 0.f ? 0.f : o2);
                        a[j + 3] = __float_as_uint(o3 < 0.f ? 0.f : o3);
                    }
                    const float mc = min32(a);
                    const unsigned mk = below_mask32(a, mc + thr);
                    if (mc < m_best) {
                        m_second = m_best;
                        m_best = mc;
                        mk_best = mk;
                        c_best = col0;
                    } else {
                        m_second = fminf(m_second, mc);
                    }
                }
                }
            }
            if (ARGMIN) {
                // group 1 hands its state to group 0 through the (otherwise unused) staging tile; group 0 merges,
                // stores the label and queues undecided rows.  NaN minima compare false everywhere: m_best stays
                // +inf with an empty mask, the row is undecided and the exact kernel takes it.
                const bool own_ok = (m_second >= m_best + thr) && (__popc(mk_best) == 1);
                const int own_idx = c_best + __ffs(mk_best) - 1;
                const uint32_t xch = a_stage + (uint32_t)row * 16;
                named_bar_sync(3, 256);  // the previous row tile's exchange has been read
                if (grp == 1) {
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(xch), "f"(m_best), "f"(__int_as_float(own_idx)),
                                 "f"(own_ok ? 1.f : 0.f), "f"(0.f)
                                 : "memory");
                }
                named_bar_sync(3, 256);
                if (grp == 0) {
                    float om, oi, ook, pad;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(om), "=f"(oi), "=f"(ook), "=f"(pad) : "r"(xch));
                    const int o_idx = __float_as_int(oi);
                    const bool take_other = om < m_best;  // equal minima: undecided below, the index does not matter
                    const float best = take_other ? om : m_best;
                    const float lose = take_other ? m_best : om;
                    const bool ok = (take_other ? (ook != 0.f) : own_ok) && (lose >= best + thr);
                    int lab = take_other ? o_idx : own_idx;
                    if (lab < 0 || lab >= (int)p.n) lab = 0;  // undecided rows only (empty mask)
                    if (grow < p.m) {
                        p.labels[p.row_base + grow] = lab;
                        if (!ok) p.queue[atomicAdd(p.qcount, 1)] = (int32_t)(p.row_base + grow);
                    }
                }
            }
        }
        if (!ARGMIN && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

struct CdLayout {
    size_t A, B, stage, bars, total;
};
CdLayout cd_layout(int nkb, int nslot, int nstg) {
    CdLayout L;
    size_t o = 0;
    L.A = o;
    o += (size_t)2 * nkb * TM * 128;
    L.B = o;
    o += (size_t)nslot * TM * 128;
    L.stage = o;
    o += (size_t)nstg * EPI_WARPS * 4096;  // per epilogue warp: nstg buffers of 32 rows x 128 B
    L.bars = o;
    o += 512;
    L.total = o + 1024;
    return L;
}

// largest B ring (2..NSLOT slots) that fits next to the A tiles
int cd_nslot(const Handle* h, int nkb, int nstg) {
    for (int ns = NSLOT; ns >= 2; --ns)
        if (cd_layout(nkb, ns, nstg).total <= (size_t)h->smem_optin) return ns;
    return 0;
}
// double-buffered output staging when the B ring keeps at least 4 slots next to it
int cd_nstg(const Handle* h, int nkb) { return cd_nslot(h, nkb, 2) >= 4 ? 2 : 1; }

}  // namespace

bool cdist_tc_supported(const Handle* h, const void* X, int64_t m, int f, int64_t ldx, const void* Y, int64_t n,
                        int64_t ldy, const void* out, int64_t ldo) {
    if (f % 32 != 0 || f < 32 || f > 128) return false;
    if (ldx % 4 != 0 || ldy % 4 != 0 || ldo % 4 != 0 || n % 4 != 0) return false;
    if ((reinterpret_cast<uintptr_t>(X) & 15) || (reinterpret_cast<uintptr_t>(Y) & 15) ||
        (reinterpret_cast<uintptr_t>(out) & 15))
        return false;
    if (m >= ((int64_t)1 << 31) - TM || n >= ((int64_t)1 << 31) - TN) return false;
    if (m < 1024 || n < 128) return false;  // small problems: the exact-FMA kernel is as good and bit-closer
    return cd_nslot(h, f / 32, 1) >= 2;
}

namespace {
struct ArgminOut {
    int32_t* labels;
    int32_t* queue;
    int* qcount;
    const float* cmax2;
    float window;
    int64_t row_base;
    const int32_t* state;
};
int launch_cdist_tc_impl(Handle* h, const void* X, int64_t m, int f, int64_t ldx, const void* Y, int64_t n, int64_t ldy,
                         void* out, int64_t ldo, int post, float gscale, const ArgminOut* am, cudaStream_t st) {
    const int nkb = f / 32;
    // scratch: xl [m x f], yl [n x f], xn [m], yn [n]
    const size_t need = ((size_t)m * f + (size_t)n * f + (size_t)m + (size_t)n + 64) * sizeof(float);
    int rc = ensure_part(h, need);
    if (rc) return rc;
    float* xl = reinterpret_cast<float*>(h->part);
    float* yl = xl + (size_t)m * f;
    float* xn = yl + (size_t)n * f;
    xn = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(xn) + 15) & ~(uintptr_t)15);
    float* yn = xn + ((m + 3) / 4) * 4;
    {
        const int lpr = f >> 2;
        const int rpw = 32 / lpr > 0 ? 32 / lpr : 1;
        const int64_t wx = (m + rpw - 1) / rpw, wy = (n + rpw - 1) / rpw;
        split_lo_kernel<<<(unsigned)((wx * 32 + 255) / 256), 256, 0, st>>>((const float*)X, m, f, ldx, xl, xn);
        split_lo_kernel<<<(unsigned)((wy * 32 + 255) / 256), 256, 0, st>>>((const float*)Y, n, f, ldy, yl, yn);
        HK_CUDA(cudaGetLastError());
        h->launches += 2;
    }
    CUtensorMap xh_map, xl_map, yh_map, yl_map, out_map;
    rc = make_tensor_map_2d(&xh_map, X, 4, (uint64_t)m, (uint64_t)f, (uint64_t)ldx, 32, TM, 128);
    if (rc) return rc;
    rc = make_tensor_map_2d(&xl_map, xl, 4, (uint64_t)m, (uint64_t)f, (uint64_t)f, 32, TM, 128);
    if (rc) return rc;
    rc = make_tensor_map_2d(&yh_map, Y, 4, (uint64_t)n, (uint64_t)f, (uint64_t)ldy, 32, TN, 128);
    if (rc) return rc;
    rc = make_tensor_map_2d(&yl_map, yl, 4, (uint64_t)n, (uint64_t)f, (uint64_t)f, 32, TN, 128);
    if (rc) return rc;
    if (am == nullptr)
        rc = make_tensor_map_2d(&out_map, out, 4, (uint64_t)m, (uint64_t)n, (uint64_t)ldo, 32, 32, 128);
    else
        out_map = xl_map;  // unused by the ARGMIN kernel
    if (rc) return rc;

    const int nstg = am != nullptr ? 1 : cd_nstg(h, nkb);  // the ARGMIN variant stores nothing (2 KB exchange area)
    const int nslot = cd_nslot(h, nkb, nstg);
    const CdLayout L = cd_layout(nkb, nslot, nstg);
    CdParams p{};
    p.nslot = nslot;
    p.nstg = nstg;
    p.m = m;
    p.n = n;
    p.f = f;
    p.nkb = nkb;
    p.num_row_tiles = (int)((m + TM - 1) / TM);
    p.num_chunks = (int)((n + TN - 1) / TN);
    p.xn = xn;
    p.yn = yn;
    p.post = post;
    p.gscale = gscale;
    p.o_A = (uint32_t)L.A;
    p.o_B = (uint32_t)L.B;
    p.o_stage = (uint32_t)L.stage;
    p.o_bars = (uint32_t)L.bars;
    int grid = h->num_sms;
    if (grid > p.num_row_tiles) grid = p.num_row_tiles;
    if (am != nullptr) {
        p.labels = am->labels;
        p.queue = am->queue;
        p.qcount = am->qcount;
        p.cmax2 = am->cmax2;
        p.window = am->window;
        p.row_base = am->row_base;
        p.state = am->state;
        HK_CUDA(cudaFuncSetAttribute(cdist_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
        prof_begin(h, st);
        cdist_tc_kernel<true><<<grid, NTHREADS, L.total, st>>>(xh_map, xl_map, yh_map, yl_map, out_map, p);
        prof_end(h, st);
    } else {
        HK_CUDA(cudaFuncSetAttribute(cdist_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
        prof_begin(h, st);
        cdist_tc_kernel<false><<<grid, NTHREADS, L.total, st>>>(xh_map, xl_map, yh_map, yl_map, out_map, p);
        prof_end(h, st);
    }
    HK_CUDA(cudaGetLastError());
    h->launches++;
    return 0;
}
}  // namespace

int launch_cdist_tc(Handle* h, const void* X, int64_t m, int f, int64_t ldx, const void* Y, int64_t n, int64_t ldy,
                    void* out, int64_t ldo, int post, float gscale, cudaStream_t st) {
    return launch_cdist_tc_impl(h, X, m, f, ldx, Y, n, ldy, out, ldo, post, gscale, nullptr, st);
}

// distances are consumed in the epilogue: labels[row_base + r] = first-index argmin_j d2(x_r, y_j); rows whose runner-up
// lies within window * (|x|^2 + *cmax2) of the minimum (or that contain NaN) are appended to queue[] (large-k Lloyd pass)
int launch_cdist_tc_argmin(Handle* h, const void* X, int64_t m, int f, int64_t ldx, const void* Y, int64_t n, int64_t ldy,
                           int32_t* labels, int64_t row_base, int32_t* queue, int* qcount, const float* cmax2, float window,
                           const int32_t* state, cudaStream_t st) {
    ArgminOut am{labels, queue, qcount, cmax2, window, row_base, state};
    return launch_cdist_tc_impl(h, X, m, f, ldx, Y, n, ldy, nullptr, 4, 0, 0.f, &am, st);
}

}  // namespace hk
