// Tensor-core Lloyd pass (fp32 data) for sm_100a: TMA-streamed row tiles, tcgen05 TF32 distance filter with
// fp32 accumulators in TMEM, exact-FMA refinement of the rows the filter cannot decide, and the same
// deterministic in-tile counting sort + segmented column sums as the exact-FMA kernel.
//
// Why a filter: at k=64, d=32 the distance contraction costs 2*k = 128 FLOP per 4-byte element, three
// times what the FP32 pipes can sustain at HBM speed, so x.c^T runs on the 5th-gen tensor cores
// (tcgen05.mma kind::tf32 reads the fp32 tile in shared memory directly and ignores the low 13 mantissa
// bits).  The TF32 result has a rigorous error bound  |s_j - (|c_j|^2 - 2 x.c_j)| <= E(x)  with
//      E = 2*beta*|x|*max_j|c_j| + (d+3)*2^-23*(|x|^2 + max_j|c_j|^2),   beta = 1.05 * 2^-9,
// so a row whose runner-up is more than 2E above the minimum has a certain label (identical to what the
// exact fp32 formula of heat/spatial/distance.py:59-64 + first-index argmin would give); every other row
// (near-ties, NaN/Inf) is re-evaluated with that exact formula over all centroids.  Labels are therefore
// those of the exact-FMA path, the tensor cores only remove work.
//
// Roles per CTA (1 CTA per SM, persistent, static tile -> CTA map):
//   warp 0   : TMA producer (cp.async.bulk.tensor, 128B swizzle, EVICT_FIRST) into an S-stage ring
//   warp 1   : tcgen05.mma issuer (one lane), accumulator buffer g = tile % G in TMEM
//   warp 2   : TMEM allocation / release
//   G groups of 4 warps: tcgen05.ld epilogue (thread == row == TMEM lane) -> label -> sort -> sums
// The accumulator is seeded with |c_j|^2 by one extra k-step (ones x three exact TF32 pieces of |c_j|^2)
// and B holds -2*c, so TMEM already contains s_j = |c_j|^2 - 2 x.c_j and the epilogue is min + sign-mask.
// |x|^2 (needed only for the bound E) is computed per row in the first pass over a matrix and cached as a
// per-tile maximum in the handle (like sklearn's x_squared_norms), later passes read one float per tile.
// Replaces _assign_to_cluster + KMeans._update_centroids for one shard
// (heat/cluster/_kcluster.py:352-370, heat/cluster/kmeans.py:76-103).
#include <math.h>

#include "hk_tma.cuh"

namespace hk {
namespace {

constexpr int TM = 128;    // rows per tile (UMMA M)
constexpr int GT = 128;    // threads per consumer group
constexpr int MISC = 128;  // warps 0-3
constexpr int GW = 4;      // warps per consumer group

struct TcParams {
    int64_t n;
    int d;
    int k;
    int nk;  // k rounded up to a multiple of 32 (UMMA N, TMEM columns per accumulator)
    const float* C;
    void* labels;
    int label_kind;
    double* part;     // [grid*G][k*(d+1)] or nullptr (assign only)
    double* fv_part;  // [grid*G] or nullptr
    int S;            // smem stages
    int nsub;
    int64_t num_tiles;
    const int32_t* state;
    uint32_t tmem_cols;
    float* bounds;   // [num_tiles] per-tile max |x|^2, followed by one int "filled" flag
    int want_write;  // 1: this launch fills `bounds`
    // shared-memory layout (byte offsets from the 1024-aligned base), computed on the host
    uint32_t o_stages, o_B, o_Aext, o_Bext, o_cn, o_grp, grp_stride, o_bars, o_misc;
    uint32_t g_sums, g_cnts, g_wcnt, g_seg, g_perm, g_gxn;  // relative to a group's base
};

struct TcLayout {
    size_t stages, B, Aext, Bext, cn, grp, grp_stride, sums, cnts, wcnt, wpre, tcnt, seg, perm, gxn, bars, misc,
        total;
};

__host__ __device__ inline size_t up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__host__ __device__ inline TcLayout tc_layout(int d, int k, int nk, int S, int G, int nsub, bool sums) {
    TcLayout L;
    size_t o = 0;
    L.stages = o;
    o += (size_t)S * TM * d * 4;
    L.B = o;
    o += up((size_t)nk * d * 4, 1024);
    L.Aext = o;
    o += (size_t)TM * 128;
    L.Bext = o;
    o += up((size_t)nk * 128, 1024);
    L.cn = o;
    o += up((size_t)nk * 4, 16);
    // per-group region (offsets relative to the group base)
    size_t g = 0;
    L.sums = g;
    if (sums) g += up((size_t)nsub * k * d * 8, 16);
    L.cnts = g;
    if (sums) g += up((size_t)k * 8, 16);
    L.wcnt = g;
    if (sums) g += up((size_t)GW * (k + 1) * 4, 16);
    L.wpre = g;
    if (sums) g += up((size_t)GW * (k + 1) * 4, 16);
    L.tcnt = g;
    if (sums) g += up((size_t)(k + 1) * 4, 16);
    L.seg = g;
    if (sums) g += up((size_t)(k + 2) * 4, 16);
    L.perm = g;
    g += up((size_t)GT * 2, 16);
    L.gxn = g;
    g += 2 * GW * 4;
    L.grp_stride = up(g, 16);
    L.grp = o;
    o += L.grp_stride * G;
    L.bars = o;
    o += 8 * 64;  // up to 64 mbarriers
    L.misc = o;
    o += 256;
    L.total = o + 1024;  // slack for manual 1024-byte alignment of the base
    return L;
}

__device__ __forceinline__ void store_label_tc(void* labels, int kind, int64_t row, int lab) {
    if (kind == HK_LABEL_I64)
        reinterpret_cast<long long*>(labels)[row] = lab;
    else if (kind == HK_LABEL_I32)
        reinterpret_cast<int*>(labels)[row] = lab;
    else if (kind == HK_LABEL_U8)
        reinterpret_cast<unsigned char*>(labels)[row] = (unsigned char)lab;
}

// exact fp32 squared distance of the reference formula for (row, centroid j): operands come from the
// swizzled tiles in shared memory (Bt holds -2*c, so the dot already carries the factor), features are
// accumulated in ascending order.  fl(-2*dot) == -2*fl(dot): scaling by two is exact.
__device__ __forceinline__ float exact_d2(const unsigned char* xt, int row, const unsigned char* Bt, int nk, int j,
                                          int d, float xn, float cnj) {
    float dotm2 = 0.f;
    for (int f = 0; f < d; f += 4) {
        const float4 xv = *reinterpret_cast<const float4*>(xt + sw128_off(TM, row, f));
        const float4 cv = *reinterpret_cast<const float4*>(Bt + sw128_off(nk, j, f));
        dotm2 = fmaf(xv.x, cv.x, dotm2);
        dotm2 = fmaf(xv.y, cv.y, dotm2);
        dotm2 = fmaf(xv.z, cv.z, dotm2);
        dotm2 = fmaf(xv.w, cv.w, dotm2);
    }
    return (xn + cnj) + dotm2;
}

__device__ __forceinline__ float row_norm2(const unsigned char* xt, int row, int d) {
    float xn = 0.f;
    for (int f = 0; f < d; f += 4) {
        const float4 xv = *reinterpret_cast<const float4*>(xt + sw128_off(TM, row, f));
        xn = fmaf(xv.x, xv.x, xn);
        xn = fmaf(xv.y, xv.y, xn);
        xn = fmaf(xv.z, xv.z, xn);
        xn = fmaf(xv.w, xv.w, xn);
    }
    return xn;
}

// wait on an mbarrier with one lane per warp, then release the warp (keeps 31 lanes out of the spin loop)
__device__ __forceinline__ void warp_mbar_wait(uint64_t* bar, uint32_t parity, int lane) {
    if (lane == 0) mbar_wait(bar, parity);
    __syncwarp();
}

// sign-bit mask of (s_j < thr) for 32 accumulator columns: bit j <-> column j (4 independent shift chains)
__device__ __forceinline__ unsigned below_mask32(const uint32_t* a, float thr) {
    unsigned m0 = 0, m1 = 0, m2 = 0, m3 = 0;
#pragma unroll
    for (int j = 7; j >= 0; --j) {
        m0 = __funnelshift_l(__float_as_uint(__uint_as_float(a[j]) - thr), m0, 1);
        m1 = __funnelshift_l(__float_as_uint(__uint_as_float(a[8 + j]) - thr), m1, 1);
        m2 = __funnelshift_l(__float_as_uint(__uint_as_float(a[16 + j]) - thr), m2, 1);
        m3 = __funnelshift_l(__float_as_uint(__uint_as_float(a[24 + j]) - thr), m3, 1);
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

enum { XN_COMPUTE = 0, XN_WRITE = 1, XN_READ = 2 };

// G consumer groups; SUMS: accumulate per-cluster sums; CPS: clusters per slice held in register accumulators
// in the sums phase (0 = generic shared-memory read-modify-write path)
template <int G, bool SUMS, int CPS>
__global__ void __launch_bounds__(MISC + G * GT, 1)
    lloyd_tc_kernel(const __grid_constant__ CUtensorMap xmap, const TcParams p) {
    extern __shared__ unsigned char smem_raw[];
    if (p.state != nullptr && p.state[0] != 0) return;  // uniform across the grid
    unsigned char* smem =
        reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int d = p.d, k = p.k, nk = p.nk, S = p.S;
    const int nkb = d >> 5;
    unsigned char* stages = smem + p.o_stages;
    unsigned char* Bt = smem + p.o_B;
    unsigned char* Aext = smem + p.o_Aext;
    unsigned char* Bext = smem + p.o_Bext;
    float* cn = reinterpret_cast<float*>(smem + p.o_cn);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.o_bars);
    uint64_t* full = bars;         // [S]  TMA -> MMA, consumers
    uint64_t* empty = bars + 16;   // [S]  consumers (one arrival per warp) -> TMA
    uint64_t* tfull = bars + 32;   // [G]  MMA -> consumers
    uint64_t* tempty = bars + 40;  // [G]  consumers (one arrival per warp) -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + p.o_misc);
    float* cmax_s = reinterpret_cast<float*>(smem + p.o_misc + 16);
    int* force_exact_s = reinterpret_cast<int*>(smem + p.o_misc + 32);
    double* fvred = reinterpret_cast<double*>(smem + p.o_misc + 64);  // [G*GW] <= 16 doubles

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const uint32_t stage_bytes = (uint32_t)TM * d * 4;
    const int ntiles = (int)p.num_tiles;
    const int xn_mode =
        p.want_write ? XN_WRITE : (reinterpret_cast<const int*>(p.bounds)[p.num_tiles] != 0 ? XN_READ : XN_COMPUTE);

    // ---------------- one-time setup -------------------------------------------------------------------
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], GW);
        }
        for (int g = 0; g < G; ++g) {
            mbar_init(&tfull[g], 1);
            mbar_init(&tempty[g], GW);
        }
        mbar_fence_init();
        tma_prefetch_desc(&xmap);
        *cmax_s = 0.f;
        *force_exact_s = 0;
    }
    if (warp == 2) tmem_alloc(tmem_slot, p.tmem_cols);
    // operand B = -2*C: K-blocked, 128B-swizzled, rows >= k zero
    for (int e = tid; e < nk * (d >> 2); e += blockDim.x) {
        const int j = e / (d >> 2), f = (e - j * (d >> 2)) << 2;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < k) {
            v = *reinterpret_cast<const float4*>(p.C + (size_t)j * d + f);
            v.x *= -2.f;
            v.y *= -2.f;
            v.z *= -2.f;
            v.w *= -2.f;
        }
        *reinterpret_cast<float4*>(Bt + sw128_off(nk, j, f)) = v;
    }
    // seed operands: A_ext[r] = (1,1,1,0,...), B_ext[j] = three exact TF32 pieces of |c_j|^2
    for (int e = tid; e < TM * 8; e += blockDim.x) {
        const int r = e >> 3, ch = e & 7;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ch == 0) v = make_float4(1.f, 1.f, 1.f, 0.f);
        *reinterpret_cast<float4*>(Aext + sw128_off(TM, r, ch << 2)) = v;
    }
    if (SUMS) {
        for (int g = 0; g < G; ++g) {
            unsigned char* gb = smem + p.o_grp + g * p.grp_stride;
            double* sums = reinterpret_cast<double*>(gb + p.g_sums);
            for (int i = tid; i < p.nsub * k * d; i += blockDim.x) sums[i] = 0.0;
            unsigned long long* cnts = reinterpret_cast<unsigned long long*>(gb + p.g_cnts);
            for (int i = tid; i < k; i += blockDim.x) cnts[i] = 0ull;
            int* wcnt = reinterpret_cast<int*>(gb + p.g_wcnt);
            for (int i = tid; i < GW * (k + 1); i += blockDim.x) wcnt[i] = 0;
        }
    }
    __syncthreads();
    for (int j = tid; j < nk; j += blockDim.x) {
        float s = 3.0e38f;  // padded centroids can never win
        if (j < k) {
            s = 0.f;
            for (int f = 0; f < d; ++f) {
                const float c = p.C[(size_t)j * d + f];
                s = fmaf(c, c, s);
            }
            if (!(s < INFINITY)) atomicExch(force_exact_s, 1);  // NaN/Inf centroid: exact path decides
            atomicMax(reinterpret_cast<int*>(cmax_s), __float_as_int(sqrtf(s)));  // s >= 0: int order == float order
        }
        cn[j] = s;
        const float p1 = __uint_as_float(__float_as_uint(s) & 0xFFFFE000u);
        const float r1 = s - p1;
        const float p2 = __uint_as_float(__float_as_uint(r1) & 0xFFFFE000u);
        const float p3 = r1 - p2;
        for (int ch = 0; ch < 8; ++ch) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ch == 0) v = make_float4(p1, p2, p3, 0.f);
            *reinterpret_cast<float4*>(Bext + sw128_off(nk, j, ch << 2)) = v;
        }
    }
    fence_proxy_async();  // operands were written with st.shared, tcgen05.mma reads them through the async proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const float cmax = *cmax_s;
    const bool force_exact = *force_exact_s != 0;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;  // phase parity of stage ring pass
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                mbar_wait(&empty[s], ph ^ 1);
                mbar_expect_tx(&full[s], stage_bytes);
                unsigned char* dst = stages + (size_t)s * stage_bytes;
                for (int kb = 0; kb < nkb; ++kb)
                    tma_load_2d(dst + (size_t)kb * TM * 128, &xmap, &full[s], kb * 32, tile * TM, kEvictFirst);
                if (++s == S) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(TM, nk);
            const uint32_t b_base = smem_u32(Bt);
            const uint64_t aext_d = umma_desc_k_sw128(smem_u32(Aext));
            const uint64_t bext_d = umma_desc_k_sw128(smem_u32(Bext));
            int s = 0, g = 0;
            uint32_t ph = 0, gph = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                mbar_wait(&tempty[g], gph ^ 1);
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t a_base = smem_u32(stages + (size_t)s * stage_bytes);
                const uint32_t dcol = tmem_base + (uint32_t)(g * nk);
                umma_tf32(dcol, aext_d, bext_d, idesc, 0u);  // D = |c_j|^2
                for (int kb = 0; kb < nkb; ++kb) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t ad = umma_desc_k_sw128(a_base + kb * TM * 128 + ks * 32);
                        const uint64_t bd = umma_desc_k_sw128(b_base + kb * nk * 128 + ks * 32);
                        umma_tf32(dcol, ad, bd, idesc, 1u);  // D += x . (-2 c_j)
                    }
                }
                umma_commit(&tfull[g]);
                if (++s == S) {
                    s = 0;
                    ph ^= 1;
                }
                if (++g == G) {
                    g = 0;
                    gph ^= 1;
                }
            }
        }
    } else if (warp >= 4) {
        // ================= consumer groups =================
        const int g = (warp - 4) / GW;
        const int gt = tid - MISC - g * GT;  // 0..127 == row in tile == TMEM lane
        const int q = gt >> 5;               // warp within the group == warp % 4 (TMEM lane quarter)
        unsigned char* gb = smem + p.o_grp + g * p.grp_stride;
        double* sums = reinterpret_cast<double*>(gb + p.g_sums);
        unsigned long long* cnts = reinterpret_cast<unsigned long long*>(gb + p.g_cnts);
        int4* wcnt4 = reinterpret_cast<int4*>(gb + p.g_wcnt);  // [k+1] x {warp 0..3}
        int* wcnt = reinterpret_cast<int*>(gb + p.g_wcnt);
        int* seg = reinterpret_cast<int*>(gb + p.g_seg);
        unsigned short* perm = reinterpret_cast<unsigned short*>(gb + p.g_perm);
        float* gxn = reinterpret_cast<float*>(gb + p.g_gxn);  // [2][GW]
        const int bar_id = 1 + g;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * nk);
        double fv_acc = 0.0;
        const float beta2 = 2.f * 1.05f * 0.001953125f;
        const float gam = (float)(d + 3) * 1.1920929e-7f;
        const int kp1 = k + 1;

        // sums phase geometry: thread == (feature quad fq, slice sl)
        const int FQ = d >> 2;
        const int SL = GT / (FQ < GT ? FQ : GT);
        const int fq = gt % FQ;  // FQ <= 64 for d <= 256
        const int sl = gt / FQ;
        const uint32_t kboff = (uint32_t)((fq >> 3) * TM * 128);
        const uint32_t cx = (uint32_t)((fq & 7) << 4);
        constexpr int NACC = CPS > 0 ? CPS : 1;
        float4 acc[NACC];
#pragma unroll
        for (int cc = 0; cc < NACC; ++cc) acc[cc] = make_float4(0.f, 0.f, 0.f, 0.f);
        int rows_since = 0;
        int c_base, sub, stride;
        if (p.nsub > 1) {
            c_base = sl / p.nsub;
            sub = sl - c_base * p.nsub;
            stride = p.nsub;
        } else {
            c_base = sl * NACC;
            sub = 0;
            stride = 1;
        }

        auto flush_acc = [&]() {
#pragma unroll
            for (int cc = 0; cc < NACC; ++cc) {
                const int c = c_base + cc;
                if (c < k) {
                    double* sp = sums + ((size_t)sub * k + c) * d + (fq << 2);
                    double2 lo = *reinterpret_cast<double2*>(sp);
                    double2 hi = *reinterpret_cast<double2*>(sp + 2);
                    lo.x += (double)acc[cc].x;
                    lo.y += (double)acc[cc].y;
                    hi.x += (double)acc[cc].z;
                    hi.y += (double)acc[cc].w;
                    *reinterpret_cast<double2*>(sp) = lo;
                    *reinterpret_cast<double2*>(sp + 2) = hi;
                }
                acc[cc] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            rows_since = 0;
        };

        // ring bookkeeping without divisions: stage s / its phase, accumulator phase, tile parity
        int s = g % S;
        uint32_t ph = (uint32_t)((g / S) & 1);
        uint32_t gph = 0;
        for (int tile = blockIdx.x + g * gridDim.x; tile < ntiles; tile += G * gridDim.x) {
            const unsigned char* xt = stages + (uint32_t)s * stage_bytes;
            const int row0 = tile * TM;
            const int rows = (p.n - row0) < (int64_t)TM ? (int)(p.n - row0) : TM;
            const bool active = gt < rows;

            warp_mbar_wait(&full[s], ph, lane);  // x tile landed (needed by |x|^2, exact path, sums)
            float xn;  // |x|^2 of this row, or an upper bound for every row of the tile
            if (xn_mode == XN_READ) {
                xn = __ldg(p.bounds + tile);
            } else {
                xn = row_norm2(xt, gt, d);
                if (xn_mode == XN_WRITE) {
                    float wm = xn;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, o));
                    if (!(wm == wm)) wm = INFINITY;  // NaN rows: make the cached bound useless, not wrong
                    if (lane == 0) gxn[gph * GW + q] = wm;
                }
            }
            const float E2 = 2.f * ((beta2 * sqrtf(xn) * cmax + gam * (xn + cmax * cmax)) * 1.001f);

            warp_mbar_wait(&tfull[g], gph, lane);  // accumulator ready: s_j = |c_j|^2 - 2 x.c_j (TF32)
            tc_fence_after();
            // two sweeps over the accumulator, 32 columns in registers at a time: min, then sign mask
            float m = INFINITY;
            for (int c0 = 0; c0 < nk; c0 += 32) {
                uint32_t a[32];
                tmem_ld32(taddr + (uint32_t)c0, a);
                tmem_wait_ld();
                m = fminf(m, min32(a));
            }
            const float thr = m + E2;
            int cnt = 0, idx = 0;
            for (int c0 = 0; c0 < nk; c0 += 32) {
                uint32_t a[32];
                tmem_ld32(taddr + (uint32_t)c0, a);
                tmem_wait_ld();
                const unsigned mk = below_mask32(a, thr);
                cnt += __popc(mk);
                if (mk) idx = c0 + __ffs(mk) - 1;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[g]);  // accumulator g may be overwritten

            int lab = k;
            if (active) {
                float best = INFINITY;
                bool have_best = false;
                float xr = xn;  // exact |x|^2 of this row (the cached value is only a bound)
                if (cnt == 1 && !force_exact && xn < INFINITY) {
                    lab = idx;
                } else {
                    // undecided (near-tie within the TF32 bound, NaN/Inf): exact formula, torch.min semantics
                    if (xn_mode == XN_READ) xr = row_norm2(xt, gt, d);
                    int bl = 0;
                    for (int j = 0; j < k; ++j) {
                        float d2 = exact_d2(xt, gt, Bt, nk, j, d, xr, cn[j]);
                        d2 = d2 < 0.f ? 0.f : d2;
                        if (d2 < best || (d2 != d2 && best == best)) {
                            best = d2;
                            bl = j;
                        }
                    }
                    lab = bl;
                    have_best = true;
                }
                if (p.label_kind != HK_LABEL_NONE) store_label_tc(p.labels, p.label_kind, (int64_t)row0 + gt, lab);
                if (p.fv_part != nullptr) {
                    if (!have_best) {
                        if (xn_mode == XN_READ) xr = row_norm2(xt, gt, d);
                        best = exact_d2(xt, gt, Bt, nk, lab, d, xr, cn[lab]);
                        best = best < 0.f ? 0.f : best;
                    }
                    const float sq = sqrtf(best);
                    fv_acc += (double)(sq * sq);
                }
            }

            if (SUMS) {
                // ---- deterministic counting sort of the tile's rows by label: two group barriers ----
                const unsigned peers = __match_any_sync(0xffffffffu, lab);
                const int rank = __popc(peers & lanemask_lt());
                const bool leader = lane == __ffs(peers) - 1;
                if (leader) wcnt[lab * 4 + q] = __popc(peers);
                named_bar_sync(bar_id, GT);  // (1) per-warp label counts visible; previous tile fully consumed
                if (xn_mode == XN_WRITE && gt == 0) {
                    const float* gx = gxn + gph * GW;
                    p.bounds[tile] = fmaxf(fmaxf(gx[0], gx[1]), fmaxf(gx[2], gx[3]));
                }
                {
                    // every warp redundantly scans the (k+1) cluster totals: lane owns `per` consecutive clusters
                    const int per = (kp1 + 31) >> 5;
                    const int b0 = lane * per;
                    int local = 0;
                    for (int ii = 0; ii < per; ++ii) {
                        const int c = b0 + ii;
                        if (c <= k) {
                            const int4 w4 = wcnt4[c];
                            local += (w4.x + w4.y) + (w4.z + w4.w);
                        }
                    }
                    int incl = local;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int vv = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += vv;
                    }
                    int run = incl - local;
                    for (int ii = 0; ii < per; ++ii) {
                        const int c = b0 + ii;
                        if (c <= k) {
                            const int4 w4 = wcnt4[c];
                            const int t = (w4.x + w4.y) + (w4.z + w4.w);
                            seg[c] = run;  // all four warps store identical values
                            if (q == 0 && c < k) cnts[c] += (unsigned long long)t;
                            run += t;
                        }
                    }
                    if (lane == 31) seg[kp1] = incl;
                }
                __syncwarp();
                {
                    const int4 w4 = wcnt4[lab];
                    int pos = seg[lab] + rank;
                    if (q > 0) pos += w4.x;
                    if (q > 1) pos += w4.y;
                    if (q > 2) pos += w4.z;
                    // perm holds the swizzled byte offset of the row inside a K-block: (r << 7) | ((r & 7) << 4)
                    perm[pos] = (unsigned short)((gt << 7) | ((gt & 7) << 4));
                }
                named_bar_sync(bar_id, GT);  // (2) perm/seg complete; wcnt reads done
                if (leader) wcnt[lab * 4 + q] = 0;  // clean table for the next tile
                // ---- segmented column sums ----
                const unsigned char* xk = xt + kboff;
                if constexpr (CPS > 0) {
                    // register accumulators: slice sl always owns clusters [c_base, c_base + CPS)
                    if (sl < SL) {
                        int e = seg[c_base < kp1 ? c_base : kp1];
#pragma unroll
                        for (int cc = 0; cc < NACC; ++cc) {
                            const int b = e;
                            const int cn1 = c_base + cc + 1;
                            e = seg[cn1 < kp1 ? cn1 : kp1];
                            if (c_base + cc < k) {
                                for (int ii = b + sub; ii < e; ii += stride) {
                                    const float4 x4 = *reinterpret_cast<const float4*>(xk + ((uint32_t)perm[ii] ^ cx));
                                    acc[cc].x += x4.x;
                                    acc[cc].y += x4.y;
                                    acc[cc].z += x4.z;
                                    acc[cc].w += x4.w;
                                }
                                rows_since += e - b;
                            }
                        }
                        if (rows_since >= 32) flush_acc();  // keep fp32 partial sums short, then widen
                    }
                } else {
                    if (sl < SL) {
                        const int nslots = p.nsub * k;
                        for (int vs = sl; vs < nslots; vs += SL) {
                            const int c = vs / p.nsub;
                            const int sb = vs - c * p.nsub;
                            const int b = seg[c], e = seg[c + 1];
                            float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
                            for (int ii = b + sb; ii < e; ii += p.nsub) {
                                const float4 x4 = *reinterpret_cast<const float4*>(xk + ((uint32_t)perm[ii] ^ cx));
                                a4.x += x4.x;
                                a4.y += x4.y;
                                a4.z += x4.z;
                                a4.w += x4.w;
                            }
                            if (b + sb < e) {
                                double* sp = sums + ((size_t)sb * k + c) * d + (fq << 2);
                                sp[0] += (double)a4.x;
                                sp[1] += (double)a4.y;
                                sp[2] += (double)a4.z;
                                sp[3] += (double)a4.w;
                            }
                        }
                    }
                }
            } else if (xn_mode == XN_WRITE) {
                named_bar_sync(bar_id, GT);
                if (gt == 0) {
                    const float* gx = gxn + gph * GW;
                    p.bounds[tile] = fmaxf(fmaxf(gx[0], gx[1]), fmaxf(gx[2], gx[3]));
                }
            }
            // this warp is done with stage s (the next tile's barrier (1) separates perm/seg reuse)
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            s += G;
            if (s >= S) {
                s -= S;
                ph ^= 1;
            }
            gph ^= 1;
        }

        if (SUMS) {
            if constexpr (CPS > 0) {
                if (sl < SL) flush_acc();
            }
            named_bar_sync(bar_id, GT);
            double* out = p.part + ((size_t)blockIdx.x * G + g) * k * (d + 1);
            for (int ii = gt; ii < k * d; ii += GT) {
                const int c = ii / d, f = ii - c * d;
                double t = 0.0;
                for (int sb = 0; sb < p.nsub; ++sb) t += sums[((size_t)sb * k + c) * d + f];
                out[(size_t)c * (d + 1) + f] = t;
            }
            for (int c = gt; c < k; c += GT) out[(size_t)c * (d + 1) + d] = (double)cnts[c];
        }
        if (p.fv_part != nullptr) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) fv_acc += __shfl_xor_sync(0xffffffffu, fv_acc, o);
            if (lane == 0) fvred[g * GW + q] = fv_acc;
            named_bar_sync(bar_id, GT);
            if (gt == 0) {
                double t = 0.0;
                for (int w = 0; w < GW; ++w) t += fvred[g * GW + w];
                p.fv_part[(size_t)blockIdx.x * G + g] = t;
            }
        }
    }

    // ---------------- teardown ---------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
    if (xn_mode == XN_WRITE && blockIdx.x == 0 && tid == 0) reinterpret_cast<int*>(p.bounds)[p.num_tiles] = 1;
}

__global__ void reduce_partials_tc_kernel(const double* __restrict__ part, int nb, int len,
                                          double* __restrict__ out, const int32_t* state) {
    if (state != nullptr && state[0] != 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    double t = 0.0;
    for (int b = 0; b < nb; ++b) t += part[(size_t)b * len + i];
    out[i] = t;
}

struct TcPlan {
    int G, S, nk, nsub, cps;
    uint32_t tmem_cols;
    size_t smem;
    bool ok;
};

TcPlan plan_tc(const Handle* h, int d, int k, bool sums) {
    TcPlan pl{};
    pl.ok = false;
    pl.nk = (k + 31) / 32 * 32;
    const int FQ = d / 4;
    const int SL = GT / (FQ < GT ? FQ : GT);
    pl.nsub = SL / k;
    if (pl.nsub < 1) pl.nsub = 1;
    // clusters per slice for the register-accumulator sums phase (1, 2, 4, 8; 0 = generic)
    int cps = pl.nsub > 1 ? 1 : (k + SL - 1) / SL;
    pl.cps = cps <= 1 ? 1 : (cps <= 2 ? 2 : (cps <= 4 ? 4 : (cps <= 8 ? 8 : 0)));
    if (!sums) pl.cps = 1;
    const size_t budget = (size_t)h->smem_optin;
    for (int G = 4; G >= 2; --G) {
        if (G * pl.nk > 512) continue;
        for (int S = 8; S >= 3; --S) {
            TcLayout L = tc_layout(d, k, pl.nk, S, G, pl.nsub, sums);
            if (L.total <= budget && (S >= G + 1 || S >= 4)) {
                pl.G = G;
                pl.S = S;
                pl.smem = L.total;
                uint32_t cols = 32;
                while (cols < (uint32_t)(G * pl.nk)) cols <<= 1;
                pl.tmem_cols = cols;
                pl.ok = true;
                return pl;
            }
        }
    }
    return pl;
}

template <int G, bool SUMS, int CPS>
int launch_inst(Handle* h, const CUtensorMap& map, TcParams& p, size_t smem, int grid, cudaStream_t st) {
    auto kern = lloyd_tc_kernel<G, SUMS, CPS>;
    HK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    prof_begin(h, st);
    kern<<<grid, MISC + G * GT, smem, st>>>(map, p);
    prof_end(h, st);
    HK_CUDA(cudaGetLastError());
    h->launches++;
    return 0;
}

template <int G>
int launch_g(Handle* h, const CUtensorMap& map, TcParams& p, const TcPlan& pl, bool sums, int grid,
             cudaStream_t st) {
    if (!sums) return launch_inst<G, false, 1>(h, map, p, pl.smem, grid, st);
    switch (pl.cps) {
        case 1: return launch_inst<G, true, 1>(h, map, p, pl.smem, grid, st);
        case 2: return launch_inst<G, true, 2>(h, map, p, pl.smem, grid, st);
        case 4: return launch_inst<G, true, 4>(h, map, p, pl.smem, grid, st);
        case 8: return launch_inst<G, true, 8>(h, map, p, pl.smem, grid, st);
        default: return launch_inst<G, true, 0>(h, map, p, pl.smem, grid, st);
    }
}

}  // namespace

bool tc_supported(const Handle* h, const LloydArgs& a) {
    if (a.dtype != HK_F32) return false;
    if (a.d % 32 != 0 || a.d < 32 || a.d > 256) return false;
    if (a.k < 1 || a.k > 256) return false;
    if (a.ldx % 4 != 0) return false;
    if ((reinterpret_cast<uintptr_t>(a.X) & 15) != 0 || (reinterpret_cast<uintptr_t>(a.C) & 15) != 0) return false;
    if (a.n >= (int64_t)1 << 31) return false;  // TMA coordinates are int32
    TcPlan pl = plan_tc(h, a.d, a.k, a.partials != nullptr);
    return pl.ok;
}

int launch_lloyd_tc(Handle* h, const LloydArgs& a) {
    const bool sums = a.partials != nullptr;
    TcPlan pl = plan_tc(h, a.d, a.k, sums);
    if (!pl.ok) {
        set_error("lloyd_tc: no feasible plan for d=%d k=%d", a.d, a.k);
        return -2;
    }
    CUtensorMap map;
    int rc = make_tensor_map_2d(&map, a.X, 4, (uint64_t)a.n, (uint64_t)a.d, (uint64_t)a.ldx, 32, TM, 128);
    if (rc) return rc;
    const int len = a.k * (a.d + 1);
    TcParams p{};
    p.n = a.n;
    p.d = a.d;
    p.k = a.k;
    p.nk = pl.nk;
    p.C = reinterpret_cast<const float*>(a.C);
    p.labels = a.labels;
    p.label_kind = a.labels ? a.label_kind : HK_LABEL_NONE;
    p.S = pl.S;
    p.nsub = pl.nsub;
    p.num_tiles = (a.n + TM - 1) / TM;
    p.state = a.state;
    p.tmem_cols = pl.tmem_cols;
    {
        const TcLayout L = tc_layout(a.d, a.k, pl.nk, pl.S, pl.G, pl.nsub, sums);
        p.o_stages = (uint32_t)L.stages;
        p.o_B = (uint32_t)L.B;
        p.o_Aext = (uint32_t)L.Aext;
        p.o_Bext = (uint32_t)L.Bext;
        p.o_cn = (uint32_t)L.cn;
        p.o_grp = (uint32_t)L.grp;
        p.grp_stride = (uint32_t)L.grp_stride;
        p.o_bars = (uint32_t)L.bars;
        p.o_misc = (uint32_t)L.misc;
        p.g_sums = (uint32_t)L.sums;
        p.g_cnts = (uint32_t)L.cnts;
        p.g_wcnt = (uint32_t)L.wcnt;
        p.g_seg = (uint32_t)L.seg;
        p.g_perm = (uint32_t)L.perm;
        p.g_gxn = (uint32_t)L.gxn;
    }

    // per-tile |x|^2 bound cache, keyed by the matrix identity (reset with hk_cache_reset)
    const size_t need = ((size_t)p.num_tiles + 4) * sizeof(float);
    const bool same = h->xb != nullptr && h->xb_X == a.X && h->xb_n == a.n && h->xb_d == a.d && h->xb_ld == a.ldx;
    if (!same) {
        if (h->xb_bytes < need) {
            if (h->xb) HK_CUDA(cudaFree(h->xb));
            h->xb = nullptr;
            h->xb_bytes = 0;
            HK_CUDA(cudaMalloc(&h->xb, need));
            h->xb_bytes = need;
        }
        HK_CUDA(cudaMemsetAsync(h->xb + p.num_tiles, 0, sizeof(int), a.stream));
        h->xb_X = a.X;
        h->xb_n = a.n;
        h->xb_d = a.d;
        h->xb_ld = a.ldx;
        p.want_write = 1;
    } else {
        p.want_write = 0;
    }
    p.bounds = h->xb;

    int64_t grid64 = h->num_sms;
    if (grid64 > p.num_tiles) grid64 = p.num_tiles;
    const int grid = (int)grid64;
    const int nb = grid * pl.G;
    rc = ensure_part(h, ((size_t)nb * len + nb) * sizeof(double));
    if (rc) return rc;
    p.part = sums ? h->part : nullptr;
    p.fv_part = a.fv_out ? h->part + (size_t)nb * len : nullptr;

    char name[112];
    snprintf(name, sizeof(name), "tc<f32,d=%d,k=%d,G=%d,S=%d,cps=%d,%s,%s>", a.d, a.k, pl.G, pl.S, pl.cps,
             sums ? "sums" : "assign", p.want_write ? "xn-write" : "xn-cached");
    h->variant = name;

    switch (pl.G) {
        case 2: rc = launch_g<2>(h, map, p, pl, sums, grid, a.stream); break;
        case 3: rc = launch_g<3>(h, map, p, pl, sums, grid, a.stream); break;
        case 4: rc = launch_g<4>(h, map, p, pl, sums, grid, a.stream); break;
        default: rc = -2;
    }
    if (rc) return rc;
    if (sums) {
        reduce_partials_tc_kernel<<<(len + 255) / 256, 256, 0, a.stream>>>(h->part, nb, len, a.partials, a.state);
        HK_CUDA(cudaGetLastError());
        h->launches++;
    }
    if (a.fv_out) {
        reduce_partials_tc_kernel<<<1, 32, 0, a.stream>>>(p.fv_part, nb, 1, a.fv_out, nullptr);
        HK_CUDA(cudaGetLastError());
        h->launches++;
    }
    return 0;
}

// ---- tensor map encoding (driver entry point fetched through the runtime) ---------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tensor_map_2d(CUtensorMap* map, const void* base, int elem_bytes, uint64_t rows, uint64_t cols,
                       uint64_t ld, uint32_t box_cols, uint32_t box_rows, int swizzle_bytes) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        HK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres));
        if (!sym || qres != cudaDriverEntryPointSuccess) {
            set_error("cuTensorMapEncodeTiled is not available from this driver");
            return -5;
        }
        fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    const cuuint64_t gdim[2] = {cols, rows};
    const cuuint64_t gstride[1] = {ld * (uint64_t)elem_bytes};
    const cuuint32_t box[2] = {box_cols, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt = elem_bytes == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    const CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                  : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                        : CU_TENSOR_MAP_SWIZZLE_NONE;
    const CUresult r = fn(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu ld=%llu box=%ux%u)", (int)r,
                  (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_cols, box_rows);
        return 3000 + (int)r;
    }
    return 0;
}

}  // namespace hk
