// Exact-FMA Lloyd pass for matrices whose rows are exactly 128 bytes (fp32 d=32, fp64 d=16): the HBM-bound
// small-k regime (BASELINE config 5: N=50M, d=16, k=8, fp64).  Same warp-specialised, mbarrier-only pipeline as
// the tensor-core kernel (hk_lloyd_tc.cu) minus the MMA:
//   warp 0      : TMA producer (cp.async.bulk.tensor.2d, 128-byte swizzle, EVICT_FIRST) into an S-stage ring
//   warps 4-19  : distance warps.  Warp (q, r) owns rows [32q, 32q+32) of the tiles i == r (mod 4); thread == row:
//                 the row is read from the swizzled tile (conflict free) into registers, the centroids are read
//                 with warp-uniform (broadcast) 16-byte loads, d2 = fl(fl(|x|^2 + |c|^2) - 2 x.c) exactly as
//                 heat/spatial/distance.py:59-64, first-index argmin with torch.min NaN semantics
//                 (heat/core/statistics.py:177); labels go to shared memory, cluster counts to private arrays
//   warps 20-27 : accumulator warps with PRIVATE [k+1][128 B] accumulators in shared memory (fp32 data: fp32,
//                 widened to fp64 slots in global memory before any row can have taken ~100 adds; fp64 data: fp64,
//                 written once at the end); lanes cover 4 rows x 8 sixteen-byte chunks per step, plain load-add-store
// Replaces _assign_to_cluster + KMeans._update_centroids for one shard
// (heat/cluster/_kcluster.py:352-370, heat/cluster/kmeans.py:76-103).
#include <math.h>

#include "hk_tma.cuh"

namespace hk {
namespace {

constexpr int TM = 128;
constexpr int MISC_WARPS = 4;
#ifndef HK_ROW_ER
#define HK_ROW_ER 4
#endif
constexpr int ER = HK_ROW_ER;           // tile residues handled by the distance warps (4 warps each)
constexpr int E_WARPS = 4 * ER;
constexpr int A_WARPS = 8;
constexpr int E_FIRST = MISC_WARPS;
constexpr int A_FIRST = MISC_WARPS + E_WARPS;
constexpr int STAGE_BYTES = TM * 128;

struct RowParams {
    int64_t n;
    int k;
    const void* C;
    void* labels;
    int label_kind;
    double* fsum;     // [grid*A_WARPS][k*d]
    double* fcnt;     // [grid][k]
    double* fv_part;  // [grid]
    int S;
    int num_tiles;
    const int32_t* state;
    uint32_t o_stages, o_C, o_cn, o_acc, o_lab, o_cnt, o_mmax, o_bars, o_misc;
};

struct RowLayout {
    size_t stages, C, cn, acc, lab, cnt, mmax, bars, misc, total;
};

inline size_t up(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline RowLayout row_layout(int k, int S, bool sums) {
    RowLayout L;
    size_t o = 0;
    L.stages = o;
    o += (size_t)S * STAGE_BYTES;
    L.C = o;
    o += (size_t)k * 128;
    L.cn = o;
    o += up((size_t)k * 8, 16);
    L.acc = o;
    if (sums) o += (size_t)A_WARPS * (k + 1) * 128;
    L.lab = o;
    if (sums) o += (size_t)S * TM * 2;
    L.cnt = o;
    if (sums) o += up((size_t)E_WARPS * k * 4, 16);
    L.mmax = o;
    if (sums) o += up((size_t)S * 16, 16);
    L.bars = o;
    o += 8 * 64;
    L.misc = o;
    o += 256;
    L.total = o + 1024;
    return L;
}

template <typename T>
struct V16;
template <>
struct V16<float> {
    using type = float4;
    static constexpr int N = 4;
    static __device__ __forceinline__ float4 lds(uint32_t a) { return lds_f4(a); }
    static __device__ __forceinline__ void sts(uint32_t a, float4 v) { sts_f4_nc(a, v); }
    static __device__ __forceinline__ float4 zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
    static __device__ __forceinline__ void add(float4& a, const float4& b) {
        a.x += b.x;
        a.y += b.y;
        a.z += b.z;
        a.w += b.w;
    }
};
template <>
struct V16<double> {
    using type = double2;
    static constexpr int N = 2;
    static __device__ __forceinline__ double2 lds(uint32_t a) { return lds_d2(a); }
    static __device__ __forceinline__ void sts(uint32_t a, double2 v) { sts_d2_nc(a, v); }
    static __device__ __forceinline__ double2 zero() { return make_double2(0.0, 0.0); }
    static __device__ __forceinline__ void add(double2& a, const double2& b) {
        a.x += b.x;
        a.y += b.y;
    }
};

__device__ __forceinline__ void store_label_r(void* labels, int kind, int64_t row, int lab) {
    if (kind == HK_LABEL_I64)
        reinterpret_cast<long long*>(labels)[row] = lab;
    else if (kind == HK_LABEL_I32)
        reinterpret_cast<int*>(labels)[row] = lab;
    else if (kind == HK_LABEL_U8)
        reinterpret_cast<unsigned char*>(labels)[row] = (unsigned char)lab;
}

template <typename T, bool SUMS>
__global__ void __launch_bounds__((MISC_WARPS + E_WARPS + (SUMS ? A_WARPS : 0)) * 32, 1)
    lloyd_row128_kernel(const __grid_constant__ CUtensorMap xmap, const RowParams p) {
    extern __shared__ unsigned char smem_raw[];
    if (p.state != nullptr && p.state[0] != 0) return;  // uniform across the grid
    using V = typename V16<T>::type;
    constexpr int EPV = V16<T>::N;       // elements per 16-byte chunk
    constexpr int D = 128 / sizeof(T);   // features
    unsigned char* smem =
        reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const int k = p.k, S = p.S;
    const uint32_t a_stages = sbase + p.o_stages;
    const uint32_t a_C = sbase + p.o_C;
    const uint32_t a_lab = sbase + p.o_lab;
    T* cn = reinterpret_cast<T*>(smem + p.o_cn);
    const uint32_t b_full = sbase + p.o_bars;  // full[S] | empty[S] | lfull[S] | mfull[S]  (16 slots each)
    const uint32_t b_empty = b_full + 16 * 8;
    const uint32_t b_lfull = b_full + 32 * 8;
    const uint32_t b_mfull = b_full + 48 * 8;
    double* fvred = reinterpret_cast<double*>(smem + p.o_misc);  // [E_WARPS]

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int ntiles = p.num_tiles;

    if (tid == 0) {
        uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.o_bars);
        for (int s = 0; s < S; ++s) {
            mbar_init(bars + s, 1);
            mbar_init(bars + 16 + s, 4 + (SUMS ? 4 : 0));
            mbar_init(bars + 32 + s, 4);
            mbar_init(bars + 48 + s, 4);
        }
        mbar_fence_init();
        tma_prefetch_desc(&xmap);
    }
    {
        const T* Cg = reinterpret_cast<const T*>(p.C);
        T* Cs = reinterpret_cast<T*>(smem + p.o_C);
        for (int i = tid; i < k * D; i += blockDim.x) Cs[i] = Cg[i];
    }
    if (SUMS) {
        float* az = reinterpret_cast<float*>(smem + p.o_acc);
        for (int i = tid; i < A_WARPS * (k + 1) * 32; i += blockDim.x) az[i] = 0.f;
        int* cz = reinterpret_cast<int*>(smem + p.o_cnt);
        for (int i = tid; i < E_WARPS * k; i += blockDim.x) cz[i] = 0;
    }
    __syncthreads();
    for (int j = tid; j < k; j += blockDim.x) {
        const T* cr = reinterpret_cast<const T*>(smem + p.o_C) + j * D;
        T s = T(0);
        for (int f = 0; f < D; ++f) s = fma(cr[f], cr[f], s);
        cn[j] = s;
    }
    __syncthreads();

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                mbar_wait_a(b_empty + s * 8, ph ^ 1);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b_full + s * 8),
                             "r"((uint32_t)STAGE_BYTES)
                             : "memory");
                asm volatile(
                    "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
                    " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(a_stages + s * STAGE_BYTES),
                    "l"(&xmap), "r"(b_full + s * 8), "r"(0), "r"(tile * TM), "l"(kEvictFirst)
                    : "memory");
                if (++s == S) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
        __syncwarp();  // lanes 1-31 must not run ahead to the block-wide barriers below while lane 0 still produces
    } else if (warp >= E_FIRST && warp < E_FIRST + E_WARPS) {
        // ================= distance warps =================
        const int we = warp - E_FIRST;
        const int q = we & 3;
        const int r = we >> 2;
        const int row = q * 32 + lane;
        const uint32_t a_ecnt = sbase + p.o_cnt + (uint32_t)we * (uint32_t)(k * 4);
        const uint32_t a_mmax = sbase + p.o_mmax;
        double fv_acc = 0.0;
        int s = r % S;
        uint32_t ph = (uint32_t)((r / S) & 1);
        for (int tile = blockIdx.x + r * gridDim.x; tile < ntiles; tile += ER * gridDim.x) {
            const uint32_t xrow = a_stages + s * STAGE_BYTES + (uint32_t)row * 128;
            const int row0 = tile * TM;
            const bool active = (int64_t)row0 + row < p.n;
            warp_wait(b_full + s * 8, ph, lane);
            // the row, chunk by chunk in feature order (physical chunk = logical ^ (row & 7): conflict free)
            T x[D];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const V v = V16<T>::lds(xrow + (uint32_t)((c ^ (row & 7)) << 4));
                const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
                for (int t = 0; t < EPV; ++t) x[c * EPV + t] = e[t];
            }
            T xn = T(0);
#pragma unroll
            for (int f = 0; f < D; ++f) xn = fma(x[f], x[f], xn);
            int lab = 0;
            T best = T(INFINITY);
#pragma unroll 2
            for (int j = 0; j < k; ++j) {
                const uint32_t cj = a_C + (uint32_t)j * 128;
                T dot = T(0);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const V v = V16<T>::lds(cj + c * 16);  // warp-uniform address: broadcast
                    const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
                    for (int t = 0; t < EPV; ++t) dot = fma(x[c * EPV + t], e[t], dot);
                }
                T d2 = (xn + cn[j]) - T(2) * dot;
                d2 = d2 < T(0) ? T(0) : d2;  // clamp(d2, 0, inf); NaN stays NaN
                if (d2 < best || (d2 != d2 && best == best)) {
                    best = d2;
                    lab = j;
                }
            }
            if (active) {
                if (p.label_kind != HK_LABEL_NONE) store_label_r(p.labels, p.label_kind, (int64_t)row0 + row, lab);
                if (p.fv_part != nullptr) {
                    const T sq = sqrt(best);
                    fv_acc += (double)(sq * sq);
                }
            } else {
                lab = k;
            }
            if (SUMS) {
                sts_u16(a_lab + s * (TM * 2) + row * 2, (uint32_t)lab);
                __syncwarp();
                if (lane == 0) mbar_arrive_a(b_lfull + s * 8);
                const unsigned peers = __match_any_sync(0xffffffffu, lab);
                const int mult = __popc(peers);
                if (lane == __ffs(peers) - 1 && lab < k) {
                    const uint32_t ca = a_ecnt + (uint32_t)lab * 4;
                    sts_s32(ca, lds_s32(ca) + mult);
                }
                const int mm = __reduce_max_sync(0xffffffffu, mult);
                if (lane == 0) {
                    sts_s32(a_mmax + (uint32_t)(s * 4 + q) * 4, mm);
                    mbar_arrive_a(b_mfull + s * 8);
                    mbar_arrive_a(b_empty + s * 8);
                }
            } else {
                __syncwarp();
                if (lane == 0) mbar_arrive_a(b_empty + s * 8);
            }
            s += ER;
            while (s >= S) {
                s -= S;
                ph ^= 1;
            }
        }
        if (p.fv_part != nullptr) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) fv_acc += __shfl_xor_sync(0xffffffffu, fv_acc, o);
            if (lane == 0) fvred[we] = fv_acc;
        }
    } else if (SUMS && warp >= A_FIRST) {
        // ================= accumulator warps =================
        const int a = warp - A_FIRST;
        const int q = a & 3;
        const int res = a >> 2;
        constexpr int NRES = A_WARPS / 4;
        const int g = lane >> 3;  // row slot inside a step (4 rows per step)
        const int ch = lane & 7;  // 16-byte chunk of the row
        const uint32_t acc_w = sbase + p.o_acc + (uint32_t)a * (uint32_t)((k + 1) * 128);
        const uint32_t acc_l = acc_w + ch * 16;
        const uint32_t a_mmax = sbase + p.o_mmax;
        double* gslot = p.fsum + ((size_t)blockIdx.x * A_WARPS + a) * (size_t)(k * D);
        bool first_flush = true;
        int run_max = 0;

        auto flush = [&]() {
            // private sums -> this warp's fp64 slot (sole owner: plain read-modify-write), then clear
            const int nq = k * 8;
            for (int e = lane; e < nq; e += 32) {
                const V v = V16<T>::lds(acc_w + e * 16);
                V16<T>::sts(acc_w + e * 16, V16<T>::zero());
                const T* ve = reinterpret_cast<const T*>(&v);
                double* gp = gslot + (size_t)e * EPV;
#pragma unroll
                for (int t = 0; t < EPV; ++t) gp[t] = (first_flush ? 0.0 : gp[t]) + (double)ve[t];
            }
            first_flush = false;
            run_max = 0;
        };

        int s = res % S;
        uint32_t ph = (uint32_t)((res / S) & 1);
        for (int tile = blockIdx.x + res * gridDim.x; tile < ntiles; tile += NRES * gridDim.x) {
            warp_wait(b_full + s * 8, ph, lane);
            warp_wait(b_lfull + s * 8, ph, lane);
            const uint32_t xq = a_stages + s * STAGE_BYTES + (uint32_t)(q * 32 * 128);
            const uint32_t mylab = lds_u16(a_lab + s * (TM * 2) + (q * 32 + lane) * 2);
            bool c = false;
#pragma unroll
            for (int x = 1; x < 4; ++x) c |= (__shfl_xor_sync(0xffffffffu, mylab, x) == mylab);
            const unsigned coll = __ballot_sync(0xffffffffu, c);
            V xr[8];
            uint32_t aa[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int rl = j * 4 + g;
                const uint32_t l = __shfl_sync(0xffffffffu, mylab, rl);
                aa[j] = acc_l + l * 128;
                xr[j] = V16<T>::lds(xq + ((uint32_t)(rl << 7) | ((uint32_t)((rl ^ ch) & 7) << 4)));
            }
            if (coll == 0u) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    V v = V16<T>::lds(aa[j]);
                    V16<T>::add(v, xr[j]);
                    V16<T>::sts(aa[j], v);
                }
            } else {
#pragma unroll 1
                for (int j = 0; j < 8; ++j) {
                    V xj = xr[0];
                    uint32_t aj = aa[0];
#pragma unroll
                    for (int t = 1; t < 8; ++t)
                        if (j == t) {
                            xj = xr[t];
                            aj = aa[t];
                        }
#pragma unroll 1
                    for (int gg = 0; gg < 4; ++gg) {
                        if (g == gg) {
                            V v = V16<T>::lds(aj);
                            V16<T>::add(v, xj);
                            V16<T>::sts(aj, v);
                        }
                        __syncwarp();
                    }
                }
            }
            warp_wait(b_mfull + s * 8, ph, lane);
            run_max += lds_s32(a_mmax + (uint32_t)(s * 4 + q) * 4);
            __syncwarp();
            if (lane == 0) mbar_arrive_a(b_empty + s * 8);
            // fp32 partial sums stay short (fp64 data needs no intermediate widening)
            if (sizeof(T) == 4 && run_max >= 72) flush();
            s += NRES;
            if (s >= S) {
                s -= S;
                ph ^= 1;
            }
        }
        flush();
    }

    __syncthreads();
    if (SUMS) {
        const int* ce = reinterpret_cast<const int*>(smem + p.o_cnt);
        for (int c = tid; c < k; c += blockDim.x) {
            int t = 0;
            for (int w = 0; w < E_WARPS; ++w) t += ce[w * k + c];
            p.fcnt[(size_t)blockIdx.x * k + c] = (double)t;
        }
    }
    if (tid == 0 && p.fv_part != nullptr) {
        double t = 0.0;
        for (int w = 0; w < E_WARPS; ++w) t += fvred[w];
        p.fv_part[blockIdx.x] = t;
    }
}

// partials[c][0..d) = sum over accumulator slots, partials[c][d] = sum over CTAs of the counts (fixed order)
__global__ void __launch_bounds__(256) reduce_row_kernel(const double* __restrict__ fsum, const double* __restrict__ fcnt,
                                                         int nslots, int nblocks, int k, int d,
                                                         double* __restrict__ out, const int32_t* state) {
    if (state != nullptr && state[0] != 0) return;
    __shared__ double sh[8][33];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lane;
    const int len = k * (d + 1);
    double t = 0.0;
    if (i < len) {
        const int c = i / (d + 1), f = i - c * (d + 1);
        if (f < d) {
            const int per = (nslots + 7) / 8;
            const int b1 = min(nslots, (grp + 1) * per);
            for (int b = grp * per; b < b1; ++b) t += fsum[(size_t)b * k * d + (size_t)c * d + f];
        } else {
            const int per = (nblocks + 7) / 8;
            const int b1 = min(nblocks, (grp + 1) * per);
            for (int b = grp * per; b < b1; ++b) t += fcnt[(size_t)b * k + c];
        }
    }
    sh[grp][lane] = t;
    __syncthreads();
    if (grp == 0 && i < len) {
        double r = sh[0][lane];
#pragma unroll
        for (int g2 = 1; g2 < 8; ++g2) r += sh[g2][lane];
        out[i] = r;
    }
}
__global__ void reduce_scalar_row_kernel(const double* __restrict__ v, int n, double* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double t = 0.0;
        for (int b = 0; b < n; ++b) t += v[b];
        *out = t;
    }
}

struct RowPlan {
    int S;
    size_t smem;
    bool ok;
};

RowPlan plan_row(const Handle* h, int k, bool sums) {
    RowPlan pl{0, 0, false};
    // a multiple of the residue counts (4 distance-warp groups, 2 accumulator groups): every stage is then always
    // consumed by the same warps, which keeps their one-bit phase parities unambiguous
    for (int S = 12; S >= 4; S -= ER) {
        RowLayout L = row_layout(k, S, sums);
        if (L.total <= (size_t)h->smem_optin) {
            pl.S = S;
            pl.smem = L.total;
            pl.ok = true;
            return pl;
        }
    }
    return pl;
}

template <typename T, bool SUMS>
int launch_row_inst(Handle* h, const CUtensorMap& map, const RowParams& p, size_t smem, int grid, cudaStream_t st) {
    auto kern = lloyd_row128_kernel<T, SUMS>;
    HK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    prof_begin(h, st);
    kern<<<grid, (MISC_WARPS + E_WARPS + (SUMS ? A_WARPS : 0)) * 32, smem, st>>>(map, p);
    prof_end(h, st);
    HK_CUDA(cudaGetLastError());
    h->launches++;
    return 0;
}

}  // namespace

bool row128_supported(const Handle* h, const LloydArgs& a) {
    const int esize = a.dtype == HK_F64 ? 8 : 4;
    if (a.d * esize != 128) return false;
    if (a.k < 1 || a.k > 128) return false;
    if ((a.ldx * esize) % 16 != 0) return false;
    if ((reinterpret_cast<uintptr_t>(a.X) & 15) != 0) return false;
    if (a.n >= (int64_t)1 << 31) return false;
    return plan_row(h, a.k, a.partials != nullptr).ok;
}

int launch_lloyd_row128(Handle* h, const LloydArgs& a) {
    const bool sums = a.partials != nullptr;
    const int esize = a.dtype == HK_F64 ? 8 : 4;
    RowPlan pl = plan_row(h, a.k, sums);
    if (!pl.ok) {
        set_error("lloyd_row128: no feasible plan for k=%d", a.k);
        return -2;
    }
    CUtensorMap map;
    int rc = make_tensor_map_2d(&map, a.X, esize, (uint64_t)a.n, (uint64_t)a.d, (uint64_t)a.ldx, (uint32_t)a.d, TM, 128);
    if (rc) return rc;
    RowParams p{};
    p.n = a.n;
    p.k = a.k;
    p.C = a.C;
    p.labels = a.labels;
    p.label_kind = a.labels ? a.label_kind : HK_LABEL_NONE;
    p.S = pl.S;
    p.num_tiles = (int)((a.n + TM - 1) / TM);
    p.state = a.state;
    {
        const RowLayout L = row_layout(a.k, pl.S, sums);
        p.o_stages = (uint32_t)L.stages;
        p.o_C = (uint32_t)L.C;
        p.o_cn = (uint32_t)L.cn;
        p.o_acc = (uint32_t)L.acc;
        p.o_lab = (uint32_t)L.lab;
        p.o_cnt = (uint32_t)L.cnt;
        p.o_mmax = (uint32_t)L.mmax;
        p.o_bars = (uint32_t)L.bars;
        p.o_misc = (uint32_t)L.misc;
    }
    int grid = h->num_sms;
    if (grid > p.num_tiles) grid = p.num_tiles;
    const int nslots = grid * A_WARPS;
    const size_t kd = (size_t)a.k * a.d;
    rc = ensure_part(h, ((size_t)nslots * kd + (size_t)grid * a.k + grid) * sizeof(double));
    if (rc) return rc;
    p.fsum = h->part;
    p.fcnt = h->part + (size_t)nslots * kd;
    p.fv_part = a.fv_out ? h->part + (size_t)nslots * kd + (size_t)grid * a.k : nullptr;

    char name[96];
    snprintf(name, sizeof(name), "row128<%s,d=%d,k=%d,S=%d,%s>", a.dtype == HK_F64 ? "f64" : "f32", a.d, a.k, pl.S,
             sums ? "sums" : "assign");
    h->variant = name;

    if (a.dtype == HK_F64)
        rc = sums ? launch_row_inst<double, true>(h, map, p, pl.smem, grid, a.stream)
                  : launch_row_inst<double, false>(h, map, p, pl.smem, grid, a.stream);
    else
        rc = sums ? launch_row_inst<float, true>(h, map, p, pl.smem, grid, a.stream)
                  : launch_row_inst<float, false>(h, map, p, pl.smem, grid, a.stream);
    if (rc) return rc;
    if (sums) {
        const int len = a.k * (a.d + 1);
        reduce_row_kernel<<<(len + 31) / 32, 256, 0, a.stream>>>(p.fsum, p.fcnt, nslots, grid, a.k, a.d, a.partials,
                                                                 a.state);
        HK_CUDA(cudaGetLastError());
        h->launches++;
    }
    if (a.fv_out) {
        reduce_scalar_row_kernel<<<1, 32, 0, a.stream>>>(p.fv_part, grid, a.fv_out);
        HK_CUDA(cudaGetLastError());
        h->launches++;
    }
    return 0;
}

}  // namespace hk
