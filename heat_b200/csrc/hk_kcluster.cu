// Other consumers of the "stream the rows of X against a small replicated codebook" pattern:
//   * L1 assignment             <- KMedians / KMedoids._assign_to_cluster with metric = manhattan(expand=True)
//                                  (heat/cluster/kmedians.py:43-50, kmedoids.py:45-52, _kcluster.py:352-370)
//   * per-cluster, per-feature medians by distributed radix selection (counts only travel between ranks)
//                               <- KMedians._update_centroids: ht.median of the rows of every cluster
//                                  (heat/cluster/kmedians.py:60-103, heat/core/statistics.py:1684-1728)
//   * nearest row to each of k points (first index) <- KMedoids._update_centroids (heat/cluster/kmedoids.py:94-110)
//   * k smallest entries per row + class vote       <- KNeighborsClassifier.predict
//                                  (heat/classification/kneighborsclassifier.py:124-135)
// Plain SIMT kernels: these paths are bound by streaming X once per pass (k small) and are not the headline.
#include <math.h>

#include "hk_common.cuh"

namespace hk {
namespace {

constexpr int NT = 256;

// order-preserving integer image of a float (ascending; -0 < +0, NaN with the sign bit clear above +inf)
template <typename T>
struct Key;
template <>
struct Key<float> {
    static constexpr int BITS = 32;
    static __device__ __forceinline__ uint64_t enc(float v) {
        const uint32_t u = __float_as_uint(v);
        return (uint64_t)(u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u));
    }
    static __device__ __forceinline__ float dec(uint64_t k64) {
        const uint32_t k = (uint32_t)k64;
        return __uint_as_float((k >> 31) ? (k ^ 0x80000000u) : ~k);
    }
};
template <>
struct Key<double> {
    static constexpr int BITS = 64;
    static __device__ __forceinline__ uint64_t enc(double v) {
        const uint64_t u = (uint64_t)__double_as_longlong(v);
        return u ^ ((u >> 63) ? 0xFFFFFFFFFFFFFFFFull : 0x8000000000000000ull);
    }
    static __device__ __forceinline__ double dec(uint64_t k) {
        return __longlong_as_double((long long)((k >> 63) ? (k ^ 0x8000000000000000ull) : ~k));
    }
};

__device__ __forceinline__ void store_label(void* labels, int kind, int64_t i, int j) {
    if (kind == HK_LABEL_I64)
        reinterpret_cast<int64_t*>(labels)[i] = j;
    else if (kind == HK_LABEL_I32)
        reinterpret_cast<int32_t*>(labels)[i] = j;
    else if (kind == HK_LABEL_U8)
        reinterpret_cast<uint8_t*>(labels)[i] = (uint8_t)j;
}

// torch.min / argmin order on (value, index): NaN counts as the smallest value, the first occurrence wins
template <typename T>
__device__ __forceinline__ bool better(T a, int64_t ia, T b, int64_t ib) {
    const bool an = a != a, bn = b != b;
    if (an != bn) return an;
    if (!an && a != b) return a < b;
    return ia < ib;
}

// labels[i] = first-index argmin_j sum_f |x_if - c_jf|; fv_part[block] = sum over the block's rows of that minimum
template <typename T, bool CSMEM>
__global__ void __launch_bounds__(NT) assign_l1_kernel(const T* __restrict__ X, int64_t n, int d, int64_t ldx,
                                                       const T* __restrict__ C, int k, void* labels, int label_kind,
                                                       double* __restrict__ fv_part) {
    extern __shared__ unsigned char sm_raw[];
    T* cs = reinterpret_cast<T*>(sm_raw);
    __shared__ double red[NT / 32];
    const int tid = threadIdx.x;
    if (CSMEM) {
        for (int i = tid; i < k * d; i += NT) cs[i] = C[i];
        __syncthreads();
    }
    double fv = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * NT + tid; i < n; i += (int64_t)gridDim.x * NT) {
        const T* x = X + i * ldx;
        T best = T(0);
        int bj = 0;
        for (int j = 0; j < k; ++j) {
            const T* c = CSMEM ? cs + (size_t)j * d : C + (size_t)j * d;
            T s = T(0);
            for (int f = 0; f < d; ++f) s += fabs(x[f] - c[f]);
            if (j == 0 || s < best || (s != s && best == best)) {
                best = s;
                bj = j;
            }
        }
        store_label(labels, label_kind, i, bj);
        fv += (double)best;
    }
    if (fv_part != nullptr) {
        for (int o = 16; o > 0; o >>= 1) fv += __shfl_xor_sync(0xffffffffu, fv, o);
        if ((tid & 31) == 0) red[tid >> 5] = fv;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int w = 0; w < NT / 32; ++w) s += red[w];
            fv_part[blockIdx.x] = s;
        }
    }
}

// Tiled variant of the two L1 passes: a block stages 256 rows (coalesced loads; rows padded to d+1 words so that
// thread == row reads them without bank conflicts) and the k points in shared memory.
//   MODE 0: labels + per-block sum of the row minima (assignment)
//   MODE 1: per point the closest row, first index on ties -> one candidate per block and point
// sum_f |x_f - c_f| with the row in registers (DREG == d, a multiple of 4) and the point read as 16-byte broadcasts
template <typename T, int DREG>
__device__ __forceinline__ T l1_reg(const T (&xr)[DREG > 0 ? DREG : 1], const T* __restrict__ c) {
    T s = T(0);
    constexpr int V = 16 / sizeof(T);
#pragma unroll
    for (int f = 0; f < DREG; f += V) {
        T cv[V];
        *reinterpret_cast<int4*>(cv) = *reinterpret_cast<const int4*>(c + f);
#pragma unroll
        for (int u = 0; u < V; ++u) s += fabs(xr[f + u] - cv[u]);
    }
    return s;
}

template <typename T, int MODE, int DREG>
__global__ void __launch_bounds__(NT) l1_tiled_kernel(const T* __restrict__ X, int64_t n, int d, int64_t ldx,
                                                      const T* __restrict__ C, int k, void* labels, int label_kind,
                                                      double* __restrict__ fv_part, int64_t row_base,
                                                      double* __restrict__ part_d, int64_t* __restrict__ part_i) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int pd = d + 1;
    T* cs = reinterpret_cast<T*>(sm_raw);  // [k][d]
    T* xs = cs + (size_t)k * d;            // [NT][d + 1]
    size_t off = ((size_t)k * d + (size_t)NT * pd) * sizeof(T);
    off = (off + 15) & ~(size_t)15;
    double* wd = reinterpret_cast<double*>(sm_raw + off);  // MODE 1: [warps][k] running best of every warp
    int64_t* wi = reinterpret_cast<int64_t*>(wd + (size_t)(NT / 32) * k);
    __shared__ double red[NT / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < k * d; i += NT) cs[i] = C[i];
    if (MODE == 1)
        for (int i = tid; i < (NT / 32) * k; i += NT) {
            wd[i] = INFINITY;
            wi[i] = INT64_MAX;
        }
    __syncthreads();
    const int64_t ntiles = (n + NT - 1) / NT;
    const int step_r = NT / d, step_f = NT - step_r * d;
    double fv = 0.0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t r0 = t * NT;
        const int rows = (int)((n - r0) < NT ? (n - r0) : NT);
        {
            int r = tid / d, f = tid - r * d;
            for (int e = tid; e < rows * d; e += NT) {
                xs[r * pd + f] = X[(r0 + r) * ldx + f];
                r += step_r;
                f += step_f;
                if (f >= d) {
                    f -= d;
                    ++r;
                }
            }
        }
        __syncthreads();
        const bool act = tid < rows;
        const T* xr = xs + (size_t)tid * pd;
        T xreg[DREG > 0 ? DREG : 1];
        if (DREG > 0) {
#pragma unroll
            for (int f = 0; f < DREG; ++f) xreg[f] = xr[f];
        }
        if (MODE == 0) {
            if (act) {
                T best = T(0);
                int bj = 0;
                for (int j = 0; j < k; ++j) {
                    const T* c = cs + (size_t)j * d;
                    T s = T(0);
                    if (DREG > 0)
                        s = l1_reg<T, DREG>(xreg, c);
                    else
                        for (int f = 0; f < d; ++f) s += fabs(xr[f] - c[f]);
                    if (j == 0 || s < best || (s != s && best == best)) {
                        best = s;
                        bj = j;
                    }
                }
                store_label(labels, label_kind, r0 + tid, bj);
                fv += (double)best;
            }
        } else {
            for (int j = 0; j < k; ++j) {
                double bd = INFINITY;
                int64_t bi = INT64_MAX;
                if (act) {
                    const T* c = cs + (size_t)j * d;
                    T s = T(0);
                    if (DREG > 0)
                        s = l1_reg<T, DREG>(xreg, c);
                    else
                        for (int f = 0; f < d; ++f) s += fabs(xr[f] - c[f]);
                    bd = (double)s;
                    bi = row_base + r0 + tid;
                }
                for (int o = 16; o > 0; o >>= 1) {
                    const double od = __shfl_xor_sync(0xffffffffu, bd, o);
                    const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (better<double>(od, oi, bd, bi)) {
                        bd = od;
                        bi = oi;
                    }
                }
                if (lane == 0 && better<double>(bd, bi, wd[warp * k + j], wi[warp * k + j])) {
                    wd[warp * k + j] = bd;
                    wi[warp * k + j] = bi;
                }
            }
        }
        __syncthreads();
    }
    if (MODE == 0) {
        if (fv_part != nullptr) {
            for (int o = 16; o > 0; o >>= 1) fv += __shfl_xor_sync(0xffffffffu, fv, o);
            if (lane == 0) red[warp] = fv;
            __syncthreads();
            if (tid == 0) {
                double s = 0.0;
                for (int w = 0; w < NT / 32; ++w) s += red[w];
                fv_part[blockIdx.x] = s;
            }
        }
    } else {
        for (int j = tid; j < k; j += NT) {
            double bd = wd[j];
            int64_t bi = wi[j];
            for (int w = 1; w < NT / 32; ++w)
                if (better<double>(wd[w * k + j], wi[w * k + j], bd, bi)) {
                    bd = wd[w * k + j];
                    bi = wi[w * k + j];
                }
            part_d[(size_t)blockIdx.x * k + j] = bd;
            part_i[(size_t)blockIdx.x * k + j] = bi;
        }
    }
}

// rows of 8 / 16 / 32 / 64 features are kept in registers (the centre reads become 16-byte broadcasts)
template <typename T, int MODE>
int launch_l1_tiled(int grid, size_t smem, cudaStream_t st, const T* X, int64_t n, int d, int64_t ldx, const T* C, int k,
                    void* labels, int label_kind, double* fv_part, int64_t row_base, double* part_d, int64_t* part_i) {
#define HK_L1_LAUNCH(DREG)                                                                                              \
    do {                                                                                                               \
        HK_CUDA(cudaFuncSetAttribute(l1_tiled_kernel<T, MODE, DREG>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                     (int)smem));                                                                      \
        l1_tiled_kernel<T, MODE, DREG><<<grid, NT, smem, st>>>(X, n, d, ldx, C, k, labels, label_kind, fv_part,         \
                                                               row_base, part_d, part_i);                              \
    } while (0)
    if (d == 8)
        HK_L1_LAUNCH(8);
    else if (d == 16)
        HK_L1_LAUNCH(16);
    else if (d == 32)
        HK_L1_LAUNCH(32);
    else if (d == 64)
        HK_L1_LAUNCH(64);
    else
        HK_L1_LAUNCH(0);
#undef HK_L1_LAUNCH
    return 0;
}

template <typename T>
size_t l1_tiled_smem(int d, int k, int mode) {
    size_t b = ((size_t)k * d + (size_t)NT * (d + 1)) * sizeof(T);
    b = (b + 15) & ~(size_t)15;
    if (mode == 1) b += (size_t)(NT / 32) * k * (sizeof(double) + sizeof(int64_t));
    return b;
}

__global__ void sum_blocks_kernel(const double* __restrict__ part, int nb, double* __restrict__ out) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double s = 0.0;
        for (int b = 0; b < nb; ++b) s += part[b];  // fixed order: repeatable
        out[0] = s;
    }
}

// keep[i] = 0 for rows that are entirely zero (the reference drops them before the median, kmedians.py:76-79).
// A warp covers rpw = max(1, 32 / d) rows per step and UNR steps per iteration (independent loads in flight: the pass
// is latency-bound otherwise); lanes stride over the features (coalesced), one ballot per step.
constexpr int UNR = 4;
template <typename T>
__global__ void __launch_bounds__(NT) row_keep_kernel(const T* __restrict__ X, int64_t n, int d, int64_t ldx,
                                                      uint8_t* __restrict__ keep) {
    const int lane = threadIdx.x & 31;
    const int dl = d < 32 ? d : 32;   // lanes per row
    const int rpw = 32 / dl;          // rows per warp step
    const int sub = lane / dl, f0 = lane - sub * dl;
    const unsigned gmask = dl == 32 ? 0xffffffffu : ((1u << dl) - 1u);
    const int64_t warp0 = ((int64_t)blockIdx.x * NT + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * NT) >> 5;
    for (int64_t base = warp0 * rpw * UNR; base < n; base += nwarps * rpw * UNR) {
        bool any[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int64_t i = base + (int64_t)u * rpw + sub;
            any[u] = false;
            if (sub < rpw && i < n)
                for (int f = f0; f < d; f += dl) any[u] |= (X[i * ldx + f] != T(0));
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int64_t i = base + (int64_t)u * rpw + sub;
            const unsigned m = __ballot_sync(0xffffffffu, any[u]);
            if (sub < rpw && i < n && f0 == 0) keep[i] = ((m >> (sub * dl)) & gmask) ? 1 : 0;
        }
    }
}

// One digit (8 bits) of the radix selection: for both order statistics w (lower / upper middle) of every (cluster,
// feature), count the kept values of that cluster whose leading digits equal prefix[w][j][f] by their next digit.
// Global-atomics version (any k); rows are taken UNR at a time so that their loads overlap.
template <typename T>
__global__ void __launch_bounds__(NT) select_hist_kernel(const T* __restrict__ X, int64_t n, int d, int64_t ldx,
                                                         const int64_t* __restrict__ labels,
                                                         const uint8_t* __restrict__ keep, int k,
                                                         const uint64_t* __restrict__ prefix, int pass,
                                                         unsigned int* __restrict__ hist) {
    const int shift = Key<T>::BITS - 8 * (pass + 1);
    const int lane = threadIdx.x & 31;
    const int dl = d < 32 ? d : 32;  // lanes per row; a warp covers 32 / dl rows at a time, lanes stride over the features
    const int rpw = 32 / dl;
    const int sub = lane / dl, f0 = lane - sub * dl;
    const int64_t warp0 = ((int64_t)blockIdx.x * NT + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * NT) >> 5;
    if (sub >= rpw) return;
    for (int64_t base = warp0 * rpw * UNR + sub; base < n; base += nwarps * rpw * UNR) {
        int64_t j[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int64_t i = base + (int64_t)u * rpw;
            j[u] = -1;
            if (i < n && keep[i]) j[u] = labels[i];
            if (j[u] >= k) j[u] = -1;
        }
        for (int f = f0; f < d; f += dl) {
            uint64_t key[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u)
                if (j[u] >= 0) key[u] = Key<T>::enc(X[(base + (int64_t)u * rpw) * ldx + f]);
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                if (j[u] < 0) continue;
                const uint64_t lead = pass == 0 ? 0 : (key[u] >> (shift + 8));
                const unsigned digit = (unsigned)((key[u] >> shift) & 255u);
#pragma unroll
                for (int w = 0; w < 2; ++w) {
                    if (pass == 0 && w == 1) break;  // both targets share the first digit's counts (the caller copies hist[0])
                    const size_t e = ((size_t)w * k + (size_t)j[u]) * d + f;
                    if (pass == 0 || lead == prefix[e]) atomicAdd(&hist[e * 256 + digit], 1u);  // 32-bit RED: no return value
                }
            }
        }
    }
}

// Pass 0 with block-private counters: every kept element is counted in this pass, which as reduction atomics to global
// memory costs ~4 ms per 320 M elements.  A block owns a group of G consecutive features (one 32-byte sector of every fp32
// row at G = 8), counts its rows in shared memory ([k][G][256] counters, thread == row) and flushes the non-zero counters
// once: 1.5 ms.  (Measured alternatives that were slower: the same kernel for the later passes with two targets in 128 KB
// of shared memory, 3.4-3.8 ms per pass; thread == (row, feature) with 64 KB, 2.5 ms; a 12-bit first digit, no change.)
template <typename T>
__global__ void __launch_bounds__(NT) select_hist0_smem_kernel(const T* __restrict__ X, int64_t n, int d, int64_t ldx,
                                                               const int64_t* __restrict__ labels,
                                                               const uint8_t* __restrict__ keep, int k, int G,
                                                               unsigned int* __restrict__ hist) {
    extern __shared__ unsigned int cnt[];  // [k][G][256]
    const int tid = threadIdx.x;
    const int f0 = blockIdx.y * G;
    const int gw = (d - f0) < G ? (d - f0) : G;  // features of this group
    const int total = k * G * 256;
    for (int i = tid; i < total; i += NT) cnt[i] = 0;
    __syncthreads();
    constexpr int shift = Key<T>::BITS - 8;
    for (int64_t i = (int64_t)blockIdx.x * NT + tid; i < n; i += (int64_t)gridDim.x * NT) {
        if (!keep[i]) continue;
        const int64_t j = labels[i];
        if (j < 0 || j >= k) continue;
        const T* x = X + i * ldx + f0;
        for (int g = 0; g < gw; ++g) {
            const unsigned digit = (unsigned)(Key<T>::enc(x[g]) >> shift);
            atomicAdd(&cnt[((int)j * G + g) * 256 + digit], 1u);
        }
    }
    __syncthreads();
    for (int i = tid; i < total; i += NT) {
        const unsigned c = cnt[i];
        if (c == 0) continue;
        const int digit = i & 255, g = (i >> 8) % G, j = (i >> 8) / G;
        if (g < gw) atomicAdd(&hist[((size_t)j * d + f0 + g) * 256 + digit], c);  // target w = 0 (the caller copies it to w = 1)
    }
}

// the shard's 32-bit counts (a bin holds at most n_local < 2^32 values) are added to the caller's 64-bit histogram
__global__ void widen_add_kernel(const unsigned int* __restrict__ h32, unsigned long long* __restrict__ h64, size_t entries) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < entries) h64[i] += h32[i];
}

// choose the digit that holds the wanted rank, descend into it
__global__ void select_step_kernel(const unsigned long long* __restrict__ hist, int64_t* __restrict__ remaining,
                                   uint64_t* __restrict__ prefix, int entries) {
    const int base = blockIdx.x * blockDim.x + threadIdx.x;
    if (base >= entries) return;
    int64_t rem = remaining[base];
    unsigned digit = 255;
    for (unsigned b = 0; b < 256; ++b) {
        const int64_t c = (int64_t)hist[(size_t)base * 256 + b];
        if (rem < c) {
            digit = b;
            break;
        }
        rem -= c;
    }
    remaining[base] = rem;
    prefix[base] = (prefix[base] << 8) | digit;
}

// median = lo + (hi - lo) * frac (heat/core/statistics.py:1708-1728); frac[j] = 0.5 for an even count, else 0
template <typename T>
__global__ void select_value_kernel(const uint64_t* __restrict__ prefix, const double* __restrict__ frac, int k, int d,
                                    T* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= k * d) return;
    const int j = e / d;
    const T lo = Key<T>::dec(prefix[e]);
    const T hi = Key<T>::dec(prefix[(size_t)k * d + e]);
    out[e] = lo + (hi - lo) * (T)frac[j];
}

// per block and point j: the row (first index) with the smallest L1 distance to P[j]
template <typename T>
__global__ void __launch_bounds__(NT) nearest_rows_l1_kernel(const T* __restrict__ X, int64_t n, int d, int64_t ldx,
                                                             const T* __restrict__ P, int k, int64_t row_base,
                                                             double* __restrict__ part_d, int64_t* __restrict__ part_i) {
    __shared__ double sd[NT / 32];
    __shared__ int64_t si[NT / 32];
    const int tid = threadIdx.x;
    for (int j = 0; j < k; ++j) {
        const T* p = P + (size_t)j * d;
        double bd = INFINITY;
        int64_t bi = INT64_MAX;
        for (int64_t i = (int64_t)blockIdx.x * NT + tid; i < n; i += (int64_t)gridDim.x * NT) {
            const T* x = X + i * ldx;
            T s = T(0);
            for (int f = 0; f < d; ++f) s += fabs(x[f] - p[f]);
            if (better<double>((double)s, row_base + i, bd, bi)) {
                bd = (double)s;
                bi = row_base + i;
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, bd, o);
            const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (better<double>(od, oi, bd, bi)) {
                bd = od;
                bi = oi;
            }
        }
        if ((tid & 31) == 0) {
            sd[tid >> 5] = bd;
            si[tid >> 5] = bi;
        }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < NT / 32; ++w)
                if (better<double>(sd[w], si[w], bd, bi)) {
                    bd = sd[w];
                    bi = si[w];
                }
            part_d[(size_t)blockIdx.x * k + j] = bd;
            part_i[(size_t)blockIdx.x * k + j] = bi;
        }
        __syncthreads();
    }
}

__global__ void nearest_final_kernel(const double* __restrict__ part_d, const int64_t* __restrict__ part_i, int nb, int k,
                                     double* __restrict__ out_d, int64_t* __restrict__ out_i) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k) return;
    double bd = part_d[j];
    int64_t bi = part_i[j];
    for (int b = 1; b < nb; ++b) {
        const double od = part_d[(size_t)b * k + j];
        const int64_t oi = part_i[(size_t)b * k + j];
        if (better<double>(od, oi, bd, bi)) {
            bd = od;
            bi = oi;
        }
    }
    out_d[j] = bd;
    out_i[j] = bi;
}

// one warp per row: the kk smallest (value, index) pairs in ascending order; NaN entries are never selected before a
// number (torch.topk(largest=False) ranks NaN as the largest value)
template <typename T>
__global__ void __launch_bounds__(NT) topk_rows_kernel(const T* __restrict__ D, int64_t m, int64_t n, int64_t ldd, int kk,
                                                       T* __restrict__ vals, int64_t* __restrict__ idx) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * NT + threadIdx.x) >> 5;
    if (row >= m) return;
    const T* r = D + row * ldd;
    T last_v = T(0);
    int64_t last_i = -1;
    for (int t = 0; t < kk; ++t) {
        T bv = T(0);
        int64_t bi = -1;  // -1: nothing found yet
        for (int64_t c = lane; c < n; c += 32) {
            const T v = r[c];
            // candidates: strictly after (last_v, last_i) in (value, index) order; NaN sorts after every number
            const bool vn = v != v;
            bool after;
            if (last_i < 0)
                after = true;
            else {
                const bool ln = last_v != last_v;
                if (vn != ln)
                    after = vn;
                else if (!vn && v != last_v)
                    after = v > last_v;
                else
                    after = c > last_i;
            }
            if (!after) continue;
            bool take;
            if (bi < 0)
                take = true;
            else {
                const bool bn = bv != bv;
                if (vn != bn)
                    take = bn;
                else if (!vn && v != bv)
                    take = v < bv;
                else
                    take = c < bi;
            }
            if (take) {
                bv = v;
                bi = c;
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const T ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oi < 0) continue;
            bool take;
            if (bi < 0)
                take = true;
            else {
                const bool on = ov != ov, bn = bv != bv;
                if (on != bn)
                    take = bn;
                else if (!on && ov != bv)
                    take = ov < bv;
                else
                    take = oi < bi;
            }
            if (take) {
                bv = ov;
                bi = oi;
            }
        }
        if (lane == 0) {
            vals[row * kk + t] = bv;
            idx[row * kk + t] = bi;
        }
        last_v = bv;
        last_i = bi;
    }
}

// classes[i] = first-index argmax_c sum_t Y[idx[i,t], c]   (one-hot or soft label rows, kneighborsclassifier.py:128-134)
template <typename T>
__global__ void __launch_bounds__(NT) knn_vote_kernel(const int64_t* __restrict__ idx, int64_t m, int kk,
                                                      const T* __restrict__ Y, int64_t n, int nc, int64_t ldy,
                                                      int64_t* __restrict__ classes) {
    const int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x;
    if (i >= m) return;
    T best = T(0);
    int bc = 0;
    for (int c = 0; c < nc; ++c) {
        T s = T(0);
        for (int t = 0; t < kk; ++t) {
            const int64_t r = idx[i * kk + t];
            if (r >= 0 && r < n) s += Y[r * ldy + c];
        }
        if (c == 0 || s > best || (s != s && best == best)) {  // torch.max: NaN wins, first occurrence
            best = s;
            bc = c;
        }
    }
    classes[i] = bc;
}

// Centre update of the batch-parallel clusterers' single-process loop (_kmex, heat/cluster/batchparallelclustering.py:67-84):
// a cluster with rows takes their mean (from the fp64 partial sums | counts of a Lloyd pass) or their median (precomputed),
// an empty cluster keeps its centre; flag[0] = torch.allclose(new, old, atol=tol) = all |new - old| <= atol + rtol |old|.
template <typename T>
__global__ void kmex_update_kernel(const double* __restrict__ partials, const T* __restrict__ medians,
                                   const int64_t* __restrict__ counts, T* __restrict__ C, int k, int d, double atol,
                                   double rtol, int* __restrict__ flag) {
    __shared__ int bad;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    int mine = 0;
    for (int e = threadIdx.x; e < k * d; e += blockDim.x) {
        const int j = e / d, f = e - j * d;
        const T old = C[e];
        T nv = old;
        if (partials != nullptr) {
            const double cnt = partials[(size_t)j * (d + 1) + d];
            if (cnt > 0.0) nv = (T)(partials[(size_t)j * (d + 1) + f] / cnt);
        } else if (counts[j] > 0) {
            nv = medians[e];
        }
        C[e] = nv;
        const double diff = fabs((double)nv - (double)old);
        if (!(diff <= atol + rtol * fabs((double)old))) mine = 1;  // NaN counts as "not close"
    }
    if (mine) atomicOr(&bad, 1);
    __syncthreads();
    if (threadIdx.x == 0) flag[0] = bad ? 0 : 1;
}

int blocks_for(const Handle* h, int64_t work_items) {
    int64_t b = (work_items + NT - 1) / NT;
    const int64_t cap = (int64_t)h->num_sms * 8;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

template <typename T>
int run_assign_l1(Handle* h, const T* X, int64_t n, int d, int64_t ldx, const T* C, int k, void* labels, int label_kind,
                  double* fv, cudaStream_t st) {
    const size_t tsz = l1_tiled_smem<T>(d, k, 0);
    const bool tiled = tsz <= 160 * 1024;
    int grid = blocks_for(h, n);
    if (tiled) {
        const int64_t ntiles = (n + NT - 1) / NT;
        grid = (int)(ntiles < (int64_t)h->num_sms * 4 ? ntiles : (int64_t)h->num_sms * 4);
    }
    double* part = nullptr;
    if (fv != nullptr) {
        int rc = ensure_part(h, (size_t)grid * sizeof(double) + 64);
        if (rc) return rc;
        part = reinterpret_cast<double*>(h->part);
    }
    const size_t csz = (size_t)k * d * sizeof(T);
    prof_begin(h, st);
    if (tiled) {
        int rc = launch_l1_tiled<T, 0>(grid, tsz, st, X, n, d, ldx, C, k, labels, label_kind, part, 0, nullptr, nullptr);
        if (rc) return rc;
    } else if (csz <= 96 * 1024) {
        HK_CUDA(cudaFuncSetAttribute(assign_l1_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csz));
        assign_l1_kernel<T, true><<<grid, NT, csz, st>>>(X, n, d, ldx, C, k, labels, label_kind, part);
    } else {
        assign_l1_kernel<T, false><<<grid, NT, 0, st>>>(X, n, d, ldx, C, k, labels, label_kind, part);
    }
    prof_end(h, st);
    HK_CUDA(cudaGetLastError());
    h->launches++;
    if (fv != nullptr) {
        sum_blocks_kernel<<<1, 32, 0, st>>>(part, grid, fv);
        HK_CUDA(cudaGetLastError());
        h->launches++;
    }
    return 0;
}

}  // namespace

int launch_assign_l1(Handle* h, const void* X, int64_t n, int d, int64_t ldx, int dtype, const void* C, int k,
                     void* labels, int label_kind, double* fv, cudaStream_t st) {
    h->variant = dtype == HK_F64 ? "assign_l1<f64>" : "assign_l1<f32>";
    if (dtype == HK_F64)
        return run_assign_l1<double>(h, (const double*)X, n, d, ldx, (const double*)C, k, labels, label_kind, fv, st);
    return run_assign_l1<float>(h, (const float*)X, n, d, ldx, (const float*)C, k, labels, label_kind, fv, st);
}

int launch_row_keep(Handle* h, const void* X, int64_t n, int d, int64_t ldx, int dtype, uint8_t* keep, cudaStream_t st) {
    const int dl = d < 32 ? d : 32;
    const unsigned grid = (unsigned)blocks_for(h, (n + (32 / dl) * UNR - 1) / ((32 / dl) * UNR) * 32);
    if (dtype == HK_F64)
        row_keep_kernel<double><<<grid, NT, 0, st>>>((const double*)X, n, d, ldx, keep);
    else
        row_keep_kernel<float><<<grid, NT, 0, st>>>((const float*)X, n, d, ldx, keep);
    HK_CUDA(cudaGetLastError());
    h->launches++;
    return 0;
}

int launch_select_hist(Handle* h, const void* X, int64_t n, int d, int64_t ldx, int dtype, const int64_t* labels,
                       const uint8_t* keep, int k, const uint64_t* prefix, int pass, unsigned long long* hist,
                       cudaStream_t st) {
    const int dl = d < 32 ? d : 32;
    const int grid = blocks_for(h, (n + (32 / dl) - 1) / (32 / dl) * 32);
    const size_t entries = (size_t)2 * k * d * 256;
    int rc = ensure_part(h, entries * sizeof(unsigned int) + 64);
    if (rc) return rc;
    unsigned int* h32 = reinterpret_cast<unsigned int*>(h->part);
    HK_CUDA(cudaMemsetAsync(h32, 0, entries * sizeof(unsigned int), st));
    // pass 0: block-private counters when a group of at least one feature fits in shared memory
    int G = 0;
    if (pass == 0) {
        const int per_sector = dtype == HK_F64 ? 4 : 8;
        for (int g = per_sector; g >= 1; g >>= 1)
            if ((size_t)k * g * 256 * sizeof(unsigned int) <= 96 * 1024) {
                G = g;
                break;
            }
    }
    if (G > 0) {
        const size_t smem = (size_t)k * G * 256 * sizeof(unsigned int);
        const int groups = (d + G - 1) / G;
        int bx = (h->num_sms * 2 + groups - 1) / groups;
        const int64_t need = (n + NT - 1) / NT;
        if (bx > need) bx = (int)(need < 1 ? 1 : need);
        dim3 g2((unsigned)bx, (unsigned)groups);
        if (dtype == HK_F64) {
            HK_CUDA(cudaFuncSetAttribute(select_hist0_smem_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            select_hist0_smem_kernel<double><<<g2, NT, smem, st>>>((const double*)X, n, d, ldx, labels, keep, k, G, h32);
        } else {
            HK_CUDA(cudaFuncSetAttribute(select_hist0_smem_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            select_hist0_smem_kernel<float><<<g2, NT, smem, st>>>((const float*)X, n, d, ldx, labels, keep, k, G, h32);
        }
    } else if (dtype == HK_F64)
        select_hist_kernel<double><<<grid, NT, 0, st>>>((const double*)X, n, d, ldx, labels, keep, k, prefix, pass, h32);
    else
        select_hist_kernel<float><<<grid, NT, 0, st>>>((const float*)X, n, d, ldx, labels, keep, k, prefix, pass, h32);
    HK_CUDA(cudaGetLastError());
    widen_add_kernel<<<(unsigned)((entries + 255) / 256), 256, 0, st>>>(h32, hist, entries);
    HK_CUDA(cudaGetLastError());
    h->launches += 2;
    h->variant = dtype == HK_F64 ? "select_hist<f64>" : "select_hist<f32>";
    return 0;
}

int launch_select_step(Handle* h, const unsigned long long* hist, int64_t* remaining, uint64_t* prefix, int entries,
                       cudaStream_t st) {
    select_step_kernel<<<(entries + 127) / 128, 128, 0, st>>>(hist, remaining, prefix, entries);
    HK_CUDA(cudaGetLastError());
    h->launches++;
    return 0;
}

int launch_select_value(Handle* h, const uint64_t* prefix, const double* frac, int k, int d, int dtype, void* out,
                        cudaStream_t st) {
    const int e = k * d;
    if (dtype == HK_F64)
        select_value_kernel<double><<<(e + 127) / 128, 128, 0, st>>>(prefix, frac, k, d, (double*)out);
    else
        select_value_kernel<float><<<(e + 127) / 128, 128, 0, st>>>(prefix, frac, k, d, (float*)out);
    HK_CUDA(cudaGetLastError());
    h->launches++;
    return 0;
}

int launch_kmex_update(Handle* h, const double* partials, const void* medians, const int64_t* counts, void* C, int k,
                       int d, int dtype, double atol, double rtol, int* flag, cudaStream_t st) {
    if (dtype == HK_F64)
        kmex_update_kernel<double><<<1, 256, 0, st>>>(partials, (const double*)medians, counts, (double*)C, k, d, atol, rtol, flag);
    else
        kmex_update_kernel<float><<<1, 256, 0, st>>>(partials, (const float*)medians, counts, (float*)C, k, d, atol, rtol, flag);
    HK_CUDA(cudaGetLastError());
    h->launches++;
    return 0;
}

int launch_nearest_rows_l1(Handle* h, const void* X, int64_t n, int d, int64_t ldx, int dtype, const void* P, int k,
                           int64_t row_base, double* out_d, int64_t* out_i, cudaStream_t st) {
    const size_t tsz = dtype == HK_F64 ? l1_tiled_smem<double>(d, k, 1) : l1_tiled_smem<float>(d, k, 1);
    const bool tiled = tsz <= 160 * 1024;
    int grid = blocks_for(h, n);
    if (tiled) {
        const int64_t ntiles = (n + NT - 1) / NT;
        grid = (int)(ntiles < (int64_t)h->num_sms * 4 ? ntiles : (int64_t)h->num_sms * 4);
    }
    const size_t need = (size_t)grid * k * (sizeof(double) + sizeof(int64_t)) + 64;
    int rc = ensure_part(h, need);
    if (rc) return rc;
    double* pd = reinterpret_cast<double*>(h->part);
    int64_t* pi = reinterpret_cast<int64_t*>(pd + (size_t)grid * k);
    if (tiled && dtype == HK_F64) {
        rc = launch_l1_tiled<double, 1>(grid, tsz, st, (const double*)X, n, d, ldx, (const double*)P, k, nullptr, 0, nullptr,
                                        row_base, pd, pi);
        if (rc) return rc;
    } else if (tiled) {
        rc = launch_l1_tiled<float, 1>(grid, tsz, st, (const float*)X, n, d, ldx, (const float*)P, k, nullptr, 0, nullptr,
                                       row_base, pd, pi);
        if (rc) return rc;
    } else if (dtype == HK_F64)
        nearest_rows_l1_kernel<double><<<grid, NT, 0, st>>>((const double*)X, n, d, ldx, (const double*)P, k, row_base, pd, pi);
    else
        nearest_rows_l1_kernel<float><<<grid, NT, 0, st>>>((const float*)X, n, d, ldx, (const float*)P, k, row_base, pd, pi);
    HK_CUDA(cudaGetLastError());
    nearest_final_kernel<<<(k + 127) / 128, 128, 0, st>>>(pd, pi, grid, k, out_d, out_i);
    HK_CUDA(cudaGetLastError());
    h->launches += 2;
    h->variant = dtype == HK_F64 ? "nearest_rows_l1<f64>" : "nearest_rows_l1<f32>";
    return 0;
}

int launch_topk_rows(Handle* h, const void* D, int64_t m, int64_t n, int64_t ldd, int dtype, int kk, void* vals,
                     int64_t* idx, cudaStream_t st) {
    const unsigned grid = (unsigned)((m * 32 + NT - 1) / NT);
    if (dtype == HK_F64)
        topk_rows_kernel<double><<<grid, NT, 0, st>>>((const double*)D, m, n, ldd, kk, (double*)vals, idx);
    else
        topk_rows_kernel<float><<<grid, NT, 0, st>>>((const float*)D, m, n, ldd, kk, (float*)vals, idx);
    HK_CUDA(cudaGetLastError());
    h->launches++;
    h->variant = dtype == HK_F64 ? "topk_rows<f64>" : "topk_rows<f32>";
    return 0;
}

int launch_knn_vote(Handle* h, const int64_t* idx, int64_t m, int kk, const void* Y, int64_t n, int nc, int64_t ldy,
                    int dtype, int64_t* classes, cudaStream_t st) {
    const unsigned grid = (unsigned)((m + NT - 1) / NT);
    if (dtype == HK_F64)
        knn_vote_kernel<double><<<grid, NT, 0, st>>>(idx, m, kk, (const double*)Y, n, nc, ldy, classes);
    else
        knn_vote_kernel<float><<<grid, NT, 0, st>>>(idx, m, kk, (const float*)Y, n, nc, ldy, classes);
    HK_CUDA(cudaGetLastError());
    h->launches++;
    return 0;
}

}  // namespace hk
