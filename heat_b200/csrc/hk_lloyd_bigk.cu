// Large-k Lloyd pass (fp32, k too large for the fused tensor-core kernel: the accumulators of k*(d+1) sums no longer
// fit in shared memory and the centroid operand no longer fits next to a row tile).  BASELINE configs[3] shape class
// (N=20M, d=128, k=1024).  Replaces the same reference code as the fused kernels
// (heat/cluster/_kcluster.py:352-370 + heat/cluster/kmeans.py:76-103), as a short sequence of passes:
//
//   1. distances with the tcgen05 3xTF32 kernel of hk_cdist_tc.cu, ARGMIN variant: the epilogue keeps a running
//      first-index minimum per row instead of writing the [rows x k] tile (1M rows per launch: bounded by the xl scratch);
//   2. (only for a last chunk below 1024 rows: exact cdist kernel + argmin_rows_kernel, every row queued.)  Rows whose
//      runner-up is within the rounding window of the 3xTF32 product, or that contain NaN, are queued for
//   3. exact_fix_kernel: exact fp32 formula of heat/spatial/distance.py:59-64 over all centroids with torch.min tie/NaN
//      semantics (one warp per queued row) - so labels are those of the exact-FMA kernels;
//   4. cluster sums without atomics on floating point data: counting sort of the row indices by label (integer atomics),
//      then gather_reduce_kernel: every cluster's rows are summed in fp64 registers by SPLIT CTAs, and a last small
//      kernel adds the SPLIT partial sums in a fixed order.  (The order of rows inside a cluster is whatever the scatter
//      produced: fp64 sums may differ in the last bit from run to run; the fused kernels are bitwise reproducible.)
#include <math.h>

#include "hk_common.cuh"

namespace hk {

int launch_cdist_tc(Handle* h, const void* X, int64_t m, int f, int64_t ldx, const void* Y, int64_t n, int64_t ldy,
                    void* out, int64_t ldo, int post, float gscale, cudaStream_t st);
int launch_cdist_tc_argmin(Handle* h, const void* X, int64_t m, int f, int64_t ldx, const void* Y, int64_t n, int64_t ldy,
                           int32_t* labels, int64_t row_base, int32_t* queue, int* qcount, const float* cmax2, float window,
                           const int32_t* state, cudaStream_t st);
bool cdist_tc_supported(const Handle* h, const void* X, int64_t m, int f, int64_t ldx, const void* Y, int64_t n,
                        int64_t ldy, const void* out, int64_t ldo);

namespace {

constexpr int SPLIT = 4;  // CTAs per cluster in the gather-reduce pass

// |c_j|^2 with features in ascending order (same arithmetic as the fused kernels) and max_j |c_j|^2
__global__ void centroid_norms_kernel(const float* __restrict__ C, int k, int d, float* __restrict__ cn,
                                      float* __restrict__ cmax2) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k) return;
    float s = 0.f;
    for (int f = 0; f < d; ++f) {
        const float c = C[(size_t)j * d + f];
        s = fmaf(c, c, s);
    }
    cn[j] = s;
    // s >= 0: int order == float order.  A NaN/Inf centroid makes the bound infinite: every row then goes through the
    // exact kernel, which applies torch.min's NaN rule over all centroids
    atomicMax(reinterpret_cast<int*>(cmax2), s == s ? __float_as_int(s) : 0x7f800000);
}

struct Cand {
    float best, second;
    int idx;
};
__device__ __forceinline__ void cand_push(Cand& c, float v, int j) {
    if (v < c.best) {
        c.second = c.best;
        c.best = v;
        c.idx = j;
    } else {
        c.second = fminf(c.second, v);
    }
}
__device__ __forceinline__ void cand_merge(Cand& a, float ob, float os, int oi) {
    // smaller value wins, ties go to the smaller index; the loser's best is a runner-up candidate
    const bool take = ob < a.best || (ob == a.best && oi < a.idx);
    const float lose = take ? a.best : ob;
    a.second = fminf(fminf(a.second, os), lose);
    if (take) {
        a.best = ob;
        a.idx = oi;
    }
}

// D: [rows x k] squared distances (already clamped at 0).  One warp per row.
// labels[row0 + r] = first-index argmin; rows that need the exact formula are appended to `queue`.
__global__ void __launch_bounds__(256) argmin_rows_kernel(const float* __restrict__ D, int rows, int k, int64_t ldd,
                                                          int64_t row0, int64_t n_total,
                                                          const float* __restrict__ xn, const float* __restrict__ cmax2,
                                                          float window, int force_all, int32_t* __restrict__ labels,
                                                          int32_t* __restrict__ queue, int* __restrict__ qcount,
                                                          double* __restrict__ fv_blocks, const int32_t* state) {
    if (state != nullptr && state[0] != 0) return;
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int r = blockIdx.x * 8 + wib;
    double fv = 0.0;
    if (r < rows) {
        const float* dr = D + (size_t)r * ldd;
        Cand c{INFINITY, INFINITY, 0};
        bool nan = false;
        for (int j0 = lane * 4; j0 < k; j0 += 128) {
            const float4 v = *reinterpret_cast<const float4*>(dr + j0);  // k % 4 == 0 (cdist_tc_supported)
            nan |= (v.x != v.x) | (v.y != v.y) | (v.z != v.z) | (v.w != v.w);
            cand_push(c, v.x, j0);
            cand_push(c, v.y, j0 + 1);
            cand_push(c, v.z, j0 + 2);
            cand_push(c, v.w, j0 + 3);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, c.best, o);
            const float os = __shfl_xor_sync(0xffffffffu, c.second, o);
            const int oi = __shfl_xor_sync(0xffffffffu, c.idx, o);
            cand_merge(c, ob, os, oi);
        }
        nan = __any_sync(0xffffffffu, nan);
        if (lane == 0) {
            // rounding window of the 3xTF32 product and of the exact formula itself, on the scale of d^2
            const float scale = (force_all ? 0.f : xn[r]) + *cmax2;
            const bool undecided = force_all || nan || !(c.second - c.best > window * scale);
            labels[row0 + r] = c.idx;
            if (undecided) queue[atomicAdd(qcount, 1)] = (int32_t)(row0 + r);
            fv = (double)c.best;
        }
    }
    if (fv_blocks != nullptr) {
        __shared__ double sh[8];
        if (lane == 0) sh[wib] = fv;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < 8; ++w) t += sh[w];
            fv_blocks[blockIdx.x] = t;  // fixed order inside the block; blocks are added in order afterwards
        }
    }
}

// exact re-evaluation of the queued rows: lane l takes centroids l, l+32, ...; first-index argmin with torch.min NaN rule
__global__ void __launch_bounds__(256) exact_fix_kernel(const float* __restrict__ X, int d, int64_t ldx,
                                                        const float* __restrict__ C, const float* __restrict__ cn, int k,
                                                        const int32_t* __restrict__ queue, const int* __restrict__ qcount,
                                                        int32_t* __restrict__ labels, const int32_t* state) {
    if (state != nullptr && state[0] != 0) return;
    const int lane = threadIdx.x & 31;
    const int nq = *qcount;
    for (int qi = blockIdx.x * 8 + (threadIdx.x >> 5); qi < nq; qi += gridDim.x * 8) {
        const int64_t row = queue[qi];
        const float* x = X + (size_t)row * ldx;
        float xn = 0.f;
        for (int f = 0; f < d; ++f) xn = fmaf(x[f], x[f], xn);
        float best = INFINITY;
        int bl = 0x7fffffff;
        bool have = false;
        for (int j = lane; j < k; j += 32) {
            const float* c = C + (size_t)j * d;
            float dot = 0.f;
            for (int f = 0; f < d; ++f) dot = fmaf(x[f], c[f], dot);
            float d2 = (xn + cn[j]) - 2.f * dot;
            d2 = d2 < 0.f ? 0.f : d2;
            if (!have || d2 < best || (d2 != d2 && best == best)) {
                best = d2;
                bl = j;
                have = true;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int ol = __shfl_xor_sync(0xffffffffu, bl, o);
            const bool on = ob != ob, bn = best != best;
            // sequential torch.min semantics: the first NaN wins, otherwise the smallest value, ties to the first index
            const bool take = (on && bn) ? (ol < bl) : (on ? true : (bn ? false : (ob < best || (ob == best && ol < bl))));
            if (take) {
                best = ob;
                bl = ol;
            }
        }
        if (lane == 0) labels[row] = bl;
    }
}

// functional value: exact fp32 distance of every row to the centroid of its final label (the 3xTF32 distances carry a
// truncation bias of ~2e-6 * |x||c| per row, which does not average out over 1e7 rows).  One thread per row, features in
// ascending order as in the fused kernels; block partials in a fixed order.
__global__ void __launch_bounds__(256) exact_fv_kernel(const float* __restrict__ X, int64_t n, int d, int64_t ldx,
                                                       const float* __restrict__ C, const float* __restrict__ cn,
                                                       const int32_t* __restrict__ labels, double* __restrict__ fv_blocks,
                                                       const int32_t* state) {
    if (state != nullptr && state[0] != 0) return;
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double v = 0.0;
    if (row < n) {
        const float4* x = reinterpret_cast<const float4*>(X + (size_t)row * ldx);
        const int l = labels[row];
        const float4* c = reinterpret_cast<const float4*>(C + (size_t)l * d);
        float xn = 0.f, dot = 0.f;
        for (int q = 0; q < (d >> 2); ++q) {
            const float4 xv = __ldg(x + q), cv = __ldg(c + q);
            xn = fmaf(xv.x, xv.x, xn);
            xn = fmaf(xv.y, xv.y, xn);
            xn = fmaf(xv.z, xv.z, xn);
            xn = fmaf(xv.w, xv.w, xn);
            dot = fmaf(xv.x, cv.x, dot);
            dot = fmaf(xv.y, cv.y, dot);
            dot = fmaf(xv.z, cv.z, dot);
            dot = fmaf(xv.w, cv.w, dot);
        }
        float d2 = (xn + cn[l]) - 2.f * dot;
        d2 = d2 < 0.f ? 0.f : d2;
        const float sq = sqrtf(d2);
        v = (double)(sq * sq);
    }
    __shared__ double sh[256];
    sh[threadIdx.x] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 256; ++i) t += sh[i];
        fv_blocks[blockIdx.x] = t;
    }
}

__global__ void sum_blocks_kernel(const double* __restrict__ v, int n, double* __restrict__ out, int accumulate,
                                  const int32_t* state) {
    if (state != nullptr && state[0] != 0) return;
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double t = accumulate ? *out : 0.0;
        for (int b = 0; b < n; ++b) t += v[b];
        *out = t;
    }
}

__global__ void store_labels_kernel(const int32_t* __restrict__ lab, int64_t n, void* out, int kind, const int32_t* state) {
    if (state != nullptr && state[0] != 0) return;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (kind == HK_LABEL_I64)
        reinterpret_cast<long long*>(out)[i] = lab[i];
    else if (kind == HK_LABEL_I32)
        reinterpret_cast<int*>(out)[i] = lab[i];
    else if (kind == HK_LABEL_U8)
        reinterpret_cast<unsigned char*>(out)[i] = (unsigned char)lab[i];
}

// ---- counting sort of the row indices by label --------------------------------------------------------
__global__ void __launch_bounds__(256) label_hist_kernel(const int32_t* __restrict__ lab, int64_t n, int k,
                                                         int* __restrict__ hist, const int32_t* state) {
    if (state != nullptr && state[0] != 0) return;
    extern __shared__ int sh[];
    for (int j = threadIdx.x; j < k; j += blockDim.x) sh[j] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&sh[lab[i]], 1);
    __syncthreads();
    for (int j = threadIdx.x; j < k; j += blockDim.x)
        if (sh[j]) atomicAdd(&hist[j], sh[j]);
}
// offsets[j] = sum_{i<j} hist[i]; cursor[j] = 0  (one block)
__global__ void __launch_bounds__(1024) label_scan_kernel(const int* __restrict__ hist, int k, int* __restrict__ offsets,
                                                          int* __restrict__ cursor, const int32_t* state) {
    if (state != nullptr && state[0] != 0) return;
    __shared__ int sh[1024];
    int carry = 0;
    for (int base = 0; base < k; base += 1024) {
        const int j = base + threadIdx.x;
        const int v = j < k ? hist[j] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const int t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        if (j < k) {
            offsets[j] = carry + sh[threadIdx.x] - v;
            cursor[j] = 0;
        }
        carry += sh[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[k] = carry;
}
__global__ void __launch_bounds__(256) label_scatter_kernel(const int32_t* __restrict__ lab, int64_t n,
                                                            const int* __restrict__ offsets, int* __restrict__ cursor,
                                                            int32_t* __restrict__ perm, const int32_t* state) {
    if (state != nullptr && state[0] != 0) return;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int l = lab[i];
        perm[offsets[l] + atomicAdd(&cursor[l], 1)] = (int32_t)i;
    }
}

// CTA (c, s): rows perm[offsets[c] + s, + SPLIT, ...) of cluster c.  Lanes cover d/4 feature quads, 32/(d/4) rows per
// warp step; fp64 accumulators in registers; the 8 warps are combined in a fixed order through shared memory.
__global__ void __launch_bounds__(256) gather_reduce_kernel(const float* __restrict__ X, int d, int64_t ldx,
                                                            const int32_t* __restrict__ perm,
                                                            const int* __restrict__ offsets, int k,
                                                            double* __restrict__ partial, const int32_t* state) {
    if (state != nullptr && state[0] != 0) return;
    const int c = blockIdx.x / SPLIT, sp = blockIdx.x - c * SPLIT;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int lpr = d >> 2;          // lanes per row (d % 4 == 0, d <= 128)
    const int rps = 32 / lpr;        // rows per warp step
    const int rslot = lane / lpr;    // row slot of this lane
    const int fq = lane - rslot * lpr;
    const bool lane_on = rslot < rps;
    const int beg = offsets[c], end = offsets[c + 1];
    // groups of rps consecutive members; group g belongs to CTA g % SPLIT, there to warp (g / SPLIT) % 8
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    for (int it0 = 0;; it0 += 4) {
        const int g0 = sp + SPLIT * (w + 8 * it0);
        if (beg + g0 * rps >= end) break;  // warp-uniform
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int m = beg + (g0 + u * SPLIT * 8) * rps + rslot;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (lane_on && m < end) {
                const int64_t row = perm[m];
                v[u] = __ldg(reinterpret_cast<const float4*>(X + (size_t)row * ldx) + fq);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            a0 += (double)v[u].x;
            a1 += (double)v[u].y;
            a2 += (double)v[u].z;
            a3 += (double)v[u].w;
        }
    }
    __shared__ double sh[8][32][4];
    sh[w][lane][0] = a0;
    sh[w][lane][1] = a1;
    sh[w][lane][2] = a2;
    sh[w][lane][3] = a3;
    __syncthreads();
    // feature f = fq*4 + t: add the row slots and the warps in a fixed order
    for (int f = threadIdx.x; f < d; f += blockDim.x) {
        const int q = f >> 2, t = f & 3;
        double s = 0.0;
        for (int ww = 0; ww < 8; ++ww)
            for (int rs = 0; rs < rps; ++rs) s += sh[ww][rs * lpr + q][t];
        partial[((size_t)c * SPLIT + sp) * d + f] = s;
    }
}
// partials[c][0..d) = sum of the SPLIT partial sums, partials[c][d] = member count
__global__ void __launch_bounds__(128) finish_sums_kernel(const double* __restrict__ partial, const int* __restrict__ offsets,
                                                          int k, int d, double* __restrict__ out, const int32_t* state) {
    if (state != nullptr && state[0] != 0) return;
    const int c = blockIdx.x;
    for (int f = threadIdx.x; f <= d; f += blockDim.x) {
        double s = 0.0;
        if (f < d) {
            for (int sp = 0; sp < SPLIT; ++sp) s += partial[((size_t)c * SPLIT + sp) * d + f];
        } else {
            s = (double)(offsets[c + 1] - offsets[c]);
        }
        out[(size_t)c * (d + 1) + f] = s;
    }
}

struct BigkScratch {
    float* dist;
    float* cn;
    float* cmax2;
    int32_t* lab;
    int32_t* queue;
    int32_t* perm;
    int* qcount;
    int* hist;
    int* offsets;
    int* cursor;
    double* partial;
    double* fvb;
};

inline size_t al(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

bool bigk_supported(const Handle* h, const LloydArgs& a) {
    if (a.dtype != HK_F32) return false;
    if (a.k < 128 || a.k % 4 != 0 || a.d % 32 != 0 || a.d < 32 || a.d > 128) return false;
    if (a.ldx % 4 != 0 || (reinterpret_cast<uintptr_t>(a.X) & 15) || (reinterpret_cast<uintptr_t>(a.C) & 15)) return false;
    if (a.n < 1024 || a.n >= ((int64_t)1 << 31) - 256) return false;
    (void)h;
    return true;
}

int launch_lloyd_bigk(Handle* h, const LloydArgs& a) {
    const int k = a.k, d = a.d;
    const int64_t n = a.n;
    cudaStream_t st = a.stream;
    // rows per launch of the distance kernel: bounded by its xl scratch (chunk * d floats), a multiple of 128 rows
    int64_t chunk = (int64_t)1 << 20;
    if (chunk > n) chunk = (n + 127) / 128 * 128;
    const int64_t tail_rows = 1024;  // a last chunk below the tensor-core kernel's minimum goes through the exact kernel
    size_t off = 0;
    auto take = [&](size_t bytes) {
        const size_t o = off;
        off += al(bytes);
        return o;
    };
    const size_t o_dist = take((size_t)tail_rows * k * 4), o_cn = take((size_t)k * 4), o_cmax = take(4),
                 o_lab = take((size_t)n * 4), o_queue = take((size_t)n * 4), o_perm = take((size_t)n * 4),
                 o_qc = take(4), o_hist = take((size_t)k * 4), o_offs = take((size_t)(k + 1) * 4),
                 o_cur = take((size_t)k * 4), o_part = take((size_t)k * SPLIT * d * 8),
                 o_fvb = take((size_t)((n + 255) / 256 + 1) * 8);
    if (h->tc_scratch_bytes < off) {
        if (h->tc_scratch) HK_CUDA(cudaFree(h->tc_scratch));
        h->tc_scratch = nullptr;
        h->tc_scratch_bytes = 0;
        HK_CUDA(cudaMalloc(&h->tc_scratch, off));
        h->tc_scratch_bytes = off;
    }
    unsigned char* base = reinterpret_cast<unsigned char*>(h->tc_scratch);
    BigkScratch s;
    s.dist = reinterpret_cast<float*>(base + o_dist);
    s.cn = reinterpret_cast<float*>(base + o_cn);
    s.cmax2 = reinterpret_cast<float*>(base + o_cmax);
    s.lab = reinterpret_cast<int32_t*>(base + o_lab);
    s.queue = reinterpret_cast<int32_t*>(base + o_queue);
    s.perm = reinterpret_cast<int32_t*>(base + o_perm);
    s.qcount = reinterpret_cast<int*>(base + o_qc);
    s.hist = reinterpret_cast<int*>(base + o_hist);
    s.offsets = reinterpret_cast<int*>(base + o_offs);
    s.cursor = reinterpret_cast<int*>(base + o_cur);
    s.partial = reinterpret_cast<double*>(base + o_part);
    s.fvb = reinterpret_cast<double*>(base + o_fvb);

    const float* X = reinterpret_cast<const float*>(a.X);
    const float* C = reinterpret_cast<const float*>(a.C);
    if (!cdist_tc_supported(h, X, chunk < n ? chunk : n, d, a.ldx, C, k, d, s.lab, 4)) {
        set_error("lloyd_bigk: shape not supported by the distance kernel (n=%lld d=%d k=%d)", (long long)n, d, k);
        return -2;
    }
    HK_CUDA(cudaMemsetAsync(s.cmax2, 0, 4, st));
    HK_CUDA(cudaMemsetAsync(s.qcount, 0, 4, st));
    centroid_norms_kernel<<<(k + 127) / 128, 128, 0, st>>>(C, k, d, s.cn, s.cmax2);
    h->launches++;
    // gap below which a label is re-evaluated exactly: (d+3) roundings of the fp32 accumulation on both sides plus the
    // dropped xl*yl term of the 3xTF32 product, on the scale |x|^2 + max|c|^2 >= 2|x||c|
    const float window = 4.f * (float)(d + 3) * 1.1920929e-7f;
    for (int64_t r0 = 0; r0 < n; r0 += chunk) {
        const int64_t rows = n - r0 < chunk ? n - r0 : chunk;
        int rc;
        if (rows >= tail_rows) {
            // distances + argmin fused: the [rows x k] tile never leaves the SM
            rc = launch_cdist_tc_argmin(h, X + (size_t)r0 * a.ldx, rows, d, a.ldx, C, k, d, s.lab, r0, s.queue, s.qcount,
                                        s.cmax2, window, a.state, st);
            if (rc) return rc;
        } else {
            rc = launch_cdist(h, X + (size_t)r0 * a.ldx, rows, d, a.ldx, C, k, d, s.dist, k, HK_F32, 1, 0, st);
            if (rc) return rc;
            const int nb = (int)((rows + 7) / 8);
            argmin_rows_kernel<<<nb, 256, 0, st>>>(s.dist, (int)rows, k, k, r0, n, nullptr, s.cmax2, window, 1, s.lab, s.queue,
                                                   s.qcount, nullptr, a.state);
            h->launches++;
        }
    }
    exact_fix_kernel<<<h->num_sms * 4, 256, 0, st>>>(X, d, a.ldx, C, s.cn, k, s.queue, s.qcount, s.lab, a.state);
    h->launches++;
    if (a.fv_out) {
        const int nb = (int)((n + 255) / 256);
        exact_fv_kernel<<<nb, 256, 0, st>>>(X, n, d, a.ldx, C, s.cn, s.lab, s.fvb, a.state);
        sum_blocks_kernel<<<1, 32, 0, st>>>(s.fvb, nb, a.fv_out, 0, a.state);
        h->launches += 2;
    }
    if (a.labels != nullptr && a.label_kind != HK_LABEL_NONE) {
        store_labels_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s.lab, n, a.labels, a.label_kind, a.state);
        h->launches++;
    }
    if (a.partials != nullptr) {
        HK_CUDA(cudaMemsetAsync(s.hist, 0, (size_t)k * 4, st));
        const int hb = h->num_sms * 4;
        label_hist_kernel<<<hb, 256, (size_t)k * 4, st>>>(s.lab, n, k, s.hist, a.state);
        label_scan_kernel<<<1, 1024, 0, st>>>(s.hist, k, s.offsets, s.cursor, a.state);
        label_scatter_kernel<<<hb, 256, 0, st>>>(s.lab, n, s.offsets, s.cursor, s.perm, a.state);
        gather_reduce_kernel<<<k * SPLIT, 256, 0, st>>>(X, d, a.ldx, s.perm, s.offsets, k, s.partial, a.state);
        finish_sums_kernel<<<k, 128, 0, st>>>(s.partial, s.offsets, k, d, a.partials, a.state);
        h->launches += 5;
    }
    HK_CUDA(cudaGetLastError());
    char name[96];
    snprintf(name, sizeof(name), "bigk<f32,d=%d,k=%d,chunk=%lld,%s>", d, k, (long long)chunk,
             a.partials ? "sums" : "assign");
    h->variant = name;
    return 0;
}

}  // namespace hk
