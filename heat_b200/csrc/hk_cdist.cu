// Pairwise distances  out[i,j] = dist(X[i,:], Y[j,:])  for one row shard of X against a replicated Y.
// Replaces cdist -> _dist -> metric(X.larray, Y.larray) for X.split in {0, None}, Y.split None
// (heat/spatial/distance.py:409-414) with metric = _euclidian_fast (quadratic expansion,
// distance.py:32-64) or _euclidian (direct, distance.py:17-29).
// Exact-FMA register-tiled kernel (fp32 / fp64); the fp32 quadratic-expansion case is routed to the
// tensor-core kernel in hk_cdist_tc.cu when the shape allows it.
#include "hk_common.cuh"

namespace hk {
namespace {

constexpr int BM = 64, BN = 64, BK = 16;

template <typename T>
__global__ void row_norms_kernel(const T* __restrict__ A, int64_t rows, int f, int64_t lda,
                                 T* __restrict__ out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const T* a = A + r * lda;
    T s = T(0);
    for (int i = 0; i < f; ++i) s = fma(a[i], a[i], s);
    out[r] = s;
}

// BODY: 0 = sum (x-y)^2, 1 = x.y for the quadratic expansion, 2 = sum |x-y| (heat/spatial/distance.py:120-133)
// post: 0 = as accumulated, 1 = sqrt, 2 = exp(-v / gden) with gden = 2 sigma^2 (distance.py:67-101)
template <typename T, int BODY>
__global__ void __launch_bounds__(256) cdist_simt_kernel(const T* __restrict__ X, int64_t m, int f,
                                                         int64_t ldx, const T* __restrict__ Y, int64_t n,
                                                         int64_t ldy, T* __restrict__ out, int64_t ldo,
                                                         const T* __restrict__ xn, const T* __restrict__ yn,
                                                         int post, T gden) {
    __shared__ T xs[BK][BM + 4];
    __shared__ T ys[BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t i0 = (int64_t)blockIdx.y * BM, j0 = (int64_t)blockIdx.x * BN;
    T acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = T(0);

    for (int k0 = 0; k0 < f; k0 += BK) {
        // 64 x 16 elements per operand, 256 threads -> 4 each; consecutive threads walk the feature axis
        for (int e = tid; e < BM * BK; e += 256) {
            const int r = e / BK, c = e - r * BK;
            const int64_t gi = i0 + r;
            const int gk = k0 + c;
            xs[c][r] = (gi < m && gk < f) ? X[gi * ldx + gk] : T(0);
            const int64_t gj = j0 + r;
            ys[c][r] = (gj < n && gk < f) ? Y[gj * ldy + gk] : T(0);
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            T xv[4], yv[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) xv[a] = xs[kk][ty * 4 + a];
#pragma unroll
            for (int b = 0; b < 4; ++b) yv[b] = ys[kk][tx * 4 + b];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    if (BODY == 1) {
                        acc[a][b] = fma(xv[a], yv[b], acc[a][b]);
                    } else if (BODY == 2) {
                        acc[a][b] += fabs(xv[a] - yv[b]);
                    } else {
                        const T df = xv[a] - yv[b];
                        acc[a][b] = fma(df, df, acc[a][b]);
                    }
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int64_t gi = i0 + ty * 4 + a;
        if (gi >= m) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int64_t gj = j0 + tx * 4 + b;
            if (gj >= n) continue;
            T v = acc[a][b];
            if (BODY == 1) {
                v = (xn[gi] + yn[gj]) - T(2) * v;
                v = v < T(0) ? T(0) : v;  // clamp(0, inf); NaN propagates
            }
            if (post == 1)
                v = sqrt(v);
            else if (post == 2)
                v = exp(-v / gden);
            out[gi * ldo + gj] = v;
        }
    }
}

template <typename T>
int run_cdist(Handle* h, const T* X, int64_t m, int f, int64_t ldx, const T* Y, int64_t n, int64_t ldy,
              T* out, int64_t ldo, int body, int post, double gden, cudaStream_t st) {
    const int quad = body == 1;
    T* xn = nullptr;
    T* yn = nullptr;
    if (quad) {
        int rc = ensure_part(h, (size_t)(m + n) * sizeof(T) + 256);
        if (rc) return rc;
        xn = reinterpret_cast<T*>(h->part);
        yn = xn + m;
        row_norms_kernel<T><<<(unsigned)((m + 255) / 256), 256, 0, st>>>(X, m, f, ldx, xn);
        row_norms_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Y, n, f, ldy, yn);
        HK_CUDA(cudaGetLastError());
        h->launches += 2;
    }
    const int64_t gy_total = (m + BM - 1) / BM;
    const unsigned gx = (unsigned)((n + BN - 1) / BN);
    // gridDim.y is limited to 65535: walk X in slabs
    for (int64_t y0 = 0; y0 < gy_total; y0 += 65535) {
        const unsigned gy = (unsigned)((gy_total - y0) < 65535 ? (gy_total - y0) : 65535);
        const int64_t r0 = y0 * BM;
        dim3 grid(gx, gy);
        prof_begin(h, st);
        if (quad)
            cdist_simt_kernel<T, 1><<<grid, 256, 0, st>>>(X + r0 * ldx, m - r0, f, ldx, Y, n, ldy, out + r0 * ldo, ldo,
                                                          xn + r0, yn, post, (T)gden);
        else if (body == 2)
            cdist_simt_kernel<T, 2><<<grid, 256, 0, st>>>(X + r0 * ldx, m - r0, f, ldx, Y, n, ldy, out + r0 * ldo, ldo,
                                                          nullptr, nullptr, post, (T)gden);
        else
            cdist_simt_kernel<T, 0><<<grid, 256, 0, st>>>(X + r0 * ldx, m - r0, f, ldx, Y, n, ldy, out + r0 * ldo, ldo,
                                                          nullptr, nullptr, post, (T)gden);
        prof_end(h, st);
        HK_CUDA(cudaGetLastError());
        h->launches++;
    }
    return 0;
}

}  // namespace

int launch_cdist_tc(Handle* h, const void* X, int64_t m, int f, int64_t ldx, const void* Y, int64_t n,
                    int64_t ldy, void* out, int64_t ldo, int post, float gscale, cudaStream_t stream);
bool cdist_tc_supported(const Handle* h, const void* X, int64_t m, int f, int64_t ldx, const void* Y,
                        int64_t n, int64_t ldy, const void* out, int64_t ldo);

int launch_cdist(Handle* h, const void* X, int64_t m, int f, int64_t ldx, const void* Y, int64_t n,
                 int64_t ldy, void* out, int64_t ldo, int dtype, int body, int post, cudaStream_t stream,
                 double gden) {
    if (dtype == HK_F32 && body == 1 && cdist_tc_supported(h, X, m, f, ldx, Y, n, ldy, out, ldo)) {
        h->variant = "cdist_tc<f32>";
        // exp(-v / gden) = 2^(v * gscale)
        const float gscale = (float)(-1.4426950408889634 / gden);
        return launch_cdist_tc(h, X, m, f, ldx, Y, n, ldy, out, ldo, post, gscale, stream);
    }
    h->variant = dtype == HK_F64 ? "cdist_simt<f64>" : "cdist_simt<f32>";
    if (dtype == HK_F64)
        return run_cdist<double>(h, (const double*)X, m, f, ldx, (const double*)Y, n, ldy, (double*)out, ldo,
                                 body, post, gden, stream);
    return run_cdist<float>(h, (const float*)X, m, f, ldx, (const float*)Y, n, ldy, (float*)out, ldo, body, post,
                            gden, stream);
}

}  // namespace hk
