// Centroid finalize: partial sums/counts (already reduced over CTAs and ranks) -> new centroids,
// squared centroid shift and the sticky convergence flag — all on device, no host sync.
// Replaces the tail of KMeans._update_centroids and the shift/tol test of KMeans.fit
// (heat/cluster/kmeans.py:94-101, 141-144) including the reference's quirks:
//   Q1 mean is evaluated in fp64 and rounded to the centroid dtype on assignment,
//   Q2 an empty cluster moves to the origin (count clipped to 1, masked sum is 0),
//   Q3 the clipped count passes through float32 (heat/core/rounding.py:156-164).
#include "hk_common.cuh"

namespace hk {
namespace {

template <typename T>
__global__ void __launch_bounds__(256) finalize_kernel(const double* __restrict__ part, const T* C_in,
                                                       T* C_out, T* C_prev, int k, int d, int use_tol,
                                                       double tol_cmp, T* shift2_out, int32_t* state) {
    __shared__ T red[256];
    if (state != nullptr && state[0] != 0) return;
    const int tid = threadIdx.x;
    T local = T(0);
    const int n = k * d;
    for (int i = tid; i < n; i += 256) {
        const int c = i / d, f = i - c * d;
        double cnt = part[(size_t)c * (d + 1) + d];
        if (cnt < 1.0) cnt = 1.0;
        const double div = (double)(float)cnt;  // Q3
        const T nv = (T)(part[(size_t)c * (d + 1) + f] / div);
        const T old = C_in[i];
        const T df = old - nv;
        local += df * df;
        if (C_prev != nullptr) C_prev[i] = old;
        C_out[i] = nv;
    }
    red[tid] = local;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) red[tid] += red[tid + o];
        __syncthreads();
    }
    if (tid == 0) {
        const T s = red[0];
        if (shift2_out != nullptr) *shift2_out = s;
        if (state != nullptr) {
            state[1] += 1;
            // `inertia <= tol` with tol rounded to float32 first (heat/core/_operations.py:117-122)
            if (use_tol && s <= (T)tol_cmp) state[0] = 1;
        }
    }
}

}  // namespace

int launch_finalize(Handle* h, const double* partials, const void* C_in, void* C_out, void* C_prev,
                    int k, int d, int dtype, int use_tol, double tol_cmp, void* shift2_out,
                    int32_t* state, cudaStream_t stream) {
    if (dtype == HK_F64)
        finalize_kernel<double><<<1, 256, 0, stream>>>(partials, (const double*)C_in, (double*)C_out,
                                                       (double*)C_prev, k, d, use_tol, tol_cmp,
                                                       (double*)shift2_out, state);
    else
        finalize_kernel<float><<<1, 256, 0, stream>>>(partials, (const float*)C_in, (float*)C_out,
                                                      (float*)C_prev, k, d, use_tol, tol_cmp,
                                                      (float*)shift2_out, state);
    HK_CUDA(cudaGetLastError());
    h->launches++;
    return 0;
}

}  // namespace hk
