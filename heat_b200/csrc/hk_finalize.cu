// Centroid finalize and the fused "finish" of a Lloyd step — all on device, no host sync.
//
// finalize_kernel:  partial sums/counts (already reduced over CTAs and ranks) -> new centroids, squared centroid
//   shift and the sticky convergence flag.  Replaces the tail of KMeans._update_centroids and the shift/tol test of
//   KMeans.fit (heat/cluster/kmeans.py:94-101, 141-144) including the reference's quirks:
//     Q1 mean is evaluated in fp64 and rounded to the centroid dtype on assignment,
//     Q2 an empty cluster moves to the origin (count clipped to 1, masked sum is 0),
//     Q3 the clipped count passes through float32 (heat/core/rounding.py:156-164).
//
// lloyd_finish_kernel:  ONE launch for everything that follows the pass over X:
//     1. fixed-order reduction of the per-CTA fp64 slots of the pass kernel (32 outputs per CTA),
//     2. cross-GPU sum of the k x (d+1) partials WITHOUT a collective library call: every CTA stores its 32
//        outputs straight into a mailbox slot of every peer GPU (peer-mapped memory over NVLink/NVSwitch,
//        double-buffered by exchange parity); the last CTA to finish (atomic ticket) publishes one flag per peer
//        (release at system scope), waits for the peers' flags on its own memory and adds the R copies in RANK
//        ORDER, so every rank obtains bit-identical sums and takes the same convergence decision,
//     3. the finalize arithmetic above, by the same CTA.
//   Replaces the 2k MPICommunication.Allreduce calls per iteration issued by heat/core/_operations.py:505-510
//   (heat/core/communication.py:1089-1110, host-staged for CUDA tensors) and the reduce -> ncclAllReduce -> finalize
//   launch sequence of the first version of this library.
#include "hk_common.cuh"

namespace hk {
namespace {

constexpr int FIN_THREADS = 1024;

// new centroids + shift^2 + convergence flag from fully reduced partials, by one CTA of FIN_THREADS threads.
// `load(i)` returns the reduced partial i (sum over ranks in rank order when there are several).
template <typename T, typename Load>
__device__ __forceinline__ void finalize_block(Load load, const T* C_in, T* C_out, T* C_prev, int k, int d, int use_tol,
                                               double tol_cmp, T* shift2_out, int32_t* state, T* red /* [FIN_THREADS] */) {
    const int tid = threadIdx.x;
    T local = T(0);
    const int n = k * d;
    for (int i = tid; i < n; i += FIN_THREADS) {
        const int c = i / d, f = i - c * d;
        double cnt = load(c * (d + 1) + d);
        if (cnt < 1.0) cnt = 1.0;
        const double div = (double)(float)cnt;  // Q3
        const T nv = (T)(load(c * (d + 1) + f) / div);
        const T old = C_in[i];
        const T df = old - nv;
        local += df * df;
        if (C_prev != nullptr) C_prev[i] = old;
        C_out[i] = nv;
    }
    red[tid] = local;
    __syncthreads();
    for (int o = FIN_THREADS / 2; o > 0; o >>= 1) {
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

template <typename T>
__global__ void __launch_bounds__(FIN_THREADS) finalize_kernel(const double* __restrict__ part, const T* C_in, T* C_out,
                                                               T* C_prev, int k, int d, int use_tol, double tol_cmp,
                                                               T* shift2_out, int32_t* state) {
    __shared__ T red[FIN_THREADS];
    if (state != nullptr && state[0] != 0) return;
    finalize_block<T>([&](int i) { return part[i]; }, C_in, C_out, C_prev, k, d, use_tol, tol_cmp, shift2_out, state, red);
}

__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <typename T>
__global__ void __launch_bounds__(FIN_THREADS) lloyd_finish_kernel(const FinishParams p) {
    __shared__ double sh[32][33];
    __shared__ T redT[FIN_THREADS];
    __shared__ unsigned s_last;
    if (p.state != nullptr && p.state[0] != 0) return;  // uniform across the grid (and across ranks)
    const int tid = threadIdx.x;
    const int lane = tid & 31, grp = tid >> 5;
    const int i = blockIdx.x * 32 + lane;
    const int len = p.k * (p.d + 1);
    const int R = p.nranks;
    // number of exchanges executed so far: identical on every rank (all ranks execute the same steps); only the
    // finishing CTA bumps it, after every other CTA of this launch has taken its ticket
    const uint32_t ep = *reinterpret_cast<volatile uint32_t*>(p.epoch);
    const uint32_t par = ep & 1u;

    // ---- 1. this rank's partials: fixed-order reduction over the pass kernel's slots ------------------------
    double t = 0.0;
    if (i < len) {
        if (p.fsum != nullptr) {
            const int c = i / (p.d + 1), f = i - c * (p.d + 1);
            if (f < p.d) {
                const int per = (p.nslots + 31) / 32;
                const int b1 = min(p.nslots, (grp + 1) * per);
                for (int b = grp * per; b < b1; ++b)
                    t += p.fsum[(size_t)b * p.slot_stride * p.k * p.d + (size_t)c * p.d + f];
            } else {
                const int per = (p.nblocks + 31) / 32;
                const int b1 = min(p.nblocks, (grp + 1) * per);
                for (int b = grp * per; b < b1; ++b) t += p.fcnt[(size_t)b * p.k + c];
            }
        } else if (grp == 0) {
            t = p.partials_in[i];
        }
    }
    sh[grp][lane] = t;
    __syncthreads();
    if (grp == 0 && i < len) {
        double r = sh[0][lane];
#pragma unroll
        for (int g2 = 1; g2 < 32; ++g2) r += sh[g2][lane];
        p.red[i] = r;
        // ---- 2a. push to every rank's mailbox (own included), slot [parity][this rank] ----------------------
        if (R > 1) {
            const size_t off = ((size_t)par * R + p.rank) * p.cap + i;
#pragma unroll 1
            for (int rr = 0; rr < R; ++rr) {
                const int dst = (p.rank + 1 + rr) % R;  // spread the first stores over the peers
                p.mbox[dst][off] = r;
            }
        }
    }
    if (R > 1)
        __threadfence_system();
    else
        __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned tk = atomicAdd(p.ticket, 1u);
        s_last = (tk == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();  // the other CTAs' stores (made before their tickets) are visible from here on

    // ---- 2b. the finishing CTA: publish, wait, then everything below sees all R copies -----------------------
    const double* mine = p.red;
    if (R > 1) {
        if (tid < R) {
            __threadfence_system();
            st_release_sys_u32(p.flags[tid] + par * R + p.rank, ep + 1u);
            const uint32_t* fl = p.flags[p.rank] + par * R + tid;
            const unsigned long long t0 = globaltimer_ns();
            while ((int32_t)(ld_acquire_sys_u32(fl) - (ep + 1u)) < 0) {
                if (globaltimer_ns() - t0 > 20000000000ull) {  // 20 s: a peer died; fail loudly instead of hanging
                    printf("hk lloyd_finish_kernel: rank %d timed out waiting for rank %d (exchange %u)\n", p.rank, tid,
                           ep + 1u);
                    __trap();
                }
            }
        }
        __syncthreads();
        __threadfence_system();
        mine = p.mbox[p.rank] + (size_t)par * R * p.cap;
    }
    const size_t cap = p.cap;
    auto load = [&](int idx) -> double {
        if (R == 1) return mine[idx];
        double s = __ldcg(mine + idx);  // L2 only: the peers wrote these lines
        for (int rr = 1; rr < R; ++rr) s += __ldcg(mine + (size_t)rr * cap + idx);
        return s;
    };
    finalize_block<T>(load, (const T*)p.C_in, (T*)p.C_out, (T*)p.C_prev, p.k, p.d, p.use_tol, p.tol_cmp, (T*)p.shift2_out,
                      p.state, redT);
    if (p.partials_out != nullptr) {  // the globally reduced partials, for callers that want them
        for (int idx = tid; idx < len; idx += FIN_THREADS) p.partials_out[idx] = load(idx);
    }
    if (tid == 0) {
        *p.ticket = 0u;
        *p.epoch = ep + 1u;
    }
}

}  // namespace

int launch_finalize(Handle* h, const double* partials, const void* C_in, void* C_out, void* C_prev,
                    int k, int d, int dtype, int use_tol, double tol_cmp, void* shift2_out,
                    int32_t* state, cudaStream_t stream) {
    if (dtype == HK_F64)
        finalize_kernel<double><<<1, FIN_THREADS, 0, stream>>>(partials, (const double*)C_in, (double*)C_out,
                                                               (double*)C_prev, k, d, use_tol, tol_cmp,
                                                               (double*)shift2_out, state);
    else
        finalize_kernel<float><<<1, FIN_THREADS, 0, stream>>>(partials, (const float*)C_in, (float*)C_out,
                                                              (float*)C_prev, k, d, use_tol, tol_cmp,
                                                              (float*)shift2_out, state);
    HK_CUDA(cudaGetLastError());
    h->launches++;
    return 0;
}

int launch_finish(Handle* h, FinishParams& p, int dtype, cudaStream_t stream) {
    const int len = p.k * (p.d + 1);
    const int grid = (len + 31) / 32;
    if (dtype == HK_F64)
        lloyd_finish_kernel<double><<<grid, FIN_THREADS, 0, stream>>>(p);
    else
        lloyd_finish_kernel<float><<<grid, FIN_THREADS, 0, stream>>>(p);
    HK_CUDA(cudaGetLastError());
    h->launches++;
    return 0;
}

}  // namespace hk
