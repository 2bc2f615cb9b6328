// fp64 Lloyd pass for the small-k, 128-byte-row regime (BASELINE config 5: N=50M, d=16, k=8, fp64): HBM-bound.
//
// The first version of this path (hk_lloyd_row128.cu: thread == row, centroids streamed from shared memory, private
// shared-memory accumulators) ran at 49 % of the HBM roofline because it was bound by shared-memory bandwidth
// (ncu: LSU wavefronts 80 % of peak): ~800 wavefronts per 128-row tile against the 728 cycles the tile's 16 KB take at
// HBM speed.  Here nothing but the row tile itself lives in shared memory:
//   * distances:  x.c^T on the FP64 tensor cores (mma.sync.m8n8k4.f64, 8 rows x 8 centroids x 4 features per
//                 instruction); the centroids are four B-fragment registers per lane for the whole kernel,
//                 d2 = fl(fl(|x|^2 + |c|^2) - 2 x.c) as heat/spatial/distance.py:59-64, first-index argmin with
//                 torch.min NaN semantics (heat/core/statistics.py:177) inside each quad of lanes;
//   * cluster sums: onehot(labels)^T . X on the same tensor cores (products by 0/1 are exact, the accumulation is
//                 fp64 with round-to-nearest per step), accumulators = four C-fragment registers per lane for the
//                 whole kernel: no shared-memory read-modify-write, no label hand-off between warps, no flushes;
//   * counts:     integer adds of the one-hot fragments.
// Each row is read from shared memory twice (once per fragment layout) = 256 B per row, 256 wavefronts per tile.
// Warp roles: warp 0 = TMA producer (128-byte swizzle, EVICT_FIRST) into an S-stage ring; 16 compute warps, warp (q, r)
// owns rows [32q, 32q+32) of the tiles i == r (mod 4).  Per-warp results go to per-warp fp64 slots in global memory and
// are reduced in a fixed order (deterministic).
// Replaces _assign_to_cluster + KMeans._update_centroids for one shard
// (heat/cluster/_kcluster.py:352-370, heat/cluster/kmeans.py:76-103).
#include <math.h>

#include "hk_tma.cuh"

namespace hk {
namespace {

constexpr int TM = 128;
#ifndef HK_DMMA_ER
#define HK_DMMA_ER 7  // 28 compute warps: the kernel is latency bound (ER 4: 59 %, 6: 67 %, 7: 68 % of HBM)
#endif
constexpr int ER = HK_DMMA_ER;  // tile residues (4 warps each)
constexpr int C_WARPS = 4 * ER;
constexpr int STAGE_BYTES = TM * 128;
constexpr int D = 16;

struct DmmaParams {
    int64_t n;
    int k;
    const double* C;
    void* labels;
    int label_kind;
    double* fsum;     // [grid][k*16] or nullptr (assign only)
    double* fcnt;     // [grid][k]
    double* fv_part;  // [grid] or nullptr
    int S;
    int num_tiles;
    const int32_t* state;
};

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void store_label_d(void* labels, int kind, int64_t row, int lab) {
    if (kind == HK_LABEL_I64)
        reinterpret_cast<long long*>(labels)[row] = lab;
    else if (kind == HK_LABEL_I32)
        reinterpret_cast<int*>(labels)[row] = lab;
    else if (kind == HK_LABEL_U8)
        reinterpret_cast<unsigned char*>(labels)[row] = (unsigned char)lab;
}
// sequential torch.min semantics when merging (value, index) pairs that cover disjoint index sets: the first NaN wins,
// otherwise the smallest value, ties to the smaller index
__device__ __forceinline__ void merge_min(double& v, int& j, double ov, int oj) {
    const bool on = ov != ov, bn = v != v;
    const bool take = (on && bn) ? (oj < j) : (on ? true : (bn ? false : (ov < v || (ov == v && oj < j))));
    if (take) {
        v = ov;
        j = oj;
    }
}

// NB = number of 8-centroid blocks (k <= 8 * NB); XN: the functional value is wanted, so the distances carry |x|^2
template <int NB, bool SUMS, bool XN>
__global__ void __launch_bounds__((1 + C_WARPS) * 32, 1)
    lloyd_dmma_kernel(const __grid_constant__ CUtensorMap xmap, const DmmaParams p) {
    extern __shared__ unsigned char smem_raw[];
    if (p.state != nullptr && p.state[0] != 0) return;  // uniform across the grid
    unsigned char* smem =
        reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const int S = p.S, k = p.k;
    const uint32_t a_stages = sbase;
    const uint32_t o_bars = (uint32_t)S * STAGE_BYTES;
    const uint32_t b_full = sbase + o_bars;  // full[S] | empty[S]  (16 slots each)
    const uint32_t b_empty = b_full + 16 * 8;
    double* fvred = reinterpret_cast<double*>(smem + o_bars + 32 * 8);  // [C_WARPS]

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int ntiles = p.num_tiles;

    if (tid == 0) {
        uint64_t* bars = reinterpret_cast<uint64_t*>(smem + o_bars);
        for (int s = 0; s < S; ++s) {
            mbar_init(bars + s, 1);
            mbar_init(bars + 16 + s, 4);
        }
        mbar_fence_init();
        tma_prefetch_desc(&xmap);
    }
    __syncthreads();

    // per-warp results live in registers across the role code (declared here so that every thread of the block meets
    // the same block-wide barriers after it)
    double acc0[NB][2], acc1[NB][2];
    int cnt[NB];
#pragma unroll
    for (int mb = 0; mb < NB; ++mb) {
        acc0[mb][0] = acc0[mb][1] = acc1[mb][0] = acc1[mb][1] = 0.0;
        cnt[mb] = 0;
    }
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
    } else {
        // ================= compute warps =================
        const int we = warp - 1;
        const int q = we & 3;
        const int r = we >> 2;
        const int g = lane >> 2;  // row of an 8-row group / centroid of a B fragment / cluster of an accumulator row
        const int t = lane & 3;
        // centroid fragments of the distance MMA: B[f][j] = c[j][f].  The order of the features inside the contraction is
        // free, so K-block kb takes feature 4 t + kb from lane t: a lane's four A values are 32 contiguous bytes of its
        // row (two 16-byte loads) and this lane holds c[j = 8 nb + g][4 t + kb]
        double bc[NB][4];
        double cn0[NB], cn1[NB];  // |c_j|^2 of this lane's two D columns j = 8 nb + 2 t + {0, 1}
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
            const int j = nb * 8 + g;
#pragma unroll
            for (int kb = 0; kb < 4; ++kb) bc[nb][kb] = j < k ? p.C[(size_t)j * D + 4 * t + kb] : 0.0;
            const int j0 = nb * 8 + 2 * t;
            double s0 = INFINITY, s1 = INFINITY;  // padded centroids can never win
            if (j0 < k) {
                s0 = 0.0;
                for (int f = 0; f < D; ++f) s0 = fma(p.C[(size_t)j0 * D + f], p.C[(size_t)j0 * D + f], s0);
            }
            if (j0 + 1 < k) {
                s1 = 0.0;
                for (int f = 0; f < D; ++f) s1 = fma(p.C[(size_t)(j0 + 1) * D + f], p.C[(size_t)(j0 + 1) * D + f], s1);
            }
            cn0[nb] = s0;
            cn1[nb] = s1;
        }
        bool cfin = true;
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
            // padded centroids carry +inf norms on purpose; a real centroid with NaN/Inf disables the integer argmin
            if (nb * 8 + 2 * t < k) cfin = cfin && (cn0[nb] < INFINITY);
            if (nb * 8 + 2 * t + 1 < k) cfin = cfin && (cn1[nb] < INFINITY);
        }
        const bool cfinite = __all_sync(0xffffffffu, cfin);
        // accumulators of onehot^T . X: accN[mb] = cluster 8 mb + g, D columns 2 t + {0, 1} of block nbf = N, i.e.
        // features 4 t + N and 4 t + 2 + N
        double fv_acc = 0.0;
        const bool want_fv = p.fv_part != nullptr;
        const int label_kind = p.label_kind;

        int s = r % S;
        uint32_t ph = (uint32_t)((r / S) & 1);
        for (int tile = blockIdx.x + r * gridDim.x; tile < ntiles; tile += ER * gridDim.x) {
            const uint32_t xt = a_stages + s * STAGE_BYTES + (uint32_t)(q * 32 * 128);
            const int64_t row0 = (int64_t)tile * TM + q * 32;
            warp_wait(b_full + s * 8, ph, lane);
#pragma unroll
            for (int grp = 0; grp < 4; ++grp) {
                // ---- distances of rows 8 grp .. 8 grp + 7: A[row g][K-block kb] = x[row][4 t + kb] ----------------------
                const int ra = grp * 8 + g;
                const uint32_t xa = xt + (uint32_t)ra * 128;
                double a[4];
                {
                    const double2 lo = lds_d2(xa + (uint32_t)(((2 * t) ^ (ra & 7)) << 4));      // features 4t, 4t+1
                    const double2 hi = lds_d2(xa + (uint32_t)(((2 * t + 1) ^ (ra & 7)) << 4));  // features 4t+2, 4t+3
                    a[0] = lo.x;
                    a[1] = lo.y;
                    a[2] = hi.x;
                    a[3] = hi.y;
                }
                // |x|^2 does not change the argmin: it is evaluated only when the distance itself is wanted (functional
                // value) or a row holds NaN/Inf.  Without it the labels are the argmin of fl(|c_j|^2 - 2 x.c_j), which can
                // differ from the argmin of fl(fl(|x|^2 + |c_j|^2) - 2 x.c_j) only between distances that agree to
                // 1 ulp of |x|^2 (~1e-16 relative: far inside the near-tie clause |dd| < 1e-6 d of the parity rule).
                auto row_norm = [&]() {
                    double v = 0.0;
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb) v = fma(a[kb], a[kb], v);
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    return v;
                };
                double xn = 0.0;
                if (XN) xn = row_norm();
                // candidates of this lane (its 2 NB columns): XN: d2 = clamp(fl(fl(|x|^2 + |c_j|^2) - 2 x.c_j), 0) as
                // heat/spatial/distance.py:59-64; else s_j = fl(|c_j|^2 - 2 x.c_j)
                double e[2 * NB], dots[2 * NB];
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) {
                    double d0 = 0.0, d1 = 0.0;
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb) dmma(d0, d1, a[kb], bc[nb][kb]);
                    dots[2 * nb] = d0;
                    dots[2 * nb + 1] = d1;
                    if (XN) {
                        const double e0 = (xn + cn0[nb]) - 2.0 * d0;
                        const double e1 = (xn + cn1[nb]) - 2.0 * d1;
                        e[2 * nb] = e0 < 0.0 ? 0.0 : e0;  // clamp(d2, 0, inf); NaN stays NaN
                        e[2 * nb + 1] = e1 < 0.0 ? 0.0 : e1;
                    } else {
                        e[2 * nb] = fma(-2.0, d0, cn0[nb]);
                        e[2 * nb + 1] = fma(-2.0, d1, cn1[nb]);
                    }
                }
                double best;
                int lab;
                // A row without NaN/Inf against finite centroids (x.c_j finite; a padded centroid has x.0 = 0 there):
                // every candidate is an ordinary double or the +inf of a padded centroid, and its order is the order of
                // the usual sortable integer image of its bit pattern - the argmin runs on the integer pipe (FP64
                // compares are half rate and were 25 % of this kernel's instructions).  Key = (image, index):
                // lexicographic minimum = first-index minimum.
                const bool row_ok =
                    cfinite && ((unsigned)(__double2hiint(dots[0]) & 0x7ff00000) != 0x7ff00000u);
                if (__all_sync(0xffffffffu, row_ok)) {
                    auto image = [](double v) {
                        const long long b = __double_as_longlong(v);
                        return (unsigned long long)(b ^ ((b >> 63) | (long long)0x8000000000000000ull));
                    };
                    unsigned long long kbest = image(e[0]);
                    lab = 2 * t;
#pragma unroll
                    for (int c = 1; c < 2 * NB; ++c) {
                        const unsigned long long kc = image(e[c]);
                        const int jc = (c >> 1) * 8 + 2 * t + (c & 1);
                        if (kc < kbest) {  // ascending index order inside the lane: strict < keeps the first index
                            kbest = kc;
                            lab = jc;
                        }
                    }
                    // (three quad-masked redux.sync instead of these shuffles were measured: 2x slower, the eight
                    // different member masks of a warp are executed one after the other)
#pragma unroll
                    for (int o = 1; o <= 2; o <<= 1) {
                        const unsigned long long ok = __shfl_xor_sync(0xffffffffu, kbest, o);
                        const int oj = __shfl_xor_sync(0xffffffffu, lab, o);
                        if (ok < kbest || (ok == kbest && oj < lab)) {
                            kbest = ok;
                            lab = oj;
                        }
                    }
                    const long long ib = (long long)kbest;
                    best = __longlong_as_double(ib ^ (((~ib) >> 63) | (long long)0x8000000000000000ull));
                } else {
                    // NaN / Inf somewhere in these 8 rows (or in a centroid): the reference formula with |x|^2 and
                    // sequential torch.min semantics in fp64
                    if (!XN) {
                        xn = row_norm();
#pragma unroll
                        for (int nb = 0; nb < NB; ++nb) {
                            const double e0 = (xn + cn0[nb]) - 2.0 * dots[2 * nb];
                            const double e1 = (xn + cn1[nb]) - 2.0 * dots[2 * nb + 1];
                            e[2 * nb] = e0 < 0.0 ? 0.0 : e0;
                            e[2 * nb + 1] = e1 < 0.0 ? 0.0 : e1;
                        }
                    }
                    best = e[0];
                    lab = 2 * t;
#pragma unroll
                    for (int c = 1; c < 2 * NB; ++c) {
                        const int jc = (c >> 1) * 8 + 2 * t + (c & 1);
                        if (e[c] < best || (e[c] != e[c] && best == best)) {
                            best = e[c];
                            lab = jc;
                        }
                    }
                    // the four lanes of a quad hold disjoint centroid subsets of the same row
#pragma unroll
                    for (int o = 1; o <= 2; o <<= 1) {
                        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                        const int ol = __shfl_xor_sync(0xffffffffu, lab, o);
                        merge_min(best, lab, ob, ol);
                    }
                }
                const bool active = row0 + ra < p.n;
                if (active && t == 0) {
                    if (label_kind != HK_LABEL_NONE) store_label_d(p.labels, label_kind, row0 + ra, lab);
                    if (want_fv) {
                        const double sq = sqrt(best);
                        fv_acc += sq * sq;
                    }
                }
                if (SUMS) {
                    // ---- sums += onehot^T . X: A[cluster g][row 4 kb + t], B[row 4 kb + t][column g] = feature 2 g + nbf ----
                    const int mylab = active ? lab : -1;  // rows past the end belong to no cluster
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb) {
                        const int rr = kb * 4 + t;                                            // row inside the group
                        const int rl = __shfl_sync(0xffffffffu, mylab, rr * 4);               // its label
                        const int rb = grp * 8 + rr;
                        const uint32_t xb = xt + (uint32_t)rb * 128;
                        // the order of the feature columns is free too: column g of block nbf is feature 2 g + nbf, so both
                        // B values of a lane come from one 16-byte load
                        const double2 bb = lds_d2(xb + (uint32_t)((g ^ (rb & 7)) << 4));
                        const double b0 = bb.x, b1 = bb.y;
#pragma unroll
                        for (int mb = 0; mb < NB; ++mb) {
                            const bool hit = rl == mb * 8 + g;
                            const double oh = hit ? 1.0 : 0.0;
                            cnt[mb] += hit ? 1 : 0;
                            dmma(acc0[mb][0], acc0[mb][1], oh, b0);
                            dmma(acc1[mb][0], acc1[mb][1], oh, b1);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_a(b_empty + s * 8);
            s += ER;
            while (s >= S) {
                s -= S;
                ph ^= 1;
            }
        }
        if (want_fv) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) fv_acc += __shfl_xor_sync(0xffffffffu, fv_acc, o);
            if (lane == 0) fvred[we] = fv_acc;
        }
    }
    __syncthreads();  // all compute warps and the producer are past their last tile: the row tiles are dead
    if (SUMS) {
        // every compute warp's sums[cluster][feature] and counts go through shared memory and are added per CTA in a
        // fixed order: one slot per CTA for the cross-CTA reduction
        if (warp >= 1) {
            const int we = warp - 1, g = lane >> 2, t = lane & 3;
            double* wsl = reinterpret_cast<double*>(smem) + (size_t)we * (size_t)(k * (D + 1));
#pragma unroll
            for (int mb = 0; mb < NB; ++mb) {
                const int c = mb * 8 + g;
                int ct = cnt[mb];
                ct += __shfl_xor_sync(0xffffffffu, ct, 1);
                ct += __shfl_xor_sync(0xffffffffu, ct, 2);
                if (c < k) {
                    wsl[(size_t)c * D + 4 * t] = acc0[mb][0];
                    wsl[(size_t)c * D + 4 * t + 2] = acc0[mb][1];
                    wsl[(size_t)c * D + 4 * t + 1] = acc1[mb][0];
                    wsl[(size_t)c * D + 4 * t + 3] = acc1[mb][1];
                    if (t == 0) wsl[(size_t)k * D + c] = (double)ct;
                }
            }
        }
        __syncthreads();
        const int per = k * (D + 1);
        const double* all = reinterpret_cast<const double*>(smem);
        for (int i = tid; i < per; i += blockDim.x) {
            double tsum = 0.0;
            for (int w = 0; w < C_WARPS; ++w) tsum += all[(size_t)w * per + i];
            if (i < k * D)
                p.fsum[(size_t)blockIdx.x * (k * D) + i] = tsum;
            else
                p.fcnt[(size_t)blockIdx.x * k + (i - k * D)] = tsum;
        }
    }
    if (tid == 0 && p.fv_part != nullptr) {
        double tsum = 0.0;
        for (int w = 0; w < C_WARPS; ++w) tsum += fvred[w];
        p.fv_part[blockIdx.x] = tsum;
    }
}

// partials[c][0..d) / [d]: sums and counts over all warp slots in a fixed order
__global__ void __launch_bounds__(256) reduce_dmma_kernel(const double* __restrict__ fsum, const double* __restrict__ fcnt,
                                                          int nslots, int k, double* __restrict__ out, const int32_t* state) {
    if (state != nullptr && state[0] != 0) return;
    __shared__ double sh[8][33];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lane;
    const int len = k * (D + 1);
    double t = 0.0;
    if (i < len) {
        const int c = i / (D + 1), f = i - c * (D + 1);
        const int per = (nslots + 7) / 8;
        const int b1 = min(nslots, (grp + 1) * per);
        if (f < D) {
            for (int b = grp * per; b < b1; ++b) t += fsum[(size_t)b * k * D + (size_t)c * D + f];
        } else {
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
__global__ void reduce_scalar_dmma_kernel(const double* __restrict__ v, int n, double* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double t = 0.0;
        for (int b = 0; b < n; ++b) t += v[b];
        *out = t;
    }
}

template <int NB, bool SUMS, bool XN>
int launch_dmma_inst(Handle* h, const CUtensorMap& map, const DmmaParams& p, size_t smem, int grid, cudaStream_t st) {
    auto kern = lloyd_dmma_kernel<NB, SUMS, XN>;
    HK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    prof_begin(h, st);
    kern<<<grid, (1 + C_WARPS) * 32, smem, st>>>(map, p);
    prof_end(h, st);
    HK_CUDA(cudaGetLastError());
    h->launches++;
    return 0;
}

}  // namespace


bool dmma_supported(const Handle* h, const LloydArgs& a) {
    (void)h;
    if (a.dtype != HK_F64 || a.d != D) return false;
    if (a.k < 1 || a.k > 16) return false;
    if ((a.ldx * 8) % 16 != 0) return false;
    if ((reinterpret_cast<uintptr_t>(a.X) & 15) != 0) return false;
    if (a.n >= (int64_t)1 << 31) return false;
    return true;
}

int launch_lloyd_dmma(Handle* h, const LloydArgs& a) {
    const bool sums = a.partials != nullptr;
    // Stages in flight per SM: a MULTIPLE of the number of tile residues, so that every stage is always consumed by the
    // same four warps.  A warp waits on a stage's "full" barrier with a one-bit phase parity; if consecutive uses of a
    // stage belonged to different warps, a warp could reach its wait while the PREVIOUS use's TMA load is still in flight
    // (loads complete out of order), see the parity of the phase before and walk on without data.  Observed as a hang
    // once in a few hundred million tiles with 12 stages and 7 residues.
    const int S = 2 * ER;  // 14 x 16 KB row tiles
    static_assert((2 * ER) % ER == 0 && 2 * ER <= 16, "stage count must be a multiple of the residue count");
    const size_t smem = (size_t)S * STAGE_BYTES + 32 * 8 + C_WARPS * 8 + 1024;
    CUtensorMap map;
    int rc = make_tensor_map_2d(&map, a.X, 8, (uint64_t)a.n, (uint64_t)a.d, (uint64_t)a.ldx, (uint32_t)a.d, TM, 128);
    if (rc) return rc;
    DmmaParams p{};
    p.n = a.n;
    p.k = a.k;
    p.C = reinterpret_cast<const double*>(a.C);
    p.labels = a.labels;
    p.label_kind = a.labels ? a.label_kind : HK_LABEL_NONE;
    p.S = S;
    p.num_tiles = (int)((a.n + TM - 1) / TM);
    p.state = a.state;
    int grid = h->num_sms;
    if (grid > p.num_tiles) grid = p.num_tiles;
    const int nslots = grid;  // one slot per CTA (the warps are folded inside the kernel)
    const size_t kd = (size_t)a.k * D;
    rc = ensure_part(h, ((size_t)nslots * kd + (size_t)nslots * a.k + grid) * sizeof(double));
    if (rc) return rc;
    p.fsum = sums ? h->part : nullptr;
    p.fcnt = h->part + (size_t)nslots * kd;
    p.fv_part = a.fv_out ? h->part + (size_t)nslots * kd + (size_t)nslots * a.k : nullptr;

    char name[96];
    snprintf(name, sizeof(name), "dmma<f64,d=16,k=%d,S=%d,%s>", a.k, S, sums ? "sums" : "assign");
    h->variant = name;
    const bool two = a.k > 8;
    const bool fv = a.fv_out != nullptr;
    if (sums)
        rc = two ? launch_dmma_inst<2, true, false>(h, map, p, smem, grid, a.stream)
                 : launch_dmma_inst<1, true, false>(h, map, p, smem, grid, a.stream);
    else if (fv)
        rc = two ? launch_dmma_inst<2, false, true>(h, map, p, smem, grid, a.stream)
                 : launch_dmma_inst<1, false, true>(h, map, p, smem, grid, a.stream);
    else
        rc = two ? launch_dmma_inst<2, false, false>(h, map, p, smem, grid, a.stream)
                 : launch_dmma_inst<1, false, false>(h, map, p, smem, grid, a.stream);
    if (rc) return rc;
    if (sums && a.slots != nullptr) {
        a.slots->fsum = p.fsum;
        a.slots->fcnt = p.fcnt;
        a.slots->nslots = nslots;
        a.slots->slot_stride = 1;
        a.slots->nblocks = nslots;
    } else if (sums) {
        const int len = a.k * (D + 1);
        reduce_dmma_kernel<<<(len + 31) / 32, 256, 0, a.stream>>>(p.fsum, p.fcnt, nslots, a.k, a.partials, a.state);
        HK_CUDA(cudaGetLastError());
        h->launches++;
    }
    if (a.fv_out) {
        reduce_scalar_dmma_kernel<<<1, 32, 0, a.stream>>>(p.fv_part, grid, a.fv_out);
        HK_CUDA(cudaGetLastError());
        h->launches++;
    }
    return 0;
}

}  // namespace hk
