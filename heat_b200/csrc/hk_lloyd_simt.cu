// Exact-FMA Lloyd pass (fp32 / fp64) — one fused pass over a rank's row shard:
//   stream X tiles into shared memory (cp.async.bulk + mbarrier ring)
//   -> per-row distances |x|^2 + |c|^2 - 2 x.c (reference formula, heat/spatial/distance.py:59-64)
//   -> first-index argmin (torch.min semantics, heat/core/statistics.py:177)
//   -> in-tile counting sort of rows by label (deterministic)
//   -> segmented column sums, flushed into per-CTA fp64 accumulators
//   -> per-CTA partials [k x (d+1)] written once at kernel end (no atomics on the fast path).
// Replaces _assign_to_cluster + KMeans._update_centroids (heat/cluster/_kcluster.py:352-370,
// heat/cluster/kmeans.py:76-103) for one shard.  This is the path for small k*d (HBM-bound regime)
// and the exact-arithmetic fallback for everything the tensor-core path does not cover.
#include "hk_common.cuh"

namespace hk {
namespace {

constexpr int NT = 256;  // threads per CTA; a tile holds at most NT rows (thread == row in phase A)
constexpr int NWARP = NT / 32;

enum { SUMS_NONE = 0, SUMS_SMEM = 1, SUMS_ATOMIC = 2 };

struct SimtParams {
    const void* X;
    int64_t n;
    int d;
    int64_t ldx;
    const void* C;
    int k;
    void* labels;
    int label_kind;
    double* part;     // SUMS_SMEM: [grid][k*(d+1)]   SUMS_ATOMIC: [k*(d+1)] pre-zeroed
    double* fv_part;  // [grid] or nullptr
    int kc;           // centroids resident in smem at a time (== k: loaded once)
    int stages;
    int tile_rows;    // <= NT, multiple of 32
    int use_bulk;     // ldx == d and X 16-byte aligned: tiles are contiguous -> 1-D bulk copies
    int nsub;         // sub-slices per cluster in phase C (>=1)
    int64_t num_tiles;
    const int32_t* state;  // optional sticky "converged" flag: skip the pass when set
};

struct Layout {
    size_t xt, cs, cn, sums, cnts, wcnt, wpre, tcnt, seg, lab, perm, mbar, red, total;
};

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__host__ __device__ inline Layout make_layout(int d, int k, int kc, int stages, int tile_rows, int nsub,
                                              int esize, int sums_mode) {
    Layout L;
    size_t o = 0;
    L.xt = o;
    o += align_up((size_t)stages * tile_rows * d * esize, 128);
    L.cs = o;
    o += align_up((size_t)kc * d * esize, 16);
    L.cn = o;
    o += align_up((size_t)kc * esize, 16);
    L.sums = o;
    if (sums_mode == SUMS_SMEM) o += align_up((size_t)nsub * k * d * 8, 16);
    L.cnts = o;
    if (sums_mode == SUMS_SMEM) o += align_up((size_t)k * 8, 16);
    L.wcnt = o;
    if (sums_mode == SUMS_SMEM) o += align_up((size_t)NWARP * (k + 1) * 4, 16);
    L.wpre = o;
    if (sums_mode == SUMS_SMEM) o += align_up((size_t)NWARP * (k + 1) * 4, 16);
    L.tcnt = o;
    if (sums_mode == SUMS_SMEM) o += align_up((size_t)(k + 1) * 4, 16);
    L.seg = o;
    if (sums_mode == SUMS_SMEM) o += align_up((size_t)(k + 2) * 4, 16);
    L.lab = o;
    o += (size_t)NT * 4;
    L.perm = o;
    o += (size_t)NT * 2;
    L.mbar = o;
    o += align_up((size_t)stages * 8, 16);
    L.red = o;
    o += NWARP * 8;
    L.total = o;
    return L;
}

template <typename T>
struct Vec16;
template <>
struct Vec16<float> {
    using type = float4;
    static constexpr int N = 4;
};
template <>
struct Vec16<double> {
    using type = double2;
    static constexpr int N = 2;
};

__device__ __forceinline__ void store_label(void* labels, int kind, int64_t row, int lab) {
    if (kind == HK_LABEL_I64)
        reinterpret_cast<long long*>(labels)[row] = lab;
    else if (kind == HK_LABEL_I32)
        reinterpret_cast<int*>(labels)[row] = lab;
    else if (kind == HK_LABEL_U8)
        reinterpret_cast<unsigned char*>(labels)[row] = (unsigned char)lab;
}

// running first-index argmin with torch.min NaN semantics (a NaN wins, the first one sticks)
template <typename T>
__device__ __forceinline__ void argmin_step(T d2, int j, T& best, int& lab) {
    d2 = d2 < T(0) ? T(0) : d2;  // clamp(d2, 0, inf) (distance.py:64); NaN stays NaN
    if (d2 < best || (d2 != d2 && best == best)) {
        best = d2;
        lab = j;
    }
}

template <typename T, int D, int SM>
__global__ void __launch_bounds__(NT) lloyd_simt_kernel(const SimtParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    if (p.state != nullptr && p.state[0] != 0) return;  // uniform across the grid
    const int d = (D > 0) ? D : p.d;
    const int k = p.k;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const Layout L = make_layout(d, k, p.kc, p.stages, p.tile_rows, p.nsub, (int)sizeof(T), SM);
    T* xt_base = reinterpret_cast<T*>(smem + L.xt);
    T* cs = reinterpret_cast<T*>(smem + L.cs);
    T* cn = reinterpret_cast<T*>(smem + L.cn);
    double* sums = reinterpret_cast<double*>(smem + L.sums);
    unsigned long long* cnts = reinterpret_cast<unsigned long long*>(smem + L.cnts);
    int* wcnt = reinterpret_cast<int*>(smem + L.wcnt);
    int* wpre = reinterpret_cast<int*>(smem + L.wpre);
    int* tcnt = reinterpret_cast<int*>(smem + L.tcnt);
    int* seg = reinterpret_cast<int*>(smem + L.seg);
    int* labs = reinterpret_cast<int*>(smem + L.lab);
    unsigned short* perm = reinterpret_cast<unsigned short*>(smem + L.perm);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + L.mbar);
    double* red = reinterpret_cast<double*>(smem + L.red);

    const T* X = reinterpret_cast<const T*>(p.X);
    const T* Cg = reinterpret_cast<const T*>(p.C);
    const int TR = p.tile_rows;
    const size_t stage_elems = (size_t)TR * d;
    const uint32_t stage_bytes = (uint32_t)(stage_elems * sizeof(T));
    const bool single_chunk = (p.kc >= k);

    if (tid == 0) {
        for (int s = 0; s < p.stages; ++s) mbar_init(&mbar[s], 1);
        mbar_fence_init();
    }
    if (SM == SUMS_SMEM) {
        for (int i = tid; i < p.nsub * k * d; i += NT) sums[i] = 0.0;
        for (int i = tid; i < k; i += NT) cnts[i] = 0ull;
        for (int i = tid; i < NWARP * (k + 1); i += NT) wcnt[i] = 0;
    }
    if (single_chunk) {
        for (int i = tid; i < k * d; i += NT) cs[i] = Cg[i];
    }
    __syncthreads();
    if (single_chunk) {
        for (int j = tid; j < k; j += NT) {
            T s = T(0);
            for (int i = 0; i < d; ++i) s = fma(cs[j * d + i], cs[j * d + i], s);
            cn[j] = s;
        }
        __syncthreads();
    }

    // prologue: fill the ring
    if (tid == 0 && p.use_bulk) {
        for (int s = 0; s < p.stages; ++s) {
            int64_t tile = (int64_t)blockIdx.x + (int64_t)s * gridDim.x;
            if (tile < p.num_tiles && (tile + 1) * TR <= p.n) {
                mbar_expect_tx(&mbar[s], stage_bytes);
                bulk_g2s(xt_base + (size_t)s * stage_elems, X + (size_t)tile * stage_elems, stage_bytes,
                         &mbar[s]);
            }
        }
    }

    double fv_acc = 0.0;

    for (int64_t it = 0;; ++it) {
        const int64_t tile = (int64_t)blockIdx.x + it * gridDim.x;
        if (tile >= p.num_tiles) break;
        const int s = (int)(it % p.stages);
        const uint32_t parity = (uint32_t)((it / p.stages) & 1);
        const int64_t row0 = tile * TR;
        const int rows = (int)((p.n - row0) < (int64_t)TR ? (p.n - row0) : (int64_t)TR);
        const bool bulk = p.use_bulk && rows == TR;
        T* xt = xt_base + (size_t)s * stage_elems;

        if (bulk) {
            mbar_wait(&mbar[s], parity);
        } else {
            // cooperative copy (tail tile, strided rows or unaligned base)
            const int tot = rows * d;
            for (int i = tid; i < tot; i += NT) {
                int r = i / d, c = i - r * d;
                xt[i] = X[(size_t)(row0 + r) * p.ldx + c];
            }
            __syncthreads();
        }

        // ---------------- phase A: thread == row, distances + first-index argmin -----------------
        int lab = k;  // dummy bin for rows past the end of the tile
        T best = T(INFINITY);
        const bool active = tid < rows;

        if constexpr (D > 0) {
            constexpr int DD = D > 0 ? D : 1;
            constexpr int EPC = Vec16<T>::N;
            constexpr int CH = DD / EPC > 0 ? DD / EPC : 1;
            using V = typename Vec16<T>::type;
            T x[DD];
            T xn = T(0);
            if (active) {
                const unsigned char* rp = reinterpret_cast<const unsigned char*>(xt + (size_t)tid * DD);
#pragma unroll
                for (int q = 0; q < CH; ++q) {
                    const int ch = (q + tid) & (CH - 1);  // rotate chunks: conflict-free row reads
                    union {
                        V v;
                        T e[EPC];
                    } u;
                    u.v = *reinterpret_cast<const V*>(rp + ch * 16);
#pragma unroll
                    for (int e = 0; e < EPC; ++e) x[q * EPC + e] = u.e[e];
                }
#pragma unroll
                for (int i = 0; i < DD; ++i) xn = fma(x[i], x[i], xn);
            }
            for (int j0 = 0; j0 < k; j0 += p.kc) {
                const int kcur = (k - j0) < p.kc ? (k - j0) : p.kc;
                if (!single_chunk) {
                    __syncthreads();
                    for (int i = tid; i < kcur * d; i += NT) cs[i] = Cg[(size_t)j0 * d + i];
                    __syncthreads();
                    for (int j = tid; j < kcur; j += NT) {
                        T sq = T(0);
                        for (int i = 0; i < d; ++i) sq = fma(cs[j * d + i], cs[j * d + i], sq);
                        cn[j] = sq;
                    }
                    __syncthreads();
                }
                if (active) {
#pragma unroll 2
                    for (int j = 0; j < kcur; ++j) {
                        const unsigned char* cp =
                            reinterpret_cast<const unsigned char*>(cs + (size_t)j * DD);
                        T dot = T(0);
#pragma unroll
                        for (int q = 0; q < CH; ++q) {
                            const int ch = (q + tid) & (CH - 1);
                            union {
                                V v;
                                T e[EPC];
                            } u;
                            u.v = *reinterpret_cast<const V*>(cp + ch * 16);
#pragma unroll
                            for (int e = 0; e < EPC; ++e) dot = fma(x[q * EPC + e], u.e[e], dot);
                        }
                        const T d2 = (xn + cn[j]) - T(2) * dot;
                        argmin_step<T>(d2, j0 + j, best, lab);
                    }
                }
            }
        } else {
            const T* xr = xt + (size_t)tid * d;
            T xn = T(0);
            if (active)
                for (int i = 0; i < d; ++i) xn = fma(xr[i], xr[i], xn);
            for (int j0 = 0; j0 < k; j0 += p.kc) {
                const int kcur = (k - j0) < p.kc ? (k - j0) : p.kc;
                if (!single_chunk) {
                    __syncthreads();
                    for (int i = tid; i < kcur * d; i += NT) cs[i] = Cg[(size_t)j0 * d + i];
                    __syncthreads();
                    for (int j = tid; j < kcur; j += NT) {
                        T sq = T(0);
                        for (int i = 0; i < d; ++i) sq = fma(cs[j * d + i], cs[j * d + i], sq);
                        cn[j] = sq;
                    }
                    __syncthreads();
                }
                if (active) {
                    for (int j = 0; j < kcur; ++j) {
                        const T* cr = cs + (size_t)j * d;
                        T dot = T(0);
                        for (int i = 0; i < d; ++i) dot = fma(xr[i], cr[i], dot);
                        const T d2 = (xn + cn[j]) - T(2) * dot;
                        argmin_step<T>(d2, j0 + j, best, lab);
                    }
                }
            }
        }

        if (active) {
            if (p.label_kind != HK_LABEL_NONE) store_label(p.labels, p.label_kind, row0 + tid, lab);
            if (p.fv_part != nullptr) {
                // reference: (||min_j sqrt(clamp(d2))||_2)^2  (heat/cluster/_kcluster.py:367-368)
                const T sq = sqrt(best);
                fv_acc += (double)(sq * sq);
            }
        }

        if (SM == SUMS_SMEM) {
            // ------------- phase B: deterministic counting sort of the tile's rows by label --------
            const unsigned peers = __match_any_sync(0xffffffffu, lab);
            const int rank = __popc(peers & lanemask_lt());
            const int leader = __ffs(peers) - 1;
            if (lane == leader) wcnt[warp * (k + 1) + lab] = __popc(peers);
            __syncthreads();
            for (int c = tid; c <= k; c += NT) {
                int run = 0;
#pragma unroll
                for (int w = 0; w < NWARP; ++w) {
                    const int v = wcnt[w * (k + 1) + c];
                    wcnt[w * (k + 1) + c] = 0;  // leave the table clean for the next tile
                    wpre[w * (k + 1) + c] = run;
                    run += v;
                }
                tcnt[c] = run;
                if (c < k) cnts[c] += (unsigned long long)run;
            }
            __syncthreads();
            if (warp == 0) {
                const int per = (k + 1 + 31) / 32;
                const int b0 = lane * per;
                int local = 0;
                for (int i = 0; i < per; ++i)
                    if (b0 + i <= k) local += tcnt[b0 + i];
                int incl = local;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                int run = incl - local;
                for (int i = 0; i < per; ++i)
                    if (b0 + i <= k) {
                        seg[b0 + i] = run;
                        run += tcnt[b0 + i];
                    }
                if (lane == 31) seg[k + 1] = incl;
            }
            __syncthreads();
            perm[seg[lab] + wpre[warp * (k + 1) + lab] + rank] = (unsigned short)tid;
            __syncthreads();

            // ------------- phase C: segmented column sums (thread == (feature, slice)) --------------
            const int FW = d < NT ? d : NT;
            const int S = NT / FW;
            const int f0 = tid % FW;
            const int sl = tid / FW;
            if (sl < S) {
                const int nslots = p.nsub * k;
                for (int v = sl; v < nslots; v += S) {
                    const int c = v / p.nsub;
                    const int sub = v - c * p.nsub;
                    const int b = seg[c], e = seg[c + 1];
                    if (b + sub >= e) continue;
                    for (int f = f0; f < d; f += FW) {
                        double tot = 0.0;
                        T acc = T(0);
                        int run = 0;
                        for (int i = b + sub; i < e; i += p.nsub) {
                            acc += xt[(size_t)perm[i] * d + f];
                            if (++run == 32) {  // keep low-precision partial sums short (<= 32 rows)
                                tot += (double)acc;
                                acc = T(0);
                                run = 0;
                            }
                        }
                        tot += (double)acc;
                        sums[((size_t)sub * k + c) * d + f] += tot;
                    }
                }
            }
            __syncthreads();
        } else if (SM == SUMS_ATOMIC) {
            labs[tid] = lab;
            __syncthreads();
            if (active) atomicAdd(&p.part[(size_t)lab * (d + 1) + d], 1.0);
            const int FW = d < NT ? d : NT;
            const int S = NT / FW;
            const int f0 = tid % FW;
            const int sl = tid / FW;
            if (sl < S) {
                for (int r = sl; r < rows; r += S) {
                    const int l = labs[r];
                    for (int f = f0; f < d; f += FW)
                        atomicAdd(&p.part[(size_t)l * (d + 1) + f], (double)xt[(size_t)r * d + f]);
                }
            }
            __syncthreads();
        } else {
            __syncthreads();
        }

        // stage s is free again: prefetch the tile that will land in it
        if (tid == 0 && p.use_bulk) {
            const int64_t nt = tile + (int64_t)p.stages * gridDim.x;
            if (nt < p.num_tiles && (nt + 1) * TR <= p.n) {
                mbar_expect_tx(&mbar[s], stage_bytes);
                bulk_g2s(xt, X + (size_t)nt * stage_elems, stage_bytes, &mbar[s]);
            }
        }
    }

    if (SM == SUMS_SMEM) {
        __syncthreads();
        double* out = p.part + (size_t)blockIdx.x * k * (d + 1);
        for (int i = tid; i < k * d; i += NT) {
            const int c = i / d, f = i - c * d;
            double t = 0.0;
            for (int sub = 0; sub < p.nsub; ++sub) t += sums[((size_t)sub * k + c) * d + f];
            out[(size_t)c * (d + 1) + f] = t;
        }
        for (int c = tid; c < k; c += NT) out[(size_t)c * (d + 1) + d] = (double)cnts[c];
    }
    if (p.fv_part != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) fv_acc += __shfl_xor_sync(0xffffffffu, fv_acc, o);
        if (lane == 0) red[warp] = fv_acc;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < NWARP; ++w) t += red[w];
            p.fv_part[blockIdx.x] = t;
        }
    }
}

// out[i] = sum_b part[b][i] in fixed order b = 0..nb-1 (deterministic); optional scalar reduce
__global__ void reduce_partials_kernel(const double* __restrict__ part, int nb, int len,
                                       double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    double t = 0.0;
    for (int b = 0; b < nb; ++b) t += part[(size_t)b * len + i];
    out[i] = t;
}

template <typename T, int D, int SM>
int launch_one(Handle* h, SimtParams& p, size_t smem, int& grid_out, bool query_only) {
    auto kern = lloyd_simt_kernel<T, D, SM>;
    HK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    HK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem));
    if (occ < 1) {
        set_error("lloyd_simt: kernel does not fit on an SM (smem %zu)", smem);
        return -2;
    }
    int64_t grid = (int64_t)h->num_sms * occ;
    if (grid > p.num_tiles) grid = p.num_tiles;
    grid_out = (int)grid;
    if (query_only) return 0;
    return 0;
}

template <typename T, int D, int SM>
int run_variant(Handle* h, SimtParams p, size_t smem, const LloydArgs& a, int len) {
    auto kern = lloyd_simt_kernel<T, D, SM>;
    int grid = 0;
    int rc = launch_one<T, D, SM>(h, p, smem, grid, false);
    if (rc) return rc;
    double* fv_part = nullptr;
    if (SM == SUMS_SMEM) {
        rc = ensure_part(h, ((size_t)grid * len + (a.fv_out ? grid : 0)) * sizeof(double));
        if (rc) return rc;
        p.part = h->part;
        if (a.fv_out) fv_part = h->part + (size_t)grid * len;
    } else {
        if (a.fv_out) {
            rc = ensure_part(h, (size_t)grid * sizeof(double));
            if (rc) return rc;
            fv_part = h->part;
        }
        if (SM == SUMS_ATOMIC) {
            HK_CUDA(cudaMemsetAsync(a.partials, 0, (size_t)len * sizeof(double), a.stream));
            p.part = a.partials;
        }
    }
    p.fv_part = fv_part;
    prof_begin(h, a.stream);
    kern<<<grid, NT, smem, a.stream>>>(p);
    prof_end(h, a.stream);
    HK_CUDA(cudaGetLastError());
    h->launches++;
    if (SM == SUMS_SMEM) {
        reduce_partials_kernel<<<(len + 255) / 256, 256, 0, a.stream>>>(h->part, grid, len, a.partials);
        HK_CUDA(cudaGetLastError());
        h->launches++;
    }
    if (a.fv_out) {
        reduce_partials_kernel<<<1, 32, 0, a.stream>>>(fv_part, grid, 1, a.fv_out);
        HK_CUDA(cudaGetLastError());
        h->launches++;
    }
    return 0;
}

template <typename T, int D>
int dispatch_sm(Handle* h, const SimtParams& p, size_t smem, const LloydArgs& a, int len, int sm) {
    switch (sm) {
        case SUMS_NONE:
            return run_variant<T, D, SUMS_NONE>(h, p, smem, a, len);
        case SUMS_SMEM:
            return run_variant<T, D, SUMS_SMEM>(h, p, smem, a, len);
        default:
            return run_variant<T, D, SUMS_ATOMIC>(h, p, smem, a, len);
    }
}

}  // namespace

int launch_lloyd_simt(Handle* h, const LloydArgs& a) {
    const int esize = a.dtype == HK_F64 ? 8 : 4;
    const int d = a.d, k = a.k;
    const int len = k * (d + 1);
    const size_t budget = (size_t)h->smem_optin;

    // tile height: keep one stage <= 48 KB, multiple of 32 rows, at most NT
    int tile_rows = NT;
    while (tile_rows > 32 && (size_t)tile_rows * d * esize > 48 * 1024) tile_rows -= 32;
    if ((size_t)tile_rows * d * esize > 96 * 1024) {
        set_error("lloyd_simt: rows of %d x %d bytes are too wide for the SIMT path", d, esize);
        return -2;
    }
    const int FW = d < NT ? d : NT;
    const int S = NT / FW;
    int nsub = S / k;
    if (nsub < 1) nsub = 1;

    int sums_mode = a.partials ? SUMS_SMEM : SUMS_NONE;
    int kc = k;
    int stages = 3;
    Layout L = make_layout(d, k, kc, stages, tile_rows, nsub, esize, sums_mode);
    auto fits = [&](const Layout& l) { return l.total <= budget; };
    if (!fits(L) || k > 4096) {
        // try fewer stages, then fall back to chunked centroids + global atomics
        stages = 2;
        L = make_layout(d, k, kc, stages, tile_rows, nsub, esize, sums_mode);
        if (!fits(L) || k > 4096) {
            if (sums_mode == SUMS_SMEM) sums_mode = SUMS_ATOMIC;
            const size_t stage_b = (size_t)tile_rows * d * esize;
            stages = stage_b * 3 <= 96 * 1024 ? 3 : (stage_b * 2 <= 128 * 1024 ? 2 : 1);
            size_t left = budget - stages * stage_b - 8 * 1024;
            kc = (int)(left / ((size_t)d * esize + esize));
            if (kc > k) kc = k;
            if (kc < 1) {
                set_error("lloyd_simt: d=%d too large for shared memory", d);
                return -2;
            }
            L = make_layout(d, k, kc, stages, tile_rows, nsub, esize, sums_mode);
            if (!fits(L)) {
                set_error("lloyd_simt: layout does not fit (%zu > %zu)", L.total, budget);
                return -2;
            }
        }
    } else {
        // prefer 2 CTAs/SM when the footprint allows it (phases of different CTAs overlap)
        if (L.total * 2 > budget) {
            Layout L2 = make_layout(d, k, kc, 2, tile_rows, nsub, esize, sums_mode);
            if (L2.total * 2 <= budget) {
                stages = 2;
                L = L2;
            }
        }
    }

    SimtParams p;
    p.X = a.X;
    p.n = a.n;
    p.d = d;
    p.ldx = a.ldx;
    p.C = a.C;
    p.k = k;
    p.labels = a.labels;
    p.label_kind = a.labels ? a.label_kind : HK_LABEL_NONE;
    p.part = nullptr;
    p.fv_part = nullptr;
    p.kc = kc;
    p.stages = stages;
    p.tile_rows = tile_rows;
    p.use_bulk = (a.ldx == d) && ((reinterpret_cast<uintptr_t>(a.X) & 15) == 0);
    p.nsub = nsub;
    p.num_tiles = (a.n + tile_rows - 1) / tile_rows;
    p.state = a.state;

    char name[96];
    int D = 0;
    if (a.dtype == HK_F32) {
        if (d == 4 || d == 8 || d == 16 || d == 32 || d == 64) D = d;
    } else {
        if (d == 2 || d == 4 || d == 8 || d == 16 || d == 32) D = d;
    }
    snprintf(name, sizeof(name), "simt<%s,D=%d,%s,kc=%d,stages=%d,tile=%d>",
             a.dtype == HK_F64 ? "f64" : "f32", D,
             sums_mode == SUMS_SMEM ? "smem" : (sums_mode == SUMS_ATOMIC ? "atomic" : "none"), kc, stages,
             tile_rows);
    h->variant = name;

#define HK_RUN(T, DV) return dispatch_sm<T, DV>(h, p, L.total, a, len, sums_mode)
    if (a.dtype == HK_F32) {
        switch (D) {
            case 4: HK_RUN(float, 4);
            case 8: HK_RUN(float, 8);
            case 16: HK_RUN(float, 16);
            case 32: HK_RUN(float, 32);
            case 64: HK_RUN(float, 64);
            default: HK_RUN(float, 0);
        }
    } else {
        switch (D) {
            case 2: HK_RUN(double, 2);
            case 4: HK_RUN(double, 4);
            case 8: HK_RUN(double, 8);
            case 16: HK_RUN(double, 16);
            case 32: HK_RUN(double, 32);
            default: HK_RUN(double, 0);
        }
    }
#undef HK_RUN
}

}  // namespace hk
