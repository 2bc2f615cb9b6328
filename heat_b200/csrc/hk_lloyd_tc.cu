// Tensor-core Lloyd pass (fp32 data) for sm_100a: TMA-streamed row tiles, tcgen05 TF32 distance filter with
// fp32 accumulators in TMEM, exact-FMA refinement of the rows the filter cannot decide, and deterministic
// per-cluster column sums in private shared-memory accumulators.
//
// Why a filter: at k=64, d=32 the distance contraction costs 2*k = 128 FLOP per 4-byte element, three
// times what the FP32 pipes can sustain at HBM speed, so x.c^T runs on the 5th-gen tensor cores
// (tcgen05.mma kind::tf32 reads the fp32 tile in shared memory directly and ignores the low 13 mantissa
// bits).  Distances are translation invariant in the centroid index: with mu = mean_j c_j and c'_j = c_j - mu,
//      |c_j|^2 - 2 x.c_j  =  (|c_j|^2 - 2 x.c'_j) - 2 x.mu,   and  -2 x.mu  does not depend on j,
// so the MMA operand holds the CENTRED centroids (the filter error scales with |x| * max|c'| instead of
// |x| * max|c|, which is what matters for data far from the origin) while every exact evaluation uses the raw
// values.  s_j = |c_j|^2 - 2 x.c'_j (TF32) differs from the exact fp32 formula value of
// heat/spatial/distance.py:59-64 (minus terms constant in j) by at most
//      E = 2*beta*|x|*max_j|c'_j| + (2d+6)*2^-23*(|x|^2 + max_j|c_j|^2 + max_j|c'_j|^2),   beta = 1.05 * 2^-9
// (TF32 truncation of both operands; fp32 rounding of the reference formula and of the tensor-core accumulation),
// so a row whose runner-up is more than 2E above the minimum has a certain label (identical to what the exact
// formula + first-index argmin would give).  For every other row the exact argmin provably lies in the CANDIDATE
// set {j : s_j < min + 2E}; only those (row, centroid) pairs are re-evaluated with the exact formula, spread over
// the lanes of the warp.  Rows with NaN/Inf go through the exact formula over all centroids.  Labels are therefore
// those of the exact-FMA path, the tensor cores only remove work.
//
// Warp roles per CTA (1 CTA per SM, persistent, static tile -> CTA map); all hand-offs are mbarriers with
// one arrival per warp, there is no CTA- or group-wide barrier in the steady state:
//   warp 0      : TMA producer (cp.async.bulk.tensor, 128B swizzle, EVICT_FIRST) into an S-stage ring
//   warps 1, 3  : tcgen05.mma issuers for even / odd local tiles (one elected lane each); accumulator buffer =
//                 tile % NBUF in TMEM
//   warp 2      : TMEM allocation / release
//   warps 4-15  : epilogue warps.  Warp (q, r) owns TMEM lane quarter q of the tiles i == r (mod 3):
//                 tcgen05.ld -> min + sign mask -> label (or exact re-evaluation) -> label to shared memory
//   warps 16-23 : accumulator warps (only when sums are wanted).  Warp (q, r') owns rows of lane quarter q of
//                 the tiles i == r' (mod NA/4) and a PRIVATE fp32 [k+1][d] accumulator in shared memory:
//                 lanes cover 128/d rows x d/4 feature quads per step, add the row into the accumulator row
//                 of its label with plain load-add-store (no atomics: the array is private, label collisions
//                 inside a step are detected up front and serialised).  The array is widened into a per-warp
//                 fp64 slot in global memory (L2) before any cluster can have received more than 160 rows,
//                 so fp32 partial sums stay short.  Fixed row order and fixed reduction order -> deterministic.
//                 The accumulator warps are the slowest stage: they wait on ONE barrier per tile (labels published;
//                 the epilogue warps observe the TMA barrier first, which orders the tile reads transitively), and
//                 the label multiplicities they need for the flush rule arrive through a barrier-free tagged ring.
// The accumulator is seeded with |c_j|^2 by one extra k-step (ones x three exact TF32 pieces of |c_j|^2)
// and B holds -2*c, so TMEM already contains s_j = |c_j|^2 - 2 x.c_j and the epilogue is min + sign-mask.
// |x|^2 (needed only for the bound E) is computed per row in the first pass over a matrix and cached as a
// per-tile maximum in the handle (like sklearn's x_squared_norms), later passes read one float per tile.
// Replaces _assign_to_cluster + KMeans._update_centroids for one shard
// (heat/cluster/_kcluster.py:352-370, heat/cluster/kmeans.py:76-103).
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "hk_tma.cuh"

namespace hk {
namespace {

// Tunables (measured at config 3 on B200: ER=3/MI=2 3.06 ms, ER=4/MI=2 3.29 ms, ER=2/MI=2 3.21 ms, ER=3/MI=1 3.15 ms):
// the kernel is bound by instruction issue, and every waiting warp polls its barrier, so fewer warps win as long
// as no role becomes the bottleneck.
#ifndef HK_TC_ER
#define HK_TC_ER 3  // tile residues handled by the epilogue warps (4 warps each)
#endif
#ifndef HK_TC_MI
#define HK_TC_MI 2  // MMA issuer warps
#endif
constexpr int TM = 128;         // rows per tile (UMMA M)
constexpr int MISC_WARPS = 4;   // producer, MMA, TMEM allocator, spare
constexpr int E_WARPS = 4 * HK_TC_ER;  // epilogue warps: 4 lane quarters x ER tile residues
constexpr int ER = HK_TC_ER;
constexpr int A_WARPS_MAX = 8;  // accumulator warps (8, or 4 when the private accumulators are large)
constexpr int E_FIRST = MISC_WARPS;
constexpr int A_FIRST = MISC_WARPS + E_WARPS;

struct TcParams {
    int64_t n;
    int d;
    int k;
    int nk;  // k rounded up to a multiple of 32 (UMMA N, TMEM columns per accumulator)
    const float* C;
    void* labels;
    int label_kind;
    double* fsum;     // [grid*NA][k*d] per-accumulator-warp fp64 sums, or nullptr (assign only)
    double* fcnt;     // [grid][k] per-CTA cluster counts
    double* fv_part;  // [grid] or nullptr
    int S;            // smem stages
    int nbuf;         // TMEM accumulator buffers (distance tiles in flight)
    int NA;           // accumulator warps (8 or 4)
    int num_tiles;
    const int32_t* state;
    uint32_t tmem_cols;
    float* bounds;   // caller-owned [num_tiles] per-tile max |x| + one int "filled" flag, or nullptr (no cache)
    unsigned long long* stats;  // [6] cumulative counters: undecided rows, exact (row, centroid) pairs, rows sent through
                                // the all-centroid formula, warps that entered the cold path, rows seen, passes
    // shared-memory layout (byte offsets from the 1024-aligned base), computed on the host
    uint32_t o_stages, o_B, o_Aext, o_Bext, o_cn, o_acc, o_lab, o_cnt, o_snap, o_bars, o_misc, o_ring;
    int ring_n, maxp;         // undecided-row ring entries (power of two) / pair-scratch entries
    unsigned long long* dbg;  // optional [grid][32 warps][8] cycle counters (HK_TC_DEBUG=1)
    unsigned long long* tl;   // optional timeline of CTA 0: [512 local tiles][8 events] clock64 stamps
};

struct TcLayout {
    size_t stages, B, Aext, Bext, cn, acc, lab, cnt, snap, bars, misc, ring, total;
};

__host__ inline size_t up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__host__ inline TcLayout tc_layout(int d, int k, int nk, int S, int NA, bool sums, int ring_n = 128, int maxp = 96) {
    TcLayout L;
    size_t o = 0;
    L.stages = o;
    o += (size_t)S * TM * d * 4;
    L.B = o;
    o += up((size_t)nk * d * 4, 1024);
    L.Aext = o;  // eight identical rows (1,1,1,0..): the descriptor's 8-row group stride is 0
    o += 1024;
    L.Bext = o;
    o += up((size_t)nk * 128, 1024);
    L.cn = o;
    o += up((size_t)nk * 4, 16);
    L.acc = o;  // NA private fp32 accumulators [k+1][d] (row k swallows the rows past the end of X)
    if (sums) o += up((size_t)NA * (k + 1) * d * 4, 16);
    L.lab = o;  // per stage: 128 labels (u16)
    if (sums) o += (size_t)S * TM * 2;
    L.cnt = o;  // private cluster counts of the 16 epilogue warps (int)
    if (sums) o += up((size_t)E_WARPS * (2 * k + 1) * 4, 16);  // + per-tile label histograms [E_WARPS][k+1]
    L.snap = o;  // multiplicity ring: [32 local tiles][4 lane quarters] largest label multiplicity among the 32 rows
    if (sums) o += 32 * 4 * 4;
    L.bars = o;
    o += 8 * 96;  // mbarriers
    L.misc = o;  // tmem slot, maxima, flags, fv partials, cold-path counters, ring control words, mu[d]
    o += 1024;
    L.ring = o;  // undecided-row ring (ring_n x 16 B) + pair scratch of the refine warp (maxp x 16 B)
    o += (size_t)ring_n * 16 + (size_t)maxp * 16;
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

// exact fp32 squared distance of the reference formula (heat/spatial/distance.py:59-64) for (row, centroid j):
// fl(fl(|x|^2 + |c_j|^2) - 2 fl(x.c_j)), features accumulated in ascending order with FMAs (the arithmetic of the
// exact-FMA kernels).  x comes from the swizzled tile in shared memory, c from global memory (raw values, L1/L2
// resident: the MMA operand in shared memory holds the centred ones).
__device__ __forceinline__ float exact_pair_inl(uint32_t xt, int row, const float* __restrict__ C, int j, int d, float cnj) {
    float dot = 0.f, xn = 0.f;
    const float4* cr = reinterpret_cast<const float4*>(C + (size_t)j * d);
#pragma unroll 8
    for (int f = 0; f < d; f += 4) {  // d is a multiple of 32: the eight loads of a block are issued before its FMA chains
        const float4 xv = lds_f4(xt + sw128_off(TM, row, f));
        const float4 cv = __ldg(cr + (f >> 2));
        xn = fmaf(xv.x, xv.x, xn);
        xn = fmaf(xv.y, xv.y, xn);
        xn = fmaf(xv.z, xv.z, xn);
        xn = fmaf(xv.w, xv.w, xn);
        dot = fmaf(xv.x, cv.x, dot);
        dot = fmaf(xv.y, cv.y, dot);
        dot = fmaf(xv.z, cv.z, dot);
        dot = fmaf(xv.w, cv.w, dot);
    }
    const float d2 = fmaf(-2.f, dot, xn + cnj);
    return d2 < 0.f ? 0.f : d2;
}

__device__ __forceinline__ float exact_pair(uint32_t xt, int row, const float* __restrict__ C, int j, int d, float cnj) {
    return exact_pair_inl(xt, row, C, j, d, cnj);
}

__device__ __noinline__ float row_norm2(uint32_t xt, int row, int d) {
    float xn = 0.f;
    for (int f = 0; f < d; f += 4) {
        const float4 xv = lds_f4(xt + sw128_off(TM, row, f));
        xn = fmaf(xv.x, xv.x, xn);
        xn = fmaf(xv.y, xv.y, xn);
        xn = fmaf(xv.z, xv.z, xn);
        xn = fmaf(xv.w, xv.w, xn);
    }
    return xn;
}

// torch.min semantics: a strictly smaller value wins, the first NaN wins and sticks
__device__ __forceinline__ void take_min(float v, int j, float& best, int& bl) {
    if (v < best || (v != v && best == best)) {
        best = v;
        bl = j;
    }
}

enum { XN_COMPUTE = 0, XN_WRITE = 1, XN_READ = 2 };

// cycle counters per role / per-tile timeline of CTA 0: compiled in with -DHK_TC_TIMING only, the hot loops of the
// shipped kernel carry no instrumentation
#ifdef HK_TC_TIMING
#define TC_T(...) __VA_ARGS__
#else
#define TC_T(...)
#endif

// ---- exact refinement of the rows the filter cannot decide -------------------------------------------------------
// The epilogue warps never evaluate the exact formula themselves (a call inside their loop costs the whole hot path
// its uniform registers: +25 % on decided-only data).  They push one 16-byte entry per undecided row into a ring in
// shared memory and move on; the otherwise idle warp 2 drains the ring, POOLING the rows of all four lane quarters
// (and of consecutive tiles): the (row, candidate) pairs of up to 32 rows are flattened over its lanes, so a tile with
// a few undecided rows of a few candidates each costs about one exact evaluation per lane instead of one long chain in
// each of four warps.  The accumulator warps cannot start on a tile before its undecided rows have their final label:
// the pushing warp raises the transaction count of the tile's "labels published" mbarrier by the number of rows it
// defers, the refine warp completes them one by one.
//
// ring entry: word0 = tag (low 16 bits of slot index + 1) | stage << 16 | chunk_lo << 20 | chunk_hi << 23 | full << 26,
//             word1 = global row, word2/3 = candidate masks of columns [32 chunk_lo, +32) / [32 chunk_hi, +32)
constexpr int RING_MAX = 128;  // ring entries (a power of two; 32 when shared memory is short, see plan_tc)
constexpr int R_WARPS = 1;  // refine warps (warp 2).  Measured and dropped: a second refine warp (25 warps leave 72
                            // registers per thread: decided-only data -7 %), and the producer / MMA / epilogue warps
                            // lending a hand between two tiles (their pipelines stall: unstructured data +35 %)


__device__ __forceinline__ uint32_t lds_u32_volatile(uint32_t a) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u32_volatile(uint32_t a, uint32_t v) {
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t bar, uint32_t n) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_complete_tx_a(uint32_t bar, uint32_t n) {
    asm volatile("mbarrier.complete_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(n) : "memory");
}

// addresses and flags a warp needs to drain the ring
struct RefineCtx {
    int k, d;
    const float* C;
    void* labels;
    int label_kind;
    bool sums, want_fv;
    uint32_t a_ring, a_qalloc, a_stages, a_lab, b_lfull, a_ecnt0;
    const float* cn;
    int ring_n, maxp;  // ring entries (power of two) and pair-scratch entries of this launch
};
struct RefineStats {
    unsigned n_und = 0, n_pairs = 0, n_full = 0, n_batches = 0;
    double fv = 0.0;
};

// One batch of a refine warp (all lanes converged; a_scr = this warp's MAXP x 16-byte pair scratch):
//   1. under the consumer lock: the leading run of published ring entries, one row per lane, cut so that their
//      candidates fit the scratch; the slots are handed back at once (payload in registers), the lock released;
//   2. every row's lane computes |x|^2 once and writes one descriptor per candidate (row address, centroid, |x|^2);
//   3. the (row, centroid) pairs are evaluated 32 at a time, one per lane, whatever row they belong to - a batch of
//      rows with a few candidates each costs a couple of exact evaluations per lane, not one long chain per row;
//   4. every row's lane takes the first-index minimum of its results (torch.min semantics) and publishes the label.
// A single warp issues these ~300 dependent instructions at one per 15-20 cycles next to the busy roles, which is why
// there are two refine warps and why the evaluation order keeps every chain as short as the formula allows.
// (Force-inlined on purpose: as a noinline function this code - never executed on decided-only data - made the whole
// kernel 25 % slower; letting the epilogue warps help between two tiles cost the decided-only case 13 %.  Both A/B
// measured, see profiles/README.md.)
constexpr int MAXP_MAX = 96;  // pair scratch entries (32 when shared memory is short)
__device__ __forceinline__ int refine_batch(const RefineCtx& c, int lane, uint32_t a_scr, RefineStats& rs) {
    constexpr unsigned FULLM = 0xffffffffu;
    const int k = c.k, d = c.d;
    const float* __restrict__ C = c.C;
    const float* cn = c.cn;
    const uint32_t a_qhead = c.a_qalloc + 4, a_qlock = c.a_qalloc + 12;
    const uint32_t stage_bytes = (uint32_t)TM * d * 4;
    // ---- 1. claim -------------------------------------------------------------------------------------------
    uint32_t got = 0;
    if (lane == 0) {
        uint32_t old;
        asm volatile("atom.shared.cas.b32 %0, [%1], 0, 1;" : "=r"(old) : "r"(a_qlock) : "memory");
        got = old == 0u ? 1u : 0u;
    }
    if (!__shfl_sync(FULLM, got, 0)) return 0;
    __threadfence_block();
    uint32_t head = lane == 0 ? lds_u32_volatile(a_qhead) : 0u;
    head = __shfl_sync(FULLM, head, 0);
    const uint32_t idx = head + (uint32_t)lane;
    const uint32_t ea = c.a_ring + (idx & (uint32_t)(c.ring_n - 1)) * 16;
    // one 16-byte load per entry (the pusher publishes tag and payload with one 16-byte store)
    uint32_t w0, grow;
    unsigned mlo, mhi;
    asm volatile("ld.volatile.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(grow), "=r"(mlo), "=r"(mhi) : "r"(ea) : "memory");
    const unsigned ready = __ballot_sync(FULLM, (w0 & 0xffffu) == ((idx + 1u) & 0xffffu));
    const int n_ready = ready == FULLM ? 32 : __ffs(~ready) - 1;
    if (n_ready == 0) {
        if (lane == 0) sts_u32_volatile(a_qlock, 0u);
        return 0;
    }
    bool full = false;
    if (lane < n_ready) {
        full = (w0 >> 26) & 1u;
    } else {
        grow = 0u;
        mlo = mhi = 0u;
    }
    int cnt = full ? 0 : __popc(mlo) + __popc(mhi);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULLM, incl, o);
        if (lane >= o) incl += t;
    }
    // rows whose candidates fit the scratch (a prefix: incl is non-decreasing); a first row that does not fit alone
    // goes through the all-centroid loop
    int n = __popc(__ballot_sync(FULLM, lane < n_ready && incl <= c.maxp));
    if (n == 0) {
        n = 1;
        if (lane == 0) {
            full = true;
            cnt = 0;
        }
    }
    __threadfence_block();  // payload read before the slots are handed back
    if (lane == 0) {
        sts_u32_volatile(a_qhead, head + (uint32_t)n);
        __threadfence_block();
        sts_u32_volatile(a_qlock, 0u);
    }
    const bool mine = lane < n;
    if (!mine) {
        cnt = 0;
        full = false;
    }
    const int excl = (mine && !full) ? incl - cnt : 0;
    int total = __shfl_sync(FULLM, incl, n - 1);
    if (__shfl_sync(FULLM, (int)full, 0) && n == 1) total = 0;
    const int stage = (int)((w0 >> 16) & 15u);
    const int clo = (int)((w0 >> 20) & 7u) * 32, chi = (int)((w0 >> 23) & 7u) * 32;
    const int row = (int)(grow & (TM - 1));
    const uint32_t xt = c.a_stages + (uint32_t)stage * stage_bytes;

    // ---- 2. |x|^2 of every row (features ascending, as the exact-FMA kernels) and one descriptor per candidate ------
    if (cnt > 0) {
        float xn = 0.f;
        for (int f = 0; f < d; f += 4) {
            const float4 xv = lds_f4(xt + sw128_off(TM, row, f));
            xn = fmaf(xv.x, xv.x, xn);
            xn = fmaf(xv.y, xv.y, xn);
            xn = fmaf(xv.z, xv.z, xn);
            xn = fmaf(xv.w, xv.w, xn);
        }
        unsigned rl = mlo, rh = mhi;
        uint32_t da = a_scr + (uint32_t)excl * 16;
        for (int i = 0; i < cnt; ++i, da += 16) {
            int j;
            if (rl) {
                j = clo + __ffs(rl) - 1;
                rl &= rl - 1;
            } else {
                j = chi + __ffs(rh) - 1;
                rh &= rh - 1;
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(da), "r"(xt), "r"((uint32_t)(row | (j << 8))),
                         "r"(__float_as_uint(xn)), "r"(0u)
                         : "memory");
        }
    }
    __syncwarp();
    // ---- 3. the pairs, 32 at a time ------------------------------------------------------------------------------
    for (int base = 0; base < total; base += 32) {
        const int pr = base + lane;
        if (pr < total) {
            const uint32_t da = a_scr + (uint32_t)pr * 16;
            uint32_t pxt, prj, pxn, pad;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(pxt), "=r"(prj), "=r"(pxn), "=r"(pad) : "r"(da) : "memory");
            (void)pad;
            const int prow = (int)(prj & 0xffu), j = (int)(prj >> 8);
            const float4* cr = reinterpret_cast<const float4*>(C + (size_t)j * d);
            float dot = 0.f;
#pragma unroll 8
            for (int f = 0; f < d; f += 4) {
                const float4 xv = lds_f4(pxt + sw128_off(TM, prow, f));
                const float4 cv = __ldg(cr + (f >> 2));
                dot = fmaf(xv.x, cv.x, dot);
                dot = fmaf(xv.y, cv.y, dot);
                dot = fmaf(xv.z, cv.z, dot);
                dot = fmaf(xv.w, cv.w, dot);
            }
            float d2 = fmaf(-2.f, dot, __uint_as_float(pxn) + cn[j]);
            d2 = d2 < 0.f ? 0.f : d2;
            sts_u32_volatile(da + 12, __float_as_uint(d2));
        }
    }
    __syncwarp();
    // ---- 4. first-index minimum per row ----------------------------------------------------------------------------
    float best = INFINITY;
    int bl = 0;
    if (cnt > 0) {
        uint32_t da = a_scr + (uint32_t)excl * 16;
        bl = (int)(lds_u32_volatile(da + 4) >> 8);
        for (int i = 0; i < cnt; ++i, da += 16)
            take_min(__uint_as_float(lds_u32_volatile(da + 12)), (int)(lds_u32_volatile(da + 4) >> 8), best, bl);
    }
    if (full) {  // NaN/Inf, or candidates in more than two chunks: every centroid (rare)
        best = INFINITY;
        bl = 0;
        for (int j = 0; j < k; ++j) take_min(exact_pair(xt, row, C, j, d, cn[j]), j, best, bl);
    }
    // ---- publish: final label, cluster count, functional value; then release the row -----------------------------
    if (mine) {
        if (c.label_kind != HK_LABEL_NONE) store_label_tc(c.labels, c.label_kind, (int64_t)grow, bl);
        if (c.sums) {
            sts_u16(c.a_lab + (uint32_t)stage * (TM * 2) + (uint32_t)row * 2, (uint32_t)bl);
            asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(c.a_ecnt0 + (uint32_t)bl * 4) : "memory");
        }
        if (c.want_fv) {
            const float sq = sqrtf(best);
            rs.fv += (double)(sq * sq);
        }
        if (c.sums) {
            __threadfence_block();  // the label store is ordered before the completion the accumulator warps wait on
            mbar_complete_tx_a(c.b_lfull + (uint32_t)stage * 8, 1u);
        }
    }
    const unsigned nf = (unsigned)__popc(__ballot_sync(FULLM, full));
    rs.n_und += (unsigned)n;
    rs.n_pairs += (unsigned)total + nf * (unsigned)k;
    rs.n_full += nf;
    rs.n_batches += 1u;
    __syncwarp();  // the scratch is reused by the next batch
    return n;
}

// a refine warp's life: drain the ring until every epilogue warp has signed off and all allocated slots are consumed
__device__ __forceinline__ void refine_finish(const RefineCtx& rc, int lane, RefineStats& rs, uint32_t* stat_s, double* fv_slot) {
    if (rc.want_fv) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rs.fv += __shfl_xor_sync(0xffffffffu, rs.fv, o);
    }
    if (lane == 0 && rs.n_batches) {
        atomicAdd(stat_s + 0, rs.n_und);
        atomicAdd(stat_s + 1, rs.n_pairs);
        atomicAdd(stat_s + 2, rs.n_full);
        atomicAdd(stat_s + 3, rs.n_batches);
        *fv_slot = rs.fv;
    }
}
__device__ __forceinline__ void refine_role(const RefineCtx& rc, int lane, uint32_t a_scr, uint32_t* stat_s, double* fv_slot) {
    RefineStats rs;
    for (;;) {
        // cheap idle test first: three plain shared-memory loads per poll, no lock traffic while the ring is empty
        uint32_t done = 0, alloc = 0, head = 0;
        if (lane == 0) {
            done = lds_u32_volatile(rc.a_qalloc + 8);
            alloc = lds_u32_volatile(rc.a_qalloc);
            head = lds_u32_volatile(rc.a_qalloc + 4);
        }
        done = __shfl_sync(0xffffffffu, done, 0);
        alloc = __shfl_sync(0xffffffffu, alloc, 0);
        head = __shfl_sync(0xffffffffu, head, 0);
        if (alloc != head) {
            refine_batch(rc, lane, a_scr, rs);
            continue;
        }
        // `done` was read before `alloc`: an epilogue warp signs off after its last allocation
        if (done == (uint32_t)E_WARPS) break;
        __nanosleep(200);
    }
    refine_finish(rc, lane, rs, stat_s, fv_slot);
}

// SUMS: accumulate per-cluster sums (adds the accumulator warps); FQL2 = log2(d/4): lanes per row in the
// accumulator warps (d = 32, 64, 128 -> 3, 4, 5)
template <bool SUMS, int FQL2>
__global__ void __launch_bounds__((MISC_WARPS + E_WARPS + (SUMS ? A_WARPS_MAX : 0)) * 32, 1)
    lloyd_tc_kernel(const __grid_constant__ CUtensorMap xmap, const TcParams p) {
    extern __shared__ unsigned char smem_raw[];
    if (p.state != nullptr && p.state[0] != 0) return;  // uniform across the grid
    unsigned char* smem =
        reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const int d = p.d, k = p.k, nk = p.nk, S = p.S;
    const int nkb = d >> 5;
    const int NBUF = p.nbuf;
    const uint32_t a_stages = sbase + p.o_stages;
    const uint32_t a_B = sbase + p.o_B;
    const uint32_t a_lab = sbase + p.o_lab;
    float* cn = reinterpret_cast<float*>(smem + p.o_cn);
    // mbarriers: full[S] | empty[S] | lfull[S] | tfull[NBUF] | tempty[NBUF]   (16 slots per array)
    const uint32_t b_full = sbase + p.o_bars;
    const uint32_t b_empty = b_full + 16 * 8;
    const uint32_t b_lfull = b_full + 32 * 8;
    const uint32_t b_tfull = b_full + 48 * 8;
    const uint32_t b_tempty = b_full + 64 * 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + p.o_misc);
    float* cmax_s = reinterpret_cast<float*>(smem + p.o_misc + 16);   // max_j |c_j|
    float* cpmax_s = reinterpret_cast<float*>(smem + p.o_misc + 20);  // max_j |c_j - mu|
    int* force_exact_s = reinterpret_cast<int*>(smem + p.o_misc + 32);
    double* fvred = reinterpret_cast<double*>(smem + p.o_misc + 64);  // [E_WARPS]
    uint32_t* stat_s = reinterpret_cast<uint32_t*>(smem + p.o_misc + 256);  // [4] cold-path counters of this CTA
    float* mu_s = reinterpret_cast<float*>(smem + p.o_misc + 512);    // [d] mean of the centroids

    const int tid = threadIdx.x;
    // warp-uniform for the compiler: the role code below computes addresses and descriptors in uniform registers
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int lane = tid & 31;
    const uint32_t stage_bytes = (uint32_t)TM * d * 4;
    const int ntiles = p.num_tiles;
    const int xn_mode = p.bounds == nullptr
                            ? XN_COMPUTE
                            : (reinterpret_cast<const int*>(p.bounds)[p.num_tiles] != 0 ? XN_READ : XN_WRITE);

    // ---------------- one-time setup -------------------------------------------------------------------
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(reinterpret_cast<uint64_t*>(smem + p.o_bars) + s, 1);
            mbar_init(reinterpret_cast<uint64_t*>(smem + p.o_bars) + 16 + s, 4 + (SUMS ? 4 : 0));
            mbar_init(reinterpret_cast<uint64_t*>(smem + p.o_bars) + 32 + s, 4);
        }
        for (int b = 0; b < 16; ++b) mbar_init(reinterpret_cast<uint64_t*>(smem + p.o_bars) + 48 + b, 1);  // tfull ring
        for (int b = 0; b < NBUF; ++b) mbar_init(reinterpret_cast<uint64_t*>(smem + p.o_bars) + 64 + b, 4);
        mbar_fence_init();
        tma_prefetch_desc(&xmap);
        *cmax_s = 0.f;
        *cpmax_s = 0.f;
        *force_exact_s = 0;
        stat_s[0] = stat_s[1] = stat_s[2] = stat_s[3] = 0u;
        for (int w = 0; w < E_WARPS + R_WARPS; ++w) fvred[w] = 0.0;
    }
    if (warp == 2) tmem_alloc(tmem_slot, p.tmem_cols);
    for (int f = tid; f < d; f += blockDim.x) mu_s[f] = 0.f;
    for (int e = tid; e < p.ring_n * 4; e += blockDim.x) reinterpret_cast<uint32_t*>(smem + p.o_ring)[e] = 0u;  // no tag matches
    if (tid < 4) reinterpret_cast<uint32_t*>(smem + p.o_misc + 272)[tid] = 0u;  // q_alloc, q_head, q_done, q_lock
    // seed operand A_ext[r] = (1,1,1,0,...)
    for (int e = tid; e < 8 * 8; e += blockDim.x) {
        const int r = e >> 3, ch = e & 7;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ch == 0) v = make_float4(1.f, 1.f, 1.f, 0.f);
        *reinterpret_cast<float4*>(smem + p.o_Aext + sw128_off(8, r, ch << 2)) = v;
    }
    if (SUMS) {
        float* accz = reinterpret_cast<float*>(smem + p.o_acc);
        for (int i = tid; i < p.NA * (k + 1) * d; i += blockDim.x) accz[i] = 0.f;
        int* cz = reinterpret_cast<int*>(smem + p.o_cnt);
        for (int i = tid; i < E_WARPS * (2 * k + 1); i += blockDim.x) cz[i] = 0;
    }
    __syncthreads();
    // mu = mean of the centroids (any vector would do: it only tightens the filter bound, so the summation order
    // of the shared-memory float atomics does not matter)
    {
        const int nw = blockDim.x >> 5;
        const float invk = 1.f / (float)k;
        for (int f = lane; f < d; f += 32) {
            float s = 0.f;
            for (int j = warp; j < k; j += nw) s += p.C[(size_t)j * d + f];
            atomicAdd(mu_s + f, s * invk);
        }
    }
    __syncthreads();
    // |c_j|^2 (raw: the exact formula and the seed use it) and |c_j - mu|^2; B_ext[j] = three exact TF32 pieces of |c_j|^2
    for (int j = tid; j < nk; j += blockDim.x) {
        float s = 3.0e38f;  // padded centroids can never win
        if (j < k) {
            s = 0.f;
            float sp = 0.f;
            for (int f = 0; f < d; ++f) {
                const float c = p.C[(size_t)j * d + f];
                const float cp = c - mu_s[f];
                s = fmaf(c, c, s);
                sp = fmaf(cp, cp, sp);
            }
            if (!(s < INFINITY)) atomicExch(force_exact_s, 1);  // NaN/Inf centroid: exact path decides
            // s, sp >= 0: int order == float order
            atomicMax(reinterpret_cast<int*>(cmax_s), __float_as_int(sqrtf(s) * 1.000001f));
            atomicMax(reinterpret_cast<int*>(cpmax_s), __float_as_int(sqrtf(sp) * 1.000001f));
        }
        cn[j] = s;
        const float p1 = __uint_as_float(__float_as_uint(s) & 0xFFFFE000u);
        const float r1 = s - p1;
        const float p2 = __uint_as_float(__float_as_uint(r1) & 0xFFFFE000u);
        const float p3 = r1 - p2;
        for (int ch = 0; ch < 8; ++ch) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ch == 0) v = make_float4(p1, p2, p3, 0.f);
            *reinterpret_cast<float4*>(smem + p.o_Bext + sw128_off(nk, j, ch << 2)) = v;
        }
    }
    __syncthreads();
    const float cmax = *cmax_s;
    const bool force_exact = *force_exact_s != 0;
    // centre only when it helps (it always does unless the centroids straddle the origin already)
    const bool use_mu = !force_exact && (*cpmax_s < cmax);
    const float cpmax = use_mu ? *cpmax_s : cmax;
    // operand B = -2 * (c - mu): K-blocked, 128B-swizzled, rows >= k zero
    for (int e = tid; e < nk * (d >> 2); e += blockDim.x) {
        const int j = e / (d >> 2), f = (e - j * (d >> 2)) << 2;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < k) {
            v = *reinterpret_cast<const float4*>(p.C + (size_t)j * d + f);
            if (use_mu) {
                const float4 m = *reinterpret_cast<const float4*>(mu_s + f);
                v.x -= m.x;
                v.y -= m.y;
                v.z -= m.z;
                v.w -= m.w;
            }
            v.x *= -2.f;
            v.y *= -2.f;
            v.z *= -2.f;
            v.w *= -2.f;
        }
        *reinterpret_cast<float4*>(smem + p.o_B + sw128_off(nk, j, f)) = v;
    }
    fence_proxy_async();  // operands were written with st.shared, tcgen05.mma reads them through the async proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    TC_T(long long tw0 = 0, tw1 = 0, tw2 = 0; long long t0 = 0, t1 = 0, t2 = 0;)

    if (warp == 0) {
        // ================= TMA producer =================
        // the whole warp runs the loop (convergent, so addresses and descriptors live in uniform registers);
        // one elected lane issues
        int s = 0;
        uint32_t ph = 0;
        TC_T(const long long tstart = clock64();)
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            TC_T(t0 = clock64();)
            mbar_wait_a(b_empty + s * 8, ph ^ 1);
            TC_T(tw0 += clock64() - t0;)
            if (elect_one()) {
                TC_T(if (p.tl && blockIdx.x == 0) {
                    const int il = (tile - blockIdx.x) / gridDim.x;
                    if (il < 512) p.tl[il * 8 + 0] = clock64();
                })
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b_full + s * 8),
                             "r"(stage_bytes)
                             : "memory");
                for (int kb = 0; kb < nkb; ++kb)
                    asm volatile(
                        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
                        " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(a_stages + s * stage_bytes + kb * TM * 128),
                        "l"(&xmap), "r"(b_full + s * 8), "r"(kb * 32), "r"(tile * TM), "l"(kEvictFirst)
                        : "memory");
            }
            __syncwarp();
            if (++s == S) {
                s = 0;
                ph ^= 1;
            }
        }
        TC_T(tw1 = clock64() - tstart;)
    } else if (warp == 1 || (HK_TC_MI == 2 && warp == 3)) {
        // ================= MMA issuers (warp 1: even local tiles, warp 3: odd) =================
        // convergent warp, one elected lane issues: descriptors are computed in uniform registers, so each
        // tcgen05.mma is a single UTCHMMA instead of a per-lane R2UR waterfall.  One issuer warp is not enough:
        // two barrier waits + elect + 5 UTCHMMA + commit cost ~600 cycles per tile.
        const int mpar = warp == 1 ? 0 : 1;
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t idesc = umma_idesc_tf32(TM, nk);
        const uint64_t aext_d = umma_desc_k_sw128_bcast(sbase + p.o_Aext);
        const uint64_t bext_d = umma_desc_k_sw128(sbase + p.o_Bext);
        int s = mpar % S, b = mpar % NBUF;
        uint32_t ph = (uint32_t)((mpar / S) & 1), bph = (uint32_t)((mpar / NBUF) & 1);
        int il = mpar;  // local tile counter of this CTA
        for (int tile = blockIdx.x + mpar * gridDim.x; tile < ntiles; tile += HK_TC_MI * gridDim.x, il += HK_TC_MI) {
            TC_T(t0 = clock64();)
            mbar_wait_a(b_tempty + b * 8, bph ^ 1);
            TC_T(t1 = clock64();)
            mbar_wait_a(b_full + s * 8, ph);
            TC_T(t2 = clock64();)
            tc_fence_after();
            const uint32_t a_base = a_stages + s * stage_bytes;
            const uint32_t dcol = tmem_u + (uint32_t)(b * nk);
            if (elect_one()) {
                TC_T(if (p.tl && blockIdx.x == 0) {
                    const int il = (tile - blockIdx.x) / gridDim.x;
                    if (il < 512) p.tl[il * 8 + 1] = clock64();
                })
                umma_tf32(dcol, aext_d, bext_d, idesc, 0u);  // D = |c_j|^2
                for (int kb = 0; kb < nkb; ++kb) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t ad = umma_desc_k_sw128(a_base + kb * TM * 128 + ks * 32);
                        const uint64_t bd = umma_desc_k_sw128(a_B + kb * nk * 128 + ks * 32);
                        umma_tf32(dcol, ad, bd, idesc, 1u);  // D += x . (-2 c_j)
                    }
                }
                // "accumulator ready" goes to a ring of 16 barriers indexed by the local tile number, not to one barrier per
                // TMEM buffer: see the wait in the epilogue
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                 b_tfull + (uint32_t)(il & 15) * 8)
                             : "memory");
            }
            __syncwarp();
            TC_T(tw0 += t1 - t0; tw1 += t2 - t1; tw2 += clock64() - t2;)
            s += HK_TC_MI;
            if (s >= S) {
                s -= S;
                ph ^= 1;
            }
            b += HK_TC_MI;
            if (b >= NBUF) {
                b -= NBUF;
                bph ^= 1;
            }
        }
    } else if (warp >= E_FIRST && warp < E_FIRST + E_WARPS) {
        // ================= epilogue warps =================
        const int we = warp - E_FIRST;
        const int q = we & 3;   // TMEM lane quarter (== warp % 4)
        const int r = we >> 2;  // tile residue mod 4
        const int row = q * 32 + lane;
        const uint32_t tlane = __shfl_sync(0xffffffffu, tmem_base, 0) + ((uint32_t)(q * 32) << 16);
        const float beta2 = 2.f * 1.05f * 0.001953125f;
        const float gam = (float)(2 * d + 6) * 1.1920929e-7f;
        const float cmax2 = cmax * cmax + cpmax * cpmax;
        const uint32_t a_qalloc = sbase + p.o_misc + 272;  // ring control: allocated slots | consumed slots | warps done
        const uint32_t a_ring = sbase + p.o_ring;
        const uint32_t a_ecnt = sbase + p.o_cnt + (uint32_t)we * (uint32_t)(k * 4);
        const uint32_t a_hist = sbase + p.o_cnt + (uint32_t)(E_WARPS * k * 4) + (uint32_t)we * (uint32_t)((k + 1) * 4);
        const uint32_t a_mmax_q = sbase + p.o_snap + (uint32_t)q * 4;
        const uint32_t a_lab_row = a_lab + (uint32_t)row * 2;
        const int n32 = (int)p.n;  // n < 2^31 (tc_supported)
        const int label_kind = p.label_kind;
        const bool want_fv = p.fv_part != nullptr;
        double fv_acc = 0.0;
        int s = r % S;
        uint32_t ph = (uint32_t)((r / S) & 1);
        int b = r % NBUF;
        int i = r;  // local tile counter of this CTA
        for (int tile = blockIdx.x + r * gridDim.x; tile < ntiles; tile += ER * gridDim.x, i += ER) {
            TC_T(t0 = clock64();)
            const uint32_t xt = a_stages + s * stage_bytes;
            const int grow = tile * TM + row;
            const bool active = grow < n32;

            float xn, xs;  // |x|^2 and |x| of this row, or upper bounds for every row of the tile
            // One-bit phase parities and shared barriers: consecutive uses of a stage belong to DIFFERENT epilogue warps
            // (3 residues over S stages), so a warp can reach a wait while the PREVIOUS use is still incomplete (TMA loads
            // land out of order), see the parity of the phase before and walk on.  Hence
            //  * this warp waits on the stage's "full" barrier only when it reads the tile itself (first pass over a
            //    matrix / no row workspace), and then first on what the producer waits for before it issues THIS tile - the
            //    stage released by all consumers of the previous use - after which the parity is unambiguous;
            //  * "accumulator ready" is a ring of 16 barriers indexed by the tile number: the previous use of the same
            //    barrier lies 16 tiles back, and a tile that far back has been consumed completely before the producer
            //    could issue the tile this warp finished last (S <= 12 stages), so there the parity cannot alias.
            // With a cached bound the tile is not touched here: the labels published below are ordered after the TMA
            // writes through the MMA (it consumed the tile before it signalled "accumulator ready").
            if (xn_mode != XN_READ) {
                if (lane == 0) {
                    mbar_wait_a(b_empty + s * 8, ph ^ 1);
                    mbar_wait_a(b_full + s * 8, ph);
                }
                __syncwarp();
            }
            if (xn_mode == XN_READ) {
                xs = __ldg(p.bounds + tile);  // the cache holds an upper bound of max |x| over the tile
                xn = xs * xs;
            } else {
                xn = row_norm2(xt, row, d);
                xs = sqrtf(xn) * 1.0000002f;
                if (xn_mode == XN_WRITE) {
                    float wm = xs * 1.0000002f;  // wm*wm must not round below xn
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, o));
                    if (!(wm == wm)) wm = INFINITY;  // NaN rows: make the cached bound useless, not wrong
                    if (lane == 0) atomicMax(reinterpret_cast<int*>(p.bounds) + tile, __float_as_int(wm));
                }
            }
            const float E2 = 2.002f * (beta2 * xs * cpmax + gam * (xn + cmax2));

            warp_wait(b_tfull + (uint32_t)(i & 15) * 8, (uint32_t)((i >> 4) & 1), lane);  // s_j = |c_j|^2 - 2 x.c'_j (TF32)
            TC_T(t1 = clock64();
                 if (p.tl && blockIdx.x == 0 && q == 0 && lane == 0 && i < 512) p.tl[i * 8 + 2] = clock64();)
            tc_fence_after();
            const uint32_t taddr = tlane + (uint32_t)(b * nk);
            // one sweep over the accumulator, 32 columns in registers at a time.  Per chunk: its minimum m_c and
            // the mask of columns below m_c + 2E.  The chunk holding the global minimum has the right mask; the
            // other chunks hold no candidate iff their minimum is at least 2E above it (else the row is
            // undecided anyway: at least one candidate per such chunk).
            float m_best = INFINITY, m_second = INFINITY;
            unsigned mk_best = 0;
            int c_best = 0;
            for (int c0 = 0; c0 < nk; c0 += 32) {
                uint32_t a[32];
                tmem_ld32(taddr + (uint32_t)c0, a);
                tmem_wait_ld();
                const float mc = min32(a);
                const unsigned mk = below_mask32(a, mc + E2);
                if (mc < m_best) {
                    m_second = m_best;
                    m_best = mc;
                    mk_best = mk;
                    c_best = c0;
                } else {
                    m_second = fminf(m_second, mc);
                }
            }
            // NaN minima compare false everywhere: the row is then undecided and the exact path takes it
            const float thr = m_best + E2;
            const bool decided = (m_second >= thr) && (__popc(mk_best) == 1) && !force_exact && (E2 < INFINITY);
            int lab = c_best + __ffs(mk_best) - 1;
            if (!active) lab = k;
            const bool cold = active && !decided;
            int n_cold = 0;
            if (__builtin_expect(__any_sync(0xffffffffu, cold), 0)) {
                // undecided rows (near-ties within the TF32 bound, NaN/Inf): hand them to the refine warp.  Candidate set
                // of a row = columns below min + 2E: the mask of the best chunk is exact; rows with candidates elsewhere
                // get theirs from a second sweep over the accumulator (still held).  Rows with candidates in more than
                // two chunks, and NaN/Inf, go through the all-centroid formula.  (Keeping the runner-up chunk's mask in
                // the first sweep instead was measured: unstructured data -15 %, decided-only data +2 %.)
                bool full = cold && (force_exact || !(thr < INFINITY) || !(fabsf(m_best) < INFINITY) || mk_best == 0u);
                const bool und = cold && !full;
                unsigned mlo = und ? mk_best : 0u, mhi = 0u;
                int clo = c_best, chi = 0;
                if (nk > 32 && __any_sync(0xffffffffu, und && m_second < thr)) {
                    int nch = 0;
                    unsigned lo = 0u, hi = 0u;
                    int cl = 0, ch = 0;
#pragma unroll 1
                    for (int c0 = 0; c0 < nk; c0 += 32) {
                        uint32_t a[32];
                        tmem_ld32(taddr + (uint32_t)c0, a);
                        tmem_wait_ld();
                        const unsigned mk = und ? below_mask32(a, thr) : 0u;
                        if (mk) {
                            if (nch == 0) {
                                lo = mk;
                                cl = c0;
                            } else if (nch == 1) {
                                hi = mk;
                                ch = c0;
                            }
                            ++nch;
                        }
                    }
                    if (und) {
                        const bool over = nch > 2 || nch == 0;
                        full = over;
                        mlo = over ? 0u : lo;
                        mhi = over ? 0u : hi;
                        clo = cl;
                        chi = ch;
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_a(b_tempty + b * 8);  // accumulator b may be overwritten
                const unsigned pm = __ballot_sync(0xffffffffu, cold);
                n_cold = __popc(pm);
                uint32_t slot0 = 0;
                if (lane == 0) {
                    // the tile's labels are complete only when the refine warps have released these rows
                    if (SUMS) mbar_expect_tx_a(b_lfull + s * 8, (uint32_t)n_cold);
                    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(slot0) : "r"(a_qalloc), "r"(n_cold) : "memory");
                }
                slot0 = __shfl_sync(0xffffffffu, slot0, 0);
                if (cold) {
                    const uint32_t idx = slot0 + (uint32_t)__popc(pm & lanemask_lt());
                    while ((int32_t)(idx - lds_u32_volatile(a_qalloc + 4)) >= p.ring_n) __nanosleep(32);  // ring full
                    // one 16-byte store publishes the entry (tag and payload land together)
                    const uint32_t w0 = ((idx + 1u) & 0xffffu) | ((uint32_t)s << 16) | ((uint32_t)(clo >> 5) << 20) |
                                        ((uint32_t)(chi >> 5) << 23) | (full ? (1u << 26) : 0u);
                    asm volatile("st.volatile.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_ring + (idx & (uint32_t)(p.ring_n - 1)) * 16),
                                 "r"(w0), "r"((uint32_t)grow), "r"(full ? 0u : mlo), "r"(full ? 0u : mhi)
                                 : "memory");
                    lab = k;  // no label yet: the refine warp stores the final one (and counts the row)
                }
            } else {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_a(b_tempty + b * 8);  // accumulator b may be overwritten
            }
            TC_T(t2 = clock64();)
            if (!SUMS && want_fv && active && !cold) {
                // functional value of a decided row: exact distance to its centroid
                mbar_wait_a(b_empty + s * 8, ph ^ 1);  // (see above) then the tile itself
                mbar_wait_a(b_full + s * 8, ph);
                const float sq = sqrtf(exact_pair_inl(xt, row, p.C, lab, d, cn[lab]));
                fv_acc += (double)(sq * sq);
            }
            if (label_kind != HK_LABEL_NONE && active && !cold) store_label_tc(p.labels, label_kind, (int64_t)grow, lab);
            if (SUMS) {
                // hand the labels to the accumulator warp of this lane quarter (rows past the end carry label k; the
                // label of an undecided row is stored by the refine warp, possibly before this point)
                if (!cold) sts_u16(a_lab_row + s * (TM * 2), (uint32_t)lab);
                __syncwarp();
                if (lane == 0) mbar_arrive_a(b_lfull + s * 8);  // the accumulator warp can start on these rows
                TC_T(if (p.tl && blockIdx.x == 0 && q == 0 && lane == 0 && i < 512) p.tl[i * 8 + 3] = clock64();)
                // rows per label among these 32 rows, through a private per-warp histogram in shared memory
                // (integer atomics: order independent): cluster counts and, for the accumulator warp's flush
                // rule, the largest multiplicity
                const uint32_t ha = a_hist + (uint32_t)lab * 4;
                int old;
                asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(ha) : "memory");
                if (lab < k) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a_ecnt + (uint32_t)lab * 4) : "memory");
                // (undecided rows sit in the dummy bucket k: each could still join the largest group)
                const int mm = min(32, __reduce_max_sync(0xffffffffu, lab < k ? old + 1 : 0) + n_cold);
                sts_s32(ha, 0);
                if (lane == 0) {
                    // 32-deep ring indexed by the local tile number, read by the accumulator warp when it starts its
                    // next own tile.  No barrier: the value carries the tile number as a tag, a reader that does not
                    // find its tag (this store not visible yet, practically never) assumes the worst case instead.
                    sts_s32(a_mmax_q + (uint32_t)(i & 31) * 16, (i << 6) | mm);
                    mbar_arrive_a(b_empty + s * 8);
                    TC_T(if (p.tl && blockIdx.x == 0 && q == 0 && i < 512) p.tl[i * 8 + 4] = clock64();)
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
            b += ER;
            while (b >= NBUF) b -= NBUF;
            TC_T(tw0 += t1 - t0; tw1 += t2 - t1; tw2 += clock64() - t2;)
        }
        if (want_fv) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) fv_acc += __shfl_xor_sync(0xffffffffu, fv_acc, o);
            if (lane == 0) fvred[we] = fv_acc;
        }
        __syncwarp();
        if (lane == 0) {  // no more entries from this warp
            __threadfence_block();
            asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a_qalloc + 8) : "memory");
        }
    } else if (warp == 2) {
        // ================= refine warp 0 =================
        const RefineCtx rc{k, d, p.C, p.labels, p.label_kind, SUMS, p.fv_part != nullptr, sbase + p.o_ring,
                           sbase + p.o_misc + 272, a_stages, a_lab, b_lfull, sbase + p.o_cnt, cn, p.ring_n, p.maxp};
        refine_role(rc, lane, sbase + p.o_ring + (uint32_t)p.ring_n * 16, stat_s, fvred + E_WARPS);
    } else if (SUMS && warp >= A_FIRST && warp < A_FIRST + p.NA) {
        // ================= accumulator warps =================
        const int a = warp - A_FIRST;
        const int q = a & 3;         // lane quarter of the tile this warp accumulates
        const int res = a >> 2;      // tile residue
        const int nres = p.NA >> 2;  // 1 or 2
        constexpr int FQ = 1 << FQL2;         // lanes per row (feature quads): 8, 16, 32
        constexpr int RPI = 32 / FQ;          // rows per step: 4, 2, 1
        constexpr int NIT = 32 / RPI;         // steps per 32-row quarter: 8, 16, 32
        constexpr int BATCH = NIT < 8 ? NIT : 8;
        const int g = lane >> FQL2;           // row slot inside a step
        const int fq = lane & (FQ - 1);
        const uint32_t rowbytes = (uint32_t)d * 4;
        const uint32_t a_mmax_q = sbase + p.o_snap + (uint32_t)q * 4;
        const uint32_t a_lab_lane = a_lab + (uint32_t)(q * 32 + lane) * 2;
        const uint32_t a_lab_grp = a_lab + (uint32_t)(q * 32 + g) * 2;  // label of row slot g of step 0
        int run_max = 0;  // upper bound on the fp32 adds any accumulator row has taken since the last flush
        const uint32_t acc_w = sbase + p.o_acc + (uint32_t)a * (uint32_t)(k + 1) * rowbytes;
        const uint32_t acc_l = acc_w + fq * 16;
        // this lane's 16 bytes inside row (q*32 + g) of a stage; step `it` adds it*RPI*128 bytes and, because
        // the 128-byte swizzle depends on (row & 7), toggles the chunk index by ((it*RPI) & 7)
        const uint32_t x_lane = (uint32_t)((fq >> 3) * TM * 128 + (q * 32 + g) * 128);
        const uint32_t fq7 = (uint32_t)(fq & 7);
        double* gslot = p.fsum + ((size_t)blockIdx.x * p.NA + a) * (size_t)(k * d);
        bool first_flush = true;

        auto flush = [&]() {
            // widen the private fp32 sums into this warp's fp64 slot (plain read-modify-write: sole owner), clear
            const int nq = (k * d) >> 2;
            for (int e = lane; e < nq; e += 32) {
                const float4 v = lds_f4(acc_w + e * 16);
                sts_f4_nc(acc_w + e * 16, make_float4(0.f, 0.f, 0.f, 0.f));
                double2* gp = reinterpret_cast<double2*>(gslot + (size_t)e * 4);
                double2 lo = make_double2(0.0, 0.0), hi = make_double2(0.0, 0.0);
                if (!first_flush) {
                    lo = gp[0];
                    hi = gp[1];
                }
                lo.x += (double)v.x;
                lo.y += (double)v.y;
                hi.x += (double)v.z;
                hi.y += (double)v.w;
                gp[0] = lo;
                gp[1] = hi;
            }
            first_flush = false;
            run_max = 0;
        };

        int s = res % S;
        uint32_t ph = (uint32_t)((res / S) & 1);
        int i = res;  // local tile counter of this CTA
        for (int tile = blockIdx.x + res * gridDim.x; tile < ntiles; tile += nres * gridDim.x, i += nres) {
            TC_T(t0 = clock64();)
            // labels of all four lane quarters published; the epilogue warps observed the x tile before publishing,
            // so this wait also orders the reads of the tile below
            warp_wait(b_lfull + s * 8, ph, lane);
            TC_T(t1 = clock64(); const int ila = (tile - blockIdx.x) / gridDim.x;
                 if (p.tl && blockIdx.x == 0 && q == 0 && lane == 0 && ila < 512) p.tl[ila * 8 + 5] = clock64();)
            const uint32_t xq = a_stages + s * stage_bytes + x_lane;
            const uint32_t mylab = lds_u16(a_lab_lane + s * (TM * 2));
            // label collisions inside a step (rows that would hit the same accumulator row), for all steps
            unsigned coll = 0;
            if (RPI > 1) {
                bool c = false;
#pragma unroll
                for (int x = 1; x < RPI; ++x) c |= (__shfl_xor_sync(0xffffffffu, mylab, x) == mylab);
                coll = __ballot_sync(0xffffffffu, c);
            }
#pragma unroll 1
            for (int it0 = 0; it0 < NIT; it0 += BATCH) {
                // fetch the rows and the accumulator addresses of a batch of steps, then run the
                // load-add-store chains (steps may share a label, so the chains stay in order)
                float4 xr[BATCH];
                uint32_t aa[BATCH];
#pragma unroll
                for (int j = 0; j < BATCH; ++j) {
                    const int rs = (it0 + j) * RPI;  // first row of the step inside the quarter
                    // straight from shared memory (an independent load) rather than a shuffle of mylab
                    const uint32_t l = lds_u16(a_lab_grp + s * (TM * 2) + (uint32_t)(rs * 2));
                    aa[j] = acc_l + l * rowbytes;
                    xr[j] = lds_f4(xq + (uint32_t)(rs << 7) + (((fq7 ^ (uint32_t)((rs + g) & 7))) << 4));
                }
                const unsigned cb = RPI > 1 ? (coll >> (it0 * RPI)) : 0u;
#pragma unroll
                for (int j = 0; j < BATCH; ++j) {
                    if (RPI == 1 || ((cb >> (j * RPI)) & ((1u << RPI) - 1u)) == 0u) {
                        // common case: no two rows of this step share a label
                        float4 v = lds_f4(aa[j]);
                        v.x += xr[j].x;
                        v.y += xr[j].y;
                        v.z += xr[j].z;
                        v.w += xr[j].w;
                        sts_f4_nc(aa[j], v);
                    } else {
#pragma unroll 1
                        for (int gg = 0; gg < RPI; ++gg) {  // one row slot at a time
                            if (g == gg) {
                                float4 v = lds_f4(aa[j]);
                                v.x += xr[j].x;
                                v.y += xr[j].y;
                                v.z += xr[j].z;
                                v.w += xr[j].w;
                                sts_f4_nc(aa[j], v);
                            }
                            __syncwarp();
                        }
                    }
                }
            }
            // multiplicity bound of this warp's previous tile: tagged ring slot, worst case (32 rows, one label) if the
            // epilogue warp's store is not visible yet
            if (i >= nres) {
                const int v = lds_s32(a_mmax_q + (uint32_t)((i - nres) & 31) * 16);
                run_max += (v >> 6) == i - nres ? (v & 63) : 32;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_a(b_empty + s * 8);
            TC_T(if (p.tl && blockIdx.x == 0 && q == 0 && lane == 0 && ila < 512) p.tl[ila * 8 + 6] = clock64();
                 t2 = clock64();)
            // widen before any accumulator row can have taken more than 96 + 64 fp32 adds: run_max lags by this tile
            // and the previous one (at most 32 adds each, ~3 in practice), so a partial sum carries at most
            // 160 * 2^-24 = 1e-5 relative rounding error in the worst case (typically 100x less).  Timing independent
            // unless a ring tag is missed, which only makes the flush earlier.  (Each flush costs ~2800 cycles of
            // global read-modify-write; a threshold of 48 costs 2.5 % of the iteration.)
            if (run_max >= 96) flush();
            TC_T(tw0 += t1 - t0; tw1 += t2 - t1; tw2 += clock64() - t2;)
            s += nres;
            if (s >= S) {
                s -= S;
                ph ^= 1;
            }
        }
        flush();
    }
    TC_T(if (p.dbg && lane == 0) {
        unsigned long long* o = p.dbg + ((size_t)blockIdx.x * 32 + warp) * 8;
        o[0] = (unsigned long long)tw0;
        o[1] = (unsigned long long)tw1;
        o[2] = (unsigned long long)tw2;
    })

    // ---------------- teardown ---------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
    if (SUMS) {
        // fold the NA per-warp fp64 slots of this CTA into its first slot (fixed order), so that the
        // cross-CTA reduction reads one slot per CTA
        const int kd = k * d;
        double* base = p.fsum + (size_t)blockIdx.x * p.NA * (size_t)kd;
        for (int e = tid; e < kd; e += blockDim.x) {
            double t = base[e];
            for (int w = 1; w < p.NA; ++w) t += base[(size_t)w * kd + e];
            base[e] = t;
        }
        const int* ce = reinterpret_cast<const int*>(smem + p.o_cnt);
        for (int c = tid; c < k; c += blockDim.x) {
            int t = 0;
            for (int w = 0; w < E_WARPS; ++w) t += ce[w * k + c];
            p.fcnt[(size_t)blockIdx.x * k + c] = (double)t;
        }
    }
    if (tid == 0) {
        if (p.fv_part != nullptr) {
            double t = 0.0;
            for (int w = 0; w < E_WARPS + R_WARPS; ++w) t += fvred[w];  // epilogue warps + refine warps
            p.fv_part[blockIdx.x] = t;
        }
        if (xn_mode == XN_WRITE && blockIdx.x == 0) reinterpret_cast<int*>(p.bounds)[p.num_tiles] = 1;
        if (p.stats != nullptr) {
            for (int c = 0; c < 4; ++c)
                if (stat_s[c]) atomicAdd(p.stats + c, (unsigned long long)stat_s[c]);
            if (blockIdx.x == 0) {
                atomicAdd(p.stats + 4, (unsigned long long)p.n);
                atomicAdd(p.stats + 5, 1ull);
            }
        }
    }
}

// partials[c][0..d) = sum over accumulator slots, partials[c][d] = sum over CTAs of the counts.
// Block = 32 outputs x 8 slot groups; every partial sum and the final 8-way combine run in a fixed order.
__global__ void __launch_bounds__(256) reduce_tc_kernel(const double* __restrict__ fsum, const double* __restrict__ fcnt,
                                                        int nslots, int slot_stride, int nblocks, int k, int d,
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
            for (int b = grp * per; b < b1; ++b) t += fsum[(size_t)b * slot_stride * k * d + (size_t)c * d + f];
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
__global__ void reduce_scalar_tc_kernel(const double* __restrict__ v, int n, double* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double t = 0.0;
        for (int b = 0; b < n; ++b) t += v[b];
        *out = t;
    }
}

struct TcPlan {
    int S, nk, NA, nbuf, fql2, ring_n, maxp;
    uint32_t tmem_cols;
    size_t smem;
    bool ok;
};

TcPlan plan_tc(const Handle* h, int d, int k, bool sums) {
    TcPlan pl{};
    pl.ok = false;
    pl.nk = (k + 31) / 32 * 32;
    pl.fql2 = d == 32 ? 3 : (d == 64 ? 4 : 5);
    // private fp32 accumulators: 8 warps when they fit in ~72 KB, else 4
    pl.NA = 8;
    if ((size_t)8 * (k + 1) * d * 4 > 80 * 1024) pl.NA = 4;
    if (sums && (size_t)pl.NA * (k + 1) * d * 4 > 100 * 1024) return pl;
    // TMEM: nbuf distance buffers of nk columns each (at most 8, 512 columns in all)
    // (an EVEN number: the two MMA warps alternate over the buffers and wait on their "empty" barriers with one-bit
    // parities, so consecutive uses of a buffer must belong to the same warp - see the note at the epilogue's wait)
    int nb = 512 / pl.nk;
    pl.nbuf = nb > 8 ? 8 : nb;
    if (pl.nbuf > 2) pl.nbuf &= ~1;
    uint32_t cols = 32;
    while (cols < (uint32_t)(pl.nbuf * pl.nk)) cols <<= 1;
    pl.tmem_cols = cols;
    const size_t budget = (size_t)h->smem_optin;
    // full-size refine ring first; shapes that are short of shared memory get a 32-entry ring rather than fewer than 4 stages
    for (int small = 0; small < 2; ++small) {
        pl.ring_n = small ? 32 : RING_MAX;
        pl.maxp = small ? 32 : MAXP_MAX;
        for (int S = 12; S >= 4; S -= 2) {  // even: the MMA / accumulator warps alternate over the stages
            TcLayout L = tc_layout(d, k, pl.nk, S, pl.NA, sums, pl.ring_n, pl.maxp);
            if (L.total <= budget) {
                pl.S = S;
                pl.smem = L.total;
                pl.ok = true;
                return pl;
            }
        }
    }
    return pl;
}

template <bool SUMS, int FQL2>
int launch_inst(Handle* h, const CUtensorMap& map, TcParams& p, size_t smem, int grid, cudaStream_t st) {
    auto kern = lloyd_tc_kernel<SUMS, FQL2>;
    HK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    prof_begin(h, st);
    kern<<<grid, (MISC_WARPS + E_WARPS + (SUMS ? p.NA : 0)) * 32, smem, st>>>(map, p);
    prof_end(h, st);
    HK_CUDA(cudaGetLastError());
    h->launches++;
    return 0;
}

}  // namespace

bool tc_supported(const Handle* h, const LloydArgs& a) {
    if (a.dtype != HK_F32) return false;
    if (a.d != 32 && a.d != 64 && a.d != 128) return false;
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
    TcParams p{};
    p.n = a.n;
    p.d = a.d;
    p.k = a.k;
    p.nk = pl.nk;
    p.C = reinterpret_cast<const float*>(a.C);
    p.labels = a.labels;
    p.label_kind = a.labels ? a.label_kind : HK_LABEL_NONE;
    p.S = pl.S;
    p.nbuf = pl.nbuf;
    p.NA = pl.NA;
    p.num_tiles = (int)((a.n + TM - 1) / TM);
    p.state = a.state;
    p.tmem_cols = pl.tmem_cols;
    {
        const TcLayout L = tc_layout(a.d, a.k, pl.nk, pl.S, pl.NA, sums, pl.ring_n, pl.maxp);
        p.ring_n = pl.ring_n;
        p.maxp = pl.maxp;
        p.o_stages = (uint32_t)L.stages;
        p.o_B = (uint32_t)L.B;
        p.o_Aext = (uint32_t)L.Aext;
        p.o_Bext = (uint32_t)L.Bext;
        p.o_cn = (uint32_t)L.cn;
        p.o_acc = (uint32_t)L.acc;
        p.o_lab = (uint32_t)L.lab;
        p.o_cnt = (uint32_t)L.cnt;
        p.o_snap = (uint32_t)L.snap;
        p.o_bars = (uint32_t)L.bars;
        p.o_misc = (uint32_t)L.misc;
        p.o_ring = (uint32_t)L.ring;
    }

    // per-tile |x| bound cache: caller-owned and tied to the content of X (see hk_row_ws_bytes); none -> per-row
    // norms are recomputed in every pass
    p.bounds = nullptr;
    if (a.row_ws != nullptr) {
        const size_t need = ((size_t)p.num_tiles + 4) * sizeof(float);
        if ((size_t)a.row_ws_bytes < need || (reinterpret_cast<uintptr_t>(a.row_ws) & 3) != 0) {
            set_error("lloyd_tc: row workspace too small or misaligned (%lld bytes given, %zu needed)",
                      (long long)a.row_ws_bytes, need);
            return -1;
        }
        p.bounds = reinterpret_cast<float*>(a.row_ws);
    }
    rc = ensure_stats(h);
    if (rc) return rc;
    p.stats = h->stats;

    int grid = h->num_sms;
    if (grid > p.num_tiles) grid = p.num_tiles;
    const int nslots = grid * pl.NA;
    const size_t kd = (size_t)a.k * a.d;
    rc = ensure_part(h, ((size_t)nslots * kd + (size_t)nslots * a.k + grid) * sizeof(double));
    if (rc) return rc;
    p.fsum = sums ? h->part : nullptr;
    p.fcnt = h->part + (size_t)nslots * kd;
    p.fv_part = a.fv_out ? h->part + (size_t)nslots * kd + (size_t)nslots * a.k : nullptr;

#ifdef HK_TC_TIMING
    static const bool dbg_on = getenv("HK_TC_DEBUG") != nullptr;  // role timing needs a -DHK_TC_TIMING build
#else
    const bool dbg_on = false;
#endif
    unsigned long long* dbg = nullptr;
    if (dbg_on) {
        HK_CUDA(cudaMalloc(&dbg, (size_t)grid * 32 * 8 * sizeof(unsigned long long)));
        HK_CUDA(cudaMemsetAsync(dbg, 0, (size_t)grid * 32 * 8 * sizeof(unsigned long long), a.stream));
    }
    p.dbg = dbg;
    unsigned long long* tl = nullptr;
    if (dbg_on) {
        HK_CUDA(cudaMalloc(&tl, 512 * 8 * sizeof(unsigned long long)));
        HK_CUDA(cudaMemsetAsync(tl, 0, 512 * 8 * sizeof(unsigned long long), a.stream));
    }
    p.tl = tl;

    char name[112];
    snprintf(name, sizeof(name), "tc<f32,d=%d,k=%d,S=%d,nbuf=%d,NA=%d,%s,%s>", a.d, a.k, pl.S, pl.nbuf, pl.NA,
             sums ? "sums" : "assign", p.bounds ? "xn-ws" : "xn-row");
    h->variant = name;

    if (!sums) {
        rc = launch_inst<false, 3>(h, map, p, pl.smem, grid, a.stream);
    } else {
        switch (pl.fql2) {
            case 3: rc = launch_inst<true, 3>(h, map, p, pl.smem, grid, a.stream); break;
            case 4: rc = launch_inst<true, 4>(h, map, p, pl.smem, grid, a.stream); break;
            default: rc = launch_inst<true, 5>(h, map, p, pl.smem, grid, a.stream); break;
        }
    }
    if (rc) return rc;
    if (dbg_on) {
        std::vector<unsigned long long> hbuf((size_t)grid * 32 * 8);
        HK_CUDA(cudaStreamSynchronize(a.stream));
        HK_CUDA(cudaMemcpy(hbuf.data(), dbg, hbuf.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        cudaFree(dbg);
        double acc[32][8] = {};
        for (int b = 0; b < grid; ++b)
            for (int w = 0; w < 32; ++w)
                for (int j = 0; j < 8; ++j) acc[w][j] += (double)hbuf[((size_t)b * 32 + w) * 8 + j] / grid;
        {
            std::vector<unsigned long long> ht(512 * 8);
            HK_CUDA(cudaMemcpy(ht.data(), tl, ht.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            cudaFree(tl);
            fprintf(stderr, "[hk tc timeline] CTA 0, cycles relative to the TMA issue of each tile: full@mma tfull@E lfull Edone lfull@A Adone | issue-to-issue\n");
            for (int il = 200; il < 216; ++il) {
                const unsigned long long* e = &ht[il * 8];
                if (!e[0]) continue;
                fprintf(stderr, "  tile %3d: %6lld %6lld %6lld %6lld %6lld %6lld | %6lld\n", il, (long long)(e[1] - e[0]),
                        (long long)(e[2] - e[0]), (long long)(e[3] - e[0]), (long long)(e[4] - e[0]), (long long)(e[5] - e[0]),
                        (long long)(e[6] - e[0]), (long long)(e[0] - ht[(il - 1) * 8]));
            }
        }
        const double tiles_cta = (double)p.num_tiles / grid;
        fprintf(stderr, "[hk tc debug] %s tiles/CTA %.0f; mean cycles per CTA (per tile in brackets)\n", name, tiles_cta);
        fprintf(stderr, "  producer: wait-empty %.0f [%.0f]  total %.0f [%.0f]\n", acc[0][0], acc[0][0] / tiles_cta, acc[0][1],
                acc[0][1] / tiles_cta);
        fprintf(stderr, "  mma warp 1: wait-tempty %.0f  wait-full %.0f  issue+commit %.0f (per own tile)\n",
                acc[1][0] / (tiles_cta / 2), acc[1][1] / (tiles_cta / 2), acc[1][2] / (tiles_cta / 2));
        for (int w = 4; w < 20; w += 5)
            fprintf(stderr, "  E warp %2d: wait %.0f [%.0f per own tile]  tmem+scan %.0f [%.0f]  label+publish %.0f [%.0f]\n", w,
                    acc[w][0], acc[w][0] / (tiles_cta / 4), acc[w][1], acc[w][1] / (tiles_cta / 4), acc[w][2],
                    acc[w][2] / (tiles_cta / 4));
        const double an = tiles_cta / (p.NA / 4);
        for (int w = 20; w < 20 + p.NA; w += 3)
            fprintf(stderr, "  A warp %2d: wait %.0f [%.0f per own tile]  rows %.0f [%.0f]  flush %.0f [%.0f]\n", w, acc[w][0],
                    acc[w][0] / an, acc[w][1], acc[w][1] / an, acc[w][2], acc[w][2] / an);
    }
    if (sums && a.slots != nullptr) {
        // the caller's finish kernel reduces the slots (one launch for reduce + exchange + finalize)
        a.slots->fsum = p.fsum;
        a.slots->fcnt = p.fcnt;
        a.slots->nslots = grid;
        a.slots->slot_stride = pl.NA;
        a.slots->nblocks = grid;
    } else if (sums) {
        const int len = a.k * (a.d + 1);
        reduce_tc_kernel<<<(len + 31) / 32, 256, 0, a.stream>>>(p.fsum, p.fcnt, grid, pl.NA, grid, a.k, a.d, a.partials,
                                                                a.state);
        HK_CUDA(cudaGetLastError());
        h->launches++;
    }
    if (a.fv_out) {
        reduce_scalar_tc_kernel<<<1, 32, 0, a.stream>>>(p.fv_part, grid, a.fv_out);
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
