// extern "C" surface of libhkmeans.so — see include/hkmeans.h for the reference interface each
// entry point replaces.
#include <stdarg.h>
#include <string.h>

#include "hk_common.cuh"

namespace hk {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static int grow(void** p, size_t* have, size_t need) {
    if (*have >= need) return 0;
    if (*p) HK_CUDA(cudaFree(*p));
    *p = nullptr;
    *have = 0;
    size_t sz = need < (1u << 20) ? (1u << 20) : need;
    HK_CUDA(cudaMalloc(p, sz));
    *have = sz;
    return 0;
}
int ensure_part(Handle* h, size_t bytes) { return grow((void**)&h->part, &h->part_bytes, bytes); }
int ensure_red(Handle* h, size_t bytes) { return grow((void**)&h->red, &h->red_bytes, bytes); }
int ensure_stats(Handle* h) {
    if (h->stats) return 0;
    HK_CUDA(cudaMalloc(&h->stats, 8 * sizeof(unsigned long long)));
    HK_CUDA(cudaMemset(h->stats, 0, 8 * sizeof(unsigned long long)));
    return 0;
}

int comm_unique_id(void* id128);
int comm_init(Handle* h, int nranks, int rank, const void* id128);
int comm_destroy(Handle* h);
int comm_peer_export(Handle* h, int64_t cap, void* handle64);
int comm_peer_import(Handle* h, const void* handles);

static int ensure_fin_sync(Handle* h) {
    if (h->fin_sync) return 0;
    HK_CUDA(cudaMalloc(&h->fin_sync, 2 * sizeof(unsigned int)));
    HK_CUDA(cudaMemset(h->fin_sync, 0, 2 * sizeof(unsigned int)));
    return 0;
}

static int check_common(const char* fn, hk_handle_t h, const void* X, int64_t n, int d, int64_t ldx,
                        int dtype, const void* C, int k) {
    HK_ARG(h != nullptr, "%s: null handle", fn);
    HK_ARG(dtype == HK_F32 || dtype == HK_F64, "%s: dtype must be HK_F32 or HK_F64", fn);
    HK_ARG(n >= 0, "%s: n_local < 0", fn);
    HK_ARG(d >= 1, "%s: d must be >= 1", fn);
    HK_ARG(k >= 1, "%s: k must be >= 1", fn);
    HK_ARG(ldx >= d, "%s: ldx (%lld) < d (%d)", fn, (long long)ldx, d);
    HK_ARG(n == 0 || X != nullptr, "%s: X is null", fn);
    HK_ARG(C != nullptr, "%s: C is null", fn);
    return 0;
}

static int run_pass(Handle* h, const LloydArgs& a) {
    int path = a.path;
    if (path == HK_PATH_AUTO) {
        if (tc_supported(h, a)) {
            path = HK_PATH_TC;
        } else if (bigk_supported(h, a) && !row128_supported(h, a)) {
            return launch_lloyd_bigk(h, a);  // large k: distances on the tensor cores, sort-based sums
        } else {
            path = HK_PATH_SIMT;
        }
    }
    if (path == HK_PATH_TC) {
        if (!tc_supported(h, a)) {
            set_error("tensor-core path does not support dtype=%d d=%d k=%d ldx=%lld", a.dtype, a.d, a.k,
                      (long long)a.ldx);
            return -2;
        }
        return launch_lloyd_tc(h, a);
    }
    if (path == HK_PATH_SIMT && dmma_supported(h, a)) return launch_lloyd_dmma(h, a);
    if ((path == HK_PATH_SIMT || path == HK_PATH_ROW128) && row128_supported(h, a)) return launch_lloyd_row128(h, a);
    if (path == HK_PATH_ROW128) {
        set_error("128-byte-row path does not support dtype=%d d=%d k=%d", a.dtype, a.d, a.k);
        return -2;
    }
    return launch_lloyd_simt(h, a);
}

}  // namespace hk

using namespace hk;

extern "C" {

int hk_version(void) { return 100; }
const char* hk_last_error(void) { return g_err; }

int hk_create(hk_handle_t* out, int device) {
    HK_ARG(out != nullptr, "hk_create: out is null");
    int ndev = 0;
    HK_CUDA(cudaGetDeviceCount(&ndev));
    HK_ARG(device >= 0 && device < ndev, "hk_create: device %d out of range (%d devices)", device, ndev);
    HK_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    HK_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("hk_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                  prop.major, prop.minor);
        return -4;
    }
    Handle* h = new Handle();
    h->device = device;
    h->num_sms = prop.multiProcessorCount;
    h->smem_optin = (int)prop.sharedMemPerBlockOptin;
    *out = reinterpret_cast<hk_handle_t>(h);
    return 0;
}

int hk_destroy(hk_handle_t hh) {
    if (!hh) return 0;
    Handle* h = reinterpret_cast<Handle*>(hh);
    comm_destroy(h);
    cudaSetDevice(h->device);
    if (h->part) cudaFree(h->part);
    if (h->red) cudaFree(h->red);
    if (h->tc_scratch) cudaFree(h->tc_scratch);
    if (h->stats) cudaFree(h->stats);
    if (h->fin_sync) cudaFree(h->fin_sync);
    if (h->run_graph) cudaGraphExecDestroy(h->run_graph);
    if (h->run_graph1) cudaGraphExecDestroy(h->run_graph1);
    if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
    delete h;
    return 0;
}

int hk_chunk(int64_t n_global, int nranks, int rank, int64_t* offset, int64_t* rows) {
    HK_ARG(nranks >= 1 && rank >= 0 && rank < nranks && n_global >= 0, "hk_chunk: bad arguments");
    int64_t c = n_global / nranks, rem = n_global % nranks, start;
    if (rem > rank) {
        c += 1;
        start = rank * c;
    } else {
        start = rank * c + rem;
    }
    if (offset) *offset = start;
    if (rows) *rows = c;
    return 0;
}

int hk_lloyd_accumulate(hk_handle_t hh, const void* X, int64_t n_local, int d, int64_t ldx, int dtype,
                        const void* C, int k, void* labels, int label_kind, double* partials, void* row_ws,
                        int64_t row_ws_bytes, int path, void* stream) {
    int rc = check_common("hk_lloyd_accumulate", hh, X, n_local, d, ldx, dtype, C, k);
    if (rc) return rc;
    HK_ARG(partials != nullptr, "hk_lloyd_accumulate: partials is null");
    Handle* h = reinterpret_cast<Handle*>(hh);
    HK_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (n_local == 0) {  // empty shard (dndarray.py:301-305): contributes nothing
        HK_CUDA(cudaMemsetAsync(partials, 0, (size_t)k * (d + 1) * sizeof(double), st));
        h->variant = "empty";
        return 0;
    }
    LloydArgs a{X,       n_local, d,      ldx,          dtype,   C,    k, labels, label_kind, partials, nullptr,
                nullptr, row_ws,  row_ws_bytes, nullptr, path, st};
    return run_pass(h, a);
}

int hk_lloyd_finalize(hk_handle_t hh, const double* partials, const void* C_in, void* C_out, int k, int d,
                      int dtype, int use_tol, double tol_cmp, void* shift2_out, int32_t* state,
                      void* stream) {
    HK_ARG(hh != nullptr && partials && C_in && C_out, "hk_lloyd_finalize: null argument");
    HK_ARG(dtype == HK_F32 || dtype == HK_F64, "hk_lloyd_finalize: bad dtype");
    HK_ARG(k >= 1 && d >= 1, "hk_lloyd_finalize: bad shape");
    Handle* h = reinterpret_cast<Handle*>(hh);
    HK_CUDA(cudaSetDevice(h->device));
    return launch_finalize(h, partials, C_in, C_out, nullptr, k, d, dtype, use_tol, tol_cmp, shift2_out,
                           state, reinterpret_cast<cudaStream_t>(stream));
}

// one Lloyd step on stream st: pass over X, then ONE finish launch (slot reduce -> peer exchange -> finalize); the
// reduce / ncclAllReduce / finalize sequence remains as the fallback when the peers are not mapped
static int step_impl(Handle* h, const void* X, int64_t n_local, int d, int64_t ldx, int dtype, void* C, void* C_prev,
                     int k, void* labels, int label_kind, int use_tol, double tol_cmp, void* shift2_out,
                     int32_t* state, int allreduce, void* row_ws, int64_t row_ws_bytes, int path, cudaStream_t st) {
    const size_t len = (size_t)k * (d + 1);
    int rc = ensure_red(h, len * sizeof(double));
    if (rc) return rc;
    rc = ensure_fin_sync(h);
    if (rc) return rc;
    const bool exchange = allreduce && h->nranks > 1;
    const bool fused = !exchange || (h->peer_ready && len <= h->peer_cap);
    SlotInfo slots;
    if (n_local == 0) {
        HK_CUDA(cudaMemsetAsync(h->red, 0, len * sizeof(double), st));
        h->variant = "empty";
    } else {
        LloydArgs a{X,     n_local, d,            ldx,                       dtype, C, k, labels, label_kind, h->red, nullptr,
                    state, row_ws,  row_ws_bytes, fused ? &slots : nullptr, path,  st};
        rc = run_pass(h, a);
        if (rc) return rc;
    }
    if (!fused) {
        rc = comm_allreduce_f64(h, h->red, (int64_t)len, st);
        if (rc) return rc;
        return launch_finalize(h, h->red, C, C, C_prev, k, d, dtype, use_tol, tol_cmp, shift2_out, state, st);
    }
    FinishParams fp{};
    fp.fsum = slots.fsum;
    fp.fcnt = slots.fcnt;
    fp.nslots = slots.nslots;
    fp.slot_stride = slots.slot_stride;
    fp.nblocks = slots.nblocks;
    fp.partials_in = h->red;
    fp.k = k;
    fp.d = d;
    fp.red = h->red;
    fp.partials_out = nullptr;
    fp.nranks = exchange ? h->nranks : 1;
    fp.rank = exchange ? h->rank : 0;
    for (int r = 0; r < fp.nranks && exchange; ++r) {
        fp.mbox[r] = h->peer_mbox[r];
        fp.flags[r] = h->peer_flags[r];
    }
    fp.cap = h->peer_cap;
    fp.ticket = h->fin_sync;
    fp.epoch = h->fin_sync + 1;
    fp.C_in = C;
    fp.C_out = C;
    fp.C_prev = C_prev;
    fp.use_tol = use_tol;
    fp.tol_cmp = tol_cmp;
    fp.shift2_out = shift2_out;
    fp.state = state;
    return launch_finish(h, fp, dtype, st);
}

int hk_lloyd_step(hk_handle_t hh, const void* X, int64_t n_local, int d, int64_t ldx, int dtype, void* C,
                  void* C_prev, int k, void* labels, int label_kind, int use_tol, double tol_cmp,
                  void* shift2_out, int32_t* state, int allreduce, void* row_ws, int64_t row_ws_bytes, int path,
                  void* stream) {
    int rc = check_common("hk_lloyd_step", hh, X, n_local, d, ldx, dtype, C, k);
    if (rc) return rc;
    Handle* h = reinterpret_cast<Handle*>(hh);
    HK_CUDA(cudaSetDevice(h->device));
    return step_impl(h, X, n_local, d, ldx, dtype, C, C_prev, k, labels, label_kind, use_tol, tol_cmp, shift2_out, state,
                     allreduce, row_ws, row_ws_bytes, path, reinterpret_cast<cudaStream_t>(stream));
}

constexpr int RUN_CHUNK = 8;  // steps per replayed graph (the default sync_every of KMeans.fit)

// `iters` Lloyd steps enqueued by one call.  The first step runs eagerly (it may grow scratch buffers and fills the
// row workspace); the remaining ones are replayed from two CUDA graphs (RUN_CHUNK steps / one step; 2 kernel nodes per
// step) that are captured once and cached in the handle for as long as the call shape stays the same, so a fit costs
// one graph launch per `sync_every` iterations instead of 2 x sync_every kernel launches from Python.
int hk_lloyd_run(hk_handle_t hh, const void* X, int64_t n_local, int d, int64_t ldx, int dtype, void* C, void* C_prev,
                 int k, int use_tol, double tol_cmp, void* shift2_out, int32_t* state, int allreduce, void* row_ws,
                 int64_t row_ws_bytes, int path, int iters, void* stream) {
    int rc = check_common("hk_lloyd_run", hh, X, n_local, d, ldx, dtype, C, k);
    if (rc) return rc;
    HK_ARG(iters >= 0, "hk_lloyd_run: iters < 0");
    Handle* h = reinterpret_cast<Handle*>(hh);
    HK_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    auto one = [&]() {
        return step_impl(h, X, n_local, d, ldx, dtype, C, C_prev, k, nullptr, HK_LABEL_NONE, use_tol, tol_cmp, shift2_out,
                         state, allreduce, row_ws, row_ws_bytes, path, st);
    };
    const size_t len = (size_t)k * (d + 1);
    const bool exchange = allreduce && h->nranks > 1;
    const bool graphable = !h->profile && !h->no_graph && iters >= 3 && (!exchange || (h->peer_ready && len <= h->peer_cap));
    if (!graphable) {
        for (int i = 0; i < iters; ++i) {
            rc = one();
            if (rc) return rc;
        }
        return 0;
    }
    uint64_t tol_bits;
    memcpy(&tol_bits, &tol_cmp, sizeof(tol_bits));
    std::vector<uint64_t> key = {(uint64_t)(uintptr_t)X, (uint64_t)n_local, (uint64_t)d, (uint64_t)ldx, (uint64_t)dtype,
                                 (uint64_t)(uintptr_t)C, (uint64_t)(uintptr_t)C_prev, (uint64_t)k, (uint64_t)use_tol,
                                 tol_bits, (uint64_t)(uintptr_t)shift2_out, (uint64_t)(uintptr_t)state,
                                 (uint64_t)allreduce, (uint64_t)(uintptr_t)row_ws, (uint64_t)row_ws_bytes, (uint64_t)path,
                                 (uint64_t)(uintptr_t)st, (uint64_t)h->nranks, (uint64_t)(uintptr_t)h->part,
                                 (uint64_t)(uintptr_t)h->red};
    int done = 0;
    if (!(h->run_graph && h->run_graph_key == key)) {
        rc = one();  // eager: allocations, function attributes, row workspace
        if (rc) return rc;
        done = 1;
        key[18] = (uint64_t)(uintptr_t)h->part;
        key[19] = (uint64_t)(uintptr_t)h->red;
    }
    if (!(h->run_graph && h->run_graph_key == key)) {
        // (re)capture: one graph of RUN_CHUNK steps and one of a single step, replayed as often as `iters` needs, so a
        // different iteration count never recaptures.  Capture happens on a stream of the library's own (the caller's may
        // be the legacy default stream, which cannot be captured); the graphs are launched into the caller's stream.
        if (h->run_graph) cudaGraphExecDestroy(h->run_graph);
        if (h->run_graph1) cudaGraphExecDestroy(h->run_graph1);
        h->run_graph = h->run_graph1 = nullptr;
        if (!h->cap_stream) HK_CUDA(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
        cudaStream_t cs = h->cap_stream;
        for (int which = 0; which < 2; ++which) {
            const int nsteps = which == 0 ? RUN_CHUNK : 1;
            cudaGraph_t g = nullptr;
            HK_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
            for (int i = 0; i < nsteps && rc == 0; ++i)
                rc = step_impl(h, X, n_local, d, ldx, dtype, C, C_prev, k, nullptr, HK_LABEL_NONE, use_tol, tol_cmp,
                               shift2_out, state, allreduce, row_ws, row_ws_bytes, path, cs);
            const cudaError_t ce = cudaStreamEndCapture(cs, &g);
            if (rc || ce != cudaSuccess) {
                if (g) cudaGraphDestroy(g);
                cudaGetLastError();
                if (rc) return rc;
                set_error("hk_lloyd_run: stream capture failed: %s", cudaGetErrorString(ce));
                return 1000 + (int)ce;
            }
            cudaGraphExec_t* dst = which == 0 ? &h->run_graph : &h->run_graph1;
            const cudaError_t ie = cudaGraphInstantiate(dst, g, 0);
            cudaGraphDestroy(g);
            if (ie != cudaSuccess) {
                *dst = nullptr;
                set_error("hk_lloyd_run: cudaGraphInstantiate failed: %s", cudaGetErrorString(ie));
                return 1000 + (int)ie;
            }
            HK_CUDA(cudaGraphUpload(*dst, st));  // pay the upload now, not at the first replay
        }
        h->run_graph_key = key;
        h->graph_builds++;
    }
    int left = iters - done;
    while (left >= RUN_CHUNK) {
        HK_CUDA(cudaGraphLaunch(h->run_graph, st));
        h->launches += 2 * (int64_t)RUN_CHUNK;
        h->graph_launches++;
        left -= RUN_CHUNK;
    }
    while (left > 0) {
        HK_CUDA(cudaGraphLaunch(h->run_graph1, st));
        h->launches += 2;
        h->graph_launches++;
        --left;
    }
    return 0;
}

int hk_assign(hk_handle_t hh, const void* X, int64_t n_local, int d, int64_t ldx, int dtype, const void* C,
              int k, void* labels, int label_kind, double* min_d2_sum, void* row_ws, int64_t row_ws_bytes, int path,
              void* stream) {
    int rc = check_common("hk_assign", hh, X, n_local, d, ldx, dtype, C, k);
    if (rc) return rc;
    Handle* h = reinterpret_cast<Handle*>(hh);
    HK_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (n_local == 0) {
        if (min_d2_sum) HK_CUDA(cudaMemsetAsync(min_d2_sum, 0, sizeof(double), st));
        h->variant = "empty";
        return 0;
    }
    LloydArgs a{X,       n_local, d,            ldx,     dtype, C,  k, labels, label_kind, nullptr, min_d2_sum,
                nullptr, row_ws,  row_ws_bytes, nullptr, path,  st};
    return run_pass(h, a);
}

int hk_cdist(hk_handle_t hh, const void* X, int64_t m, int f, int64_t ldx, const void* Y, int64_t n,
             int64_t ldy, void* out, int64_t ldo, int dtype, int quadratic_expansion, int sqrt_flag,
             void* stream) {
    HK_ARG(hh != nullptr, "hk_cdist: null handle");
    HK_ARG(dtype == HK_F32 || dtype == HK_F64, "hk_cdist: bad dtype");
    HK_ARG(m >= 0 && n >= 0 && f >= 1, "hk_cdist: bad shape");
    HK_ARG(ldx >= f && ldy >= f && ldo >= n, "hk_cdist: bad leading dimension");
    if (m == 0 || n == 0) return 0;
    HK_ARG(X && Y && out, "hk_cdist: null pointer");
    Handle* h = reinterpret_cast<Handle*>(hh);
    HK_CUDA(cudaSetDevice(h->device));
    return launch_cdist(h, X, m, f, ldx, Y, n, ldy, out, ldo, dtype, quadratic_expansion, sqrt_flag,
                        reinterpret_cast<cudaStream_t>(stream));
}

int hk_pairwise(hk_handle_t hh, const void* X, int64_t m, int f, int64_t ldx, const void* Y, int64_t n,
                int64_t ldy, void* out, int64_t ldo, int dtype, int metric, int expand, double sigma,
                void* stream) {
    HK_ARG(hh != nullptr, "hk_pairwise: null handle");
    HK_ARG(dtype == HK_F32 || dtype == HK_F64, "hk_pairwise: bad dtype");
    HK_ARG(metric == HK_METRIC_EUCLIDEAN || metric == HK_METRIC_GAUSSIAN || metric == HK_METRIC_MANHATTAN,
           "hk_pairwise: bad metric");
    HK_ARG(m >= 0 && n >= 0 && f >= 1, "hk_pairwise: bad shape");
    HK_ARG(ldx >= f && ldy >= f && ldo >= n, "hk_pairwise: bad leading dimension");
    HK_ARG(metric != HK_METRIC_GAUSSIAN || (sigma == sigma && sigma != 0.0), "hk_pairwise: bad sigma");
    if (m == 0 || n == 0) return 0;
    HK_ARG(X && Y && out, "hk_pairwise: null pointer");
    Handle* h = reinterpret_cast<Handle*>(hh);
    HK_CUDA(cudaSetDevice(h->device));
    const int body = metric == HK_METRIC_MANHATTAN ? 2 : (expand ? 1 : 0);
    const int post = metric == HK_METRIC_EUCLIDEAN ? 1 : (metric == HK_METRIC_GAUSSIAN ? 2 : 0);
    return launch_cdist(h, X, m, f, ldx, Y, n, ldy, out, ldo, dtype, body, post,
                        reinterpret_cast<cudaStream_t>(stream), 2.0 * sigma * sigma);
}

// ---- other consumers of the assignment pattern --------------------------------------------------
#define HK_HANDLE(name)                                   \
    HK_ARG(hh != nullptr, name ": null handle");          \
    Handle* h = reinterpret_cast<Handle*>(hh);            \
    HK_CUDA(cudaSetDevice(h->device));                    \
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream)

int hk_assign_l1(hk_handle_t hh, const void* X, int64_t n_local, int d, int64_t ldx, int dtype, const void* C, int k,
                 void* labels, int label_kind, double* min_d_sum, void* stream) {
    HK_ARG(dtype == HK_F32 || dtype == HK_F64, "hk_assign_l1: bad dtype");
    HK_ARG(n_local >= 0 && d >= 1 && k >= 1 && ldx >= d, "hk_assign_l1: bad shape");
    HK_ARG(label_kind >= HK_LABEL_NONE && label_kind <= HK_LABEL_I64, "hk_assign_l1: bad label kind");
    HK_ARG(label_kind != HK_LABEL_U8 || k <= 256, "hk_assign_l1: uint8 labels need k <= 256");
    HK_HANDLE("hk_assign_l1");
    if (n_local == 0) {
        if (min_d_sum != nullptr) HK_CUDA(cudaMemsetAsync(min_d_sum, 0, sizeof(double), st));
        return 0;
    }
    HK_ARG(X && C, "hk_assign_l1: null pointer");
    HK_ARG(label_kind == HK_LABEL_NONE || labels != nullptr, "hk_assign_l1: null labels");
    return launch_assign_l1(h, X, n_local, d, ldx, dtype, C, k, labels, label_kind, min_d_sum, st);
}

int hk_row_keep(hk_handle_t hh, const void* X, int64_t n_local, int d, int64_t ldx, int dtype, void* keep_u8,
                void* stream) {
    HK_ARG(dtype == HK_F32 || dtype == HK_F64, "hk_row_keep: bad dtype");
    HK_ARG(n_local >= 0 && d >= 1 && ldx >= d, "hk_row_keep: bad shape");
    HK_HANDLE("hk_row_keep");
    if (n_local == 0) return 0;
    HK_ARG(X && keep_u8, "hk_row_keep: null pointer");
    return launch_row_keep(h, X, n_local, d, ldx, dtype, reinterpret_cast<uint8_t*>(keep_u8), st);
}

int hk_select_passes(int dtype) { return dtype == HK_F64 ? 8 : 4; }

int hk_select_hist(hk_handle_t hh, const void* X, int64_t n_local, int d, int64_t ldx, int dtype, const void* labels_i64,
                   const void* keep_u8, int k, const void* prefix_u64, int pass, void* hist_i64, void* stream) {
    HK_ARG(dtype == HK_F32 || dtype == HK_F64, "hk_select_hist: bad dtype");
    HK_ARG(n_local >= 0 && d >= 1 && k >= 1 && ldx >= d, "hk_select_hist: bad shape");
    HK_ARG(pass >= 0 && pass < hk_select_passes(dtype), "hk_select_hist: bad pass");
    HK_HANDLE("hk_select_hist");
    if (n_local == 0) return 0;
    HK_ARG(X && labels_i64 && keep_u8 && prefix_u64 && hist_i64, "hk_select_hist: null pointer");
    return launch_select_hist(h, X, n_local, d, ldx, dtype, reinterpret_cast<const int64_t*>(labels_i64),
                              reinterpret_cast<const uint8_t*>(keep_u8), k, reinterpret_cast<const uint64_t*>(prefix_u64),
                              pass, reinterpret_cast<unsigned long long*>(hist_i64), st);
}

int hk_select_step(hk_handle_t hh, const void* hist_i64, void* remaining_i64, void* prefix_u64, int k, int d,
                   void* stream) {
    HK_ARG(k >= 1 && d >= 1, "hk_select_step: bad shape");
    HK_HANDLE("hk_select_step");
    HK_ARG(hist_i64 && remaining_i64 && prefix_u64, "hk_select_step: null pointer");
    return launch_select_step(h, reinterpret_cast<const unsigned long long*>(hist_i64),
                              reinterpret_cast<int64_t*>(remaining_i64), reinterpret_cast<uint64_t*>(prefix_u64), 2 * k * d,
                              st);
}

int hk_select_value(hk_handle_t hh, const void* prefix_u64, const double* frac, int k, int d, int dtype, void* medians,
                    void* stream) {
    HK_ARG(dtype == HK_F32 || dtype == HK_F64, "hk_select_value: bad dtype");
    HK_ARG(k >= 1 && d >= 1, "hk_select_value: bad shape");
    HK_HANDLE("hk_select_value");
    HK_ARG(prefix_u64 && frac && medians, "hk_select_value: null pointer");
    return launch_select_value(h, reinterpret_cast<const uint64_t*>(prefix_u64), frac, k, d, dtype, medians, st);
}

int hk_kmex_update(hk_handle_t hh, const double* partials, const void* medians, const void* counts_i64, void* C, int k,
                   int d, int dtype, double atol, double rtol, void* flag_i32, void* stream) {
    HK_ARG(dtype == HK_F32 || dtype == HK_F64, "hk_kmex_update: bad dtype");
    HK_ARG(k >= 1 && d >= 1, "hk_kmex_update: bad shape");
    HK_HANDLE("hk_kmex_update");
    HK_ARG(C && flag_i32, "hk_kmex_update: null pointer");
    HK_ARG((partials != nullptr) != (medians != nullptr && counts_i64 != nullptr),
           "hk_kmex_update: pass either the partial sums or medians + counts");
    return launch_kmex_update(h, partials, medians, reinterpret_cast<const int64_t*>(counts_i64), C, k, d, dtype, atol, rtol,
                              reinterpret_cast<int*>(flag_i32), st);
}

int hk_nearest_rows_l1(hk_handle_t hh, const void* X, int64_t n_local, int d, int64_t ldx, int dtype, const void* P,
                       int k, int64_t row_base, double* best_dist, void* best_index_i64, void* stream) {
    HK_ARG(dtype == HK_F32 || dtype == HK_F64, "hk_nearest_rows_l1: bad dtype");
    HK_ARG(n_local >= 1 && d >= 1 && k >= 1 && ldx >= d, "hk_nearest_rows_l1: bad shape");
    HK_HANDLE("hk_nearest_rows_l1");
    HK_ARG(X && P && best_dist && best_index_i64, "hk_nearest_rows_l1: null pointer");
    return launch_nearest_rows_l1(h, X, n_local, d, ldx, dtype, P, k, row_base, best_dist,
                                  reinterpret_cast<int64_t*>(best_index_i64), st);
}

int hk_topk_rows(hk_handle_t hh, const void* D, int64_t m, int64_t n, int64_t ldd, int dtype, int kk, void* values,
                 void* indices_i64, void* stream) {
    HK_ARG(dtype == HK_F32 || dtype == HK_F64, "hk_topk_rows: bad dtype");
    HK_ARG(m >= 0 && n >= 1 && ldd >= n && kk >= 1 && kk <= n, "hk_topk_rows: bad shape");
    HK_HANDLE("hk_topk_rows");
    if (m == 0) return 0;
    HK_ARG(D && values && indices_i64, "hk_topk_rows: null pointer");
    return launch_topk_rows(h, D, m, n, ldd, dtype, kk, values, reinterpret_cast<int64_t*>(indices_i64), st);
}

int hk_knn_vote(hk_handle_t hh, const void* indices_i64, int64_t m, int kk, const void* Y, int64_t n, int n_classes,
                int64_t ldy, int dtype, void* classes_i64, void* stream) {
    HK_ARG(dtype == HK_F32 || dtype == HK_F64, "hk_knn_vote: bad dtype");
    HK_ARG(m >= 0 && n >= 1 && kk >= 1 && n_classes >= 1 && ldy >= n_classes, "hk_knn_vote: bad shape");
    HK_HANDLE("hk_knn_vote");
    if (m == 0) return 0;
    HK_ARG(indices_i64 && Y && classes_i64, "hk_knn_vote: null pointer");
    return launch_knn_vote(h, reinterpret_cast<const int64_t*>(indices_i64), m, kk, Y, n, n_classes, ldy, dtype,
                           reinterpret_cast<int64_t*>(classes_i64), st);
}

int hk_comm_unique_id(void* id128) {
    HK_ARG(id128 != nullptr, "hk_comm_unique_id: null");
    return comm_unique_id(id128);
}
int hk_comm_init(hk_handle_t hh, int nranks, int rank, const void* id128) {
    HK_ARG(hh != nullptr, "hk_comm_init: null handle");
    HK_ARG(nranks == 1 || id128 != nullptr, "hk_comm_init: null id");
    return comm_init(reinterpret_cast<Handle*>(hh), nranks, rank, id128);
}
int hk_comm_peer_export(hk_handle_t hh, int64_t cap_doubles, void* handle64) {
    HK_ARG(hh != nullptr, "hk_comm_peer_export: null handle");
    return comm_peer_export(reinterpret_cast<Handle*>(hh), cap_doubles, handle64);
}
int hk_comm_peer_import(hk_handle_t hh, const void* handles) {
    HK_ARG(hh != nullptr, "hk_comm_peer_import: null handle");
    return comm_peer_import(reinterpret_cast<Handle*>(hh), handles);
}
int hk_comm_mode(hk_handle_t hh) {
    if (!hh) return 0;
    Handle* h = reinterpret_cast<Handle*>(hh);
    if (h->nranks <= 1) return 0;
    return h->peer_ready ? 2 : 1;
}
int hk_comm_destroy(hk_handle_t hh) {
    if (!hh) return 0;
    return comm_destroy(reinterpret_cast<Handle*>(hh));
}
int hk_allreduce_f64(hk_handle_t hh, double* buf, int64_t count, void* stream) {
    HK_ARG(hh != nullptr && buf != nullptr && count >= 0, "hk_allreduce_f64: bad argument");
    return comm_allreduce_f64(reinterpret_cast<Handle*>(hh), buf, count,
                              reinterpret_cast<cudaStream_t>(stream));
}

int hk_graph_enable(hk_handle_t hh, int enable) {
    HK_ARG(hh != nullptr, "hk_graph_enable: null handle");
    reinterpret_cast<Handle*>(hh)->no_graph = enable == 0;
    return 0;
}
int64_t hk_graph_launch_count(hk_handle_t hh) { return hh ? reinterpret_cast<Handle*>(hh)->graph_launches : 0; }

int64_t hk_launch_count(hk_handle_t hh) { return hh ? reinterpret_cast<Handle*>(hh)->launches : 0; }
const char* hk_last_variant(hk_handle_t hh) {
    return hh ? reinterpret_cast<Handle*>(hh)->variant.c_str() : "";
}

int64_t hk_row_ws_bytes(int64_t n_local) {
    if (n_local < 0) return 0;
    return (int64_t)(((n_local + 127) / 128 + 4) * sizeof(float));
}

int hk_stats_read(hk_handle_t hh, int64_t* out6) {
    HK_ARG(hh != nullptr && out6 != nullptr, "hk_stats_read: null argument");
    Handle* h = reinterpret_cast<Handle*>(hh);
    for (int i = 0; i < 6; ++i) out6[i] = 0;
    if (!h->stats) return 0;
    HK_CUDA(cudaSetDevice(h->device));
    HK_CUDA(cudaDeviceSynchronize());
    unsigned long long v[6];
    HK_CUDA(cudaMemcpy(v, h->stats, sizeof(v), cudaMemcpyDeviceToHost));
    HK_CUDA(cudaMemset(h->stats, 0, sizeof(v)));
    for (int i = 0; i < 6; ++i) out6[i] = (int64_t)v[i];
    return 0;
}

int hk_profile_enable(hk_handle_t hh, int enable) {
    HK_ARG(hh != nullptr, "hk_profile_enable: null handle");
    reinterpret_cast<Handle*>(hh)->profile = enable != 0;
    return 0;
}
int hk_profile_read(hk_handle_t hh, double* total_ms, int64_t* launches) {
    HK_ARG(hh != nullptr, "hk_profile_read: null handle");
    Handle* h = reinterpret_cast<Handle*>(hh);
    HK_CUDA(cudaSetDevice(h->device));
    double tot = 0.0;
    int64_t n = 0;
    for (auto& ev : h->prof_events) {
        HK_CUDA(cudaEventSynchronize(ev.second));
        float ms = 0.f;
        HK_CUDA(cudaEventElapsedTime(&ms, ev.first, ev.second));
        tot += ms;
        ++n;
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
    }
    h->prof_events.clear();
    if (total_ms) *total_ms = tot;
    if (launches) *launches = n;
    return 0;
}

}  // extern "C"
