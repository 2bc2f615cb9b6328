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

int comm_unique_id(void* id128);
int comm_init(Handle* h, int nranks, int rank, const void* id128);
int comm_destroy(Handle* h);

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
    if (path == HK_PATH_SIMT && row128_supported(h, a)) return launch_lloyd_row128(h, a);
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
    if (h->xb) cudaFree(h->xb);
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
                        const void* C, int k, void* labels, int label_kind, double* partials, int path,
                        void* stream) {
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
    LloydArgs a{X, n_local, d, ldx, dtype, C, k, labels, label_kind, partials, nullptr, nullptr, path, st};
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

int hk_lloyd_step(hk_handle_t hh, const void* X, int64_t n_local, int d, int64_t ldx, int dtype, void* C,
                  void* C_prev, int k, void* labels, int label_kind, int use_tol, double tol_cmp,
                  void* shift2_out, int32_t* state, int allreduce, int path, void* stream) {
    int rc = check_common("hk_lloyd_step", hh, X, n_local, d, ldx, dtype, C, k);
    if (rc) return rc;
    Handle* h = reinterpret_cast<Handle*>(hh);
    HK_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t len = (size_t)k * (d + 1);
    rc = ensure_red(h, len * sizeof(double));
    if (rc) return rc;
    if (n_local == 0) {
        HK_CUDA(cudaMemsetAsync(h->red, 0, len * sizeof(double), st));
        h->variant = "empty";
    } else {
        LloydArgs a{X, n_local, d, ldx, dtype, C, k, labels, label_kind, h->red, nullptr, state, path, st};
        rc = run_pass(h, a);
        if (rc) return rc;
    }
    if (allreduce) {
        rc = comm_allreduce_f64(h, h->red, (int64_t)len, st);
        if (rc) return rc;
    }
    return launch_finalize(h, h->red, C, C, C_prev, k, d, dtype, use_tol, tol_cmp, shift2_out, state, st);
}

int hk_assign(hk_handle_t hh, const void* X, int64_t n_local, int d, int64_t ldx, int dtype, const void* C,
              int k, void* labels, int label_kind, double* min_d2_sum, int path, void* stream) {
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
    LloydArgs a{X, n_local, d, ldx, dtype, C, k, labels, label_kind, nullptr, min_d2_sum, nullptr, path, st};
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

int hk_comm_unique_id(void* id128) {
    HK_ARG(id128 != nullptr, "hk_comm_unique_id: null");
    return comm_unique_id(id128);
}
int hk_comm_init(hk_handle_t hh, int nranks, int rank, const void* id128) {
    HK_ARG(hh != nullptr, "hk_comm_init: null handle");
    HK_ARG(nranks == 1 || id128 != nullptr, "hk_comm_init: null id");
    return comm_init(reinterpret_cast<Handle*>(hh), nranks, rank, id128);
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

int64_t hk_launch_count(hk_handle_t hh) { return hh ? reinterpret_cast<Handle*>(hh)->launches : 0; }
const char* hk_last_variant(hk_handle_t hh) {
    return hh ? reinterpret_cast<Handle*>(hh)->variant.c_str() : "";
}

int hk_cache_reset(hk_handle_t hh) {
    HK_ARG(hh != nullptr, "hk_cache_reset: null handle");
    reinterpret_cast<Handle*>(hh)->xb_X = nullptr;
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
