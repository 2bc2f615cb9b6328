// Placeholder until the tcgen05 kernels land: report "unsupported" so that HK_PATH_AUTO uses the
// exact-FMA kernels.  (Replaced by hk_lloyd_tc.cu / hk_cdist_tc.cu.)
#include "hk_common.cuh"
namespace hk {
bool tc_supported(const Handle*, const LloydArgs&) { return false; }
int launch_lloyd_tc(Handle*, const LloydArgs&) {
    set_error("tensor-core path not built");
    return -2;
}
bool cdist_tc_supported(const Handle*, const void*, int64_t, int, int64_t, const void*, int64_t, int64_t,
                        const void*, int64_t) {
    return false;
}
int launch_cdist_tc(Handle*, const void*, int64_t, int, int64_t, const void*, int64_t, int64_t, void*,
                    int64_t, int, cudaStream_t) {
    set_error("tensor-core cdist not built");
    return -2;
}
}  // namespace hk
