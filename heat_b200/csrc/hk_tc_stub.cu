// Placeholder until the tcgen05 cdist kernel lands: report "unsupported" so that hk_cdist uses the
// exact-FMA kernel.  (Replaced by hk_cdist_tc.cu.)
#include "hk_common.cuh"
namespace hk {
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
