// tcgen05 / TMEM graph-linear path (bf16 and 3-plane split-bf16).  Placeholder until the
// tensor-core kernel lands: requesting it fails loudly instead of silently using the FFMA path.
#include "sd_internal.h"

namespace sd {
int glin_forward_tc(const sd_glin* L, const GlinCall& c, int precision, cudaStream_t st) {
    (void)L; (void)c; (void)st;
    set_error("precision %d (tcgen05 path) is not built in this library", precision);
    return SD_ERR_UNSUPPORTED;
}
}  // namespace sd
