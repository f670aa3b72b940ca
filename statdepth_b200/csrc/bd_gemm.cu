// bd_gemm.cu -- strict band depth as an int8 violation Gram on tcgen05 / TMEM (placeholder until the
// tensor-core kernel lands; SD_BD_AUTO never routes here).
#include "common.cuh"

namespace sd {

int bd_strict_gemm_device(sd_ctx *ctx, const double *dX, i64 T, i64 n, i64 ld, const i64 *d_q, i64 nq,
                          i64 *d_out) {
    (void)ctx; (void)dX; (void)T; (void)n; (void)ld; (void)d_q; (void)nq; (void)d_out;
    set_error("strict band depth: the tcgen05 Gram kernel is not built in this revision");
    return SD_ERR_UNSUPPORTED;
}

}  // namespace sd
