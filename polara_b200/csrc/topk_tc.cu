// placeholder until the tcgen05 kernel lands
#include "topk_common.cuh"
int pb_score_tc(pb200_ctx* ctx, const float*, int64_t, const float*, int64_t, int64_t, int64_t, int,
                const int64_t*, const int32_t*, int64_t, int, int*, pb200_cand**, Scratch&) {
    ctx->err = "tcgen05 scoring kernel not built";
    return PB200_ENOTIMPL;
}
