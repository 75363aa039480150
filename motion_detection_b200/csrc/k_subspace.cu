// k_subspace.cu -- placeholder until the fitSubspace scorer lands.
#include "md_internal.h"
extern "C" int md_fit_subspace(md_ctx *ctx, const float *, int32_t, int32_t, int32_t, double, uint32_t, const int32_t *, int32_t,
                               float *, int32_t *, uint8_t *, int32_t *, int)
{
    if (!ctx) return MD_ERR_INVALID;
    ctx->err = "md_fit_subspace: not implemented yet";
    return MD_ERR_UNSUPPORTED;
}
