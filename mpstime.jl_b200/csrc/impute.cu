// K8 placeholder: replaced below in this round by the batched imputation kernels.
#include "mpst_common.cuh"
int impute_batch(mpst_ctx* c, int, const double*, const uint8_t*, int64_t, int, const double*, int, const double*, int,
                 double, double*) {
    c->err = "impute_batch: not built yet";
    return MPST_E_UNSUPPORTED;
}
