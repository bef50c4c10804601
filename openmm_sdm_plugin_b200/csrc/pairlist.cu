// pairlist.cu -- cluster-pair list path.  (placeholder until the cluster kernel lands)
#include "sdm_ctx.h"

int sdm_ctx_init_pairlist(sdm_ctx* c) {
    if (c->pair_mode == SDM_PAIR_CLUSTER) return sdm_fail(SDM_ERR_INVALID, "cluster pair mode not built");
    return SDM_OK;
}
void sdm_ctx_free_pairlist(sdm_ctx*) {}
int sdm_ctx_pairlist_eval(sdm_ctx*) { return sdm_fail(SDM_ERR_INVALID, "cluster pair mode not built"); }
int sdm_ctx_pairlist_emit(sdm_ctx*, int, int*, int*, int) { return sdm_fail(SDM_ERR_INVALID, "cluster pair mode not built"); }
int sdm_ctx_pairlist_info(sdm_ctx*, const char*, double*) { return SDM_ERR_INVALID; }
