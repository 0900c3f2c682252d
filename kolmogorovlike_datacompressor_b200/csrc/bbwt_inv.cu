// bbwt_inv.cu — placeholder until the inverse transform lands
#include "common.cuh"
int kolm_bbwt_inv_impl(kolm_ctx* c, const u8* in, u8* out, cudaStream_t s) { (void)c; (void)in; (void)out; (void)s; return KOLM_E_UNSUPPORTED; }
