// rice.cu — placeholder
#include "common.cuh"
int kolm_rice_kf_enc_impl(kolm_ctx* c, const u8* mtf, u8* out, size_t out_cap, i64* out_off, int* params, cudaStream_t s) { return KOLM_E_UNSUPPORTED; }
int kolm_rice_kf_dec_impl(kolm_ctx* c, const u8* pay, const i64* pay_off, u8* mtf_out, cudaStream_t s) { return KOLM_E_UNSUPPORTED; }
int kolm_rice_k2_enc_impl(kolm_ctx* c, const u8* mtf, int flags, u8* out, size_t out_cap, i64* out_off, i64* sizes, cudaStream_t s) { return KOLM_E_UNSUPPORTED; }
int kolm_rice_k2_dec_impl(kolm_ctx* c, const u8* pay, const i64* pay_off, int flags, u8* mtf_out, cudaStream_t s) { return KOLM_E_UNSUPPORTED; }
