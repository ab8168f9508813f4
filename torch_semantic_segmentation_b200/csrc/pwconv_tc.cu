// Pointwise convolution on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
// Placeholder until the kernel lands: the entry point refuses loudly (no fallback).
#include "common.cuh"

int tss_pwconv_fwd_tc(const void* x, const void* wp, void* y, int64_t M, int K, int Nc, int64_t ldx,
                      int64_t ldy, const float* scale, const float* shift, const void* res, int64_t ldr,
                      int flags, float* stats, cudaStream_t st) {
    tss_set_error("pwconv impl 1 (tcgen05) is not built yet");
    return TSS_ERR_ARG;
}
