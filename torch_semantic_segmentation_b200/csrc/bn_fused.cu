// BatchNorm (training) forward with the finalize step folded into the apply kernel: 44 launches of a C-thread
// kernel per step sit on the critical chain conv -> finalize -> apply of every layer (bn.cu); here the apply pass
// derives scale / shift from the fp64 statistics itself, CTA 0 also writes mean / rstd for the backward pass and
// updates the running statistics, and the per-layer scratch is cleared ("consume and clear", bn.cu:
// bn_finalize_kernel) by whichever CTA is the LAST to finish -- a ticket counter that the same CTA resets.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

// Thread = (8-channel group, row lane) like bn.cu's apply kernel.  Prologue, ONCE PER CTA: thread c derives scale / shift of
// channel c from the fp64 statistics with bn_finalize_kernel's own arithmetic (bit-identical constants) into shared memory --
// ~100 instructions per channel including the fp64 sqrt and division.  (The first version let every thread derive the
// constants of its eight channels: 8 x that chain in each of up to 300 k threads, 9 us per launch slower than the separate
// finalize kernel.)  CTA 0 also publishes mean / rstd / running statistics.  The ticket is taken at the END of the CTA:
// by then every CTA has long read the statistics, and the last one clears the scratch.
template <typename T, int U, bool kRes>
__global__ void __launch_bounds__(kThreads)
bn_finalize_apply_kernel(double* stats, double inv_count, double unbias, const float* __restrict__ gamma,
                         const float* __restrict__ beta, float* __restrict__ running_mean,
                         float* __restrict__ running_var, int64_t* __restrict__ nbt, float momentum, float eps,
                         float* __restrict__ mean_out, float* __restrict__ rstd_out, int* ticket, int clear_n,
                         const T* __restrict__ y, const T* __restrict__ res, T* __restrict__ z, int64_t M, int C,
                         int64_t ldy, int64_t ldr, int64_t ldz, int relu, int PL) {
    __shared__ int s_last;
    TSS_DYN_SMEM(float, s_const);                      // [2][C]: scale | shift
    const int CG = C >> 3;
    const int cg = threadIdx.x % CG;
    const int pl = threadIdx.x / CG;
    const int c0 = cg * 8;
    pdl_wait();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const double st1 = stats[c], st2 = stats[C + c];
        const double mean = st1 * inv_count;
        double var = st2 * inv_count - mean * mean;
        if (var < 0.0) var = 0.0;
        const float rstd = (float)(1.0 / sqrt(var + (double)eps));
        const float sc = (gamma != nullptr ? __ldg(gamma + c) : 1.f) * rstd;
        s_const[c] = sc;
        s_const[C + c] = (beta != nullptr ? __ldg(beta + c) : 0.f) - (float)mean * sc;
        if (blockIdx.x == 0) {
            mean_out[c] = (float)mean;
            rstd_out[c] = rstd;
            if (running_mean != nullptr) {
                running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
                running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(var * unbias);
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && nbt != nullptr) *nbt += 1;
    __syncthreads();
    if (pl < PL) {
        float sc[8], sh[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { sc[e] = s_const[c0 + e]; sh[e] = s_const[C + c0 + e]; }
        const int64_t step = (int64_t)gridDim.x * PL;
        for (int64_t m0 = (int64_t)blockIdx.x * PL + pl; m0 < M; m0 += U * step) {
            Raw8<T> ry[U], rr[kRes ? U : 1];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t m = m0 + u * step;
                if (m < M) {
                    ry[u].ld(y + m * ldy + c0);
                    if (kRes) rr[u].ld(res + m * ldr + c0);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t m = m0 + u * step;
                if (m >= M) break;
                float v[8];
                ry[u].get(v);
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = fmaf(v[e], sc[e], sh[e]);
                if (kRes) {
                    float w[8];
                    rr[u].get(w);
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] += w[e];
                }
                if (relu) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
                }
                store8(z + m * ldz + c0, v);
            }
        }
    }
    __syncthreads();                                   // every thread of this CTA has read the statistics long ago
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(ticket, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (s_last) {                                      // every CTA has: the scratch is zero again for the next step
        __threadfence();
        for (int i = threadIdx.x; i < clear_n; i += blockDim.x) stats[i] = 0.0;
        if (threadIdx.x == 0) *ticket = 0;
    }
}

}  // namespace

extern "C" int tss_bn_finalize_apply(double* stats, int64_t count, const float* gamma, const float* beta,
                                     float* running_mean, float* running_var, int64_t* num_batches_tracked,
                                     float momentum, float eps, float* mean, float* rstd, int* ticket,
                                     int64_t clear_n, const void* y, const void* res, void* z, int64_t M, int C,
                                     int64_t ldy, int64_t ldr, int64_t ldz, int flags, int dtype, void* stream) {
    TSS_REQUIRE(M > 0 && C > 0 && C % 8 == 0 && C <= 2048 && count > 0, "bn_finalize_apply: M=%lld C=%d count=%lld", (long long)M, C, (long long)count);
    TSS_REQUIRE(count > 1, "bn_finalize_apply: Expected more than 1 value per channel when training");
    TSS_REQUIRE(stats != nullptr && mean != nullptr && rstd != nullptr && ticket != nullptr && y != nullptr && z != nullptr,
                "bn_finalize_apply: missing buffer");
    TSS_REQUIRE(ldy % 8 == 0 && ldz % 8 == 0 && (res == nullptr || ldr % 8 == 0) && clear_n >= 0, "bn_finalize_apply: bad pitch");
    TSS_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "bn_finalize_apply: running_mean and running_var go together");
    const double unbias = (double)count / (double)(count - 1);
    const int CG = C / 8;
    TSS_REQUIRE(CG <= kThreads, "bn_finalize_apply: C=%d too large", C);
    const int PL = kThreads / CG, threads = PL * CG;
    TSS_DISPATCH_DTYPE(dtype, "bn_finalize_apply", {
        constexpr int U = sizeof(T) == 2 ? 4 : 2;
        int64_t grid = ceil_div64(M, (int64_t)PL * U);
        const int64_t cap = (int64_t)tss_num_sms() * 4;          // the resident CTAs (64 registers): the prologue runs once per CTA
        if (grid > cap) grid = cap;
        if (grid < 1) grid = 1;
        auto kern = res != nullptr ? bn_finalize_apply_kernel<T, U, true> : bn_finalize_apply_kernel<T, U, false>;
        tss_launch(kern, (unsigned)grid, threads, (size_t)2 * C * sizeof(float), (cudaStream_t)stream, stats,
                   1.0 / (double)count, unbias, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps, mean,
                   rstd, ticket, (int)clear_n, (const T*)y, (const T*)res, (T*)z, M, C, ldy, ldr, ldz, flags & TSS_EPI_RELU, PL);
        TSS_LAUNCH_CHECK("bn_finalize_apply");
        return TSS_OK;
    });
}
