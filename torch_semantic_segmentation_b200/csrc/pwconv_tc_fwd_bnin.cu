// Pointwise-conv FORWARD (training) on tcgen05 whose input is the RAW conv output of the producing layer: the
// producer's BatchNorm + ReLU is applied by the threads that build the GEMM's A operand, the depthwise->pointwise
// counterpart of dwconv_bnin.cu (conv2 -> conv3 of a bottleneck, fastscnn.py:153-157):
//
//   prologue  z = relu?(x * in_scale + in_shift)     128 rows x 64 channels per chunk, bf16, written K-major /
//             128-byte swizzled (pwconv_tc_bwd.cu); the column-tile-0 CTAs also store z: the weight gradient of this
//             layer needs it as a dense operand
//   GEMM      y[M][Nc] = z[M][K] . W[Nc][K]^T        weights by TMA, accumulator in TMEM
//   epilogue  BatchNorm statistics of y (pwconv_tc.cu) + bf16 store of the raw output
//
// Against bn_apply + pwconv_fwd: the activated tensor is written once and never read in the forward pass
// (4 instead of 6 bytes per element of the wide tensor) and one launch less.
#include <stdlib.h>

#include "tc_ptx.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kThreads = 192;
constexpr uint32_t kABytes = BM * BK * 2;
constexpr int kStagesB = 2;
constexpr int kRowsInFlight = 4;

// lanes 2j / 2j+1 end with the sum over the 32 lanes of v[j], j = lane >> 1
__device__ __forceinline__ float warp_transpose_sum16(float (&v)[16], int lane) {
#pragma unroll
    for (int step = 16, n = 16; step >= 2; step >>= 1, n >>= 1) {
        const bool upper = (lane & step) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float send = upper ? v[i] : v[i + n / 2];
            const float keep = upper ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}




__global__ void __launch_bounds__(kThreads)
pw_tc_fwd_bnin_kernel(const __grid_constant__ CUtensorMap tmB, const bf16* __restrict__ x, int64_t ldx,
                      const float* __restrict__ in_scale, const float* __restrict__ in_shift, int in_relu,
                      bf16* __restrict__ z_out, int64_t ldz, bf16* __restrict__ Y, int64_t ldy, int64_t M, int K,
                      int block_n, uint32_t tmem_cols, double* __restrict__ stats, int stats_stride) {
    TSS_DYN_SMEM(uint8_t, smem_raw);
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t b_bytes = (uint32_t)block_n * BK * 2;
    const uint32_t b_pad = (b_bytes + 1023) & ~1023u;
    uint8_t* sA = smem;
    uint8_t* sB = sA + (size_t)kStagesB * kABytes;
    uint64_t* bars = (uint64_t*)(sB + (size_t)kStagesB * b_pad);
    uint64_t* b_full = bars, *a_full = bars + kStagesB, *ab_empty = bars + 2 * kStagesB, *tmem_full = bars + 3 * kStagesB;
    uint32_t* tmem_slot = (uint32_t*)(bars + 3 * kStagesB + 1);
    float* s_stat = (float*)(tmem_slot + 2);                           // [2][block_n]
    const int KP = (K + 63) & ~63;
    float* s_c = s_stat + 2 * block_n;                                 // [2][KP]: scale, shift of the producer

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * block_n;
    const int num_kb = (K + BK - 1) / BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStagesB; ++s) {
            mbar_init(smem_u32(b_full + s), 1);
            mbar_init(smem_u32(a_full + s), 128);
            mbar_init(smem_u32(ab_empty + s), 1);
        }
        mbar_init(smem_u32(tmem_full), 1);
        mbar_init_fence();
    }
    if (warp == 1) tc_alloc(smem_u32(tmem_slot), tmem_cols);
    for (int i = threadIdx.x; i < 2 * block_n; i += kThreads) s_stat[i] = 0.f;
    if (threadIdx.x == 0) {tma_prefetch_desc(&tmB); }      // descriptor fetch (~0.5 us) under the predecessor's tail
    pdl_wait();
    for (int c = threadIdx.x; c < KP; c += kThreads) {
        s_c[c] = c < K ? __ldg(in_scale + c) : 0.f;
        s_c[KP + c] = c < K ? __ldg(in_shift + c) : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {                                   // ---------------- TMA producer: weight slices
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kStagesB;
                const uint32_t phase = (kb / kStagesB) & 1;
                mbar_wait(smem_u32(ab_empty + s), phase ^ 1);
                mbar_expect_tx(smem_u32(b_full + s), b_bytes);
                tma_load_2d(smem_u32(sB + (size_t)s * b_pad), &tmB, smem_u32(b_full + s), kb * BK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                   // ---------------- MMA issuer
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(block_n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kStagesB;
                const uint32_t phase = (kb / kStagesB) & 1;
                mbar_wait(smem_u32(a_full + s), phase);
                mbar_wait(smem_u32(b_full + s), phase);
                tc_fence_after();
                const uint64_t adesc = make_desc_k_sw128(smem_u32(sA + (size_t)s * kABytes));
                const uint64_t bdesc = make_desc_k_sw128(smem_u32(sB + (size_t)s * b_pad));
                const int rem = K - kb * BK;
                const int k16 = rem >= BK ? BK / 16 : (rem + 15) / 16;
                for (int k = 0; k < k16; ++k)
                    umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (uint32_t)(kb > 0 || k > 0));
                umma_commit(smem_u32(ab_empty + s));
            }
            umma_commit(smem_u32(tmem_full));
        }
    } else {                                               // ---------------- A producers (BN + ReLU), then epilogue
        const int tt = threadIdx.x - 64;
        const int cg = tt & 7, r0 = tt >> 3;
        const bool col0 = z_out != nullptr && blockIdx.y == 0;
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % kStagesB;
            const uint32_t phase = (kb / kStagesB) & 1;
            const int c0 = kb * BK + cg * 8;
            const bool ch_in = c0 < K;
            float sc[8], sh[8];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                *reinterpret_cast<float4*>(sc + 4 * h) = *reinterpret_cast<const float4*>(s_c + c0 + 4 * h);
                *reinterpret_cast<float4*>(sh + 4 * h) = *reinterpret_cast<const float4*>(s_c + KP + c0 + 4 * h);
            }
            mbar_wait(smem_u32(ab_empty + s), phase ^ 1);
            uint8_t* a_tile = sA + (size_t)s * kABytes;
#pragma unroll
            for (int pass = 0; pass < 8 / kRowsInFlight; ++pass) {
                Raw8<bf16> rx[kRowsInFlight];
#pragma unroll
                for (int i = 0; i < kRowsInFlight; ++i) {
                    const int64_t m = m0 + r0 + 16 * (pass * kRowsInFlight + i);
                    if (ch_in && m < M) rx[i].ld(x + m * ldx + c0); else rx[i].zero();
                }
#pragma unroll
                for (int i = 0; i < kRowsInFlight; ++i) {
                    const int p = r0 + 16 * (pass * kRowsInFlight + i);
                    const int64_t m = m0 + p;
                    const bool live = ch_in && m < M;
                    float v[8], o[8];
                    rx[i].get(v);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        float t = fmaf(v[e], sc[e], sh[e]);
                        if (in_relu) t = fmaxf(t, 0.f);
                        o[e] = live ? t : 0.f;
                    }
                    uint4 u;
                    u.x = pack_bf16x2(o[0], o[1]); u.y = pack_bf16x2(o[2], o[3]);
                    u.z = pack_bf16x2(o[4], o[5]); u.w = pack_bf16x2(o[6], o[7]);
                    *reinterpret_cast<uint4*>(a_tile + (size_t)p * 128 + (size_t)((cg ^ (p & 7)) << 4)) = u;
                    if (col0 && live) *reinterpret_cast<uint4*>(z_out + m * ldz + c0) = u;
                }
            }
            fence_async_smem();
            mbar_arrive(smem_u32(a_full + s));
        }

        // ---------------- epilogue: statistics + raw output (rows >= M are exact zeros: their A rows are zero)
        const int q = warp & 3;
        const int row_in_tile = q * 32 + lane;
        const int64_t row = m0 + row_in_tile;
        const bool row_ok = row < M;
        mbar_wait(smem_u32(tmem_full), 0);
        tc_fence_after();
        for (int c = 0; c < block_n; c += 16) {
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
            if (stats != nullptr) {
                float sq[16], sm[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) { sm[i] = v[i]; sq[i] = v[i] * v[i]; }
                const float s1 = warp_transpose_sum16(sm, lane);
                const float s2 = warp_transpose_sum16(sq, lane);
                if ((lane & 1) == 0) {
                    atomicAdd(&s_stat[c + (lane >> 1)], s1);
                    atomicAdd(&s_stat[block_n + c + (lane >> 1)], s2);
                }
            }
            if (row_ok) {
                uint32_t o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
                bf16* dst = Y + row * ldy + n0 + c;
                *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<uint4*>(dst + 8) = make_uint4(o[4], o[5], o[6], o[7]);
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (stats != nullptr) {
        for (int i = threadIdx.x; i < block_n; i += kThreads) {
            atomicAdd(stats + n0 + i, (double)s_stat[i]);
            atomicAdd(stats + stats_stride + n0 + i, (double)s_stat[block_n + i]);
        }
    }
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        tc_dealloc(tmem_base, tmem_cols);
    }
}

}  // namespace

extern "C" int tss_pwconv_fwd_bnin(const void* x, int64_t ldx, const float* in_scale, const float* in_shift, int in_flags,
                                   void* z, int64_t ldz, const void* wp, void* y, int64_t ldy, int64_t M, int K, int Nc,
                                   double* stats, void* stream) {
    TSS_REQUIRE(M > 0 && K > 0 && Nc > 0 && K % 8 == 0 && K <= 2048 && Nc % 16 == 0, "pwconv_fwd_bnin: M=%lld K=%d Nc=%d", (long long)M, K, Nc);
    TSS_REQUIRE(ldx % 8 == 0 && ldy % 8 == 0 && (z == nullptr || ldz % 8 == 0), "pwconv_fwd_bnin: pitches must be multiples of 8");
    TSS_REQUIRE((((uintptr_t)x | (uintptr_t)y | (uintptr_t)z | (uintptr_t)wp) & 15) == 0, "pwconv_fwd_bnin: buffers must be 16-byte aligned");
    TSS_REQUIRE(in_scale != nullptr && in_shift != nullptr, "pwconv_fwd_bnin: missing input affine");
    int bn = 0;
    for (int b = 64; b >= 16; b -= 16)
        if (Nc % b == 0) { bn = b; break; }
    TSS_REQUIRE(bn >= 16, "pwconv_fwd_bnin: no tile width for Nc=%d", Nc);
    typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    tss_bind_context();
    static EncodeTiledFn enc = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    TSS_REQUIRE(enc != nullptr, "pwconv_fwd_bnin: cuTensorMapEncodeTiled is not available from the driver");
    CUtensorMap tmB;                                      // weights wp[Nc][K], K-major
    {
        cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)Nc};
        cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
        cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)bn};
        cuuint32_t estr[2] = {1, 1};
        TSS_REQUIRE((K * 2) % 16 == 0, "pwconv_fwd_bnin: weight pitch must be a multiple of 16 bytes");
        CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(wp), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        TSS_REQUIRE(r == CUDA_SUCCESS, "pwconv_fwd_bnin: cuTensorMapEncodeTiled failed (%d)", (int)r);
    }
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < bn) tmem_cols <<= 1;
    const int KP = (K + 63) & ~63;
    const uint32_t b_pad = ((uint32_t)bn * BK * 2 + 1023) & ~1023u;
    const size_t smem = 1024 + (size_t)kStagesB * (kABytes + b_pad) + (3 * kStagesB + 1) * 8 + 8 + (2 * bn + 2 * KP) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        TSS_CUDA(cudaFuncSetAttribute(pw_tc_fwd_bnin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    dim3 grid((unsigned)ceil_div64(M, BM), (unsigned)(Nc / bn));
    tss_launch(pw_tc_fwd_bnin_kernel, grid, kThreads, smem, (cudaStream_t)stream, tmB, (const bf16*)x, ldx, in_scale, in_shift,
               in_flags & TSS_EPI_RELU, (bf16*)z, ldz, (bf16*)y, ldy, M, K, bn, tmem_cols, stats, Nc);
    TSS_LAUNCH_CHECK("pwconv_fwd_bnin");
    return TSS_OK;
}
