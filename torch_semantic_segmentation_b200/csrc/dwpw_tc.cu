// Fused depthwise 3x3 (+folded BatchNorm, +ReLU) -> pointwise 1x1 (+folded BatchNorm, +residual, +ReLU)
// for INFERENCE (eval-mode BatchNorm): the tail of the inverted-residual bottleneck (conv2 + conv3,
// fastscnn.py:149-161 / contextnet.py:138-147) and the DS-conv block (fastscnn.py:188-199) as ONE
// kernel.  The depthwise output -- at 6x the block's width in a bottleneck -- never goes to HBM: it is
// produced tile by tile straight into the shared-memory A operand of the tcgen05 GEMM.
//
// One CTA = 8 x 16 output pixels (= the 128 rows of the UMMA tile) x all Nc <= 256 output channels,
// looping over the input channels in chunks of 64 (= one 128-byte swizzle row of the GEMM's K):
//   warp 0 (one lane)  TMA: the chunk's halo tile {64 ch, 18, 10} of x (zero fill = padding) and the
//                      chunk's Nc x 64 slice of the packed pointwise weights;
//   warps 2-5          depthwise: thread = (8 channels, tile column), 3-row register window down the 8
//                      rows (FFMA2), folded BN + ReLU, bf16, written as the K-major / 128B-swizzled A
//                      tile (row = pixel, 16-byte chunk j of row r at chunk j ^ (r & 7)), then
//                      fence.proxy.async + mbarrier arrive;  after the last chunk the same warps run
//                      the GEMM epilogue (tcgen05.ld, affine, residual, ReLU, 128-bit stores);
//   warp 1 (one lane)  tcgen05.mma (M=128, N=Nc, K=16 x 4 per chunk) into a TMEM accumulator.
// Two stages of {halo tile 23 KB, A tile 16 KB, B slice <= 32 KB}.  Stride 1, dilation 1, C % 64 == 0.
#include <cuda.h>

#include "tma.cuh"

namespace {

constexpr int TH = 8, TW = 16;               // output tile = 128 GEMM rows
constexpr int IH = TH + 2, IW = TW + 2;      // halo tile
constexpr int KC = 64;                       // channels per chunk = one 128-byte K row
constexpr int kStages = 2;
constexpr int kThreads = 192;
constexpr uint32_t kXBytes = IH * IW * KC * 2;
constexpr uint32_t kXPad = (kXBytes + 1023) & ~1023u;
constexpr uint32_t kABytes = 128 * KC * 2;

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row groups 1024 B apart (as pwconv_tc.cu)
__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__global__ void __launch_bounds__(kThreads, 2)
dwpw_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmB,
               const float* __restrict__ w_dw, const float* __restrict__ scale1, const float* __restrict__ shift1,
               int relu1, bf16* __restrict__ Y, int Ho, int Wo, int C, int Nc, int64_t ldy, int tiles_w, int tiles_h,
               const float* __restrict__ scale2, const float* __restrict__ shift2, const bf16* __restrict__ res,
               int64_t ldr, int relu2, uint32_t tmem_cols) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t b_bytes = (uint32_t)Nc * KC * 2;
    const uint32_t b_pad = (b_bytes + 1023) & ~1023u;
    uint8_t* sX = smem;                                        // [kStages][kXPad]
    uint8_t* sA = sX + (size_t)kStages * kXPad;                // [kStages][kABytes]   (1024-byte aligned)
    uint8_t* sB = sA + (size_t)kStages * kABytes;              // [kStages][b_pad]
    uint64_t* bars = (uint64_t*)(sB + (size_t)kStages * b_pad);
    // per stage: x_full, b_full, x_empty, a_full, ab_empty ; then tmem_full
    uint64_t* x_full = bars, *b_full = bars + kStages, *x_empty = bars + 2 * kStages, *a_full = bars + 3 * kStages,
            *ab_empty = bars + 4 * kStages, *tmem_full = bars + 5 * kStages;
    uint32_t* tmem_slot = (uint32_t*)(bars + 5 * kStages + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int t = blockIdx.x;
    const int tw = t % tiles_w; t /= tiles_w;
    const int th = t % tiles_h;
    const int n = t / tiles_h;
    const int ho0 = th * TH, wo0 = tw * TW;
    const int num_kb = C / KC;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(smem_u32(x_full + s), 1);
            mbar_init(smem_u32(b_full + s), 1);
            mbar_init(smem_u32(x_empty + s), 128);
            mbar_init(smem_u32(a_full + s), 128);
            mbar_init(smem_u32(ab_empty + s), 1);
        }
        mbar_init(smem_u32(tmem_full), 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) {tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmB); }      // descriptor fetch (~0.5 us) under the predecessor's tail
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {                                   // ---------------- TMA producer
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kStages;
                const uint32_t phase = (kb / kStages) & 1;
                mbar_wait(smem_u32(x_empty + s), phase ^ 1);
                mbar_expect_tx(smem_u32(x_full + s), kXBytes);
                tma_load_4d(smem_u32(sX + (size_t)s * kXPad), &tmX, smem_u32(x_full + s), kb * KC, wo0 - 1, ho0 - 1, n);
                mbar_wait(smem_u32(ab_empty + s), phase ^ 1);
                mbar_expect_tx(smem_u32(b_full + s), b_bytes);
                tma_load_2d(smem_u32(sB + (size_t)s * b_pad), &tmB, smem_u32(b_full + s), kb * KC, 0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                   // ---------------- MMA issuer
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(Nc >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kStages;
                const uint32_t phase = (kb / kStages) & 1;
                mbar_wait(smem_u32(a_full + s), phase);
                mbar_wait(smem_u32(b_full + s), phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t adesc = make_desc_k_sw128(smem_u32(sA + (size_t)s * kABytes));
                const uint64_t bdesc = make_desc_k_sw128(smem_u32(sB + (size_t)s * b_pad));
#pragma unroll
                for (int k = 0; k < KC / 16; ++k)
                    umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (uint32_t)(kb > 0 || k > 0));
                umma_commit(smem_u32(ab_empty + s));               // A and B of this stage are free once these retire
            }
            umma_commit(smem_u32(tmem_full));
        }
    } else {                                               // ---------------- depthwise producers, then epilogue
        const int tt = threadIdx.x - 64;                   // 0..127
        const int cg = tt & 7, col = tt >> 3;              // 8 channel groups x 16 tile columns
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % kStages;
            const uint32_t phase = (kb / kStages) & 1;
            const int c0 = kb * KC + cg * 8;
            float2 wr[9][4];
#pragma unroll
            for (int k = 0; k < 9; ++k)
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    wr[k][e] = make_float2(__ldg(w_dw + (c0 + 2 * e) * 9 + k), __ldg(w_dw + (c0 + 2 * e + 1) * 9 + k));
            mbar_wait(smem_u32(x_full + s), phase);
            const bf16* tp = (const bf16*)(sX + (size_t)s * kXPad) + (size_t)col * KC + cg * 8;
            float2 acc[TH][4];
#pragma unroll
            for (int r = 0; r < TH; ++r) zero8p(acc[r]);
#pragma unroll
            for (int j = 0; j < IH; ++j) {
                float2 v[3][4];
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) load8p_smem(tp + ((size_t)j * IW + kx) * KC, v[kx]);
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const int r = j - ky;
                    if (r >= 0 && r < TH) {
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                            for (int e = 0; e < 4; ++e) acc[r][e] = ffma2(v[kx][e], wr[ky * 3 + kx][e], acc[r][e]);
                    }
                }
            }
            mbar_arrive(smem_u32(x_empty + s));            // the halo tile has been read
            float2 sc[4], sh[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                sc[e] = scale1 != nullptr ? make_float2(__ldg(scale1 + c0 + 2 * e), __ldg(scale1 + c0 + 2 * e + 1)) : make_float2(1.f, 1.f);
                sh[e] = shift1 != nullptr ? make_float2(__ldg(shift1 + c0 + 2 * e), __ldg(shift1 + c0 + 2 * e + 1)) : make_float2(0.f, 0.f);
            }
            mbar_wait(smem_u32(ab_empty + s), phase ^ 1);  // the MMAs that read A[s] last time have retired
            uint8_t* a_tile = sA + (size_t)s * kABytes;
#pragma unroll
            for (int r = 0; r < TH; ++r) {
                uint4 u;
                uint32_t* up = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float2 o = ffma2(acc[r][e], sc[e], sh[e]);
                    if (relu1) o = make_float2(fmaxf(o.x, 0.f), fmaxf(o.y, 0.f));
                    up[e] = pack_bf16x2(o.x, o.y);
                }
                const int p = r * TW + col;                // GEMM row = pixel of the tile
                *reinterpret_cast<uint4*>(a_tile + (size_t)p * 128 + (size_t)((cg ^ (p & 7)) << 4)) = u;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to tcgen05
            mbar_arrive(smem_u32(a_full + s));
        }

        const int q = warp & 3;                            // TMEM lane quarter this warp may access
        const int p = q * 32 + lane;                       // GEMM row handled in the epilogue
        const int ho = ho0 + p / TW, wo = wo0 + p % TW;
        const bool row_ok = ho < Ho && wo < Wo;
        const int64_t row = ((int64_t)n * Ho + ho) * Wo + wo;
        mbar_wait(smem_u32(tmem_full), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int c = 0; c < Nc; c += 16) {
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
            if (row_ok) {
                if (shift2 != nullptr) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        v[i] = fmaf(v[i], scale2 != nullptr ? __ldg(scale2 + c + i) : 1.f, __ldg(shift2 + c + i));
                }
                if (res != nullptr) {
                    float r0[8], r1[8];
                    load8(res + row * ldr + c, r0);
                    load8(res + row * ldr + c + 8, r1);
#pragma unroll
                    for (int i = 0; i < 8; ++i) { v[i] += r0[i]; v[8 + i] += r1[i]; }
                }
                if (relu2) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
                }
                uint32_t o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
                bf16* dst = Y + row * ldy + c;
                *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<uint4*>(dst + 8) = make_uint4(o[4], o[5], o[6], o[7]);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

}  // namespace

extern "C" int tss_dwpw_fwd(const void* x, const float* w_dw, const float* scale1, const float* shift1, int flags1,
                            const void* wp, void* y, int N, int H, int W, int C, int Nc, int64_t ldy,
                            const float* scale2, const float* shift2, const void* res, int64_t ldr, int flags2,
                            void* stream) {
    TSS_REQUIRE(N > 0 && H > 0 && W > 0, "dwpw_fwd: empty tensor N=%d H=%d W=%d", N, H, W);
    TSS_REQUIRE(C > 0 && C % KC == 0, "dwpw_fwd: C=%d must be a multiple of %d", C, KC);
    TSS_REQUIRE(Nc >= 16 && Nc <= 256 && Nc % 16 == 0, "dwpw_fwd: Nc=%d must be a multiple of 16 in [16, 256]", Nc);
    TSS_REQUIRE(ldy % 8 == 0 && (res == nullptr || ldr % 8 == 0), "dwpw_fwd: output / residual pitch must be a multiple of 8");
    TSS_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)res & 15) == 0 && ((uintptr_t)wp & 15) == 0,
                "dwpw_fwd: buffers must be 16-byte aligned");
    TSS_REQUIRE(scale2 == nullptr || shift2 != nullptr, "dwpw_fwd: scale2 without shift2");
    TssEncodeTiledFn enc = tss_encode_tiled();
    TSS_REQUIRE(enc != nullptr, "dwpw_fwd: cuTensorMapEncodeTiled is not available from the driver");
    CUtensorMap mx, mb;
    {
        cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
        cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
        cuuint32_t box[4] = {KC, IW, IH, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&mx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        TSS_REQUIRE(r == CUDA_SUCCESS, "dwpw_fwd: cuTensorMapEncodeTiled(x) failed (%d)", (int)r);
    }
    {
        cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)Nc};
        cuuint64_t gstr[1] = {(cuuint64_t)C * 2};
        cuuint32_t box[2] = {KC, (cuuint32_t)Nc};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(wp), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        TSS_REQUIRE(r == CUDA_SUCCESS, "dwpw_fwd: cuTensorMapEncodeTiled(w) failed (%d)", (int)r);
    }
    const int tiles_w = (W + TW - 1) / TW, tiles_h = (H + TH - 1) / TH;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < Nc) tmem_cols <<= 1;
    const uint32_t b_pad = ((uint32_t)Nc * KC * 2 + 1023) & ~1023u;
    const size_t smem = 1024 + (size_t)kStages * (kXPad + kABytes + b_pad) + (5 * kStages + 1) * 8 + 16;
    static bool attr_set = false;
    if (!attr_set) {
        TSS_CUDA(cudaFuncSetAttribute(dwpw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    tss_launch(dwpw_tc_kernel, (unsigned)(N * tiles_h * tiles_w), kThreads, smem, (cudaStream_t)stream, mx, mb, w_dw, scale1,
               shift1, flags1 & TSS_EPI_RELU, (bf16*)y, H, W, C, Nc, ldy, tiles_w, tiles_h, scale2, shift2, (const bf16*)res,
               ldr, flags2 & TSS_EPI_RELU, tmem_cols);
    TSS_LAUNCH_CHECK("dwpw_fwd");
    return TSS_OK;
}
