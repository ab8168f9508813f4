// Pointwise-conv backward on tcgen05 with BOTH BatchNorm-backward passes of the neighbouring layers fused
// around the dgrad GEMM (training):
//
//   prologue  dy = gamma*rstd * (g - sum(g)/m - xhat * sum(g*xhat)/m)      BatchNorm-backward APPLY of THIS layer
//             (g = dz * relu-mask, mask recomputed from the raw conv output y; bn.cu: bn_bwd_apply_kernel)
//   GEMM      dz_in[M][K] = dy[M][Nc] . W[Nc][K]                             (pwconv_tc.cu main loop)
//   epilogue  g_in = dz_in * mask(yp), sums_p += ...                         BatchNorm-backward REDUCTION of the
//             PRODUCER of this conv's input (pwconv_tc_bnred.cu), optional
//
// dy is never read back from HBM by the dgrad: the four epilogue warps build the GEMM's A operand themselves,
// chunk by chunk (128 rows x 64 channels): 128-bit loads of dz and y, the BatchNorm arithmetic in registers,
// bf16, written in the K-major 128-byte-swizzled UMMA layout (chunk j of row r at j ^ (r & 7)), then
// fence.proxy.async + mbarrier (the pattern of dwpw_tc.cu).  The column-tile-0 CTAs also store dy for the
// weight-gradient kernel, which runs on the side stream.  One launch (instead of apply + dgrad) and one full
// read of dy less per pointwise layer.
#include <stdlib.h>

#include "tc_ptx.cuh"

namespace {

constexpr int BM = 128;          // rows per tile = UMMA M
constexpr int BK = 64;           // bf16 elements per k-block = 128 bytes = one swizzle row
constexpr int kThreads = 192;
constexpr uint32_t kABytes = BM * BK * 2;

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



constexpr int kStagesB = 2;
constexpr int kRowsInFlight = 4;      // rows of (dz, y) a producer thread has in flight per pass (of its 8 per chunk)

__global__ void __launch_bounds__(kThreads)
pw_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmB, const bf16* __restrict__ dz, const bf16* __restrict__ y,
                 int64_t lddz, int64_t ldy, const float* __restrict__ mean, const float* __restrict__ rstd,
                 const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ sums,
                 int relu, float inv_count, bf16* __restrict__ dy_out, int64_t lddy, float* __restrict__ dgamma,
                 float* __restrict__ dbeta, bf16* __restrict__ DX, int64_t M, int Nc, int64_t lddx, int block_n,
                 uint32_t tmem_cols,
                 const bf16* __restrict__ yp, int64_t ldyp, const float* __restrict__ pmean, const float* __restrict__ prstd,
                 const float* __restrict__ pgamma, const float* __restrict__ pbeta, int prelu, float* __restrict__ psums,
                 int psums_stride) {
    TSS_DYN_SMEM(uint8_t, smem_raw);
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t b_bytes = (uint32_t)block_n * BK * 2;
    const uint32_t b_pad = (b_bytes + 1023) & ~1023u;
    uint8_t* sA = smem;                                               // [kStagesB][kABytes]
    uint8_t* sB = sA + (size_t)kStagesB * kABytes;                     // [kStagesB][b_pad]
    uint64_t* bars = (uint64_t*)(sB + (size_t)kStagesB * b_pad);
    uint64_t* b_full = bars, *a_full = bars + kStagesB, *ab_empty = bars + 2 * kStagesB, *tmem_full = bars + 3 * kStagesB;
    uint32_t* tmem_slot = (uint32_t*)(bars + 3 * kStagesB + 1);
    float* s_stat = (float*)(tmem_slot + 2);                           // [2][block_n]          (producer reduction)
    float* s_pc = s_stat + 2 * block_n;                                // [4][block_n]          (producer constants)
    const int NcP = (Nc + 63) & ~63;
    float* s_c = s_pc + 4 * block_n;                                   // [4][NcP]: A, SH, B, D of THIS layer (below)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * block_n;
    const int num_kb = (Nc + BK - 1) / BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStagesB; ++s) {
            mbar_init(smem_u32(b_full + s), 1);
            mbar_init(smem_u32(a_full + s), 128);
            mbar_init(smem_u32(ab_empty + s), 1);
        }
        mbar_init(smem_u32(tmem_full), 1);
        mbar_init_fence();
    }
    if (warp == 1) {
        tc_alloc(smem_u32(tmem_slot), tmem_cols);
    }
    for (int i = threadIdx.x; i < 2 * block_n; i += kThreads) s_stat[i] = 0.f;
    if (threadIdx.x == 0) {tma_prefetch_desc(&tmB); }      // descriptor fetch (~0.5 us) under the predecessor's tail
    pdl_wait();
    // this layer's BatchNorm backward as dy = A*g + B*y + D per channel (A = gamma*rstd, B = -rstd*k2,
    // D = mean*rstd*k2 - k1 with k1 = A*sum(g)/m, k2 = A*sum(g*xhat)/m), and SH for the ReLU mask fma(y, A, SH) > 0
    for (int c = threadIdx.x; c < NcP; c += kThreads) {
        float A = 0.f, SH = 0.f, B = 0.f, D = 0.f;
        if (c < Nc) {
            const float mu = __ldg(mean + c), rs = __ldg(rstd + c);
            A = (gamma != nullptr ? __ldg(gamma + c) : 1.f) * rs;
            SH = (beta != nullptr ? __ldg(beta + c) : 0.f) - mu * A;
            const float a1 = __ldg(sums + c), a2 = __ldg(sums + Nc + c);
            const float k2 = A * a2 * inv_count;
            B = -rs * k2;
            D = fmaf(mu * rs, k2, -A * a1 * inv_count);
            if (blockIdx.x == 0 && blockIdx.y == 0) {
                if (dbeta != nullptr) dbeta[c] += a1;
                if (dgamma != nullptr) dgamma[c] += a2;
            }
        }
        s_c[c] = A; s_c[NcP + c] = SH; s_c[2 * NcP + c] = B; s_c[3 * NcP + c] = D;
    }
    if (yp != nullptr) {
        for (int i = threadIdx.x; i < block_n; i += kThreads) {        // constants of the producer's BatchNorm (reduction)
            const float mu = __ldg(pmean + n0 + i), rs = __ldg(prstd + n0 + i);
            const float sc = (pgamma != nullptr ? __ldg(pgamma + n0 + i) : 1.f) * rs;
            s_pc[i] = mu;
            s_pc[block_n + i] = rs;
            s_pc[2 * block_n + i] = sc;
            s_pc[3 * block_n + i] = (pbeta != nullptr ? __ldg(pbeta + n0 + i) : 0.f) - mu * sc;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {                                   // ---------------- TMA producer: weight slices only
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
                int rem = Nc - kb * BK;
                const int k16 = rem >= BK ? BK / 16 : (rem + 15) / 16;
                for (int k = 0; k < k16; ++k)
                    umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (uint32_t)(kb > 0 || k > 0));
                umma_commit(smem_u32(ab_empty + s));
            }
            umma_commit(smem_u32(tmem_full));
        }
    } else {                                               // ---------------- A producers (BN apply), then epilogue
        const int tt = threadIdx.x - 64;                   // 0..127
        const int cg = tt & 7, r0 = tt >> 3;               // 8 channel groups x 16 rows; 8 row passes per chunk
        const bool col0 = dy_out != nullptr && blockIdx.y == 0;
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % kStagesB;
            const uint32_t phase = (kb / kStagesB) & 1;
            const int c0 = kb * BK + cg * 8;
            const bool ch_in = c0 < Nc;
            float A[8], SH[8], B[8], D[8];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                *reinterpret_cast<float4*>(A + 4 * h) = *reinterpret_cast<const float4*>(s_c + c0 + 4 * h);
                *reinterpret_cast<float4*>(SH + 4 * h) = *reinterpret_cast<const float4*>(s_c + NcP + c0 + 4 * h);
                *reinterpret_cast<float4*>(B + 4 * h) = *reinterpret_cast<const float4*>(s_c + 2 * NcP + c0 + 4 * h);
                *reinterpret_cast<float4*>(D + 4 * h) = *reinterpret_cast<const float4*>(s_c + 3 * NcP + c0 + 4 * h);
            }
            mbar_wait(smem_u32(ab_empty + s), phase ^ 1);  // the MMAs that read A[s] last time have retired
            uint8_t* a_tile = sA + (size_t)s * kABytes;
#pragma unroll
            for (int pass = 0; pass < 8 / kRowsInFlight; ++pass) {
                Raw8<bf16> rg[kRowsInFlight], ry[kRowsInFlight];
#pragma unroll
                for (int i = 0; i < kRowsInFlight; ++i) {
                    const int64_t m = m0 + r0 + 16 * (pass * kRowsInFlight + i);
                    if (ch_in && m < M) {
                        rg[i].ld(dz + m * lddz + c0);
                        ry[i].ld(y + m * ldy + c0);
                    } else {
                        rg[i].zero(); ry[i].zero();
                    }
                }
#pragma unroll
                for (int i = 0; i < kRowsInFlight; ++i) {
                    const int p = r0 + 16 * (pass * kRowsInFlight + i);        // GEMM row in the tile
                    const int64_t m = m0 + p;
                    const bool live = ch_in && m < M;
                    float g[8], yy[8], o[8];
                    rg[i].get(g);
                    ry[i].get(yy);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        float gg = g[e];
                        if (relu && !(fmaf(yy[e], A[e], SH[e]) > 0.f)) gg = 0.f;
                        o[e] = live ? fmaf(A[e], gg, fmaf(B[e], yy[e], D[e])) : 0.f;
                    }
                    uint4 u;
                    u.x = pack_bf16x2(o[0], o[1]); u.y = pack_bf16x2(o[2], o[3]);
                    u.z = pack_bf16x2(o[4], o[5]); u.w = pack_bf16x2(o[6], o[7]);
                    *reinterpret_cast<uint4*>(a_tile + (size_t)p * 128 + (size_t)((cg ^ (p & 7)) << 4)) = u;
                    if (col0 && live) *reinterpret_cast<uint4*>(dy_out + m * lddy + c0) = u;
                }
            }
            fence_async_smem();     // generic-proxy writes -> visible to tcgen05
            mbar_arrive(smem_u32(a_full + s));
        }

        // ---------------- epilogue: dz_in (or g_in + the producer's reduction)
        const int q = warp & 3;
        const int row_in_tile = q * 32 + lane;
        const int64_t row = m0 + row_in_tile;
        const bool row_ok = row < M;
        mbar_wait(smem_u32(tmem_full), 0);
        tc_fence_after();
        for (int c = 0; c < block_n; c += 16) {
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
            const int col = n0 + c;
            if (yp != nullptr) {
                float y0[8], y1[8];
                if (row_ok) {
                    load8(yp + row * ldyp + col, y0);
                    load8(yp + row * ldyp + col + 8, y1);
                } else {
                    zero8(y0); zero8(y1);
                }
                float gx[16], sm[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float yy = i < 8 ? y0[i] : y1[i - 8];
                    if (prelu && !(fmaf(yy, s_pc[2 * block_n + c + i], s_pc[3 * block_n + c + i]) > 0.f)) v[i] = 0.f;
                    if (!row_ok) v[i] = 0.f;
                    gx[i] = v[i] * ((yy - s_pc[c + i]) * s_pc[block_n + c + i]);
                    sm[i] = v[i];
                }
                const float s1 = warp_transpose_sum16(sm, lane);
                const float s2 = warp_transpose_sum16(gx, lane);
                if ((lane & 1) == 0) {
                    atomicAdd(&s_stat[c + (lane >> 1)], s1);
                    atomicAdd(&s_stat[block_n + c + (lane >> 1)], s2);
                }
            }
            if (row_ok) {
                uint32_t o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
                bf16* dst = DX + row * lddx + col;
                *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<uint4*>(dst + 8) = make_uint4(o[4], o[5], o[6], o[7]);
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (yp != nullptr) {
        for (int i = threadIdx.x; i < block_n; i += kThreads) {
            atomicAdd(psums + n0 + i, s_stat[i]);
            atomicAdd(psums + psums_stride + n0 + i, s_stat[block_n + i]);
        }
    }
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        tc_dealloc(tmem_base, tmem_cols);
    }
}

int make_map_bwd(CUtensorMap* map, const void* base, int64_t rows, int cols, int64_t ld, int box_rows) {
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
    TSS_REQUIRE(enc != nullptr, "pwconv_bwd_fused: cuTensorMapEncodeTiled is not available from the driver");
    TSS_REQUIRE(((uintptr_t)base & 15) == 0 && (ld * 2) % 16 == 0, "pwconv_bwd_fused: TMA needs 16-byte aligned base and pitch");
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TSS_REQUIRE(r == CUDA_SUCCESS, "pwconv_bwd_fused: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return TSS_OK;
}

}  // namespace

extern "C" int tss_pwconv_bwd_fused(const void* dz, const void* y, int64_t lddz, int64_t ldy, const float* mean,
                                    const float* rstd, const float* gamma, const float* beta, const float* sums,
                                    int flags, int64_t count, void* dy, int64_t lddy, float* dgamma, float* dbeta,
                                    const void* wpT, void* dx, int64_t M, int K, int Nc, int64_t lddx,
                                    const void* yp, int64_t ldyp, const float* pmean, const float* prstd,
                                    const float* pgamma, const float* pbeta, int pflags, float* psums, void* stream) {
    TSS_REQUIRE(M > 0 && K > 0 && Nc > 0 && K % 16 == 0 && Nc % 8 == 0 && Nc <= 1024, "pwconv_bwd_fused: M=%lld K=%d Nc=%d", (long long)M, K, Nc);
    TSS_REQUIRE(lddz % 8 == 0 && ldy % 8 == 0 && lddx % 8 == 0 && (dy == nullptr || lddy % 8 == 0) && (yp == nullptr || ldyp % 8 == 0),
                "pwconv_bwd_fused: pitches must be multiples of 8");
    TSS_REQUIRE((((uintptr_t)dz | (uintptr_t)y | (uintptr_t)dy | (uintptr_t)dx | (uintptr_t)yp | (uintptr_t)wpT) & 15) == 0,
                "pwconv_bwd_fused: buffers must be 16-byte aligned");
    TSS_REQUIRE(mean != nullptr && rstd != nullptr && sums != nullptr, "pwconv_bwd_fused: missing BatchNorm operands");
    TSS_REQUIRE(yp == nullptr || (pmean != nullptr && prstd != nullptr && psums != nullptr), "pwconv_bwd_fused: missing producer operands");
    if (count <= 0) count = M;
    int bn = 0;
    for (int b = 64; b >= 16; b -= 16)
        if (K % b == 0) { bn = b; break; }
    TSS_REQUIRE(bn >= 16, "pwconv_bwd_fused: no tile width for K=%d", K);
    CUtensorMap tmB;
    if (int e = make_map_bwd(&tmB, wpT, K, Nc, Nc, bn)) return e;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < bn) tmem_cols <<= 1;
    const int NcP = (Nc + 63) & ~63;
    const uint32_t b_pad = ((uint32_t)bn * BK * 2 + 1023) & ~1023u;
    const size_t smem = 1024 + (size_t)kStagesB * (kABytes + b_pad) + (3 * kStagesB + 1) * 8 + 8 + (6 * bn + 4 * NcP) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        TSS_CUDA(cudaFuncSetAttribute(pw_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    dim3 grid((unsigned)ceil_div64(M, BM), (unsigned)(K / bn));
    tss_launch(pw_tc_bwd_kernel, grid, kThreads, smem, (cudaStream_t)stream, tmB, (const bf16*)dz, (const bf16*)y, lddz, ldy, mean, rstd,
               gamma, beta, sums, flags & TSS_EPI_RELU, (float)(1.0 / (double)count), (bf16*)dy, lddy, dgamma, dbeta, (bf16*)dx, M, Nc,
               lddx, bn, tmem_cols, (const bf16*)yp, ldyp, pmean, prstd, pgamma, pbeta, pflags & TSS_EPI_RELU, psums, K);
    TSS_LAUNCH_CHECK("pwconv_bwd_fused");
    return TSS_OK;
}
