// Depthwise 3x3 (stride 1) backward with BOTH neighbouring BatchNorm-backward passes fused around the dgrad
// (training), the depthwise counterpart of pwconv_tc_bwd.cu:
//
//   prologue  dy = gamma*rstd * (g - sum(g)/m - xhat * sum(g*xhat)/m)      BatchNorm-backward APPLY of THIS layer,
//             g = dz * relu-mask(y); dz and y arrive as two TMA halo tiles and dy replaces dz in shared memory
//             (halo positions outside the image are forced to zero: a zero-filled (dz, y) pair would give D != 0)
//   stencil   dz_in = dy (*) flipped taps                                    (dwconv_tma.cu / dwconv_bnred.cu)
//   epilogue  g_in = dz_in * mask(yp), sums_p += ...                         reduction of the PRODUCER (dwconv_bnred.cu)
//
// The interior of the dy tile is also stored for the weight-gradient kernel (side stream).  One launch instead of
// bn_bwd_apply + dgrad, and dy is read back from HBM once (wgrad) instead of twice.
#include "tma.cuh"

namespace {

constexpr int TH = 8;
constexpr int IH = TH + 2;

template <typename T>
__global__ void __launch_bounds__(192, 2)
dw_bwd_fused_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmY,
                    const float* __restrict__ w, T* __restrict__ g_out, T* __restrict__ dy_out, int H, int W, int C,
                    int CB, int TW, int tiles_w, int tiles_h, const float* __restrict__ mean, const float* __restrict__ rstd,
                    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ sums,
                    int relu, float inv_count, float* __restrict__ dgamma, float* __restrict__ dbeta,
                    const T* __restrict__ yp, const float* __restrict__ pmean, const float* __restrict__ prstd,
                    const float* __restrict__ pgamma, const float* __restrict__ pbeta, int prelu, float* __restrict__ psums) {
    TSS_DYN_SMEM(uint8_t, smem_raw);
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    const int IW = TW + 2;
    const uint32_t tile_bytes = (uint32_t)IH * IW * CB * sizeof(T);
    const uint32_t tile_pad = (tile_bytes + 127) & ~127u;
    T* tile = (T*)smem;                                   // dz, then dy in place
    T* tile_y = (T*)(smem + tile_pad);
    uint64_t* bar = (uint64_t*)(smem + 2 * (size_t)tile_pad);

    int t = blockIdx.x;
    const int tw = t % tiles_w; t /= tiles_w;
    const int th = t % tiles_h;
    const int n = t / tiles_h;
    const int cb0 = blockIdx.y * CB;
    const int h0 = th * TH, w0 = tw * TW;

    if (threadIdx.x == 0) {
        mbar_init(smem_u32(bar), 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {tma_prefetch_desc(&tmG); tma_prefetch_desc(&tmY); }      // descriptor fetch (~0.5 us) under the predecessor's tail
    pdl_wait();
    if (threadIdx.x == 0) {
        mbar_expect_tx(smem_u32(bar), 2 * tile_bytes);
        tma_load_4d(smem_u32(tile), &tmG, smem_u32(bar), cb0, w0 - 1, h0 - 1, n);
        tma_load_4d(smem_u32(tile_y), &tmY, smem_u32(bar), cb0, w0 - 1, h0 - 1, n);
    }

    const int CGB = CB >> 3;
    const int cg = threadIdx.x % CGB, col = threadIdx.x / CGB;
    const int c0 = cb0 + cg * 8;
    {
        // this layer's BatchNorm backward as dy = A*g + B*y + D per channel (pwconv_tc_bwd.cu), SH for the mask
        float A[8], SH[8], B[8], D[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float mu = __ldg(mean + c0 + e), rs = __ldg(rstd + c0 + e);
            A[e] = (gamma != nullptr ? __ldg(gamma + c0 + e) : 1.f) * rs;
            SH[e] = (beta != nullptr ? __ldg(beta + c0 + e) : 0.f) - mu * A[e];
            const float a1 = __ldg(sums + c0 + e), a2 = __ldg(sums + C + c0 + e);
            const float k2 = A[e] * a2 * inv_count;
            B[e] = -rs * k2;
            D[e] = fmaf(mu * rs, k2, -A[e] * a1 * inv_count);
            if (blockIdx.x == 0 && col == 0) {            // one thread per channel and channel block
                if (dbeta != nullptr) dbeta[c0 + e] += a1;
                if (dgamma != nullptr) dgamma[c0 + e] += a2;
            }
        }
        mbar_wait(smem_u32(bar), 0);
        const int ncols = blockDim.x / CGB;               // = TW
        for (int p = col; p < IH * IW; p += ncols) {      // every halo-tile position of this thread's channel group
            const int j = p / IW, i = p - j * IW;
            const int h = h0 - 1 + j, x = w0 - 1 + i;
            const bool inside = h >= 0 && h < H && x >= 0 && x < W;
            T* gp = tile + (size_t)p * CB + cg * 8;
            float g[8], yy[8], o[8];
            load8_smem(gp, g);
            load8_smem(tile_y + (size_t)p * CB + cg * 8, yy);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float gg = g[e];
                if (relu && !(fmaf(yy[e], A[e], SH[e]) > 0.f)) gg = 0.f;
                o[e] = inside ? fmaf(A[e], gg, fmaf(B[e], yy[e], D[e])) : 0.f;
            }
            store8(gp, o);
            if (dy_out != nullptr && inside && j >= 1 && j <= TH && i >= 1 && i <= TW)
                store8(dy_out + (((int64_t)n * H + h) * W + x) * C + c0, o);
        }
    }
    __syncthreads();                                      // the dy tile is complete

    float2 wr[9][4];                                      // flipped taps: dgrad of a stride-1 correlation
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
        for (int e = 0; e < 4; ++e)
            wr[k][e] = make_float2(__ldg(w + (c0 + 2 * e) * 9 + (8 - k)), __ldg(w + (c0 + 2 * e + 1) * 9 + (8 - k)));
    float2 acc[TH][4];
#pragma unroll
    for (int r = 0; r < TH; ++r) zero8p(acc[r]);
    const T* tp = tile + (size_t)col * CB + cg * 8;
#pragma unroll
    for (int j = 0; j < IH; ++j) {
        float2 v[3][4];
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) load8p_smem(tp + ((size_t)j * IW + kx) * CB, v[kx]);
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

    const int wo = w0 + col;
    if (yp == nullptr) {                                  // no producer to reduce for: the plain input gradient
        if (wo < W) {
            const int64_t base = (((int64_t)n * H + h0) * W + wo) * C + c0;
#pragma unroll
            for (int r = 0; r < TH; ++r)
                if (h0 + r < H) store8p(g_out + base + (int64_t)r * W * C, acc[r]);
        }
        return;                                           // uniform over the grid
    }
    float mu[8], rs[8], sc[8], sh[8];                     // the producer's BatchNorm (reduction)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        mu[e] = __ldg(pmean + c0 + e);
        rs[e] = __ldg(prstd + c0 + e);
        sc[e] = (pgamma != nullptr ? __ldg(pgamma + c0 + e) : 1.f) * rs[e];
        sh[e] = (pbeta != nullptr ? __ldg(pbeta + c0 + e) : 0.f) - mu[e] * sc[e];
    }
    float s1[8], s2[8];
    zero8(s1); zero8(s2);
    if (wo < W) {
        const int64_t base = (((int64_t)n * H + h0) * W + wo) * C + c0;
#pragma unroll
        for (int r = 0; r < TH; ++r) {
            if (h0 + r < H) {
                float yy[8], g[8];
                load8(yp + base + (int64_t)r * W * C, yy);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float dz = (e & 1) ? acc[r][e >> 1].y : acc[r][e >> 1].x;
                    const bool on = !prelu || fmaf(yy[e], sc[e], sh[e]) > 0.f;
                    g[e] = on ? dz : 0.f;
                    s1[e] += g[e];
                    s2[e] = fmaf(g[e], (yy[e] - mu[e]) * rs[e], s2[e]);
                }
                store8(g_out + base + (int64_t)r * W * C, g);
            }
        }
    }
    // column partials -> one atomic per channel and CTA (the dy tile is dead: reuse it)
    __syncthreads();
    float* part = (float*)tile;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        part[(size_t)col * CB + cg * 8 + e] = s1[e];
        part[(size_t)(TW + col) * CB + cg * 8 + e] = s2[e];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * CB; i += blockDim.x) {
        const int which = i / CB, ch = i - which * CB;
        float s = 0.f;
        for (int cidx = 0; cidx < TW; ++cidx) s += part[(size_t)(which * TW + cidx) * CB + ch];
        atomicAdd(psums + which * C + cb0 + ch, s);
    }
}

template <typename T> struct TmaTypeF;
template <> struct TmaTypeF<float> { static constexpr CUtensorMapDataType v = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; };
template <> struct TmaTypeF<bf16> { static constexpr CUtensorMapDataType v = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; };

}  // namespace

bool tss_dw_tma_config(int C, int* CB, int* TW);     // dwconv_tma.cu

extern "C" int tss_dwconv3x3_bwd_fused(const void* dz, const void* y, const float* w, const float* mean,
                                       const float* rstd, const float* gamma, const float* beta, const float* sums,
                                       int flags, int64_t count, void* dy, float* dgamma, float* dbeta, void* g, int N,
                                       int H, int W, int C, const void* yp, const float* pmean, const float* prstd,
                                       const float* pgamma, const float* pbeta, int pflags, float* psums, int dtype,
                                       void* stream) {
    TSS_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "dwconv3x3_bwd_fused: bad shape N=%d H=%d W=%d C=%d", N, H, W, C);
    TSS_REQUIRE(mean != nullptr && rstd != nullptr && sums != nullptr, "dwconv3x3_bwd_fused: missing BatchNorm operands");
    TSS_REQUIRE(yp == nullptr || (pmean != nullptr && prstd != nullptr && psums != nullptr), "dwconv3x3_bwd_fused: missing producer operands");
    TSS_REQUIRE((((uintptr_t)dz | (uintptr_t)y | (uintptr_t)dy | (uintptr_t)g | (uintptr_t)yp) & 15) == 0,
                "dwconv3x3_bwd_fused: buffers must be 16-byte aligned");
    if (count <= 0) count = (int64_t)N * H * W;
    int CB, TW;
    TSS_REQUIRE(tss_dw_tma_config(C, &CB, &TW), "dwconv3x3_bwd_fused: no channel block for C=%d", C);
    TssEncodeTiledFn enc = tss_encode_tiled();
    TSS_REQUIRE(enc != nullptr, "dwconv3x3_bwd_fused: cuTensorMapEncodeTiled is not available from the driver");
    TSS_DISPATCH_DTYPE(dtype, "dwconv3x3_bwd_fused", {
        if (sizeof(T) == 4 && TW == 32) TW = 16;
        const int IW = TW + 2;
        CUtensorMap maps[2];
        const void* bases[2] = {dz, y};
        for (int i = 0; i < 2; ++i) {
            cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
            cuuint64_t gstr[3] = {(cuuint64_t)C * sizeof(T), (cuuint64_t)W * C * sizeof(T), (cuuint64_t)H * W * C * sizeof(T)};
            cuuint32_t box[4] = {(cuuint32_t)CB, (cuuint32_t)IW, (cuuint32_t)IH, 1};
            cuuint32_t estr[4] = {1, 1, 1, 1};
            CUresult r = enc(&maps[i], TmaTypeF<T>::v, 4, const_cast<void*>(bases[i]), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            TSS_REQUIRE(r == CUDA_SUCCESS, "dwconv3x3_bwd_fused: cuTensorMapEncodeTiled failed (%d)", (int)r);
        }
        const int tiles_w = (W + TW - 1) / TW, tiles_h = (H + TH - 1) / TH;
        const int threads = (CB / 8) * TW;
        const size_t tile_pad = ((size_t)IH * IW * CB * sizeof(T) + 127) & ~(size_t)127;
        size_t smem = 128 + 2 * tile_pad + 16;
        const size_t part = (size_t)2 * TW * CB * sizeof(float) + 128;
        if (smem < part) smem = part;
        auto kern = dw_bwd_fused_kernel<T>;
        static bool attr_set = false;
        if (!attr_set) {
            TSS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            attr_set = true;
        }
        dim3 grid((unsigned)((int64_t)N * tiles_h * tiles_w), (unsigned)(C / CB));
        tss_launch(kern, grid, threads, smem, (cudaStream_t)stream, maps[0], maps[1], w, (T*)g, (T*)dy, H, W, C, CB, TW, tiles_w, tiles_h,
                   mean, rstd, gamma, beta, sums, flags & TSS_EPI_RELU, (float)(1.0 / (double)count), dgamma, dbeta, (const T*)yp, pmean,
                   prstd, pgamma, pbeta, pflags & TSS_EPI_RELU, psums);
        TSS_LAUNCH_CHECK("dwconv3x3_bwd_fused");
        return TSS_OK;
    });
}
