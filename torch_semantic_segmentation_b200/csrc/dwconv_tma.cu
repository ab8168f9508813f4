// Depthwise 3x3 forward (and stride-1 dgrad) with TMA-staged halo tiles.
//
// One CTA = TH x TW output pixels x CB channels.  A single elected thread issues ONE 4-D
// cp.async.bulk.tensor box load {CB, IW, IH, 1} of the input halo tile into shared memory;
// the box may start at negative coordinates / run past the edge -- TMA zero-fills, which IS
// the convolution's zero padding.  While the copy is in flight every thread loads its 72
// weights; then thread = (8-channel group, output column) slides a 3-row register window
// down the TH rows reading 128-bit vectors from shared memory (conflict-free: consecutive
// lanes read consecutive 16 B), and stores 128-bit results.  No register staging of global
// loads, so HBM latency is covered by the other CTAs resident on the SM (3-4 per SM).
// Algorithmic traffic: 2*C*(in + out pixels) bytes in bf16; halo re-reads hit L2.
#include <stdlib.h>

#include "tma.cuh"
#include "dw_fwd_persistent.cuh"

namespace {

template <int S, int D, int TH>
struct Geo {
    static constexpr int IH = (TH - 1) * S + 2 * D + 1;
};

template <typename T, int S, int D, int TH, bool FLIP>
__global__ void __launch_bounds__(192, 2)
dw_tma_kernel(const __grid_constant__ CUtensorMap tmX, const float* __restrict__ w, T* __restrict__ y,
              int Ho, int Wo, int C, int CB, int TW, int tiles_w, int tiles_h,
              const float* __restrict__ scale, const float* __restrict__ shift, int flags,
              double* __restrict__ stats) {
    constexpr int IH = Geo<S, D, TH>::IH;
    TSS_DYN_SMEM(uint8_t, smem_raw);
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    const int IW = (TW - 1) * S + 2 * D + 1;
    const uint32_t tile_bytes = (uint32_t)IH * IW * CB * sizeof(T);
    T* tile = (T*)smem;
    uint64_t* bar = (uint64_t*)(smem + ((tile_bytes + 15) & ~15u));
    float* s_stat = (float*)(bar + 1);                       // [2][CB]

    int t = blockIdx.x;
    const int tw = t % tiles_w; t /= tiles_w;
    const int th = t % tiles_h;
    const int n = t / tiles_h;
    const int cb0 = blockIdx.y * CB;
    const int ho0 = th * TH, wo0 = tw * TW;

    if (threadIdx.x == 0) {
        mbar_init(smem_u32(bar), 1);
        mbar_fence_init();
    }
    if (stats != nullptr)
        for (int i = threadIdx.x; i < 2 * CB; i += blockDim.x) s_stat[i] = 0.f;
    __syncthreads();
    if (threadIdx.x == 0) {tma_prefetch_desc(&tmX); }      // descriptor fetch (~0.5 us) under the predecessor's tail
    pdl_wait();              // barrier init / smem zeroing above overlap the previous kernel's tail
    if (threadIdx.x == 0) {
        mbar_expect_tx(smem_u32(bar), tile_bytes);
        tma_load_4d(smem_u32(tile), &tmX, smem_u32(bar), cb0, wo0 * S - D, ho0 * S - D, n);
    }

    const int CGB = CB >> 3;
    const int cg = threadIdx.x % CGB, col = threadIdx.x / CGB;     // col < TW by construction
    const int c0 = cb0 + cg * 8;
    float2 wr[9][4];
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
        for (int e = 0; e < 4; ++e)
            wr[k][e] = make_float2(__ldg(w + (c0 + 2 * e) * 9 + (FLIP ? 8 - k : k)), __ldg(w + (c0 + 2 * e + 1) * 9 + (FLIP ? 8 - k : k)));

    mbar_wait(smem_u32(bar), 0);

    float2 acc[TH][4];
#pragma unroll
    for (int r = 0; r < TH; ++r) zero8p(acc[r]);
    const T* tp = tile + ((size_t)col * S) * CB + cg * 8;
#pragma unroll
    for (int j = 0; j < IH; ++j) {
        bool used = false;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int tt = j - ky * D;
            if (tt >= 0 && tt % S == 0 && tt / S < TH) used = true;
        }
        if (!used) continue;
        float2 v[3][4];
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) load8p_smem(tp + ((size_t)j * IW + kx * D) * CB, v[kx]);
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int tt = j - ky * D;
            if (tt >= 0 && tt % S == 0 && tt / S < TH) {
                const int r = tt / S;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[r][e] = ffma2(v[kx][e], wr[ky * 3 + kx][e], acc[r][e]);
            }
        }
    }

    const int wo = wo0 + col;
    const bool relu = (flags & TSS_EPI_RELU) != 0;
    float2 s1[4], s2[4];
    zero8p(s1); zero8p(s2);
    if (wo < Wo) {
        T* yp = y + (((int64_t)n * Ho + ho0) * Wo + wo) * C + c0;
        float2 sc[4], sh[4];
        if (shift != nullptr) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                sc[e] = scale != nullptr ? make_float2(__ldg(scale + c0 + 2 * e), __ldg(scale + c0 + 2 * e + 1)) : make_float2(1.f, 1.f);
                sh[e] = make_float2(__ldg(shift + c0 + 2 * e), __ldg(shift + c0 + 2 * e + 1));
            }
        }
#pragma unroll
        for (int r = 0; r < TH; ++r) {
            if (ho0 + r < Ho) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    s1[e].x += acc[r][e].x; s1[e].y += acc[r][e].y;
                    s2[e] = ffma2(acc[r][e], acc[r][e], s2[e]);
                }
                if (shift != nullptr) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[r][e] = ffma2(acc[r][e], sc[e], sh[e]);
                }
                if (relu) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[r][e] = make_float2(fmaxf(acc[r][e].x, 0.f), fmaxf(acc[r][e].y, 0.f));
                }
                store8p(yp + (int64_t)r * Wo * C, acc[r]);
            }
        }
    }
    if (stats != nullptr) {
        // two-stage reduction without shared atomics: the input tile is dead, so every thread
        // parks its 16 partial sums there ([2][col][CB]); then one thread per channel sums the
        // TW columns and issues the CTA's single global atomic for that channel
        __syncthreads();
        float* part = (float*)tile;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            part[(size_t)col * CB + cg * 8 + 2 * e] = s1[e].x;
            part[(size_t)col * CB + cg * 8 + 2 * e + 1] = s1[e].y;
            part[(size_t)(TW + col) * CB + cg * 8 + 2 * e] = s2[e].x;
            part[(size_t)(TW + col) * CB + cg * 8 + 2 * e + 1] = s2[e].y;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * CB; i += blockDim.x) {
            const int which = i / CB, ch = i - which * CB;
            float s = 0.f;
            for (int cidx = 0; cidx < TW; ++cidx) s += part[(size_t)(which * TW + cidx) * CB + ch];
            atomicAdd(stats + which * C + cb0 + ch, (double)s);
        }
    }
}


// ---------------------------------------------------------------------------------------------
// wgrad: dw[c][ky][kx] += sum_{n,ho,wo} x[n][ho*S-D+ky*D][wo*S-D+kx*D][c] * dy[n][ho][wo][c]
//
// Persistent CTAs (grid.y = channel block, grid.x strides over the spatial tiles) with a 2-stage
// TMA pipeline: while the threads consume stage s, one elected thread has already issued the box
// loads of the next tile -- the x halo tile {CB, IW, IH} and the dy tile {CB, TW, TH}; both are
// zero-filled outside the maps, which is the padding for x and "no contribution" for dy.
// thread = (8-channel group, tile column): its 72 partial sums (9 taps x 8 channels) live in
// registers across ALL tiles of the CTA; they are reduced over the TW columns through shared
// memory once, and one fp32 atomic per (channel, tap) per CTA goes to global memory.
template <typename T, int S, int D, int TH>
__global__ void __launch_bounds__(192, 2)
dw_wgrad_tma_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG,
                    float* __restrict__ dw, int CB, int TW, int tiles_w, int tiles_h, int ntiles, uint32_t stage_bytes) {
    constexpr int IH = Geo<S, D, TH>::IH;
    TSS_DYN_SMEM(uint8_t, smem_raw);
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    const int IW = (TW - 1) * S + 2 * D + 1;
    const uint32_t x_bytes = (uint32_t)IH * IW * CB * sizeof(T);
    const uint32_t x_pad = (x_bytes + 127) & ~127u;
    const uint32_t g_bytes = (uint32_t)TH * TW * CB * sizeof(T);
    uint64_t* bars = (uint64_t*)(smem + 2 * (size_t)stage_bytes);

    const int cb0 = blockIdx.y * CB;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(bars), 1);
        mbar_init(smem_u32(bars + 1), 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmG); }      // descriptor fetch (~0.5 us) under the predecessor's tail
    pdl_wait();

    auto issue = [&](int tile, int stage) {
        int t = tile;
        const int tw = t % tiles_w; t /= tiles_w;
        const int th = t % tiles_h;
        const int n = t / tiles_h;
        uint8_t* base = smem + (size_t)stage * stage_bytes;
        const uint32_t bar = smem_u32(bars + stage);
        mbar_expect_tx(bar, x_bytes + g_bytes);
        tma_load_4d(smem_u32(base), &tmX, bar, cb0, tw * TW * S - D, th * TH * S - D, n);
        tma_load_4d(smem_u32(base + x_pad), &tmG, bar, cb0, tw * TW, th * TH, n);
    };

    int tile = blockIdx.x;
    if (threadIdx.x == 0 && tile < ntiles) issue(tile, 0);

    const int CGB = CB >> 3;
    const int cg = threadIdx.x % CGB, col = threadIdx.x / CGB;     // col < TW by construction
    float2 acc[9][4];
#pragma unroll
    for (int k = 0; k < 9; ++k) zero8p(acc[k]);

    for (int it = 0; tile < ntiles; ++it, tile += gridDim.x) {
        const int stage = it & 1;
        const int next = tile + gridDim.x;
        if (threadIdx.x == 0 && next < ntiles) issue(next, stage ^ 1);    // that stage was released by the
                                                                          // __syncthreads of iteration it-1
        mbar_wait(smem_u32(bars + stage), (uint32_t)(it >> 1) & 1);
        const T* sx = (const T*)(smem + (size_t)stage * stage_bytes) + ((size_t)col * S) * CB + cg * 8;
        const T* sg = (const T*)(smem + (size_t)stage * stage_bytes + x_pad) + (size_t)col * CB + cg * 8;
        float2 g[TH][4];
#pragma unroll
        for (int r = 0; r < TH; ++r) load8p_smem(sg + (size_t)r * TW * CB, g[r]);
#pragma unroll
        for (int j = 0; j < IH; ++j) {
            bool used = false;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int tt = j - ky * D;
                if (tt >= 0 && tt % S == 0 && tt / S < TH) used = true;
            }
            if (!used) continue;
            float2 v[3][4];
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) load8p_smem(sx + ((size_t)j * IW + kx * D) * CB, v[kx]);
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int tt = j - ky * D;
                if (tt >= 0 && tt % S == 0 && tt / S < TH) {
                    const int r = tt / S;
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc[ky * 3 + kx][e] = ffma2(v[kx][e], g[r][e], acc[ky * 3 + kx][e]);
                }
            }
        }
        __syncthreads();                       // everyone is done with this stage: it may be refilled
    }

    // reduce the TW column partials: part[col][k][CB] in the (now idle) stage buffers
    float* part = (float*)smem;
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            part[((size_t)col * 9 + k) * CB + cg * 8 + 2 * e] = acc[k][e].x;
            part[((size_t)col * 9 + k) * CB + cg * 8 + 2 * e + 1] = acc[k][e].y;
        }
    __syncthreads();
    for (int i = threadIdx.x; i < 9 * CB; i += blockDim.x) {
        const int k = i / CB, ch = i - k * CB;
        float s = 0.f;
        for (int c = 0; c < TW; ++c) s += part[((size_t)c * 9 + k) * CB + ch];
        atomicAdd(dw + (size_t)(cb0 + ch) * 9 + k, s);
    }
}

template <typename T> struct TmaType;
template <> struct TmaType<float> { static constexpr CUtensorMapDataType v = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; };
template <> struct TmaType<bf16> { static constexpr CUtensorMapDataType v = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; };

template <typename T, int S, int D, int TH, bool FLIP>
int launch(const void* x, const float* w, void* y, int N, int Hi, int Wi, int Ho, int Wo, int C, int CB, int TW,
           const float* scale, const float* shift, int flags, double* stats, cudaStream_t st) {
    TssEncodeTiledFn enc = tss_encode_tiled();
    TSS_REQUIRE(enc != nullptr, "dwconv_tma: cuTensorMapEncodeTiled is not available from the driver");
    constexpr int IH = Geo<S, D, TH>::IH;
    const int IW = (TW - 1) * S + 2 * D + 1;
    CUtensorMap map;
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)Wi, (cuuint64_t)Hi, (cuuint64_t)N};
    cuuint64_t gstr[3] = {(cuuint64_t)C * sizeof(T), (cuuint64_t)Wi * C * sizeof(T), (cuuint64_t)Hi * Wi * C * sizeof(T)};
    cuuint32_t box[4] = {(cuuint32_t)CB, (cuuint32_t)IW, (cuuint32_t)IH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&map, TmaType<T>::v, 4, const_cast<void*>(x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TSS_REQUIRE(r == CUDA_SUCCESS, "dwconv_tma: cuTensorMapEncodeTiled failed (%d) C=%d W=%d H=%d N=%d box=%dx%dx%d",
                (int)r, C, Wi, Hi, N, CB, IW, IH);
    const int tiles_w = (Wo + TW - 1) / TW, tiles_h = (Ho + TH - 1) / TH;
    const int threads = (CB / 8) * TW;
    const size_t tile_bytes = (size_t)IH * IW * CB * sizeof(T);
    {
        // persistent CTAs with a 2-stage TMA ring (TSS_DW_PERSIST=0: the one-tile kernel below)
        bool launched = false;
        if (int e = dw_launch_persistent<T, S, D, TH, FLIP, false>(map, w, y, N, Hi, Wi, Ho, Wo, C, CB, TW, scale, shift, flags, stats,
                                                                  nullptr, nullptr, 0, 0, st, &launched))
            return e;
        if (launched) return TSS_OK;
    }
    const size_t smem = 128 + ((tile_bytes + 15) & ~(size_t)15) + 8 + 2 * CB * sizeof(float);
    auto kern = dw_tma_kernel<T, S, D, TH, FLIP>;
    static bool attr_set = false;          // per template instance; idempotent, benign race
    if (!attr_set) {
        TSS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    dim3 grid((unsigned)((int64_t)N * tiles_h * tiles_w), (unsigned)(C / CB));
    tss_launch(kern, grid, threads, smem, st, map, w, (T*)y, Ho, Wo, C, CB, TW, tiles_w, tiles_h, scale, shift, flags, stats);
    TSS_LAUNCH_CHECK("dwconv3x3(tma)");
    return TSS_OK;
}


template <typename T>
int make_map4(CUtensorMap* map, const void* base, int C, int W, int H, int N, int bc, int bw, int bh) {
    TssEncodeTiledFn enc = tss_encode_tiled();
    TSS_REQUIRE(enc != nullptr, "dwconv_tma: cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t gstr[3] = {(cuuint64_t)C * sizeof(T), (cuuint64_t)W * C * sizeof(T), (cuuint64_t)H * W * C * sizeof(T)};
    cuuint32_t box[4] = {(cuuint32_t)bc, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, TmaType<T>::v, 4, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TSS_REQUIRE(r == CUDA_SUCCESS, "dwconv_tma: cuTensorMapEncodeTiled failed (%d) C=%d W=%d H=%d N=%d box=%dx%dx%d",
                (int)r, C, W, H, N, bc, bw, bh);
    return TSS_OK;
}

template <typename T, int S, int D, int TH>
int launch_wgrad(const void* x, const void* dy, float* dw, int N, int Hi, int Wi, int Ho, int Wo, int C, int CB, int TW,
                 cudaStream_t st) {
    constexpr int IH = Geo<S, D, TH>::IH;
    int IW = (TW - 1) * S + 2 * D + 1;
    auto stage_size = [&](int tw) {
        const int iw = (tw - 1) * S + 2 * D + 1;
        const size_t xb = ((size_t)IH * iw * CB * sizeof(T) + 127) & ~(size_t)127;
        const size_t gb = ((size_t)TH * tw * CB * sizeof(T) + 127) & ~(size_t)127;
        return xb + gb;
    };
    while (2 * stage_size(TW) > 200 * 1024 && TW > 8) TW >>= 1;
    if (2 * stage_size(TW) > 200 * 1024 || IW > 256) return -1;
    IW = (TW - 1) * S + 2 * D + 1;
    size_t stage = stage_size(TW);
    const size_t part = (size_t)TW * 9 * CB * sizeof(float);
    if (2 * stage < part) stage = (part / 2 + 127) & ~(size_t)127;
    CUtensorMap mx, mg;
    if (int e = make_map4<T>(&mx, x, C, Wi, Hi, N, CB, IW, IH)) return e;
    if (int e = make_map4<T>(&mg, dy, C, Wo, Ho, N, CB, TW, TH)) return e;
    const int tiles_w = (Wo + TW - 1) / TW, tiles_h = (Ho + TH - 1) / TH;
    const int ntiles = N * tiles_h * tiles_w;
    const int cblocks = C / CB;
    int gx = (2 * tss_num_sms()) / cblocks;       // rounded down: a CTA beyond the resident set would run alone, after the others
    if (gx > ntiles) gx = ntiles;
    if (gx < 1) gx = 1;
    const int threads = (CB / 8) * TW;
    const size_t smem = 128 + 2 * stage + 16;
    auto kern = dw_wgrad_tma_kernel<T, S, D, TH>;
    static bool attr_set = false;
    if (!attr_set) {
        TSS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    tss_launch(kern, dim3((unsigned)gx, (unsigned)cblocks), threads, smem, st, mx, mg, dw, CB, TW, tiles_w, tiles_h, ntiles,
                                                                      (uint32_t)stage);
    TSS_LAUNCH_CHECK("dwconv3x3_wgrad(tma)");
    return TSS_OK;
}

}  // namespace

// Channel block / tile width for the TMA path, or false if C has no suitable block.
bool tss_dw_tma_config(int C, int* CB, int* TW) {
    // (A/B: TSS_DW_TW64=24 gives the 64-channel block 192 threads per CTA -- 12 warps per SM in the two-CTA persistent kernels)
    static const int tw64 = [] { const char* e = getenv("TSS_DW_TW64"); const int v = e ? atoi(e) : 16; return (v == 8 || v == 16 || v == 24) ? v : 16; }();
    if (C % 64 == 0) { *CB = 64; *TW = tw64; return true; }
    if (C % 96 == 0) { *CB = 96; *TW = 16; return true; }
    if (C % 48 == 0) { *CB = 48; *TW = 32; return true; }
    if (C % 32 == 0) { *CB = 32; *TW = 32; return true; }
    return false;
}

// flip = stride-1 dgrad (correlation with flipped taps).  Returns -1 if this shape is not covered.
int tss_dwconv3x3_tma(const void* x, const float* w, void* y, int N, int Hi, int Wi, int C, int stride, int dilation,
                      bool flip, const float* scale, const float* shift, int flags, double* stats, int dtype,
                      cudaStream_t st) {
    int CB, TW;
    if (!tss_dw_tma_config(C, &CB, &TW)) return -1;
    if (((uintptr_t)x & 15) != 0) return -1;
    const int Ho = (Hi - 1) / stride + 1, Wo = (Wi - 1) / stride + 1;
    TSS_DISPATCH_DTYPE(dtype, "dwconv3x3(tma)", {
        if (sizeof(T) == 4 && TW == 32) TW = 16;      // keep fp32 tiles under the smem budget
        if (stride == 1 && dilation == 1) {
            if (flip) return launch<T, 1, 1, 8, true>(x, w, y, N, Hi, Wi, Ho, Wo, C, CB, TW, scale, shift, flags, stats, st);
            return launch<T, 1, 1, 8, false>(x, w, y, N, Hi, Wi, Ho, Wo, C, CB, TW, scale, shift, flags, stats, st);
        }
        if (stride == 2 && dilation == 1 && !flip)
            return launch<T, 2, 1, 4, false>(x, w, y, N, Hi, Wi, Ho, Wo, C, CB, TW, scale, shift, flags, stats, st);
        if (stride == 1 && dilation == 4) {
            if (flip) return launch<T, 1, 4, 8, true>(x, w, y, N, Hi, Wi, Ho, Wo, C, CB, TW, scale, shift, flags, stats, st);
            return launch<T, 1, 4, 8, false>(x, w, y, N, Hi, Wi, Ho, Wo, C, CB, TW, scale, shift, flags, stats, st);
        }
        return -1;
    });
}

// Returns -1 if this shape is not covered by the TMA wgrad kernel.
int tss_dwconv3x3_wgrad_tma(const void* x, const void* dy, float* dw, int N, int Hi, int Wi, int C, int stride,
                            int dilation, int dtype, cudaStream_t st) {
    int CB, TW;
    if (!tss_dw_tma_config(C, &CB, &TW)) return -1;
    if (((uintptr_t)x & 15) != 0 || ((uintptr_t)dy & 15) != 0) return -1;
    const int Ho = (Hi - 1) / stride + 1, Wo = (Wi - 1) / stride + 1;
    TSS_DISPATCH_DTYPE(dtype, "dwconv3x3_wgrad(tma)", {
        if (sizeof(T) == 4 && TW == 32) TW = 16;
        if (stride == 1 && dilation == 1) return launch_wgrad<T, 1, 1, 8>(x, dy, dw, N, Hi, Wi, Ho, Wo, C, CB, TW, st);
        if (stride == 2 && dilation == 1) return launch_wgrad<T, 2, 1, 4>(x, dy, dw, N, Hi, Wi, Ho, Wo, C, CB, TW, st);
        if (stride == 1 && dilation == 4) return launch_wgrad<T, 1, 4, 8>(x, dy, dw, N, Hi, Wi, Ho, Wo, C, CB, TW, st);
        return -1;
    });
}
