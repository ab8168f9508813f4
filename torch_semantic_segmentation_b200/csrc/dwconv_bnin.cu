// Depthwise 3x3 forward and weight gradient (training) that read the RAW conv output of the producing layer and
// apply its BatchNorm + ReLU on the fly -- the activated tensor between conv1 and conv2 of a bottleneck
// (fastscnn.py:149-156; 6x the block width) is never materialised: conv1's bn_apply pass (read + write of that
// tensor) disappears from the forward pass, and the backward pass is unchanged in traffic (the weight gradient
// re-derives the activation from the raw tensor it reads anyway; the dgrad's fused reduction already masks with it).
// The kernels are the TMA-staged ones of dwconv_tma.cu; the only change is in_affine() behind every shared-memory
// read of the input tile.  TMA zero-fill cannot provide the padding any more (BN(0) != 0): positions outside the
// image are forced to zero from their coordinates.
#include "tma.cuh"
#include "dw_fwd_persistent.cuh"

namespace {

template <int S, int D, int TH>
struct Geo {
    static constexpr int IH = (TH - 1) * S + 2 * D + 1;
};

// z = relu?(y * scale + shift) inside the image, 0 outside (the convolution's zero padding applies to z)
__device__ __forceinline__ void in_affine(float2 (&v)[4], const float2 (&sc)[4], const float2 (&sh)[4], int relu, bool inside) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        float2 t = ffma2(v[e], sc[e], sh[e]);
        if (relu) t = make_float2(fmaxf(t.x, 0.f), fmaxf(t.y, 0.f));
        v[e] = inside ? t : make_float2(0.f, 0.f);
    }
}

template <typename T, int S, int D, int TH>
__global__ void __launch_bounds__(192, 2)
dw_tma_bnin_kernel(const __grid_constant__ CUtensorMap tmX, const float* __restrict__ w, T* __restrict__ y,
                   int Ho, int Wo, int C, int CB, int TW, int tiles_w, int tiles_h,
                   const float* __restrict__ scale, const float* __restrict__ shift, int flags,
                   double* __restrict__ stats, const float* __restrict__ in_scale, const float* __restrict__ in_shift,
                   int in_relu, int Hi, int Wi) {
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
            wr[k][e] = make_float2(__ldg(w + (c0 + 2 * e) * 9 + k), __ldg(w + (c0 + 2 * e + 1) * 9 + k));

    float2 isc[4], ish[4];                                   // the producer's BatchNorm, applied while reading the tile
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        isc[e] = make_float2(__ldg(in_scale + c0 + 2 * e), __ldg(in_scale + c0 + 2 * e + 1));
        ish[e] = make_float2(__ldg(in_shift + c0 + 2 * e), __ldg(in_shift + c0 + 2 * e + 1));
    }

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
        for (int kx = 0; kx < 3; ++kx) {
            load8p_smem(tp + ((size_t)j * IW + kx * D) * CB, v[kx]);
            const int hq = ho0 * S - D + j, wq = wo0 * S - D + col * S + kx * D;
            in_affine(v[kx], isc, ish, in_relu, hq >= 0 && hq < Hi && wq >= 0 && wq < Wi);
        }
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



template <typename T, int S, int D, int TH>
__global__ void __launch_bounds__(192, 2)
dw_wgrad_tma_bnin_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG,
                         float* __restrict__ dw, int CB, int TW, int tiles_w, int tiles_h, int ntiles, uint32_t stage_bytes,
                         const float* __restrict__ in_scale, const float* __restrict__ in_shift, int in_relu, int Hi, int Wi,
                         int pad_nan) {
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
    float2 isc[4], ish[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        isc[e] = make_float2(__ldg(in_scale + cb0 + cg * 8 + 2 * e), __ldg(in_scale + cb0 + cg * 8 + 2 * e + 1));
        ish[e] = make_float2(__ldg(in_shift + cb0 + cg * 8 + 2 * e), __ldg(in_shift + cb0 + cg * 8 + 2 * e + 1));
    }

    for (int it = 0; tile < ntiles; ++it, tile += gridDim.x) {
        const int stage = it & 1;
        const int next = tile + gridDim.x;
        if (threadIdx.x == 0 && next < ntiles) issue(next, stage ^ 1);    // that stage was released by the
                                                                          // __syncthreads of iteration it-1
        mbar_wait(smem_u32(bars + stage), (uint32_t)(it >> 1) & 1);
        const int tw_ = tile % tiles_w, th_ = (tile / tiles_w) % tiles_h;      // where this tile sits in the image
        const T* sx = (const T*)(smem + (size_t)stage * stage_bytes) + ((size_t)col * S) * CB + cg * 8;
        const T* sg = (const T*)(smem + (size_t)stage * stage_bytes + x_pad) + (size_t)col * CB + cg * 8;
        float2 g[TH][4];
#pragma unroll
        for (int r = 0; r < TH; ++r) load8p_smem(sg + (size_t)r * TW * CB, g[r]);
        // kPadNan: out-of-image elements arrive as NaN (tensor-map fill) and the ReLU behind the BatchNorm makes them the
        // zero padding -- no coordinate test and no select behind the loads (see dw_fwd_persistent.cuh)
        auto correlate = [&](auto pad_nan_tag) {
            constexpr bool kPadNan = decltype(pad_nan_tag)::value;
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
                for (int kx = 0; kx < 3; ++kx) {
                    load8p_smem(sx + ((size_t)j * IW + kx * D) * CB, v[kx]);
                    if (kPadNan) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 a = ffma2(v[kx][e], isc[e], ish[e]);
                            v[kx][e] = make_float2(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f));
                        }
                    } else {
                        const int hq = th_ * TH * S - D + j, wq = tw_ * TW * S - D + col * S + kx * D;
                        in_affine(v[kx], isc, ish, in_relu, hq >= 0 && hq < Hi && wq >= 0 && wq < Wi);
                    }
                }
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
        };
        if (pad_nan) correlate(DwTrue{}); else correlate(DwFalse{});
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

template <typename T> struct TmaTypeI;
template <> struct TmaTypeI<float> { static constexpr CUtensorMapDataType v = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; };
template <> struct TmaTypeI<bf16> { static constexpr CUtensorMapDataType v = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; };

template <typename T>
int make_map4i(CUtensorMap* map, const void* base, int C, int W, int H, int N, int bc, int bw, int bh, const char* name,
               bool nan_fill = false) {
    TssEncodeTiledFn enc = tss_encode_tiled();
    TSS_REQUIRE(enc != nullptr, "%s: cuTensorMapEncodeTiled is not available from the driver", name);
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t gstr[3] = {(cuuint64_t)C * sizeof(T), (cuuint64_t)W * C * sizeof(T), (cuuint64_t)H * W * C * sizeof(T)};
    cuuint32_t box[4] = {(cuuint32_t)bc, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, TmaTypeI<T>::v, 4, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     nan_fill ? CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA : CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TSS_REQUIRE(r == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled failed (%d)", name, (int)r);
    return TSS_OK;
}

template <typename T, int S, int TH>
int launch_fwd(const void* x, const float* w, void* y, int N, int Hi, int Wi, int Ho, int Wo, int C, int CB, int TW, double* stats,
               const float* in_scale, const float* in_shift, int in_relu, cudaStream_t st) {
    constexpr int IH = Geo<S, 1, TH>::IH;
    const int IW = (TW - 1) * S + 3;
    CUtensorMap map;
    if (in_relu) {
        // out-of-image elements arrive as NaN; the ReLU behind the BatchNorm turns them into the convolution's zero padding
        static const int nan_pad = [] { const char* e = getenv("TSS_DW_NAN_PAD"); return (e != nullptr && e[0] == '0') ? 0 : 1; }();
        CUtensorMap map_nan;
        if (int e = make_map4i<T>(&map_nan, x, C, Wi, Hi, N, CB, IW, IH, "dwconv3x3_fwd_bnin", nan_pad != 0)) return e;
        bool launched = false;
        if (int e = dw_launch_persistent<T, S, 1, TH, false, true>(map_nan, w, y, N, Hi, Wi, Ho, Wo, C, CB, TW, nullptr, nullptr, 0, stats,
                                                                  in_scale, in_shift, in_relu, nan_pad, st, &launched))
            return e;
        if (launched) return TSS_OK;
    }
    if (int e = make_map4i<T>(&map, x, C, Wi, Hi, N, CB, IW, IH, "dwconv3x3_fwd_bnin")) return e;
    if (!in_relu) {
        bool launched = false;
        if (int e = dw_launch_persistent<T, S, 1, TH, false, true>(map, w, y, N, Hi, Wi, Ho, Wo, C, CB, TW, nullptr, nullptr, 0, stats,
                                                                  in_scale, in_shift, in_relu, 0, st, &launched))
            return e;
        if (launched) return TSS_OK;
    }
    const int tiles_w = (Wo + TW - 1) / TW, tiles_h = (Ho + TH - 1) / TH;
    const int threads = (CB / 8) * TW;
    const size_t tile_bytes = (size_t)IH * IW * CB * sizeof(T);
    const size_t smem = 128 + ((tile_bytes + 15) & ~(size_t)15) + 8 + 2 * CB * sizeof(float);
    auto kern = dw_tma_bnin_kernel<T, S, 1, TH>;
    static bool attr_set = false;
    if (!attr_set) {
        TSS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    dim3 grid((unsigned)((int64_t)N * tiles_h * tiles_w), (unsigned)(C / CB));
    tss_launch(kern, grid, threads, smem, st, map, w, (T*)y, Ho, Wo, C, CB, TW, tiles_w, tiles_h, (const float*)nullptr,
               (const float*)nullptr, 0, stats, in_scale, in_shift, in_relu, Hi, Wi);
    TSS_LAUNCH_CHECK("dwconv3x3_fwd_bnin");
    return TSS_OK;
}

template <typename T, int S, int TH>
int launch_wgrad(const void* x, const void* dy, float* dw, int N, int Hi, int Wi, int Ho, int Wo, int C, int CB, int TW,
                 const float* in_scale, const float* in_shift, int in_relu, cudaStream_t st) {
    constexpr int IH = Geo<S, 1, TH>::IH;
    auto stage_size = [&](int tw) {
        const int iw = (tw - 1) * S + 3;
        const size_t xb = ((size_t)IH * iw * CB * sizeof(T) + 127) & ~(size_t)127;
        const size_t gb = ((size_t)TH * tw * CB * sizeof(T) + 127) & ~(size_t)127;
        return xb + gb;
    };
    while (2 * stage_size(TW) > 200 * 1024 && TW > 8) TW >>= 1;
    TSS_REQUIRE(2 * stage_size(TW) <= 200 * 1024, "dwconv3x3_wgrad_bnin: tile does not fit in shared memory");
    const int IW = (TW - 1) * S + 3;
    size_t stage = stage_size(TW);
    const size_t part = (size_t)TW * 9 * CB * sizeof(float);
    if (2 * stage < part) stage = (part / 2 + 127) & ~(size_t)127;
    CUtensorMap mx, mg;
    static const int nan_env = [] { const char* e = getenv("TSS_DW_NAN_PAD"); return (e != nullptr && e[0] == '0') ? 0 : 1; }();
    const int pad_nan = (in_relu && nan_env) ? 1 : 0;
    if (int e = make_map4i<T>(&mx, x, C, Wi, Hi, N, CB, IW, IH, "dwconv3x3_wgrad_bnin", pad_nan != 0)) return e;
    if (int e = make_map4i<T>(&mg, dy, C, Wo, Ho, N, CB, TW, TH, "dwconv3x3_wgrad_bnin")) return e;
    const int tiles_w = (Wo + TW - 1) / TW, tiles_h = (Ho + TH - 1) / TH;
    const int ntiles = N * tiles_h * tiles_w;
    const int cblocks = C / CB;
    int gx = (2 * tss_num_sms()) / cblocks;       // rounded down: a CTA beyond the resident set would run alone, after the others
    if (gx > ntiles) gx = ntiles;
    if (gx < 1) gx = 1;
    const int threads = (CB / 8) * TW;
    const size_t smem = 128 + 2 * stage + 16;
    auto kern = dw_wgrad_tma_bnin_kernel<T, S, 1, TH>;
    static bool attr_set = false;
    if (!attr_set) {
        TSS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    tss_launch(kern, dim3((unsigned)gx, (unsigned)cblocks), threads, smem, st, mx, mg, dw, CB, TW, tiles_w, tiles_h, ntiles,
               (uint32_t)stage, in_scale, in_shift, in_relu, Hi, Wi, pad_nan);
    TSS_LAUNCH_CHECK("dwconv3x3_wgrad_bnin");
    return TSS_OK;
}

}  // namespace

bool tss_dw_tma_config(int C, int* CB, int* TW);     // dwconv_tma.cu

extern "C" int tss_dwconv3x3_fwd_bnin(const void* x, const float* in_scale, const float* in_shift, int in_flags,
                                      const float* w, void* y, int N, int Hi, int Wi, int C, int stride, double* stats,
                                      int dtype, void* stream) {
    TSS_REQUIRE(N > 0 && Hi > 0 && Wi > 0 && C > 0 && (stride == 1 || stride == 2), "dwconv3x3_fwd_bnin: bad shape");
    TSS_REQUIRE(in_scale != nullptr && in_shift != nullptr, "dwconv3x3_fwd_bnin: missing input affine");
    TSS_REQUIRE((((uintptr_t)x | (uintptr_t)y) & 15) == 0, "dwconv3x3_fwd_bnin: tensors must be 16-byte aligned");
    int CB, TW;
    TSS_REQUIRE(tss_dw_tma_config(C, &CB, &TW), "dwconv3x3_fwd_bnin: no channel block for C=%d", C);
    const int Ho = (Hi - 1) / stride + 1, Wo = (Wi - 1) / stride + 1;
    TSS_DISPATCH_DTYPE(dtype, "dwconv3x3_fwd_bnin", {
        if (sizeof(T) == 4 && TW == 32) TW = 16;
        if (stride == 1)
            return launch_fwd<T, 1, 8>(x, w, y, N, Hi, Wi, Ho, Wo, C, CB, TW, stats, in_scale, in_shift, in_flags & TSS_EPI_RELU,
                                       (cudaStream_t)stream);
        return launch_fwd<T, 2, 4>(x, w, y, N, Hi, Wi, Ho, Wo, C, CB, TW, stats, in_scale, in_shift, in_flags & TSS_EPI_RELU,
                                   (cudaStream_t)stream);
    });
}

extern "C" int tss_dwconv3x3_wgrad_bnin(const void* x, const float* in_scale, const float* in_shift, int in_flags,
                                        const void* dy, float* dw, int N, int Hi, int Wi, int C, int stride, int dtype,
                                        void* stream) {
    TSS_REQUIRE(N > 0 && Hi > 0 && Wi > 0 && C > 0 && (stride == 1 || stride == 2), "dwconv3x3_wgrad_bnin: bad shape");
    TSS_REQUIRE(in_scale != nullptr && in_shift != nullptr, "dwconv3x3_wgrad_bnin: missing input affine");
    TSS_REQUIRE((((uintptr_t)x | (uintptr_t)dy) & 15) == 0, "dwconv3x3_wgrad_bnin: tensors must be 16-byte aligned");
    int CB, TW;
    TSS_REQUIRE(tss_dw_tma_config(C, &CB, &TW), "dwconv3x3_wgrad_bnin: no channel block for C=%d", C);
    const int Ho = (Hi - 1) / stride + 1, Wo = (Wi - 1) / stride + 1;
    TSS_DISPATCH_DTYPE(dtype, "dwconv3x3_wgrad_bnin", {
        if (sizeof(T) == 4 && TW == 32) TW = 16;
        if (stride == 1)
            return launch_wgrad<T, 1, 8>(x, dy, dw, N, Hi, Wi, Ho, Wo, C, CB, TW, in_scale, in_shift, in_flags & TSS_EPI_RELU,
                                         (cudaStream_t)stream);
        return launch_wgrad<T, 2, 4>(x, dy, dw, N, Hi, Wi, Ho, Wo, C, CB, TW, in_scale, in_shift, in_flags & TSS_EPI_RELU,
                                     (cudaStream_t)stream);
    });
}
