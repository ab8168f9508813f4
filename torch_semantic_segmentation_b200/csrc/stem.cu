// Stem: dense 3x3 stride-2 pad-1 convolution 3 -> 32 channels.
// Input is the caller's NCHW fp32 image batch, read in place (no layout or precision
// pre-pass over the largest input tensor); output is NHWC in the activation dtype.
// K = 27 is too small/awkward for a tensor-core tile, and the layer is bound by its
// 2 x 113 MB (TRN) of output traffic plus 864 FMA per pixel, so this is a SIMT kernel:
//   fwd  : one thread per pair of adjacent output pixels, 2 x 16 float2 accumulators (FFMA2 over
//          channel pairs), weights broadcast from smem as 128-bit loads shared by both pixels.
//   wgrad: persistent CTAs; per 8x32 pixel tile the input patch and dy tile are staged in
//          shared memory, thread = (4 output channels, 1 input channel, pixel split) keeps
//          36 accumulators; one cross-split reduction and 864 atomics per CTA at the end.
#include "common.cuh"

namespace {

constexpr int CO = 32;            // output channels (both nets)
constexpr int kFwdThreads = 128;

template <typename T>
__global__ void __launch_bounds__(kFwdThreads)
stem_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, T* __restrict__ y,
                int N, int H, int W, int Ho, int Wo, const float* __restrict__ scale,
                const float* __restrict__ shift, int relu, double* __restrict__ stats) {
    pdl_wait();
    __shared__ __align__(16) float ws[27][CO];
    __shared__ float s_stat[2 * CO];
    for (int i = threadIdx.x; i < 27 * CO; i += kFwdThreads) {
        const int co = i / 27, t = i - co * 27;        // w[co][ci][ky][kx], t = ci*9+ky*3+kx
        ws[t][co] = w[i];
    }
    if (threadIdx.x < 2 * CO) s_stat[threadIdx.x] = 0.f;
    __syncthreads();

    // a thread owns TWO horizontally adjacent output pixels: every 128-bit weight load from shared
    // memory (the limiter of this layer: 216 of them per pixel) feeds 8 FMAs instead of 4, and the
    // two 3x3 windows share one of their three input columns
    const int Wp = (Wo + 1) >> 1;
    const int64_t total = (int64_t)N * Ho * Wp;
    const int64_t p = (int64_t)blockIdx.x * kFwdThreads + threadIdx.x;
    const bool valid = p < total;
    float2 acc[2][CO / 2];                     // [pixel][channel pair]: FFMA2
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int c = 0; c < CO / 2; ++c) acc[q][c] = make_float2(0.f, 0.f);
    int wo = 0, ho = 0, n = 0;
    bool second = false;
    if (valid) {
        int64_t t = p;
        wo = (int)(t % Wp) * 2; t /= Wp;
        ho = (int)(t % Ho);
        n = (int)(t / Ho);
        second = wo + 1 < Wo;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int hi = 2 * ho - 1 + ky;
                const bool row_ok = hi >= 0 && hi < H;
                const float* xr = x + (((int64_t)n * 3 + ci) * H + hi) * W;
                float in[5];
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    const int wi = 2 * wo - 1 + j;
                    in[j] = (row_ok && wi >= 0 && wi < W) ? __ldg(xr + wi) : 0.f;
                }
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int tp = ci * 9 + ky * 3 + kx;
                    const float2 i0 = make_float2(in[kx], in[kx]), i1 = make_float2(in[kx + 2], in[kx + 2]);
#pragma unroll
                    for (int c4 = 0; c4 < CO / 4; ++c4) {
                        const float4 wv = *reinterpret_cast<const float4*>(&ws[tp][c4 * 4]);
                        const float2 wa = make_float2(wv.x, wv.y), wb = make_float2(wv.z, wv.w);
                        acc[0][c4 * 2 + 0] = ffma2(i0, wa, acc[0][c4 * 2 + 0]);
                        acc[0][c4 * 2 + 1] = ffma2(i0, wb, acc[0][c4 * 2 + 1]);
                        acc[1][c4 * 2 + 0] = ffma2(i1, wa, acc[1][c4 * 2 + 0]);
                        acc[1][c4 * 2 + 1] = ffma2(i1, wb, acc[1][c4 * 2 + 1]);
                    }
                }
            }
        if (!second) {
#pragma unroll
            for (int c = 0; c < CO / 2; ++c) acc[1][c] = make_float2(0.f, 0.f);
        }
        T* yp = y + (((int64_t)n * Ho + ho) * Wo + wo) * CO;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            if (q == 1 && !second) break;
#pragma unroll
            for (int c8 = 0; c8 < CO / 8; ++c8) {
                float o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float2 a2 = acc[q][c8 * 4 + (e >> 1)];
                    float v = (e & 1) ? a2.y : a2.x;
                    if (shift != nullptr) v = fmaf(v, scale != nullptr ? __ldg(scale + c8 * 8 + e) : 1.f, __ldg(shift + c8 * 8 + e));
                    if (relu) v = fmaxf(v, 0.f);
                    o[e] = v;
                }
                store8(yp + q * CO + c8 * 8, o);
            }
        }
    }

    if (stats != nullptr) {     // raw (pre-affine) output statistics; invalid threads / pixels hold zeros
        const int lane = threadIdx.x & 31;
        float sm[CO], sq[CO];
#pragma unroll
        for (int c = 0; c < CO / 2; ++c) {
            sm[2 * c] = acc[0][c].x + acc[1][c].x;
            sm[2 * c + 1] = acc[0][c].y + acc[1][c].y;
            sq[2 * c] = fmaf(acc[0][c].x, acc[0][c].x, acc[1][c].x * acc[1][c].x);
            sq[2 * c + 1] = fmaf(acc[0][c].y, acc[0][c].y, acc[1][c].y * acc[1][c].y);
        }
        const float s1 = warp_transpose_sum32(sm, lane);
        const float s2 = warp_transpose_sum32(sq, lane);
        atomicAdd(&s_stat[lane], s1);
        atomicAdd(&s_stat[CO + lane], s2);
        __syncthreads();
        if (threadIdx.x < 2 * CO) atomicAdd(stats + threadIdx.x, (double)s_stat[threadIdx.x]);
    }
}

// ---------------------------------------------------------------- wgrad ----------------
constexpr int TH = 8, TW = 32;                 // output-pixel tile
constexpr int PH = 2 * TH + 1, PW = 2 * TW + 1; // input patch 17 x 65
constexpr int PWP = PW + 2;                    // pitch
constexpr int kSplits = 10;
constexpr int kWgThreads = 24 * kSplits;       // (8 co-quads x 3 ci) x 10 pixel splits
constexpr int kPatch = (3 * PH * PWP + 3) & ~3; // floats, rounded so that dy_s stays 16-byte aligned
constexpr int kDyTile = TH * TW * CO;          // floats
static_assert(kPatch + kDyTile >= kSplits * 864, "scratch reuse");

template <typename T>
__global__ void __launch_bounds__(kWgThreads)
stem_wgrad_kernel(const float* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw,
                  int N, int H, int W, int Ho, int Wo, int tiles_h, int tiles_w) {
    pdl_wait();
    TSS_DYN_SMEM(float, smem);
    float* in_s = smem;                 // [3][PH][PWP]
    float* dy_s = smem + kPatch;        // [TH*TW][CO]
    const int tid = threadIdx.x;
    const int quad = tid % 8, ci = (tid / 8) % 3, split = tid / 24;

    float acc[4][9];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[i][t] = 0.f;

    const int64_t ntiles = (int64_t)N * tiles_h * tiles_w;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int64_t tt = tile;
        const int tw = (int)(tt % tiles_w); tt /= tiles_w;
        const int th = (int)(tt % tiles_h);
        const int n = (int)(tt / tiles_h);
        const int ho0 = th * TH, wo0 = tw * TW;
        const int hi0 = 2 * ho0 - 1, wi0 = 2 * wo0 - 1;
        for (int i = tid; i < 3 * PH * PW; i += kWgThreads) {
            const int c = i % PW;
            const int r = (i / PW) % PH;
            const int cc = i / (PW * PH);
            const int hi = hi0 + r, wi = wi0 + c;
            const bool ok = hi >= 0 && hi < H && wi >= 0 && wi < W;
            in_s[(cc * PH + r) * PWP + c] = ok ? __ldg(x + (((int64_t)n * 3 + cc) * H + hi) * W + wi) : 0.f;
        }
        for (int i = tid; i < TH * TW * (CO / 8); i += kWgThreads) {
            const int v8 = i % (CO / 8);
            const int px = (i / (CO / 8)) % TW;
            const int py = i / ((CO / 8) * TW);
            const int ho = ho0 + py, wo = wo0 + px;
            float v[8];
            if (ho < Ho && wo < Wo) load8(dy + (((int64_t)n * Ho + ho) * Wo + wo) * CO + v8 * 8, v);
            else zero8(v);
            float* d = dy_s + (py * TW + px) * CO + v8 * 8;
            *reinterpret_cast<float4*>(d) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(d + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
        __syncthreads();
        for (int p = split; p < TH * TW; p += kSplits) {
            const int py = p / TW, px = p - py * TW;
            const float4 g4 = *reinterpret_cast<const float4*>(dy_s + p * CO + quad * 4);
            const float g[4] = {g4.x, g4.y, g4.z, g4.w};
            const float* ip = in_s + (ci * PH + 2 * py) * PWP + 2 * px;
            float iv[9];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) iv[ky * 3 + kx] = ip[ky * PWP + kx];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int t = 0; t < 9; ++t) acc[i][t] = fmaf(g[i], iv[t], acc[i][t]);
        }
        __syncthreads();
    }

    // cross-split reduction through shared memory, then one atomic per weight per CTA
    float* scratch = smem;              // [kSplits][864], index co*27 + ci*9 + t
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int t = 0; t < 9; ++t) scratch[split * 864 + (quad * 4 + i) * 27 + ci * 9 + t] = acc[i][t];
    __syncthreads();
    for (int o = tid; o < 864; o += kWgThreads) {
        float s = 0.f;
#pragma unroll
        for (int sp = 0; sp < kSplits; ++sp) s += scratch[sp * 864 + o];
        atomicAdd(dw + o, s);
    }
}

}  // namespace

extern "C" int tss_stem3x3s2_fwd(const float* x, const float* w, void* y, int N, int H, int W, int Cout,
                                 const float* scale, const float* shift, int flags, double* stats,
                                 int dtype, void* stream) {
    TSS_REQUIRE(N > 0 && H > 0 && W > 0, "stem3x3s2_fwd: empty input");
    TSS_REQUIRE(Cout == CO, "stem3x3s2_fwd: Cout=%d unsupported (only %d)", Cout, CO);
    TSS_REQUIRE(scale == nullptr || shift != nullptr, "stem3x3s2_fwd: scale without shift");
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const int64_t total = (int64_t)N * Ho * ((Wo + 1) / 2);        // a thread owns two adjacent output pixels
    TSS_DISPATCH_DTYPE(dtype, "stem3x3s2_fwd", {
        tss_launch(stem_fwd_kernel<T>, (unsigned)ceil_div64(total, kFwdThreads), kFwdThreads, 0, (cudaStream_t)stream, 
            x, w, (T*)y, N, H, W, Ho, Wo, scale, shift, flags & TSS_EPI_RELU, stats);
        TSS_LAUNCH_CHECK("stem3x3s2_fwd");
        return TSS_OK;
    });
}

extern "C" int tss_stem3x3s2_wgrad(const float* x, const void* dy, float* dw, int N, int H, int W, int Cout,
                                   int dtype, void* stream) {
    TSS_REQUIRE(N > 0 && H > 0 && W > 0, "stem3x3s2_wgrad: empty input");
    TSS_REQUIRE(Cout == CO, "stem3x3s2_wgrad: Cout=%d unsupported (only %d)", Cout, CO);
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const int tiles_h = (Ho + TH - 1) / TH, tiles_w = (Wo + TW - 1) / TW;
    const int64_t ntiles = (int64_t)N * tiles_h * tiles_w;
    int64_t cap = (int64_t)tss_num_sms() * 3;
    const int grid = (int)(ntiles < cap ? ntiles : cap);
    const size_t smem = (size_t)(kPatch + kDyTile) * sizeof(float);
    TSS_DISPATCH_DTYPE(dtype, "stem3x3s2_wgrad", {
        static bool attr_set = false;   // benign race: idempotent attribute
        if (!attr_set) {
            TSS_CUDA(cudaFuncSetAttribute(stem_wgrad_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set = true;
        }
        tss_launch(stem_wgrad_kernel<T>, grid, kWgThreads, smem, (cudaStream_t)stream, 
            x, (const T*)dy, dw, N, H, W, Ho, Wo, tiles_h, tiles_w);
        TSS_LAUNCH_CHECK("stem3x3s2_wgrad");
        return TSS_OK;
    });
}
