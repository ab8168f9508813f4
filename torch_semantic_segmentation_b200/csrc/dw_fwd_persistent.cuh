// The persistent TMA-staged depthwise forward kernel (also the stride-1 dgrad: flipped taps), shared by dwconv_tma.cu
// (plain input) and dwconv_bnin.cu (input = raw output of the producing conv, its BatchNorm applied while reading).
#pragma once
#include <stdlib.h>

#include "tma.cuh"

template <int S, int D, int TH>
struct DwGeo {
    static constexpr int IH = (TH - 1) * S + 2 * D + 1;
};
struct DwTrue { static constexpr bool value = true; };
struct DwFalse { static constexpr bool value = false; };

// ---------------------------------------------------------------------------------------------
// Persistent variant of dw_tma_kernel: a CTA keeps its 72 weights in registers and walks the spatial tiles
// blockIdx.x, blockIdx.x + gridDim.x, ... of its channel block with a 2-stage TMA ring (the halo tile of tile i+1 is in
// flight while tile i is computed, as in the weight-gradient kernel below); the BatchNorm statistics live in per-thread
// shared-memory slots across ALL tiles and are reduced once per CTA.  The one-tile kernel pays barrier init, weight
// loads, the TMA round trip, a shared-memory reduction and 2*CB fp64 atomics PER TILE.
//   BNIN: the input is the RAW output of the producing conv; its BatchNorm (+ReLU) is applied behind every shared-memory
//   read (dwconv_bnin.cu explains why the zero padding then has to come from coordinates instead of the TMA fill).
// The launcher asks for enough shared memory that exactly `ctas_per_sm` CTAs fit on an SM: with programmatic dependent
// launch the CTAs of a grid this small become resident while the previous kernel drains, wherever a slot is free first --
// without the cap 297 CTAs piled up three deep on 101 of the 148 SMs (tools/trace_kernels.py).
// MAXT = 128 or 192 threads: two CTAs of 128 threads may use up to 255 registers each, two of 192 threads 168.
template <typename T, int S, int D, int TH, bool FLIP, bool BNIN, int MAXT>
__global__ void __launch_bounds__(MAXT, 2)
dw_tma_persistent_kernel(const __grid_constant__ CUtensorMap tmX, const float* __restrict__ w, T* __restrict__ y,
                         int Ho, int Wo, int C, int CB, int TW, int tiles_w, int tiles_h, int ntiles, uint32_t stage_bytes,
                         const float* __restrict__ scale, const float* __restrict__ shift, int flags,
                         double* __restrict__ stats, const float* __restrict__ in_scale, const float* __restrict__ in_shift,
                         int in_relu, int Hi, int Wi, int pad_nan) {
    constexpr int IH = DwGeo<S, D, TH>::IH;
    TSS_DYN_SMEM(uint8_t, smem_raw);
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    const int IW = (TW - 1) * S + 2 * D + 1;
    const uint32_t tile_bytes = (uint32_t)IH * IW * CB * sizeof(T);
    uint64_t* bars = (uint64_t*)(smem + 2 * (size_t)stage_bytes);
    float* part = (float*)(bars + 2);          // [2][TW][CB]: every thread's running statistics (its private 16 slots)
    float* s_w = part + (size_t)2 * TW * CB;   // [9][CB] taps
    const int cb0 = blockIdx.y * CB;
    TSS_MARK(0);
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(bars), 1);
        mbar_init(smem_u32(bars + 1), 1);
        mbar_fence_init();
    }
    dw_stage_taps<FLIP>(w, cb0, CB, s_w);
    const int CGB = CB >> 3;
    const int cg = threadIdx.x % CGB, col = threadIdx.x / CGB;     // col < TW by construction
    const int c0 = cb0 + cg * 8;
    float* my1 = part + (size_t)col * CB + cg * 8;           // the statistics live in shared memory between tiles: registers
    float* my2 = part + (size_t)(TW + col) * CB + cg * 8;    // hold 72 weights + 64 accumulators already
    if (stats != nullptr) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { my1[e] = 0.f; my2[e] = 0.f; }
    }
    __syncthreads();
    float2 wr[9][4];
    dw_take_taps(s_w, CB, cg, wr);
    TSS_MARK(1);
    if (threadIdx.x == 0) {tma_prefetch_desc(&tmX); }      // descriptor fetch (~0.5 us) under the predecessor's tail
    pdl_wait();
    TSS_MARK(2);

    auto issue = [&](int tile, int stage) {
        int t = tile;
        const int tw = t % tiles_w; t /= tiles_w;
        const int th = t % tiles_h;
        const int n = t / tiles_h;
        const uint32_t bar = smem_u32(bars + stage);
        mbar_expect_tx(bar, tile_bytes);
        tma_load_4d(smem_u32(smem + (size_t)stage * stage_bytes), &tmX, bar, cb0, tw * TW * S - D, th * TH * S - D, n);
    };
    int tile = blockIdx.x;
    if (threadIdx.x == 0 && tile < ntiles) issue(tile, 0);

    float2 isc[4], ish[4];                                   // BNIN: the producer's BatchNorm, written by the kernel before this one
    if (BNIN) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            isc[e] = make_float2(__ldg(in_scale + c0 + 2 * e), __ldg(in_scale + c0 + 2 * e + 1));
            ish[e] = make_float2(__ldg(in_shift + c0 + 2 * e), __ldg(in_shift + c0 + 2 * e + 1));
        }
    }
    const bool relu = (flags & TSS_EPI_RELU) != 0;
    TSS_MARK(3);

    for (int it = 0; tile < ntiles; ++it, tile += gridDim.x) {
        const int stage = it & 1;
        const int next = tile + gridDim.x;
        if (threadIdx.x == 0 && next < ntiles) issue(next, stage ^ 1);    // released by the __syncthreads of iteration it-1
        int t = tile;
        const int tw = t % tiles_w; t /= tiles_w;
        const int th = t % tiles_h;
        const int n = t / tiles_h;
        const int ho0 = th * TH, wo0 = tw * TW;
        mbar_wait(smem_u32(bars + stage), (uint32_t)(it >> 1) & 1);
        TSS_MARK_IF(threadIdx.x == 0 && it < 3, 4 + 3 * it);           // tile `it` has landed
        float2 acc[TH][4];
#pragma unroll
        for (int r = 0; r < TH; ++r) zero8p(acc[r]);
        const T* tp = (const T*)(smem + (size_t)stage * stage_bytes) + ((size_t)col * S) * CB + cg * 8;
        // kPadNan (BNIN with a ReLU): the tensor map fills out-of-image elements with NaN and fmaxf(NaN, 0) = 0 IS the
        // convolution's zero padding -- no coordinate test and no select behind the loads
        auto convolve = [&](auto pad_nan) {
            constexpr bool kPadNan = decltype(pad_nan)::value;
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
                    if (BNIN && kPadNan) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 a = ffma2(v[kx][e], isc[e], ish[e]);
                            v[kx][e] = make_float2(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f));
                        }
                    } else if (BNIN) {
                        const int hq = ho0 * S - D + j, wq = wo0 * S - D + col * S + kx * D;
                        const bool inside = hq >= 0 && hq < Hi && wq >= 0 && wq < Wi;
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float2 a = ffma2(v[kx][e], isc[e], ish[e]);
                            if (in_relu) a = make_float2(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f));
                            v[kx][e] = inside ? a : make_float2(0.f, 0.f);
                        }
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
                            for (int e = 0; e < 4; ++e) acc[r][e] = ffma2(v[kx][e], wr[ky * 3 + kx][e], acc[r][e]);
                    }
                }
            }
        };
        if (BNIN && pad_nan) convolve(DwTrue{}); else convolve(DwFalse{});
        TSS_MARK_IF(threadIdx.x == 0 && it < 3, 5 + 3 * it);           // ... computed
        const int wo = wo0 + col;
        if (wo < Wo) {
            T* yp = y + (((int64_t)n * Ho + ho0) * Wo + wo) * C + c0;
            float2 s1[4], s2[4];
            zero8p(s1); zero8p(s2);
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
                        for (int e = 0; e < 4; ++e) {
                            const float2 sc = scale != nullptr ? make_float2(__ldg(scale + c0 + 2 * e), __ldg(scale + c0 + 2 * e + 1)) : make_float2(1.f, 1.f);
                            const float2 sh = make_float2(__ldg(shift + c0 + 2 * e), __ldg(shift + c0 + 2 * e + 1));
                            acc[r][e] = ffma2(acc[r][e], sc, sh);
                        }
                    }
                    if (relu) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc[r][e] = make_float2(fmaxf(acc[r][e].x, 0.f), fmaxf(acc[r][e].y, 0.f));
                    }
                    store8p(yp + (int64_t)r * Wo * C, acc[r]);
                }
            }
            if (stats != nullptr) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    my1[2 * e] += s1[e].x; my1[2 * e + 1] += s1[e].y;
                    my2[2 * e] += s2[e].x; my2[2 * e + 1] += s2[e].y;
                }
            }
        }
        __syncthreads();                       // everyone is done with this stage: it may be refilled
        TSS_MARK_IF(threadIdx.x == 0 && it < 3, 6 + 3 * it);           // ... stored
    }
    TSS_MARK(13);
    if (stats != nullptr) {                    // (the loop's last __syncthreads published every thread's slots)
        for (int i = threadIdx.x; i < 2 * CB; i += blockDim.x) {
            const int which = i / CB, ch = i - which * CB;
            float s = 0.f;
            for (int cidx = 0; cidx < TW; ++cidx) s += part[(size_t)(which * TW + cidx) * CB + ch];
            atomicAdd(stats + which * C + cb0 + ch, (double)s);
        }
    }
    TSS_MARK(14);
}


// Launches the persistent kernel if the shape fits (`*launched`), otherwise leaves the call to the one-tile kernels.
template <typename T, int S, int D, int TH, bool FLIP, bool BNIN>
int dw_launch_persistent(const CUtensorMap& map, const float* w, void* y, int N, int Hi, int Wi, int Ho, int Wo, int C, int CB, int TW,
                         const float* scale, const float* shift, int flags, double* stats, const float* in_scale,
                         const float* in_shift, int in_relu, int pad_nan, cudaStream_t st, bool* launched) {
    *launched = false;
    static const int persist = [] { const char* e = getenv("TSS_DW_PERSIST"); return (e != nullptr && e[0] == '0') ? 0 : 1; }();
    static const int per_sm = [] { const char* e = getenv("TSS_DW_CTAS_PER_SM"); const int v = e != nullptr ? atoi(e) : 2; return v >= 1 && v <= 4 ? v : 2; }();
    constexpr int IH = DwGeo<S, D, TH>::IH;
    const int IW = (TW - 1) * S + 2 * D + 1;
    const int tiles_w = (Wo + TW - 1) / TW, tiles_h = (Ho + TH - 1) / TH;
    const int threads = (CB / 8) * TW;
    const size_t tile_bytes = (size_t)IH * IW * CB * sizeof(T);
    const size_t stage = (tile_bytes + 127) & ~(size_t)127;
    const int64_t ntiles = (int64_t)N * tiles_h * tiles_w;
    const size_t part = (size_t)2 * TW * CB * sizeof(float);
    size_t smem = 128 + 2 * stage + 16 + part + (size_t)9 * CB * sizeof(float);
    if (!persist || smem > 100 * 1024 || ntiles >= (1ll << 30)) return TSS_OK;
    // exactly `per_sm` CTAs per SM: one more must not fit (228 KB of shared memory per SM, 1 KB reserved per CTA)
    const size_t cap = (size_t)(228 * 1024) / (per_sm + 1) - 1024 + 256;
    if (smem < cap) smem = cap;
    if (threads > 192) return TSS_OK;
    auto kp = threads <= 128 ? dw_tma_persistent_kernel<T, S, D, TH, FLIP, BNIN, 128> : dw_tma_persistent_kernel<T, S, D, TH, FLIP, BNIN, 192>;
    static bool attr_set[2] = {false, false};          // per template instance; idempotent, benign race
    if (!attr_set[threads <= 128]) {
        TSS_CUDA(cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
        attr_set[threads <= 128] = true;
    }
    const int cblocks = C / CB;
    int64_t gx = ((int64_t)tss_num_sms() * per_sm) / cblocks;       // rounded DOWN: a CTA beyond the resident set would run alone, after the others
    if (gx > ntiles) gx = ntiles;
    if (gx < 1) gx = 1;
    dim3 grid((unsigned)gx, (unsigned)cblocks);
    tss_launch(kp, grid, threads, smem, st, map, w, (T*)y, Ho, Wo, C, CB, TW, tiles_w, tiles_h, (int)ntiles, (uint32_t)stage,
               scale, shift, flags, stats, in_scale, in_shift, in_relu, Hi, Wi, pad_nan);
    TSS_LAUNCH_CHECK(BNIN ? "dwconv3x3_fwd_bnin(tma, persistent)" : "dwconv3x3(tma, persistent)");
    *launched = true;
    return TSS_OK;
}
