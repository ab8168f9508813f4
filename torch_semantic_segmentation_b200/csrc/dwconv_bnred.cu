// Depthwise 3x3 stride-1 dgrad with the BatchNorm-backward REDUCTION of the producing layer fused into
// its epilogue (training): the kernel of dwconv_tma.cu (TMA-staged halo tile of dy, flipped taps,
// FFMA2), whose output dz = gradient w.r.t. the activated input z = relu(BN(yp)) of this depthwise
// conv.  The epilogue reads the matching yp values, applies the ReLU mask recomputed from yp, stores
// g = dz * mask instead of dz and accumulates  sums[c] += sum g,  sums[C+c] += sum g * xhat  (what
// bn_bwd_reduce_kernel would compute in a pass of its own over dz and yp).
#include <stdlib.h>

#include "tma.cuh"

namespace {

constexpr int TH = 8;
constexpr int IH = TH + 2;

template <typename T>
__global__ void __launch_bounds__(192, 2)
dw_dgrad_bnred_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmY,
                      const float* __restrict__ w, T* __restrict__ g_out,
                      int H, int W, int C, int CB, int TW, int tiles_w, int tiles_h, const T* __restrict__ yp,
                      const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                      const float* __restrict__ beta, int relu, float* __restrict__ sums) {
    TSS_DYN_SMEM(uint8_t, smem_raw);
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    const int IW = TW + 2;
    const uint32_t tile_bytes = (uint32_t)IH * IW * CB * sizeof(T);
    T* tile = (T*)smem;
    // the producer's raw output for the tile's TH x TW pixels arrives by TMA too, under the same barrier: the epilogue
    // reads it from shared memory instead of starting a second round of global loads after the convolution
    const uint32_t ytile_bytes = (uint32_t)TH * TW * CB * sizeof(T);
    T* ytile = (T*)(smem + ((tile_bytes + 127) & ~127u));
    uint64_t* bar = (uint64_t*)((uint8_t*)ytile + ((ytile_bytes + 15) & ~15u));

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
        mbar_expect_tx(smem_u32(bar), tile_bytes + ytile_bytes);
        tma_load_4d(smem_u32(tile), &tmG, smem_u32(bar), cb0, w0 - 1, h0 - 1, n);
        tma_load_4d(smem_u32(ytile), &tmY, smem_u32(bar), cb0, w0, h0, n);
    }

    const int CGB = CB >> 3;
    const int cg = threadIdx.x % CGB, col = threadIdx.x / CGB;
    const int c0 = cb0 + cg * 8;
    float2 wr[9][4];                                      // flipped taps: dgrad of a stride-1 correlation
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
        for (int e = 0; e < 4; ++e)
            wr[k][e] = make_float2(__ldg(w + (c0 + 2 * e) * 9 + (8 - k)), __ldg(w + (c0 + 2 * e + 1) * 9 + (8 - k)));
    float mu[8], rs[8], sc[8], sh[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        mu[e] = __ldg(mean + c0 + e);
        rs[e] = __ldg(rstd + c0 + e);
        sc[e] = (gamma != nullptr ? __ldg(gamma + c0 + e) : 1.f) * rs[e];
        sh[e] = (beta != nullptr ? __ldg(beta + c0 + e) : 0.f) - mu[e] * sc[e];
    }

    mbar_wait(smem_u32(bar), 0);

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
    float s1[8], s2[8];
    zero8(s1); zero8(s2);
    if (wo < W) {
        const int64_t base = (((int64_t)n * H + h0) * W + wo) * C + c0;
#pragma unroll
        for (int r = 0; r < TH; ++r) {
            if (h0 + r < H) {
                float yy[8], g[8];
                load8_smem(ytile + ((size_t)r * TW + col) * CB + cg * 8, yy);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float dz = (e & 1) ? acc[r][e >> 1].y : acc[r][e >> 1].x;
                    const bool on = !relu || fmaf(yy[e], sc[e], sh[e]) > 0.f;
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
        atomicAdd(sums + which * C + cb0 + ch, s);
    }
}

// ---------------------------------------------------------------------------------------------
// Persistent variant (bf16, 128-thread CTAs): a CTA keeps its taps and the producer's BatchNorm constants in registers and
// walks the spatial tiles blockIdx.x, blockIdx.x + gridDim.x, ... of its channel block with a 2-stage TMA ring (gradient
// halo tile + producer tile of tile i+1 in flight while tile i is computed); the two sums live in per-thread shared-memory
// slots across all tiles and leave the CTA as one atomic per channel.  Two CTAs per SM by construction (2 x ~90 KB of
// shared memory): the one-tile kernel above re-loads 72 taps and 32 constants per tile with scalar loads (~2 us per CTA,
// tools/trace_kernels.py) and its 648 CTAs at 1/32 resolution land unevenly on the SMs under programmatic dependent launch.
template <typename T>
__global__ void __launch_bounds__(128, 2)
dw_dgrad_bnred_persistent_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmY,
                                 const float* __restrict__ w, T* __restrict__ g_out,
                                 int H, int W, int C, int CB, int TW, int tiles_w, int tiles_h, int ntiles,
                                 uint32_t stage_bytes, uint32_t ytile_off,
                                 const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, int relu, float* __restrict__ sums) {
    TSS_DYN_SMEM(uint8_t, smem_raw);
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    const int IW = TW + 2;
    const uint32_t tile_bytes = (uint32_t)IH * IW * CB * sizeof(T);
    const uint32_t ytile_bytes = (uint32_t)TH * TW * CB * sizeof(T);
    uint64_t* bars = (uint64_t*)(smem + 2 * (size_t)stage_bytes);
    float* part = (float*)(bars + 2);          // [2][TW][CB]: every thread's running sums (its private 16 slots)
    float* s_w = part + (size_t)2 * TW * CB;   // [9][CB] flipped taps
    const int cb0 = blockIdx.y * CB;
    TSS_MARK(0);
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(bars), 1);
        mbar_init(smem_u32(bars + 1), 1);
        mbar_fence_init();
    }
    dw_stage_taps<true>(w, cb0, CB, s_w);
    const int CGB = CB >> 3;
    const int cg = threadIdx.x % CGB, col = threadIdx.x / CGB;
    const int c0 = cb0 + cg * 8;
    float* my1 = part + (size_t)col * CB + cg * 8;
    float* my2 = part + (size_t)(TW + col) * CB + cg * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) { my1[e] = 0.f; my2[e] = 0.f; }
    __syncthreads();
    float2 wr[9][4];
    dw_take_taps(s_w, CB, cg, wr);
    TSS_MARK(1);
    if (threadIdx.x == 0) {tma_prefetch_desc(&tmG); tma_prefetch_desc(&tmY); }      // descriptor fetch (~0.5 us) under the predecessor's tail
    pdl_wait();
    TSS_MARK(2);

    auto issue = [&](int tile, int stage) {
        int t = tile;
        const int tw = t % tiles_w; t /= tiles_w;
        const int th = t % tiles_h;
        const int n = t / tiles_h;
        const uint32_t bar = smem_u32(bars + stage);
        uint8_t* base = smem + (size_t)stage * stage_bytes;
        mbar_expect_tx(bar, tile_bytes + ytile_bytes);
        tma_load_4d(smem_u32(base), &tmG, bar, cb0, tw * TW - 1, th * TH - 1, n);
        tma_load_4d(smem_u32(base + ytile_off), &tmY, bar, cb0, tw * TW, th * TH, n);
    };
    int tile = blockIdx.x;
    if (threadIdx.x == 0 && tile < ntiles) issue(tile, 0);

    float mu[8], rs[8], sc[8], sh[8];
    load8(mean + c0, mu);
    load8(rstd + c0, rs);
    if (gamma != nullptr) load8(gamma + c0, sc);
    if (beta != nullptr) load8(beta + c0, sh);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        sc[e] = (gamma != nullptr ? sc[e] : 1.f) * rs[e];
        sh[e] = (beta != nullptr ? sh[e] : 0.f) - mu[e] * sc[e];
    }
    TSS_MARK(3);

    for (int it = 0; tile < ntiles; ++it, tile += gridDim.x) {
        const int stage = it & 1;
        const int next = tile + gridDim.x;
        if (threadIdx.x == 0 && next < ntiles) issue(next, stage ^ 1);    // released by the __syncthreads of iteration it-1
        int t = tile;
        const int tw = t % tiles_w; t /= tiles_w;
        const int th = t % tiles_h;
        const int n = t / tiles_h;
        const int h0 = th * TH, w0 = tw * TW;
        mbar_wait(smem_u32(bars + stage), (uint32_t)(it >> 1) & 1);
        TSS_MARK_IF(threadIdx.x == 0 && it < 3, 4 + 3 * it);
        const T* halo = (const T*)(smem + (size_t)stage * stage_bytes);
        const T* ytile = (const T*)(smem + (size_t)stage * stage_bytes + ytile_off);
        float2 acc[TH][4];
#pragma unroll
        for (int r = 0; r < TH; ++r) zero8p(acc[r]);
        const T* tp = halo + (size_t)col * CB + cg * 8;
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
        TSS_MARK_IF(threadIdx.x == 0 && it < 3, 5 + 3 * it);
        const int wo = w0 + col;
        if (wo < W) {
            float s1[8], s2[8];
            zero8(s1); zero8(s2);
            const int64_t base = (((int64_t)n * H + h0) * W + wo) * C + c0;
#pragma unroll
            for (int r = 0; r < TH; ++r) {
                if (h0 + r < H) {
                    float yy[8], g[8];
                    load8_smem(ytile + ((size_t)r * TW + col) * CB + cg * 8, yy);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float dz = (e & 1) ? acc[r][e >> 1].y : acc[r][e >> 1].x;
                        const bool on = !relu || fmaf(yy[e], sc[e], sh[e]) > 0.f;
                        g[e] = on ? dz : 0.f;
                        s1[e] += g[e];
                        s2[e] = fmaf(g[e], (yy[e] - mu[e]) * rs[e], s2[e]);
                    }
                    store8(g_out + base + (int64_t)r * W * C, g);
                }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) { my1[e] += s1[e]; my2[e] += s2[e]; }
        }
        __syncthreads();                       // everyone is done with this stage: it may be refilled
        TSS_MARK_IF(threadIdx.x == 0 && it < 3, 6 + 3 * it);
    }
    TSS_MARK(13);
    for (int i = threadIdx.x; i < 2 * CB; i += blockDim.x) {      // (the loop's last __syncthreads published every thread's slots)
        const int which = i / CB, ch = i - which * CB;
        float s = 0.f;
        for (int cidx = 0; cidx < TW; ++cidx) s += part[(size_t)(which * TW + cidx) * CB + ch];
        atomicAdd(sums + which * C + cb0 + ch, s);
    }
    TSS_MARK(14);
}

// ---------------------------------------------------------------------------------------------
// Stride 2: the quad kernel of dwconv.cu (2x2 input pixels per 2x2 gradient neighbourhood, vertical
// strips of R quads, persistent grid with a loop-invariant channel group per thread) with the same fused
// reduction.  These are the largest BatchNorm-backward instances of the network (the producers are the
// stem and the first expand conv).
constexpr int kQuadThreads = 128;

template <typename T, int R, int kMinBlocks>
__global__ void __launch_bounds__(kQuadThreads, kMinBlocks)
dw_dgrad_s2_bnred_kernel(const T* __restrict__ dy, const float* __restrict__ w, T* __restrict__ g_out,
                         int N, int Hi, int Wi, int Ho, int Wo, int C, const T* __restrict__ yp,
                         const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                         const float* __restrict__ beta, int relu, float* __restrict__ sums) {
    TSS_DYN_SMEM(float, s_sum);        // [2*C] sums, [9][C] taps
    float* s_w = s_sum + 2 * C;
    for (int i = threadIdx.x; i < 2 * C; i += kQuadThreads) s_sum[i] = 0.f;
    // taps through shared memory with coalesced loads, before the wait (parameters: see dw_stage_taps in tma.cuh)
    dw_stage_taps<false>(w, 0, C, s_w);
    __syncthreads();
    const int CG = C >> 3;
    const int nstrips = (Ho + R - 1) / R;
    const int64_t total = (int64_t)N * nstrips * Wo * CG;
    const int64_t gstride = (int64_t)gridDim.x * kQuadThreads;
    int64_t item = (int64_t)blockIdx.x * kQuadThreads + threadIdx.x;
    const int c0 = (int)(item % CG) * 8;          // loop-invariant: gstride % CG == 0
    float wr[9][8];
#pragma unroll
    for (int k = 0; k < 9; ++k) load8_smem(s_w + k * C + c0, wr[k]);
    pdl_wait();
    // the loop accumulates sum g and sum g*y; x-hat comes in at the end: sum g*xhat = rstd*(sum g*y - mean*sum g)
    float sc[8], sh[8], s1[8], s2[8];
    zero8(s1); zero8(s2);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const float m_ = __ldg(mean + c0 + e), r_ = __ldg(rstd + c0 + e);
        sc[e] = (gamma != nullptr ? __ldg(gamma + c0 + e) : 1.f) * r_;
        sh[e] = (beta != nullptr ? __ldg(beta + c0 + e) : 0.f) - m_ * sc[e];
    }
    // mask, accumulate and store one output pixel's 8 channels (the producer's raw output arrives as a raw vector that
    // was requested a whole quad earlier)
    auto emit = [&](float (&o)[8], const Raw8<T>& ryy, int64_t off) {
        float yy[8];
        ryy.get(yy);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            if (relu && !(fmaf(yy[e], sc[e], sh[e]) > 0.f)) o[e] = 0.f;
            s1[e] += o[e];
            s2[e] = fmaf(o[e], yy[e], s2[e]);
        }
        store8(g_out + off, o);
    };

    for (; item < total; item += gstride) {
        int64_t t = item / CG;
        const int q = (int)(t % Wo); t /= Wo;
        const int strip = (int)(t % nstrips);
        const int n = (int)(t / nstrips);
        const int p0 = strip * R;
        const T* dyn = dy + (int64_t)n * Ho * Wo * C + c0;
        const int64_t xn = (int64_t)n * Hi * Wi * C + c0;
        const bool q1 = q + 1 < Wo;
        const bool col1 = 2 * q + 1 < Wi;
        float g0[2][8], g1[2][8];
        load8(dyn + ((int64_t)p0 * Wo + q) * C, g0[0]);
        if (q1) load8(dyn + ((int64_t)p0 * Wo + q + 1) * C, g0[1]); else zero8(g0[1]);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int p = p0 + r;
            if (p >= Ho) break;
            // all six loads of this quad row (2 gradient vectors of the next row, 4 producer outputs) are issued before
            // the first of them is used: six 128-bit requests in flight per thread instead of one or two
            const bool row1 = 2 * p + 1 < Hi;
            const int64_t off00 = xn + ((int64_t)(2 * p) * Wi + 2 * q) * C;
            const int64_t off10 = off00 + (int64_t)Wi * C;
            Raw8<T> rn[2], ry[4];
            rn[0].zero(); rn[1].zero();
            ry[0].ld(yp + off00);
            if (col1) ry[1].ld(yp + off00 + C);
            if (row1) {
                ry[2].ld(yp + off10);
                if (col1) ry[3].ld(yp + off10 + C);
            }
            if (p + 1 < Ho) {
                rn[0].ld(dyn + ((int64_t)(p + 1) * Wo + q) * C);
                if (q1) rn[1].ld(dyn + ((int64_t)(p + 1) * Wo + q + 1) * C);
            }
            rn[0].get(g1[0]);
            rn[1].get(g1[1]);
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = g0[0][e] * wr[4][e];
            emit(o, ry[0], off00);
            if (col1) {
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = fmaf(g0[1][e], wr[3][e], g0[0][e] * wr[5][e]);
                emit(o, ry[1], off00 + C);
            }
            if (row1) {
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = fmaf(g1[0][e], wr[1][e], g0[0][e] * wr[7][e]);
                emit(o, ry[2], off10);
                if (col1) {
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        o[e] = fmaf(g1[1][e], wr[0][e], fmaf(g1[0][e], wr[2][e], fmaf(g0[1][e], wr[6][e], g0[0][e] * wr[8][e])));
                    emit(o, ry[3], off10 + C);
                }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) { g0[0][e] = g1[0][e]; g0[1][e] = g1[1][e]; }
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const float m_ = __ldg(mean + c0 + e), r_ = __ldg(rstd + c0 + e);
        atomicAdd(&s_sum[c0 + e], s1[e]);
        atomicAdd(&s_sum[C + c0 + e], r_ * (s2[e] - m_ * s1[e]));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += kQuadThreads) {
        const float v = s_sum[i];
        if (v != 0.f) atomicAdd(sums + i, v);
    }
}

// ---------------------------------------------------------------------------------------------
// Stride 2 through TMA (bf16): the quad arithmetic above on the structure of the persistent stride-1 kernel.  A CTA owns a
// channel block and walks tiles of TQH x TQ quads (2 TQH x 2 TQ producer pixels) with a 2-stage ring: the gradient tile
// (TQH+1 x TQ+1, the +1 row / column are the lower / right neighbours of the last quads; beyond the map the TMA unit
// fills zeros) and the producer's raw-output tile of tile i+1 are in flight while tile i is computed, so the memory
// system sees two bulk requests of ~40 KB per CTA instead of six 16-byte loads per thread (the quad kernel ran at 0.40
// of the HBM peak on the 254 MB instance behind the stem: tools/time_ops.py).  Taps, BatchNorm constants and the two
// running sums stay in registers across tiles (a thread's channel group is fixed).
template <typename T, int TQH, int MAXT>
__global__ void __launch_bounds__(MAXT, 2)
dw_dgrad_s2_bnred_tma_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmY,
                             const float* __restrict__ w, T* __restrict__ g_out,
                             int Hi, int Wi, int Ho, int Wo, int C, int CB, int TQ, int tiles_w, int tiles_h, int ntiles,
                             uint32_t stage_bytes, uint32_t ytile_off,
                             const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                             const float* __restrict__ beta, int relu, float* __restrict__ sums) {
    TSS_DYN_SMEM(uint8_t, smem_raw);
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    const int GW = TQ + 1;                                   // gradient tile width
    const uint32_t gtile_bytes = (uint32_t)(TQH + 1) * GW * CB * sizeof(T);
    const uint32_t ytile_bytes = (uint32_t)(2 * TQH) * (2 * TQ) * CB * sizeof(T);
    uint64_t* bars = (uint64_t*)(smem + 2 * (size_t)stage_bytes);
    float* s_sum = (float*)(bars + 2);                       // [2][CB]
    float* s_w = s_sum + 2 * CB;                             // [9][CB] taps (not flipped: the quad formulas index them directly)
    const int cb0 = blockIdx.y * CB;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(bars), 1);
        mbar_init(smem_u32(bars + 1), 1);
        mbar_fence_init();
    }
    dw_stage_taps<false>(w, cb0, CB, s_w);
    for (int i = threadIdx.x; i < 2 * CB; i += blockDim.x) s_sum[i] = 0.f;
    const int CGB = CB >> 3;
    const int cg = threadIdx.x % CGB, col = threadIdx.x / CGB;
    const int c0 = cb0 + cg * 8;
    __syncthreads();
    float2 wr[9][4];
    dw_take_taps(s_w, CB, cg, wr);
    if (threadIdx.x == 0) { tma_prefetch_desc(&tmG); tma_prefetch_desc(&tmY); }
    pdl_wait();

    auto issue = [&](int tile, int stage) {
        int t = tile;
        const int tw = t % tiles_w; t /= tiles_w;
        const int th = t % tiles_h;
        const int n = t / tiles_h;
        const uint32_t bar = smem_u32(bars + stage);
        uint8_t* base = smem + (size_t)stage * stage_bytes;
        mbar_expect_tx(bar, gtile_bytes + ytile_bytes);
        tma_load_4d(smem_u32(base), &tmG, bar, cb0, tw * TQ, th * TQH, n);
        tma_load_4d(smem_u32(base + ytile_off), &tmY, bar, cb0, 2 * tw * TQ, 2 * th * TQH, n);
    };
    int tile = blockIdx.x;
    if (threadIdx.x == 0 && tile < ntiles) issue(tile, 0);

    float2 sc[4], sh[4], s1[4], s2[4];
    zero8p(s1); zero8p(s2);
    {
        float mu[8], rs[8], ga[8], be[8];
        load8(mean + c0, mu);
        load8(rstd + c0, rs);
        if (gamma != nullptr) load8(gamma + c0, ga);
        if (beta != nullptr) load8(beta + c0, be);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float a = (gamma != nullptr ? ga[e] : 1.f) * rs[e];
            const float b = (beta != nullptr ? be[e] : 0.f) - mu[e] * a;
            if (e & 1) { sc[e >> 1].y = a; sh[e >> 1].y = b; } else { sc[e >> 1].x = a; sh[e >> 1].x = b; }
        }
    }
    const float2 zero2 = make_float2(0.f, 0.f);
    // mask, accumulate (sum g and sum g*y; x-hat comes in at the end: sum g*xhat = rstd*(sum g*y - mean*sum g)) and store
    // one producer pixel's 8 channels
    auto emit = [&](float2 (&o)[4], const T* yptr, T* gptr) {
        float2 yy[4];
        load8p_smem(yptr, yy);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 z = ffma2(yy[e], sc[e], sh[e]);
            if (relu && !(z.x > 0.f)) o[e].x = 0.f;
            if (relu && !(z.y > 0.f)) o[e].y = 0.f;
            s1[e].x += o[e].x;
            s1[e].y += o[e].y;
            s2[e] = ffma2(o[e], yy[e], s2[e]);
        }
        store8p(gptr, o);
    };

    for (int it = 0; tile < ntiles; ++it, tile += gridDim.x) {
        const int stage = it & 1;
        const int next = tile + gridDim.x;
        if (threadIdx.x == 0 && next < ntiles) issue(next, stage ^ 1);    // released by the __syncthreads of iteration it-1
        int t = tile;
        const int tw = t % tiles_w; t /= tiles_w;
        const int th = t % tiles_h;
        const int n = t / tiles_h;
        mbar_wait(smem_u32(bars + stage), (uint32_t)(it >> 1) & 1);
        const T* gt = (const T*)(smem + (size_t)stage * stage_bytes) + (size_t)col * CB + cg * 8;
        const T* yt = (const T*)(smem + (size_t)stage * stage_bytes + ytile_off) + (size_t)(2 * col) * CB + cg * 8;
        const int q = tw * TQ + col;
        if (q < Wo) {
            const bool col1 = 2 * q + 1 < Wi;
            float2 g0[2][4], g1[2][4];
            load8p_smem(gt, g0[0]);
            load8p_smem(gt + CB, g0[1]);
#pragma unroll 1
            for (int pl = 0; pl < TQH; ++pl) {
                const int p = th * TQH + pl;
                if (p >= Ho) break;
                load8p_smem(gt + (size_t)(pl + 1) * GW * CB, g1[0]);
                load8p_smem(gt + (size_t)(pl + 1) * GW * CB + CB, g1[1]);
                const bool row1 = 2 * p + 1 < Hi;
                const T* y0 = yt + (size_t)(2 * pl) * (2 * TQ) * CB;
                const T* y1 = y0 + (size_t)(2 * TQ) * CB;
                T* o0 = g_out + (((int64_t)n * Hi + 2 * p) * Wi + 2 * q) * C + c0;
                T* o1 = o0 + (int64_t)Wi * C;
                float2 o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = ffma2(g0[0][e], wr[4][e], zero2);
                emit(o, y0, o0);
                if (col1) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) o[e] = ffma2(g0[1][e], wr[3][e], ffma2(g0[0][e], wr[5][e], zero2));
                    emit(o, y0 + CB, o0 + C);
                }
                if (row1) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) o[e] = ffma2(g1[0][e], wr[1][e], ffma2(g0[0][e], wr[7][e], zero2));
                    emit(o, y1, o1);
                    if (col1) {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            o[e] = ffma2(g1[1][e], wr[0][e], ffma2(g1[0][e], wr[2][e], ffma2(g0[1][e], wr[6][e], ffma2(g0[0][e], wr[8][e], zero2))));
                        emit(o, y1 + CB, o1 + C);
                    }
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) { g0[0][e] = g1[0][e]; g0[1][e] = g1[1][e]; }
            }
        }
        __syncthreads();                       // everyone is done with this stage: it may be refilled
    }
    {
        float mu[8], rs[8];
        load8(mean + c0, mu);
        load8(rstd + c0, rs);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float a = (e & 1) ? s1[e >> 1].y : s1[e >> 1].x;
            const float b = (e & 1) ? s2[e >> 1].y : s2[e >> 1].x;
            atomicAdd(&s_sum[cg * 8 + e], a);
            atomicAdd(&s_sum[CB + cg * 8 + e], rs[e] * (b - mu[e] * a));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * CB; i += blockDim.x) {
        const int which = i / CB, ch = i - which * CB;
        const float v = s_sum[i];
        if (v != 0.f) atomicAdd(sums + which * C + cb0 + ch, v);
    }
}

int gcd_int2(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }

template <typename T> struct TmaTypeB;
template <> struct TmaTypeB<float> { static constexpr CUtensorMapDataType v = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; };
template <> struct TmaTypeB<bf16> { static constexpr CUtensorMapDataType v = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; };

}  // namespace

bool tss_dw_tma_config(int C, int* CB, int* TW);     // dwconv_tma.cu

// stride 2, dilation 1: g has the INPUT geometry (Hi, Wi), dy the output geometry ((Hi-1)/2+1, (Wi-1)/2+1)
extern "C" int tss_dwconv3x3_dgrad_s2_bnred(const void* dy, const float* w, void* g, int N, int Hi, int Wi, int C,
                                            const void* yp, const float* mean, const float* rstd, const float* gamma,
                                            const float* beta, int flags, float* sums, int dtype, void* stream) {
    TSS_REQUIRE(N > 0 && Hi > 0 && Wi > 0 && C > 0 && C % 8 == 0 && C <= 2048, "dwconv3x3_dgrad_s2_bnred: bad shape N=%d H=%d W=%d C=%d", N, Hi, Wi, C);
    TSS_REQUIRE(yp != nullptr && mean != nullptr && rstd != nullptr && sums != nullptr, "dwconv3x3_dgrad_s2_bnred: missing BatchNorm operands");
    TSS_REQUIRE(((uintptr_t)dy & 15) == 0 && ((uintptr_t)g & 15) == 0 && ((uintptr_t)yp & 15) == 0, "dwconv3x3_dgrad_s2_bnred: buffers must be 16-byte aligned");
    const int Ho = (Hi - 1) / 2 + 1, Wo = (Wi - 1) / 2 + 1;
    {
        // bf16 with a TMA channel block: the persistent 2-stage TMA kernel (TSS_S2_TMA=0: the quad kernel below)
        static const int tma_env = [] { const char* e = getenv("TSS_S2_TMA"); return (e != nullptr && e[0] == '0') ? 0 : 1; }();
        int CB, TQ;
        TssEncodeTiledFn enc = tss_encode_tiled();
        if (tma_env && dtype == TSS_BF16 && enc != nullptr && tss_dw_tma_config(C, &CB, &TQ) && (int64_t)N * Ho * Wo < (1ll << 30)) {
            typedef bf16 T;
            const int threads = (CB / 8) * TQ;
            auto stage_of = [&](int tqh, size_t* yoff) {
                const size_t gt = (size_t)(tqh + 1) * (TQ + 1) * CB * sizeof(T), yt = (size_t)(2 * tqh) * (2 * TQ) * CB * sizeof(T);
                *yoff = (gt + 127) & ~(size_t)127;
                return *yoff + ((yt + 127) & ~(size_t)127);
            };
            static const int tqh_env = [] { const char* e = getenv("TSS_S2_TQH"); return e ? atoi(e) : 0; }();
            size_t yoff;
            int tqh = 4;
            if (tqh_env == 2 || 2 * stage_of(4, &yoff) > 100 * 1024) tqh = 2;      // two CTAs per SM
            const size_t stage = stage_of(tqh, &yoff);
            const size_t smem = 128 + 2 * stage + 16 + (size_t)11 * CB * sizeof(float);
            CUtensorMap mg, my;
            cuuint32_t estr[4] = {1, 1, 1, 1};
            {
                cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)N};
                cuuint64_t gstr[3] = {(cuuint64_t)C * sizeof(T), (cuuint64_t)Wo * C * sizeof(T), (cuuint64_t)Ho * Wo * C * sizeof(T)};
                cuuint32_t box[4] = {(cuuint32_t)CB, (cuuint32_t)(TQ + 1), (cuuint32_t)(tqh + 1), 1};
                CUresult r = enc(&mg, TmaTypeB<T>::v, 4, const_cast<void*>(dy), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                TSS_REQUIRE(r == CUDA_SUCCESS, "dwconv3x3_dgrad_s2_bnred: cuTensorMapEncodeTiled (dy) failed (%d)", (int)r);
            }
            {
                cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)Wi, (cuuint64_t)Hi, (cuuint64_t)N};
                cuuint64_t gstr[3] = {(cuuint64_t)C * sizeof(T), (cuuint64_t)Wi * C * sizeof(T), (cuuint64_t)Hi * Wi * C * sizeof(T)};
                cuuint32_t box[4] = {(cuuint32_t)CB, (cuuint32_t)(2 * TQ), (cuuint32_t)(2 * tqh), 1};
                CUresult r = enc(&my, TmaTypeB<T>::v, 4, const_cast<void*>(yp), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                TSS_REQUIRE(r == CUDA_SUCCESS, "dwconv3x3_dgrad_s2_bnred: cuTensorMapEncodeTiled (yp) failed (%d)", (int)r);
            }
            const int tiles_w = (Wo + TQ - 1) / TQ, tiles_h = (Ho + tqh - 1) / tqh;
            const int64_t ntiles = (int64_t)N * tiles_h * tiles_w;
            const int cblocks = C / CB;
            int64_t gx = ((int64_t)tss_num_sms() * 2) / cblocks;
            if (gx > ntiles) gx = ntiles;
            if (gx < 1) gx = 1;
            // 128-thread blocks (C % 64 == 0, C % 32 == 0) may use 255 registers at two CTAs per SM, 192-thread ones 168
            auto kern = threads <= 128 ? (tqh == 4 ? dw_dgrad_s2_bnred_tma_kernel<T, 4, 128> : dw_dgrad_s2_bnred_tma_kernel<T, 2, 128>)
                                       : (tqh == 4 ? dw_dgrad_s2_bnred_tma_kernel<T, 4, 192> : dw_dgrad_s2_bnred_tma_kernel<T, 2, 192>);
            static bool attr_set_t[4] = {false, false, false, false};
            const int ki = (threads <= 128 ? 0 : 2) + (tqh == 4 ? 1 : 0);
            if (!attr_set_t[ki]) {
                TSS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
                attr_set_t[ki] = true;
            }
            tss_launch(kern, dim3((unsigned)gx, (unsigned)cblocks), threads, smem, (cudaStream_t)stream, mg, my, w, (T*)g, Hi, Wi, Ho, Wo, C, CB, TQ,
                       tiles_w, tiles_h, (int)ntiles, (uint32_t)stage, (uint32_t)yoff, mean, rstd, gamma, beta, flags & TSS_EPI_RELU, sums);
            TSS_LAUNCH_CHECK("dwconv3x3_dgrad_s2_bnred(tma)");
            return TSS_OK;
        }
    }
    constexpr int R = 4;
    const int CG = C / 8;
    const int64_t total = (int64_t)N * ((Ho + R - 1) / R) * Wo * CG;
    // persistent grid whose total thread count is a multiple of CG (loop-invariant channel group per thread)
    const int qd = CG / gcd_int2(CG, kQuadThreads);
    int64_t grid = ceil_div64(total, kQuadThreads);
    // 3 resident CTAs of 128 threads per SM (168 registers, a few spills) or 2 (no spills): TSS_S2_OCC picks, default 3
    static const int occ = [] { const char* e = getenv("TSS_S2_OCC"); return (e != nullptr && e[0] == '2') ? 2 : 3; }();
    const int64_t cap = (int64_t)tss_num_sms() * occ;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    grid = (grid + qd - 1) / qd * qd;
    TSS_DISPATCH_DTYPE(dtype, "dwconv3x3_dgrad_s2_bnred", {
        auto kern = occ == 2 ? dw_dgrad_s2_bnred_kernel<T, R, 2> : dw_dgrad_s2_bnred_kernel<T, R, 3>;
        if ((size_t)11 * C * sizeof(float) > 48 * 1024) {
            static bool attr_set[2] = {false, false};
            if (!attr_set[occ == 2]) {
                TSS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 11 * 2048 * (int)sizeof(float)));
                attr_set[occ == 2] = true;
            }
        }
        tss_launch(kern, (unsigned)grid, kQuadThreads, (size_t)11 * C * sizeof(float), (cudaStream_t)stream,
                   (const T*)dy, w, (T*)g, N, Hi, Wi, Ho, Wo, C, (const T*)yp, mean, rstd, gamma, beta, flags & TSS_EPI_RELU, sums);
        TSS_LAUNCH_CHECK("dwconv3x3_dgrad_s2_bnred");
        return TSS_OK;
    });
}

extern "C" int tss_dwconv3x3_dgrad_bnred(const void* dy, const float* w, void* g, int N, int H, int W, int C,
                                         const void* yp, const float* mean, const float* rstd, const float* gamma,
                                         const float* beta, int flags, float* sums, int dtype, void* stream) {
    TSS_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "dwconv3x3_dgrad_bnred: bad shape N=%d H=%d W=%d C=%d", N, H, W, C);
    TSS_REQUIRE(yp != nullptr && mean != nullptr && rstd != nullptr && sums != nullptr, "dwconv3x3_dgrad_bnred: missing BatchNorm operands");
    TSS_REQUIRE(((uintptr_t)dy & 15) == 0 && ((uintptr_t)g & 15) == 0 && ((uintptr_t)yp & 15) == 0, "dwconv3x3_dgrad_bnred: buffers must be 16-byte aligned");
    int CB, TW;
    TSS_REQUIRE(tss_dw_tma_config(C, &CB, &TW), "dwconv3x3_dgrad_bnred: no channel block for C=%d", C);
    TssEncodeTiledFn enc = tss_encode_tiled();
    TSS_REQUIRE(enc != nullptr, "dwconv3x3_dgrad_bnred: cuTensorMapEncodeTiled is not available from the driver");
    TSS_DISPATCH_DTYPE(dtype, "dwconv3x3_dgrad_bnred", {
        if (sizeof(T) == 4 && TW == 32) TW = 16;
        const int IW = TW + 2;
        CUtensorMap map;
        cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
        cuuint64_t gstr[3] = {(cuuint64_t)C * sizeof(T), (cuuint64_t)W * C * sizeof(T), (cuuint64_t)H * W * C * sizeof(T)};
        cuuint32_t box[4] = {(cuuint32_t)CB, (cuuint32_t)IW, (cuuint32_t)IH, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&map, TmaTypeB<T>::v, 4, const_cast<void*>(dy), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        TSS_REQUIRE(r == CUDA_SUCCESS, "dwconv3x3_dgrad_bnred: cuTensorMapEncodeTiled failed (%d)", (int)r);
        CUtensorMap mapy;
        cuuint32_t boxy[4] = {(cuuint32_t)CB, (cuuint32_t)TW, (cuuint32_t)TH, 1};
        r = enc(&mapy, TmaTypeB<T>::v, 4, const_cast<void*>(yp), gdim, gstr, boxy, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        TSS_REQUIRE(r == CUDA_SUCCESS, "dwconv3x3_dgrad_bnred: cuTensorMapEncodeTiled (yp) failed (%d)", (int)r);
        const int tiles_w = (W + TW - 1) / TW, tiles_h = (H + TH - 1) / TH;
        const int threads = (CB / 8) * TW;
        const size_t tile_bytes = (size_t)IH * IW * CB * sizeof(T);
        const size_t ytile_bytes = (size_t)TH * TW * CB * sizeof(T);
        {
            // persistent CTAs with a 2-stage TMA ring, two per SM (TSS_DW_PERSIST=0: the one-tile kernel below)
            static const int persist = [] { const char* e = getenv("TSS_DW_PERSIST"); return (e != nullptr && e[0] == '0') ? 0 : 1; }();
            const size_t ytile_off = (tile_bytes + 127) & ~(size_t)127;
            const size_t stage = ytile_off + ((ytile_bytes + 127) & ~(size_t)127);
            const int64_t ntiles = (int64_t)N * tiles_h * tiles_w;
            size_t smem_p = 128 + 2 * stage + 16 + (size_t)2 * TW * CB * sizeof(float) + (size_t)9 * CB * sizeof(float);
            if (persist && sizeof(T) == 2 && threads <= 128 && smem_p <= 112 * 1024 && ntiles < (1ll << 30)) {
                const size_t cap = (size_t)(228 * 1024) / 3 - 1024 + 256;        // a third CTA must not fit on the SM
                if (smem_p < cap) smem_p = cap;
                auto kp = dw_dgrad_bnred_persistent_kernel<T>;
                static bool attr_set_p = false;
                if (!attr_set_p) {
                    TSS_CUDA(cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
                    attr_set_p = true;
                }
                const int cblocks = C / CB;
                int64_t gx = ((int64_t)tss_num_sms() * 2) / cblocks;     // rounded DOWN: a CTA beyond the resident set would run alone
                if (gx > ntiles) gx = ntiles;
                if (gx < 1) gx = 1;
                tss_launch(kp, dim3((unsigned)gx, (unsigned)cblocks), threads, smem_p, (cudaStream_t)stream, map, mapy, w, (T*)g, H, W, C, CB, TW,
                           tiles_w, tiles_h, (int)ntiles, (uint32_t)stage, (uint32_t)ytile_off, mean, rstd, gamma, beta,
                           flags & TSS_EPI_RELU, sums);
                TSS_LAUNCH_CHECK("dwconv3x3_dgrad_bnred(persistent)");
                return TSS_OK;
            }
        }
        size_t smem = 128 + ((tile_bytes + 127) & ~(size_t)127) + ((ytile_bytes + 15) & ~(size_t)15) + 16;
        const size_t part = (size_t)2 * TW * CB * sizeof(float) + 128;
        if (smem < part) smem = part;
        auto kern = dw_dgrad_bnred_kernel<T>;
        static bool attr_set = false;
        if (!attr_set) {
            TSS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            attr_set = true;
        }
        dim3 grid((unsigned)((int64_t)N * tiles_h * tiles_w), (unsigned)(C / CB));
        tss_launch(kern, grid, threads, smem, (cudaStream_t)stream, map, mapy, w, (T*)g, H, W, C, CB, TW, tiles_w, tiles_h, (const T*)yp,
                   mean, rstd, gamma, beta, flags & TSS_EPI_RELU, sums);
        TSS_LAUNCH_CHECK("dwconv3x3_dgrad_bnred");
        return TSS_OK;
    });
}
