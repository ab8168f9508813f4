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
