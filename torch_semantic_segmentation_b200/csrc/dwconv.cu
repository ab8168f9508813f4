// Depthwise 3x3 convolution, NHWC, padding == dilation.  fwd / dgrad / wgrad.
//
// Memory-bound (9 MAC per output element): every thread owns 8 channels (one 128-bit
// access in bf16) and a vertical strip of R output rows, sliding a 3-row register window
// down the strip so that each input row is fetched once per strip; horizontal overlap is
// served by L1.  The grid is persistent and sized so that a thread's channel group is
// loop-invariant: the 72 weights stay in registers and the BatchNorm statistics are
// flushed once per CTA.
#include "common.cuh"

// dwconv_tma.cu: TMA-staged forward / stride-1 dgrad; returns -1 when the shape is not covered
int tss_dwconv3x3_tma(const void* x, const float* w, void* y, int N, int Hi, int Wi, int C, int stride, int dilation,
                      bool flip, const float* scale, const float* shift, int flags, double* stats, int dtype,
                      cudaStream_t st);

int tss_dwconv3x3_wgrad_tma(const void* x, const void* dy, float* dw, int N, int Hi, int Wi, int C, int stride,
                            int dilation, int dtype, cudaStream_t st);

namespace {

constexpr int kThreads = 256;

template <int S, int D, int R>
struct DwGeom {
    static constexpr int NR = (R - 1) * S + 2 * D + 1;   // input rows touched by a strip
};

// ---------------------------------------------------------------- forward ------------
// FLIP=true turns the kernel into the stride-1 dgrad (correlation with the flipped taps).
template <typename T, int S, int D, int R, bool FLIP>
__global__ void __launch_bounds__(kThreads)
dw_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w, T* __restrict__ y,
              int N, int Hi, int Wi, int Ho, int Wo, int C,
              const float* __restrict__ scale, const float* __restrict__ shift, int flags,
              double* __restrict__ stats) {
    pdl_wait();
    TSS_DYN_SMEM(float, s_stats);        // [2*C] when stats != nullptr
    const int CG = C >> 3;
    const int nstrips = (Ho + R - 1) / R;
    const int64_t total = (int64_t)N * nstrips * Wo * CG;
    const int64_t gstride = (int64_t)gridDim.x * kThreads;
    int64_t item = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int c0 = (int)(item % CG) * 8;          // loop-invariant: gstride % CG == 0

    float wr[9][8];
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
        for (int e = 0; e < 8; ++e) wr[k][e] = __ldg(w + (c0 + e) * 9 + (FLIP ? 8 - k : k));

    float s1[8], s2[8];
    zero8(s1); zero8(s2);
    if (stats != nullptr) {
        for (int i = threadIdx.x; i < 2 * C; i += kThreads) s_stats[i] = 0.f;
        __syncthreads();
    }
    const bool relu = (flags & TSS_EPI_RELU) != 0;

    for (; item < total; item += gstride) {
        int64_t t = item / CG;
        const int wo = (int)(t % Wo); t /= Wo;
        const int strip = (int)(t % nstrips);
        const int n = (int)(t / nstrips);
        const int ho0 = strip * R;
        const int hi_base = ho0 * S - D;
        const T* xn = x + (int64_t)n * Hi * Wi * C + c0;

        float acc[R][8];
#pragma unroll
        for (int r = 0; r < R; ++r) zero8(acc[r]);

#pragma unroll
        for (int j = 0; j < DwGeom<S, D, R>::NR; ++j) {
            bool used = false;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int tt = j - ky * D;
                if (tt >= 0 && tt % S == 0 && tt / S < R) used = true;
            }
            if (!used) continue;
            const int hi = hi_base + j;
            const bool row_ok = hi >= 0 && hi < Hi;
            float v[3][8];
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int wi = wo * S - D + kx * D;
                if (row_ok && wi >= 0 && wi < Wi) load8(xn + ((int64_t)hi * Wi + wi) * C, v[kx]);
                else zero8(v[kx]);
            }
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int tt = j - ky * D;
                if (tt >= 0 && tt % S == 0 && tt / S < R) {
                    const int r = tt / S;
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                        for (int e = 0; e < 8; ++e)
                            acc[r][e] = fmaf(v[kx][e], wr[ky * 3 + kx][e], acc[r][e]);
                }
            }
        }

        T* yn = y + (((int64_t)n * Ho + ho0) * Wo + wo) * C + c0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (ho0 + r < Ho) {
                if (stats != nullptr) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) { s1[e] += acc[r][e]; s2[e] = fmaf(acc[r][e], acc[r][e], s2[e]); }
                }
                if (shift != nullptr) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float sc = scale != nullptr ? __ldg(scale + c0 + e) : 1.f;
                        acc[r][e] = fmaf(acc[r][e], sc, __ldg(shift + c0 + e));
                    }
                }
                if (relu) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[r][e] = fmaxf(acc[r][e], 0.f);
                }
                store8(yn + (int64_t)r * Wo * C, acc[r]);
            }
        }
    }

    if (stats != nullptr) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            atomicAdd(&s_stats[c0 + e], s1[e]);
            atomicAdd(&s_stats[C + c0 + e], s2[e]);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * C; i += kThreads) {
            const float v = s_stats[i];
            if (v != 0.f) atomicAdd(stats + i, (double)v);
        }
    }
}

// ---------------------------------------------------------------- dgrad, stride 2, quads ----
// Stride 2 / dilation 1: the four input pixels (2p+a, 2q+b) of a quad see disjoint tap subsets of
// the 2x2 output neighbourhood g[p..p+1][q..q+1] (9 FMAs per quad per channel, no divisibility
// tests, no divergence).  A thread owns 8 channels and a vertical strip of R quads, sliding the
// two gradient rows down the strip; horizontal neighbours share their loads through L1.
template <typename T, int R>
__global__ void __launch_bounds__(kThreads, 2)
dw_dgrad_s2_quad_kernel(const T* __restrict__ dy, const float* __restrict__ w, T* __restrict__ dx,
                        int N, int Hi, int Wi, int Ho, int Wo, int C) {
    pdl_wait();
    const int CG = C >> 3;
    const int nstrips = (Ho + R - 1) / R;
    const int64_t total = (int64_t)N * nstrips * Wo * CG;
    const int64_t gstride = (int64_t)gridDim.x * kThreads;
    int64_t item = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int c0 = (int)(item % CG) * 8;          // loop-invariant: gstride % CG == 0
    float wr[9][8];
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
        for (int e = 0; e < 8; ++e) wr[k][e] = __ldg(w + (c0 + e) * 9 + k);

    for (; item < total; item += gstride) {
        int64_t t = item / CG;
        const int q = (int)(t % Wo); t /= Wo;
        const int strip = (int)(t % nstrips);
        const int n = (int)(t / nstrips);
        const int p0 = strip * R;
        const T* dyn = dy + (int64_t)n * Ho * Wo * C + c0;
        T* dxn = dx + (int64_t)n * Hi * Wi * C + c0;
        const bool q1 = q + 1 < Wo;
        const bool col1 = 2 * q + 1 < Wi;
        float g0[2][8], g1[2][8];                  // rows p and p+1, columns q and q+1
        load8(dyn + ((int64_t)p0 * Wo + q) * C, g0[0]);
        if (q1) load8(dyn + ((int64_t)p0 * Wo + q + 1) * C, g0[1]); else zero8(g0[1]);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int p = p0 + r;
            if (p >= Ho) break;
            if (p + 1 < Ho) {
                load8(dyn + ((int64_t)(p + 1) * Wo + q) * C, g1[0]);
                if (q1) load8(dyn + ((int64_t)(p + 1) * Wo + q + 1) * C, g1[1]); else zero8(g1[1]);
            } else { zero8(g1[0]); zero8(g1[1]); }
            float o[8];
            T* row0 = dxn + ((int64_t)(2 * p) * Wi + 2 * q) * C;
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = g0[0][e] * wr[4][e];
            store8(row0, o);
            if (col1) {
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = fmaf(g0[1][e], wr[3][e], g0[0][e] * wr[5][e]);
                store8(row0 + C, o);
            }
            if (2 * p + 1 < Hi) {
                T* row1 = row0 + (int64_t)Wi * C;
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = fmaf(g1[0][e], wr[1][e], g0[0][e] * wr[7][e]);
                store8(row1, o);
                if (col1) {
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        o[e] = fmaf(g1[1][e], wr[0][e], fmaf(g1[0][e], wr[2][e], fmaf(g0[1][e], wr[6][e], g0[0][e] * wr[8][e])));
                    store8(row1 + C, o);
                }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) { g0[0][e] = g1[0][e]; g0[1][e] = g1[1][e]; }
        }
    }
}

// ---------------------------------------------------------------- wgrad --------------
// Block = CG channel groups x PL pixel lanes.  Each thread accumulates its 8 channels x 9
// taps over a grid-strided set of output pixels, then the PL lanes are reduced through
// shared memory (warp shuffles first when lanes of one warp share a channel group) and
// one fp32 atomic per (channel, tap) per CTA goes to global memory.
template <typename T, int S, int D>
__global__ void __launch_bounds__(kThreads)
dw_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw,
                int N, int Hi, int Wi, int Ho, int Wo, int C, int PL) {
    pdl_wait();
    TSS_DYN_SMEM(float, s_red);        // [PL][C]
    const int CG = C >> 3;
    const int cg = threadIdx.x % CG;
    const int pl = threadIdx.x / CG;
    const int c0 = cg * 8;
    const int64_t npix = (int64_t)N * Ho * Wo;
    const int64_t lanes = (int64_t)gridDim.x * PL;

    float acc[9][8];
#pragma unroll
    for (int k = 0; k < 9; ++k) zero8(acc[k]);

    if (pl < PL) {
        for (int64_t p = (int64_t)blockIdx.x * PL + pl; p < npix; p += lanes) {
            int64_t t = p;
            const int wo = (int)(t % Wo); t /= Wo;
            const int ho = (int)(t % Ho);
            const int n = (int)(t / Ho);
            float g[8];
            load8(dy + p * C + c0, g);
            const T* xn = x + (int64_t)n * Hi * Wi * C + c0;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int hi = ho * S - D + ky * D;
                if (hi < 0 || hi >= Hi) continue;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int wi = wo * S - D + kx * D;
                    if (wi < 0 || wi >= Wi) continue;
                    float v[8];
                    load8(xn + ((int64_t)hi * Wi + wi) * C, v);
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[ky * 3 + kx][e] = fmaf(v[e], g[e], acc[ky * 3 + kx][e]);
                }
            }
        }
    }

    // warp-shuffle pre-reduction when CG divides 32: lanes l and l+CG share channels
    const bool shuffle_ok = (CG <= 16) && ((32 % CG) == 0) && ((blockDim.x & 31) == 0);
    if (shuffle_ok) {
#pragma unroll
        for (int k = 0; k < 9; ++k)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float v = acc[k][e];
                for (int o = 16; o >= CG; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                acc[k][e] = v;
            }
    }
    const int lane = threadIdx.x & 31;
    const bool writer = pl < PL && (!shuffle_ok || lane < CG);
    const int rows = shuffle_ok ? (blockDim.x >> 5) : PL;          // partial rows in s_red
    const int row = shuffle_ok ? (threadIdx.x >> 5) : pl;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        if (writer) {
#pragma unroll
            for (int e = 0; e < 8; ++e) s_red[row * C + c0 + e] = acc[k][e];
        }
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            float s = 0.f;
            for (int r = 0; r < rows; ++r) s += s_red[r * C + c];
            atomicAdd(dw + c * 9 + k, s);
        }
        __syncthreads();
    }
}

int gcd_int(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }

// Persistent grid whose total thread count is a multiple of CG = C/8, so that every
// thread keeps the same channel group across its grid-stride loop.
int persistent_grid(int64_t needed_ctas, int per_sm, int CG) {
    const int q = CG / gcd_int(CG, kThreads);
    int64_t cap = (int64_t)tss_num_sms() * per_sm;
    int64_t g = needed_ctas < cap ? needed_ctas : cap;
    if (g < 1) g = 1;
    g = (g + q - 1) / q * q;
    return (int)g;
}

template <typename T, int S, int D, int R, bool FLIP>
int launch_fwd(const void* x, const float* w, void* y, int N, int Hi, int Wi, int Ho, int Wo, int C,
               const float* scale, const float* shift, int flags, double* stats, cudaStream_t st) {
    const int nstrips = (Ho + R - 1) / R;
    const int64_t total = (int64_t)N * nstrips * Wo * (C / 8);
    const int grid = persistent_grid(ceil_div64(total, kThreads), 4, C / 8);
    const size_t smem = stats ? (size_t)2 * C * sizeof(float) : 0;
    tss_launch(dw_fwd_kernel<T, S, D, R, FLIP>, grid, kThreads, smem, st, 
        (const T*)x, w, (T*)y, N, Hi, Wi, Ho, Wo, C, scale, shift, flags, stats);
    TSS_LAUNCH_CHECK("dwconv3x3_fwd");
    return TSS_OK;
}

int check_common(const char* name, int N, int Hi, int Wi, int C, int stride, int dilation) {
    TSS_REQUIRE(N > 0 && Hi > 0 && Wi > 0, "%s: empty tensor N=%d H=%d W=%d", name, N, Hi, Wi);
    TSS_REQUIRE(C > 0 && C % 8 == 0 && C <= 2048, "%s: C=%d must be a multiple of 8 (<= 2048)", name, C);
    TSS_REQUIRE((stride == 1 && (dilation == 1 || dilation == 2 || dilation == 4)) || (stride == 2 && dilation == 1),
                "%s: unsupported stride=%d dilation=%d (supported: s1 d1/d2/d4, s2 d1)", name, stride, dilation);
    return TSS_OK;
}

}  // namespace

extern "C" int tss_dwconv3x3_fwd(const void* x, const float* w, void* y, int N, int Hi, int Wi, int C,
                                 int stride, int dilation, const float* scale, const float* shift,
                                 int flags, double* stats, int dtype, void* stream) {
    if (int e = check_common("dwconv3x3_fwd", N, Hi, Wi, C, stride, dilation)) return e;
    TSS_REQUIRE(scale == nullptr || shift != nullptr, "dwconv3x3_fwd: scale without shift");
    cudaStream_t st = (cudaStream_t)stream;
    const int Ho = (Hi - 1) / stride + 1, Wo = (Wi - 1) / stride + 1;
    {   // TMA-staged halo-tile kernel for every shape it covers; the register kernel below is the
        // general path (channel counts without a 32/48/64/96 block, dilation 2)
        const int r = tss_dwconv3x3_tma(x, w, y, N, Hi, Wi, C, stride, dilation, false, scale, shift, flags, stats, dtype, st);
        if (r >= 0) return r;
    }
    TSS_DISPATCH_DTYPE(dtype, "dwconv3x3_fwd", {
        if (stride == 1 && dilation == 1) return launch_fwd<T, 1, 1, 4, false>(x, w, y, N, Hi, Wi, Ho, Wo, C, scale, shift, flags, stats, st);
        if (stride == 2 && dilation == 1) return launch_fwd<T, 2, 1, 4, false>(x, w, y, N, Hi, Wi, Ho, Wo, C, scale, shift, flags, stats, st);
        if (stride == 1 && dilation == 2) return launch_fwd<T, 1, 2, 4, false>(x, w, y, N, Hi, Wi, Ho, Wo, C, scale, shift, flags, stats, st);
        return launch_fwd<T, 1, 4, 4, false>(x, w, y, N, Hi, Wi, Ho, Wo, C, scale, shift, flags, stats, st);
    });
}

extern "C" int tss_dwconv3x3_dgrad(const void* dy, const float* w, void* dx, int N, int Hi, int Wi, int C,
                                   int stride, int dilation, int dtype, void* stream) {
    if (int e = check_common("dwconv3x3_dgrad", N, Hi, Wi, C, stride, dilation)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    const int Ho = (Hi - 1) / stride + 1, Wo = (Wi - 1) / stride + 1;
    if (stride == 1) {
        const int r = tss_dwconv3x3_tma(dy, w, dx, N, Hi, Wi, C, 1, dilation, true, nullptr, nullptr, 0, nullptr, dtype, st);
        if (r >= 0) return r;
    }
    TSS_DISPATCH_DTYPE(dtype, "dwconv3x3_dgrad", {
        // stride 1: dx = conv(dy, flipped taps), same padding
        if (stride == 1 && dilation == 1) return launch_fwd<T, 1, 1, 4, true>(dy, w, dx, N, Hi, Wi, Hi, Wi, C, nullptr, nullptr, 0, nullptr, st);
        if (stride == 1 && dilation == 2) return launch_fwd<T, 1, 2, 4, true>(dy, w, dx, N, Hi, Wi, Hi, Wi, C, nullptr, nullptr, 0, nullptr, st);
        if (stride == 1 && dilation == 4) return launch_fwd<T, 1, 4, 4, true>(dy, w, dx, N, Hi, Wi, Hi, Wi, C, nullptr, nullptr, 0, nullptr, st);
        constexpr int R = 4;
        const int64_t total = (int64_t)N * ((Ho + R - 1) / R) * Wo * (C / 8);
        const int grid = persistent_grid(ceil_div64(total, kThreads), 6, C / 8);
        tss_launch(dw_dgrad_s2_quad_kernel<T, R>, grid, kThreads, 0, st, (const T*)dy, w, (T*)dx, N, Hi, Wi, Ho, Wo, C);
        TSS_LAUNCH_CHECK("dwconv3x3_dgrad");
        return TSS_OK;
    });
}

extern "C" int tss_dwconv3x3_wgrad(const void* x, const void* dy, float* dw, int N, int Hi, int Wi, int C,
                                   int stride, int dilation, int dtype, void* stream) {
    if (int e = check_common("dwconv3x3_wgrad", N, Hi, Wi, C, stride, dilation)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    const int Ho = (Hi - 1) / stride + 1, Wo = (Wi - 1) / stride + 1;
    {   // persistent TMA-pipelined kernel for every shape it covers
        const int r = tss_dwconv3x3_wgrad_tma(x, dy, dw, N, Hi, Wi, C, stride, dilation, dtype, st);
        if (r >= 0) return r;
    }
    const int CG = C / 8;
    TSS_REQUIRE(CG <= kThreads, "dwconv3x3_wgrad: C=%d too large", C);
    const int PL = kThreads / CG;
    int threads = PL * CG;
    if ((32 % CG) == 0) threads = kThreads;       // power-of-two groups: full warps for shuffles
    const int64_t npix = (int64_t)N * Ho * Wo;
    int64_t want = ceil_div64(npix, (int64_t)PL * 8);          // >= 8 pixels per lane
    int64_t cap = (int64_t)tss_num_sms() * 3;
    const int grid = (int)(want < 1 ? 1 : (want < cap ? want : cap));
    const size_t smem = (size_t)PL * C * sizeof(float);
    TSS_DISPATCH_DTYPE(dtype, "dwconv3x3_wgrad", {
        if (stride == 1 && dilation == 1) tss_launch(dw_wgrad_kernel<T, 1, 1>, grid, threads, smem, st, (const T*)x, (const T*)dy, dw, N, Hi, Wi, Ho, Wo, C, PL);
        else if (stride == 2) tss_launch(dw_wgrad_kernel<T, 2, 1>, grid, threads, smem, st, (const T*)x, (const T*)dy, dw, N, Hi, Wi, Ho, Wo, C, PL);
        else if (dilation == 2) tss_launch(dw_wgrad_kernel<T, 1, 2>, grid, threads, smem, st, (const T*)x, (const T*)dy, dw, N, Hi, Wi, Ho, Wo, C, PL);
        else tss_launch(dw_wgrad_kernel<T, 1, 4>, grid, threads, smem, st, (const T*)x, (const T*)dy, dw, N, Hi, Wi, Ho, Wo, C, PL);
        TSS_LAUNCH_CHECK("dwconv3x3_wgrad");
        return TSS_OK;
    });
}
