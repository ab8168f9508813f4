// Pyramid pooling (adaptive average pool, all bins in one launch) and bilinear resize with
// align_corners=True semantics identical to ATen (scale = (in-1)/(out-1) in fp32,
// src = scale*dst, i0 = (int)src, lambda = src - i0).  All HBM-streaming kernels.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxBins = 8;

struct Bins {
    int n;
    int b[kMaxBins];
    int off[kMaxBins + 1];   // prefix sums of b*b
};

__device__ __forceinline__ int win_start(int i, int in, int b) { return (i * in) / b; }
__device__ __forceinline__ int win_end(int i, int in, int b) { return ((i + 1) * in + b - 1) / b; }

// one CTA per (n, cell): 256 threads = CG channel groups x PL pixel lanes
template <typename T>
__global__ void __launch_bounds__(kThreads)
pool_fwd_kernel(const T* __restrict__ x, T* __restrict__ out, int N, int H, int W, int C, Bins bins, int PL) {
    pdl_wait();
    TSS_DYN_SMEM(float, s_acc);   // [C]
    const int cells = bins.off[bins.n];
    const int n = blockIdx.x / cells;
    int cell = blockIdx.x - n * cells;
    int bi = 0;
    while (bi + 1 < bins.n && cell >= bins.off[bi + 1]) ++bi;
    cell -= bins.off[bi];
    const int b = bins.b[bi];
    const int ci = cell / b, cj = cell - ci * b;
    const int h0 = win_start(ci, H, b), h1 = win_end(ci, H, b);
    const int w0 = win_start(cj, W, b), w1 = win_end(cj, W, b);
    const int CG = C >> 3;
    const int cg = threadIdx.x % CG, pl = threadIdx.x / CG;
    for (int i = threadIdx.x; i < C; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
    float acc[8];
    zero8(acc);
    const int ww = w1 - w0, npx = (h1 - h0) * ww;
    if (pl < PL) {
        for (int p = pl; p < npx; p += PL) {
            const int hh = h0 + p / ww, wx = w0 + p % ww;
            float v[8];
            load8(x + (((int64_t)n * H + hh) * W + wx) * C + cg * 8, v);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] += v[e];
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) atomicAdd(&s_acc[cg * 8 + e], acc[e]);
    }
    __syncthreads();
    const float inv = 1.f / (float)npx;
    for (int c = threadIdx.x; c < C; c += blockDim.x)
        out[((int64_t)N * bins.off[bi] + (int64_t)n * b * b + cell) * C + c] = from_f32<T>(s_acc[c] * inv);
}

// dx[n][h][w][c] (+)= sum over bins, over cells containing (h,w): dout[cell][c] / area(cell)
template <typename T>
__global__ void __launch_bounds__(kThreads)
pool_bwd_kernel(const T* __restrict__ dout, T* __restrict__ dx, int N, int H, int W, int C, Bins bins,
                int accumulate) {
    pdl_wait();
    const int CG = C >> 3;
    const int64_t total = (int64_t)N * H * W * CG;
    for (int64_t item = (int64_t)blockIdx.x * kThreads + threadIdx.x; item < total;
         item += (int64_t)gridDim.x * kThreads) {
        int64_t t = item / CG;
        const int c0 = (int)(item - t * CG) * 8;
        const int w = (int)(t % W); t /= W;
        const int h = (int)(t % H);
        const int n = (int)(t / H);
        float acc[8];
        if (accumulate) load8(dx + item * 8, acc);
        else zero8(acc);
        for (int bi = 0; bi < bins.n; ++bi) {
            const int b = bins.b[bi];
            for (int i = 0; i < b; ++i) {
                const int h0 = win_start(i, H, b), h1 = win_end(i, H, b);
                if (h < h0 || h >= h1) continue;
                for (int j = 0; j < b; ++j) {
                    const int w0 = win_start(j, W, b), w1 = win_end(j, W, b);
                    if (w < w0 || w >= w1) continue;
                    const float inv = 1.f / (float)((h1 - h0) * (w1 - w0));
                    float v[8];
                    load8(dout + ((int64_t)N * bins.off[bi] + (int64_t)n * b * b + i * b + j) * C + c0, v);
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[e] = fmaf(v[e], inv, acc[e]);
                }
            }
        }
        store8(dx + item * 8, acc);
    }
}

// ------------------------------------------------------------ bilinear, NHWC, C%8==0 ---
template <typename T>
__global__ void __launch_bounds__(kThreads)
bilinear_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int Hi, int Wi, int Ho, int Wo, int C,
                    int64_t ldx, int64_t ldy, float sh, float sw) {
    pdl_wait();
    const int CG = C >> 3;
    const int64_t total = (int64_t)N * Ho * Wo * CG;
    for (int64_t item = (int64_t)blockIdx.x * kThreads + threadIdx.x; item < total;
         item += (int64_t)gridDim.x * kThreads) {
        int64_t t = item / CG;
        const int c0 = (int)(item - t * CG) * 8;
        const int64_t opix = t;
        const int wo = (int)(t % Wo); t /= Wo;
        const int ho = (int)(t % Ho);
        const int n = (int)(t / Ho);
        int h0, h1, w0, w1;
        float lh, lw;
        ac_source(sh, ho, Hi, h0, h1, lh);
        ac_source(sw, wo, Wi, w0, w1, lw);
        const T* xn = x + (int64_t)n * Hi * Wi * ldx + c0;
        float p00[8], p01[8], p10[8], p11[8], o[8];
        load8(xn + ((int64_t)h0 * Wi + w0) * ldx, p00);
        load8(xn + ((int64_t)h0 * Wi + w1) * ldx, p01);
        load8(xn + ((int64_t)h1 * Wi + w0) * ldx, p10);
        load8(xn + ((int64_t)h1 * Wi + w1) * ldx, p11);
        const float a = 1.f - lh, b = 1.f - lw;
#pragma unroll
        for (int e = 0; e < 8; ++e)
            o[e] = a * (b * p00[e] + lw * p01[e]) + lh * (b * p10[e] + lw * p11[e]);   // ATen's association
        store8(y + opix * ldy + c0, o);
    }
}

// gather form of the transpose: each input pixel sums the output pixels that read it
template <typename T>
__global__ void __launch_bounds__(kThreads)
bilinear_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int N, int Hi, int Wi, int Ho, int Wo,
                    int C, int64_t lddy, int64_t lddx, float sh, float sw) {
    pdl_wait();
    const int CG = C >> 3;
    const int64_t total = (int64_t)N * Hi * Wi * CG;
    for (int64_t item = (int64_t)blockIdx.x * kThreads + threadIdx.x; item < total;
         item += (int64_t)gridDim.x * kThreads) {
        int64_t t = item / CG;
        const int c0 = (int)(item - t * CG) * 8;
        const int64_t ipix = t;
        const int wi = (int)(t % Wi); t /= Wi;
        const int hi = (int)(t % Hi);
        const int n = (int)(t / Hi);
        const int ho_a = first_candidate(sh, hi, Ho), ho_b = last_candidate(sh, hi, Ho);
        const int wo_a = first_candidate(sw, wi, Wo), wo_b = last_candidate(sw, wi, Wo);
        const T* dyn = dy + (int64_t)n * Ho * Wo * lddy + c0;
        float acc[8];
        zero8(acc);
        for (int ho = ho_a; ho <= ho_b; ++ho) {
            int h0, h1; float lh;
            ac_source(sh, ho, Hi, h0, h1, lh);
            float wh = 0.f;
            if (h0 == hi) wh += 1.f - lh;
            if (h1 == hi) wh += lh;
            if (wh == 0.f) continue;
            for (int wo = wo_a; wo <= wo_b; ++wo) {
                int w0, w1; float lw;
                ac_source(sw, wo, Wi, w0, w1, lw);
                float ww = 0.f;
                if (w0 == wi) ww += 1.f - lw;
                if (w1 == wi) ww += lw;
                if (ww == 0.f) continue;
                float v[8];
                load8(dyn + ((int64_t)ho * Wo + wo) * lddy, v);
                const float k = wh * ww;
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] = fmaf(v[e], k, acc[e]);
            }
        }
        store8(dx + ipix * lddx + c0, acc);
    }
}

// transpose of a 1-D align_corners resize along one axis: the tensor is [outer][S][inner][C]
// (C pitched); every thread sums the <= 2*ceil(So/Si)+3 source rows that read its row.
template <typename Tin, typename Tout>
__global__ void __launch_bounds__(kThreads)
resize_bwd_axis_kernel(const Tin* __restrict__ dy, Tout* __restrict__ dx, int64_t outer, int So, int Si,
                       int inner, int C, int64_t ld_in, int64_t ld_out, float scale) {
    pdl_wait();
    const int CG = C >> 3;
    const int64_t total = outer * Si * inner * CG;
    for (int64_t item = (int64_t)blockIdx.x * kThreads + threadIdx.x; item < total;
         item += (int64_t)gridDim.x * kThreads) {
        int64_t t = item / CG;
        const int c0 = (int)(item - t * CG) * 8;
        const int p = (int)(t % inner); t /= inner;
        const int si = (int)(t % Si);
        const int64_t o = t / Si;
        const int a = first_candidate(scale, si, So), b = last_candidate(scale, si, So);
        float acc[8];
        zero8(acc);
        for (int so = a; so <= b; ++so) {
            int i0, i1; float lam;
            ac_source(scale, so, Si, i0, i1, lam);
            float w = 0.f;
            if (i0 == si) w += 1.f - lam;
            if (i1 == si) w += lam;
            if (w == 0.f) continue;
            float v[8];
            load8(dy + ((o * So + so) * inner + p) * ld_in + c0, v);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = fmaf(v[e], w, acc[e]);
        }
        store8(dx + ((o * Si + si) * inner + p) * ld_out + c0, acc);
    }
}

// ------------------------------------------------------------ final logits x8 ----------
// out NCHW y[n][c][ho][wo].  One CTA per (n, ho): the two source rows are blended
// vertically into shared memory (fp32 [Wi][C]) with coalesced reads, then every thread
// produces 8 consecutive wo of one class plane (one 128-bit store in bf16).
template <typename T>
__global__ void __launch_bounds__(kThreads)
upsample_logits_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int Hi, int Wi, int Ho, int Wo, int C,
                           int64_t ldx, float sh, float sw) {
    pdl_wait();
    TSS_DYN_SMEM(float, s_row);   // [Wi][C]
    const int n = blockIdx.x / Ho, ho = blockIdx.x - n * Ho;
    int h0, h1; float lh;
    ac_source(sh, ho, Hi, h0, h1, lh);
    const T* r0 = x + ((int64_t)n * Hi + h0) * Wi * ldx;
    const T* r1 = x + ((int64_t)n * Hi + h1) * Wi * ldx;
    for (int i = threadIdx.x; i < Wi * C; i += kThreads) {
        const int w = i / C, c = i - w * C;
        s_row[i] = (1.f - lh) * to_f32(r0[w * ldx + c]) + lh * to_f32(r1[w * ldx + c]);
    }
    __syncthreads();
    // a thread keeps ONE group of 8 output columns and walks the classes: the column geometry (source
    // offsets and weights) is computed once into registers and reused for every class plane
    const int groups = Wo >> 3;
    const int lanes = kThreads / groups > 0 ? kThreads / groups : 1;     // threads sharing one column group
    const int lane = threadIdx.x / groups;                               // (0 when there are more groups than threads)
    if (lane < lanes) {
        for (int g = threadIdx.x % groups; g < groups; g += kThreads) {
            int o0[8], o1[8];
            float lw[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                int w0, w1;
                ac_source(sw, g * 8 + e, Wi, w0, w1, lw[e]);
                o0[e] = w0 * C;
                o1[e] = w1 * C;
            }
            for (int c = lane; c < C; c += lanes) {
                float o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = (1.f - lw[e]) * s_row[o0[e] + c] + lw[e] * s_row[o1[e] + c];
                store8(y + (((int64_t)n * C + c) * Ho + ho) * Wo + g * 8, o);
            }
        }
    }
}

// Backward: one CTA per (n, c, source-row interval hi): the output rows whose h0 == hi are
// read exactly once (coalesced 128-bit), blended vertically into two fp32 rows in shared
// memory (weights 1-lh -> row hi, lh -> row hi+1), reduced horizontally by gathering, and
// added to the fp32 gradient with one atomic per element.
template <typename T>
__global__ void __launch_bounds__(kThreads)
upsample_logits_bwd_kernel(const T* __restrict__ dy, float* __restrict__ dx, int Hi, int Wi, int Ho, int Wo,
                           int C, int64_t lddx, float sh, float sw) {
    pdl_wait();
    TSS_DYN_SMEM(float, s_t);     // [2][Wo]
    const int hi = blockIdx.x % Hi;
    const int c = (blockIdx.x / Hi) % C;
    const int n = blockIdx.x / (Hi * C);
    const int ho_a = first_candidate(sh, hi, Ho), ho_b = last_candidate(sh, hi, Ho);
    const T* plane = dy + ((int64_t)n * C + c) * Ho * Wo;
    const bool has_next = hi < Hi - 1;
    const int groups = Wo >> 3;
    for (int g = threadIdx.x; g < groups; g += kThreads) {
        float t0[8], t1[8];
        zero8(t0); zero8(t1);
        for (int ho = ho_a; ho <= ho_b; ++ho) {
            int h0, h1; float lh;
            ac_source(sh, ho, Hi, h0, h1, lh);
            if (h0 != hi) continue;
            float v[8];
            load8(plane + (int64_t)ho * Wo + g * 8, v);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                if (h1 == hi) t0[e] += v[e];                      // clamped last row: both weights land on hi
                else { t0[e] = fmaf(v[e], 1.f - lh, t0[e]); t1[e] = fmaf(v[e], lh, t1[e]); }
            }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) { s_t[g * 8 + e] = t0[e]; s_t[Wo + g * 8 + e] = t1[e]; }
    }
    __syncthreads();
    for (int wi = threadIdx.x; wi < Wi; wi += kThreads) {
        const int wo_a = first_candidate(sw, wi, Wo), wo_b = last_candidate(sw, wi, Wo);
        float a0 = 0.f, a1 = 0.f;
        for (int wo = wo_a; wo <= wo_b; ++wo) {
            int w0, w1; float lw;
            ac_source(sw, wo, Wi, w0, w1, lw);
            float ww = 0.f;
            if (w0 == wi) ww += 1.f - lw;
            if (w1 == wi) ww += lw;
            if (ww == 0.f) continue;
            a0 = fmaf(s_t[wo], ww, a0);
            a1 = fmaf(s_t[Wo + wo], ww, a1);
        }
        atomicAdd(dx + (((int64_t)n * Hi + hi) * Wi + wi) * lddx + c, a0);
        if (has_next) atomicAdd(dx + (((int64_t)n * Hi + hi + 1) * Wi + wi) * lddx + c, a1);
    }
}

__global__ void __launch_bounds__(kThreads)
bilinear_nchw_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int NC, int Hi, int Wi, int Ho,
                         int Wo, float sh, float sw) {
    pdl_wait();
    const int64_t total = (int64_t)NC * Ho * Wo;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
        int64_t t = i;
        const int wo = (int)(t % Wo); t /= Wo;
        const int ho = (int)(t % Ho);
        const int64_t p = t / Ho;
        int h0, h1, w0, w1; float lh, lw;
        ac_source(sh, ho, Hi, h0, h1, lh);
        ac_source(sw, wo, Wi, w0, w1, lw);
        const float* xp = x + p * Hi * Wi;
        const float a = 1.f - lh, b = 1.f - lw;
        y[i] = a * (b * __ldg(xp + (int64_t)h0 * Wi + w0) + lw * __ldg(xp + (int64_t)h0 * Wi + w1)) +
               lh * (b * __ldg(xp + (int64_t)h1 * Wi + w0) + lw * __ldg(xp + (int64_t)h1 * Wi + w1));
    }
}

int make_bins(const char* name, const int* bins, int nbins, Bins& out) {
    TSS_REQUIRE(nbins > 0 && nbins <= kMaxBins, "%s: nbins=%d (max %d)", name, nbins, kMaxBins);
    out.n = nbins;
    out.off[0] = 0;
    for (int i = 0; i < nbins; ++i) {
        TSS_REQUIRE(bins[i] > 0 && bins[i] <= 64, "%s: bin %d out of range", name, bins[i]);
        out.b[i] = bins[i];
        out.off[i + 1] = out.off[i] + bins[i] * bins[i];
    }
    return TSS_OK;
}

inline int stream_grid(int64_t items, int per_sm = 8) {
    int64_t want = ceil_div64(items, kThreads);
    int64_t cap = (int64_t)tss_num_sms() * per_sm;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

}  // namespace

extern "C" int tss_adaptive_pool_fwd(const void* x, void* out, int N, int H, int W, int C, const int* bins,
                                     int nbins, int dtype, void* stream) {
    Bins b;
    if (int e = make_bins("adaptive_pool_fwd", bins, nbins, b)) return e;
    TSS_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && C / 8 <= kThreads, "adaptive_pool_fwd: bad shape");
    const int CG = C / 8, PL = kThreads / CG;
    TSS_DISPATCH_DTYPE(dtype, "adaptive_pool_fwd", {
        tss_launch(pool_fwd_kernel<T>, N * b.off[b.n], PL * CG, (size_t)C * sizeof(float), (cudaStream_t)stream, 
            (const T*)x, (T*)out, N, H, W, C, b, PL);
        TSS_LAUNCH_CHECK("adaptive_pool_fwd");
        return TSS_OK;
    });
}

extern "C" int tss_adaptive_pool_bwd(const void* dout, void* dx, int N, int H, int W, int C, const int* bins,
                                     int nbins, int accumulate, int dtype, void* stream) {
    Bins b;
    if (int e = make_bins("adaptive_pool_bwd", bins, nbins, b)) return e;
    TSS_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "adaptive_pool_bwd: bad shape");
    TSS_DISPATCH_DTYPE(dtype, "adaptive_pool_bwd", {
        tss_launch(pool_bwd_kernel<T>, stream_grid((int64_t)N * H * W * (C / 8)), kThreads, 0, (cudaStream_t)stream, 
            (const T*)dout, (T*)dx, N, H, W, C, b, accumulate);
        TSS_LAUNCH_CHECK("adaptive_pool_bwd");
        return TSS_OK;
    });
}

extern "C" int tss_bilinear_fwd(const void* x, void* y, int N, int Hi, int Wi, int Ho, int Wo, int C,
                                int64_t ldx, int64_t ldy, int dtype, void* stream) {
    TSS_REQUIRE(N > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, "bilinear_fwd: empty tensor");
    TSS_REQUIRE(C > 0 && C % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0 && ldx >= C && ldy >= C, "bilinear_fwd: C=%d ldx=%lld ldy=%lld", C, (long long)ldx, (long long)ldy);
    TSS_DISPATCH_DTYPE(dtype, "bilinear_fwd", {
        tss_launch(bilinear_fwd_kernel<T>, stream_grid((int64_t)N * Ho * Wo * (C / 8)), kThreads, 0, (cudaStream_t)stream, 
            (const T*)x, (T*)y, N, Hi, Wi, Ho, Wo, C, ldx, ldy, ac_scale(Hi, Ho), ac_scale(Wi, Wo));
        TSS_LAUNCH_CHECK("bilinear_fwd");
        return TSS_OK;
    });
}

extern "C" int tss_bilinear_bwd(const void* dy, void* dx, float* workspace, int N, int Hi, int Wi, int Ho,
                                int Wo, int C, int64_t lddy, int64_t lddx, int dtype, void* stream) {
    TSS_REQUIRE(N > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, "bilinear_bwd: empty tensor");
    TSS_REQUIRE(C > 0 && C % 8 == 0 && lddx % 8 == 0 && lddy % 8 == 0 && lddx >= C && lddy >= C, "bilinear_bwd: C=%d", C);
    cudaStream_t st = (cudaStream_t)stream;
    TSS_DISPATCH_DTYPE(dtype, "bilinear_bwd", {
        if (workspace == nullptr) {      // single gather pass (fine for small maps)
            tss_launch(bilinear_bwd_kernel<T>, stream_grid((int64_t)N * Hi * Wi * (C / 8)), kThreads, 0, st, 
                (const T*)dy, (T*)dx, N, Hi, Wi, Ho, Wo, C, lddy, lddx, ac_scale(Hi, Ho), ac_scale(Wi, Wo));
            TSS_LAUNCH_CHECK("bilinear_bwd");
            return TSS_OK;
        }
        // separable: rows first (Ho -> Hi) into the fp32 workspace [N][Hi][Wo][C], then columns
        tss_launch(resize_bwd_axis_kernel<T, float>, stream_grid((int64_t)N * Hi * Wo * (C / 8)), kThreads, 0, st, 
            (const T*)dy, workspace, N, Ho, Hi, Wo, C, lddy, C, ac_scale(Hi, Ho));
        TSS_LAUNCH_CHECK("bilinear_bwd(rows)");
        tss_launch(resize_bwd_axis_kernel<float, T>, stream_grid((int64_t)N * Hi * Wi * (C / 8)), kThreads, 0, st, 
            workspace, (T*)dx, (int64_t)N * Hi, Wo, Wi, 1, C, C, lddx, ac_scale(Wi, Wo));
        TSS_LAUNCH_CHECK("bilinear_bwd(cols)");
        return TSS_OK;
    });
}

extern "C" int tss_upsample_logits_fwd(const void* x, void* y, int N, int Hi, int Wi, int Ho, int Wo, int C,
                                       int64_t ldx, int dtype, void* stream) {
    TSS_REQUIRE(N > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, "upsample_logits_fwd: empty tensor");
    TSS_REQUIRE(C > 0 && C <= 64 && ldx >= C, "upsample_logits_fwd: C=%d ldx=%lld", C, (long long)ldx);
    TSS_REQUIRE(Wo % 8 == 0, "upsample_logits_fwd: Wo=%d must be a multiple of 8", Wo);
    const size_t smem = (size_t)Wi * C * sizeof(float);
    TSS_REQUIRE(smem <= 48 * 1024, "upsample_logits_fwd: Wi*C=%d too large", Wi * C);
    TSS_DISPATCH_DTYPE(dtype, "upsample_logits_fwd", {
        tss_launch(upsample_logits_fwd_kernel<T>, N * Ho, kThreads, smem, (cudaStream_t)stream, 
            (const T*)x, (T*)y, Hi, Wi, Ho, Wo, C, ldx, ac_scale(Hi, Ho), ac_scale(Wi, Wo));
        TSS_LAUNCH_CHECK("upsample_logits_fwd");
        return TSS_OK;
    });
}

extern "C" int tss_upsample_logits_bwd(const void* dy, float* dx32, int N, int Hi, int Wi, int Ho, int Wo,
                                       int C, int64_t lddx, int dtype, void* stream) {
    TSS_REQUIRE(N > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, "upsample_logits_bwd: empty tensor");
    TSS_REQUIRE(C > 0 && C <= 64 && lddx >= C, "upsample_logits_bwd: C=%d lddx=%lld", C, (long long)lddx);
    TSS_REQUIRE(Wo % 8 == 0, "upsample_logits_bwd: Wo=%d must be a multiple of 8", Wo);
    const size_t smem = (size_t)2 * Wo * sizeof(float);
    TSS_REQUIRE(smem <= 48 * 1024, "upsample_logits_bwd: Wo=%d too large", Wo);
    TSS_DISPATCH_DTYPE(dtype, "upsample_logits_bwd", {
        tss_launch(upsample_logits_bwd_kernel<T>, N * C * Hi, kThreads, smem, (cudaStream_t)stream, 
            (const T*)dy, dx32, Hi, Wi, Ho, Wo, C, lddx, ac_scale(Hi, Ho), ac_scale(Wi, Wo));
        TSS_LAUNCH_CHECK("upsample_logits_bwd");
        return TSS_OK;
    });
}

extern "C" int tss_bilinear_nchw_f32(const float* x, float* y, int NC, int Hi, int Wi, int Ho, int Wo,
                                     void* stream) {
    TSS_REQUIRE(NC > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, "bilinear_nchw_f32: empty tensor");
    tss_launch(bilinear_nchw_f32_kernel, stream_grid((int64_t)NC * Ho * Wo), kThreads, 0, (cudaStream_t)stream, 
        x, y, NC, Hi, Wi, Ho, Wo, ac_scale(Hi, Ho), ac_scale(Wi, Wo));
    TSS_LAUNCH_CHECK("bilinear_nchw_f32");
    return TSS_OK;
}
