// Pyramid pooling branches as grouped launches (training).  reference: PyramidPoolingModule, models/fastscnn.py:
// 101-123 -- 4 x [AdaptiveAvgPool2d(b) -> Conv2d(128, 32, 1) -> BatchNorm2d -> ReLU] -> bilinear up-sampling
// (align_corners=True) -> cat([x, *branches]).  The branches hold N*(1+4+9+36) pixels: no bytes, no flops, but
// taken layer by layer they cost 16 launches forward and ~26 backward on the critical chain.  Here:
//
//   forward   pool (resample.cu, all bins in one pass)
//             ppm_branches_fwd   grid (branch, 8-channel group): 1x1 conv + batch statistics + finalize (running
//                                statistics included) + affine + ReLU.  A CTA owns ALL pixels of its channels, so
//                                the statistics never leave the CTA: no atomics, no second kernel.
//             ppm_concat_fwd     x and the four up-sampled branches into the concat buffer, one pass.
//   backward  ppm_concat_bwd     transpose of the up-samplings for all branches: CTA per (image, branch),
//                                separable (rows, then columns) through shared memory.
//             ppm_bn_bwd         grid (branch, 8-channel group): ReLU mask from the raw conv output, both
//                                reductions, dy, dgamma, dbeta.
//             ppm_dgrad_wgrad    dgrad over row chunks and, in two more CTAs per branch, the weight gradient.
//             (+ the existing pool backward, accumulating onto the copy of the pass-through gradient.)
//
// Parameters of the branches live wherever the optimizer put them: the kernels read a small device table of
// addresses (kPpmTableCols int64 per branch, see tss_b200.h).
#include <stdint.h>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxBins = 8;
constexpr int kPpmTableCols = 10;

struct PpmBins {
    int n;
    int b[kMaxBins];
    int row0[kMaxBins + 1];      // first row of the branch in the pooled buffer: N * sum of b*b of the previous bins
};

enum { T_W = 0, T_GAMMA, T_BETA, T_RMEAN, T_RVAR, T_NBT, T_DW, T_DGAMMA, T_DBETA };

template <typename P> __device__ __forceinline__ P* table_ptr(const int64_t* table, int branch, int col) {
    return reinterpret_cast<P*>((uintptr_t)table[branch * kPpmTableCols + col]);
}

template <typename T> __device__ __forceinline__ void load8_plain(const T* p, float (&v)[8]) { load8_smem(p, v); }

__device__ __forceinline__ float scalar_f32(float v) { return v; }
__device__ __forceinline__ float scalar_f32(bf16 v) { return __bfloat162float(v); }

// block-wide sums of 16 per-thread values -> s_out[16] (double); every thread calls
__device__ __forceinline__ void block_sum16(const float (&v)[16], float* s_part /*[8 warps][16]*/, double* s_out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        float x = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) s_part[warp * 16 + i] = x;
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        double a = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) a += (double)s_part[w * 16 + threadIdx.x];
        s_out[threadIdx.x] = a;
    }
    __syncthreads();
}

// ------------------------------------------------------------------ forward: conv + BN(train) + ReLU
template <typename T>
__global__ void __launch_bounds__(kThreads)
ppm_branches_fwd_kernel(const T* __restrict__ pool, const int64_t* __restrict__ table, T* y, T* __restrict__ z,
                        float* __restrict__ mean_out, float* __restrict__ rstd_out, int C, int Cb, PpmBins bins,
                        float momentum, float eps, int eval_mode) {
    TSS_DYN_SMEM(float, s_w);                            // [8][C]
    __shared__ float s_part[(kThreads / 32) * 16];
    __shared__ double s_sum[16];
    __shared__ float s_aff[16];                          // scale[8], shift[8]
    pdl_wait();
    const int br = blockIdx.x, c0 = blockIdx.y * 8;
    const int row0 = bins.row0[br], M = bins.row0[br + 1] - row0;
    const float* w = table_ptr<const float>(table, br, T_W);
    for (int i = threadIdx.x; i < 8 * C; i += kThreads) s_w[i] = w[(size_t)c0 * C + i];
    __syncthreads();
    float st[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) st[i] = 0.f;
    for (int r = threadIdx.x; r < M; r += kThreads) {
        const T* xr = pool + (size_t)(row0 + r) * C;
        float acc[8];
        zero8(acc);
        for (int k = 0; k < C; k += 8) {
            float xv[8];
            load8(xr + k, xv);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 wa = *reinterpret_cast<const float4*>(s_w + c * C + k);
                const float4 wb = *reinterpret_cast<const float4*>(s_w + c * C + k + 4);
                acc[c] = fmaf(xv[0], wa.x, acc[c]); acc[c] = fmaf(xv[1], wa.y, acc[c]);
                acc[c] = fmaf(xv[2], wa.z, acc[c]); acc[c] = fmaf(xv[3], wa.w, acc[c]);
                acc[c] = fmaf(xv[4], wb.x, acc[c]); acc[c] = fmaf(xv[5], wb.y, acc[c]);
                acc[c] = fmaf(xv[6], wb.z, acc[c]); acc[c] = fmaf(xv[7], wb.w, acc[c]);
            }
        }
        if (eval_mode) {          // folded BatchNorm: table columns 1, 2 hold scale / shift (running statistics)
            const float* sc = table_ptr<const float>(table, br, T_GAMMA);
            const float* sh = table_ptr<const float>(table, br, T_BETA);
            float o[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) o[c] = fmaxf(fmaf(acc[c], __ldg(sc + c0 + c), __ldg(sh + c0 + c)), 0.f);
            store8(z + (size_t)(row0 + r) * Cb + c0, o);
            continue;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            st[c] += acc[c];
            st[8 + c] = fmaf(acc[c], acc[c], st[8 + c]);
        }
        store8(y + (size_t)(row0 + r) * Cb + c0, acc);
    }
    if (eval_mode) return;                                  // uniform over the grid
    block_sum16(st, s_part, s_sum);
    if (threadIdx.x < 8) {
        const int c = c0 + threadIdx.x;
        const double inv = 1.0 / (double)M;
        const double mean = s_sum[threadIdx.x] * inv;
        double var = s_sum[8 + threadIdx.x] * inv - mean * mean;
        if (var < 0.0) var = 0.0;
        const float rstd = (float)(1.0 / sqrt(var + (double)eps));
        const float* gamma = table_ptr<const float>(table, br, T_GAMMA);
        const float* beta = table_ptr<const float>(table, br, T_BETA);
        const float sc = (gamma != nullptr ? gamma[c] : 1.f) * rstd;
        s_aff[threadIdx.x] = sc;
        s_aff[8 + threadIdx.x] = (beta != nullptr ? beta[c] : 0.f) - (float)mean * sc;
        mean_out[br * Cb + c] = (float)mean;
        rstd_out[br * Cb + c] = rstd;
        float* rm = table_ptr<float>(table, br, T_RMEAN);
        float* rv = table_ptr<float>(table, br, T_RVAR);
        if (rm != nullptr && rv != nullptr) {
            const double unbias = M > 1 ? (double)M / (double)(M - 1) : 1.0;
            rm[c] = (1.f - momentum) * rm[c] + momentum * (float)mean;
            rv[c] = (1.f - momentum) * rv[c] + momentum * (float)(var * unbias);
        }
        int64_t* nbt = table_ptr<int64_t>(table, br, T_NBT);
        if (nbt != nullptr && c == 0) *nbt += 1;
    }
    __syncthreads();
    for (int r = threadIdx.x; r < M; r += kThreads) {      // each thread re-reads the rows it wrote itself
        float v[8], o[8];
        load8_plain(y + (size_t)(row0 + r) * Cb + c0, v);
#pragma unroll
        for (int c = 0; c < 8; ++c) o[c] = fmaxf(fmaf(v[c], s_aff[c], s_aff[8 + c]), 0.f);
        store8(z + (size_t)(row0 + r) * Cb + c0, o);
    }
}

// ------------------------------------------------------------------ forward: cat([x, up(z_b)...])
template <typename T>
__global__ void __launch_bounds__(kThreads)
ppm_concat_fwd_kernel(const T* __restrict__ x, const T* __restrict__ z, T* __restrict__ cat, int N, int H, int W,
                      int C, int Cb, PpmBins bins) {
    pdl_wait();
    const int ldc = C + bins.n * Cb;
    const int CG = ldc >> 3;
    const int64_t total = (int64_t)N * H * W * CG;
    for (int64_t item = (int64_t)blockIdx.x * kThreads + threadIdx.x; item < total; item += (int64_t)gridDim.x * kThreads) {
        const int64_t pix = item / CG;
        const int c = (int)(item - pix * CG) * 8;
        float o[8];
        if (c < C) {
            load8(x + pix * C + c, o);
        } else {
            const int br = (c - C) / Cb, cb = (c - C) - br * Cb;
            const int b = bins.b[br];
            const int wo = (int)(pix % W);
            const int ho = (int)((pix / W) % H);
            const int n = (int)(pix / ((int64_t)W * H));
            int h0, h1, w0, w1;
            float lh, lw;
            ac_source(ac_scale(b, H), ho, b, h0, h1, lh);
            ac_source(ac_scale(b, W), wo, b, w0, w1, lw);
            const T* zn = z + ((size_t)bins.row0[br] + (size_t)n * b * b) * Cb + cb;
            float p00[8], p01[8], p10[8], p11[8];
            load8(zn + (size_t)(h0 * b + w0) * Cb, p00);
            load8(zn + (size_t)(h0 * b + w1) * Cb, p01);
            load8(zn + (size_t)(h1 * b + w0) * Cb, p10);
            load8(zn + (size_t)(h1 * b + w1) * Cb, p11);
            const float a = 1.f - lh, bb = 1.f - lw;
#pragma unroll
            for (int e = 0; e < 8; ++e)
                o[e] = a * (bb * p00[e] + lw * p01[e]) + lh * (bb * p10[e] + lw * p11[e]);   // ATen's association
        }
        store8(cat + pix * ldc + c, o);
    }
}

// ------------------------------------------------------------------ backward: transpose of the up-samplings
// CTA per (image, branch).  Rows first: tmp[i][w][c] = sum_h wh(i <- h) * dcat[n][h][w][C + br*Cb + c], then
// columns: dz[i][j][c] = sum_w ww(j <- w) * tmp[i][w][c].
template <typename T>
__global__ void __launch_bounds__(kThreads)
ppm_concat_bwd_kernel(const T* __restrict__ dcat, T* __restrict__ dz, int N, int H, int W, int C, int Cb,
                      int64_t lddcat, PpmBins bins) {
    TSS_DYN_SMEM(float, s_tmp);                          // [b][W][Cb]
    pdl_wait();
    const int n = blockIdx.x, br = blockIdx.y;
    const int b = bins.b[br];
    const int CG = Cb >> 3;
    const float sh = ac_scale(b, H), sw = ac_scale(b, W);
    const T* g = dcat + (size_t)n * H * W * lddcat + C + br * Cb;
    for (int item = threadIdx.x; item < W * CG; item += kThreads) {
        const int w = item / CG, c = (item - w * CG) * 8;
        for (int i = 0; i < b; ++i) {
            float acc[8];
            zero8(acc);
            const int lo = first_candidate(sh, i, H), hi = last_candidate(sh, i, H);
            for (int h = lo; h <= hi; ++h) {
                int i0, i1;
                float lam;
                ac_source(sh, h, b, i0, i1, lam);
                float wt = 0.f;
                if (i0 == i) wt += 1.f - lam;
                if (i1 == i) wt += lam;
                if (wt != 0.f) {
                    float v[8];
                    load8(g + ((size_t)h * W + w) * lddcat + c, v);
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[e] = fmaf(wt, v[e], acc[e]);
                }
            }
            float* t = s_tmp + ((size_t)i * W + w) * Cb + c;
#pragma unroll
            for (int e = 0; e < 8; ++e) t[e] = acc[e];
        }
    }
    __syncthreads();
    T* out = dz + ((size_t)bins.row0[br] + (size_t)n * b * b) * Cb;
    for (int item = threadIdx.x; item < b * b * CG; item += kThreads) {
        const int cell = item / CG, c = (item - cell * CG) * 8;
        const int i = cell / b, j = cell - i * b;
        float acc[8];
        zero8(acc);
        const int lo = first_candidate(sw, j, W), hi = last_candidate(sw, j, W);
        for (int w = lo; w <= hi; ++w) {
            int j0, j1;
            float lam;
            ac_source(sw, w, b, j0, j1, lam);
            float wt = 0.f;
            if (j0 == j) wt += 1.f - lam;
            if (j1 == j) wt += lam;
            if (wt != 0.f) {
                const float* t = s_tmp + ((size_t)i * W + w) * Cb + c;
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] = fmaf(wt, t[e], acc[e]);
            }
        }
        store8(out + (size_t)cell * Cb + c, acc);
    }
}

// ------------------------------------------------------------------ backward: BatchNorm + ReLU of the branches
template <typename T>
__global__ void __launch_bounds__(kThreads)
ppm_bn_bwd_kernel(const T* __restrict__ dz, const T* __restrict__ y, const int64_t* __restrict__ table,
                  const float* __restrict__ mean, const float* __restrict__ rstd, T* __restrict__ dy, int Cb,
                  PpmBins bins) {
    __shared__ float s_part[(kThreads / 32) * 16];
    __shared__ double s_sum[16];
    pdl_wait();
    const int br = blockIdx.x, c0 = blockIdx.y * 8;
    const int row0 = bins.row0[br], M = bins.row0[br + 1] - row0;
    const float* gamma = table_ptr<const float>(table, br, T_GAMMA);
    const float* beta = table_ptr<const float>(table, br, T_BETA);
    float mu[8], rs[8], sc[8], shf[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        mu[c] = __ldg(mean + br * Cb + c0 + c);
        rs[c] = __ldg(rstd + br * Cb + c0 + c);
        sc[c] = (gamma != nullptr ? gamma[c0 + c] : 1.f) * rs[c];
        shf[c] = (beta != nullptr ? beta[c0 + c] : 0.f) - mu[c] * sc[c];
    }
    float st[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) st[i] = 0.f;
    for (int r = threadIdx.x; r < M; r += kThreads) {
        float g[8], yy[8];
        load8(dz + (size_t)(row0 + r) * Cb + c0, g);
        load8(y + (size_t)(row0 + r) * Cb + c0, yy);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (!(fmaf(yy[c], sc[c], shf[c]) > 0.f)) g[c] = 0.f;
            st[c] += g[c];
            st[8 + c] = fmaf(g[c], (yy[c] - mu[c]) * rs[c], st[8 + c]);
        }
    }
    block_sum16(st, s_part, s_sum);
    const float inv = 1.f / (float)M;
    float k1[8], k2[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        k1[c] = (float)s_sum[c] * inv;
        k2[c] = (float)s_sum[8 + c] * inv;
    }
    if (threadIdx.x < 8) {
        float* dgamma = table_ptr<float>(table, br, T_DGAMMA);
        float* dbeta = table_ptr<float>(table, br, T_DBETA);
        if (dbeta != nullptr) dbeta[c0 + threadIdx.x] += (float)s_sum[threadIdx.x];
        if (dgamma != nullptr) dgamma[c0 + threadIdx.x] += (float)s_sum[8 + threadIdx.x];
    }
    for (int r = threadIdx.x; r < M; r += kThreads) {
        float g[8], yy[8], o[8];
        load8(dz + (size_t)(row0 + r) * Cb + c0, g);
        load8(y + (size_t)(row0 + r) * Cb + c0, yy);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (!(fmaf(yy[c], sc[c], shf[c]) > 0.f)) g[c] = 0.f;
            o[c] = sc[c] * (g[c] - k1[c] - (yy[c] - mu[c]) * rs[c] * k2[c]);
        }
        store8(dy + (size_t)(row0 + r) * Cb + c0, o);
    }
}

// ------------------------------------------------------------------ backward: dgrad and wgrad of the 1x1 convs
// grid (chunks + 2, branch): CTAs 0..chunks-1 make dpool for 16 rows each (thread = row x 8 input channels);
// the last two CTAs make dw[c][k] += sum_r dy[r][c] * pool[r][k] for 16 output channels each.
template <typename T>
__global__ void __launch_bounds__(kThreads)
ppm_dgrad_wgrad_kernel(const T* __restrict__ dy, const T* __restrict__ pool, const int64_t* __restrict__ table,
                       T* __restrict__ dpool, int C, int Cb, PpmBins bins, int chunks) {
    TSS_DYN_SMEM(float, s_w);                            // dgrad: [Cb][C]
    pdl_wait();
    const int br = blockIdx.y;
    const int row0 = bins.row0[br], M = bins.row0[br + 1] - row0;
    const int KG = C >> 3;
    if ((int)blockIdx.x < chunks) {
        const int rows_per = kThreads / KG;              // 16 for C = 128
        const int r_base = blockIdx.x * rows_per;
        if (r_base >= M) return;
        const float* w = table_ptr<const float>(table, br, T_W);
        for (int i = threadIdx.x; i < Cb * C; i += kThreads) s_w[i] = w[i];
        __syncthreads();
        const int r = r_base + threadIdx.x / KG, k = (threadIdx.x % KG) * 8;
        if (threadIdx.x / KG >= rows_per || r >= M) return;
        float acc[8];
        zero8(acc);
        const T* g = dy + (size_t)(row0 + r) * Cb;
        for (int c = 0; c < Cb; c += 8) {
            float gv[8];
            load8(g + c, gv);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float4 wa = *reinterpret_cast<const float4*>(s_w + (c + e) * C + k);
                const float4 wb = *reinterpret_cast<const float4*>(s_w + (c + e) * C + k + 4);
                acc[0] = fmaf(gv[e], wa.x, acc[0]); acc[1] = fmaf(gv[e], wa.y, acc[1]);
                acc[2] = fmaf(gv[e], wa.z, acc[2]); acc[3] = fmaf(gv[e], wa.w, acc[3]);
                acc[4] = fmaf(gv[e], wb.x, acc[4]); acc[5] = fmaf(gv[e], wb.y, acc[5]);
                acc[6] = fmaf(gv[e], wb.z, acc[6]); acc[7] = fmaf(gv[e], wb.w, acc[7]);
            }
        }
        store8(dpool + (size_t)(row0 + r) * C + k, acc);
    } else {
        float* dw = table_ptr<float>(table, br, T_DW);
        if (dw == nullptr) return;
        const int half = blockIdx.x - chunks;            // 0 or 1: which half of the output channels
        const int c_per = (Cb + 1) / 2;
        for (int item = threadIdx.x; item < c_per * KG; item += kThreads) {
            const int c = half * c_per + item / KG, k = (item % KG) * 8;
            if (c >= Cb) continue;
            float acc[8];
            zero8(acc);
            for (int r = 0; r < M; ++r) {
                const float gv = scalar_f32(dy[(size_t)(row0 + r) * Cb + c]);
                float xv[8];
                load8(pool + (size_t)(row0 + r) * C + k, xv);
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] = fmaf(gv, xv[e], acc[e]);
            }
            float* d = dw + (size_t)c * C + k;
#pragma unroll
            for (int e = 0; e < 8; ++e) d[e] += acc[e];
        }
    }
}

int make_bins(PpmBins& pb, const int* bins, int nbins, int N, const char* name) {
    TSS_REQUIRE(bins != nullptr && nbins > 0 && nbins <= kMaxBins, "%s: nbins=%d", name, nbins);
    pb.n = nbins;
    pb.row0[0] = 0;
    for (int i = 0; i < nbins; ++i) {
        TSS_REQUIRE(bins[i] > 0 && bins[i] <= 64, "%s: bin %d", name, bins[i]);
        pb.b[i] = bins[i];
        pb.row0[i + 1] = pb.row0[i] + N * bins[i] * bins[i];
    }
    return TSS_OK;
}

}  // namespace

extern "C" int tss_ppm_branches_fwd(const void* pool, const int64_t* table, void* y, void* z, float* mean,
                                    float* rstd, int N, int C, int Cb, const int* bins, int nbins, float momentum,
                                    float eps, int dtype, void* stream) {
    PpmBins pb;
    if (int e = make_bins(pb, bins, nbins, N, "ppm_branches_fwd")) return e;
    TSS_REQUIRE(N > 0 && C > 0 && C % 8 == 0 && Cb > 0 && Cb % 8 == 0, "ppm_branches_fwd: N=%d C=%d Cb=%d", N, C, Cb);
    for (int i = 0; i < nbins; ++i)
        TSS_REQUIRE(N * bins[i] * bins[i] > 1, "ppm_branches_fwd: Expected more than 1 value per channel when training (bin %d, batch %d)", bins[i], N);
    const size_t smem = (size_t)8 * C * sizeof(float);
    TSS_REQUIRE(smem <= 40 * 1024, "ppm_branches_fwd: C=%d too large", C);
    TSS_DISPATCH_DTYPE(dtype, "ppm_branches_fwd", {
        tss_launch(ppm_branches_fwd_kernel<T>, dim3(nbins, Cb / 8), kThreads, smem, (cudaStream_t)stream, (const T*)pool, table,
                   (T*)y, (T*)z, mean, rstd, C, Cb, pb, momentum, eps, 0);
        TSS_LAUNCH_CHECK("ppm_branches_fwd");
        return TSS_OK;
    });
}

extern "C" int tss_ppm_branches_eval(const void* pool, const int64_t* table, void* z, int N, int C, int Cb,
                                     const int* bins, int nbins, int dtype, void* stream) {
    PpmBins pb;
    if (int e = make_bins(pb, bins, nbins, N, "ppm_branches_eval")) return e;
    TSS_REQUIRE(N > 0 && C > 0 && C % 8 == 0 && Cb > 0 && Cb % 8 == 0, "ppm_branches_eval: N=%d C=%d Cb=%d", N, C, Cb);
    const size_t smem = (size_t)8 * C * sizeof(float);
    TSS_REQUIRE(smem <= 40 * 1024, "ppm_branches_eval: C=%d too large", C);
    TSS_DISPATCH_DTYPE(dtype, "ppm_branches_eval", {
        tss_launch(ppm_branches_fwd_kernel<T>, dim3(nbins, Cb / 8), kThreads, smem, (cudaStream_t)stream, (const T*)pool, table,
                   (T*)nullptr, (T*)z, (float*)nullptr, (float*)nullptr, C, Cb, pb, 0.f, 0.f, 1);
        TSS_LAUNCH_CHECK("ppm_branches_eval");
        return TSS_OK;
    });
}

extern "C" int tss_ppm_concat_fwd(const void* x, const void* z, void* cat, int N, int H, int W, int C, int Cb,
                                  const int* bins, int nbins, int dtype, void* stream) {
    PpmBins pb;
    if (int e = make_bins(pb, bins, nbins, N, "ppm_concat_fwd")) return e;
    TSS_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && Cb > 0 && Cb % 8 == 0, "ppm_concat_fwd: bad shape");
    const int64_t items = (int64_t)N * H * W * ((C + nbins * Cb) / 8);
    int64_t grid = ceil_div64(items, kThreads);
    const int64_t cap = (int64_t)tss_num_sms() * 8;
    if (grid > cap) grid = cap;
    TSS_DISPATCH_DTYPE(dtype, "ppm_concat_fwd", {
        tss_launch(ppm_concat_fwd_kernel<T>, (int)grid, kThreads, 0, (cudaStream_t)stream, (const T*)x, (const T*)z, (T*)cat, N, H,
                   W, C, Cb, pb);
        TSS_LAUNCH_CHECK("ppm_concat_fwd");
        return TSS_OK;
    });
}

extern "C" int tss_ppm_concat_bwd(const void* dcat, void* dz, int N, int H, int W, int C, int Cb, int64_t lddcat,
                                  const int* bins, int nbins, int dtype, void* stream) {
    PpmBins pb;
    if (int e = make_bins(pb, bins, nbins, N, "ppm_concat_bwd")) return e;
    TSS_REQUIRE(N > 0 && H > 0 && W > 0 && C % 8 == 0 && Cb > 0 && Cb % 8 == 0 && lddcat >= C + nbins * Cb && lddcat % 8 == 0,
                "ppm_concat_bwd: bad shape");
    int bmax = 0;
    for (int i = 0; i < nbins; ++i) bmax = bins[i] > bmax ? bins[i] : bmax;
    const size_t smem = (size_t)bmax * W * Cb * sizeof(float);
    TSS_REQUIRE(smem <= 200 * 1024, "ppm_concat_bwd: %d x %d x %d does not fit in shared memory", bmax, W, Cb);
    TSS_DISPATCH_DTYPE(dtype, "ppm_concat_bwd", {
        if (smem > 48 * 1024)
            TSS_CUDA(cudaFuncSetAttribute(ppm_concat_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tss_launch(ppm_concat_bwd_kernel<T>, dim3(N, nbins), kThreads, smem, (cudaStream_t)stream, (const T*)dcat, (T*)dz, N, H, W,
                   C, Cb, lddcat, pb);
        TSS_LAUNCH_CHECK("ppm_concat_bwd");
        return TSS_OK;
    });
}

extern "C" int tss_ppm_branches_bwd(const void* dz, const void* y, const void* pool, const int64_t* table,
                                    const float* mean, const float* rstd, void* dy, void* dpool, int N, int C, int Cb,
                                    const int* bins, int nbins, int dtype, void* stream) {
    PpmBins pb;
    if (int e = make_bins(pb, bins, nbins, N, "ppm_branches_bwd")) return e;
    TSS_REQUIRE(N > 0 && C > 0 && C % 8 == 0 && C <= 2048 && Cb > 0 && Cb % 8 == 0, "ppm_branches_bwd: N=%d C=%d Cb=%d", N, C, Cb);
    const int KG = C / 8;
    TSS_REQUIRE(KG <= kThreads, "ppm_branches_bwd: C=%d too large", C);
    const size_t smem = (size_t)Cb * C * sizeof(float);
    TSS_REQUIRE(smem <= 200 * 1024, "ppm_branches_bwd: weights do not fit in shared memory");
    int mmax = 0;
    for (int i = 0; i < nbins; ++i) mmax = N * bins[i] * bins[i] > mmax ? N * bins[i] * bins[i] : mmax;
    const int rows_per = kThreads / KG;
    const int chunks = (mmax + rows_per - 1) / rows_per;
    TSS_DISPATCH_DTYPE(dtype, "ppm_branches_bwd", {
        tss_launch(ppm_bn_bwd_kernel<T>, dim3(nbins, Cb / 8), kThreads, 0, (cudaStream_t)stream, (const T*)dz, (const T*)y, table,
                   mean, rstd, (T*)dy, Cb, pb);
        TSS_LAUNCH_CHECK("ppm_branches_bwd(bn)");
        if (smem > 48 * 1024)
            TSS_CUDA(cudaFuncSetAttribute(ppm_dgrad_wgrad_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tss_launch(ppm_dgrad_wgrad_kernel<T>, dim3(chunks + 2, nbins), kThreads, smem, (cudaStream_t)stream, (const T*)dy,
                   (const T*)pool, table, (T*)dpool, C, Cb, pb, chunks);
        TSS_LAUNCH_CHECK("ppm_branches_bwd(dgrad/wgrad)");
        return TSS_OK;
    });
}
