// Dense 3x3 convolution (stride 1, padding 1) between equal-sized NHWC maps, expressed as the
// pointwise GEMM of pwconv_tc.cu over a tap-major patch matrix:
//
//   col[m][tap*C + c] = x[n][h + ky - 1][w + kx - 1][c]     (zero outside the map), tap = ky*3 + kx
//   y[m][co]          = sum_k col[m][k] . wk[co][k],         wk[co][tap*C + c] = w[co][c][ky][kx]
//
// The only dense 3x3 of the hot path besides the 3-channel stem is ContextNet's 128->128 layer on
// the 1/32-resolution context map (contextnet.py:55): its patch matrix is 9 x a 1/32-resolution
// activation, i.e. < 1% of the step's HBM traffic, so staging it once buys the tcgen05 GEMM
// (forward, dgrad and wgrad) unchanged.  All three kernels here are pure 128-bit streaming copies.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

inline int stream_grid(int64_t items) {
    int64_t want = ceil_div64(items, kThreads);
    const int64_t cap = (int64_t)tss_num_sms() * 8;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
im2col3x3_kernel(const T* __restrict__ x, T* __restrict__ col, int N, int H, int W, int C) {
    pdl_wait();
    const int CG = C >> 3;
    const int64_t total = (int64_t)N * H * W * 9 * CG;
    for (int64_t item = (int64_t)blockIdx.x * kThreads + threadIdx.x; item < total;
         item += (int64_t)gridDim.x * kThreads) {
        int64_t t = item;
        const int cg = (int)(t % CG); t /= CG;
        const int tap = (int)(t % 9); t /= 9;           // t = pixel index m
        const int w = (int)(t % W);
        const int h = (int)((t / W) % H);
        const int n = (int)(t / ((int64_t)W * H));
        const int hi = h + tap / 3 - 1, wi = w + tap % 3 - 1;
        uint4 v = make_uint4(0, 0, 0, 0), v2 = make_uint4(0, 0, 0, 0);
        if (hi >= 0 && hi < H && wi >= 0 && wi < W) {
            const T* src = x + (((int64_t)n * H + hi) * W + wi) * C + cg * 8;
            v = __ldg(reinterpret_cast<const uint4*>(src));
            if (sizeof(T) == 4) v2 = __ldg(reinterpret_cast<const uint4*>(src) + 1);
        }
        T* dst = col + (t * 9 + tap) * C + cg * 8;
        *reinterpret_cast<uint4*>(dst) = v;
        if (sizeof(T) == 4) *(reinterpret_cast<uint4*>(dst) + 1) = v2;
    }
}

// transpose of the gather above, as a gather: dx[n][h][w][c] = sum_tap dcol[n][h-ky+1][w-kx+1][tap][c]
template <typename T>
__global__ void __launch_bounds__(kThreads)
col2im3x3_kernel(const T* __restrict__ dcol, T* __restrict__ dx, int N, int H, int W, int C) {
    pdl_wait();
    const int CG = C >> 3;
    const int64_t total = (int64_t)N * H * W * CG;
    for (int64_t item = (int64_t)blockIdx.x * kThreads + threadIdx.x; item < total;
         item += (int64_t)gridDim.x * kThreads) {
        int64_t t = item;
        const int cg = (int)(t % CG); t /= CG;
        const int w = (int)(t % W);
        const int h = (int)((t / W) % H);
        const int n = (int)(t / ((int64_t)W * H));
        float acc[8];
        zero8(acc);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int ho = h - (tap / 3) + 1, wo = w - (tap % 3) + 1;
            if (ho < 0 || ho >= H || wo < 0 || wo >= W) continue;
            float v[8];
            load8(dcol + ((((int64_t)n * H + ho) * W + wo) * 9 + tap) * C + cg * 8, v);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] += v[e];
        }
        store8(dx + t * C + cg * 8, acc);
    }
}

// forward: wk[co][tap][c] = w[co][c][tap];  backward (accumulate): dw[co][c][tap] += dwk[co][tap][c]
__global__ void permute_w3x3_kernel(const float* __restrict__ src, float* __restrict__ dst, int Cout, int Cin,
                                    int backward) {
    pdl_wait();
    const int64_t total = (int64_t)Cout * Cin * 9;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int tap = (int)(i % 9);
        const int c = (int)((i / 9) % Cin);
        const int co = (int)(i / (9 * (int64_t)Cin));
        const int64_t k = ((int64_t)co * 9 + tap) * Cin + c;     // index in the tap-major matrix
        if (backward) dst[i] += src[k];
        else dst[k] = src[i];
    }
}

}  // namespace

extern "C" int tss_im2col3x3(const void* x, void* col, int N, int H, int W, int C, int dtype, void* stream) {
    TSS_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "im2col3x3: bad shape N=%d H=%d W=%d C=%d", N, H, W, C);
    TSS_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)col & 15) == 0, "im2col3x3: buffers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    TSS_DISPATCH_DTYPE(dtype, "im2col3x3", {
        tss_launch(im2col3x3_kernel<T>, stream_grid((int64_t)N * H * W * 9 * (C / 8)), kThreads, 0, st, (const T*)x, (T*)col, N, H, W, C);
        TSS_LAUNCH_CHECK("im2col3x3");
        return TSS_OK;
    });
}

extern "C" int tss_col2im3x3(const void* dcol, void* dx, int N, int H, int W, int C, int dtype, void* stream) {
    TSS_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "col2im3x3: bad shape N=%d H=%d W=%d C=%d", N, H, W, C);
    TSS_REQUIRE(((uintptr_t)dx & 15) == 0 && ((uintptr_t)dcol & 15) == 0, "col2im3x3: buffers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    TSS_DISPATCH_DTYPE(dtype, "col2im3x3", {
        tss_launch(col2im3x3_kernel<T>, stream_grid((int64_t)N * H * W * (C / 8)), kThreads, 0, st, (const T*)dcol, (T*)dx, N, H, W, C);
        TSS_LAUNCH_CHECK("col2im3x3");
        return TSS_OK;
    });
}

extern "C" int tss_permute_weights3x3(const float* src, float* dst, int Cout, int Cin, int backward, void* stream) {
    TSS_REQUIRE(Cout > 0 && Cin > 0, "permute_weights3x3: bad shape Cout=%d Cin=%d", Cout, Cin);
    const int64_t total = (int64_t)Cout * Cin * 9;
    int64_t grid = ceil_div64(total, 256);
    if (grid > 1184) grid = 1184;
    tss_launch(permute_w3x3_kernel, (int)grid, 256, 0, (cudaStream_t)stream, src, dst, Cout, Cin, backward);
    TSS_LAUNCH_CHECK("permute_weights3x3");
    return TSS_OK;
}
