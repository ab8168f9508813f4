// Softmax cross-entropy with ignore_index: forward and gradient fused in ONE pass over
// the class planes of NCHW logits (the layout the reference's model returns).
// Every thread owns 4 consecutive pixels: per class plane one 64-bit (bf16) / 128-bit (fp32)
// coalesced load, the C logits of its pixels stay in registers, and the gradient
// (softmax - onehot)/n_valid is written back plane by plane with the same access pattern.
// Algorithmic traffic in bf16: 2C (logits) + 8 (int64 label) + 2C (gradient) bytes / pixel.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxC = 32;

template <typename T> struct Quad;
template <> struct Quad<float> {
    __device__ static void ld(const float* p, float (&v)[4]) {
        float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ static void st(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct Quad<bf16> {
    __device__ static void ld(const bf16* p, float (&v)[4]) {
        uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
        v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
        v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
    }
    __device__ static void st(bf16* p, const float (&v)[4]) {
        uint2 u; u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
        *reinterpret_cast<uint2*>(p) = u;
    }
};

__global__ void __launch_bounds__(kThreads)
count_valid_kernel(const int64_t* __restrict__ target, int64_t n, int64_t ignore_index,
                   unsigned long long* __restrict__ nvalid) {
    unsigned int cnt = 0;
    const int64_t n2 = n >> 1;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n2; i += (int64_t)gridDim.x * kThreads) {
        const longlong2 t = __ldg(reinterpret_cast<const longlong2*>(target) + i);
        cnt += (t.x != ignore_index) + (t.y != ignore_index);
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) cnt += target[n - 1] != ignore_index;
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    __shared__ unsigned int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_cnt, cnt);
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt) atomicAdd(nvalid, (unsigned long long)s_cnt);
}

template <typename T, int C>
__global__ void __launch_bounds__(kThreads)
ce_fwd_kernel(const T* __restrict__ logits, const int64_t* __restrict__ target, int N, int64_t HW,
              int64_t ignore_index, const int64_t* __restrict__ nvalid, double* __restrict__ loss_sum,
              float* __restrict__ pixel_loss, T* __restrict__ dlogits) {
    const int64_t quads_per_img = HW >> 2;
    const int64_t total = (int64_t)N * quads_per_img;
    const float inv_n = dlogits != nullptr ? 1.f / (float)(*nvalid) : 0.f;   // inf when no valid pixel; masked below
    float lsum = 0.f;
    for (int64_t q = (int64_t)blockIdx.x * kThreads + threadIdx.x; q < total; q += (int64_t)gridDim.x * kThreads) {
        const int64_t n = q / quads_per_img;
        const int64_t hw = (q - n * quads_per_img) << 2;
        const T* lp = logits + n * C * HW + hw;
        float x[C][4];
#pragma unroll
        for (int c = 0; c < C; ++c) Quad<T>::ld(lp + (int64_t)c * HW, x[c]);
        const longlong2 ta = __ldg(reinterpret_cast<const longlong2*>(target + n * HW + hw));
        const longlong2 tb = __ldg(reinterpret_cast<const longlong2*>(target + n * HW + hw) + 1);
        const int64_t tg[4] = {ta.x, ta.y, tb.x, tb.y};
        float mx[4], se[4], xt[4], pl[4];
        bool valid[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            mx[p] = x[0][p];
#pragma unroll
            for (int c = 1; c < C; ++c) mx[p] = fmaxf(mx[p], x[c][p]);
            se[p] = 0.f;
            xt[p] = 0.f;
            valid[p] = tg[p] != ignore_index && tg[p] >= 0 && tg[p] < C;
        }
#pragma unroll
        for (int c = 0; c < C; ++c)
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                if ((int64_t)c == tg[p]) xt[p] = x[c][p];
                x[c][p] = __expf(x[c][p] - mx[p]);
                se[p] += x[c][p];
            }
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            pl[p] = valid[p] ? (__logf(se[p]) + mx[p] - xt[p]) : 0.f;
            lsum += pl[p];
        }
        if (pixel_loss != nullptr) *reinterpret_cast<float4*>(pixel_loss + n * HW + hw) = make_float4(pl[0], pl[1], pl[2], pl[3]);
        if (dlogits != nullptr) {
            float k[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) k[p] = valid[p] ? inv_n / se[p] : 0.f;
            T* gp = dlogits + n * C * HW + hw;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                float g[4];
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    g[p] = x[c][p] * k[p];
                    if (valid[p] && (int64_t)c == tg[p]) g[p] -= inv_n;
                }
                Quad<T>::st(gp + (int64_t)c * HW, g);
            }
        }
    }
    if (loss_sum != nullptr) {
        lsum = warp_sum(lsum);
        __shared__ float s_part[kThreads / 32];
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = lsum;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w) s += (double)s_part[w];
            atomicAdd(loss_sum, s);
        }
    }
}

__global__ void ce_finalize_kernel(const double* __restrict__ loss_sum, const int64_t* __restrict__ nvalid,
                                   float* __restrict__ loss) {
    *loss = (float)(*loss_sum / (double)(*nvalid));   // 0/0 -> NaN like the reference
}

template <typename T, int C>
int launch_ce(const void* logits, const int64_t* target, int N, int64_t HW, int64_t ignore_index,
              const int64_t* nvalid, double* loss_sum, float* pixel_loss, void* dlogits, cudaStream_t st) {
    const int64_t total = (int64_t)N * (HW >> 2);
    int64_t want = ceil_div64(total, kThreads);
    int64_t cap = (int64_t)tss_num_sms() * 8;
    const int grid = (int)(want < 1 ? 1 : (want < cap ? want : cap));
    ce_fwd_kernel<T, C><<<grid, kThreads, 0, st>>>((const T*)logits, target, N, HW, ignore_index, nvalid,
                                                   loss_sum, pixel_loss, (T*)dlogits);
    TSS_LAUNCH_CHECK("ce_fwd");
    return TSS_OK;
}

}  // namespace

extern "C" int tss_ce_count_valid(const int64_t* target, int64_t n, int64_t ignore_index, int64_t* nvalid,
                                  void* stream) {
    TSS_REQUIRE(n > 0, "ce_count_valid: empty target");
    TSS_REQUIRE(((uintptr_t)target & 15) == 0, "ce_count_valid: target must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    TSS_CUDA(cudaMemsetAsync(nvalid, 0, sizeof(int64_t), st));
    int64_t want = ceil_div64(n / 2 + 1, kThreads * 4);
    int64_t cap = (int64_t)tss_num_sms() * 8;
    const int grid = (int)(want < 1 ? 1 : (want < cap ? want : cap));
    count_valid_kernel<<<grid, kThreads, 0, st>>>(target, n, ignore_index, (unsigned long long*)nvalid);
    TSS_LAUNCH_CHECK("ce_count_valid");
    return TSS_OK;
}

extern "C" int tss_ce_fwd(const void* logits, const int64_t* target, int N, int C, int64_t HW,
                          int64_t ignore_index, const int64_t* nvalid, double* loss_sum, float* pixel_loss,
                          void* dlogits, int dtype, void* stream) {
    TSS_REQUIRE(N > 0 && HW > 0, "ce_fwd: empty input");
    TSS_REQUIRE(HW % 4 == 0, "ce_fwd: H*W=%lld must be a multiple of 4", (long long)HW);
    TSS_REQUIRE(dlogits == nullptr || nvalid != nullptr, "ce_fwd: gradient needs nvalid");
    cudaStream_t st = (cudaStream_t)stream;
    TSS_DISPATCH_DTYPE(dtype, "ce_fwd", {
        switch (C) {
            case 19: return launch_ce<T, 19>(logits, target, N, HW, ignore_index, nvalid, loss_sum, pixel_loss, dlogits, st);
            case 2: return launch_ce<T, 2>(logits, target, N, HW, ignore_index, nvalid, loss_sum, pixel_loss, dlogits, st);
            case 8: return launch_ce<T, 8>(logits, target, N, HW, ignore_index, nvalid, loss_sum, pixel_loss, dlogits, st);
            case 11: return launch_ce<T, 11>(logits, target, N, HW, ignore_index, nvalid, loss_sum, pixel_loss, dlogits, st);
            case 12: return launch_ce<T, 12>(logits, target, N, HW, ignore_index, nvalid, loss_sum, pixel_loss, dlogits, st);
            case 20: return launch_ce<T, 20>(logits, target, N, HW, ignore_index, nvalid, loss_sum, pixel_loss, dlogits, st);
            case 21: return launch_ce<T, 21>(logits, target, N, HW, ignore_index, nvalid, loss_sum, pixel_loss, dlogits, st);
            default:
                tss_set_error("ce_fwd: C=%d not instantiated (19 Cityscapes/BDD, 11/12 CamVid, 20/21 VOC, 2, 8)", C);
                return TSS_ERR_ARG;
        }
    });
}

extern "C" int tss_ce_finalize(const double* loss_sum, const int64_t* nvalid, float* loss, void* stream) {
    ce_finalize_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(loss_sum, nvalid, loss);
    TSS_LAUNCH_CHECK("ce_finalize");
    return TSS_OK;
}
