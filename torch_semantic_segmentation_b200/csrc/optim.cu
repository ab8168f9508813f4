// torch.optim.AdamW over one flat fp32 parameter arena: a single streaming kernel per step
// (16 B/param read, 12 B/param written) instead of one multi-tensor launch chain per
// parameter group.  Hyper-parameters live in device memory so that a captured CUDA graph
// can be replayed while the host changes the learning rate between replays.
#include "common.cuh"

namespace {

// hyper: [0] lr [1] beta1 [2] beta2 [3] eps [4] weight_decay [5] step [6] 1/bias_correction1
//        [7] sqrt(bias_correction2)
__global__ void adamw_prepare_kernel(float* __restrict__ hyper) {
    pdl_wait();
    const double step = (double)hyper[5] + 1.0;
    hyper[5] = (float)step;
    hyper[6] = (float)(1.0 / (1.0 - pow((double)hyper[1], step)));
    hyper[7] = (float)sqrt(1.0 - pow((double)hyper[2], step));
}

__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             int64_t n4, int64_t n, const float* __restrict__ hyper, float grad_scale) {
    pdl_wait_hold();        // writes parameters: no dependent grid beside it
    const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4];
    const float step_size = lr * hyper[6], bc2s = hyper[7];
    const float decay = 1.f - lr * wd;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
        float4 pp = reinterpret_cast<float4*>(p)[i];
        const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
        float4 mm = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
        float* pa = &pp.x; const float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float gr = ga[e] * grad_scale;
            pa[e] *= decay;
            ma[e] = ma[e] + (gr - ma[e]) * (1.f - b1);
            va[e] = va[e] * b2 + gr * gr * (1.f - b2);
            const float denom = sqrtf(va[e]) / bc2s + eps;
            pa[e] -= step_size * (ma[e] / denom);
        }
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
    // scalar tail (n % 4)
    if (blockIdx.x == 0 && threadIdx.x < (int)(n - n4 * 4)) {
        const int64_t i = n4 * 4 + threadIdx.x;
        const float gr = g[i] * grad_scale;
        float pv = p[i] * decay;
        const float mv = m[i] + (gr - m[i]) * (1.f - b1);
        const float vv = v[i] * b2 + gr * gr * (1.f - b2);
        pv -= step_size * (mv / (sqrtf(vv) / bc2s + eps));
        p[i] = pv; m[i] = mv; v[i] = vv;
    }
}

}  // namespace

extern "C" int tss_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float* hyper,
                              float grad_scale, void* stream) {
    TSS_REQUIRE(n > 0, "adamw_step: n=%lld", (long long)n);
    TSS_REQUIRE((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0, "adamw_step: arenas must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    tss_launch(adamw_prepare_kernel, 1, 1, 0, st, hyper);
    TSS_LAUNCH_CHECK("adamw_prepare");
    const int64_t n4 = n / 4;
    int64_t want = ceil_div64(n4 > 0 ? n4 : 1, 256);
    int64_t cap = (int64_t)tss_num_sms() * 8;
    tss_launch(adamw_kernel, (int)(want < cap ? want : cap), 256, 0, st, p, g, m, v, n4, n, hyper, grad_scale);
    TSS_LAUNCH_CHECK("adamw_step");
    return TSS_OK;
}
