// Inverted dropout (nn.Dropout(0.1) in front of the class-score conv; fastscnn.py:96, contextnet.py:84) without a
// mask tensor: keep/drop comes from a counter-based generator (Philox4x32-10) keyed by (seed, step offset, element
// group), so the backward pass regenerates the very mask of the forward pass from two integers.  The offset
// lives on the device and is advanced by the forward kernel itself (last CTA to have read it, ticket counter), so
// a captured CUDA graph draws a fresh mask at every replay.  One 8-channel group per thread: one Philox call
// yields 4 x 32 bits = eight 16-bit uniforms, keep iff u16 >= p * 65536.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
    const unsigned long long p = (unsigned long long)a * b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
}

__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0, l0, h1, l1;
        mulhilo(0xD2511F53u, c[0], h0, l0);
        mulhilo(0xCD9E8D57u, c[2], h1, l1);
        const uint32_t n0 = h1 ^ c[1] ^ k0, n2 = h0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = l1; c[2] = n2; c[3] = l0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

// keep mask (bit e = element e of the group) of 8-element group `group` under (seed, offset)
__device__ __forceinline__ uint32_t keep_mask8(unsigned long long seed, unsigned long long offset, unsigned long long group,
                                               uint32_t thresh) {
    uint32_t c[4] = {(uint32_t)group, (uint32_t)(group >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    uint32_t m = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const uint32_t u16 = (c[e >> 1] >> ((e & 1) * 16)) & 0xffffu;
        m |= (u16 >= thresh ? 1u : 0u) << e;
    }
    return m;
}

// rng: {seed, offset, ticket}; `used` (may alias nothing else) receives the offset this launch drew its mask with
template <typename T, bool kForward>
__global__ void __launch_bounds__(kThreads)
dropout_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t groups, uint32_t thresh, float scale,
               unsigned long long* rng, unsigned long long* used) {
    __shared__ unsigned long long s_state[2];
    pdl_wait();
    if (threadIdx.x == 0) {
        s_state[0] = rng[0];
        s_state[1] = kForward ? rng[1] : used[0];
    }
    __syncthreads();
    const unsigned long long seed = s_state[0], offset = s_state[1];
    if (kForward) {
        if (threadIdx.x == 0) {                       // this CTA has read the offset: the last one advances it
            if (blockIdx.x == 0) used[0] = offset;
            __threadfence();
            const unsigned long long t = atomicAdd(rng + 2, 1ull);
            if (t == (unsigned long long)gridDim.x - 1) {
                rng[1] = offset + 1;
                rng[2] = 0;
            }
        }
    }
    for (int64_t g = (int64_t)blockIdx.x * kThreads + threadIdx.x; g < groups; g += (int64_t)gridDim.x * kThreads) {
        const uint32_t m = keep_mask8(seed, offset, (unsigned long long)g, thresh);
        float v[8];
        load8(x + g * 8, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = (m >> e) & 1u ? v[e] * scale : 0.f;
        store8(y + g * 8, v);
    }
}

template <bool kForward>
int launch(const void* x, void* y, int64_t n, float p, int64_t* rng, int64_t* used, int dtype, void* stream, const char* name) {
    TSS_REQUIRE(n > 0 && n % 8 == 0, "%s: n=%lld must be a positive multiple of 8", name, (long long)n);
    TSS_REQUIRE(p >= 0.f && p < 1.f, "%s: p=%f", name, (double)p);
    TSS_REQUIRE(x != nullptr && y != nullptr && rng != nullptr && used != nullptr, "%s: missing buffer", name);
    TSS_REQUIRE((((uintptr_t)x | (uintptr_t)y) & 15) == 0, "%s: tensors must be 16-byte aligned", name);
    const uint32_t thresh = (uint32_t)(p * 65536.f + 0.5f);
    const float scale = 1.f / (1.f - (float)thresh / 65536.f);       // exactly the keep probability the mask realises
    int64_t grid = ceil_div64(n / 8, kThreads);
    const int64_t cap = (int64_t)tss_num_sms() * 8;
    if (grid > cap) grid = cap;
    TSS_DISPATCH_DTYPE(dtype, name, {
        tss_launch(dropout_kernel<T, kForward>, (unsigned)grid, kThreads, 0, (cudaStream_t)stream, (const T*)x, (T*)y, n / 8, thresh,
                   scale, (unsigned long long*)rng, (unsigned long long*)used);
        TSS_LAUNCH_CHECK(name);
        return TSS_OK;
    });
}

}  // namespace

extern "C" int tss_dropout_fwd(const void* x, void* y, int64_t n, float p, int64_t* rng, int64_t* used, int dtype,
                               void* stream) {
    return launch<true>(x, y, n, p, rng, used, dtype, stream, "dropout_fwd");
}

extern "C" int tss_dropout_bwd(const void* dy, void* dx, int64_t n, float p, int64_t* rng, int64_t* used, int dtype,
                               void* stream) {
    return launch<false>(dy, dx, n, p, rng, used, dtype, stream, "dropout_bwd");
}
