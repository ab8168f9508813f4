// Online hard example mining (losses/ohem_loss.py:10-21) without the reference's full descending
// sort of ~7 M per-pixel losses and without its host synchronisation (`if loss[n] > thresh`):
//
//   v      = the (n+1)-th largest per-pixel loss (sorted[n]), found by a 3-pass radix select
//            (11 + 11 + 10 bits of the order-preserving integer image of the fp32 values);
//   case A (v > thresh):  mean of the losses above thresh
//   case B (otherwise):   mean of the n largest = (sum_{l > v} l + (n - #{l > v}) * v) / n
//
// Everything stays on the device: the select state, the case decision and the per-pixel weights
// of the backward pass ([cut, w_above, tie, w_tie], consumed by the CE kernels) are device scalars.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kBins = 2048;

struct SelectState {
    unsigned int prefix;    // bits decided so far (high bits of the key)
    unsigned int rank;      // descending rank still to resolve inside the prefix
    unsigned int done;      // blocks that finished the final reduction (self-resetting)
    unsigned int pad;
    double sum_gt_cut[2];   // [0]: l > thresh, [1]: l > v
    unsigned long long cnt_gt_cut[2];
    unsigned long long cnt_eq_v;
};

__device__ __forceinline__ unsigned int order_key(float f) {      // monotone: a < b  <=>  key(a) < key(b)
    const unsigned int b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(unsigned int k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__host__ __device__ __forceinline__ int pass_shift(int pass) { return pass == 0 ? 21 : (pass == 1 ? 10 : 0); }
__host__ __device__ __forceinline__ int pass_bits(int pass) { return pass == 2 ? 10 : 11; }

__global__ void __launch_bounds__(kThreads)
select_hist_kernel(const float* __restrict__ v, int64_t n, const SelectState* __restrict__ st,
                   unsigned int* __restrict__ hist, int pass) {
    pdl_wait();
    __shared__ unsigned int s_hist[kBins];
    for (int i = threadIdx.x; i < kBins; i += kThreads) s_hist[i] = 0;
    __syncthreads();
    const int shift = pass_shift(pass), bits = pass_bits(pass);
    const unsigned int prefix = pass == 0 ? 0u : st->prefix;
    const int above = shift + bits;                        // bits above this pass' digit
    const int64_t n4 = n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kThreads) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(v) + i);
        const float f[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const unsigned int k = order_key(f[e]);
            if (pass == 0 || (k >> above) == prefix) atomicAdd(&s_hist[(k >> shift) & ((1u << bits) - 1)], 1u);
        }
    }
    if (blockIdx.x == 0)
        for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += kThreads) {
            const unsigned int k = order_key(v[i]);
            if (pass == 0 || (k >> above) == prefix) atomicAdd(&s_hist[(k >> shift) & ((1u << bits) - 1)], 1u);
        }
    __syncthreads();
    for (int i = threadIdx.x; i < kBins; i += kThreads)
        if (s_hist[i]) atomicAdd(hist + i, s_hist[i]);
}

// one block: walk the bins from the top until the cumulative count passes the rank
__global__ void __launch_bounds__(kThreads)
select_scan_kernel(unsigned int* __restrict__ hist, SelectState* __restrict__ st, int pass, unsigned int rank0) {
    pdl_wait();
    __shared__ unsigned int s_cnt[kBins];
    __shared__ unsigned int s_chunk[kThreads];
    const int nb = 1 << pass_bits(pass);
    for (int i = threadIdx.x; i < kBins; i += kThreads) {
        s_cnt[i] = i < nb ? hist[i] : 0;
        hist[i] = 0;                                        // ready for the next pass / next call
    }
    __syncthreads();
    // chunk sums (8 bins per thread), descending order = bins from the top
    const int per = kBins / kThreads;
    unsigned int c = 0;
    for (int j = 0; j < per; ++j) c += s_cnt[kBins - 1 - (threadIdx.x * per + j)];
    s_chunk[threadIdx.x] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int rank = pass == 0 ? rank0 : st->rank;
        unsigned int cum = 0;
        int t = 0;
        while (t < kThreads - 1 && cum + s_chunk[t] <= rank) cum += s_chunk[t++];
        int b = kBins - 1 - t * per;
        while (b > kBins - 1 - (t * per + per - 1) && cum + s_cnt[b] <= rank) cum += s_cnt[b--];
        st->prefix = ((pass == 0 ? 0u : st->prefix) << pass_bits(pass)) | (unsigned int)b;
        st->rank = rank - cum;
        if (pass == 2) {
            st->done = 0;
            st->sum_gt_cut[0] = st->sum_gt_cut[1] = 0.0;
            st->cnt_gt_cut[0] = st->cnt_gt_cut[1] = 0ull;
            st->cnt_eq_v = 0ull;
        }
    }
}

// sums above the two cuts; the last block to finish decides the case and publishes the result
__global__ void __launch_bounds__(kThreads)
ohem_reduce_kernel(const float* __restrict__ v, int64_t n, SelectState* __restrict__ st, float thresh,
                   unsigned int n_keep, float* __restrict__ loss, float* __restrict__ weights) {
    pdl_wait();
    const float vk = key_to_float(st->prefix);
    float s0 = 0.f, s1 = 0.f;
    unsigned int c0 = 0, c1 = 0, ce = 0;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        const float f = __ldg(v + i);
        if (f > thresh) { s0 += f; ++c0; }
        if (f > vk) { s1 += f; ++c1; }
        if (f == vk) ++ce;
    }
    s0 = warp_sum(s0); s1 = warp_sum(s1);
    c0 = __reduce_add_sync(0xffffffffu, c0); c1 = __reduce_add_sync(0xffffffffu, c1); ce = __reduce_add_sync(0xffffffffu, ce);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&st->sum_gt_cut[0], (double)s0);
        atomicAdd(&st->sum_gt_cut[1], (double)s1);
        atomicAdd(&st->cnt_gt_cut[0], (unsigned long long)c0);
        atomicAdd(&st->cnt_gt_cut[1], (unsigned long long)c1);
        atomicAdd(&st->cnt_eq_v, (unsigned long long)ce);
    }
    __threadfence();
    __syncthreads();
    __shared__ bool s_last;
    if (threadIdx.x == 0) s_last = atomicAdd(&st->done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        const volatile SelectState* vs = st;
        if (vk > thresh) {                                  // losses/ohem_loss.py:18-19
            const double cnt = (double)vs->cnt_gt_cut[0];
            *loss = (float)(vs->sum_gt_cut[0] / cnt);
            weights[0] = thresh; weights[1] = (float)(1.0 / cnt); weights[2] = -1.f; weights[3] = 0.f;
        } else {                                            // losses/ohem_loss.py:20-21: top n
            const double gt = (double)vs->cnt_gt_cut[1], eq = (double)vs->cnt_eq_v, k = (double)n_keep;
            *loss = (float)((vs->sum_gt_cut[1] + (k - gt) * (double)vk) / k);       // n_keep == 0 -> NaN like mean([])
            weights[0] = vk; weights[1] = (float)(1.0 / k); weights[2] = vk;
            weights[3] = eq > 0.0 ? (float)((k - gt) / (eq * k)) : 0.f;            // ties share the remaining slots
        }
    }
}

}  // namespace

extern "C" int64_t tss_ohem_workspace_bytes(void) {
    return (int64_t)(kBins * sizeof(unsigned int) + sizeof(SelectState));
}

extern "C" int tss_ohem_select(const float* pixel_loss, int64_t n, int64_t n_keep, float thresh, void* workspace,
                               float* loss, float* weights, void* stream) {
    TSS_REQUIRE(n > 0 && n_keep >= 0 && n_keep < n, "ohem_select: need 0 <= n_keep < n (n=%lld n_keep=%lld)",
                (long long)n, (long long)n_keep);
    TSS_REQUIRE(n < ((int64_t)1 << 32), "ohem_select: n=%lld too large", (long long)n);
    TSS_REQUIRE(((uintptr_t)pixel_loss & 15) == 0 && ((uintptr_t)workspace & 15) == 0, "ohem_select: buffers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned int* hist = (unsigned int*)workspace;          // zero on entry (the scan leaves it zero again)
    SelectState* state = (SelectState*)(hist + kBins);
    int64_t want = ceil_div64(n / 4 + 1, kThreads * 4);
    const int64_t cap = (int64_t)tss_num_sms() * 4;
    const int grid = (int)(want < 1 ? 1 : (want < cap ? want : cap));
    for (int pass = 0; pass < 3; ++pass) {
        tss_launch(select_hist_kernel, grid, kThreads, 0, st, pixel_loss, n, (const SelectState*)state, hist, pass);
        TSS_LAUNCH_CHECK("ohem_select(hist)");
        tss_launch(select_scan_kernel, 1, kThreads, 0, st, hist, state, pass, (unsigned int)n_keep);
        TSS_LAUNCH_CHECK("ohem_select(scan)");
    }
    tss_launch(ohem_reduce_kernel, grid, kThreads, 0, st, pixel_loss, n, state, thresh, (unsigned int)n_keep, loss, weights);
    TSS_LAUNCH_CHECK("ohem_select(reduce)");
    return TSS_OK;
}
