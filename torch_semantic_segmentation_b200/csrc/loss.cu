// Softmax cross-entropy with ignore_index: forward and gradient fused in ONE pass over
// the class planes of NCHW logits (the layout the reference's model returns).
// Every thread owns 4 consecutive pixels: per class plane one 64-bit (bf16) / 128-bit (fp32)
// coalesced load, the C logits of its pixels stay in registers, and the gradient
// (softmax - onehot)/n_valid is written back plane by plane with the same access pattern.
// Algorithmic traffic in bf16: 2C (logits) + 8 (int64 label) + 2C (gradient) bytes / pixel.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

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
count_valid_kernel(const int64_t* __restrict__ target, int64_t n, int64_t ignore_index, int64_t num_classes,
                   unsigned long long* __restrict__ nvalid) {
    pdl_wait();
    // the same predicate as ce_fwd_kernel / upsample_ce_kernel: a label outside [0, C) carries no loss and no
    // gradient there, so it must not sit in the mean's denominator either
    auto counts = [=](int64_t t) { return (unsigned int)(t != ignore_index && t >= 0 && t < num_classes); };
    unsigned int cnt = 0;
    const int64_t n2 = n >> 1;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n2; i += (int64_t)gridDim.x * kThreads) {
        const longlong2 t = __ldg(reinterpret_cast<const longlong2*>(target) + i);
        cnt += counts(t.x) + counts(t.y);
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) cnt += counts(target[n - 1]);
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
              float* __restrict__ pixel_loss, T* __restrict__ dlogits, const float* __restrict__ ohem) {
    pdl_wait();
    const int64_t quads_per_img = HW >> 2;
    const int64_t total = (int64_t)N * quads_per_img;
    const float inv_n = (dlogits != nullptr && ohem == nullptr) ? 1.f / (float)(*nvalid) : 0.f;   // inf when no valid pixel; masked below
    float o_cut = 0.f, o_above = 0.f, o_tie = 0.f, o_wtie = 0.f;       // OHEM: per-pixel weight from the pixel's own loss
    if (ohem != nullptr) { o_cut = __ldg(ohem); o_above = __ldg(ohem + 1); o_tie = __ldg(ohem + 2); o_wtie = __ldg(ohem + 3); }
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
            float k[4], wgt[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                wgt[p] = ohem == nullptr ? inv_n : (pl[p] > o_cut ? o_above : (pl[p] == o_tie ? o_wtie : 0.f));
                k[p] = valid[p] ? wgt[p] / se[p] : 0.f;
            }
            T* gp = dlogits + n * C * HW + hw;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                float g[4];
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    g[p] = x[c][p] * k[p];
                    if (valid[p] && (int64_t)c == tg[p]) g[p] -= wgt[p];
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
    pdl_wait();
    *loss = (float)(*loss_sum / (double)(*nvalid));   // 0/0 -> NaN like the reference
}

template <typename T, int C>
int launch_ce(const void* logits, const int64_t* target, int N, int64_t HW, int64_t ignore_index,
              const int64_t* nvalid, double* loss_sum, float* pixel_loss, void* dlogits, const float* ohem, cudaStream_t st) {
    const int64_t total = (int64_t)N * (HW >> 2);
    int64_t want = ceil_div64(total, kThreads);
    int64_t cap = (int64_t)tss_num_sms() * 8;
    const int grid = (int)(want < 1 ? 1 : (want < cap ? want : cap));
    tss_launch(ce_fwd_kernel<T, C>, grid, kThreads, 0, st, (const T*)logits, target, N, HW, ignore_index, nvalid,
                                                   loss_sum, pixel_loss, (T*)dlogits, ohem);
    TSS_LAUNCH_CHECK("ce_fwd");
    return TSS_OK;
}

// ---------------------------------------------------------------------------------------------
// Fused head: x8 bilinear up-sampling (align_corners=True) + softmax cross-entropy + the
// gradient w.r.t. the LOW-resolution class scores, in one pass that reads only the labels.
// The full-resolution logits are recomputed on the fly from the 1/8-resolution NHWC scores and the
// full-resolution gradient never exists: (softmax - onehot) is folded straight through the
// transpose of the interpolation.
//   CTA = (image n, source row interval hi, chunk of kCols output columns); thread = output column.
//   The thread interpolates horizontally once (a[c], b[c] for source rows hi / hi+1), then walks the
//   8-9 output rows whose source row is hi: vertical blend -> softmax -> loss -> g; g is blended
//   back vertically into t0[c] (row hi) / t1[c] (row hi+1) in registers.  The column partials go to
//   shared memory and (source column, class) threads gather their ~17 output columns: two fp32
//   atomics per (source pixel, class) per CTA.  Gradients are accumulated UNSCALED; the valid
//   count is produced by the same pass and applied by tss_upsample_ce_finalize.
constexpr int kHeadCols = 256;
constexpr int kHeadMaxRows = 40;      // output rows per source-row interval: ceil(1/scale) + 1 (x32 heads: 34)

__device__ __forceinline__ float ex2_approx(float x) {
#ifdef TSS_HOST_EMU
    return exp2f(x);            // host build for tests/simt_emu
#else
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#endif
}

template <typename T, int C>
__global__ void __launch_bounds__(kHeadCols, 2)
upsample_ce_kernel(const T* __restrict__ x, const int64_t* __restrict__ target, float* __restrict__ dx32,
                   double* __restrict__ loss_sum, unsigned long long* __restrict__ nvalid,
                   float* __restrict__ pixel_loss, const float* __restrict__ ohem, int Hi, int Wi, int Ho, int Wo,
                   int64_t ldx, int64_t lddx, int64_t ignore_index, float sh, float sw, int chunks) {
    pdl_wait();
    TSS_DYN_SMEM(float, s_mem);
    float o_cut = 0.f, o_above = 0.f, o_tie = 0.f, o_wtie = 0.f;       // OHEM: per-pixel weight from the pixel's own loss
    if (ohem != nullptr) { o_cut = __ldg(ohem); o_above = __ldg(ohem + 1); o_tie = __ldg(ohem + 2); o_wtie = __ldg(ohem + 3); }
    constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
    constexpr int TP = 2 * C + 1;
    __shared__ int s_nrows;
    __shared__ int s_row_ho[kHeadMaxRows];
    __shared__ float s_row_k0[kHeadMaxRows], s_row_k1[kHeadMaxRows], s_row_lh[kHeadMaxRows];
    __shared__ int s_w0[kHeadCols];
    __shared__ float s_lw[kHeadCols];
    int b = blockIdx.x;
    const int chunk = b % chunks; b /= chunks;
    const int hi = b % Hi;
    const int n = b / Hi;
    const int wo_lo = chunk * kHeadCols;
    const int wo_hi = min(wo_lo + kHeadCols, Wo) - 1;             // inclusive
    int wl0, wl1, wh0, wh1; float tmp;
    ac_source(sw, wo_lo, Wi, wl0, wl1, tmp);
    ac_source(sw, wo_hi, Wi, wh0, wh1, tmp);
    const int wi_lo = wl0, ncols = wh1 - wl0 + 1;                 // source columns touched by this chunk
    float* s_src = s_mem;                                         // [2][ncols][C], pre-scaled by log2(e)
    float* s_t = s_mem + 2 * ncols * C;                           // [kHeadCols][TP]: t0 | t1 per output column
    const int h1src = hi + (hi < Hi - 1 ? 1 : 0);
    for (int i = threadIdx.x; i < 2 * ncols * C; i += kHeadCols) {
        const int r = i / (ncols * C), rem = i - r * ncols * C;
        const int w = rem / C, c = rem - w * C;
        s_src[i] = kLog2e * to_f32(x[(((int64_t)n * Hi + (r ? h1src : hi)) * Wi + wi_lo + w) * ldx + c]);
    }
    if (threadIdx.x == 0) {                                       // the output rows whose upper source row is hi
        int cnt = 0;
        const int ho_a = first_candidate(sh, hi, Ho), ho_b = last_candidate(sh, hi, Ho);
        for (int ho = ho_a; ho <= ho_b && cnt < kHeadMaxRows; ++ho) {
            int h0, h1; float lh;
            ac_source(sh, ho, Hi, h0, h1, lh);
            if (h0 != hi) continue;
            s_row_ho[cnt] = ho;
            s_row_lh[cnt] = lh;
            s_row_k0[cnt] = (h1 == hi) ? 1.f : 1.f - lh;          // clamped last row: both weights land on hi
            s_row_k1[cnt] = (h1 == hi) ? 0.f : lh;
            ++cnt;
        }
        s_nrows = cnt;
    }
    const int wo = wo_lo + threadIdx.x;
    const bool col_ok = wo <= wo_hi;
    int w0 = 0, w1 = 0; float lw = 0.f;
    if (col_ok) ac_source(sw, wo, Wi, w0, w1, lw);
    s_w0[threadIdx.x] = col_ok ? w0 : -4;
    s_lw[threadIdx.x] = lw;
    float* my_t = s_t + threadIdx.x * TP;
#pragma unroll
    for (int c = 0; c < 2 * C; ++c) my_t[c] = 0.f;                // the -onehot part is accumulated here
    __syncthreads();

    float t0[C], t1[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { t0[c] = 0.f; t1[c] = 0.f; }
    float lsum = 0.f;
    unsigned int cnt = 0;
    if (col_ok) {
        float a[C], d[C];                                         // row hi, and (row hi+1) - (row hi), in log2 units
        const float* p00 = s_src + (w0 - wi_lo) * C;
        const float* p01 = s_src + (w1 - wi_lo) * C;
        const float* p10 = p00 + ncols * C;
        const float* p11 = p01 + ncols * C;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            a[c] = fmaf(lw, p01[c] - p00[c], p00[c]);
            d[c] = fmaf(lw, p11[c] - p10[c], p10[c]) - a[c];
        }
        const int nrows = s_nrows;
        // the labels of the rows are requested four at a time, before the first of them is used (r2: the dependent
        // 8-byte load at the top of every row iteration was the longest stall of this kernel)
        for (int r0 = 0; r0 < nrows; r0 += 4) {
            int64_t tg4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                tg4[j] = r0 + j < nrows ? __ldg(target + ((int64_t)n * Ho + s_row_ho[r0 + j]) * Wo + wo) : ignore_index;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
            const int r = r0 + j;
            if (r >= nrows) break;
            const int ho = s_row_ho[r];
            const float lh = s_row_lh[r];
            const int64_t pix = ((int64_t)n * Ho + ho) * Wo + wo;
            const int64_t tg64 = tg4[j];
            const bool valid = tg64 != ignore_index && tg64 >= 0 && tg64 < C;
            const int tg = valid ? (int)tg64 : 0;
            float v[C];
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                v[c] = fmaf(lh, d[c], a[c]);
                mx = fmaxf(mx, v[c]);
            }
            float se = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                v[c] = ex2_approx(v[c] - mx);
                se += v[c];
            }
            float pl = 0.f;
            if (valid || pixel_loss != nullptr) {
                // the target class' logit again, from the staged source (dynamic index: shared memory)
                const float xa = fmaf(lw, p01[tg] - p00[tg], p00[tg]);
                const float xb = fmaf(lw, p11[tg] - p10[tg], p10[tg]);
                const float xt = fmaf(lh, xb - xa, xa);
                pl = valid ? kLn2 * (__log2f(se) + mx - xt) : 0.f;
                if (pixel_loss != nullptr) pixel_loss[pix] = pl;
                lsum += pl;
            }
            if (valid) {
                ++cnt;
                const float wgt = ohem == nullptr ? 1.f : (pl > o_cut ? o_above : (pl == o_tie ? o_wtie : 0.f));
                const float inv = 1.f / se;
                const float k0 = s_row_k0[r] * wgt, k1 = s_row_k1[r] * wgt;
                const float ik0 = inv * k0, ik1 = inv * k1;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    t0[c] = fmaf(v[c], ik0, t0[c]);
                    t1[c] = fmaf(v[c], ik1, t1[c]);
                }
                my_t[tg] -= k0;
                my_t[C + tg] -= k1;
            }
            }   // j
        }
    }
    if (dx32 != nullptr) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            my_t[c] += t0[c];
            my_t[C + c] += t1[c];
        }
        __syncthreads();
        const bool has_next = hi < Hi - 1;
        for (int i = threadIdx.x; i < ncols * C; i += kHeadCols) {
            const int w = i / C, c = i - w * C;
            const int wi = wi_lo + w;
            int ca = first_candidate(sw, wi, Wo), cb = last_candidate(sw, wi, Wo);
            ca = max(ca, wo_lo) - wo_lo; cb = min(cb, wo_hi) - wo_lo;
            float a0 = 0.f, a1 = 0.f;
            for (int o = ca; o <= cb; ++o) {
                const int ow0 = s_w0[o];
                const float olw = s_lw[o];
                // weight of output column o on source column wi (w1 == w0 at the clamped right edge)
                const int ow1 = ow0 + (ow0 < Wi - 1 ? 1 : 0);
                const float ww = (ow0 == wi ? 1.f - olw : 0.f) + (ow1 == wi ? olw : 0.f);
                a0 = fmaf(s_t[o * TP + c], ww, a0);
                a1 = fmaf(s_t[o * TP + C + c], ww, a1);
            }
            if (a0 != 0.f) atomicAdd(dx32 + (((int64_t)n * Hi + hi) * Wi + wi) * lddx + c, a0);
            if (has_next && a1 != 0.f) atomicAdd(dx32 + (((int64_t)n * Hi + hi + 1) * Wi + wi) * lddx + c, a1);
        }
    }
    lsum = warp_sum(lsum);
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    __shared__ float s_part[kHeadCols / 32];
    __shared__ unsigned int s_cnt[kHeadCols / 32];
    if ((threadIdx.x & 31) == 0) { s_part[threadIdx.x >> 5] = lsum; s_cnt[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        unsigned int k = 0;
#pragma unroll
        for (int w = 0; w < kHeadCols / 32; ++w) { s += (double)s_part[w]; k += s_cnt[w]; }
        if (loss_sum != nullptr) atomicAdd(loss_sum, s);
        if (k) atomicAdd(nvalid, (unsigned long long)k);
    }
}

// loss = loss_sum / nvalid; dx (activation dtype) = dx32 / nvalid * upstream  (n % 8 == 0 elements)
template <typename T>
__global__ void __launch_bounds__(kThreads)
upsample_ce_finalize_kernel(const double* __restrict__ loss_sum, const int64_t* __restrict__ nvalid,
                            float* __restrict__ loss, const float* __restrict__ dx32, T* __restrict__ dx, int64_t n8) {
    pdl_wait();
    const double nv = (double)(*nvalid);
    if (blockIdx.x == 0 && threadIdx.x == 0 && loss != nullptr) *loss = (float)(*loss_sum / nv);
    if (dx == nullptr) return;
    const float k = (float)(1.0 / nv);
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n8; i += (int64_t)gridDim.x * kThreads) {
        float v[8];
        load8(dx32 + i * 8, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] *= k;
        store8(dx + i * 8, v);
    }
}

template <typename T, int C>
int launch_head(const void* x, const int64_t* target, float* dx32, double* loss_sum, int64_t* nvalid, float* pixel_loss,
                const float* ohem, int N, int Hi, int Wi, int Ho, int Wo, int64_t ldx, int64_t lddx, int64_t ignore_index, cudaStream_t st) {
    const int chunks = (Wo + kHeadCols - 1) / kHeadCols;
    const float sw = ac_scale(Wi, Wo);
    int ncols_max = (int)((float)kHeadCols * sw) + 4;
    if (ncols_max > Wi) ncols_max = Wi;
    const size_t smem = ((size_t)2 * ncols_max * C + (size_t)kHeadCols * (2 * C + 1)) * sizeof(float);
    TSS_REQUIRE(smem <= 200 * 1024, "upsample_ce: tile does not fit in shared memory (Wi=%d Wo=%d)", Wi, Wo);
    auto kern = upsample_ce_kernel<T, C>;
    static bool attr_set = false;
    if (!attr_set) {
        TSS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    tss_launch(kern, N * Hi * chunks, kHeadCols, smem, st, (const T*)x, target, dx32, loss_sum, (unsigned long long*)nvalid,
                                                   pixel_loss, ohem, Hi, Wi, Ho, Wo, ldx, lddx, ignore_index,
                                                   ac_scale(Hi, Ho), sw, chunks);
    TSS_LAUNCH_CHECK("upsample_ce_fwd");
    return TSS_OK;
}

}  // namespace

extern "C" int tss_ce_count_valid(const int64_t* target, int64_t n, int64_t ignore_index, int64_t num_classes,
                                  int64_t* nvalid, void* stream) {
    TSS_REQUIRE(n > 0, "ce_count_valid: empty target");
    TSS_REQUIRE(num_classes > 0, "ce_count_valid: num_classes must be positive");
    TSS_REQUIRE(((uintptr_t)target & 15) == 0, "ce_count_valid: target must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    TSS_CUDA(cudaMemsetAsync(nvalid, 0, sizeof(int64_t), st));
    int64_t want = ceil_div64(n / 2 + 1, kThreads * 4);
    int64_t cap = (int64_t)tss_num_sms() * 8;
    const int grid = (int)(want < 1 ? 1 : (want < cap ? want : cap));
    tss_launch(count_valid_kernel, grid, kThreads, 0, st, target, n, ignore_index, num_classes, (unsigned long long*)nvalid);
    TSS_LAUNCH_CHECK("ce_count_valid");
    return TSS_OK;
}

extern "C" int tss_ce_fwd(const void* logits, const int64_t* target, int N, int C, int64_t HW,
                          int64_t ignore_index, const int64_t* nvalid, double* loss_sum, float* pixel_loss,
                          void* dlogits, const float* ohem, int dtype, void* stream) {
    TSS_REQUIRE(N > 0 && HW > 0, "ce_fwd: empty input");
    TSS_REQUIRE(HW % 4 == 0, "ce_fwd: H*W=%lld must be a multiple of 4", (long long)HW);
    TSS_REQUIRE(dlogits == nullptr || nvalid != nullptr || ohem != nullptr, "ce_fwd: gradient needs nvalid (or OHEM weights)");
    cudaStream_t st = (cudaStream_t)stream;
    TSS_DISPATCH_DTYPE(dtype, "ce_fwd", {
        switch (C) {
            case 19: return launch_ce<T, 19>(logits, target, N, HW, ignore_index, nvalid, loss_sum, pixel_loss, dlogits, ohem, st);
            case 2: return launch_ce<T, 2>(logits, target, N, HW, ignore_index, nvalid, loss_sum, pixel_loss, dlogits, ohem, st);
            case 8: return launch_ce<T, 8>(logits, target, N, HW, ignore_index, nvalid, loss_sum, pixel_loss, dlogits, ohem, st);
            case 11: return launch_ce<T, 11>(logits, target, N, HW, ignore_index, nvalid, loss_sum, pixel_loss, dlogits, ohem, st);
            case 12: return launch_ce<T, 12>(logits, target, N, HW, ignore_index, nvalid, loss_sum, pixel_loss, dlogits, ohem, st);
            case 20: return launch_ce<T, 20>(logits, target, N, HW, ignore_index, nvalid, loss_sum, pixel_loss, dlogits, ohem, st);
            case 21: return launch_ce<T, 21>(logits, target, N, HW, ignore_index, nvalid, loss_sum, pixel_loss, dlogits, ohem, st);
            default:
                tss_set_error("ce_fwd: C=%d not instantiated (19 Cityscapes/BDD, 11/12 CamVid, 20/21 VOC, 2, 8)", C);
                return TSS_ERR_ARG;
        }
    });
}

extern "C" int tss_ce_finalize(const double* loss_sum, const int64_t* nvalid, float* loss, void* stream) {
    tss_launch(ce_finalize_kernel, 1, 1, 0, (cudaStream_t)stream, loss_sum, nvalid, loss);
    TSS_LAUNCH_CHECK("ce_finalize");
    return TSS_OK;
}

extern "C" int tss_upsample_ce_fwd(const void* x, const int64_t* target, int N, int C, int Hi, int Wi, int Ho, int Wo,
                                   int64_t ldx, int64_t ignore_index, double* loss_sum, int64_t* nvalid,
                                   float* pixel_loss, float* dx32, int64_t lddx, const float* ohem, int dtype,
                                   void* stream) {
    TSS_REQUIRE(N > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, "upsample_ce_fwd: empty input");
    TSS_REQUIRE(ldx >= C && (dx32 == nullptr || lddx >= C), "upsample_ce_fwd: pitch smaller than C=%d", C);
    TSS_REQUIRE(nvalid != nullptr, "upsample_ce_fwd: nvalid is required");
    cudaStream_t st = (cudaStream_t)stream;
    TSS_DISPATCH_DTYPE(dtype, "upsample_ce_fwd", {
        switch (C) {
            case 19: return launch_head<T, 19>(x, target, dx32, loss_sum, nvalid, pixel_loss, ohem, N, Hi, Wi, Ho, Wo, ldx, lddx, ignore_index, st);
            case 11: return launch_head<T, 11>(x, target, dx32, loss_sum, nvalid, pixel_loss, ohem, N, Hi, Wi, Ho, Wo, ldx, lddx, ignore_index, st);
            case 12: return launch_head<T, 12>(x, target, dx32, loss_sum, nvalid, pixel_loss, ohem, N, Hi, Wi, Ho, Wo, ldx, lddx, ignore_index, st);
            case 21: return launch_head<T, 21>(x, target, dx32, loss_sum, nvalid, pixel_loss, ohem, N, Hi, Wi, Ho, Wo, ldx, lddx, ignore_index, st);
            default:
                tss_set_error("upsample_ce_fwd: C=%d not instantiated (19 Cityscapes/BDD, 11/12 CamVid, 21 VOC)", C);
                return TSS_ERR_ARG;
        }
    });
}

extern "C" int tss_upsample_ce_finalize(const double* loss_sum, const int64_t* nvalid, float* loss, const float* dx32,
                                        void* dx, int64_t n, int dtype, void* stream) {
    TSS_REQUIRE(dx == nullptr || (dx32 != nullptr && n > 0 && n % 8 == 0), "upsample_ce_finalize: n=%lld must be a positive multiple of 8", (long long)n);
    TSS_DISPATCH_DTYPE(dtype, "upsample_ce_finalize", {
        int64_t want = dx != nullptr ? ceil_div64(n / 8, kThreads) : 1;
        const int64_t cap = (int64_t)tss_num_sms() * 8;
        const int grid = (int)(want < 1 ? 1 : (want < cap ? want : cap));
        tss_launch(upsample_ce_finalize_kernel<T>, grid, kThreads, 0, (cudaStream_t)stream, loss_sum, nvalid, loss, dx32, (T*)dx, n / 8);
        TSS_LAUNCH_CHECK("upsample_ce_finalize");
        return TSS_OK;
    });
}
