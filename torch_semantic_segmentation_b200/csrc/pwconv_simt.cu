// Pointwise (1x1) convolution as a GEMM -- SIMT fp32-accumulate implementation (impl 0).
//
// This is the reference-precision path: it serves the fp32 verification mode (fp32 parity
// at 1e-4 needs true fp32 products; tcgen05 kind::tf32 cannot give that) and any shape the
// tensor-core path does not cover.  The bf16 production path is pwconv_tc.cu (tcgen05).
//
//   fwd  : Y[M][Nc]  = X[M][K]   . W[Nc][K]^T   (+ affine / residual / ReLU / BN statistics)
//   dgrad: dX[M][K]  = dY[M][Nc] . W[Nc][K]
//   wgrad: dW[Nc][K] += dY[M][Nc]^T . X[M][K]   (split over M, fp32 atomics)
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 64, BK = 16, kThreads = 256;

__host__ __device__ inline int round_up8(int v) { return (v + 7) & ~7; }

template <typename T> struct Vec4Store;
template <> struct Vec4Store<float> {
    __device__ static void st(float* p, float a, float b, float c, float d) {
        *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
    }
    __device__ static void ld(const float* p, float (&v)[4]) {
        float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
};
template <> struct Vec4Store<bf16> {
    __device__ static void st(bf16* p, float a, float b, float c, float d) {
        uint2 u; u.x = pack_bf16x2(a, b); u.y = pack_bf16x2(c, d);
        *reinterpret_cast<uint2*>(p) = u;
    }
    __device__ static void ld(const bf16* p, float (&v)[4]) {
        uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
        v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
        v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
    }
};

// C[M][Nout] = A[M][Kred] . B ; B(j, r) = TRANS_B ? W[r*ldw + j] : W[j*ldw + r]
// (fwd: Nout=Nc, Kred=K, W[Nc][K], TRANS_B=false; dgrad: Nout=K, Kred=Nc, TRANS_B=true)
template <typename T, bool TRANS_B>
__global__ void __launch_bounds__(kThreads)
gemm_simt_kernel(const T* __restrict__ A, const float* __restrict__ W, T* __restrict__ Cout,
                 int64_t M, int Kred, int Nout, int64_t lda, int ldw, int64_t ldc, int Nstore,
                 const float* __restrict__ scale, const float* __restrict__ shift,
                 const T* __restrict__ res, int64_t ldr, int relu, double* __restrict__ stats) {
    pdl_wait();
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    __shared__ float s_col[2][BN];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int Kpad = round_up8(Kred);

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int a_row = tid >> 1, a_k = (tid & 1) * 8;
    for (int k0 = 0; k0 < Kpad; k0 += BK) {
        {   // A tile: 128 rows x 16 k, transposed into As[k][row]
            float v[8];
            const int64_t m = m0 + a_row;
            if (m < M && k0 + a_k < Kpad) load8(A + m * lda + k0 + a_k, v);
            else zero8(v);
#pragma unroll
            for (int e = 0; e < 8; ++e) As[a_k + e][a_row] = v[e];
        }
        if (!TRANS_B) {   // W[n][k]: 64 n x 16 k
            const int n = tid >> 2, kq = (tid & 3) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + n < Nout && k0 + kq < Kred) v = __ldg(reinterpret_cast<const float4*>(W + (int64_t)(n0 + n) * ldw + k0 + kq));
            Bs[kq + 0][n] = v.x; Bs[kq + 1][n] = v.y; Bs[kq + 2][n] = v.z; Bs[kq + 3][n] = v.w;
        } else {          // W[r][j]: 16 r x 64 j
            const int r = tid >> 4, jq = (tid & 15) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k0 + r < Kred && n0 + jq < Nout) v = __ldg(reinterpret_cast<const float4*>(W + (int64_t)(k0 + r) * ldw + n0 + jq));
            *reinterpret_cast<float4*>(&Bs[r][jq]) = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[8], b[4];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    const int nc = n0 + tx * 4;   // first of this thread's 4 output columns
    if (stats != nullptr) {       // BatchNorm statistics of the raw output (rows >= M are zero)
        if (tid < 2 * BN) (&s_col[0][0])[tid] = 0.f;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) { s1 += acc[i][j]; s2 = fmaf(acc[i][j], acc[i][j], s2); }
            atomicAdd(&s_col[0][tx * 4 + j], s1);
            atomicAdd(&s_col[1][tx * 4 + j], s2);
        }
        __syncthreads();
        if (tid < BN && n0 + tid < Nout) {
            atomicAdd(stats + n0 + tid, (double)s_col[0][tid]);
            atomicAdd(stats + Nout + n0 + tid, (double)s_col[1][tid]);
        }
    }

    if (nc >= Nstore) return;
    float sc[4], sh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const bool ok = nc + j < Nout;
        sc[j] = (ok && scale != nullptr) ? __ldg(scale + nc + j) : 1.f;
        sh[j] = (ok && shift != nullptr) ? __ldg(shift + nc + j) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t m = m0 + ty * 8 + i;
        if (m >= M) break;
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = fmaf(acc[i][j], sc[j], sh[j]);
        if (res != nullptr) {
            float r[4];
            Vec4Store<T>::ld(res + m * ldr + nc, r);
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] += r[j];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (relu) o[j] = fmaxf(o[j], 0.f);
            if (nc + j >= Nout) o[j] = 0.f;     // keep the pad columns zero
        }
        Vec4Store<T>::st(Cout + m * ldc + nc, o[0], o[1], o[2], o[3]);
    }
}

// dW[Nc][K] += sum_m dY[m][n] X[m][k].  CTA tile 64(n) x 64(k), rows [blockIdx.z*rows_per, +rows_per)
constexpr int WT = 64, WK = 16;
template <typename T>
__global__ void __launch_bounds__(kThreads)
wgrad_simt_kernel(const T* __restrict__ X, const T* __restrict__ dY, float* __restrict__ dW,
                  float* __restrict__ db, int64_t M, int K, int Nc, int64_t ldx, int64_t lddy,
                  int64_t rows_per) {
    pdl_wait();
    __shared__ __align__(16) float Gs[WK][WT];   // dY rows
    __shared__ __align__(16) float Xs[WK][WT];   // X rows
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;      // ty -> n, tx -> k
    const int n0 = blockIdx.x * WT, k0 = blockIdx.y * WT;
    const int64_t mb = (int64_t)blockIdx.z * rows_per;
    const int64_t me = (mb + rows_per < M) ? mb + rows_per : M;
    const int Npad = round_up8(Nc);

    float acc[4][4];
    float bsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int lr = (tid & 127) >> 3, lc = (tid & 7) * 8;   // loader: row within 16, 8-col group
    for (int64_t m = mb; m < me; m += WK) {
        float v[8];
        if (tid < 128) {
            if (m + lr < me && n0 + lc < Npad) load8(dY + (m + lr) * lddy + n0 + lc, v);
            else zero8(v);
            *reinterpret_cast<float4*>(&Gs[lr][lc]) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(&Gs[lr][lc + 4]) = make_float4(v[4], v[5], v[6], v[7]);
        } else {
            if (m + lr < me && k0 + lc < K) load8(X + (m + lr) * ldx + k0 + lc, v);
            else zero8(v);
            *reinterpret_cast<float4*>(&Xs[lr][lc]) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(&Xs[lr][lc + 4]) = make_float4(v[4], v[5], v[6], v[7]);
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < WK; ++r) {
            const float4 g = *reinterpret_cast<const float4*>(&Gs[r][ty * 4]);
            const float4 x = *reinterpret_cast<const float4*>(&Xs[r][tx * 4]);
            const float ga[4] = {g.x, g.y, g.z, g.w};
            const float xa[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                bsum[i] += ga[i];
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ga[i], xa[j], acc[i][j]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n0 + ty * 4 + i;
        if (n >= Nc) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + tx * 4 + j;
            if (k < K) atomicAdd(dW + (int64_t)n * K + k, acc[i][j]);
        }
        if (db != nullptr && blockIdx.y == 0 && tx == 0) atomicAdd(db + n, bsum[i]);
    }
}

__global__ void pack_weights_kernel(const float* __restrict__ w, bf16* __restrict__ wp, bf16* __restrict__ wpT,
                                    int Nc, int K) {
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Nc * K) return;
    const int n = i / K, k = i - n * K;
    const bf16 v = __float2bfloat16_rn(w[i]);
    if (wp != nullptr) wp[i] = v;
    if (wpT != nullptr) wpT[(int64_t)k * Nc + n] = v;
}

// all pointwise weights of a model in ONE launch: blockIdx.y = table row {offset of the fp32 (Nc,K)
// weight in the parameter arena, Nc, K, wp pointer, wpT pointer}
__global__ void pack_weights_multi_kernel(const float* __restrict__ arena, const int64_t* __restrict__ table) {
    pdl_wait();
    const int64_t* row = table + (int64_t)blockIdx.y * 5;
    const float* w = arena + row[0];
    const int Nc = (int)row[1], K = (int)row[2];
    bf16* wp = reinterpret_cast<bf16*>(row[3]);
    bf16* wpT = reinterpret_cast<bf16*>(row[4]);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Nc * K; i += gridDim.x * blockDim.x) {
        const int n = i / K, k = i - n * K;
        const bf16 v = __float2bfloat16_rn(w[i]);
        wp[i] = v;
        wpT[(int64_t)k * Nc + n] = v;
    }
}

// Class-score convolution nn.Conv2d(K, Nc, 1) with bias and Nc not a multiple of 16 (19 classes; fastscnn.py:97) on the
// tensor-core GEMMs: zero-padded bf16 operands wp[Np][K] (forward B operand), wpT[K][Npt] (dgrad B operand, Npt = the
// gradient's channel pitch) and the bias as a padded fp32 shift vector.  One tiny launch per forward.
__global__ void class_pack_kernel(const float* __restrict__ w, const float* __restrict__ bias, bf16* __restrict__ wp,
                                  bf16* __restrict__ wpT, float* __restrict__ bias_pad, int Nc, int K, int Np, int Npt) {
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < Np * K) {
        const int n = i / K, k = i - n * K;
        wp[i] = __float2bfloat16_rn(n < Nc ? w[n * K + k] : 0.f);
    }
    if (i < K * Npt) {
        const int k = i / Npt, n = i - k * Npt;
        wpT[i] = __float2bfloat16_rn(n < Nc ? w[n * K + k] : 0.f);
    }
    if (i < Np) bias_pad[i] = (bias != nullptr && i < Nc) ? bias[i] : 0.f;
}

// out[c] += sum over rows of x[m][c]  (bias gradient of the class-score conv; C <= 32, x row pitch ld)
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ x, float* __restrict__ out, int64_t M, int C, int64_t ld) {
    pdl_wait();
    __shared__ float s_sum[32];
    if (threadIdx.x < 32) s_sum[threadIdx.x] = 0.f;
    __syncthreads();
    const int c = threadIdx.x & 31;              // lane = channel: a warp reads one row (<= 64 contiguous bytes)
    const int r = threadIdx.x >> 5;
    float acc = 0.f;
    if (c < C)
        for (int64_t m = (int64_t)blockIdx.x * 8 + r; m < M; m += (int64_t)gridDim.x * 8) acc += to_f32(x[m * ld + c]);
    if (c < C) atomicAdd(&s_sum[c], acc);
    __syncthreads();
    if (threadIdx.x < C) atomicAdd(out + threadIdx.x, s_sum[threadIdx.x]);
}

}  // namespace

// implemented in pwconv_tc.cu
int tss_pwconv_fwd_tc(const void* x, const void* wp, void* y, int64_t M, int K, int Nc, int64_t ldx,
                      int64_t ldy, const float* scale, const float* shift, const void* res, int64_t ldr,
                      int flags, double* stats, cudaStream_t st);

int tss_pwconv_wgrad_tc(const void* x, const void* dy, float* dw, int64_t M, int K, int Nc, int64_t ldx,
                        int64_t lddy, cudaStream_t st);

static int check_gemm(const char* name, int64_t M, int K, int Nc, int64_t lda, int64_t ldc, int red, int out) {
    TSS_REQUIRE(M > 0 && K > 0 && Nc > 0, "%s: empty problem M=%lld K=%d Nc=%d", name, (long long)M, K, Nc);
    TSS_REQUIRE(K % 8 == 0, "%s: K=%d must be a multiple of 8", name, K);
    TSS_REQUIRE(lda >= round_up8(red) && lda % 8 == 0, "%s: input pitch %lld must be a multiple of 8 and >= %d",
                name, (long long)lda, round_up8(red));
    TSS_REQUIRE(ldc % 4 == 0 && (out % 4 == 0 || ldc >= round_up8(out)),
                "%s: output pitch %lld incompatible with %d columns", name, (long long)ldc, out);
    return TSS_OK;
}

extern "C" int tss_pwconv_fwd(const void* x, const float* w, const void* wp, void* y, int64_t M, int K, int Nc,
                              int64_t ldx, int64_t ldy, const float* scale, const float* shift,
                              const void* res, int64_t ldr, int flags, double* stats, int impl,
                              int dtype, void* stream) {
    if (int e = check_gemm("pwconv_fwd", M, K, Nc, ldx, ldy, K, Nc)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    if (impl == 1) {
        TSS_REQUIRE(dtype == TSS_BF16 && wp != nullptr, "pwconv_fwd: impl 1 needs bf16 activations and packed weights");
        return tss_pwconv_fwd_tc(x, wp, y, M, K, Nc, ldx, ldy, scale, shift, res, ldr, flags, stats, st);
    }
    TSS_REQUIRE(impl == 0, "pwconv_fwd: unknown impl %d", impl);
    const int Nstore = (Nc % 4 == 0) ? Nc : (int)(ldy < round_up8(Nc) ? ldy : round_up8(Nc));
    dim3 grid((unsigned)ceil_div64(M, BM), (unsigned)((Nstore + BN - 1) / BN));
    TSS_DISPATCH_DTYPE(dtype, "pwconv_fwd", {
        tss_launch(gemm_simt_kernel<T, false>, grid, kThreads, 0, st, 
            (const T*)x, w, (T*)y, M, K, Nc, ldx, K, ldy, Nstore, scale, shift, (const T*)res, ldr,
            flags & TSS_EPI_RELU, stats);
        TSS_LAUNCH_CHECK("pwconv_fwd");
        return TSS_OK;
    });
}

extern "C" int tss_pwconv_dgrad(const void* dy, const float* w, const void* wpT, void* dx, int64_t M, int K,
                                int Nc, int64_t lddy, int64_t lddx, int impl, int dtype, void* stream) {
    if (int e = check_gemm("pwconv_dgrad", M, K, Nc, lddy, lddx, Nc, K)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    if (impl == 1) {
        TSS_REQUIRE(dtype == TSS_BF16 && wpT != nullptr, "pwconv_dgrad: impl 1 needs bf16 activations and packed weights");
        TSS_REQUIRE(Nc % 8 == 0, "pwconv_dgrad: impl 1 needs Nc %% 8 == 0 (got %d)", Nc);
        // dX[M][K] = dY[M][Nc] . (W^T)[K][Nc]^T : the forward kernel with the transposed pack
        return tss_pwconv_fwd_tc(dy, wpT, dx, M, Nc, K, lddy, lddx, nullptr, nullptr, nullptr, 0, 0, nullptr, st);
    }
    TSS_REQUIRE(impl == 0, "pwconv_dgrad: unknown impl %d", impl);
    dim3 grid((unsigned)ceil_div64(M, BM), (unsigned)((K + BN - 1) / BN));
    TSS_DISPATCH_DTYPE(dtype, "pwconv_dgrad", {
        tss_launch(gemm_simt_kernel<T, true>, grid, kThreads, 0, st, 
            (const T*)dy, w, (T*)dx, M, Nc, K, lddy, K, lddx, K, nullptr, nullptr, (const T*)nullptr, 0, 0, nullptr);
        TSS_LAUNCH_CHECK("pwconv_dgrad");
        return TSS_OK;
    });
}

extern "C" int tss_pwconv_wgrad(const void* x, const void* dy, float* dw, float* db, int64_t M, int K, int Nc,
                                int64_t ldx, int64_t lddy, int impl, int dtype, void* stream) {
    TSS_REQUIRE(M > 0 && K > 0 && Nc > 0, "pwconv_wgrad: empty problem");
    TSS_REQUIRE(K % 8 == 0 && ldx % 8 == 0 && ldx >= K, "pwconv_wgrad: K=%d ldx=%lld", K, (long long)ldx);
    TSS_REQUIRE(lddy % 8 == 0 && lddy >= round_up8(Nc), "pwconv_wgrad: lddy=%lld Nc=%d", (long long)lddy, Nc);
    TSS_REQUIRE(impl == 0 || impl == 1, "pwconv_wgrad: unknown impl %d", impl);
    cudaStream_t st = (cudaStream_t)stream;
    if (impl == 1) {
        TSS_REQUIRE(dtype == TSS_BF16, "pwconv_wgrad: impl 1 needs bf16 activations");
        if (int e = tss_pwconv_wgrad_tc(x, dy, dw, M, K, Nc, ldx, lddy, st)) return e;
        if (db != nullptr) {
            TSS_REQUIRE(Nc <= 32, "pwconv_wgrad: impl 1 bias gradient needs Nc <= 32 (got %d)", Nc);
            int64_t grid = ceil_div64(M, 8 * 16);
            const int64_t cap = (int64_t)tss_num_sms() * 4;
            if (grid > cap) grid = cap;
            tss_launch(colsum_kernel<bf16>, (unsigned)grid, 256, 0, st, (const bf16*)dy, db, M, Nc, lddy);
            TSS_LAUNCH_CHECK("pwconv_wgrad(bias)");
        }
        return TSS_OK;
    }
    const int gx = (Nc + WT - 1) / WT, gy = (K + WT - 1) / WT;
    int64_t target = (int64_t)tss_num_sms() * 4;
    int64_t nsplit = target / ((int64_t)gx * gy);
    if (nsplit < 1) nsplit = 1;
    int64_t rows_per = ceil_div64(ceil_div64(M, nsplit), WK) * WK;
    if (rows_per < 64) rows_per = 64;
    nsplit = ceil_div64(M, rows_per);
    dim3 grid(gx, gy, (unsigned)nsplit);
    TSS_DISPATCH_DTYPE(dtype, "pwconv_wgrad", {
        tss_launch(wgrad_simt_kernel<T>, grid, kThreads, 0, st, (const T*)x, (const T*)dy, dw, db, M, K, Nc, ldx, lddy, rows_per);
        TSS_LAUNCH_CHECK("pwconv_wgrad");
        return TSS_OK;
    });
}

extern "C" int tss_pack_weights_multi(const float* arena, const int64_t* table, int n_entries, int64_t max_elems,
                                      void* stream) {
    TSS_REQUIRE(n_entries > 0 && max_elems > 0, "pack_weights_multi: n_entries=%d max_elems=%lld", n_entries, (long long)max_elems);
    int64_t gx = ceil_div64(max_elems, 256);
    if (gx > 64) gx = 64;
    tss_launch(pack_weights_multi_kernel, dim3((unsigned)gx, (unsigned)n_entries), 256, 0, (cudaStream_t)stream, arena, table);
    TSS_LAUNCH_CHECK("pack_weights_multi");
    return TSS_OK;
}

extern "C" int tss_class_scores_pack(const float* w, const float* bias, void* wp, void* wpT, float* bias_pad, int Nc,
                                     int K, int Np, int Npt, void* stream) {
    TSS_REQUIRE(Nc > 0 && K > 0 && Np >= Nc && Np % 16 == 0 && Npt >= Nc && Npt % 8 == 0,
                "class_scores_pack: Nc=%d K=%d Np=%d Npt=%d", Nc, K, Np, Npt);
    TSS_REQUIRE(w != nullptr && wp != nullptr && wpT != nullptr && bias_pad != nullptr, "class_scores_pack: missing buffer");
    const int n = (Np > Npt ? Np : Npt) * K;
    tss_launch(class_pack_kernel, (n + 255) / 256, 256, 0, (cudaStream_t)stream, w, bias, (bf16*)wp, (bf16*)wpT, bias_pad, Nc,
               K, Np, Npt);
    TSS_LAUNCH_CHECK("class_scores_pack");
    return TSS_OK;
}

extern "C" int tss_pack_weights_bf16(const float* w, void* wp, void* wpT, int Nc, int K, void* stream) {
    TSS_REQUIRE(Nc > 0 && K > 0, "pack_weights_bf16: Nc=%d K=%d", Nc, K);
    tss_launch(pack_weights_kernel, (Nc * K + 255) / 256, 256, 0, (cudaStream_t)stream, w, (bf16*)wp, (bf16*)wpT, Nc, K);
    TSS_LAUNCH_CHECK("pack_weights_bf16");
    return TSS_OK;
}
