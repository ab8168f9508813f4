// Stem convolution (3 -> 32, 3x3, stride 2, padding 1; fastscnn.py:30, contextnet.py:38,48) on tcgen05, bf16.
// The SIMT kernel (stem.cu) spends 864 FMAs per output pixel on the CUDA cores and runs at ~5x its HBM time.
// Here the 27-tap patch of each output pixel becomes one row of an implicit-GEMM A operand that the threads
// build themselves: 128 consecutive output pixels of one row = the 128 rows of a UMMA tile, K = 27 padded to 32
// (two K=16 instructions), N = 32 output channels.
//
//   threads 0..127   build their pixel's row from the NCHW fp32 image (27 loads, zero padding by predicate),
//                    convert to bf16 and store it K-major / 128-byte swizzled (chunk j of row r at j ^ (r & 7)),
//                    fence.proxy.async; after the MMA they read their row of the TMEM accumulator back
//                    (tcgen05.ld), accumulate the BatchNorm statistics (or apply scale/shift/ReLU in eval mode)
//                    and store 32 bf16 channels (64 contiguous bytes per pixel).
//   warp 4           allocates TMEM (32 columns) and issues the two tcgen05.mma per tile from one lane.
//   weights          (32,3,3,3) fp32 -> bf16 [32][32] K-major swizzled tile, built once per CTA by the threads.
// No TMA: both operands are thread-built.  A CTA walks kTilesPerCta tiles with one A buffer and one TMEM
// accumulator; overlap comes from ~10 co-resident CTAs per SM (20 KB of shared memory, 32 TMEM columns each).
// Algorithmic bytes: 12 B per input pixel + 64 B per output pixel, as the SIMT kernel.
#include "tc_ptx.cuh"

namespace {

constexpr int CO = 32;
constexpr int kTaps = 27;
constexpr int kRows = 128;                    // output pixels per tile
constexpr int kThreads = 160;
constexpr int kTilesPerCta = 4;
constexpr uint32_t kABytes = kRows * 128;     // 64 bf16 per swizzled row, only the first 32 are used
constexpr uint32_t kBBytes = CO * 128;

__global__ void __launch_bounds__(kThreads)
stem_tc_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, bf16* __restrict__ y, int N, int H, int W,
                   int Ho, int Wo, int tiles_w, int64_t ntiles, const float* __restrict__ scale,
                   const float* __restrict__ shift, int relu, double* __restrict__ stats) {
    TSS_DYN_SMEM(uint8_t, smem_raw);
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = sA + kABytes;
    uint64_t* mma_done = (uint64_t*)(sB + kBBytes);
    uint32_t* tmem_slot = (uint32_t*)(mma_done + 1);
    float* s_stat = (float*)(tmem_slot + 2);          // [2][CO]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(smem_u32(mma_done), 1);
        mbar_init_fence();
    }
    if (warp == 4) tc_alloc(smem_u32(tmem_slot), 32);
    if (threadIdx.x < 2 * CO) s_stat[threadIdx.x] = 0.f;
    pdl_wait();
    // weights: B[co][t] = w[co][ci][ky][kx], t = ci*9 + ky*3 + kx, zero for t >= 27
    for (int i = threadIdx.x; i < CO * 32; i += kThreads) {
        const int co = i >> 5, t = i & 31;
        const bf16 v = __float2bfloat16_rn(t < kTaps ? __ldg(w + co * kTaps + t) : 0.f);
        *reinterpret_cast<bf16*>(sB + co * 128 + (((t >> 3) ^ (co & 7)) << 4) + (t & 7) * 2) = v;
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(CO >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24);

    for (int it = 0; it < kTilesPerCta; ++it) {
        const int64_t tile = (int64_t)blockIdx.x * kTilesPerCta + it;
        if (tile >= ntiles) break;                                       // uniform over the CTA
        const int tw = (int)(tile % tiles_w);
        const int ho = (int)((tile / tiles_w) % Ho);
        const int n = (int)(tile / ((int64_t)tiles_w * Ho));
        const int wo = tw * kRows + (int)threadIdx.x;                    // producers: threadIdx.x < 128
        if (threadIdx.x < kRows) {
            const int r = threadIdx.x;
            float v[32];
#pragma unroll
            for (int t = kTaps; t < 32; ++t) v[t] = 0.f;
#pragma unroll
            for (int ci = 0; ci < 3; ++ci)
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const int hi = 2 * ho - 1 + ky;
                    const bool row_ok = wo < Wo && hi >= 0 && hi < H;
                    const float* xr = x + (((int64_t)n * 3 + ci) * H + (row_ok ? hi : 0)) * W;
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const int wi = 2 * wo - 1 + kx;
                        v[ci * 9 + ky * 3 + kx] = (row_ok && wi >= 0 && wi < W) ? __ldg(xr + wi) : 0.f;
                    }
                }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint4 u;
                u.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]); u.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
                u.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]); u.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
                *reinterpret_cast<uint4*>(sA + r * 128 + ((j ^ (r & 7)) << 4)) = u;
            }
            fence_async_smem();
        }
        tc_fence_before();
        __syncthreads();                                                 // the A tile is complete
        if (warp == 4 && lane == 0) {
            tc_fence_after();
            const uint64_t adesc = make_desc_k_sw128(smem_u32(sA)), bdesc = make_desc_k_sw128(smem_u32(sB));
            umma_bf16(tmem_base, adesc, bdesc, idesc, 0u);
            umma_bf16(tmem_base, adesc + 2, bdesc + 2, idesc, 1u);
            umma_commit(smem_u32(mma_done));
        }
        if (threadIdx.x < kRows) {
            mbar_wait(smem_u32(mma_done), (uint32_t)(it & 1));
            tc_fence_after();
            float acc[CO];
            {
                float a[16], b[16];
                tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16), a);
                tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + 16u, b);
#pragma unroll
                for (int c = 0; c < 16; ++c) { acc[c] = a[c]; acc[16 + c] = b[c]; }
            }
            const bool valid = wo < Wo;
            if (valid) {
                bf16* yp = y + (((int64_t)n * Ho + ho) * Wo + wo) * CO;
#pragma unroll
                for (int c8 = 0; c8 < CO / 8; ++c8) {
                    float o[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        float t = acc[c8 * 8 + e];
                        if (shift != nullptr) t = fmaf(t, scale != nullptr ? __ldg(scale + c8 * 8 + e) : 1.f, __ldg(shift + c8 * 8 + e));
                        if (relu) t = fmaxf(t, 0.f);
                        o[e] = t;
                    }
                    store8(yp + c8 * 8, o);
                }
            }
            if (stats != nullptr) {                                      // raw (pre-affine) output statistics
                float sq[CO];
#pragma unroll
                for (int c = 0; c < CO; ++c) {
                    if (!valid) acc[c] = 0.f;
                    sq[c] = acc[c] * acc[c];
                }
                const float s1 = warp_transpose_sum32(acc, lane);
                const float s2 = warp_transpose_sum32(sq, lane);
                atomicAdd(&s_stat[lane], s1);
                atomicAdd(&s_stat[CO + lane], s2);
            }
            tc_fence_before();
        }
        __syncthreads();                                                 // accumulator and A tile are free again
    }
    if (stats != nullptr && threadIdx.x < 2 * CO) atomicAdd(stats + threadIdx.x, (double)s_stat[threadIdx.x]);
    if (warp == 4) {
        __syncwarp();
        tc_fence_after();
        tc_dealloc(tmem_base, 32);
    }
}

// ------------------------------------------------------------------ weight gradient on tcgen05
// dw[co][t] += sum over output pixels p of dy[p][co] * patch[p][t]: a GEMM whose reduction dimension is the pixel
// index.  Both operands are built K-major by the threads, 64 pixels (= one 128-byte swizzle row) per k-block:
//   A[m = co][k = pixel]   dy transposed on the way into shared memory (rows 32..127 of the M = 128 tile stay zero)
//   B[n = t ][k = pixel]   the 27 taps of each pixel from the fp32 image (rows 27..31 zero)
// and one TMEM accumulator per CTA collects all k-blocks of the CTA's share (persistent grid, 2-stage ring:
// producers -> full[s] -> 4 x tcgen05.mma (K = 16) -> commit -> empty[s]).  At the end warp 0 reads rows 0..31
// of the accumulator (lane = output channel) and adds its 27 values to dw with red.global.add.f32.
constexpr int kWgThreads = 160;
constexpr int kWgStages = 2;
constexpr int kWgPix = 64;                      // pixels per k-block

__global__ void __launch_bounds__(kWgThreads)
stem_tc_wgrad_kernel(const float* __restrict__ x, const bf16* __restrict__ dy, float* __restrict__ dw, int N, int H, int W,
                     int Ho, int Wo, int chunks_w, int64_t nblocks, const bf16* __restrict__ yraw,
                     const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                     const float* __restrict__ beta, const float* __restrict__ sums, int relu, float inv_count,
                     float* __restrict__ dgamma, float* __restrict__ dbeta) {
    TSS_DYN_SMEM(uint8_t, smem_raw);
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                                   // [kWgStages][128 rows][128 B]
    uint8_t* sB = sA + (size_t)kWgStages * kABytes;        // [kWgStages][32 rows][128 B]
    uint64_t* full = (uint64_t*)(sB + (size_t)kWgStages * kBBytes);
    uint64_t* empty = full + kWgStages;
    uint64_t* done = empty + kWgStages;
    uint32_t* tmem_slot = (uint32_t*)(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kWgStages; ++s) {
            mbar_init(smem_u32(full + s), 128);
            mbar_init(smem_u32(empty + s), 1);
        }
        mbar_init(smem_u32(done), 1);
        mbar_init_fence();
    }
    if (warp == 4) tc_alloc(smem_u32(tmem_slot), 32);
    // rows that are never rewritten: A rows 32..127 and B rows 27..31 are zero in both stages
    for (int i = threadIdx.x; i < kWgStages * (int)(kABytes + kBBytes) / 16; i += kWgThreads)
        reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();
    const uint32_t tmem_base = *tmem_slot;
    // yraw != nullptr: `dy` holds the gradient AFTER the stem's BatchNorm/ReLU and the BatchNorm-backward apply happens
    // here, while the operand is built: dy_c = A*g + B*y + D per output channel (pwconv_tc_bwd.cu), SH for the mask
    float cA[2] = {1.f, 1.f}, cSH[2] = {0.f, 0.f}, cB[2] = {0.f, 0.f}, cD[2] = {0.f, 0.f};
    if (yraw != nullptr && warp < 4) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = ((int)threadIdx.x + 128 * h) >> 3;      // the output channel of this thread's chunk task h
            const float mu = __ldg(mean + c), rs = __ldg(rstd + c);
            cA[h] = (gamma != nullptr ? __ldg(gamma + c) : 1.f) * rs;
            cSH[h] = (beta != nullptr ? __ldg(beta + c) : 0.f) - mu * cA[h];
            const float a1 = __ldg(sums + c), a2 = __ldg(sums + CO + c);
            const float k2 = cA[h] * a2 * inv_count;
            cB[h] = -rs * k2;
            cD[h] = fmaf(mu * rs, k2, -cA[h] * a1 * inv_count);
            if (blockIdx.x == 0 && (threadIdx.x & 7) == 0) {      // one thread per channel, first CTA only
                if (dbeta != nullptr) dbeta[c] += a1;
                if (dgamma != nullptr) dgamma[c] += a2;
            }
        }
    }
    // this CTA's contiguous share of the k-blocks
    const int64_t per = (nblocks + gridDim.x - 1) / gridDim.x;
    const int64_t kb0 = (int64_t)blockIdx.x * per;
    const int64_t kb1 = kb0 + per < nblocks ? kb0 + per : nblocks;
    const int nkb = kb1 > kb0 ? (int)(kb1 - kb0) : 0;

    if (warp < 4) {                                        // ---------------- producers
        const int tt = threadIdx.x;                        // 0..127
        for (int i = 0; i < nkb; ++i) {
            const int s = i % kWgStages;
            const uint32_t phase = (i / kWgStages) & 1;
            const int64_t kb = kb0 + i;
            const int cw = (int)(kb % chunks_w);
            const int ho = (int)((kb / chunks_w) % Ho);
            const int n = (int)(kb / ((int64_t)chunks_w * Ho));
            const int wo0 = cw * kWgPix;
            uint4 a_chunk[2], b_chunk[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int task = tt + 128 * h;             // 256 chunk tasks per operand: (row, 8-pixel chunk)
                const int row = task >> 3, pc = task & 7;
                // A: row = output channel, 8 consecutive pixels of dy[..][row]
                uint32_t pk[4];
#pragma unroll
                for (int e = 0; e < 8; e += 2) {
                    const int w0 = wo0 + pc * 8 + e;
                    const int64_t off = (((int64_t)n * Ho + ho) * Wo + w0) * CO + row;
                    uint16_t lo = w0 < Wo ? __ldg(reinterpret_cast<const unsigned short*>(dy + off)) : (uint16_t)0;
                    uint16_t hi = w0 + 1 < Wo ? __ldg(reinterpret_cast<const unsigned short*>(dy + off + CO)) : (uint16_t)0;
                    if (yraw != nullptr) {
                        float v[2];
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const bool in = w0 + q < Wo;
                            const float g = __uint_as_float((uint32_t)(q ? hi : lo) << 16);
                            const float yy = in ? __uint_as_float((uint32_t)__ldg(reinterpret_cast<const unsigned short*>(yraw + off + q * CO)) << 16) : 0.f;
                            const float gg = (relu && !(fmaf(yy, cA[h], cSH[h]) > 0.f)) ? 0.f : g;
                            v[q] = in ? fmaf(cA[h], gg, fmaf(cB[h], yy, cD[h])) : 0.f;
                        }
                        const uint32_t p2 = pack_bf16x2(v[0], v[1]);
                        lo = (uint16_t)(p2 & 0xffffu);
                        hi = (uint16_t)(p2 >> 16);
                    }
                    pk[e >> 1] = (uint32_t)lo | ((uint32_t)hi << 16);
                }
                a_chunk[h] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                // B: row = tap t (27 real rows), the same 8 pixels of the image patch
                float v[8];
                if (row < kTaps) {
                    const int ci = row / 9, ky = (row % 9) / 3, kx = row % 3;
                    const int hi_ = 2 * ho - 1 + ky;
                    const bool row_ok = hi_ >= 0 && hi_ < H;
                    const float* xr = x + (((int64_t)n * 3 + ci) * H + (row_ok ? hi_ : 0)) * W;
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int wo = wo0 + pc * 8 + e, wi = 2 * wo - 1 + kx;
                        v[e] = (row_ok && wo < Wo && wi >= 0 && wi < W) ? __ldg(xr + wi) : 0.f;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = 0.f;
                }
                b_chunk[h] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
            }
            mbar_wait(smem_u32(empty + s), phase ^ 1);     // the MMAs that read stage s last time have retired
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int task = tt + 128 * h;
                const int row = task >> 3, pc = task & 7;
                *reinterpret_cast<uint4*>(sA + (size_t)s * kABytes + row * 128 + ((pc ^ (row & 7)) << 4)) = a_chunk[h];
                *reinterpret_cast<uint4*>(sB + (size_t)s * kBBytes + row * 128 + ((pc ^ (row & 7)) << 4)) = b_chunk[h];
            }
            fence_async_smem();
            mbar_arrive(smem_u32(full + s));
        }
    } else if (lane == 0) {                                // ---------------- MMA issuer (warp 4, one lane)
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int i = 0; i < nkb; ++i) {
            const int s = i % kWgStages;
            const uint32_t phase = (i / kWgStages) & 1;
            mbar_wait(smem_u32(full + s), phase);
            tc_fence_after();
            const uint64_t adesc = make_desc_k_sw128(smem_u32(sA + (size_t)s * kABytes));
            const uint64_t bdesc = make_desc_k_sw128(smem_u32(sB + (size_t)s * kBBytes));
#pragma unroll
            for (int k = 0; k < kWgPix / 16; ++k) umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (uint32_t)(i > 0 || k > 0));
            umma_commit(smem_u32(empty + s));
        }
        umma_commit(smem_u32(done));
    }
    if (warp == 0 && nkb > 0) {                            // ---------------- epilogue: rows 0..31 = output channels
        mbar_wait(smem_u32(done), 0);
        tc_fence_after();
        float a[16], b[16];
        tmem_ld16(tmem_base, a);
        tmem_ld16(tmem_base + 16u, b);
        float* d = dw + lane * kTaps;
#pragma unroll
        for (int t = 0; t < 16; ++t) atomicAdd(d + t, a[t]);
#pragma unroll
        for (int t = 16; t < kTaps; ++t) atomicAdd(d + t, b[t - 16]);
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 4) {
        __syncwarp();
        tc_fence_after();
        tc_dealloc(tmem_base, 32);
    }
}

// ------------------------------------------------------------------ weight gradient through a patch matrix -------
// The thread-built operands above make the weight gradient the slowest kernel of the step per byte (2-byte strided loads
// of dy for the A operand: 200-250 us for 198 MB).  Alternative: write the 27-tap patches once as a dense bf16 matrix
// P[pixel][32] (taps 27..31 zero) with fully coalesced traffic, then dW = dy^T . P is exactly the pointwise weight
// gradient, whose tensor-core kernel takes BOTH operands from HBM with TMA as they sit there (pwconv_tc.cu:
// wgrad_tc_kernel, MN-major boxes).  +226 MB of traffic (P written and read once), all of it at streaming speed.
constexpr int kPatchPix = 128;

__global__ void __launch_bounds__(kPatchPix)
stem_patches_kernel(const float* __restrict__ x, bf16* __restrict__ P, int H, int W, int Ho, int Wo, int tiles_w) {
    __shared__ float s_x[9][2 * kPatchPix + 2];          // rows (ci, ky); input columns 2*wo0-1 .. 2*wo0+256
    pdl_wait();
    const int tw = blockIdx.x % tiles_w;
    const int ho = (blockIdx.x / tiles_w) % Ho;
    const int n = blockIdx.x / (tiles_w * Ho);
    const int wo0 = tw * kPatchPix, wi0 = 2 * wo0 - 1;
    constexpr int kSpan = 2 * kPatchPix + 1;              // 257 input columns feed 128 output pixels
    for (int idx = threadIdx.x; idx < 9 * kSpan; idx += kPatchPix) {
        const int r = idx / kSpan, j = idx - r * kSpan;
        const int ci = r / 3, ky = r - ci * 3;
        const int hi = 2 * ho - 1 + ky, wi = wi0 + j;
        float v = 0.f;
        if (hi >= 0 && hi < H && wi >= 0 && wi < W) v = __ldg(x + (((int64_t)n * 3 + ci) * H + hi) * W + wi);
        s_x[r][j] = v;
    }
    __syncthreads();
    const int wo = wo0 + (int)threadIdx.x;
    if (wo >= Wo) return;
    float v[32];
#pragma unroll
    for (int t = kTaps; t < 32; ++t) v[t] = 0.f;
#pragma unroll
    for (int r = 0; r < 9; ++r)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) v[r * 3 + kx] = s_x[r][2 * threadIdx.x + kx];
    uint4* dst = reinterpret_cast<uint4*>(P + (((int64_t)n * Ho + ho) * Wo + wo) * 32);
#pragma unroll
    for (int j = 0; j < 4; ++j)
        dst[j] = make_uint4(pack_bf16x2(v[8 * j + 0], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                            pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
}

// dw[co][t] += dw32[co][t], t < 27  (the GEMM's [32][32] result -> the (32,3,3,3) parameter gradient)
__global__ void stem_wgrad_unpack_kernel(const float* __restrict__ dw32, float* __restrict__ dw) {
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < CO * kTaps) dw[i] += dw32[(i / kTaps) * 32 + (i % kTaps)];
}

}  // namespace

// pwconv_tc.cu
int tss_pwconv_wgrad_tc(const void* x, const void* dy, float* dw, int64_t M, int K, int Nc, int64_t ldx,
                        int64_t lddy, cudaStream_t st);

extern "C" int tss_stem3x3s2_patches(const float* x, void* patches, int N, int H, int W, void* stream) {
    TSS_REQUIRE(N > 0 && H > 0 && W > 0, "stem3x3s2_patches: empty input");
    TSS_REQUIRE(patches != nullptr && ((uintptr_t)patches & 15) == 0, "stem3x3s2_patches: patches must be 16-byte aligned");
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const int tiles_w = (Wo + kPatchPix - 1) / kPatchPix;
    TSS_REQUIRE((int64_t)N * Ho * tiles_w < (1ll << 31), "stem3x3s2_patches: grid too large");
    tss_launch(stem_patches_kernel, (unsigned)((int64_t)N * Ho * tiles_w), kPatchPix, 0, (cudaStream_t)stream, x, (bf16*)patches, H, W,
               Ho, Wo, tiles_w);
    TSS_LAUNCH_CHECK("stem3x3s2_patches");
    return TSS_OK;
}

extern "C" int tss_stem3x3s2_wgrad_from_patches(const void* patches, const void* dy, float* dw32, float* dw, int64_t M, int Cout,
                                                void* stream) {
    TSS_REQUIRE(M > 0, "stem3x3s2_wgrad_from_patches: empty input");
    TSS_REQUIRE(Cout == CO, "stem3x3s2_wgrad_from_patches: Cout=%d unsupported (only %d)", Cout, CO);
    TSS_REQUIRE(patches != nullptr && dw32 != nullptr && dw != nullptr, "stem3x3s2_wgrad_from_patches: missing buffer");
    TSS_REQUIRE((((uintptr_t)patches | (uintptr_t)dw32 | (uintptr_t)dy) & 15) == 0, "stem3x3s2_wgrad_from_patches: buffers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    TSS_CUDA(cudaMemsetAsync(dw32, 0, (size_t)CO * 32 * sizeof(float), st));
    if (int e = tss_pwconv_wgrad_tc(patches, dy, dw32, M, 32, CO, 32, CO, st)) return e;
    tss_launch(stem_wgrad_unpack_kernel, (CO * kTaps + 255) / 256, 256, 0, st, (const float*)dw32, dw);
    TSS_LAUNCH_CHECK("stem3x3s2_wgrad_from_patches(unpack)");
    return TSS_OK;
}

extern "C" int tss_stem3x3s2_wgrad_patches(const float* x, const void* dy, void* patches, float* dw32, float* dw, int N,
                                           int H, int W, int Cout, void* stream) {
    if (int e = tss_stem3x3s2_patches(x, patches, N, H, W, stream)) return e;
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    return tss_stem3x3s2_wgrad_from_patches(patches, dy, dw32, dw, (int64_t)N * Ho * Wo, Cout, stream);
}

extern "C" int tss_stem3x3s2_fwd_tc(const float* x, const float* w, void* y, int N, int H, int W, int Cout,
                                    const float* scale, const float* shift, int flags, double* stats, void* stream) {
    TSS_REQUIRE(N > 0 && H > 0 && W > 0, "stem3x3s2_fwd_tc: empty input");
    TSS_REQUIRE(Cout == CO, "stem3x3s2_fwd_tc: Cout=%d unsupported (only %d)", Cout, CO);
    TSS_REQUIRE(scale == nullptr || shift != nullptr, "stem3x3s2_fwd_tc: scale without shift");
    TSS_REQUIRE(((uintptr_t)y & 15) == 0, "stem3x3s2_fwd_tc: y must be 16-byte aligned");
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const int tiles_w = (Wo + kRows - 1) / kRows;
    const int64_t ntiles = (int64_t)N * Ho * tiles_w;
    const size_t smem = 1024 + kABytes + kBBytes + 8 + 8 + 2 * CO * sizeof(float);
    tss_launch(stem_tc_fwd_kernel, (unsigned)ceil_div64(ntiles, kTilesPerCta), kThreads, smem, (cudaStream_t)stream, x, w, (bf16*)y, N,
               H, W, Ho, Wo, tiles_w, ntiles, scale, shift, flags & TSS_EPI_RELU, stats);
    TSS_LAUNCH_CHECK("stem3x3s2_fwd_tc");
    return TSS_OK;
}

extern "C" int tss_stem3x3s2_wgrad_tc(const float* x, const void* dy, float* dw, int N, int H, int W, int Cout,
                                      void* stream) {
    TSS_REQUIRE(N > 0 && H > 0 && W > 0, "stem3x3s2_wgrad_tc: empty input");
    TSS_REQUIRE(Cout == CO, "stem3x3s2_wgrad_tc: Cout=%d unsupported (only %d)", Cout, CO);
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const int chunks_w = (Wo + kWgPix - 1) / kWgPix;
    const int64_t nblocks = (int64_t)N * Ho * chunks_w;
    int64_t grid = (int64_t)tss_num_sms() * 2;
    if (grid > nblocks) grid = nblocks;
    const size_t smem = 1024 + (size_t)kWgStages * (kABytes + kBBytes) + (2 * kWgStages + 1) * 8 + 8;
    tss_launch(stem_tc_wgrad_kernel, (unsigned)grid, kWgThreads, smem, (cudaStream_t)stream, x, (const bf16*)dy, dw, N, H, W, Ho, Wo,
               chunks_w, nblocks, (const bf16*)nullptr, (const float*)nullptr, (const float*)nullptr, (const float*)nullptr,
               (const float*)nullptr, (const float*)nullptr, 0, 0.f, (float*)nullptr, (float*)nullptr);
    TSS_LAUNCH_CHECK("stem3x3s2_wgrad_tc");
    return TSS_OK;
}

extern "C" int tss_stem3x3s2_wgrad_tc_bn(const float* x, const void* dz, const void* y, const float* mean, const float* rstd,
                                         const float* gamma, const float* beta, const float* sums, int flags, int64_t count,
                                         float* dw, float* dgamma, float* dbeta, int N, int H, int W, int Cout, void* stream) {
    TSS_REQUIRE(N > 0 && H > 0 && W > 0, "stem3x3s2_wgrad_tc_bn: empty input");
    TSS_REQUIRE(Cout == CO, "stem3x3s2_wgrad_tc_bn: Cout=%d unsupported (only %d)", Cout, CO);
    TSS_REQUIRE(dz != nullptr && y != nullptr && mean != nullptr && rstd != nullptr && sums != nullptr,
                "stem3x3s2_wgrad_tc_bn: missing BatchNorm operands");
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    if (count <= 0) count = (int64_t)N * Ho * Wo;
    const int chunks_w = (Wo + kWgPix - 1) / kWgPix;
    const int64_t nblocks = (int64_t)N * Ho * chunks_w;
    int64_t grid = (int64_t)tss_num_sms() * 2;
    if (grid > nblocks) grid = nblocks;
    const size_t smem = 1024 + (size_t)kWgStages * (kABytes + kBBytes) + (2 * kWgStages + 1) * 8 + 8;
    tss_launch(stem_tc_wgrad_kernel, (unsigned)grid, kWgThreads, smem, (cudaStream_t)stream, x, (const bf16*)dz, dw, N, H, W, Ho, Wo,
               chunks_w, nblocks, (const bf16*)y, mean, rstd, gamma, beta, sums, flags & TSS_EPI_RELU, (float)(1.0 / (double)count),
               dgamma, dbeta);
    TSS_LAUNCH_CHECK("stem3x3s2_wgrad_tc_bn");
    return TSS_OK;
}
