// Pointwise-conv dgrad on tcgen05 with the BatchNorm-backward REDUCTION of the producing layer fused
// into its epilogue (training).
//
//   dz[M][C] = dY[M][Nc] . W[Nc][C]          gradient w.r.t. the activated input z of this 1x1 conv
//   z = relu(BN(yp))  was produced by the previous layer from its raw conv output yp[M][C]; that
//   layer's BatchNorm backward starts with   g = dz * (z > 0),  sums[c] += sum_m g,
//   sums[C+c] += sum_m g * xhat,  xhat = (yp - mean) * rstd      (bn.cu: bn_bwd_reduce_kernel).
// The dgrad epilogue already holds the dz tile in registers: it reads the matching yp tile (the ReLU
// mask is recomputed from yp with the forward's arithmetic), stores g instead of dz and accumulates
// the two sums through a warp-private shared-memory transposition (colsum.cuh).
// The producer's BatchNorm backward then runs its apply pass only (no mask, no reduction): one full
// read of dz and yp and one launch less per layer.  Main loop identical to pwconv_tc.cu.
#include <stdlib.h>

#include "tc_ptx.cuh"
#include "colsum.cuh"

namespace {

constexpr int BM = 128;          // rows per tile = UMMA M
constexpr int BK = 64;           // bf16 elements per k-block = 128 bytes = one swizzle row
constexpr int kThreads = 192;
constexpr uint32_t kABytes = BM * BK * 2;

__global__ void __launch_bounds__(kThreads, 4)     // 4 x 192 threads x 85 registers: a whole 1/32-resolution layer (486 tiles) in one wave
pw_tc_bnred_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   bf16* __restrict__ G, int64_t M, int K, int64_t ldg, int block_n, int stages, uint32_t tmem_cols,
                   const bf16* __restrict__ yp, int64_t ldyp, const float* __restrict__ mean,
                   const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                   int relu, float* __restrict__ sums, int sums_stride) {
    TSS_DYN_SMEM(uint8_t, smem_raw);
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t b_bytes = (uint32_t)block_n * BK * 2;
    uint8_t* sA = smem;
    uint8_t* sB = smem + (size_t)stages * kABytes;
    uint64_t* bars = (uint64_t*)(sB + (size_t)stages * b_bytes);     // full[stages], empty[stages], tmem_full
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * stages + 1);
    float* s_stat = (float*)(((uintptr_t)(tmem_slot + 2) + 15) & ~(uintptr_t)15);   // [4 epilogue warps][2][block_n]
    float* s_const = s_stat + 8 * block_n;                            // [4][block_n]: mean, rstd, scale, shift

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * block_n;
    const int num_kb = (K + BK - 1) / BK;
    TSS_MARK(0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(smem_u32(bars + s), 1);
            mbar_init(smem_u32(bars + stages + s), 1);
        }
        mbar_init(smem_u32(bars + 2 * stages), 1);
        mbar_init_fence();
    }
    if (warp == 1) {
        tc_alloc(smem_u32(tmem_slot), tmem_cols);
    }
    for (int i = threadIdx.x; i < 8 * block_n; i += kThreads) s_stat[i] = 0.f;
    TSS_MARK(1);
    if (threadIdx.x == 0) {tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }      // descriptor fetch (~0.5 us) under the predecessor's tail
    pdl_wait();
    TSS_MARK(2);
    // the first pass over the ring needs no empty-slot wait: thread 0 (which initialised the barriers) issues those loads
    // NOW, so that they fly while the CTA fetches the BatchNorm constants and meets at the barrier below (~1.5 us)
    const int pre_kb = num_kb < stages ? num_kb : stages;
    if (threadIdx.x == 0) {
        for (int kb = 0; kb < pre_kb; ++kb) {
            const uint32_t full = smem_u32(bars + kb);
            mbar_expect_tx(full, kABytes + b_bytes);
            tma_load_2d(smem_u32(sA + (size_t)kb * kABytes), &tmA, full, kb * BK, (int)m0);
            tma_load_2d(smem_u32(sB + (size_t)kb * b_bytes), &tmB, full, kb * BK, n0);
        }
    }
    for (int i = threadIdx.x; i < block_n; i += kThreads) {           // per-column constants of the producer's BatchNorm
        const float mu = __ldg(mean + n0 + i), rs = __ldg(rstd + n0 + i);
        const float sc = (gamma != nullptr ? __ldg(gamma + n0 + i) : 1.f) * rs;
        s_const[i] = mu;
        s_const[block_n + i] = rs;
        s_const[2 * block_n + i] = sc;
        s_const[3 * block_n + i] = (beta != nullptr ? __ldg(beta + n0 + i) : 0.f) - mu * sc;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {                                   // ---------------- TMA producer
            for (int kb = pre_kb; kb < num_kb; ++kb) {
                const int s = kb % stages;
                const uint32_t phase = (kb / stages) & 1;
                mbar_wait(smem_u32(bars + stages + s), phase ^ 1);
                const uint32_t full = smem_u32(bars + s);
                mbar_expect_tx(full, kABytes + b_bytes);
                tma_load_2d(smem_u32(sA + (size_t)s * kABytes), &tmA, full, kb * BK, (int)m0);
                tma_load_2d(smem_u32(sB + (size_t)s * b_bytes), &tmB, full, kb * BK, n0);
            }
            TSS_MARK_IF(true, 3);
        }
    } else if (warp == 1) {
        if (lane == 0) {                                   // ---------------- MMA issuer
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(block_n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % stages;
                const uint32_t phase = (kb / stages) & 1;
                mbar_wait(smem_u32(bars + s), phase);
                TSS_MARK_IF(kb == 0, 4);
                tc_fence_after();
                const uint64_t adesc = make_desc_k_sw128(smem_u32(sA + (size_t)s * kABytes));
                const uint64_t bdesc = make_desc_k_sw128(smem_u32(sB + (size_t)s * b_bytes));
                int rem = K - kb * BK;
                const int k16 = rem >= BK ? BK / 16 : (rem + 15) / 16;
                for (int k = 0; k < k16; ++k)
                    umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (uint32_t)(kb > 0 || k > 0));
                umma_commit(smem_u32(bars + stages + s));
            }
            umma_commit(smem_u32(bars + 2 * stages));
            TSS_MARK_IF(true, 5);
        }
    } else {                                               // ---------------- epilogue warps 2..5
        const int q = warp & 3;
        const int row_in_tile = q * 32 + lane;
        const int64_t row = m0 + row_in_tile;
        const bool row_ok = row < M;
        // 256-bit accesses when every 16-column chunk of a row starts on a 32-byte boundary (block_n is a multiple of 16)
        const bool wide = (((uintptr_t)yp | (uintptr_t)G) & 31) == 0 && ((ldyp | ldg) & 15) == 0;
        // this thread's row of the producer's raw output (block_n <= 64 columns = 8 x 16 bytes) is requested NOW, while the
        // TMA / MMA main loop is still running, instead of chunk by chunk after the accumulator is complete
        uint4 ypre[8];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            ypre[j] = ypre[j + 1] = make_uint4(0, 0, 0, 0);     // rows >= M: dz is an exact zero (TMA zero fill)
            if (row_ok && j * 8 < block_n) {
                if (wide) {
                    ldg256(yp + row * ldyp + n0 + j * 8, ypre[j], ypre[j + 1]);
                } else {
                    ypre[j] = __ldg(reinterpret_cast<const uint4*>(yp + row * ldyp + n0 + j * 8));
                    ypre[j + 1] = __ldg(reinterpret_cast<const uint4*>(yp + row * ldyp + n0 + j * 8 + 8));
                }
            }
        }
        TSS_MARK_IF(threadIdx.x == 64, 10);
        mbar_wait(smem_u32(bars + 2 * stages), 0);
        TSS_MARK_IF(threadIdx.x == 64, 6);
        tc_fence_after();
        // column sums through this warp's scratch: the pipeline stages are dead once the accumulator is complete (every
        // TMA write has landed and every MMA has read its operands), so the scratch aliases them (the launcher keeps
        // the ring at >= 4 x kCsPair floats)
        float* scratch = reinterpret_cast<float*>(smem) + q * kCsPair;
        float* mine = s_stat + q * 2 * block_n;            // this warp's private slice: plain read-modify-write
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const int c = cc * 16;
            if (c >= block_n) break;
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
            TSS_MARK_IF(threadIdx.x == 64 && cc == 0, 11);
            const int col = n0 + c;
            float y0[8], y1[8];
            unpack8(ypre[2 * cc], y0);
            unpack8(ypre[2 * cc + 1], y1);
            float gx[16];                                  // g * (yp - mean); rstd comes in once per column at the end
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
                const float4 mu4 = *reinterpret_cast<const float4*>(s_const + c + 4 * i4);
                const float4 sc4 = *reinterpret_cast<const float4*>(s_const + 2 * block_n + c + 4 * i4);
                const float4 sh4 = *reinterpret_cast<const float4*>(s_const + 3 * block_n + c + 4 * i4);
                const float mu[4] = {mu4.x, mu4.y, mu4.z, mu4.w}, sc[4] = {sc4.x, sc4.y, sc4.z, sc4.w}, sh[4] = {sh4.x, sh4.y, sh4.z, sh4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int i = 4 * i4 + e;
                    const float yy = i < 8 ? y0[i] : y1[i - 8];
                    if (relu && !(fmaf(yy, sc[e], sh[e]) > 0.f)) v[i] = 0.f;
                    gx[i] = v[i] * (yy - mu[e]);
                }
            }
            cs_store16(scratch, lane, v);
            cs_store16(scratch + kCsArray + 16, lane, gx);
            __syncwarp();
            mine[(lane >> 4) * block_n + c + (lane & 15)] += cs_sum_pair(scratch, lane);   // lanes 0..15: sum g, 16..31: sum g (yp - mean)
            __syncwarp();                                  // the scratch may be rewritten
            TSS_MARK_IF(threadIdx.x == 64 && cc == 0, 12);
            if (row_ok) {
                uint32_t o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
                bf16* dst = G + row * ldg + col;
                if (wide) {
                    stg256(dst, make_uint4(o[0], o[1], o[2], o[3]), make_uint4(o[4], o[5], o[6], o[7]));
                } else {
                    *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
                    *reinterpret_cast<uint4*>(dst + 8) = make_uint4(o[4], o[5], o[6], o[7]);
                }
            }
            TSS_MARK_IF(threadIdx.x == 64 && cc == 0, 13);
        }
        TSS_MARK_IF(threadIdx.x == 64, 7);
        tc_fence_before();
    }
    __syncthreads();
    TSS_MARK(8);
    for (int i = threadIdx.x; i < block_n; i += kThreads) {
        atomicAdd(sums + n0 + i, (s_stat[i] + s_stat[2 * block_n + i]) + (s_stat[4 * block_n + i] + s_stat[6 * block_n + i]));
        atomicAdd(sums + sums_stride + n0 + i, s_const[block_n + i] *
                  ((s_stat[block_n + i] + s_stat[3 * block_n + i]) + (s_stat[5 * block_n + i] + s_stat[7 * block_n + i])));
    }
    TSS_MARK(9);
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        tc_dealloc(tmem_base, tmem_cols);
    }
}

// ------------------------------------------------------------------ the persistent kernel ------------
// The same tile and epilogue on the schedule of pw_tc_persistent_kernel (pwconv_tc.cu): a CTA walks row tiles
// blockIdx.x, blockIdx.x + gridDim.x, ... of its column tile, the accumulator is double-buffered in TMEM (the MMA warp fills
// buffer (j+1) & 1 while the epilogue warps drain buffer j & 1), the TMA ring runs ahead across tiles, and an epilogue
// thread requests its row of the producer's output for tile j+1 before it starts on tile j.  The one-tile kernel pays
// barrier init + TMEM allocation + constants + TMA + MMA latency (~5 us, tools/trace_kernels.py) in front of every 4.5 us
// epilogue: three waves of such CTAs made the fused dgrad into 128 channels at 1/8 resolution slower (41 us) than the plain
// persistent dgrad plus a stand-alone reduction.  The two sums stay in shared memory over all the CTA's tiles.
__global__ void __launch_bounds__(kThreads, 3)
pw_tc_bnred_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                              bf16* __restrict__ G, int64_t M, int K, int64_t ldg, int block_n, int stages, uint32_t tmem_cols,
                              const bf16* __restrict__ yp, int64_t ldyp, const float* __restrict__ mean,
                              const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                              int relu, float* __restrict__ sums, int sums_stride, int m_tiles) {
    TSS_DYN_SMEM(uint8_t, smem_raw);
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t b_bytes = (uint32_t)block_n * BK * 2;
    uint8_t* sA = smem;
    uint8_t* sB = smem + (size_t)stages * kABytes;
    uint64_t* bars = (uint64_t*)(sB + (size_t)stages * b_bytes);     // full[stages], empty[stages], tmem_full[2], tmem_empty[2]
    uint64_t* tmem_full = bars + 2 * stages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);
    float* s_stat = (float*)(((uintptr_t)(tmem_slot + 2) + 15) & ~(uintptr_t)15);   // [4 epilogue warps][2][block_n]
    float* s_const = s_stat + 8 * block_n;                            // [4][block_n]: mean, rstd, scale, shift
    float* s_scratch = s_const + 4 * block_n;                         // [4 epilogue warps][kCsPair] (colsum.cuh)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.y * block_n;
    const int num_kb = (K + BK - 1) / BK;
    const int my_tiles = ((int)blockIdx.x < m_tiles) ? (m_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(smem_u32(bars + s), 1);
            mbar_init(smem_u32(bars + stages + s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(smem_u32(tmem_full + a), 1);
            mbar_init(smem_u32(tmem_empty + a), 4);                  // one arrival per epilogue warp
        }
        mbar_init_fence();
    }
    if (warp == 1) {
        tc_alloc(smem_u32(tmem_slot), tmem_cols);
    }
    for (int i = threadIdx.x; i < 8 * block_n; i += kThreads) s_stat[i] = 0.f;
    if (threadIdx.x == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
    pdl_wait();
    for (int i = threadIdx.x; i < block_n; i += kThreads) {           // per-column constants of the producer's BatchNorm
        const float mu = __ldg(mean + n0 + i), rs = __ldg(rstd + n0 + i);
        const float sc = (gamma != nullptr ? __ldg(gamma + n0 + i) : 1.f) * rs;
        s_const[i] = mu;
        s_const[block_n + i] = rs;
        s_const[2 * block_n + i] = sc;
        s_const[3 * block_n + i] = (beta != nullptr ? __ldg(beta + n0 + i) : 0.f) - mu * sc;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {                                   // ---------------- TMA producer
            int it = 0;
            for (int j = 0; j < my_tiles; ++j) {
                const int m0 = ((int)blockIdx.x + j * (int)gridDim.x) * BM;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % stages;
                    const uint32_t phase = (it / stages) & 1;
                    mbar_wait(smem_u32(bars + stages + s), phase ^ 1);
                    const uint32_t full = smem_u32(bars + s);
                    mbar_expect_tx(full, kABytes + b_bytes);
                    tma_load_2d(smem_u32(sA + (size_t)s * kABytes), &tmA, full, kb * BK, m0);
                    tma_load_2d(smem_u32(sB + (size_t)s * b_bytes), &tmB, full, kb * BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                   // ---------------- MMA issuer
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(block_n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            int it = 0;
            for (int j = 0; j < my_tiles; ++j) {
                const int a = j & 1;
                mbar_wait(smem_u32(tmem_empty + a), ((uint32_t)(j >> 1) & 1) ^ 1);   // the epilogue has drained this buffer
                tc_fence_after();
                const uint32_t acc = tmem_base + (uint32_t)(a * block_n);
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % stages;
                    const uint32_t phase = (it / stages) & 1;
                    mbar_wait(smem_u32(bars + s), phase);
                    tc_fence_after();
                    const uint64_t adesc = make_desc_k_sw128(smem_u32(sA + (size_t)s * kABytes));
                    const uint64_t bdesc = make_desc_k_sw128(smem_u32(sB + (size_t)s * b_bytes));
                    int rem = K - kb * BK;
                    const int k16 = rem >= BK ? BK / 16 : (rem + 15) / 16;
                    for (int k = 0; k < k16; ++k)
                        umma_bf16(acc, adesc + 2 * k, bdesc + 2 * k, idesc, (uint32_t)(kb > 0 || k > 0));
                    umma_commit(smem_u32(bars + stages + s));
                }
                umma_commit(smem_u32(tmem_full + a));                                 // accumulator of tile j complete
            }
        }
    } else {                                               // ---------------- epilogue warps 2..5
        const int q = warp & 3;
        const int row_in_tile = q * 32 + lane;
        float* scratch = s_scratch + q * kCsPair;
        float* mine = s_stat + q * 2 * block_n;            // this warp's private slice: plain read-modify-write
        const bool wide = (((uintptr_t)yp | (uintptr_t)G) & 31) == 0 && ((ldyp | ldg) & 15) == 0;
        auto request = [&](int j, uint4 (&dst)[8]) {       // this thread's row of the producer's raw output for tile j
            const int64_t row = (int64_t)((int)blockIdx.x + j * (int)gridDim.x) * BM + row_in_tile;
#pragma unroll
            for (int u = 0; u < 8; u += 2) {
                dst[u] = dst[u + 1] = make_uint4(0, 0, 0, 0);   // rows >= M: dz is an exact zero (TMA zero fill)
                if (j < my_tiles && row < M && u * 8 < block_n) {
                    if (wide) {
                        ldg256(yp + row * ldyp + n0 + u * 8, dst[u], dst[u + 1]);
                    } else {
                        dst[u] = __ldg(reinterpret_cast<const uint4*>(yp + row * ldyp + n0 + u * 8));
                        dst[u + 1] = __ldg(reinterpret_cast<const uint4*>(yp + row * ldyp + n0 + u * 8 + 8));
                    }
                }
            }
        };
        uint4 ycur[8], ynext[8];
        request(0, ynext);
        for (int j = 0; j < my_tiles; ++j) {
            const int a = j & 1;
            const int64_t row = (int64_t)((int)blockIdx.x + j * (int)gridDim.x) * BM + row_in_tile;
            const bool row_ok = row < M;
#pragma unroll
            for (int u = 0; u < 8; ++u) ycur[u] = ynext[u];
            request(j + 1, ynext);
            mbar_wait(smem_u32(tmem_full + a), (uint32_t)(j >> 1) & 1);
            tc_fence_after();
            const uint32_t acc = tmem_base + (uint32_t)(a * block_n) + ((uint32_t)(q * 32) << 16);
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const int c = cc * 16;
                if (c >= block_n) break;
                float v[16];
                tmem_ld16(acc + (uint32_t)c, v);
                float y0[8], y1[8];
                unpack8(ycur[2 * cc], y0);
                unpack8(ycur[2 * cc + 1], y1);
                float gx[16];                              // g * (yp - mean); rstd comes in once per column at the end
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                    const float4 mu4 = *reinterpret_cast<const float4*>(s_const + c + 4 * i4);
                    const float4 sc4 = *reinterpret_cast<const float4*>(s_const + 2 * block_n + c + 4 * i4);
                    const float4 sh4 = *reinterpret_cast<const float4*>(s_const + 3 * block_n + c + 4 * i4);
                    const float mu[4] = {mu4.x, mu4.y, mu4.z, mu4.w}, sc[4] = {sc4.x, sc4.y, sc4.z, sc4.w}, sh[4] = {sh4.x, sh4.y, sh4.z, sh4.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int i = 4 * i4 + e;
                        const float yy = i < 8 ? y0[i] : y1[i - 8];
                        if (relu && !(fmaf(yy, sc[e], sh[e]) > 0.f)) v[i] = 0.f;
                        gx[i] = v[i] * (yy - mu[e]);
                    }
                }
                cs_store16(scratch, lane, v);
                cs_store16(scratch + kCsArray + 16, lane, gx);
                __syncwarp();
                mine[(lane >> 4) * block_n + c + (lane & 15)] += cs_sum_pair(scratch, lane);
                __syncwarp();                              // the scratch may be rewritten
                if (row_ok) {
                    uint32_t o[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) o[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
                    bf16* dst = G + row * ldg + n0 + c;
                    if (wide) {
                        stg256(dst, make_uint4(o[0], o[1], o[2], o[3]), make_uint4(o[4], o[5], o[6], o[7]));
                    } else {
                        *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
                        *reinterpret_cast<uint4*>(dst + 8) = make_uint4(o[4], o[5], o[6], o[7]);
                    }
                }
            }
            // this warp has read its quarter of the buffer (tcgen05.wait::ld inside tmem_ld16): hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(tmem_empty + a));
        }
        tc_fence_before();
    }
    __syncthreads();
    if (my_tiles > 0) {
        for (int i = threadIdx.x; i < block_n; i += kThreads) {
            atomicAdd(sums + n0 + i, (s_stat[i] + s_stat[2 * block_n + i]) + (s_stat[4 * block_n + i] + s_stat[6 * block_n + i]));
            atomicAdd(sums + sums_stride + n0 + i, s_const[block_n + i] *
                      ((s_stat[block_n + i] + s_stat[3 * block_n + i]) + (s_stat[5 * block_n + i] + s_stat[7 * block_n + i])));
        }
    }
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        tc_dealloc(tmem_base, tmem_cols);
    }
}

int make_map_bnred(CUtensorMap* map, const void* base, int64_t rows, int cols, int64_t ld, int box_rows) {
    typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    tss_bind_context();
    static EncodeTiledFn enc = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    TSS_REQUIRE(enc != nullptr, "pwconv_dgrad_bnred: cuTensorMapEncodeTiled is not available from the driver");
    TSS_REQUIRE(((uintptr_t)base & 15) == 0 && (ld * 2) % 16 == 0, "pwconv_dgrad_bnred: TMA needs 16-byte aligned base and pitch");
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TSS_REQUIRE(r == CUDA_SUCCESS, "pwconv_dgrad_bnred: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return TSS_OK;
}

}  // namespace

extern "C" int tss_pwconv_dgrad_bnred(const void* dy, const void* wpT, void* g, int64_t M, int K, int Nc, int64_t lddy,
                                      int64_t ldg, const void* yp, int64_t ldyp, const float* mean, const float* rstd,
                                      const float* gamma, const float* beta, int flags, float* sums, void* stream) {
    // GEMM: G[M][K] = dY[M][Nc] . (W^T)[K][Nc]^T  -- reduction over Nc, K output columns (= the producer's channels)
    TSS_REQUIRE(M > 0 && K > 0 && Nc > 0 && K % 16 == 0 && Nc % 8 == 0, "pwconv_dgrad_bnred: M=%lld K=%d Nc=%d", (long long)M, K, Nc);
    TSS_REQUIRE(ldg % 8 == 0 && ldyp % 8 == 0 && ((uintptr_t)g & 15) == 0 && ((uintptr_t)yp & 15) == 0,
                "pwconv_dgrad_bnred: output / yp must be 16-byte aligned with pitches that are multiples of 8");
    TSS_REQUIRE(yp != nullptr && mean != nullptr && rstd != nullptr && sums != nullptr, "pwconv_dgrad_bnred: missing BatchNorm operands");
    int bn = 0;
    for (int b = 64; b >= 16; b -= 16)
        if (K % b == 0) { bn = b; break; }
    TSS_REQUIRE(bn >= 16, "pwconv_dgrad_bnred: no tile width for K=%d", K);
    CUtensorMap tmA, tmB;
    if (int e = make_map_bnred(&tmA, dy, M, Nc, lddy, BM)) return e;
    if (int e = make_map_bnred(&tmB, wpT, K, Nc, Nc, bn)) return e;
    const int num_kb = (Nc + BK - 1) / BK;
    {
        // persistent CTAs (double-buffered TMEM accumulator, sums flushed once per CTA): 0 = never, 1 = at most two column
        // tiles and at least four waves of row tiles (the rule of the forward kernel, pwconv_tc.cu), 2 = always
        static const int persist = [] { const char* e = getenv("TSS_PW_BNRED_PERSIST"); return e != nullptr ? atoi(e) : 0; }();
        const int64_t m_tiles = ceil_div64(M, BM);
        const int n_tiles = K / bn;
        if ((persist == 2 || (persist == 1 && n_tiles <= 2 && m_tiles >= 4 * (int64_t)tss_num_sms())) && m_tiles < (1ll << 30)) {
            const int stages = 2;
            uint32_t tmem_cols = 32;
            while ((int)tmem_cols < 2 * bn) tmem_cols <<= 1;
            const size_t smem = 1024 + (size_t)stages * (kABytes + (size_t)bn * BK * 2) + (2 * stages + 4) * 8 + 8 + 16 + 12 * bn * sizeof(float) +
                                (size_t)4 * kCsPair * sizeof(float);
            static bool attr_set_p = false;
            if (!attr_set_p) {
                TSS_CUDA(cudaFuncSetAttribute(pw_tc_bnred_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
                attr_set_p = true;
            }
            int64_t gx = ((int64_t)tss_num_sms() * 3 + n_tiles - 1) / n_tiles;        // three resident CTAs per SM (74 KB, 85 registers)
            if (gx > m_tiles) gx = m_tiles;
            if (gx < 1) gx = 1;
            tss_launch(pw_tc_bnred_persistent_kernel, dim3((unsigned)gx, (unsigned)n_tiles), kThreads, smem, (cudaStream_t)stream, tmA, tmB,
                       (bf16*)g, M, Nc, ldg, bn, stages, tmem_cols, (const bf16*)yp, ldyp, mean, rstd, gamma, beta, flags & TSS_EPI_RELU,
                       sums, K, (int)m_tiles);
            TSS_LAUNCH_CHECK("pwconv_dgrad_bnred(persistent)");
            return TSS_OK;
        }
    }
    int stages = num_kb < 4 ? num_kb : 4;
    if (stages < 2) stages = 2;        // the epilogue's column-sum scratch (4 x kCsPair floats = 20.3 KB) aliases the ring
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < bn) tmem_cols <<= 1;
    const size_t smem = 1024 + (size_t)stages * (kABytes + (size_t)bn * BK * 2) + (2 * stages + 1) * 8 + 8 + 16 + 12 * bn * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        TSS_CUDA(cudaFuncSetAttribute(pw_tc_bnred_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    dim3 grid((unsigned)ceil_div64(M, BM), (unsigned)(K / bn));
    tss_launch(pw_tc_bnred_kernel, grid, kThreads, smem, (cudaStream_t)stream, tmA, tmB, (bf16*)g, M, Nc, ldg, bn, stages, tmem_cols,
               (const bf16*)yp, ldyp, mean, rstd, gamma, beta, flags & TSS_EPI_RELU, sums, K);
    TSS_LAUNCH_CHECK("pwconv_dgrad_bnred");
    return TSS_OK;
}
