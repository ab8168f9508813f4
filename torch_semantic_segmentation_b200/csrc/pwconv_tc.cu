// Pointwise (1x1) convolution on the 5th-generation tensor cores (sm_100a only).
//
//   Y[M][Nc] = X[M][K] . W[Nc][K]^T  (+ BatchNorm statistics | affine + residual + ReLU), bf16 in,
//   fp32 accumulate in TMEM, bf16 out.   dgrad is the same kernel with the transposed weight pack.
//
// One CTA per 128-row x BLOCK_N tile (BLOCK_N = largest multiple-of-16 divisor of Nc <= 256):
//   warp 0 (one lane): TMA producer  -- cp.async.bulk.tensor.2d boxes of 128 x 64 (A) and
//                      BLOCK_N x 64 (B) bf16, 128-byte swizzle, mbarrier complete_tx;
//                      K tails and the M tail are zero-filled by TMA.
//   warp 1 (one lane): MMA issuer    -- tcgen05.mma.cta_group::1.kind::f16, M=128, N=BLOCK_N,
//                      K=16 per instruction, SS operands through shared-memory descriptors;
//                      tcgen05.commit releases smem stages / publishes the accumulator.
//                      (the whole warp allocates / frees the TMEM columns)
//   warps 2-5:         epilogue      -- tcgen05.ld 32x32b (one output row per thread), fused
//                      BatchNorm statistics (warp-transposed shuffle sums -> smem -> one global
//                      atomic per column per CTA) or scale/shift/residual/ReLU, bf16 stores.
// These GEMMs are HBM-bound (arithmetic intensity 19..110 flop/B vs a ridge of ~246): several
// CTAs stay resident per SM (smem = stages x (16 KB + BLOCK_N x 128 B), TMEM = BLOCK_N columns)
// so that loads, MMAs and stores of different tiles overlap.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "colsum.cuh"

namespace {

constexpr int BM = 128;          // rows per tile = UMMA M
constexpr int BK = 64;           // bf16 elements per k-block = 128 bytes = one swizzle row
constexpr int kThreads = 192;
constexpr uint32_t kABytes = BM * BK * 2;

// ------------------------------------------------------------------ PTX wrappers ------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);        // start address
    d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset
    d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
    return d;
}

// lanes 2j / 2j+1 end with the sum over the 32 lanes of v[j], j = lane >> 1
__device__ __forceinline__ float warp_transpose_sum16(float (&v)[16], int lane) {
#pragma unroll
    for (int step = 16, n = 16; step >= 2; step >>= 1, n >>= 1) {
        const bool upper = (lane & step) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float send = upper ? v[i] : v[i + n / 2];
            const float keep = upper ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// ------------------------------------------------------------------ the kernel ---------
__global__ void __launch_bounds__(kThreads)
pw_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             bf16* __restrict__ Y, int64_t M, int K, int64_t ldy, int block_n, int stages, uint32_t tmem_cols,
             const float* __restrict__ scale, const float* __restrict__ shift, const bf16* __restrict__ res,
             int64_t ldr, int relu, double* __restrict__ stats, int stats_stride) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SW128 needs 1024-B alignment
    const uint32_t b_bytes = (uint32_t)block_n * BK * 2;
    uint8_t* sA = smem;
    uint8_t* sB = smem + (size_t)stages * kABytes;
    uint64_t* bars = (uint64_t*)(sB + (size_t)stages * b_bytes);     // full[stages], empty[stages], tmem_full
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * stages + 1);
    float* s_stat = (float*)(tmem_slot + 2);                          // [4 epilogue warps][2][block_n]

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
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (stats != nullptr)
        for (int i = threadIdx.x; i < 8 * block_n; i += kThreads) s_stat[i] = 0.f;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    TSS_MARK(1);
    if (threadIdx.x == 0) {tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }      // descriptor fetch (~0.5 us) under the predecessor's tail
    pdl_wait();          // everything above is on-chip setup; global memory is touched from here on
    TSS_MARK(2);

    if (warp == 0) {
        if (lane == 0) {                                   // ---------------- TMA producer
            for (int kb = 0; kb < num_kb; ++kb) {
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
            // instruction descriptor: D=f32, A=B=bf16, both K-major, N=block_n, M=128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(block_n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % stages;
                const uint32_t phase = (kb / stages) & 1;
                mbar_wait(smem_u32(bars + s), phase);
                TSS_MARK_IF(kb == 0, 4);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t adesc = make_desc_k_sw128(smem_u32(sA + (size_t)s * kABytes));
                const uint64_t bdesc = make_desc_k_sw128(smem_u32(sB + (size_t)s * b_bytes));
                int rem = K - kb * BK;
                const int k16 = rem >= BK ? BK / 16 : (rem + 15) / 16;     // skip all-zero K slices of the tail
                for (int k = 0; k < k16; ++k)                                // +32 bytes (= 2 x 16 B) per K=16 slice
                    umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (uint32_t)(kb > 0 || k > 0));
                umma_commit(smem_u32(bars + stages + s));                    // smem stage free once these MMAs retire
            }
            umma_commit(smem_u32(bars + 2 * stages));                        // accumulator complete
            TSS_MARK_IF(true, 5);
        }
    } else {                                               // ---------------- epilogue warps 2..5
        const int q = warp & 3;                            // TMEM lane quarter this warp may access
        const int row_in_tile = q * 32 + lane;
        const int64_t row = m0 + row_in_tile;
        const bool row_ok = row < M;
        mbar_wait(smem_u32(bars + 2 * stages), 0);
        TSS_MARK_IF(threadIdx.x == 64, 6);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int c = 0; c < block_n; c += 16) {
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
            if (stats != nullptr) {                        // rows >= M are exact zeros (TMA zero fill)
                // column sums through this warp's scratch (colsum.cuh), double-buffered: one __syncwarp per block.  The
                // scratch aliases the pipeline stages, which are dead once the accumulator is complete (the launcher
                // keeps the ring at >= 4 x 2 x kCsArray floats).
                float* buf = reinterpret_cast<float*>(smem) + (q * 2 + ((c >> 4) & 1)) * kCsArray;
                cs_store16(buf, lane, v);
                __syncwarp();
                float s1, s2;
                cs_sum_sq(buf, lane, s1, s2);
                if (lane < 16) {                           // this warp's private slice: plain read-modify-write, one lane per column
                    float* mine = s_stat + q * 2 * block_n;   // (a shared-memory float atomicAdd is a CAS spin loop in SASS)
                    mine[c + lane] += s1;
                    mine[block_n + c + lane] += s2;
                }
            }
            if (row_ok) {
                const int col = n0 + c;
                if (shift != nullptr) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        v[i] = fmaf(v[i], scale != nullptr ? __ldg(scale + col + i) : 1.f, __ldg(shift + col + i));
                }
                if (res != nullptr) {
                    float r0[8], r1[8];
                    load8(res + row * ldr + col, r0);
                    load8(res + row * ldr + col + 8, r1);
#pragma unroll
                    for (int i = 0; i < 8; ++i) { v[i] += r0[i]; v[8 + i] += r1[i]; }
                }
                if (relu) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
                }
                // 16 bf16 = one full 32-byte sector per thread and store instruction (st.global.v8.b32)
                uint32_t o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
                bf16* dst = Y + row * ldy + col;
                if ((((uintptr_t)dst) & 31) == 0) {
                    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                                 ::"l"(dst), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7])
                                 : "memory");
                } else {
                    *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
                    *reinterpret_cast<uint4*>(dst + 8) = make_uint4(o[4], o[5], o[6], o[7]);
                }
            }
        }
        TSS_MARK_IF(threadIdx.x == 64, 7);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    TSS_MARK(8);
    if (stats != nullptr) {
        for (int i = threadIdx.x; i < block_n; i += kThreads) {
            const float a = (s_stat[i] + s_stat[2 * block_n + i]) + (s_stat[4 * block_n + i] + s_stat[6 * block_n + i]);
            const float b = (s_stat[block_n + i] + s_stat[3 * block_n + i]) + (s_stat[5 * block_n + i] + s_stat[7 * block_n + i]);
            atomicAdd(stats + n0 + i, (double)a);
            atomicAdd(stats + stats_stride + n0 + i, (double)b);
        }
    }
    TSS_MARK(9);
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------ the persistent kernel ------------
// Same tile, same roles, but a CTA walks row tiles m = blockIdx.x, blockIdx.x + gridDim.x, ... of its column tile:
//   * the accumulator is DOUBLE-BUFFERED in TMEM (2 x BLOCK_N columns): the MMA warp fills buffer (j+1) & 1 while the
//     epilogue warps drain buffer j & 1 (tmem_full[2] / tmem_empty[2] mbarriers); the TMA ring runs ahead across tiles;
//   * the BatchNorm statistics are accumulated in shared memory over ALL the CTA's tiles and leave it as ONE fp64 atomic
//     per column per CTA.  The one-tile-per-CTA kernel above issues m_tiles atomics on each of the 2 x Nc addresses --
//     3456 of them per address at 1/4 resolution -- and the serialisation of same-address atomics in L2 made the forward
//     pass 3 x slower than the dgrad of the same shape (ncu r1z: 67 us against 23 us for 32->48 @ 1/4).
__global__ void __launch_bounds__(kThreads)
pw_tc_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        bf16* __restrict__ Y, int64_t M, int K, int64_t ldy, int block_n, int stages, uint32_t tmem_cols,
                        const float* __restrict__ scale, const float* __restrict__ shift, const bf16* __restrict__ res,
                        int64_t ldr, int relu, double* __restrict__ stats, int stats_stride, int m_tiles, int cs_smem) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t b_bytes = (uint32_t)block_n * BK * 2;
    uint8_t* sA = smem;
    uint8_t* sB = smem + (size_t)stages * kABytes;
    uint64_t* bars = (uint64_t*)(sB + (size_t)stages * b_bytes);     // full[stages], empty[stages], tmem_full[2], tmem_empty[2]
    uint64_t* tmem_full = bars + 2 * stages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);
    float* s_stat = (float*)(((uintptr_t)(tmem_slot + 2) + 15) & ~(uintptr_t)15);   // [4 epilogue warps][2][block_n]
    float* cs_scratch = cs_smem ? s_stat + 8 * block_n : nullptr;     // [4 epilogue warps][kCsArray] (colsum.cuh)

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
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (stats != nullptr)
        for (int i = threadIdx.x; i < 8 * block_n; i += kThreads) s_stat[i] = 0.f;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) {tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }      // descriptor fetch (~0.5 us) under the predecessor's tail
    pdl_wait();

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
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc = tmem_base + (uint32_t)(a * block_n);
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % stages;
                    const uint32_t phase = (it / stages) & 1;
                    mbar_wait(smem_u32(bars + s), phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
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
        for (int j = 0; j < my_tiles; ++j) {
            const int a = j & 1;
            const int64_t m0 = (int64_t)((int)blockIdx.x + j * (int)gridDim.x) * BM;
            const int64_t row = m0 + row_in_tile;
            const bool row_ok = row < M;
            mbar_wait(smem_u32(tmem_full + a), (uint32_t)(j >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t acc = tmem_base + (uint32_t)(a * block_n) + ((uint32_t)(q * 32) << 16);
            for (int c = 0; c < block_n; c += 16) {
                float v[16];
                tmem_ld16(acc + (uint32_t)c, v);
                if (stats != nullptr && cs_scratch != nullptr) {     // rows >= M are exact zeros (TMA zero fill)
                    // column sums through this warp's dedicated scratch (colsum.cuh; the ring is live here), single-buffered
                    float* buf = cs_scratch + q * kCsArray;
                    cs_store16(buf, lane, v);
                    __syncwarp();
                    float s1, s2;
                    cs_sum_sq(buf, lane, s1, s2);
                    __syncwarp();                          // the scratch may be rewritten
                    if (lane < 16) {                       // this warp's private slice: plain read-modify-write
                        float* mine = s_stat + q * 2 * block_n;
                        mine[c + lane] += s1;
                        mine[block_n + c + lane] += s2;
                    }
                } else if (stats != nullptr) {
                    float sq[16], sm[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) { sm[i] = v[i]; sq[i] = v[i] * v[i]; }
                    const float s1 = warp_transpose_sum16(sm, lane);
                    const float s2 = warp_transpose_sum16(sq, lane);
                    if ((lane & 1) == 0) {                 // this warp's private slice: plain read-modify-write
                        float* mine = s_stat + q * 2 * block_n;
                        mine[c + (lane >> 1)] += s1;
                        mine[block_n + c + (lane >> 1)] += s2;
                    }
                }
                if (row_ok) {
                    const int col = n0 + c;
                    if (shift != nullptr) {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            v[i] = fmaf(v[i], scale != nullptr ? __ldg(scale + col + i) : 1.f, __ldg(shift + col + i));
                    }
                    if (res != nullptr) {
                        float r0[8], r1[8];
                        load8(res + row * ldr + col, r0);
                        load8(res + row * ldr + col + 8, r1);
#pragma unroll
                        for (int i = 0; i < 8; ++i) { v[i] += r0[i]; v[8 + i] += r1[i]; }
                    }
                    if (relu) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
                    }
                    uint32_t o[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) o[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
                    bf16* dst = Y + row * ldy + col;
                    if ((((uintptr_t)dst) & 31) == 0) {
                        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                                     ::"l"(dst), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7])
                                     : "memory");
                    } else {
                        *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
                        *reinterpret_cast<uint4*>(dst + 8) = make_uint4(o[4], o[5], o[6], o[7]);
                    }
                }
            }
            // this warp has read its quarter of the buffer (tcgen05.wait::ld inside tmem_ld16): hand it back to the MMA warp
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(tmem_empty + a)) : "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (stats != nullptr && my_tiles > 0) {
        for (int i = threadIdx.x; i < block_n; i += kThreads) {
            const float a = (s_stat[i] + s_stat[2 * block_n + i]) + (s_stat[4 * block_n + i] + s_stat[6 * block_n + i]);
            const float b = (s_stat[block_n + i] + s_stat[3 * block_n + i]) + (s_stat[5 * block_n + i] + s_stat[7 * block_n + i]);
            atomicAdd(stats + n0 + i, (double)a);
            atomicAdd(stats + stats_stride + n0 + i, (double)b);
        }
    }
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------ wgrad --------------
// dW[Nc][K] += dY[M][Nc]^T . X[M][K].  The reduction runs over pixels (M), so both operands are
// "MN-major" as they sit in HBM: a TMA box of 64 pixels x 64 channels (128-byte swizzle) IS the
// canonical MN-major UMMA atom stack (8 pixel rows x 128 B per atom, atoms 1024 B apart, the
// next 64-channel group one box further = leading byte offset).  One CTA owns a 128 (n) x
// BLOCK_K (k <= 256) tile of dW and a contiguous range of pixels; the fp32 accumulator lives
// in TMEM for the whole range and is added to global memory once, with 128-bit vector reductions.
constexpr int WM = 64;                           // pixels per stage
constexpr uint32_t kBoxBytes = WM * 64 * 2;      // one 64-pixel x 64-channel box

__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(kBoxBytes >> 4) << 16;       // leading byte offset: next 64-channel group
    d |= (uint64_t)(1024 >> 4) << 32;            // stride byte offset: next 8-pixel atom
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kThreads)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX,
                float* __restrict__ dW, int64_t M, int K, int Nc, int block_k, int n_groups, int k_groups,
                int64_t rows_per, int stages, uint32_t tmem_cols) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t a_bytes = (uint32_t)n_groups * kBoxBytes, b_bytes = (uint32_t)k_groups * kBoxBytes;
    uint8_t* sA = smem;
    uint8_t* sB = smem + (size_t)stages * a_bytes;
    uint64_t* bars = (uint64_t*)(sB + (size_t)stages * b_bytes);
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * stages + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * 128, k0 = blockIdx.y * block_k;
    const int64_t mb = (int64_t)blockIdx.z * rows_per;
    const int64_t me = mb + rows_per < M ? mb + rows_per : M;
    const int num_it = (int)((me - mb + WM - 1) / WM);

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(smem_u32(bars + s), 1);
            mbar_init(smem_u32(bars + stages + s), 1);
        }
        mbar_init(smem_u32(bars + 2 * stages), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) {tma_prefetch_desc(&tmG); tma_prefetch_desc(&tmX); }      // descriptor fetch (~0.5 us) under the predecessor's tail
    pdl_wait();          // everything above is on-chip setup; global memory is touched from here on

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < num_it; ++it) {
                const int s = it % stages;
                const uint32_t phase = (it / stages) & 1;
                mbar_wait(smem_u32(bars + stages + s), phase ^ 1);
                const uint32_t full = smem_u32(bars + s);
                mbar_expect_tx(full, a_bytes + b_bytes);
                const int m = (int)(mb + (int64_t)it * WM);
                for (int g = 0; g < n_groups; ++g)
                    tma_load_2d(smem_u32(sA + (size_t)s * a_bytes + (size_t)g * kBoxBytes), &tmG, full, n0 + g * 64, m);
                for (int g = 0; g < k_groups; ++g)
                    tma_load_2d(smem_u32(sB + (size_t)s * b_bytes + (size_t)g * kBoxBytes), &tmX, full, k0 + g * 64, m);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // D=f32, A=B=bf16, both MN-major (bits 15, 16), N=block_k, M=128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                   ((uint32_t)(block_k >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            for (int it = 0; it < num_it; ++it) {
                const int s = it % stages;
                const uint32_t phase = (it / stages) & 1;
                mbar_wait(smem_u32(bars + s), phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t adesc = make_desc_mn_sw128(smem_u32(sA + (size_t)s * a_bytes));
                const uint64_t bdesc = make_desc_mn_sw128(smem_u32(sB + (size_t)s * b_bytes));
                int64_t left = me - (mb + (int64_t)it * WM);
                const int k16 = left >= WM ? WM / 16 : (int)((left + 15) / 16);
                for (int k = 0; k < k16; ++k)                 // 16 pixels = 2 atoms = 2048 B per slice
                    umma_bf16(tmem_base, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), idesc, (uint32_t)(it > 0 || k > 0));
                umma_commit(smem_u32(bars + stages + s));
            }
            umma_commit(smem_u32(bars + 2 * stages));
        }
    } else {
        const int q = warp & 3;
        const int n = n0 + q * 32 + lane;
        mbar_wait(smem_u32(bars + 2 * stages), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int c = 0; c < block_k; c += 16) {
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
            if (n < Nc && num_it > 0) {
                float* dst = dW + (int64_t)n * K + k0 + c;
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    if (k0 + c + j < K) red_add_v4(dst + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------ host side ----------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_tiled() {
    tss_bind_context();
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// 2-D bf16 row-major [rows][cols] with row pitch ld (elements); box = box_rows x 64 columns, 128-B swizzle
int make_map(CUtensorMap* map, const void* base, int64_t rows, int cols, int64_t ld, int box_rows) {
    EncodeTiledFn enc = get_encode_tiled();
    TSS_REQUIRE(enc != nullptr, "pwconv_tc: cuTensorMapEncodeTiled is not available from the driver");
    TSS_REQUIRE(((uintptr_t)base & 15) == 0 && (ld * 2) % 16 == 0, "pwconv_tc: TMA needs 16-byte aligned base and pitch");
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TSS_REQUIRE(r == CUDA_SUCCESS, "pwconv_tc: cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%d ld=%lld box_rows=%d",
                (int)r, (long long)rows, cols, (long long)ld, box_rows);
    return TSS_OK;
}

// Tile width along the output channels: the largest multiple-of-16 divisor of Nc up to a cap.  The
// forward / dgrad GEMMs are bound by their epilogue (TMEM -> registers -> statistics -> stores), so
// narrow tiles (cap 64: more, smaller CTAs co-resident per SM, 64 TMEM columns each) beat wide ones:
// measured on B200 for the whole training step 4.80 ms (cap 256), 4.74 (128), 4.69 (96), 4.65 (64),
// 4.75 (48), 4.82 (32).  The re-read of the A tile by the other column tiles comes from L2.
int pick_block_n(int Nc, int cap) {
    for (int bn = cap; bn >= 16; bn -= 16)
        if (Nc % bn == 0) return bn;
    return 0;
}
int fwd_tile_cap() {
    static int cap = [] { const char* e = getenv("TSS_PW_BN_CAP"); return e ? atoi(e) : 64; }();
    return cap;
}
int wgrad_tile_cap() {
    static int cap = [] { const char* e = getenv("TSS_PW_WGRAD_CAP"); return e ? atoi(e) : 256; }();
    return cap;
}

}  // namespace

int tss_pwconv_wgrad_tc(const void* x, const void* dy, float* dw, int64_t M, int K, int Nc, int64_t ldx,
                        int64_t lddy, cudaStream_t st) {
    // Nc is the M dimension of the MMA (128 rows per CTA): any Nc works, TMA zero-fills the columns of dy beyond it
    TSS_REQUIRE(Nc >= 1 && K % 16 == 0, "pwconv_wgrad_tc: needs K %% 16 == 0 (Nc=%d K=%d)", Nc, K);
    TSS_REQUIRE(((uintptr_t)dw & 15) == 0, "pwconv_wgrad_tc: dw must be 16-byte aligned");
    const int bk = pick_block_n(K, wgrad_tile_cap());
    TSS_REQUIRE(bk >= 16, "pwconv_wgrad_tc: no tile width for K=%d", K);
    CUtensorMap tmG, tmX;
    if (int e = make_map(&tmG, dy, M, Nc, lddy, WM)) return e;
    if (int e = make_map(&tmX, x, M, K, ldx, WM)) return e;
    const int gx = (Nc + 127) / 128, gy = K / bk;
    const int n_groups = 2, k_groups = (bk + 63) / 64;
    static const int split_env = [] { const char* e = getenv("TSS_WGRAD_SPLIT"); return e ? atoi(e) : 2; }();       // CTAs per SM (A/B)
    int64_t nsplit = ((int64_t)tss_num_sms() * split_env) / ((int64_t)gx * gy);
    const int64_t max_split = ceil_div64(M, 512);
    if (nsplit > max_split) nsplit = max_split;
    if (nsplit < 1) nsplit = 1;
    const int64_t rows_per = ceil_div64(ceil_div64(M, nsplit), WM) * WM;
    nsplit = ceil_div64(M, rows_per);
    // 3 stages: two CTAs of the K = 128 layers (32 KB per stage) fit one SM, so the epilogue (REDs) of one overlaps the stream of the
    // other; whole step on B200 3.176 ms (4 stages) / 3.167 (3) / 3.179 (2) / 3.232 (6).  TSS_WGRAD_SPLIT (CTAs per SM): 1 and 2 equal, 3 slower.
    static const int stages_env = [] { const char* e = getenv("TSS_WGRAD_STAGES"); return e ? atoi(e) : 3; }();
    int stages = stages_env;
    while (stages > 2 && 1024 + (size_t)stages * (n_groups + k_groups) * kBoxBytes + (2 * stages + 1) * 8 + 16 > 200 * 1024) --stages;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < bk) tmem_cols <<= 1;
    const size_t smem = 1024 + (size_t)stages * (n_groups + k_groups) * kBoxBytes + (2 * stages + 1) * 8 + 16;
    static bool attr_set = false;
    if (!attr_set) {
        TSS_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    dim3 grid(gx, gy, (unsigned)nsplit);
    tss_launch(wgrad_tc_kernel, grid, kThreads, smem, st, tmG, tmX, dw, M, K, Nc, bk, n_groups, k_groups, rows_per, stages, tmem_cols);
    TSS_LAUNCH_CHECK("pwconv_wgrad_tc");
    return TSS_OK;
}

int tss_pwconv_fwd_tc(const void* x, const void* wp, void* y, int64_t M, int K, int Nc, int64_t ldx,
                      int64_t ldy, const float* scale, const float* shift, const void* res, int64_t ldr,
                      int flags, double* stats, cudaStream_t st) {
    TSS_REQUIRE(Nc % 16 == 0 && K % 8 == 0, "pwconv_tc: needs Nc %% 16 == 0 and K %% 8 == 0 (Nc=%d K=%d)", Nc, K);
    TSS_REQUIRE(ldy % 8 == 0 && (res == nullptr || ldr % 8 == 0), "pwconv_tc: output / residual pitch must be a multiple of 8");
    TSS_REQUIRE(((uintptr_t)y & 15) == 0 && ((uintptr_t)res & 15) == 0, "pwconv_tc: output / residual must be 16-byte aligned");
    TSS_REQUIRE(scale == nullptr || shift != nullptr, "pwconv_tc: scale without shift");
    int bn = pick_block_n(Nc, fwd_tile_cap());
    TSS_REQUIRE(bn >= 16, "pwconv_tc: no tile width for Nc=%d", Nc);
    // Small maps (1/16, 1/32 resolution: 216 / 54 row tiles) leave most of the 148 SMs without a CTA at the default tile
    // width.  TSS_PW_FILL=f narrows the column tile until there are at least f CTAs per SM (the A tile is re-read from
    // L2); measured on B200 for the whole step: 3.798 ms (off, the default), 3.814 (f = 2), 3.842 (f = 4) -- these
    // layers are bound by the latency of one CTA's TMA -> MMA -> epilogue chain, not by the number of CTAs.
    {
        static const int fill = [] { const char* e = getenv("TSS_PW_FILL"); return e ? atoi(e) : 0; }();
        const int64_t m_tiles = ceil_div64(M, BM);
        while (fill > 0 && bn > 16 && m_tiles * (Nc / bn) < (int64_t)fill * tss_num_sms()) {
            const int nb = pick_block_n(Nc, bn - 16);
            if (nb < 16) break;
            bn = nb;
        }
    }
    CUtensorMap tmA, tmB;
    if (int e = make_map(&tmA, x, M, K, ldx, BM)) return e;
    if (int e = make_map(&tmB, wp, Nc, K, K, bn)) return e;
    const int num_kb = (K + BK - 1) / BK;
    // 0 = never, 1 = where it measured faster (the default), 2 = always.  Per-shape times on B200 (tools/time_ops.py, us,
    // one-tile-per-CTA -> persistent): forward 32->48 @1/4 62 -> 35, 48->64 @1/8 19 -> 15, dgrad into 64 channels @1/8 26 -> 22,
    // 128->128 @1/8 dgrad 18 -> 14; but forward 64->384 @1/8 (6 column tiles) 50 -> 105, 96->576 @1/32 8 -> 12: with several
    // column tiles the epilogue of a tile is the bound and 2 resident persistent CTAs drain fewer tiles at a time than 6
    // one-tile CTAs.  So: at most two column tiles and at least two waves of row tiles.
    static const int persist = [] { const char* e = getenv("TSS_PW_PERSIST"); return e != nullptr ? atoi(e) : 1; }();
    const int64_t m_tiles = ceil_div64(M, BM);
    // ... and, since the 2-stage ring (three or four resident CTAs when a tile is ONE k-block), K <= 64 with any number of column
    // tiles: forward 64->384 @1/8 47.7 -> 43.7, @1/16 14.5 -> 12.9 (TSS_PW_PERSIST_K64=0 for the A/B)
    static const int persist_k64 = [] { const char* e = getenv("TSS_PW_PERSIST_K64"); return e != nullptr ? atoi(e) : 1; }();
    const bool persist_wins = (Nc / bn <= 2 && m_tiles >= 4 * (int64_t)tss_num_sms()) ||
                              (persist_k64 && num_kb == 1 && m_tiles >= 200);
    if ((persist == 2 || (persist == 1 && persist_wins)) && m_tiles < (1ll << 30)) {
        // persistent CTAs, double-buffered TMEM accumulator, statistics flushed once per CTA
        static const int stages_env = [] { const char* e = getenv("TSS_PW_STAGES"); return e ? atoi(e) : 0; }();
        // the ring only has to cover the TMA latency across tiles: 2 stages when a tile is one k-block (4 CTAs of 48 KB
        // per SM instead of 2 of 96 KB: twice the epilogue warps), 4 otherwise
        const int stages = stages_env > 0 ? stages_env : (num_kb == 1 ? 2 : 4);
        uint32_t tmem_cols = 32;
        while ((int)tmem_cols < 2 * bn) tmem_cols <<= 1;
        // statistics: column sums through a dedicated shared-memory scratch (TSS_PW_CS=0: the shuffle transposes)
        static const int cs_env = [] { const char* e = getenv("TSS_PW_CS"); return e ? atoi(e) : 1; }();
        const int cs_smem = (stats != nullptr && cs_env != 0) ? 1 : 0;
        const size_t smem = 1024 + (size_t)stages * (kABytes + (size_t)bn * BK * 2) + (2 * stages + 4) * 8 + 8 + 16 + 8 * bn * sizeof(float) +
                            (cs_smem ? (size_t)4 * kCsArray * sizeof(float) : 0);
        static bool attr_set_p = false;
        if (!attr_set_p) {
            TSS_CUDA(cudaFuncSetAttribute(pw_tc_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            attr_set_p = true;
        }
        static const int per_sm_env = [] { const char* e = getenv("TSS_PW_CTAS_PER_SM"); return e ? atoi(e) : 0; }();
        const int per_sm = per_sm_env > 0 ? per_sm_env : (stages <= 2 ? (cs_smem ? 3 : 4) : 2);      // (the scratch leaves room for 3)
        const int n_tiles = Nc / bn;
        int64_t gx = ((int64_t)tss_num_sms() * per_sm + n_tiles - 1) / n_tiles;     // resident CTAs shared by the column tiles
        if (gx > m_tiles) gx = m_tiles;
        if (gx < 1) gx = 1;
        dim3 grid((unsigned)gx, (unsigned)n_tiles);
        tss_launch(pw_tc_persistent_kernel, grid, kThreads, smem, st, tmA, tmB, (bf16*)y, M, K, ldy, bn, stages, tmem_cols, scale, shift,
                   (const bf16*)res, ldr, flags & TSS_EPI_RELU, stats, Nc, (int)m_tiles, cs_smem);
        TSS_LAUNCH_CHECK("pwconv_fwd_tc(persistent)");
        return TSS_OK;
    }
    int stages = num_kb < 4 ? num_kb : 4;
    // the statistics epilogue's column-sum scratch (4 warps x 2 x kCsArray floats = 20 KB) aliases the ring
    while (stats != nullptr && (size_t)stages * (kABytes + (size_t)bn * BK * 2) < (size_t)8 * kCsArray * sizeof(float)) ++stages;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < bn) tmem_cols <<= 1;
    const size_t smem = 1024 + (size_t)stages * (kABytes + (size_t)bn * BK * 2) + (2 * stages + 1) * 8 + 8 + 8 * bn * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        TSS_CUDA(cudaFuncSetAttribute(pw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    dim3 grid((unsigned)ceil_div64(M, BM), (unsigned)(Nc / bn));
    tss_launch(pw_tc_kernel, grid, kThreads, smem, st, tmA, tmB, (bf16*)y, M, K, ldy, bn, stages, tmem_cols, scale, shift,
                                               (const bf16*)res, ldr, flags & TSS_EPI_RELU, stats, Nc);
    TSS_LAUNCH_CHECK("pwconv_fwd_tc");
    return TSS_OK;
}
