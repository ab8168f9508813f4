// Functional emulation of the tcgen05 / TMA / mbarrier subset wrapped by csrc/tc_ptx.cuh, for HOST builds on the
// SIMT emulation (cuda_runtime.h next to this file).  Test infrastructure.
//   * shared addresses are byte offsets into the emulated dynamic shared memory (256 KB: 18 bits, exactly what a
//     UMMA descriptor's start-address field holds);
//   * an mbarrier is {expected arrivals, pending arrivals, pending transaction bytes, phase} under one mutex;
//     waiters spin with yield;
//   * a TMA 2-D load copies the box into shared memory in the 128-byte-swizzled layout (16-byte chunk index XOR
//     row & 7 inside 1024-byte atoms, out-of-bounds elements zero) and completes the transaction bytes;
//   * tcgen05.mma decodes the instruction and matrix descriptors (K-major, SWIZZLE_128B, M x N x 16 per call, bf16
//     inputs, fp32 accumulate) and multiplies synchronously at issue into an emulated TMEM [128 lanes][512 cols];
//   * tcgen05.ld 32x32b.x16 gives thread i of a warp 16 consecutive columns of lane (lane base + i).
#pragma once
#include <mutex>

#include "cuda_runtime.h"
#include "cuda_bf16.h"
#include "../../torch_semantic_segmentation_b200/csrc/common.cuh"

// ---- the slice of the driver API the launchers touch (cuTensorMapEncodeTiled through cudaGetDriverEntryPoint)
typedef uint64_t cuuint64_t;
typedef uint32_t cuuint32_t;
typedef int CUresult;
enum { CUDA_SUCCESS = 0 };
enum CUtensorMapDataType { CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 = 9, CU_TENSOR_MAP_DATA_TYPE_FLOAT32 = 7 };
enum CUtensorMapInterleave { CU_TENSOR_MAP_INTERLEAVE_NONE = 0 };
enum CUtensorMapSwizzle { CU_TENSOR_MAP_SWIZZLE_NONE = 0, CU_TENSOR_MAP_SWIZZLE_128B = 3 };
enum CUtensorMapL2promotion { CU_TENSOR_MAP_L2_PROMOTION_L2_128B = 2 };
enum CUtensorMapFloatOOBfill { CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE = 0, CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA = 1 };
struct CUtensorMap {
    const unsigned char* base;
    uint32_t rank;
    uint64_t dim[4], stride[4];      // dim[0] = innermost extent (elements); stride[i] = byte pitch of dimension i (stride[0] = element)
    uint32_t box[4];
    int elem_bytes, swizzle, nan_fill;
};
inline CUresult tss_emu_encode_tiled(CUtensorMap* m, CUtensorMapDataType dt, cuuint32_t rank, void* base, const cuuint64_t* gdim,
                                     const cuuint64_t* gstr, const cuuint32_t* box, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle sw, CUtensorMapL2promotion, CUtensorMapFloatOOBfill fill) {
    if (rank < 2 || rank > 4) return 1;
    m->elem_bytes = dt == CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 ? 2 : 4;
    if (sw == CU_TENSOR_MAP_SWIZZLE_128B && (rank != 2 || box[0] * (uint32_t)m->elem_bytes != 128)) return 1;
    if (sw != CU_TENSOR_MAP_SWIZZLE_128B && sw != CU_TENSOR_MAP_SWIZZLE_NONE) return 1;
    if (((uintptr_t)base & 15) != 0) return 1;
    m->base = (const unsigned char*)base;
    m->rank = rank;
    m->stride[0] = (uint64_t)m->elem_bytes;
    for (uint32_t i = 0; i < 4; ++i) {
        m->dim[i] = i < rank ? gdim[i] : 1;
        m->box[i] = i < rank ? box[i] : 1;
        if (i >= 1) m->stride[i] = i < rank ? gstr[i - 1] : 0;
        if (i >= 1 && i < rank && gstr[i - 1] % 16 != 0) return 1;
    }
    m->swizzle = sw == CU_TENSOR_MAP_SWIZZLE_128B;
    m->nan_fill = fill == CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA;
    return CUDA_SUCCESS;
}
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0 };
enum { cudaEnableDefault = 0 };
inline cudaError_t cudaGetDriverEntryPoint(const char*, void** fn, int, cudaDriverEntryPointQueryResult* q) {
    *fn = (void*)&tss_emu_encode_tiled;
    *q = cudaDriverEntryPointSuccess;
    return cudaSuccess;
}

namespace tss_emu {
struct MBar { int count, pending, tx; uint32_t phase; };
extern MBar mbars[sizeof(dyn_smem) / 8];
extern std::mutex mbar_mutex;
extern float tmem[128][512];
inline void mbar_check(MBar& b) {                   // caller holds the mutex
    if (b.pending == 0 && b.tx == 0) { b.phase ^= 1u; b.pending = b.count; }
}
inline uint32_t swz128(uint32_t addr) { return addr ^ (((addr >> 7) & 7u) << 4); }
inline float bf16_at(uint32_t addr) {
    uint16_t h;
    memcpy(&h, dyn_smem + addr, 2);
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
}  // namespace tss_emu

inline uint32_t smem_u32(const void* p) { return (uint32_t)((const unsigned char*)p - tss_emu::dyn_smem); }
inline void mbar_init(uint32_t bar, uint32_t count) {
    std::lock_guard<std::mutex> l(tss_emu::mbar_mutex);
    tss_emu::mbars[bar >> 3] = {(int)count, (int)count, 0, 0u};
}
inline void mbar_init_fence() {}
inline void mbar_arrive(uint32_t bar) {
    std::lock_guard<std::mutex> l(tss_emu::mbar_mutex);
    tss_emu::MBar& b = tss_emu::mbars[bar >> 3];
    b.pending -= 1;
    tss_emu::mbar_check(b);
}
inline void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    std::lock_guard<std::mutex> l(tss_emu::mbar_mutex);
    tss_emu::MBar& b = tss_emu::mbars[bar >> 3];
    b.tx += (int)bytes;
    b.pending -= 1;
    tss_emu::mbar_check(b);
}
inline void mbar_wait(uint32_t bar, uint32_t parity) {
    for (;;) {
        {
            std::lock_guard<std::mutex> l(tss_emu::mbar_mutex);
            if ((tss_emu::mbars[bar >> 3].phase & 1u) != (parity & 1u)) return;     // the phase with this parity completed
        }
        std::this_thread::yield();
    }
}
inline void tma_load_nd(uint32_t dst, const CUtensorMap* map, uint32_t bar, const int (&c)[4]) {
    const uint32_t eb = (uint32_t)map->elem_bytes;
    uint32_t off = 0;                                    // dense box in shared memory, innermost dimension first
    for (uint32_t i3 = 0; i3 < map->box[3]; ++i3)
        for (uint32_t i2 = 0; i2 < map->box[2]; ++i2)
            for (uint32_t i1 = 0; i1 < map->box[1]; ++i1)
                for (uint32_t i0 = 0; i0 < map->box[0]; ++i0, off += eb) {
                    const int64_t x[4] = {(int64_t)c[0] + i0, (int64_t)c[1] + i1, (int64_t)c[2] + i2, (int64_t)c[3] + i3};
                    unsigned char v[4] = {0, 0, 0, 0};
                    bool in = true;
                    uint64_t src = 0;
                    for (int d = 0; d < 4; ++d) {
                        in = in && x[d] >= 0 && (uint64_t)x[d] < map->dim[d];
                        src += (uint64_t)x[d] * map->stride[d];
                    }
                    if (in) memcpy(v, map->base + src, eb);
                    else if (map->nan_fill) { const uint32_t q = eb == 2 ? 0x7fffu : 0x7fffffffu; memcpy(v, &q, eb); }   // quiet NaN
                    memcpy(tss_emu::dyn_smem + (map->swizzle ? tss_emu::swz128(dst + off) : dst + off), v, eb);
                }
    std::lock_guard<std::mutex> l(tss_emu::mbar_mutex);
    tss_emu::MBar& b = tss_emu::mbars[bar >> 3];
    b.tx -= (int)off;
    tss_emu::mbar_check(b);
}
inline void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    const int c[4] = {c0, c1, 0, 0};
    tma_load_nd(dst, map, bar, c);
}
inline void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    const int c[4] = {c0, c1, c2, c3};
    tma_load_nd(dst, map, bar, c);
}
inline void mbar_fence_init() {}
inline void fence_async_smem() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void tc_alloc(uint32_t slot, uint32_t) {
    const uint32_t base = 0;
    memcpy(tss_emu::dyn_smem + slot, &base, 4);
}
inline void tc_dealloc(uint32_t, uint32_t) {}
inline void tc_fence_before() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void tc_fence_after() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    const int N = (int)((idesc >> 17) & 0x3Fu) << 3, M = (int)((idesc >> 24) & 0x1Fu) << 4;
    const uint32_t a0 = (uint32_t)(adesc & 0x3FFFu) << 4, b0 = (uint32_t)(bdesc & 0x3FFFu) << 4;
    const uint32_t sbo_a = (uint32_t)((adesc >> 32) & 0x3FFFu) << 4, sbo_b = (uint32_t)((bdesc >> 32) & 0x3FFFu) << 4;
    const uint32_t col0 = tmem_d & 0xFFFFu;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            float acc = accumulate ? tss_emu::tmem[m][col0 + n] : 0.f;
            for (int k = 0; k < 16; ++k) {
                const float a = tss_emu::bf16_at(tss_emu::swz128(a0 + (m >> 3) * sbo_a + (m & 7) * 128 + k * 2));
                const float b = tss_emu::bf16_at(tss_emu::swz128(b0 + (n >> 3) * sbo_b + (n & 7) * 128 + k * 2));
                acc += a * b;
            }
            tss_emu::tmem[m][col0 + n] = acc;
        }
}
inline void umma_commit(uint32_t bar) { mbar_arrive(bar); }
inline void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    const uint32_t lane = (taddr >> 16) + (threadIdx.x & 31), col = taddr & 0xFFFFu;
    for (int i = 0; i < 16; ++i) v[i] = tss_emu::tmem[lane][col + i];
}
