// TMA / mbarrier helpers shared by the tensor-core GEMMs and the TMA-staged depthwise kernels.
#pragma once
#ifdef TSS_HOST_EMU            // tests/simt_emu: functional emulation of the same functions for host builds
#include "tcgen05_emu.h"
#else
#include <cuda.h>

#include "common.cuh"
#endif

typedef CUresult (*TssEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
static inline TssEncodeTiledFn tss_encode_tiled() {
    tss_bind_context();
    static TssEncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (TssEncodeTiledFn)p;
    }();
    return fn;
}

#ifndef TSS_HOST_EMU
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

#endif  // !TSS_HOST_EMU

// ---------------------------------------------------------------- depthwise taps --------
// The CB x 9 taps of one channel block, for the TMA-staged depthwise kernels.  Every thread used to fetch its 72 taps
// with scalar loads 36 bytes apart (8 sectors per request, ~2 us of L1 traffic per CTA: tools/trace_kernels.py); now the
// CTA copies the block with coalesced loads into shared memory as [9][CB] (tap-major, flipped for a dgrad) and each
// thread takes 8 channels x 9 taps with 128-bit reads.  Parameters are written by the optimizer's kernels only, which
// do not release their dependents early (optim.cu), so the copy may run BEFORE griddepcontrol.wait.
template <bool FLIP>
__device__ __forceinline__ void dw_stage_taps(const float* __restrict__ w, int cb0, int CB, float* s_w) {
    const float* src = w + (size_t)cb0 * 9;
    const int n = CB * 9;
    // eight independent loads in flight per thread and round (a plain load -> store loop pays one global round trip per
    // iteration: 2 us for 5 iterations in the trace)
    for (int base = 0; base < n; base += 8 * (int)blockDim.x) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = base + u * (int)blockDim.x + (int)threadIdx.x;
            v[u] = i < n ? __ldg(src + i) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = base + u * (int)blockDim.x + (int)threadIdx.x;
            if (i < n) {
                const int c = i / 9, k = i - c * 9;
                s_w[(FLIP ? 8 - k : k) * CB + c] = v[u];
            }
        }
    }
}
__device__ __forceinline__ void dw_take_taps(const float* s_w, int CB, int cg, float2 (&wr)[9][4]) {
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(s_w + k * CB + cg * 8);
        const float4 b = *reinterpret_cast<const float4*>(s_w + k * CB + cg * 8 + 4);
        wr[k][0] = make_float2(a.x, a.y);
        wr[k][1] = make_float2(a.z, a.w);
        wr[k][2] = make_float2(b.x, b.y);
        wr[k][3] = make_float2(b.z, b.w);
    }
}

