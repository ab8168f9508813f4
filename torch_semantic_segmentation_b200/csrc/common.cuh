// Shared device/host helpers for libtss_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/tss_b200.h"

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------- errors ------------
void tss_set_error(const char* fmt, ...);
void tss_count_launch(int n);

#define TSS_REQUIRE(cond, ...)                                                        \
    do {                                                                              \
        if (!(cond)) {                                                                \
            tss_set_error(__VA_ARGS__);                                               \
            return TSS_ERR_ARG;                                                       \
        }                                                                             \
    } while (0)

#define TSS_LAUNCH_CHECK(name)                                                        \
    do {                                                                              \
        cudaError_t e_ = cudaGetLastError();                                          \
        if (e_ != cudaSuccess) {                                                      \
            tss_set_error("%s: launch failed: %s", name, cudaGetErrorString(e_));     \
            return TSS_ERR_CUDA;                                                      \
        }                                                                             \
        tss_count_launch(1);                                                          \
    } while (0)

#define TSS_CUDA(call)                                                                \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) {                                                      \
            tss_set_error("%s failed: %s", #call, cudaGetErrorString(e_));            \
            return TSS_ERR_CUDA;                                                      \
        }                                                                             \
    } while (0)

// Dispatch a launcher body on the activation dtype. `T` is float or bf16 inside BODY.
#define TSS_DISPATCH_DTYPE(dtype, name, ...)                                          \
    do {                                                                              \
        if ((dtype) == TSS_F32) { typedef float T; __VA_ARGS__ }                      \
        else if ((dtype) == TSS_BF16) { typedef bf16 T; __VA_ARGS__ }                 \
        else { tss_set_error("%s: unsupported dtype %d", name, (int)(dtype)); return TSS_ERR_ARG; } \
    } while (0)

// ---------------------------------------------------------------- launches ----------
// Every kernel of the library is launched with programmatic stream serialization (PDL): the next
// kernel of the stream (or graph) may be scheduled while the previous one drains, and each
// kernel executes griddepcontrol.wait (pdl_wait) before it touches global memory, which returns
// once the preceding grid has completed and its writes are visible.  On-chip setup (barrier
// init, TMEM allocation, shared-memory zeroing) sits in front of the wait and overlaps with the
// predecessor's tail.  A step is ~370 dependent launches of 5..50 us each, so the per-edge
// latency is a first-order term.  tss_set_pdl(0) falls back to plain serialized launches.
bool tss_pdl_enabled();

// wait for the preceding grid, then let the NEXT grid's CTAs become resident as soon as every CTA
// of this grid has started (its last wave is running): they fill the SM slots this grid's tail
// frees, parked in their own griddepcontrol.wait, and never starve this grid's own CTAs.
__device__ __forceinline__ void pdl_wait() {
#ifndef TSS_HOST_EMU
    asm volatile("griddepcontrol.wait;" ::: "memory");
#ifndef TSS_NO_PDL_TRIGGER
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
#endif
}

// The first TMA instruction that names a tensor map fetches its 128-byte descriptor (a kernel parameter, independent of the
// preceding grid): ~0.5 us in front of the first load of every CTA (tools/trace_kernels.py: "predecessor complete" ->
// "last TMA issued").  One thread requests it before griddepcontrol.wait so that the fetch overlaps the predecessor's tail.
__device__ __forceinline__ void tma_prefetch_desc(const void* map) {
#ifndef TSS_HOST_EMU
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<unsigned long long>(map)) : "memory");
#endif
}

// ---------------------------------------------------------------- in-kernel timeline ----
// `make trace` builds libtss_b200_trace.so with -DTSS_TRACE: TSS_MARK(slot) then stores %globaltimer (ns) of the calling
// thread into trace[cta * 16 + slot] (tools/trace_kernels.py reads the buffer and prints where a CTA's time goes).
// The product library is built without it: the marks compile to nothing.
#ifdef TSS_TRACE
#define TSS_TRACE_SLOTS 16
#define TSS_TRACE_MAX_CTAS 65536
unsigned long long* tss_trace_buffer_host();                 // api.cu (set by tss_trace_set)
static __device__ unsigned long long* g_tss_trace;           // one copy per translation unit, bound by tss_launch
__device__ __forceinline__ void tss_trace_mark(int slot) {
    unsigned long long* buf = g_tss_trace;
    if (buf == nullptr) return;
    const unsigned cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    if (cta >= TSS_TRACE_MAX_CTAS) return;
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    buf[(size_t)cta * TSS_TRACE_SLOTS + slot] = t;
    if (slot == 0) {
        unsigned sm;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        buf[(size_t)cta * TSS_TRACE_SLOTS + TSS_TRACE_SLOTS - 1] = sm;
    }
}
#define TSS_MARK(slot) do { if (threadIdx.x == 0) tss_trace_mark(slot); } while (0)
#define TSS_MARK_IF(cond, slot) do { if (cond) tss_trace_mark(slot); } while (0)
#else
#define TSS_MARK(slot) ((void)0)
#define TSS_MARK_IF(cond, slot) ((void)0)
#endif

// The same wait without releasing the dependents early, for the kernels that WRITE PARAMETERS (the optimizer): kernels may
// read parameters in front of their own wait (dw_stage_taps in tma.cuh), which is only sound if a kernel that is still
// writing them never has a dependent grid running beside it.
__device__ __forceinline__ void pdl_wait_hold() {
#ifndef TSS_HOST_EMU
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

// Dynamic shared memory of the CTA.  (tests/simt_emu/ redefines this for host builds of the plain SIMT kernels.)
#ifndef TSS_DYN_SMEM
#define TSS_DYN_SMEM(type, name) extern __shared__ __align__(16) type name[]
#endif

template <typename... KArgs, typename... Args>
static inline cudaError_t tss_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
#ifdef TSS_TRACE
    {   // bind this translation unit's trace pointer (synchronous copy: set the buffer and warm up outside graph capture)
        static unsigned long long* bound = reinterpret_cast<unsigned long long*>(-1);
        unsigned long long* want = tss_trace_buffer_host();
        if (want != bound) {
            cudaMemcpyToSymbol(g_tss_trace, &want, sizeof(want));
            bound = want;
        }
    }
#endif
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = tss_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Driver-API entry points (cuTensorMapEncodeTiled) need the primary context bound to the CALLING thread.  The runtime binds
// it lazily at a thread's first runtime call, and PyTorch does not call cudaSetDevice on a thread that stays on device 0:
// a descriptor encode that is the first CUDA call of autograd's backward worker fails with CUDA_ERROR_INVALID_CONTEXT
// (201).  One cudaFree(nullptr) per thread binds it.
#ifndef TSS_HOST_EMU
static inline void tss_bind_context() {
    static thread_local bool bound = false;
    if (!bound) { cudaFree(nullptr); bound = true; }
}
#else
static inline void tss_bind_context() {}
#endif

static inline int tss_num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
}

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------- scalar conversion --
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// ---------------------------------------------------------------- 8-wide vectors -----
// All NHWC activation kernels move 8 channels per thread: one 128-bit access for bf16,
// two for fp32.  Pointers must be 16-byte aligned (C % 8 == 0 and 16B-aligned base).
__device__ __forceinline__ void load8(const float* __restrict__ p, float (&v)[8]) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p));
    float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
    v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
    v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
    v[4] = __uint_as_float(u.z << 16); v[5] = __uint_as_float(u.z & 0xffff0000u);
    v[6] = __uint_as_float(u.w << 16); v[7] = __uint_as_float(u.w & 0xffff0000u);
}
__device__ __forceinline__ void load8(const bf16* __restrict__ p, float (&v)[8]) {
    uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    unpack8(u, v);
}
// same, from shared memory (plain loads: the read-only/LDG path is for global memory)
__device__ __forceinline__ void load8_smem(const float* p, float (&v)[8]) {
    const float4 a = reinterpret_cast<const float4*>(p)[0];
    const float4 b = reinterpret_cast<const float4*>(p)[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8_smem(const bf16* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    unpack8(u, v);
}
// ---------------------------------------------------------------- packed fp32x2 ------
// Blackwell issues two fp32 FMAs per lane with one FFMA2 (fma.rn.f32x2): the depthwise kernels are
// bound by instruction issue (9 FMAs per output element against ~5 bytes of traffic), so their
// inner loops run on float2 pairs.  Same IEEE result as two scalar fmaf.
struct f2x4 { float2 p[4]; };
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ void load8p_smem(const bf16* p, float2 (&v)[4]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    v[0] = make_float2(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u));
    v[1] = make_float2(__uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
    v[2] = make_float2(__uint_as_float(u.z << 16), __uint_as_float(u.z & 0xffff0000u));
    v[3] = make_float2(__uint_as_float(u.w << 16), __uint_as_float(u.w & 0xffff0000u));
}
__device__ __forceinline__ void load8p_smem(const float* p, float2 (&v)[4]) {
    const float4 a = reinterpret_cast<const float4*>(p)[0];
    const float4 b = reinterpret_cast<const float4*>(p)[1];
    v[0] = make_float2(a.x, a.y); v[1] = make_float2(a.z, a.w);
    v[2] = make_float2(b.x, b.y); v[3] = make_float2(b.z, b.w);
}
__device__ __forceinline__ void zero8p(float2 (&v)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = make_float2(0.f, 0.f);
}
__device__ __forceinline__ void store8p(float* __restrict__ p, const float2 (&v)[4]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[2].x, v[2].y, v[3].x, v[3].y);
}
__device__ __forceinline__ void store8p(bf16* __restrict__ p, const float2 (&v)[4]) {
    uint4 u;
    __nv_bfloat162 h;
    h = __float22bfloat162_rn(v[0]); u.x = *reinterpret_cast<uint32_t*>(&h);
    h = __float22bfloat162_rn(v[1]); u.y = *reinterpret_cast<uint32_t*>(&h);
    h = __float22bfloat162_rn(v[2]); u.z = *reinterpret_cast<uint32_t*>(&h);
    h = __float22bfloat162_rn(v[3]); u.w = *reinterpret_cast<uint32_t*>(&h);
    *reinterpret_cast<uint4*>(p) = u;
}

// raw 8-element vectors (loads issued early, unpacked late: keeps many loads in flight per thread
// without holding their fp32 expansions in registers)
template <typename T> struct Raw8;
template <> struct Raw8<bf16> {
    uint4 a;
    __device__ __forceinline__ void ld(const bf16* p) { a = __ldg(reinterpret_cast<const uint4*>(p)); }
    __device__ __forceinline__ void zero() { a = make_uint4(0, 0, 0, 0); }
    __device__ __forceinline__ void get(float (&v)[8]) const { unpack8(a, v); }
};
template <> struct Raw8<float> {
    float4 a, b;
    __device__ __forceinline__ void ld(const float* p) {
        a = __ldg(reinterpret_cast<const float4*>(p));
        b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    }
    __device__ __forceinline__ void zero() { a = make_float4(0, 0, 0, 0); b = a; }
    __device__ __forceinline__ void get(float (&v)[8]) const {
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
};

// 256-bit global accesses (sm_100: LDG.256 / STG.256): one full 32-byte sector per thread and request.  A tensor-core
// epilogue thread owns a ROW of the tile, so a warp-wide access touches 32 different lines whatever its width; the
// 256-bit form halves the number of requests the SM sends to L2 for the same bytes.  `p` must be 32-byte aligned.
__device__ __forceinline__ void ldg256(const void* p, uint4& a, uint4& b) {
#ifndef TSS_HOST_EMU
    asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
#else
    a = reinterpret_cast<const uint4*>(p)[0];
    b = reinterpret_cast<const uint4*>(p)[1];
#endif
}
__device__ __forceinline__ void stg256(void* p, const uint4& a, const uint4& b) {
#ifndef TSS_HOST_EMU
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
#else
    reinterpret_cast<uint4*>(p)[0] = a;
    reinterpret_cast<uint4*>(p)[1] = b;
#endif
}

__device__ __forceinline__ void zero8(float (&v)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void store8(float* __restrict__ p, const float (&v)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(bf16* __restrict__ p, const float (&v)[8]) {
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
    u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = u;
}
// value as it will be read back after the store (rounding to the storage type)
__device__ __forceinline__ float round_as(float v, const float*) { return v; }
__device__ __forceinline__ float round_as(float v, const bf16*) { return __bfloat162float(__float2bfloat16_rn(v)); }

// ---------------------------------------------------------------- reductions ---------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// lane l ends with sum over the 32 lanes of v[l]  (31 shuffles instead of 160)
__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32], int lane) {
#pragma unroll
    for (int step = 16, n = 32; step >= 1; step >>= 1, n >>= 1) {
        const bool upper = (lane & step) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            float send = upper ? v[i] : v[i + n / 2];
            float keep = upper ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
        }
    }
    return v[0];
}

// bilinear align_corners=True source index, same float arithmetic as ATen
// (area_pixel_compute_scale / compute_source_index): scale = (in-1)/(out-1) in fp32.
__host__ __device__ __forceinline__ float ac_scale(int in, int out) {
    return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
}
__device__ __forceinline__ void ac_source(float scale, int dst, int in_size, int& i0, int& i1, float& lam) {
    float s = scale * (float)dst;
    i0 = (int)s;
    if (i0 > in_size - 1) i0 = in_size - 1;   // guards fp rounding at the last pixel
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    lam = fminf(fmaxf(s - (float)i0, 0.f), 1.f);
}

// first output index whose source index floor can be >= i (conservative), given scale
__device__ __forceinline__ int first_candidate(float scale, int i, int out_size) {
    if (scale <= 0.f) return 0;
    int o = (int)floorf((float)(i - 1) / scale) - 1;
    return o < 0 ? 0 : (o > out_size - 1 ? out_size - 1 : o);
}
__device__ __forceinline__ int last_candidate(float scale, int i, int out_size) {
    if (scale <= 0.f) return out_size - 1;
    int o = (int)ceilf((float)(i + 1) / scale) + 1;
    return o > out_size - 1 ? out_size - 1 : o;
}

