// SIMT emulation of the CUDA subset the library's plain (non-TMA, non-tcgen05) kernels use, for HOST builds of
// the .cu files.  Test infrastructure (tests/test_simt_emulation.py): the kernels' index arithmetic, shared
// memory choreography and launch configurations are exercised on a box without a GPU by running every CUDA
// thread of a block as an OS thread: __syncthreads is a barrier over the live threads of the block (a thread
// that returns drops out, like on hardware), warp shuffles exchange through per-warp slots, blocks run one
// after the other.  Not a performance model and not a product path; the product has no CPU path.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <barrier>
#include <memory>
#include <thread>
#include <tuple>
#include <utility>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __maxnreg__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static
#define __grid_constant__

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(16) longlong2 { long long x, y; };
inline float2 make_float2(float x, float y) { return {x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return {x, y, z, w}; }
inline longlong2 make_longlong2(long long x, long long y) { return {x, y}; }

namespace tss_emu {
struct Block {
    std::barrier<> all;
    std::vector<std::unique_ptr<std::barrier<>>> warps;
    std::vector<uint32_t> slots;
    explicit Block(int threads) : all(threads), slots(2 * ((threads + 31) / 32) * 32) {
        for (int w = 0; w < (threads + 31) / 32; ++w) {
            const int n = threads - w * 32 < 32 ? threads - w * 32 : 32;
            warps.emplace_back(new std::barrier<>(n));
        }
    }
};
extern thread_local Block* block;
extern thread_local unsigned slot_parity;
extern unsigned char dyn_smem[256 * 1024];
}  // namespace tss_emu
extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;
// dynamic shared memory of the running block (kernels declare it through TSS_DYN_SMEM, common.cuh)
#define TSS_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(tss_emu::dyn_smem)

inline void __syncthreads() { tss_emu::block->all.arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { tss_emu::block->warps[threadIdx.x >> 5]->arrive_and_wait(); }

template <typename V> inline V __shfl_xor_sync(unsigned, V v, int lane_mask) {
    static_assert(sizeof(V) == 4, "32-bit shuffles only");
    // one barrier per shuffle: the exchange slots are double-buffered on the call parity, so a lane that races
    // ahead into the next shuffle writes the other buffer and cannot clobber a value a slower lane still reads
    unsigned& parity = tss_emu::slot_parity;             // shared by every warp-exchange primitive of the thread
    tss_emu::Block* b = tss_emu::block;
    const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint32_t* slots = b->slots.data() + (size_t)parity * (b->slots.size() / 2) + warp * 32;
    parity ^= 1u;
    memcpy(&slots[lane], &v, 4);
    b->warps[warp]->arrive_and_wait();
    V out;
    memcpy(&out, &slots[lane ^ (unsigned)lane_mask], 4);
    return out;
}

// all 32 values of the warp (lanes of a partial last warp that do not exist read as absent)
inline void tss_emu_warp_gather(uint32_t v, uint32_t (&all)[32], int& nlanes) {
    unsigned& parity = tss_emu::slot_parity;
    tss_emu::Block* b = tss_emu::block;
    const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint32_t* slots = b->slots.data() + (size_t)parity * (b->slots.size() / 2) + warp * 32;
    parity ^= 1u;
    slots[lane] = v;
    b->warps[warp]->arrive_and_wait();
    const int total = (int)blockDim.x - (int)warp * 32;
    nlanes = total < 32 ? total : 32;
    for (int i = 0; i < nlanes; ++i) all[i] = slots[i];
    b->warps[warp]->arrive_and_wait();            // gathers read every slot: nobody may overwrite before all have read
}
inline unsigned __ballot_sync(unsigned, int pred) {
    uint32_t all[32]; int n;
    tss_emu_warp_gather(pred ? 1u : 0u, all, n);
    unsigned r = 0;
    for (int i = 0; i < n; ++i) r |= (all[i] ? 1u : 0u) << i;
    return r;
}
inline unsigned __match_any_sync(unsigned, int v) {
    uint32_t all[32]; int n;
    tss_emu_warp_gather((uint32_t)v, all, n);
    unsigned r = 0;
    for (int i = 0; i < n; ++i) r |= (all[i] == (uint32_t)v ? 1u : 0u) << i;
    return r;
}
inline unsigned __match_all_sync(unsigned, int v, int* pred) {
    uint32_t all[32]; int n;
    tss_emu_warp_gather((uint32_t)v, all, n);
    bool same = true;
    for (int i = 0; i < n; ++i) same = same && all[i] == (uint32_t)v;
    *pred = same ? 1 : 0;
    return same ? (n == 32 ? 0xffffffffu : ((1u << n) - 1u)) : 0u;
}
inline unsigned __reduce_add_sync(unsigned, unsigned v) {
    uint32_t all[32]; int n;
    tss_emu_warp_gather(v, all, n);
    unsigned r = 0;
    for (int i = 0; i < n; ++i) r += all[i];
    return r;
}
inline int __reduce_add_sync(unsigned m, int v) { return (int)__reduce_add_sync(m, (unsigned)v); }
inline float __expf(float x) { return expf(x); }
inline float __logf(float x) { return logf(x); }
inline float __log2f(float x) { return log2f(x); }
inline float __exp2f(float x) { return exp2f(x); }
inline int __ffs(unsigned v) { return v ? __builtin_ctz(v) + 1 : 0; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return std::atomic_ref<unsigned>(*p).fetch_add(v); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return std::atomic_ref<unsigned long long>(*p).fetch_add(v); }

template <typename V> inline V __ldg(const V* p) { return *p; }
template <typename V> inline V __ldcg(const V* p) { return *p; }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return {fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline int __float2int_rn(float a) { return (int)lrintf(a); }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline float atomicAdd(float* p, float v) {
    std::atomic_ref<float> r(*p);
    float old = r.load();
    while (!r.compare_exchange_weak(old, old + v)) {}
    return old;
}
inline int atomicAdd(int* p, int v) { return std::atomic_ref<int>(*p).fetch_add(v); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline double atomicAdd(double* p, double v) {
    std::atomic_ref<double> r(*p);
    double old = r.load();
    while (!r.compare_exchange_weak(old, old + v)) {}
    return old;
}

// ---- the slice of the runtime API the launchers touch
typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
enum cudaLaunchAttributeID { cudaLaunchAttributeProgrammaticStreamSerialization = 3 };
struct cudaLaunchAttribute {
    cudaLaunchAttributeID id;
    struct { int programmaticStreamSerializationAllowed; } val;
};
struct cudaLaunchConfig_t {
    dim3 gridDim, blockDim;
    size_t dynamicSmemBytes;
    cudaStream_t stream;
    cudaLaunchAttribute* attrs;
    unsigned numAttrs;
};
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 4; return cudaSuccess; }   // 4 "SMs": small grids
template <typename F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

template <typename... KArgs, typename... Args>
inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t* cfg, void (*kernel)(KArgs...), Args&&... args) {
    std::tuple<KArgs...> params(static_cast<KArgs>(args)...);
    const dim3 grid = cfg->gridDim, blk = cfg->blockDim;
    if (blk.y != 1 || blk.z != 1 || cfg->dynamicSmemBytes > sizeof(tss_emu::dyn_smem)) return 1;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                tss_emu::Block block((int)blk.x);
                std::vector<std::thread> threads;
                threads.reserve(blk.x);
                for (unsigned t = 0; t < blk.x; ++t)
                    threads.emplace_back([&, t] {
                        tss_emu::block = &block;
                        threadIdx = {t, 0, 0};
                        blockIdx = {bx, by, bz};
                        blockDim = blk;
                        gridDim = grid;
                        std::apply(kernel, params);
                        block.warps[t >> 5]->arrive_and_drop();      // an exited thread no longer takes part in barriers
                        block.all.arrive_and_drop();
                    });
                for (auto& th : threads) th.join();
            }
    return cudaSuccess;
}
