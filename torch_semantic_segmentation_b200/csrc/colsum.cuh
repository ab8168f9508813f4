// Column sums over the 32 rows a warp holds in a tensor-core epilogue (one accumulator row per lane, 16 columns per
// tcgen05.ld), through a warp-private shared-memory scratch instead of shuffle trees.
//
// The shuffle transpose (warp_transpose_sum16) costs 15 SHFL + 30 SEL + 15 FADD per thread for every 16 x 32 block and
// sum, and an epilogue needs two sums per block (BatchNorm statistics: sum y, sum y^2; BatchNorm backward: sum g, sum g x-hat):
// ~135 issue slots per thread and block against ~30 for everything else the epilogue does (tools/trace_kernels.py: the
// 64-column epilogue of the fused dgrad took 6.1 us of a 13.9 us kernel at 1/32 resolution).  Here every lane writes its
// 16 values as four 128-bit stores into a [32][20]-float array (row pitch 20 words: the eight lanes of a store phase hit
// eight distinct bank quads), and after a __syncwarp every lane sums one column over 16 or 32 rows with conflict-free
// 32-bit loads: 4 STS.128 + 16..32 LDS + as many FADD/FFMA per thread.
#pragma once

constexpr int kCsPitch = 20;                    // floats per scratch row: 16 values + 4 pad
constexpr int kCsArray = 32 * kCsPitch;         // one 32 x 16 block
constexpr int kCsPair = 2 * kCsArray + 16;      // two blocks, the second shifted by 16 banks

__device__ __forceinline__ void cs_store16(float* buf, int lane, const float (&v)[16]) {
    float4* p = reinterpret_cast<float4*>(buf + lane * kCsPitch);
    p[0] = make_float4(v[0], v[1], v[2], v[3]);
    p[1] = make_float4(v[4], v[5], v[6], v[7]);
    p[2] = make_float4(v[8], v[9], v[10], v[11]);
    p[3] = make_float4(v[12], v[13], v[14], v[15]);
}

// One block: lane (c = lane & 15, h = lane >> 4) sums column c over the 16 rows {4h + (i & 3) + 8 (i >> 2)} -- the two
// half-warps read rows 4 apart = 80 words = 16 banks apart -- and the halves are combined with one shuffle per sum.
// Every lane returns the full sums of column (lane & 15): s1 = sum x, s2 = sum x^2.
__device__ __forceinline__ void cs_sum_sq(const float* buf, int lane, float& s1, float& s2) {
    const float* p = buf + 4 * (lane >> 4) * kCsPitch + (lane & 15);
    float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
        const float x = p[((i & 3) + 8 * (i >> 2)) * kCsPitch];
        const float y = p[(((i + 1) & 3) + 8 * ((i + 1) >> 2)) * kCsPitch];
        a0 += x;
        b0 = fmaf(x, x, b0);
        a1 += y;
        b1 = fmaf(y, y, b1);
    }
    const float a = a0 + a1, b = b0 + b1;
    s1 = a + __shfl_xor_sync(0xffffffffu, a, 16);
    s2 = b + __shfl_xor_sync(0xffffffffu, b, 16);
}

// Two blocks (buf and buf + kCsArray + 16): lanes 0..15 return the sum of column `lane` of the first block over all 32
// rows, lanes 16..31 that of column `lane - 16` of the second block.
__device__ __forceinline__ float cs_sum_pair(const float* buf, int lane) {
    const float* p = buf + (lane >> 4) * (kCsArray + 16) + (lane & 15);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int r = 0; r < 32; r += 4) {
        a0 += p[r * kCsPitch];
        a1 += p[(r + 1) * kCsPitch];
        a2 += p[(r + 2) * kCsPitch];
        a3 += p[(r + 3) * kCsPitch];
    }
    return (a0 + a1) + (a2 + a3);
}
