// bfloat16 for the SIMT emulation (tests/simt_emu/cuda_runtime.h): round-to-nearest-even conversions.
#pragma once
#include <stdint.h>
#include <string.h>

struct __nv_bfloat16 { uint16_t x; };
struct __nv_bfloat162 { __nv_bfloat16 x, y; };

inline __nv_bfloat16 __float2bfloat16_rn(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    __nv_bfloat16 h;
    if ((u & 0x7fffffffu) > 0x7f800000u) { h.x = 0x7fff; return h; }       // NaN
    u += 0x7fffu + ((u >> 16) & 1u);
    h.x = (uint16_t)(u >> 16);
    return h;
}
inline float __bfloat162float(__nv_bfloat16 h) {
    uint32_t u = (uint32_t)h.x << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
inline __nv_bfloat162 __floats2bfloat162_rn(float lo, float hi) { return {__float2bfloat16_rn(lo), __float2bfloat16_rn(hi)}; }
struct float2;
template <typename F2> inline __nv_bfloat162 __float22bfloat162_rn(F2 v) { return {__float2bfloat16_rn(v.x), __float2bfloat16_rn(v.y)}; }
