// Per-pixel arithmetic of the device input pipeline (csrc/augment.cu), written once for the kernel and for a
// host build: tests/test_data_cpu.py compiles this header with g++ (-ffp-contract=off) and checks it against the
// OpenCV-generated golden vectors, so the fixed-point formulas are pinned even on a box without a GPU.  On the
// device every floating-point step uses the round-to-nearest intrinsics, which the compiler never contracts
// into FMAs; on the host the same steps are plain IEEE operations.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDA_ARCH__
#define TSS_HD __host__ __device__ __forceinline__
#define TSS_DADD(a, b) __dadd_rn(a, b)
#define TSS_DSUB(a, b) __dsub_rn(a, b)
#define TSS_DMUL(a, b) __dmul_rn(a, b)
#define TSS_DDIV(a, b) __ddiv_rn(a, b)
#define TSS_FSUB(a, b) __fsub_rn(a, b)
#define TSS_FMUL(a, b) __fmul_rn(a, b)
#define TSS_F2I_RN(a) __float2int_rn(a)
#define TSS_LDG(p) __ldg(p)
#else
#ifdef __CUDACC__
#define TSS_HD __host__ __device__ inline
#else
#define TSS_HD inline
#endif
#define TSS_DADD(a, b) ((a) + (b))
#define TSS_DSUB(a, b) ((a) - (b))
#define TSS_DMUL(a, b) ((a) * (b))
#define TSS_DDIV(a, b) ((a) / (b))
#define TSS_FSUB(a, b) ((a) - (b))
#define TSS_FMUL(a, b) ((a) * (b))
#define TSS_F2I_RN(a) ((int)lrintf(a))          /* default rounding mode: nearest even */
#define TSS_LDG(p) (*(p))
#endif

struct TssNorm3 {
    float mean[3];      // mean * 255
    float inv[3];       // 1 / (std * 255), rounded to fp32 on the host like np.reciprocal(float32)
};

struct TssTap {
    int i0, i1;         // clipped source indices
    int w0, w1;         // 11-bit weights
};

TSS_HD int tss_clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// cv::resize: scale = 1. / ((double)dsize / ssize)
TSS_HD double tss_resize_scale(int dn, int sn) { return TSS_DDIV(1.0, TSS_DDIV((double)dn, (double)sn)); }

// OpenCV INTER_LINEAR: fx = (float)((d + 0.5) * scale - 0.5); s = floor(fx); fx -= s; each operation rounded on
// its own.  Horizontal taps zero the fraction at a clamped index; vertical taps only clamp the row index.
TSS_HD TssTap tss_linear_tap(int d, double scale, int sn, bool clamp_fraction) {
    float f = (float)TSS_DSUB(TSS_DMUL(TSS_DADD((double)d, 0.5), scale), 0.5);
    int s = (int)floorf(f);
    f = TSS_FSUB(f, (float)s);
    if (clamp_fraction) {
        if (s < 0) { f = 0.f; s = 0; }
        if (s >= sn - 1) { f = 0.f; s = sn - 1; }
    }
    TssTap t;
    t.w1 = TSS_F2I_RN(TSS_FMUL(f, 2048.f));
    t.w0 = TSS_F2I_RN(TSS_FMUL(TSS_FSUB(1.f, f), 2048.f));
    t.i0 = tss_clampi(s, 0, sn - 1);
    t.i1 = tss_clampi(s + 1, 0, sn - 1);
    return t;
}

// OpenCV INTER_NEAREST: min(floor(d * scale), sn - 1) in double
TSS_HD int tss_nearest_tap(int d, double scale, int sn) {
    int s = (int)floor(TSS_DMUL((double)d, scale));
    return s < sn - 1 ? s : sn - 1;
}

// One channel of one output pixel: horizontal pass in int32, vertical pass with OpenCV's >>4, >>16, +2, >>2
// rounding to uint8, then albumentations' normalize as two separately rounded fp32 operations.
TSS_HD float tss_augment_value(const uint8_t* r0, const uint8_t* r1, const TssTap& tx, const TssTap& ty, int c,
                               float mean, float inv) {
    const int h0 = (int)TSS_LDG(r0 + tx.i0 * 3 + c) * tx.w0 + (int)TSS_LDG(r0 + tx.i1 * 3 + c) * tx.w1;
    const int h1 = (int)TSS_LDG(r1 + tx.i0 * 3 + c) * tx.w0 + (int)TSS_LDG(r1 + tx.i1 * 3 + c) * tx.w1;
    int v = (((ty.w0 * (h0 >> 4)) >> 16) + ((ty.w1 * (h1 >> 4)) >> 16) + 2) >> 2;
    v = tss_clampi(v, 0, 255);
    return TSS_FMUL(TSS_FSUB((float)v, mean), inv);
}
