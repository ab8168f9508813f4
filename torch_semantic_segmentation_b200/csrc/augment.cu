// Input pipeline on the device: what the reference's DataLoader workers do to one
// decoded Cityscapes sample on the CPU -- TRAIN_MAPPING (data/cityscapes.py:17-20,88), albu.RandomScale ->
// RandomCrop -> HorizontalFlip -> Normalize -> ToTensor (scripts/train_fastscnn.py:62-68) -- as ONE gather kernel
// over the output crop: each output pixel is traced back through flip, crop and scale to its 2x2 source taps in
// the uint8 frame.  The scaled image is never materialised, yet every output value is bit-identical to the CPU
// pipeline, because the arithmetic of OpenCV's uint8 INTER_LINEAR resize is reproduced exactly: coordinates in
// double then float, 11-bit coefficients rounded half-to-even, the horizontal pass in int32, the vertical pass
// with its (b * (row >> 4)) >> 16, +2, >> 2 rounding to uint8; then (u8 - mean*255) * (1 / (std*255)) in fp32
// as two separately rounded operations (no FMA).  Labels take OpenCV's INTER_NEAREST index and the id table.
//
// HBM: writes 12 B (fp32 x 3 planes) + 8 B (int64 label) per output pixel; reads <= 4 B / scale^2 per pixel of
// uint8 source through L1/L2.  A thread makes 4 consecutive output columns: one 16-byte store per plane and two
// for the labels; the random draws arrive as a small device table, so a batch is one launch.
#include <stdint.h>

#include "augment_math.h"
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

inline int stream_grid(int64_t items, int per_sm = 8) {
    int64_t want = ceil_div64(items, kThreads);
    int64_t cap = (int64_t)tss_num_sms() * per_sm;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

__global__ void __launch_bounds__(kThreads)
augment_kernel(const uint8_t* __restrict__ images, const uint8_t* __restrict__ labels, const int* __restrict__ geom,
               const int64_t* __restrict__ lut, TssNorm3 norm, float* __restrict__ out_image,
               int64_t* __restrict__ out_label, int N, int H, int W, int ch, int cw) {
    pdl_wait();
    const int groups = cw >> 2;
    const int64_t total = (int64_t)N * ch * groups;
    for (int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * kThreads) {
        const int gx = (int)(idx % groups);
        const int oy = (int)((idx / groups) % ch);
        const int n = (int)(idx / ((int64_t)groups * ch));
        const int* gm = geom + n * 5;
        const int nh = __ldg(gm), nw = __ldg(gm + 1), cy = __ldg(gm + 2), cx = __ldg(gm + 3), flip = __ldg(gm + 4);
        const double scale_y = tss_resize_scale(nh, H), scale_x = tss_resize_scale(nw, W);
        const int dy = oy + cy;
        const TssTap ty = tss_linear_tap(dy, scale_y, H, false);
        const uint8_t* img = images + (size_t)n * H * W * 3;
        const uint8_t* r0 = img + (size_t)ty.i0 * W * 3;
        const uint8_t* r1 = img + (size_t)ty.i1 * W * 3;
        const int ly = labels != nullptr ? tss_nearest_tap(dy, scale_y, H) : 0;
        float px[3][4];
        long long lab[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ox = gx * 4 + j;
            const int dx = (flip ? cw - 1 - ox : ox) + cx;
            const TssTap tx = tss_linear_tap(dx, scale_x, W, true);
#pragma unroll
            for (int c = 0; c < 3; ++c) px[c][j] = tss_augment_value(r0, r1, tx, ty, c, norm.mean[c], norm.inv[c]);
            if (labels != nullptr) {
                const int id = __ldg(labels + ((size_t)n * H + ly) * W + tss_nearest_tap(dx, scale_x, W));
                lab[j] = lut != nullptr ? __ldg(lut + id) : (long long)id;
            }
        }
        const size_t plane = (size_t)ch * cw;
        float* o = out_image + (size_t)n * 3 * plane + (size_t)oy * cw + gx * 4;
#pragma unroll
        for (int c = 0; c < 3; ++c)
            *reinterpret_cast<float4*>(o + c * plane) = make_float4(px[c][0], px[c][1], px[c][2], px[c][3]);
        if (labels != nullptr) {
            longlong2* l = reinterpret_cast<longlong2*>(out_label + (size_t)n * plane + (size_t)oy * cw + gx * 4);
            l[0] = make_longlong2(lab[0], lab[1]);
            l[1] = make_longlong2(lab[2], lab[3]);
        }
    }
}

}  // namespace

extern "C" int tss_augment_batch(const void* images, const void* labels, const int* geom, const int64_t* lut,
                                 const float* norm, float* out_image, int64_t* out_label, int N, int H, int W,
                                 int ch, int cw, void* stream) {
    TSS_REQUIRE(N > 0 && H > 1 && W > 1 && ch > 0 && cw > 0, "augment_batch: N=%d H=%d W=%d crop=%dx%d", N, H, W, ch, cw);
    TSS_REQUIRE(cw % 4 == 0, "augment_batch: the crop width %d must be a multiple of 4", cw);
    TSS_REQUIRE(images != nullptr && geom != nullptr && norm != nullptr && out_image != nullptr, "augment_batch: missing buffer");
    TSS_REQUIRE((labels == nullptr) == (out_label == nullptr), "augment_batch: labels and out_label go together");
    TSS_REQUIRE((((uintptr_t)out_image | (uintptr_t)out_label) & 15) == 0, "augment_batch: outputs must be 16-byte aligned");
    TssNorm3 nm;                                   // `norm` is a HOST array of 6 floats: mean*255 (3), 1/(std*255) (3)
    for (int c = 0; c < 3; ++c) {
        nm.mean[c] = norm[c];
        nm.inv[c] = norm[3 + c];
    }
    tss_launch(augment_kernel, stream_grid((int64_t)N * ch * (cw / 4)), kThreads, 0, (cudaStream_t)stream,
               (const uint8_t*)images, (const uint8_t*)labels, geom, lut, nm, out_image, out_label, N, H, W, ch, cw);
    TSS_LAUNCH_CHECK("augment_batch");
    return TSS_OK;
}
