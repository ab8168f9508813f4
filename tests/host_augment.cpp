// Host build of the per-pixel arithmetic of csrc/augment.cu (csrc/augment_math.h), for tests/test_data_cpu.py:
// the same loop the kernel's threads run, on the CPU, so that the fixed-point formulas can be checked against the
// OpenCV-generated golden vectors without a GPU.  Test infrastructure; built by the test with
//   g++ -O2 -ffp-contract=off -shared -fPIC
#include "augment_math.h"

extern "C" void host_augment_batch(const uint8_t* images, const uint8_t* labels, const int* geom, const int64_t* lut,
                                   const float* norm, float* out_image, int64_t* out_label, int N, int H, int W,
                                   int ch, int cw) {
    for (int n = 0; n < N; ++n) {
        const int* gm = geom + n * 5;
        const int nh = gm[0], nw = gm[1], cy = gm[2], cx = gm[3], flip = gm[4];
        const double scale_y = tss_resize_scale(nh, H), scale_x = tss_resize_scale(nw, W);
        const uint8_t* img = images + (size_t)n * H * W * 3;
        for (int oy = 0; oy < ch; ++oy) {
            const int dy = oy + cy;
            const TssTap ty = tss_linear_tap(dy, scale_y, H, false);
            const uint8_t* r0 = img + (size_t)ty.i0 * W * 3;
            const uint8_t* r1 = img + (size_t)ty.i1 * W * 3;
            const int ly = tss_nearest_tap(dy, scale_y, H);
            for (int ox = 0; ox < cw; ++ox) {
                const int dx = (flip ? cw - 1 - ox : ox) + cx;
                const TssTap tx = tss_linear_tap(dx, scale_x, W, true);
                for (int c = 0; c < 3; ++c)
                    out_image[(((size_t)n * 3 + c) * ch + oy) * cw + ox] = tss_augment_value(r0, r1, tx, ty, c, norm[c], norm[3 + c]);
                if (labels != nullptr) {
                    const int id = labels[((size_t)n * H + ly) * W + tss_nearest_tap(dx, scale_x, W)];
                    out_label[((size_t)n * ch + oy) * cw + ox] = lut != nullptr ? lut[id] : id;
                }
            }
        }
    }
}
