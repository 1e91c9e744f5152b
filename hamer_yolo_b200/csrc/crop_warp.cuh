// Box -> 256x256 normalised crop, integer-exact with OpenCV.
//
// Replaces the per-box CPU loop of hamer_inference.prepare_batch_bbox (hamer/infer.py:154-259):
//   generate_image_patch_cv2 -> cv2.warpAffine(INTER_LINEAR, BORDER_CONSTANT 0)  (datasets/utils.py:359)
//   BGR->RGB (infer.py:228), horizontal flip for left hands (:229-230), HWC->CHW float32
//   (utils.py:379-392), (x - 255*mean) / (255*std) (infer.py:235-238).
// The affine matrix itself (cv2.getAffineTransform, utils.py:127) is computed on the host in
// double (hm_api.cu: affine_from_box) because its LU arithmetic must match OpenCV bit for bit.
//
// cv2.warpAffine for 8-bit images is fixed point: inverse map in double, coordinates in 1/1024 px
// (AB_BITS = 10) reduced to 1/32 px (INTER_BITS = 5), int16 bilinear weights scaled by 2^15,
// result = (sum + 2^14) >> 15.  With 1/32 fractions the weight products are exact multiples of 32,
// so the table OpenCV builds is w = (32-fx|fx) * (32-fy|fy) * 32 with no sum correction.
// Only 3 x 256 distinct normalised values exist, so normalisation is a 768-entry LUT built with
// the reference's float32 arithmetic.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"

namespace hm {

struct CropParams {
    double m[6];      // inverse affine (dst -> src), as cv::warpAffine derives it from M
    int32_t flip;     // 1 = left hand: mirror the patch horizontally (infer.py:229-230); 2 = sample the horizontally
                      // mirrored frame instead (generate_image_patch_cv2 do_flip, datasets/utils.py:347-349: the
                      // affine is built for c_x' = W - c_x - 1 and source column x is frame column W - 1 - x)
    int32_t pad;
};

__device__ __forceinline__ int cv_round_scaled(double v) {
    // saturate_cast<int>(v * AB_SCALE): cvRound = round-half-even; the scale is a power of two
    return (int)__double2ll_rn(__dmul_rn(v, 1024.0));
}

// grid (256 rows, n crops), block 256 (one output row per block; thread = output column)
__global__ void __launch_bounds__(256)
crop_warp_normalize_kernel(const uint8_t* __restrict__ frame, int H, int W, int64_t row_stride,
                           const CropParams* __restrict__ params, const float* __restrict__ lut,
                           float* __restrict__ out) {
    ptx::pdl_prologue();
    const int n = blockIdx.y, y = blockIdx.x, x = threadIdx.x;
    const CropParams p = params[n];
    const int adelta = cv_round_scaled(__dmul_rn(p.m[0], (double)x));
    const int bdelta = cv_round_scaled(__dmul_rn(p.m[3], (double)x));
    const int X0 = cv_round_scaled(__dadd_rn(__dmul_rn(p.m[1], (double)y), p.m[2])) + 16;
    const int Y0 = cv_round_scaled(__dadd_rn(__dmul_rn(p.m[4], (double)y), p.m[5])) + 16;
    const int X = (X0 + adelta) >> 5;
    const int Y = (Y0 + bdelta) >> 5;
    int sx = X >> 5, sy = Y >> 5;
    sx = max(-32768, min(32767, sx));
    sy = max(-32768, min(32767, sy));
    const int fx = X & 31, fy = Y & 31;
    const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32;
    const int w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;

    int acc[3] = {0, 0, 0};
    const bool x0in = (sx >= 0) && (sx < W), x1in = (sx + 1 >= 0) && (sx + 1 < W);
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
        const int yy = sy + dy;
        if (yy < 0 || yy >= H) continue;
        const uint8_t* rowp = frame + (size_t)yy * row_stride;
        const int wl = dy ? w10 : w00, wr = dy ? w11 : w01;
        if (x0in) {
            const uint8_t* q = rowp + (size_t)(p.flip == 2 ? W - 1 - sx : sx) * 3;
            acc[0] += wl * q[0]; acc[1] += wl * q[1]; acc[2] += wl * q[2];
        }
        if (x1in) {
            const uint8_t* q = rowp + (size_t)(p.flip == 2 ? W - 2 - sx : sx + 1) * 3;
            acc[0] += wr * q[0]; acc[1] += wr * q[1]; acc[2] += wr * q[2];
        }
    }
    const int xo = p.flip == 1 ? (255 - x) : x;
    float* o = out + (size_t)n * 3 * 65536 + (size_t)y * 256 + xo;
#pragma unroll
    for (int c = 0; c < 3; ++c) {                        // c = BGR index; output plane = 2 - c
        const int v = (acc[c] + (1 << 14)) >> 15;
        o[(size_t)(2 - c) * 65536] = __ldg(lut + (2 - c) * 256 + v);
    }
}

}  // namespace hm
