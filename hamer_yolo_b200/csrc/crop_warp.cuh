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

// ------------------------------------------------------------------------------------------------
// Anti-aliased crop.  ViTDetDataset.__getitem__ (hamer/hamer/datasets/vitdet_dataset.py:62-71) and
// prepare_item (hamer/infer.py:312-316) blur the WHOLE frame before the warp when the crop window is more than 2.2 x the
// patch (bbox_size > 563 px):   cvimg = skimage.filters.gaussian(cvimg, sigma=(f - 1) / 2, channel_axis=2, preserve_range=True)
// = scipy.ndimage.gaussian_filter(float64 image, (sigma, sigma, 0), mode='nearest', truncate=4.0), and cv2.warpAffine then runs
// on the float64 image (float bilinear weights from the 1/32-pixel table, double accumulation) instead of the 8-bit fixed-point path.
// Only 4 x 65,536 blurred pixels per crop are ever read, so the blur is evaluated at those taps only, with scipy's own
// arithmetic (ni_filters.c NI_Correlate1D, symmetric weights: centre first, then (x[-k] + x[+k]) w_k for k = r .. 1; rows
// first, then columns; separately rounded double operations) - bit-identical to blurring the frame.  The weights come from
// the caller (numpy's exp, like scipy: hamer_yolo_b200/datasets.py) so that they match to the last bit as well.
// ------------------------------------------------------------------------------------------------
constexpr int CROP_BLUR_MAX_R = 32;     // radius int(4 sigma + 0.5): sigma <= 8 <-> crop windows up to 8,700 px

struct BlurParams {
    double w[CROP_BLUR_MAX_R + 1];      // w[k] = kernel weight at offset +-k
    int32_t radius;                     // 0: this crop takes the 8-bit path (the blur kernel skips it)
    int32_t pad;
};

// rows pass at (y, x): V = x[y] w0 + sum_{k = r..1} (x[y - k] + x[y + k]) w_k, rows clamped ('nearest'); 3 channels
__device__ __forceinline__ void blur_rows_at(const uint8_t* __restrict__ frame, int H, int64_t row_stride, int y, int x,
                                             const BlurParams& bp, double (&v)[3]) {
    const uint8_t* col = frame + (size_t)x * 3;
    const uint8_t* q = col + (size_t)y * row_stride;
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = __dmul_rn((double)q[c], bp.w[0]);
    for (int k = bp.radius; k >= 1; --k) {
        const uint8_t* qa = col + (size_t)max(y - k, 0) * row_stride;
        const uint8_t* qb = col + (size_t)min(y + k, H - 1) * row_stride;
        const double wk = bp.w[k];
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = __dadd_rn(v[c], __dmul_rn((double)((int)qa[c] + (int)qb[c]), wk));
    }
}

// blurred frame at (y, x): columns pass over the rows-pass values, columns clamped
__device__ __forceinline__ void blur_at(const uint8_t* __restrict__ frame, int H, int W, int64_t row_stride, int y, int x,
                                        const BlurParams& bp, double (&t)[3]) {
    double a[3], b[3];
    blur_rows_at(frame, H, row_stride, y, x, bp, a);
#pragma unroll
    for (int c = 0; c < 3; ++c) t[c] = __dmul_rn(a[c], bp.w[0]);
    for (int k = bp.radius; k >= 1; --k) {
        blur_rows_at(frame, H, row_stride, y, max(x - k, 0), bp, a);
        blur_rows_at(frame, H, row_stride, y, min(x + k, W - 1), bp, b);
        const double wk = bp.w[k];
#pragma unroll
        for (int c = 0; c < 3; ++c) t[c] = __dadd_rn(t[c], __dmul_rn(__dadd_rn(a[c], b[c]), wk));
    }
}

// grid (256 rows, n crops), block 256; crops with radius 0 return at once (they went through the 8-bit kernel).
// mean / sd: the three float32 normalisation constants (RGB order) of the LUT.
__global__ void __launch_bounds__(256)
crop_warp_blur_kernel(const uint8_t* __restrict__ frame, int H, int W, int64_t row_stride,
                      const CropParams* __restrict__ params, const BlurParams* __restrict__ blur, float3 mean, float3 sd,
                      float* __restrict__ out) {
    ptx::pdl_prologue();
    const int n = blockIdx.y, y = blockIdx.x, x = threadIdx.x;
    __shared__ BlurParams bp;
    if (blur[n].radius == 0) return;
    for (int i = threadIdx.x; i < (int)(sizeof(BlurParams) / 4); i += 256)
        reinterpret_cast<uint32_t*>(&bp)[i] = reinterpret_cast<const uint32_t*>(blur + n)[i];
    __syncthreads();
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
    // cv2's float interpolation table (initInterTab2D, INTER_LINEAR): products of the 1-D weights {1 - f / 32, f / 32} in float
    const float wx1 = (float)fx * 0.03125f, wx0 = 1.0f - wx1, wy1 = (float)fy * 0.03125f, wy0 = 1.0f - wy1;
    const float wt[4] = {wy0 * wx0, wy0 * wx1, wy1 * wx0, wy1 * wx1};
    double acc[3] = {0.0, 0.0, 0.0};
#pragma unroll 1
    for (int tp = 0; tp < 4; ++tp) {                       // remapBilinear: S00 w00 + S01 w01 + S10 w10 + S11 w11, left to right
        const int yy = sy + (tp >> 1), xx = sx + (tp & 1);
        double t[3] = {0.0, 0.0, 0.0};
        if (yy >= 0 && yy < H && xx >= 0 && xx < W)        // BORDER_CONSTANT 0 outside the frame
            blur_at(frame, H, W, row_stride, yy, p.flip == 2 ? W - 1 - xx : xx, bp, t);
        const double w = (double)wt[tp];
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] = tp == 0 ? __dmul_rn(t[c], w) : __dadd_rn(acc[c], __dmul_rn(t[c], w));
    }
    const int xo = p.flip == 1 ? (255 - x) : x;
    float* o = out + (size_t)n * 3 * 65536 + (size_t)y * 256 + xo;
    const float mn[3] = {mean.x, mean.y, mean.z}, sdv[3] = {sd.x, sd.y, sd.z};
#pragma unroll
    for (int c = 0; c < 3; ++c) {                          // c = BGR index; output plane = 2 - c; float32 normalisation
        const float v = (float)acc[c];                     // convert_cvimg_to_tensor: astype(np.float32)
        o[(size_t)(2 - c) * 65536] = __fdiv_rn(__fsub_rn(v, mn[2 - c]), sdv[2 - c]);
    }
}

}  // namespace hm
