// Detector hand-off on the device: the frame that hm_preprocess crops from is also what the
// YOLOv7 detector sees, so its letterbox and the box rescaling run here and the frame never returns to the host.
//
// Replaces, bit for bit:
//   LoadImage.process_img (yolo/yolov7/utils/datasets.py:137-141) = letterbox(img0, 640, stride 32, auto=True)
//   (datasets.py:999-1030: cv2.resize INTER_LINEAR to round(shape * r), cv2.copyMakeBorder 114), BGR -> RGB, HWC -> CHW,
//   and the uint8 -> float32 / float16 conversion and division by 255 of Detector.detect (yolo/detector.py:121-123);
//   scale_coords(...).round() + the 'right' / 'left' labelling (yolo/detector.py:139-145, utils/general.py:323-345).
//
// cv2.resize(INTER_LINEAR) for 8-bit images is fixed point (imgproc/resize.cpp, HResizeLinear / VResizeLinear):
//   per axis  f = (float)((d + 0.5) * (src / dst) - 0.5),  s = floor(f),  f -= s          (scale in double, f in float)
//   columns   s < 0 -> (s, f) = (0, 0);  s >= src - 1 -> (s, f) = (src - 1, 0)            (weights clamped with the tap)
//   rows      taps s, s + 1 clamped into [0, src), weights kept
//   weights   a0 = rint((1 - f) * 2048), a1 = rint(f * 2048)                               (short, 11 bits)
//   value     h(y) = S[y][s_x] a0x + S[y][s_x + 1] a1x                                     (int, per source row)
//             out = (((a0y (h(s_y) >> 4)) >> 16) + ((a1y (h(s_y + 1) >> 4)) >> 16) + 2) >> 2
// oracle/letterbox.py restates the same formulas in numpy and is pinned to cv2.resize itself (tests/test_oracle_letterbox.py).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"

namespace hm {

struct LetterboxGeom {
    int H, W;            // source frame
    int rh, rw;          // resized (un-padded) size: round(H r), round(W r)
    int top, left;       // border
    int oh, ow;          // output size = resized + borders
    int resize;          // 0: the frame already has the un-padded size (cv2.resize is skipped, datasets.py:1023)
    double scale_x, scale_y;   // W / rw, H / rh (cv2: inv_scale)
};

// one axis of cv2.resize's coefficient table
__device__ __forceinline__ void lb_axis(int d, double scale, int ssize, bool clamp_weights, int* s0, int* s1, int* a0, int* a1) {
    float f = (float)__dadd_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), -0.5);
    int s = (int)floorf(f);
    f = __fsub_rn(f, (float)s);
    if (clamp_weights) {
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
    }
    *a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.0f, f), 2048.0f));
    *a1 = __float2int_rn(__fmul_rn(f, 2048.0f));
    *s0 = min(max(s, 0), ssize - 1);
    *s1 = min(max(s + 1, 0), ssize - 1);
}

// grid (ceil(ow / 128), oh), block 128: thread = output pixel; OutT = float or __half ([3][oh][ow], RGB, value / 255)
template <typename OutT>
__global__ void __launch_bounds__(128)
letterbox_kernel(const uint8_t* __restrict__ frame, int64_t row_stride, const LetterboxGeom g, OutT* __restrict__ out) {
    ptx::pdl_prologue();
    const int x = blockIdx.x * 128 + threadIdx.x, y = blockIdx.y;
    if (x >= g.ow) return;
    int v[3] = {114, 114, 114};                                  // copyMakeBorder value, all channels
    const int dx = x - g.left, dy = y - g.top;
    if (dx >= 0 && dx < g.rw && dy >= 0 && dy < g.rh) {
        if (!g.resize) {
            const uint8_t* q = frame + (size_t)dy * row_stride + (size_t)dx * 3;
            v[0] = q[0]; v[1] = q[1]; v[2] = q[2];
        } else {
            int sx0, sx1, ax0, ax1, sy0, sy1, ay0, ay1;
            lb_axis(dx, g.scale_x, g.W, true, &sx0, &sx1, &ax0, &ax1);
            lb_axis(dy, g.scale_y, g.H, false, &sy0, &sy1, &ay0, &ay1);
            const uint8_t* r0 = frame + (size_t)sy0 * row_stride;
            const uint8_t* r1 = frame + (size_t)sy1 * row_stride;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int h0 = r0[sx0 * 3 + c] * ax0 + r0[sx1 * 3 + c] * ax1;
                const int h1 = r1[sx0 * 3 + c] * ax0 + r1[sx1 * 3 + c] * ax1;
                v[c] = (((ay0 * (h0 >> 4)) >> 16) + ((ay1 * (h1 >> 4)) >> 16) + 2) >> 2;
            }
        }
    }
    const size_t plane = (size_t)g.oh * g.ow, o = (size_t)y * g.ow + x;
#pragma unroll
    for (int c = 0; c < 3; ++c) {                                // c = BGR index; output plane = 2 - c (RGB)
        const float f = __fdiv_rn((float)v[c], 255.0f);          // img.float() / 255.0; fp16: the half division rounds once more
        if constexpr (sizeof(OutT) == 2) out[(size_t)(2 - c) * plane + o] = __float2half_rn(f);
        else out[(size_t)(2 - c) * plane + o] = f;
    }
}

// scale_coords + clip_coords + round (yolov7/utils/general.py:323-345, yolo/detector.py:139) in the tensor's own precision:
// float32 rows, or float16 rows with every torch op rounding back to half (CUDA half kernels compute in float and round).
// det: [n][stride] = x1 y1 x2 y2 conf cls; boxes: [n][4] double (what det.tolist() hands to prepare_batch_bbox),
// is_right[n] = (cls == 1) (detector.py:141-144).  One thread per detection.
template <typename T>
__global__ void scale_boxes_kernel(const T* __restrict__ det, int n, int stride, float pad_x, float pad_y, float gain,
                                   float max_x, float max_y, double* __restrict__ boxes, uint8_t* __restrict__ is_right) {
    ptx::pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const T* d = det + (size_t)i * stride;
    auto rnd = [](float v) -> float {
        if constexpr (sizeof(T) == 2) return __half2float(__float2half_rn(v)); else return v;
    };
    auto ld = [&](int k) -> float {
        if constexpr (sizeof(T) == 2) return __half2float(d[k]); else return d[k];
    };
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const bool is_x = (k & 1) == 0;
        float v = rnd(__fsub_rn(ld(k), is_x ? pad_x : pad_y));   // coords[:, [0, 2]] -= pad[0] ...
        v = rnd(__fdiv_rn(v, gain));                             // coords[:, :4] /= gain
        v = fminf(fmaxf(v, 0.f), is_x ? max_x : max_y);          // clamp_(0, img_shape)
        boxes[(size_t)i * 4 + k] = (double)rnd(rintf(v));        // .round(): half to even
    }
    is_right[i] = ld(5) == 1.0f ? 1 : 0;
}

}  // namespace hm
