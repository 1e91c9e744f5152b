// RootNet / SAR root-depth estimate: the call infer.py makes per hand right after the HaMeR forward
// (hamer/infer.py:1275 -> EstimateRGB.estimate_root_depth_custom, rootnet/Model_RGB.py:572-638).
//
//   crop        : process_bbox'ed square window -> 256 x 256 patch (cv2.warpAffine with the float32 matrix of
//                 generate_patch_image, rootnet/preprocessing.py:39-79), RGB, ToTensor + Normalize - the crop kernel of
//                 crop_warp.cuh with a second 768-entry table ((v / 255 - mean) / std in torchvision's float32 order)
//   backbone    : SARresnet34 (Model_RGB.py:179-197) = torchvision ResNet-34 stem + layer1..4.  Every convolution is
//                 im2col (fp16, NHWC, K = (ky, kx, c)) + the tcgen05 GEMM of gemm_tcgen05.cuh with BatchNorm folded into
//                 weight and bias at load time and ReLU / residual-add + ReLU as GEMM epilogues; activations stay NHWC fp16
//                 with a row pitch of max(C, 128) (64-channel layers pad their GEMM N to one 128-column tile).
//   head        : ResRootNet (Model_RGB.py:240-287): global average pool, 1 x 1 conv 512 -> 1, depth = gamma * k_value.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"

namespace hm {

constexpr int RN_CHUNK = 8;          // hands per pass (the reference runs one; the workspace is sized for this many)
constexpr int RN_NCONV = 36;

// conv1 (7 x 7, stride 2, pad 3, 3 -> 64) im2col straight from the crop kernel's fp32 planes [n][3][256][256]:
// row m = (b, oy, ox) of 128 x 128 outputs, k = (ky 7 + kx) 3 + c, 147 values zero-padded to 192.
// grid (128 oy, n), block 128 (thread = ox)
__global__ void __launch_bounds__(128)
rn_im2col_conv1_kernel(const float* __restrict__ img, __half* __restrict__ col) {
    ptx::pdl_prologue();
    const int b = blockIdx.y, oy = blockIdx.x, ox = threadIdx.x;
    const float* src = img + (size_t)b * 3 * 65536;
    __half* dst = col + ((size_t)(b * 128 + oy) * 128 + ox) * 192;
#pragma unroll
    for (int v = 0; v < 24; ++v) {                       // 24 vectors of 8 halves; every index below is a compile-time constant
        __half2 h2[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float f[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int k = v * 8 + e * 2 + q;
                f[q] = 0.f;
                if (k < 147) {
                    const int tap = k / 3, c = k % 3, ky = tap / 7, kx = tap % 7;
                    const int iy = oy * 2 - 3 + ky, ix = ox * 2 - 3 + kx;
                    if (iy >= 0 && iy < 256 && ix >= 0 && ix < 256) f[q] = src[(size_t)c * 65536 + iy * 256 + ix];
                }
            }
            h2[e] = __floats2half2_rn(f[0], f[1]);
        }
        reinterpret_cast<uint4*>(dst)[v] = *reinterpret_cast<const uint4*>(h2);
    }
}

// generic NHWC fp16 im2col: in [n][H][W][ld] (C valid channels, C % 8 == 0), k x k window, stride s, padding p ->
// col [n Ho Wo][k k C].  One warp per output pixel, lanes over the 16-byte vectors of the row.
__global__ void __launch_bounds__(256)
rn_im2col_kernel(const __half* __restrict__ in, __half* __restrict__ col, int n, int H, int W, int ld, int C, int k, int s, int p,
                 int Ho, int Wo) {
    ptx::pdl_prologue();
    const int m = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (m >= n * Ho * Wo) return;
    const int b = m / (Ho * Wo), r = m - b * Ho * Wo, oy = r / Wo, ox = r - oy * Wo;
    const int c8n = C >> 3, nv = k * k * c8n;
    uint4* dst = reinterpret_cast<uint4*>(col + (size_t)m * (k * k * C));
    for (int v = lane; v < nv; v += 32) {
        const int tap = v / c8n, c8 = v - tap * c8n, ky = tap / k, kx = tap - ky * k;
        const int iy = oy * s - p + ky, ix = ox * s - p + kx;
        uint4 q = make_uint4(0u, 0u, 0u, 0u);
        if (iy >= 0 && iy < H && ix >= 0 && ix < W)
            q = *reinterpret_cast<const uint4*>(in + ((size_t)(b * H + iy) * W + ix) * ld + c8 * 8);
        dst[v] = q;
    }
}

// MaxPool2d(3, stride 2, padding 1) on NHWC fp16 (padding = -inf: the maximum over the valid taps): [n][H][W][ld] ->
// [n][H/2][W/2][ld], all ld channels (the padded ones are zeros and stay zeros).  Thread = (output pixel, 8 channels).
__global__ void __launch_bounds__(256)
rn_maxpool_kernel(const __half* __restrict__ in, __half* __restrict__ out, int n, int H, int W, int ld) {
    ptx::pdl_prologue();
    const int Ho = H / 2, Wo = W / 2, c8n = ld >> 3;
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= (size_t)n * Ho * Wo * c8n) return;
    const int c8 = (int)(i % c8n);
    const size_t m = i / c8n;
    const int b = (int)(m / (Ho * Wo)), r = (int)(m - (size_t)b * Ho * Wo), oy = r / Wo, ox = r - oy * Wo;
    __half2 best[4];
    bool any = false;
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy * 2 - 1 + ky;
        if (iy < 0 || iy >= H) continue;
        for (int kx = 0; kx < 3; ++kx) {
            const int ix = ox * 2 - 1 + kx;
            if (ix < 0 || ix >= W) continue;
            const uint4 q = *reinterpret_cast<const uint4*>(in + ((size_t)(b * H + iy) * W + ix) * ld + c8 * 8);
            const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
            for (int e = 0; e < 4; ++e) best[e] = any ? __hmax2(best[e], h[e]) : h[e];
            any = true;
        }
    }
    *reinterpret_cast<uint4*>(out + m * ld + c8 * 8) = *reinterpret_cast<const uint4*>(best);
}

// ResRootNet: gamma = depth_layer(mean over the 8 x 8 positions), depth = gamma * k_value.  feats [n][64][512] fp16.
// One block of 512 threads per hand (thread = channel), fp32 accumulation.
__global__ void __launch_bounds__(512)
rn_head_kernel(const __half* __restrict__ feats, const float* __restrict__ w, const float* __restrict__ bias,
               const float* __restrict__ k_value, float* __restrict__ depth) {
    ptx::pdl_prologue();
    __shared__ float part[16];
    const int b = blockIdx.x, c = threadIdx.x;
    const __half* f = feats + (size_t)b * 64 * 512 + c;
    float s = 0.f;
#pragma unroll 8
    for (int pix = 0; pix < 64; ++pix) s += __half2float(f[(size_t)pix * 512]);
    float v = (s * (1.0f / 64.0f)) * w[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((c & 31) == 0) part[c >> 5] = v;
    __syncthreads();
    if (c < 32) {
        float t = c < 16 ? part[c] : 0.f;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (c == 0) depth[b] = (t + bias[0]) * k_value[b];
    }
}

}  // namespace hm
