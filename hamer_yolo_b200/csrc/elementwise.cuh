// HBM-bound row kernels of the backbone: LayerNorm (fp32 stats) and the patch-embed im2col.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"

namespace hm {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over the last dimension, one warp per row, two-pass statistics held in registers.
// Reference: nn.LayerNorm(1280, eps=1e-6) in every Block (vit.py:141,145,222) and last_norm
// (vit.py:335); nn.LayerNorm(1024, eps=1e-5) in the decoder's PreNorm (pose_transformer.py:30).
// D = 32 * 4 * V elements.  OutT = __nv_bfloat16 (feeds a GEMM) or float (last_norm / decoder).
// ------------------------------------------------------------------------------------------------
// EARLY_GB: gamma / beta are requested together with the row (80 more registers) instead of after the statistics - with a
// few hundred rows (video frames) the kernel is one latency chain and that second trip to L2 is a fifth of it; with
// thousands of rows occupancy matters more and the late loads hit L1.
template <int V, typename OutT, bool EARLY_GB = false>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                 const float* __restrict__ beta, OutT* __restrict__ y, int rows, float eps) {
    ptx::pdl_prologue();
    constexpr int D = 128 * V;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * D);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
    float4 v[V];
    float4 ge[EARLY_GB ? V : 1], be[EARLY_GB ? V : 1];
    if constexpr (EARLY_GB) {
#pragma unroll
        for (int i = 0; i < V; ++i) {       // volatile + memory clobber: ptxas otherwise sinks these loads back to their use
            asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[i].x), "=f"(v[i].y), "=f"(v[i].z), "=f"(v[i].w) : "l"(xr + i * 32 + lane) : "memory");
        }
#pragma unroll
        for (int i = 0; i < V; ++i) {
            asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(ge[i].x), "=f"(ge[i].y), "=f"(ge[i].z), "=f"(ge[i].w) : "l"(g4 + i * 32 + lane) : "memory");
            asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(be[i].x), "=f"(be[i].y), "=f"(be[i].z), "=f"(be[i].w) : "l"(b4 + i * 32 + lane) : "memory");
        }
    } else {
#pragma unroll
        for (int i = 0; i < V; ++i) v[i] = xr[i * 32 + lane];
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    if constexpr (EARLY_GB) {               // a data dependence ptxas cannot schedule around: the loads above stay above
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i) t += ge[i].x + be[i].x;
        s = fmaf(t, 0.0f, s);               // (adds zero: only t's availability before the reductions matters)
    }
    const float mean = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
        q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
#pragma unroll
    for (int i = 0; i < V; ++i) {
        float4 g, b;
        if constexpr (EARLY_GB) { g = ge[i]; b = be[i]; }
        else { g = __ldg(g4 + i * 32 + lane); b = __ldg(b4 + i * 32 + lane); }
        float4 o;
        o.x = v[i].x * rstd * g.x + b.x;
        o.y = v[i].y * rstd * g.y + b.y;
        o.z = v[i].z * rstd * g.z + b.z;
        o.w = v[i].w * rstd * g.w + b.w;
        if constexpr (sizeof(OutT) == 2) {
            uint2 p;
            p.x = ptx::pack_half2_sat(o.x, o.y);
            p.y = ptx::pack_half2_sat(o.z, o.w);
            reinterpret_cast<uint2*>(y + (size_t)row * D)[i * 32 + lane] = p;
        } else {
            reinterpret_cast<float4*>(y + (size_t)row * D)[i * 32 + lane] = o;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Patch-embed im2col.  Reference: backbone(x[:, :, :, 32:-32]) (hamer.py:119) followed by
// Conv2d(3,1280,k=16,s=16,p=2) (vit.py:168-175): token t = 12*i + j covers slice rows
// 16i-2 .. 16i+13 and slice columns 16j-2 .. 16j+13 (zero outside the 256x192 slice).
// Output row (b*192 + t) holds K = 2*768 fp16: [hi(768) | lo(768)] with hi = fp16(x),
// lo = fp16(x - hi); the GEMM weight is [W | W], so the pixel operand carries ~22 mantissa bits
// (the conv stays effectively fp32 although it runs on the 16-bit tensor pipe).
// k index inside each half = c*256 + ky*16 + kx (the Conv2d weight's own flattening).
// One block per (token row i, crop): 12 tokens x 768 values.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
im2col_patch16_kernel(const float* __restrict__ img, __half* __restrict__ out, int B) {
    ptx::pdl_prologue();
    const int b = blockIdx.y, i = blockIdx.x;           // i: 0..15 patch row
    const float* src = img + (size_t)b * 3 * 256 * 256;
    // 3 channels x 16 rows x 192 slice columns = 9216 source pixels; thread handles pairs along x
    for (int e = threadIdx.x; e < 3 * 16 * 96; e += blockDim.x) {
        const int c = e / (16 * 96);
        const int ky = (e / 96) % 16;
        const int xs = (e % 96) * 2;                    // slice column of the pair (even)
        const int y = 16 * i - 2 + ky;
        float2 px = make_float2(0.f, 0.f);
        if (y >= 0) px = *reinterpret_cast<const float2*>(src + ((size_t)c * 256 + y) * 256 + 32 + xs);
        // slice column xs -> patch column j = (xs+2)/16, kx = (xs+2)%16 (xs+2 even, pair stays in
        // the same patch since kx is even and <= 14)
        const int j = (xs + 2) >> 4, kx = (xs + 2) & 15;
        if (j < 12) {
            const size_t row = (size_t)b * 192 + i * 12 + j;
            const int k = c * 256 + ky * 16 + kx;
            const __half2 hi = __floats2half2_rn(px.x, px.y);
            const float2 hf = __half22float2(hi);
            const __half2 lo = __floats2half2_rn(px.x - hf.x, px.y - hf.y);
            *reinterpret_cast<__half2*>(out + row * 1536 + k) = hi;
            *reinterpret_cast<__half2*>(out + row * 1536 + 768 + k) = lo;
        }
    }
    // left zero padding: kx = 0,1 of patch column j = 0 (slice columns -2,-1)
    for (int e = threadIdx.x; e < 3 * 16; e += blockDim.x) {
        const int c = e / 16, ky = e % 16;
        const size_t row = (size_t)b * 192 + i * 12;
        const int k = c * 256 + ky * 16;
        const __half2 z = __floats2half2_rn(0.f, 0.f);
        *reinterpret_cast<__half2*>(out + row * 1536 + k) = z;
        *reinterpret_cast<__half2*>(out + row * 1536 + 768 + k) = z;
    }
}

}  // namespace hm
