// MANO transformer-decoder head in fp32 (precision-critical, ~0.1 GFLOP/crop after re-association).
//
// Replaces MANOTransformerDecoderHead.forward (heads/mano_head.py:61-115) and the
// TransformerDecoder / TransformerCrossAttn / Attention / CrossAttention / FeedForward stack
// (components/pose_transformer.py:55-124,160-201,301-357) plus rot6d_to_rotmat
// (utils/geometry.py:47-70) and pred_cam_t (hamer.py:131-134).
//
// The decoder has ONE query token per crop, so
//   * self-attention: softmax over a single logit is 1  =>  SA(y) = to_out(W_v y)
//   * cross-attention: (W_k c_t) . q_h = c_t . (W_k,h^T q_h) and sum_t a_t W_v c_t = W_v sum_t a_t c_t
//     => per layer 8 score vectors against the 192 context tokens and 8 pooled contexts replace the
//     503 MFLOP K/V projection of the context (exact in real arithmetic).
// Everything is fp32 FMA on CUDA cores; the dense pieces are "skinny" GEMMs with M = batch.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "elementwise.cuh"

namespace hm {

// ------------------------------------------------------------------------------------------------
// Y[b, n] (+)= act( scale * sum_k X[b, k] * W[n, k] + bias[n] ),  b < B, n < N, K % 64 == 0, K <= 1280.
// Batched over blockIdx.z with element strides (per-head products of the cross-attention).
// Block = 8 warps = 8 output columns x a tile of 16 rows (crops): the whole [16 x K] X tile is staged
// once in shared memory (cp.async), each warp keeps its W row in registers (K/32 floats per lane),
// so there is one global round trip and one barrier per block; 64 KB of smem lets 3 blocks share an SM.
// ------------------------------------------------------------------------------------------------
enum SkinnyEpi : int { SK_NONE = 0, SK_GELU = 1, SK_RESID = 2 };

struct SkinnyParams {
    const float* X; int ldx; long long x_bs;
    const float* W; long long w_bs;      // W row pitch = K
    const float* bias; int bias_bs;      // may be null
    float* Y; int ldy; long long y_bs;
    int B, N, K;
    float scale;                         // applied to the dot product before bias
};

constexpr int SK_ROWS = 16;     // crops per block tile
constexpr int SK_COLS = 8;      // output columns (warps) per block
constexpr int SK_KMAX = 1280;   // largest K (cross-attention value projection)

__device__ __forceinline__ void sk_cp_async16(float* smem_dst, const float* g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"
                 ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(g) : "memory");
}

// Reduce 16 per-lane partial sums for 16 rows at once: afterwards lanes i and i + 16 hold the full sum
// of row i (15 + 1 shuffles instead of 16 x 5).
__device__ __forceinline__ float warp_transpose_reduce16(float (&v)[16], int lane) {
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int j = 0; j < o; ++j) {
            const float keep = up ? v[j + o] : v[j];
            const float send = up ? v[j] : v[j + o];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16);
}

template <int EPI>
__global__ void __launch_bounds__(SK_COLS * 32)
skinny_linear_kernel(const SkinnyParams p) {
    ptx::pdl_prologue();
    extern __shared__ __align__(16) float sk_xs[];              // [SK_ROWS][K]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.x * SK_COLS + warp;
    const int b0 = blockIdx.y * SK_ROWS;
    const int z = blockIdx.z;
    const float* X = p.X + z * p.x_bs;
    const float* W = p.W + z * p.w_bs;
    float* Y = p.Y + z * p.y_bs;
    const int rows = min(SK_ROWS, p.B - b0);
    const bool active = n < p.N;
    const int K = p.K, K4 = K >> 2;

    // X tile -> smem (rows >= rows are zero-filled so the inner loop has no predicates)
    for (int e = threadIdx.x; e < SK_ROWS * K4; e += blockDim.x) {
        const int r = e / K4, c = e - r * K4;
        float* dst = sk_xs + (size_t)r * K + c * 4;
        if (r < rows) sk_cp_async16(dst, X + (size_t)(b0 + r) * p.ldx + c * 4);
        else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    // this warp's W row: lane holds float4 chunks lane, lane + 32, ... (K / 128 of them, <= 10)
    float4 w[SK_KMAX / 128];
#pragma unroll
    for (int i = 0; i < SK_KMAX / 128; ++i) {
        const int c = lane + 32 * i;
        w[i] = (active && c < K4) ? __ldg(reinterpret_cast<const float4*>(W + (size_t)n * K) + c)
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    float acc[SK_ROWS];
#pragma unroll
    for (int r = 0; r < SK_ROWS; ++r) acc[r] = 0.f;
#pragma unroll
    for (int i = 0; i < SK_KMAX / 128; ++i) {
        const int c = lane + 32 * i;
        if (i * 32 < K4) {                                      // block-uniform
            const int cc = c < K4 ? c : 0;                      // lanes past a 64-wide tail re-read chunk 0 (w = 0)
#pragma unroll
            for (int r = 0; r < SK_ROWS; ++r) {
                const float4 x = *reinterpret_cast<const float4*>(sk_xs + (size_t)r * K + cc * 4);
                acc[r] = fmaf(x.x, w[i].x, fmaf(x.y, w[i].y, fmaf(x.z, w[i].z, fmaf(x.w, w[i].w, acc[r]))));
            }
        }
    }
    float v = warp_transpose_reduce16(acc, lane);               // lanes i, i+16: row i
    if (!active) return;
    const int r = lane & 15;
    if (lane < 16 && r < rows) {
        const float bias = p.bias ? __ldg(p.bias + z * p.bias_bs + n) : 0.f;
        v = v * p.scale + bias;
        if constexpr (EPI == SK_GELU) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
        float* y = Y + (size_t)(b0 + r) * p.ldy + n;
        if constexpr (EPI == SK_RESID) v += *y;
        *y = v;
    }
}

// ------------------------------------------------------------------------------------------------
// Cross-attention scores: sc[b, h, t] = sum_c feats[b, t, c] * u[b, h, c]   (u already carries the
// 64^-0.5 scale).  grid (192/32, B), block 256: the 8 x 1280 u-vectors sit in smem, each warp owns
// 4 context tokens.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
head_scores_kernel(const float* __restrict__ feats, const float* __restrict__ u,
                   float* __restrict__ sc) {
    ptx::pdl_prologue();
    __shared__ __align__(16) float us[8][1280];
    const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float4* ub = reinterpret_cast<const float4*>(u + (size_t)b * 8 * 1280);
    for (int e = threadIdx.x; e < 8 * 320; e += 256) reinterpret_cast<float4*>(&us[0][0])[e] = ub[e];
    __syncthreads();
    // a warp owns 4 context tokens and walks them together: every u float4 read from shared memory
    // feeds 4 tokens (the kernel was bound by shared-memory reads with one token at a time)
    const int t0 = blockIdx.x * 32 + warp * 4;
    const float4* f = reinterpret_cast<const float4*>(feats + ((size_t)b * 192 + t0) * 1280);
    float acc[4][8];
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int h = 0; h < 8; ++h) acc[t][h] = 0.f;
#pragma unroll 2
    for (int i = 0; i < 10; ++i) {
        float4 x[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) x[t] = f[(size_t)t * 320 + i * 32 + lane];
#pragma unroll
        for (int h = 0; h < 8; ++h) {
            const float4 w = reinterpret_cast<const float4*>(&us[h][0])[i * 32 + lane];
#pragma unroll
            for (int t = 0; t < 4; ++t)
                acc[t][h] = fmaf(x[t].x, w.x, fmaf(x[t].y, w.y, fmaf(x[t].z, w.z, fmaf(x[t].w, w.w, acc[t][h]))));
        }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int h = 0; h < 8; ++h) {
            const float v = warp_sum(acc[t][h]);
            if (lane == 0) sc[((size_t)b * 8 + h) * 192 + t0 + t] = v;
        }
}

// ------------------------------------------------------------------------------------------------
// softmax over the 192 tokens + attention-weighted pooling of the context:
//   pooled[b, h, c] = sum_t softmax_t(sc[b, h, :])[t] * feats[b, t, c]
// grid (1280/128, B), block 128: thread = context channel c (coalesced), 8 heads in registers.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
head_pool_kernel(const float* __restrict__ feats, const float* __restrict__ sc,
                 float* __restrict__ pooled) {
    ptx::pdl_prologue();
    __shared__ __align__(16) float a[192][8];          // [token][head]: two float4 broadcasts per token
    const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int h = warp; h < 8; h += 4) {               // each warp normalises two heads
        const float* s = sc + ((size_t)b * 8 + h) * 192;
        float v[6], mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 6; ++i) { v[i] = s[i * 32 + lane]; mx = fmaxf(mx, v[i]); }
        mx = warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 6; ++i) { v[i] = expf(v[i] - mx); sum += v[i]; }
        const float inv = 1.0f / warp_sum(sum);
#pragma unroll
        for (int i = 0; i < 6; ++i) a[i * 32 + lane][h] = v[i] * inv;
    }
    __syncthreads();
    const int c = blockIdx.x * 128 + threadIdx.x;
    const float* f = feats + (size_t)b * 192 * 1280 + c;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 8
    for (int t = 0; t < 192; ++t) {
        const float x = f[(size_t)t * 1280];
        const float4 a0 = *reinterpret_cast<const float4*>(&a[t][0]);
        const float4 a1 = *reinterpret_cast<const float4*>(&a[t][4]);
        acc[0] = fmaf(a0.x, x, acc[0]); acc[1] = fmaf(a0.y, x, acc[1]);
        acc[2] = fmaf(a0.z, x, acc[2]); acc[3] = fmaf(a0.w, x, acc[3]);
        acc[4] = fmaf(a1.x, x, acc[4]); acc[5] = fmaf(a1.y, x, acc[5]);
        acc[6] = fmaf(a1.z, x, acc[6]); acc[7] = fmaf(a1.w, x, acc[7]);
    }
#pragma unroll
    for (int h = 0; h < 8; ++h) pooled[((size_t)b * 8 + h) * 1280 + c] = acc[h];
}

// Broadcast the constant layer-0 token (to_token_embedding.bias + pos_embedding) to every crop.
__global__ void head_init_token_kernel(const float* __restrict__ tok0, float* __restrict__ x, int B) {
    ptx::pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B * 1024) x[i] = tok0[i & 1023];
}

// ------------------------------------------------------------------------------------------------
// Read-out epilogue: 6D -> rotation matrices (Gram-Schmidt, geometry.py:47-70), camera
// translation (hamer.py:131-134), focal length; writes the reference's output tensors.
// raw[b] = [pose6d(96) | betas(10) | cam(3)] (decpose/decshape/deccam + init_*, mano_head.py:93-95)
// ------------------------------------------------------------------------------------------------
struct HeadOut {
    float* global_orient;   // [B,1,3,3]
    float* hand_pose;       // [B,15,3,3]
    float* betas;           // [B,10]
    float* pred_cam;        // [B,3]
    float* pred_cam_t;      // [B,3]
    float* focal_length;    // [B,2]
    float* rotmats;         // [B,16,3,3] packed copy for the MANO kernel
};

__global__ void head_finalize_kernel(const float* __restrict__ raw, HeadOut o, int B,
                                     float focal, float image_size) {
    ptx::pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * 16) return;
    const int b = i >> 4, j = i & 15;
    const float* r = raw + (size_t)b * 109;
    const float* p = r + j * 6;
    // x.reshape(-1,2,3).permute(0,2,1): a1 = p[0:3], a2 = p[3:6]
    const float a1x = p[0], a1y = p[1], a1z = p[2], a2x = p[3], a2y = p[4], a2z = p[5];
    const float n1 = fmaxf(sqrtf(a1x * a1x + a1y * a1y + a1z * a1z), 1e-12f);   // F.normalize eps
    const float b1x = a1x / n1, b1y = a1y / n1, b1z = a1z / n1;
    const float d = b1x * a2x + b1y * a2y + b1z * a2z;
    const float cx = a2x - d * b1x, cy = a2y - d * b1y, cz = a2z - d * b1z;
    const float n2 = fmaxf(sqrtf(cx * cx + cy * cy + cz * cz), 1e-12f);
    const float b2x = cx / n2, b2y = cy / n2, b2z = cz / n2;
    const float b3x = b1y * b2z - b1z * b2y, b3y = b1z * b2x - b1x * b2z, b3z = b1x * b2y - b1y * b2x;
    // columns (b1, b2, b3): R[row][col]
    const float R[9] = {b1x, b2x, b3x, b1y, b2y, b3y, b1z, b2z, b3z};
    float* dst = (j == 0) ? (o.global_orient + (size_t)b * 9) : (o.hand_pose + ((size_t)b * 15 + j - 1) * 9);
    float* pk = o.rotmats + ((size_t)b * 16 + j) * 9;
#pragma unroll
    for (int e = 0; e < 9; ++e) { dst[e] = R[e]; pk[e] = R[e]; }
    if (j == 0) {
#pragma unroll
        for (int e = 0; e < 10; ++e) o.betas[(size_t)b * 10 + e] = r[96 + e];
        const float c0 = r[106], c1 = r[107], c2 = r[108];
        o.pred_cam[b * 3 + 0] = c0; o.pred_cam[b * 3 + 1] = c1; o.pred_cam[b * 3 + 2] = c2;
        o.pred_cam_t[b * 3 + 0] = c1;
        o.pred_cam_t[b * 3 + 1] = c2;
        o.pred_cam_t[b * 3 + 2] = 2.0f * focal / (image_size * c0 + 1e-9f);
        o.focal_length[b * 2 + 0] = focal;
        o.focal_length[b * 2 + 1] = focal;
    }
}

}  // namespace hm
