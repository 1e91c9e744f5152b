// MANO transformer-decoder head in fp32 (precision-critical, ~0.1 GFLOP/crop after re-association).
//
// Replaces MANOTransformerDecoderHead.forward (heads/mano_head.py:61-115) and the
// TransformerDecoder / TransformerCrossAttn / Attention / CrossAttention / FeedForward stack
// (components/pose_transformer.py:55-124,160-201,301-357) plus rot6d_to_rotmat
// (utils/geometry.py:47-70) and pred_cam_t (hamer.py:131-134).
//
// The decoder has ONE query token per crop, so
//   * self-attention: softmax over a single logit is 1  =>  SA(y) = to_out(W_v y) = (W_o W_v) y + b_o (merged at load time)
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
// Y[b, n] = [R[b, n] +] act( scale * sum_k X'[b, k] * W[n, k] + bias[n] ),  b < B, n < N, K % 128 == 0, K <= 1280,
// X' = X or LayerNorm(X) (PreNorm of the decoder, pose_transformer.py:27-40, eps 1e-5).
// Batched over blockIdx.z with element strides (per-head products of the cross-attention).
// Block = 8 warps, tile = 16 rows (crops) x 8 CPW columns: the [16 x K] X tile is staged once in shared memory
// (cp.async) and - with LNF - normalised there, two rows per warp.  A warp owns CPW output columns; their W rows are
// copied by cp.async into lane-private shared-memory slots, all of them in flight from the first instruction on (the
// block then only waits on its own copy groups, no barrier), so every X float4 read from shared memory feeds CPW x 4
// FMAs: with CPW = 1 the kernel is bound by shared-memory reads (each warp re-reads the whole tile for one column);
// CPW = 4 makes it FMA-bound.  The host picks the widest CPW that still gives every SM a block and fits 227 KB.
// ------------------------------------------------------------------------------------------------
enum SkinnyEpi : int { SK_NONE = 0, SK_GELU = 1, SK_RESID = 2 };

struct SkinnyParams {
    const float* X; int ldx; long long x_bs;
    const float* W; long long w_bs;      // W row pitch = K
    const float* bias; int bias_bs;      // may be null
    float* Y; int ldy; long long y_bs;
    const float* R; int ldr;             // SK_RESID: Y = R + (...); R == Y updates in place, R != Y is needed when X is R
                                         // itself (PreNorm sub-layers: other blocks still read the rows this one updates)
    int B, N, K;
    float scale;                         // applied to the dot product before bias
    const float *ln_g, *ln_b;            // LNF: LayerNorm weight / bias over the K inputs
    float ln_eps;
};

constexpr int SK_ROWS = 16;     // crops per block tile
constexpr int SK_WARPS = 8;
constexpr int SK_KMAX = 1280;   // largest K (cross-attention value projection)
constexpr int SK_SMEM_MAX = 200 * 1024;
// X tile + the lane-private W slots (8 warps x CPW columns x K floats)
__host__ __device__ constexpr size_t sk_smem_bytes(int K, int cpw) { return (size_t)(SK_ROWS + SK_WARPS * cpw + 2) * K * sizeof(float); }   // + LayerNorm gamma / beta

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

__device__ __forceinline__ void sk_cp_async_wait(int pending) {   // cp.async.wait_group takes an immediate
    switch (pending) {
        case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
        case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
        case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
        default: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
    }
}

// RT = rows of the tile that are computed (4, 8 or 16): with a handful of crops (video frames) the other rows of the 16-row
// tile are padding, and every row costs a shared-memory read of the whole K per warp.
template <int CPW, int EPI, bool LNF, int RT = SK_ROWS>
__global__ void __launch_bounds__(SK_WARPS * 32)
skinny_linear_kernel(const SkinnyParams p) {
    static_assert(RT == 4 || RT == 8 || RT == 16, "row tile");
    ptx::griddep_launch_dependents();   // (the wait for the previous kernel comes after the weight copies are in flight)
    extern __shared__ __align__(16) float sk_smem[];            // X tile [SK_ROWS][K], then the W slots
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = (blockIdx.x * SK_WARPS + warp) * CPW;
    const int b0 = blockIdx.y * SK_ROWS;
    const int z = blockIdx.z;
    const float* X = p.X + z * p.x_bs;
    const float* W = p.W + z * p.w_bs;
    float* Y = p.Y + z * p.y_bs;
    const int rows = min(SK_ROWS, p.B - b0);
    const int K = p.K, K4 = K >> 2, KI = K >> 7, NG = (KI + 1) >> 1;   // W arrives in NG groups of two 128-wide K steps
    float* sk_xs = sk_smem;
    // this lane's private W slots: [step i][column c] float4 - written by its own cp.async, read back by itself, so the
    // only synchronisation the W stream needs is cp.async.wait_group
    float4* wslot = reinterpret_cast<float4*>(sk_smem + (size_t)SK_ROWS * K) + (size_t)warp * (KI * CPW * 32) + lane;

    // Everything the block will read goes in flight at once (X tile 16 K floats, W 8 CPW K floats: up to 192 KB): the
    // kernel is a latency chain otherwise (one block per SM, a handful of dependent round trips to L2 / HBM).  The
    // weights and the LayerNorm parameters are constants, so their copies are issued BEFORE the wait for the previous
    // kernel (programmatic dependent launch: this block is often resident while a smaller predecessor still runs); the
    // activations follow after it, as the last copy group.
    float* sk_ln = sk_smem + (size_t)(SK_ROWS + SK_WARPS * CPW) * K;   // LNF: [gamma (K) | beta (K)], behind the W slots
    if constexpr (LNF) {
        for (int e = threadIdx.x; e < 2 * K4; e += blockDim.x)
            sk_cp_async16(sk_ln + e * 4, (e < K4 ? p.ln_g : p.ln_b - K) + e * 4);
    }
    const float* wrow[CPW];                                     // columns past N re-read the last valid row (results dropped)
#pragma unroll
    for (int c = 0; c < CPW; ++c) wrow[c] = W + (size_t)min(n0 + c, p.N - 1) * K + lane * 4;
    for (int g = 0; g < NG; ++g) {
#pragma unroll
        for (int ii = 0; ii < 2; ++ii) {
            const int i = 2 * g + ii;
            if (i < KI) {
#pragma unroll
                for (int c = 0; c < CPW; ++c)
                    sk_cp_async16(reinterpret_cast<float*>(wslot + (i * CPW + c) * 32), wrow[c] + i * 128);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    ptx::griddep_wait();
    for (int e = threadIdx.x; e < SK_ROWS * K4; e += blockDim.x) {
        const int r = e / K4, c = e - r * K4;
        float* dst = sk_xs + (size_t)r * K + c * 4;
        if (r < rows) sk_cp_async16(dst, X + (size_t)(b0 + r) * p.ldx + c * 4);
        else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);   // no predicates in the inner loop
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    // bias and residual of the thread's output elements: issued now, consumed in the epilogue (lanes i and i + 16 own row i;
    // read at the end they were one more exposed L2 round trip per kernel)
    float bias_r[CPW], resid_r[CPW];
    {
        const int r = lane & 15;
#pragma unroll
        for (int c = 0; c < CPW; ++c) {
            const int n = n0 + c;
            const bool mine = lane < 16 && r < rows && n < p.N;
            bias_r[c] = (mine && p.bias) ? __ldg(p.bias + z * p.bias_bs + n) : 0.f;
            resid_r[c] = 0.f;
            if constexpr (EPI == SK_RESID) { if (mine) resid_r[c] = p.R[(size_t)(b0 + r) * p.ldr + n]; }
        }
    }
    sk_cp_async_wait(0);                                        // the X tile (the last group, so everything) has landed
    __syncthreads();
    if constexpr (LNF) {
        // PreNorm: rows 2 warp, 2 warp + 1 of the staged tile, two-pass statistics (same arithmetic as layernorm_kernel)
        const float4* g4 = reinterpret_cast<const float4*>(sk_ln);
        const float4* b4 = reinterpret_cast<const float4*>(sk_ln + K);
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            if (2 * warp + rr >= RT) continue;                  // padding rows of a small batch (warp-uniform)
            float4* xr = reinterpret_cast<float4*>(sk_xs + (size_t)(2 * warp + rr) * K);
            float4 v[SK_KMAX / 128];
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < SK_KMAX / 128; ++i)
                if (i < KI) { v[i] = xr[i * 32 + lane]; s += (v[i].x + v[i].y) + (v[i].z + v[i].w); }
            const float mean = warp_sum(s) / (float)K;
            float q = 0.f;
#pragma unroll
            for (int i = 0; i < SK_KMAX / 128; ++i)
                if (i < KI) {
                    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
                    q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
                }
            const float rstd = rsqrtf(warp_sum(q) / (float)K + p.ln_eps);
#pragma unroll
            for (int i = 0; i < SK_KMAX / 128; ++i)
                if (i < KI) {
                    const float4 g = g4[i * 32 + lane], b = b4[i * 32 + lane];
                    float4 o;
                    o.x = v[i].x * rstd * g.x + b.x; o.y = v[i].y * rstd * g.y + b.y;
                    o.z = v[i].z * rstd * g.z + b.z; o.w = v[i].w * rstd * g.w + b.w;
                    xr[i * 32 + lane] = o;
                }
        }
        __syncthreads();
    }
    float acc[CPW][SK_ROWS];                                    // rows >= RT stay zero
#pragma unroll
    for (int c = 0; c < CPW; ++c)
#pragma unroll
        for (int r = 0; r < SK_ROWS; ++r) acc[c][r] = 0.f;
    const float4* xs4 = reinterpret_cast<const float4*>(sk_xs) + lane;
    for (int g = 0; g < NG; ++g) {
        // (every copy group has completed by now: X was the last one)
#pragma unroll
        for (int ii = 0; ii < 2; ++ii) {
            const int i = 2 * g + ii;
            if (i < KI) {
                float4 wc[CPW];
#pragma unroll
                for (int c = 0; c < CPW; ++c) wc[c] = wslot[(i * CPW + c) * 32];
#pragma unroll
                for (int r = 0; r < RT; ++r) {
                    const float4 x = xs4[(size_t)r * K4 + i * 32];
#pragma unroll
                    for (int c = 0; c < CPW; ++c)
                        acc[c][r] = fmaf(x.x, wc[c].x, fmaf(x.y, wc[c].y, fmaf(x.z, wc[c].z, fmaf(x.w, wc[c].w, acc[c][r]))));
                }
            }
        }
    }
    const int r = lane & 15;
#pragma unroll
    for (int c = 0; c < CPW; ++c) {
        float v = warp_transpose_reduce16(acc[c], lane);        // lanes i, i + 16: row i
        const int n = n0 + c;
        if (lane < 16 && r < rows && n < p.N) {
            v = v * p.scale + bias_r[c];
            if constexpr (EPI == SK_GELU) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
            if constexpr (EPI == SK_RESID) v += resid_r[c];
            Y[(size_t)(b0 + r) * p.ldy + n] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Cross-attention, step 1: u[b, h, c] = scale * sum_d W_k[h 64 + d, c] * q[b, h 64 + d]  (the key projection moved onto
// the query, (W_k c_t) . q_h = c_t . (W_k,h^T q_h); W_k = to_kv.weight rows 0..511 in the checkpoint's own [512][1280] layout).
// grid (1280 / 128, 8 heads, ceil(B / 16)), block 128: thread = context channel c (coalesced W reads), the 16 x 64 query
// tile of the head sits in shared memory and is read as broadcasts; 64 independent weight loads per thread.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
head_u_kernel(const float* __restrict__ q, const float* __restrict__ wk, float* __restrict__ u, int B, float scale) {
    ptx::pdl_prologue();
    __shared__ __align__(16) float qs[SK_ROWS][64];
    const int h = blockIdx.y, b0 = blockIdx.z * SK_ROWS, rows = min(SK_ROWS, B - b0);
    const int c = blockIdx.x * 128 + threadIdx.x;
    const float* w = wk + (size_t)h * 64 * 1280 + c;
    // all 64 weight loads of the thread go out before the query tile is staged: the barrier below keeps the FMAs (which
    // need the tile) behind them, otherwise ptxas interleaves loads and FMAs eight at a time - eight exposed round trips
    float wv[64];
#pragma unroll
    for (int d = 0; d < 64; ++d) asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(wv[d]) : "l"(w + (size_t)d * 1280));
    for (int e = threadIdx.x; e < SK_ROWS * 16; e += 128) {
        const int r = e >> 4, c4 = e & 15;
        reinterpret_cast<float4*>(&qs[r][0])[c4] =
            r < rows ? *reinterpret_cast<const float4*>(q + (size_t)(b0 + r) * 512 + h * 64 + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    float acc[SK_ROWS];
#pragma unroll
    for (int r = 0; r < SK_ROWS; ++r) acc[r] = 0.f;
#pragma unroll
    for (int d4 = 0; d4 < 16; ++d4) {
#pragma unroll
        for (int r = 0; r < SK_ROWS; ++r) {
            const float4 x = reinterpret_cast<const float4*>(&qs[r][0])[d4];
            acc[r] = fmaf(x.x, wv[4 * d4], fmaf(x.y, wv[4 * d4 + 1], fmaf(x.z, wv[4 * d4 + 2], fmaf(x.w, wv[4 * d4 + 3], acc[r]))));
        }
    }
#pragma unroll
    for (int r = 0; r < SK_ROWS; ++r)
        if (r < rows) u[((size_t)(b0 + r) * 8 + h) * 1280 + c] = acc[r] * scale;
}

// ------------------------------------------------------------------------------------------------
// Cross-attention, step 2, one pass over the context ("flash decoding" with a single query token per head):
//   s[h, t] = feats[b, t, :] . u[b, h, :],  a = softmax_t(s),  pooled[b, h, :] = sum_t a[h, t] feats[b, t, :]
// Persistent: one CTA per SM walks a contiguous range of work items, item = 16 context tokens of one crop (80 KB of
// fp32 features = ONE contiguous run of the feature tensor).  The items stream through a two-deep shared-memory ring
// by cp.async.bulk (the next item lands while this one is computed: the features leave HBM once per layer, nothing is
// re-read through L2).  Per item:
//   scores : warp = (8-token group, fifth of the 1280 channels): 8 tokens x 8 heads per lane, the u vectors from shared
//            memory - every u float4 feeds 8 tokens, every x float4 8 heads (shared-memory reads balance the FMAs);
//   softmax: running (max, sum) per head over the CTA's consecutive items of the same crop (online rescaling);
//   pooling: thread j owns float4 column j of the 8 pooled vectors, accumulated in registers over the items.
// When the crop (or the CTA's range) ends, the segment's un-normalised [8][1280] + (m, l) go to global memory; the
// CTA whose segment completes the crop's 12 chunks (atomic ticket) merges the segments in chunk order.  The ranges
// are a static function of (B, grid), so the summation order is fixed run to run; PER_CHUNK (hm_config.deterministic)
// makes every item its own segment, which also makes it independent of the batch size.
// ------------------------------------------------------------------------------------------------
constexpr int XA_T = 16, XA_NCH = 192 / XA_T;
constexpr int XA_THREADS = 320;
constexpr uint32_t XA_CHUNK_BYTES = XA_T * 1280 * 4;
constexpr uint32_t XA_OFF_U = 2 * XA_CHUNK_BYTES;                 // [8][1280] floats
constexpr uint32_t XA_OFF_SC = XA_OFF_U + 8 * 1280 * 4;           // [5][16][8] partial scores
constexpr uint32_t XA_OFF_PS = XA_OFF_SC + 5 * 16 * 8 * 4;        // [16][8] exp(s - m)
constexpr uint32_t XA_OFF_WSM = XA_OFF_PS + 16 * 8 * 4;           // [12][8] merge weights
constexpr uint32_t XA_OFF_RUN = XA_OFF_WSM + XA_NCH * 8 * 4;      // m_run[8], l_run[8], rescale[8]
constexpr uint32_t XA_OFF_BAR = XA_OFF_RUN + 24 * 4;              // 2 mbarriers, then the is_last flag
constexpr uint32_t XA_SMEM_BYTES = XA_OFF_BAR + 16 + 16;

// first item of CTA c's range: items are split evenly and contiguously over the grid
__device__ __forceinline__ int xa_range_start(int c, int items, int grid) { return (int)(((long long)c * items) / grid); }

template <bool PER_CHUNK>
__global__ void __launch_bounds__(XA_THREADS, 1)
head_xattn_kernel(const float* __restrict__ feats, const float* __restrict__ u, float* __restrict__ part,
                  float* __restrict__ ml, unsigned int* __restrict__ ticket, float* __restrict__ pooled, int B) {
    extern __shared__ __align__(128) uint8_t xa_smem[];
    float* us = reinterpret_cast<float*>(xa_smem + XA_OFF_U);
    float* sc = reinterpret_cast<float*>(xa_smem + XA_OFF_SC);
    float* ps = reinterpret_cast<float*>(xa_smem + XA_OFF_PS);
    float* wsm = reinterpret_cast<float*>(xa_smem + XA_OFF_WSM);
    float* mrun = reinterpret_cast<float*>(xa_smem + XA_OFF_RUN);
    float* lrun = mrun + 8;
    float* resc = mrun + 16;
    uint64_t* full = reinterpret_cast<uint64_t*>(xa_smem + XA_OFF_BAR);
    int* is_last = reinterpret_cast<int*>(xa_smem + XA_OFF_BAR + 16);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    const int items = XA_NCH * B, G = gridDim.x;
    const int i0 = xa_range_start(blockIdx.x, items, G), i1 = xa_range_start(blockIdx.x + 1, items, G);
    ptx::griddep_launch_dependents();
    if (tid == 0) {
        ptx::mbar_init(&full[0], 1);
        ptx::mbar_init(&full[1], 1);
        ptx::fence_mbar_init();
    }
    __syncthreads();
    ptx::griddep_wait();
    auto issue = [&](int it, int buf) {          // one thread: the item's 80 KB as four bulk copies
        const float* src = feats + (size_t)it * (XA_T * 1280);      // item it = tokens [16 it, 16 it + 16) of the flat [B 192] rows
        ptx::mbar_arrive_expect_tx(&full[buf], XA_CHUNK_BYTES);
#pragma unroll
        for (int q = 0; q < 4; ++q)
            ptx::bulk_load_1d(xa_smem + buf * XA_CHUNK_BYTES + q * (XA_CHUNK_BYTES / 4), src + q * (XA_T / 4) * 1280,
                              XA_CHUNK_BYTES / 4, &full[buf]);
    };
    if (tid == 0 && i0 < i1) issue(i0, 0);
    int u_b = -1, seg_first = 0, seg_n = 0;
    float4 acc[8];
    for (int it = i0; it < i1; ++it) {
        const int k = it - i0, buf = k & 1, b = it / XA_NCH, ch = it - b * XA_NCH;
        if (tid == 0 && it + 1 < i1) issue(it + 1, buf ^ 1);       // the other slot was released by the barrier that ended item it - 1
        if (b != u_b) {                                           // block-uniform: a new crop starts
            const float4* ub = reinterpret_cast<const float4*>(u + (size_t)b * 8 * 1280);
            for (int e = tid; e < 8 * 320; e += XA_THREADS) reinterpret_cast<float4*>(us)[e] = __ldg(ub + e);
            u_b = b;
            seg_n = 0;
        }
        if (seg_n == 0) {                                         // a new segment: reset the running softmax state
            seg_first = ch;
#pragma unroll
            for (int h = 0; h < 8; ++h) acc[h] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (tid < 8) { mrun[tid] = -INFINITY; lrun[tid] = 0.f; }
        }
        __syncthreads();
        ptx::mbar_wait(&full[buf], (uint32_t)(k >> 1) & 1u);
        const float4* xc = reinterpret_cast<const float4*>(xa_smem + buf * XA_CHUNK_BYTES);
        {   // ---- scores: warp = (token group tg, channel fifth kf) ----
            const int tg = warp / 5, kf = warp - tg * 5;
            float a8[8][8];
#pragma unroll
            for (int t = 0; t < 8; ++t)
#pragma unroll
                for (int h = 0; h < 8; ++h) a8[t][h] = 0.f;
#pragma unroll
            for (int ii = 0; ii < 2; ++ii) {
                const int i = kf * 2 + ii;
                float4 w[8];
#pragma unroll
                for (int h = 0; h < 8; ++h) w[h] = reinterpret_cast<const float4*>(us)[h * 320 + i * 32 + lane];
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const float4 x = xc[(tg * 8 + t) * 320 + i * 32 + lane];
#pragma unroll
                    for (int h = 0; h < 8; ++h)
                        a8[t][h] = fmaf(x.x, w[h].x, fmaf(x.y, w[h].y, fmaf(x.z, w[h].z, fmaf(x.w, w[h].w, a8[t][h]))));
                }
            }
#pragma unroll
            for (int tp = 0; tp < 4; ++tp) {                      // two tokens x 8 heads = 16 values per reduction
                float v16[16];
#pragma unroll
                for (int h = 0; h < 8; ++h) { v16[h] = a8[2 * tp][h]; v16[8 + h] = a8[2 * tp + 1][h]; }
                const float r = warp_transpose_reduce16(v16, lane);   // lanes i, i + 16: value i
                if (lane < 16) sc[(kf * 16 + tg * 8 + 2 * tp + (lane >> 3)) * 8 + (lane & 7)] = r;
            }
        }
        __syncthreads();
        if (warp < 8) {     // ---- warp = head: this item's weights against the running maximum ----
            float v = -INFINITY;
            if (lane < XA_T) {
                v = 0.f;
#pragma unroll
                for (int kf = 0; kf < 5; ++kf) v += sc[(kf * 16 + lane) * 8 + warp];
            }
            const float m_old = mrun[warp];
            const float m_new = fmaxf(m_old, warp_max(v));
            const float e = lane < XA_T ? expf(v - m_new) : 0.f;
            const float l_item = warp_sum(e);
            if (lane < XA_T) ps[lane * 8 + warp] = e;
            if (lane == 0) {
                const float so = expf(m_old - m_new);             // exp(-inf) = 0 on the first item of a segment
                resc[warp] = so;
                lrun[warp] = fmaf(lrun[warp], so, l_item);
                mrun[warp] = m_new;
            }
        }
        __syncthreads();
        {   // ---- pooling: thread = float4 column ----
            const float4 r0 = *reinterpret_cast<const float4*>(resc), r1 = *reinterpret_cast<const float4*>(resc + 4);
            const float so[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
            for (int h = 0; h < 8; ++h) { acc[h].x *= so[h]; acc[h].y *= so[h]; acc[h].z *= so[h]; acc[h].w *= so[h]; }
#pragma unroll 4
            for (int t = 0; t < XA_T; ++t) {
                const float4 x = xc[t * 320 + tid];
                const float4 a0 = *reinterpret_cast<const float4*>(ps + t * 8);
                const float4 a1 = *reinterpret_cast<const float4*>(ps + t * 8 + 4);
                const float w[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                for (int h = 0; h < 8; ++h) {
                    acc[h].x = fmaf(w[h], x.x, acc[h].x); acc[h].y = fmaf(w[h], x.y, acc[h].y);
                    acc[h].z = fmaf(w[h], x.z, acc[h].z); acc[h].w = fmaf(w[h], x.w, acc[h].w);
                }
            }
        }
        ++seg_n;
        const bool seg_end = PER_CHUNK || it + 1 == i1 || (it + 1) / XA_NCH != b;      // block-uniform
        if (seg_end) {
            const size_t slot = (size_t)b * XA_NCH + seg_first;
            float4* dst = reinterpret_cast<float4*>(part + slot * 8 * 1280) + tid;
#pragma unroll
            for (int h = 0; h < 8; ++h) __stcg(dst + h * 320, acc[h]);
            if (tid < 8) { ml[(slot * 8 + tid) * 2] = mrun[tid]; ml[(slot * 8 + tid) * 2 + 1] = lrun[tid]; }
            __threadfence();
            __syncthreads();
            if (tid == 0) *is_last = (atomicAdd(ticket + b, (unsigned)seg_n) + (unsigned)seg_n == (unsigned)XA_NCH);
            __syncthreads();
            if (*is_last) {       // this segment completed the crop: merge the segments in chunk order
                __threadfence();
                if (tid < 8) {
                    // all 24 (m, l) values of the head's 12 chunk slots in flight at once: as two rolled loops of dependent
                    // L2 reads this was ~16 us of serial latency on 8 threads at the very end of the kernel
                    float mv[XA_NCH], lv[XA_NCH];
#pragma unroll
                    for (int s = 0; s < XA_NCH; ++s) {
                        const int idx = b * XA_NCH + s;
                        const int cta = (int)(((long long)idx * G + items - 1) / items);
                        const bool start = PER_CHUNK || s == 0 || (cta < G && xa_range_start(cta, items, G) == idx);
                        const float2 v = __ldcg(reinterpret_cast<const float2*>(ml + ((size_t)idx * 8 + tid) * 2));
                        mv[s] = start ? v.x : -INFINITY;
                        lv[s] = v.y;
                    }
                    float M = -INFINITY, L = 0.f;
#pragma unroll
                    for (int s = 0; s < XA_NCH; ++s) M = fmaxf(M, mv[s]);
                    float ev[XA_NCH];
#pragma unroll
                    for (int s = 0; s < XA_NCH; ++s) {
                        ev[s] = mv[s] == -INFINITY ? 0.f : expf(mv[s] - M);
                        if (ev[s] != 0.f) L = fmaf(ev[s], lv[s], L);
                    }
                    const float inv = 1.0f / L;
#pragma unroll
                    for (int s = 0; s < XA_NCH; ++s) wsm[s * 8 + tid] = mv[s] == -INFINITY ? -1.f : ev[s] * inv;   // -1: no segment starts here
                }
                if (tid == 0) ticket[b] = 0u;
                __syncthreads();
                // Two heads at a time with all 12 segments' loads in flight (24 independent L2 reads per thread): the merge is
                // the serial tail of the kernel when a crop's segments come from 12 different CTAs (few crops), and a loop over
                // the segments with 8 loads each was 12 dependent round trips.  Same summation order per head as before.
                const float4* src = reinterpret_cast<const float4*>(part + (size_t)b * XA_NCH * 8 * 1280) + tid;
                float4* dstp = reinterpret_cast<float4*>(pooled + (size_t)b * 8 * 1280) + tid;
#pragma unroll 1
                for (int h0 = 0; h0 < 8; h0 += 2) {
                    float4 v[XA_NCH][2];
#pragma unroll
                    for (int s = 0; s < XA_NCH; ++s) {
                        const bool on = wsm[s * 8] >= 0.f;                       // block-uniform: a segment starts at chunk s
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh)
                            v[s][hh] = on ? __ldcg(src + ((size_t)s * 8 + h0 + hh) * 320) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int s = 0; s < XA_NCH; ++s) {
                            const float w = wsm[s * 8 + h0 + hh];
                            if (w >= 0.f) {                                      // block-uniform (same segments for every head)
                                o.x = fmaf(w, v[s][hh].x, o.x); o.y = fmaf(w, v[s][hh].y, o.y);
                                o.z = fmaf(w, v[s][hh].z, o.z); o.w = fmaf(w, v[s][hh].w, o.w);
                            }
                        }
                        dstp[(h0 + hh) * 320] = o;
                    }
                }
            }
            seg_n = 0;
        }
        __syncthreads();          // every thread is done with slot buf (and with sc / ps / wsm)
    }
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
