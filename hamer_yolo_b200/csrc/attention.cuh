// Multi-head self-attention core of one ViT block: softmax(q k^T / sqrt(80)) v over 192 tokens.
//
// Replaces Attention.forward lines vit.py:112-122 of the reference (reshape to (B,16,192,80),
// q*scale, q@k^T, softmax, @v, transpose back).  The qkv projection and the output projection are
// the tcgen05 GEMMs in gemm_tcgen05.cuh; this kernel is 2.4 % of the block's FLOPs.
//
// One CTA per (head, crop): Q, K, V (192 x 80 fp16 each) are staged in shared memory with
// cp.async, 6 warps each own 2 x 16 query rows (two CTAs per SM so one's loads overlap the other's MMAs).  S = Q K^T and O = P V run on the tensor cores
// (mma.sync m16n8k16 fp16 -> fp32); scores, max, exp and the row sum stay in fp32 registers;
// P is rounded to fp16 only as the MMA operand.  O is staged back through the warp's own Q rows so
// the global store is 16-byte coalesced.
//   qkv  : [B*192, 3840] fp16, column = which*1280 + head*80 + d   (vit.py:113 reshape order)
//   out  : [B*192, 1280] fp16, column = head*80 + d                 (vit.py:122)
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"

namespace hm {

constexpr int ATT_T = 192;
constexpr int ATT_D = 80;
constexpr int ATT_PITCH = 88;  // fp16 elements per smem row: 176 B keeps ldmatrix conflict-free
constexpr int ATT_THREADS = 192;   // 6 warps x 2 query tiles of 16 rows; 2 CTAs per SM overlap load and math
constexpr int ATT_SMEM_BYTES = 3 * ATT_T * ATT_PITCH * 2;

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_f16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                               uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 "
        "{%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(g) : "memory");
}
__device__ __forceinline__ uint32_t pack2_f16(float lo, float hi) {
    __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(ATT_THREADS, 2)
vit_attention_kernel(const __half* __restrict__ qkv, __half* __restrict__ out,
                     float scale_log2e) {
    ptx::pdl_prologue();
    extern __shared__ __align__(16) uint8_t att_smem[];
    __half* sQ = reinterpret_cast<__half*>(att_smem);
    const int h = blockIdx.x, b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- stage Q, K, V -------------------------------------------------------------------------
    const __half* base = qkv + (size_t)b * ATT_T * 3840 + h * ATT_D;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(att_smem);
    for (int e = threadIdx.x; e < 3 * ATT_T * 10; e += ATT_THREADS) {
        const int m = e / (ATT_T * 10), r = (e / 10) % ATT_T, ch = e % 10;
        cp_async16(sbase + (uint32_t)(((m * ATT_T + r) * ATT_PITCH + ch * 8) * 2),
                   base + (size_t)r * 3840 + m * 1280 + ch * 8);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    const uint32_t q_addr = sbase;
    const uint32_t k_addr = sbase + ATT_T * ATT_PITCH * 2;
    const uint32_t v_addr = k_addr + ATT_T * ATT_PITCH * 2;

#pragma unroll 1
  for (int mt = 0; mt < 2; ++mt) {
    const int m0 = (warp * 2 + mt) * 16;
    // ---- S = Q K^T  (16 x 192 per warp: 24 n-tiles of 8) ----------------------------------------
    float s[24][4];
#pragma unroll
    for (int i = 0; i < 24; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < ATT_D / 16; ++ks) {
        uint32_t a[4];
        ldmatrix_x4(a, q_addr + (uint32_t)(((m0 + (lane & 7) + 8 * ((lane >> 3) & 1)) * ATT_PITCH +
                                            ks * 16 + 8 * (lane >> 4)) * 2));
#pragma unroll
        for (int np = 0; np < 12; ++np) {
            uint32_t kb[4];
            ldmatrix_x4(kb, k_addr + (uint32_t)(((np * 16 + (lane & 7) + 8 * (lane >> 4)) * ATT_PITCH +
                                                 ks * 16 + 8 * ((lane >> 3) & 1)) * 2));
            mma_f16_16816(s[2 * np], a, kb[0], kb[1]);
            mma_f16_16816(s[2 * np + 1], a, kb[2], kb[3]);
        }
    }

    // ---- softmax over the 192 keys (rows r = lane/4 and r + 8) ----------------------------------
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int i = 0; i < 24; ++i) {
        mx0 = fmaxf(mx0, fmaxf(s[i][0], s[i][1]));
        mx1 = fmaxf(mx1, fmaxf(s[i][2], s[i][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int i = 0; i < 24; ++i) {
        s[i][0] = exp2f((s[i][0] - mx0) * scale_log2e);
        s[i][1] = exp2f((s[i][1] - mx0) * scale_log2e);
        s[i][2] = exp2f((s[i][2] - mx1) * scale_log2e);
        s[i][3] = exp2f((s[i][3] - mx1) * scale_log2e);
        sum0 += s[i][0] + s[i][1];
        sum1 += s[i][2] + s[i][3];
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);

    // ---- O = P V  (16 x 80 per warp: 10 n-tiles) ------------------------------------------------
    float o[10][4];
#pragma unroll
    for (int i = 0; i < 10; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < ATT_T / 16; ++kk) {
        uint32_t a[4];
        a[0] = pack2_f16(s[2 * kk][0], s[2 * kk][1]);
        a[1] = pack2_f16(s[2 * kk][2], s[2 * kk][3]);
        a[2] = pack2_f16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        a[3] = pack2_f16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
        for (int np = 0; np < 5; ++np) {
            uint32_t vb[4];
            ldmatrix_x4_trans(vb, v_addr + (uint32_t)(((kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * ATT_PITCH +
                                                       np * 16 + 8 * (lane >> 4)) * 2));
            mma_f16_16816(o[2 * np], a, vb[0], vb[1]);
            mma_f16_16816(o[2 * np + 1], a, vb[2], vb[3]);
        }
    }

    // ---- normalise, stage through this warp's own Q rows, coalesced store ------------------------
    const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
    __syncwarp();
    {
        const int r = lane >> 2, c = (lane & 3) * 2;
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            *reinterpret_cast<uint32_t*>(sQ + (m0 + r) * ATT_PITCH + i * 8 + c) =
                pack2_f16(o[i][0] * inv0, o[i][1] * inv0);
            *reinterpret_cast<uint32_t*>(sQ + (m0 + r + 8) * ATT_PITCH + i * 8 + c) =
                pack2_f16(o[i][2] * inv1, o[i][3] * inv1);
        }
    }
    __syncwarp();
    __half* obase = out + ((size_t)b * ATT_T + m0) * 1280 + h * ATT_D;
    for (int e = lane; e < 160; e += 32) {
        const int r = e / 10, ch = e % 10;
        const uint4 v = *reinterpret_cast<const uint4*>(sQ + (m0 + r) * ATT_PITCH + ch * 8);
        *reinterpret_cast<uint4*>(obase + (size_t)r * 1280 + ch * 8) = v;
    }
  }  // mt
}

}  // namespace hm
