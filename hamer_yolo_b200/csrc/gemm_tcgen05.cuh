// fp16 x fp16 -> f32 GEMM for the ViT-H backbone:  C[M,N] = A[M,K] * W[N,K]^T  (+ fused epilogue)
// Operands are fp16 rather than bf16 because the 0.5 mm / 1e-3 parity bar needs the 11-bit
// mantissa (tests/tools/emulate_bf16.py: bf16 activations give 1.9e-3 on the rotation matrices,
// fp16 3e-4).  tcgen05.mma kind::f16 runs fp16 and bf16 at the same rate but needs A and B of the
// same type (a mixed fp16 x bf16 descriptor raises an illegal-instruction fault on B200 - measured),
// so the weights are fp16 as well; every product is exact in the fp32 accumulator.
//
// Replaces every nn.Linear / the patch-embed Conv2d of the reference backbone
// (hamer/hamer/models/backbones/vit.py:77-79 fc1/fc2, :104 qkv, :107 proj, :168 patch-embed conv),
// which the reference runs through cuBLAS/MKL via ATen.
//
// sm_100a design
//   * persistent, warp-specialised: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (one lane),
//     warps 2.. = epilogue (TMEM -> registers -> global): 8 warps, 16 for the GELU epilogue whose
//     erf evaluation is issue-bound.
//   * operands: TMA 128B-swizzled K-major tiles, BK = 64 (one swizzle atom per row), STAGES-deep
//     mbarrier ring.
//   * accumulators: fp32 in TMEM, two stages (2 x BN columns) so the epilogue of tile i overlaps the
//     MMAs of tile i+1.
//   * CTA_GROUP == 2: a CTA pair (cluster 2x1x1) computes a 256 x BN tile with UMMA M = 256; each
//     CTA loads its own 128 rows of A and half (BN/2 rows) of W, the leader CTA issues the MMAs,
//     tcgen05.commit multicasts the "slot free" / "accumulator full" signals to both CTAs.
//   * epilogues (template EPI): bias -> fp16 | bias + erf-GELU -> fp16 |
//     bias + residual add into the fp32 stream | bias + positional table -> fp32.
//     The residual epilogue never reads the stream: each warp writes its 32x32 (acc + bias) block
//     into a 128B-swizzled shared-memory tile and issues a TMA reduce-add (cp.reduce.async.bulk
//     .add.f32), so x += tile happens at L2 and the SM only produces write traffic (the plain
//     read-modify-write version was bound by the latency of re-reading x: 34 % tensor-pipe
//     utilisation at K = 1280).  The positional epilogue (patch embed, 0.15 % of the FLOPs)
//     transposes through the same tile for 128-byte coalesced stores.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "ptx.cuh"

namespace hm {

enum GemmEpilogue : int {
    EPI_BIAS_BF16 = 0,       // out_f16[m,n] = acc + bias[n]        (16-bit activation output)
    EPI_BIAS_GELU_BF16 = 1,  // out_f16[m,n] = gelu(acc + bias[n])
    EPI_BIAS_RESID_F32 = 2,  // out_f32[m,n] += acc + bias[n]
    EPI_BIAS_POS_F32 = 3,    // out_f32[m,n]  = acc + bias[n] + pos[(m % pos_rows), n]
    EPI_BIAS_RELU_F16 = 4,        // out_f16[m,n] = relu(acc + bias[n])                    (RootNet convolutions, BatchNorm folded)
    EPI_BIAS_RESID_RELU_F16 = 5,  // out_f16[m,n] = relu(acc + bias[n] + resid_f16[m,n])   (BasicBlock tail)
};

struct GemmParams {
    int M, N, K;
    const float* bias;  // [N]
    void* out;          // [M, ldo] fp16 or f32
    uint32_t b_format;  // operand format of A and B: 0 = fp16 (the hot path), 1 = bf16
    int ldo;            // elements
    const float* pos;   // [pos_rows, N] (EPI_BIAS_POS_F32)
    int pos_rows;
    const void* resid;  // [M, ldo] fp16 (EPI_BIAS_RESID_RELU_F16)
    int ksplit_workers; // worker count the split was planned for (set by gemm_launch)
    int ksplit;         // EPI_BIAS_RESID_F32 only: the tiles of the last, partial wave are split into `ksplit`
                        // K-slices (each slice reduce-adds its partial product; slice 0 carries the bias); <= 1 = off
    int a_kwrap;        // > 0: A has only a_kwrap k-blocks and is re-read for k-blocks >= a_kwrap, i.e. the product is
                        // A * [W_hi | W_lo]^T with W stored as [N, 2K] (hm_config.split_weights: fp32-exact weights as
                        // the sum of two fp16 matrices); 0 = plain GEMM
};

// Work items of the persistent loop: full tiles first, then the K-slices of the tail tiles.
struct GemmItem { int tile, kb0, kb1; };
__device__ __forceinline__ GemmItem gemm_item(int item, int full_tiles, int ksplit, int num_k) {
    GemmItem it;
    if (item < full_tiles) { it.tile = item; it.kb0 = 0; it.kb1 = num_k; return it; }
    const int r = item - full_tiles, t = r / ksplit, sl = r - t * ksplit;
    it.tile = full_tiles + t;
    it.kb0 = sl * num_k / ksplit;
    it.kb1 = (sl + 1) * num_k / ksplit;
    return it;
}

constexpr int GEMM_BM = 128;  // rows per CTA (UMMA M = 128 * CTA_GROUP)
constexpr int GEMM_BK = 64;   // 64 x 16-bit = 128 B = one swizzle atom row
constexpr int GEMM_UMMA_K = 16;
constexpr uint32_t GEMM_SMEM_LIMIT = 227 * 1024;

template <int CTA_GROUP, int BN, int EPI>
struct GemmCfg {
    static constexpr bool F32_OUT = (EPI == EPI_BIAS_RESID_F32 || EPI == EPI_BIAS_POS_F32);
    static constexpr int EPI_WARPS = (EPI == EPI_BIAS_GELU_BF16) ? 16 : 8;
    static constexpr int THREADS = 64 + 32 * EPI_WARPS;
    static constexpr int COL_PARTS = EPI_WARPS / 4;             // warps sharing one TMEM lane quadrant
    static constexpr int COLS_PER_WARP = BN / COL_PARTS;
    static constexpr int B_ROWS = BN / CTA_GROUP;               // rows of W each CTA loads per stage
    static constexpr uint32_t A_BYTES = GEMM_BM * GEMM_BK * 2;
    static constexpr uint32_t B_BYTES = B_ROWS * GEMM_BK * 2;
    static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
    // fp32 epilogues stage 32x32 blocks per warp (double-buffered for the TMA reduce-add path)
    static constexpr uint32_t STAGING_BYTES = F32_OUT ? EPI_WARPS * 2 * 32 * 32 * 4 : 0;
    static constexpr uint32_t FIXED_BYTES = 1024 /*align*/ + 256 /*barriers*/ + STAGING_BYTES;
    static constexpr int STAGES_FIT = (GEMM_SMEM_LIMIT - FIXED_BYTES) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_FIT >= 8 ? 8 : STAGES_FIT;
    static constexpr uint32_t TMEM_COLS = 2 * BN;               // two accumulator stages
    static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + FIXED_BYTES;
    static_assert(TMEM_COLS == 512 || TMEM_COLS == 256 || TMEM_COLS == 128, "TMEM cols");
    static_assert(STAGES >= 3, "pipeline too shallow");
    static_assert(COLS_PER_WARP % 32 == 0, "epilogue chunking");
};

// erf-GELU as nn.GELU() computes it, 0.5 x (1 + erf(x / sqrt 2)), with erf from Abramowitz-Stegun
// 7.1.26 (|error| <= 1.5e-7 absolute - three orders below the fp16 rounding of the result) instead
// of erff(): ~14 issue slots instead of ~25, which is what bounds the fc1 epilogue.
__device__ __forceinline__ float gelu_erf(float x) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    float t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));   // 1 ulp, arg in [1, inf)
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * z * z));  // exp(-z^2), arg <= 0
    const float q = poly * t * e;                                // 1 - erf(|x|/sqrt2)  in (0, 1]
    // 0.5 x (1 + sign(x) (1 - q)) = x - 0.5 x q  for x >= 0,  0.5 x q  for x < 0
    const float hxq = 0.5f * x * q;
    return x >= 0.f ? x - hxq : hxq;
}

// Two elements per instruction with the packed fp32 pipe of sm_100 (FFMA2 / FMUL2): the fc1 epilogue
// is issue-bound, so halving the FMA/MUL count of the polynomial is what matters (MUFU stays scalar).
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
    // z = |x| / sqrt 2;  q = 1 - erf(z) = poly(t) t exp(-z^2), t = 1 / (1 + p z);  gelu(x) = max(x, 0) - (z / sqrt 2) q
    // (x >= 0: x - x q / 2;  x < 0: x q / 2 = -|x| q / 2): no select, one packed FMA for the final combination
    const float2 z = __fmul2_rn(make_float2(fabsf(x.x), fabsf(x.y)), make_float2(0.70710678118654752440f, 0.70710678118654752440f));
    const float2 d = __ffma2_rn(make_float2(0.3275911f, 0.3275911f), z, make_float2(1.0f, 1.0f));
    float2 t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(d.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(d.y));
    float2 poly = __ffma2_rn(make_float2(1.061405429f, 1.061405429f), t, make_float2(-1.453152027f, -1.453152027f));
    poly = __ffma2_rn(poly, t, make_float2(1.421413741f, 1.421413741f));
    poly = __ffma2_rn(poly, t, make_float2(-0.284496736f, -0.284496736f));
    poly = __ffma2_rn(poly, t, make_float2(0.254829592f, 0.254829592f));
    const float2 a = __fmul2_rn(__fmul2_rn(z, z), make_float2(-1.4426950408889634f, -1.4426950408889634f));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(a.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(a.y));
    const float2 q = __fmul2_rn(__fmul2_rn(poly, t), e);
    const float2 nz = __fmul2_rn(z, make_float2(-0.70710678118654752440f, -0.70710678118654752440f));   // -|x| / 2
    return __ffma2_rn(nz, q, make_float2(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f)));
}

__device__ __forceinline__ uint32_t pack_act2(float lo, float hi) { return ptx::pack_half2_sat(lo, hi); }

template <int CTA_GROUP, int BN, int EPI>
__global__ void __launch_bounds__(GemmCfg<CTA_GROUP, BN, EPI>::THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a,
                    const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_out, const GemmParams p) {
    using Cfg = GemmCfg<CTA_GROUP, BN, EPI>;
    constexpr int STAGES = Cfg::STAGES;
    ptx::griddep_launch_dependents();   // the next kernel may start its launch / set-up (PDL)

    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment. The dynamic-smem base offset is identical in
    // both CTAs of a pair, so the aligned offsets match (required: the MMA uses one descriptor
    // for both CTAs' shared memories).
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
    float* staging = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);   // 1024-aligned
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::STAGING_BYTES);
    uint64_t* full_bar = bars;                 // [STAGES]  TMA bytes landed (leader CTA's is used)
    uint64_t* empty_bar = bars + STAGES;       // [STAGES]  MMAs that read the slot retired
    uint64_t* tmem_full = bars + 2 * STAGES;   // [2]       accumulator stage complete
    uint64_t* tmem_empty = tmem_full + 2;      // [2]       accumulator stage drained (leader's)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = (CTA_GROUP == 2) ? ptx::cluster_ctarank() : 0u;
    const bool is_leader = (cta_rank == 0);

    const int num_k = p.K / GEMM_BK;
    const int n_blocks = p.N / BN;
    const int tile_m = GEMM_BM * CTA_GROUP;
    const int m_blocks = (p.M + tile_m - 1) / tile_m;
    const int num_tiles = m_blocks * n_blocks;
    const int worker = blockIdx.x / CTA_GROUP;
    const int num_workers = gridDim.x / CTA_GROUP;
    // wave quantisation: N = 1280 gives 240 tiles for 74 CTA pairs (3.24 waves); the 18 tiles of the last wave
    // are split along K so that every worker gets a quarter tile instead of 18 workers a whole one.  With fewer
    // tiles than workers (small batches) every tile is split.  `ksplit_workers` = the worker count the host
    // planned the split for (the grid may be larger than the tile count).
    const bool split = (EPI == EPI_BIAS_RESID_F32) && p.ksplit > 1;
    const int ksplit = split ? p.ksplit : 1;
    const int full_tiles = split ? (num_tiles / p.ksplit_workers) * p.ksplit_workers : num_tiles;
    const int num_items = full_tiles + (num_tiles - full_tiles) * ksplit;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_a);
        ptx::prefetch_tmap(&tmap_b);
        if constexpr (EPI == EPI_BIAS_RESID_F32) ptx::prefetch_tmap(&tmap_out);
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&tmem_full[s], 1);
            ptx::mbar_init(&tmem_empty[s], Cfg::EPI_WARPS * CTA_GROUP);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc<CTA_GROUP>(tmem_slot, Cfg::TMEM_COLS);
    }
    ptx::tc_fence_before();
    if constexpr (CTA_GROUP == 2) ptx::cluster_sync_all(); else __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // The W operand is constant: its tiles for the first k-blocks of this CTA's first item are requested BEFORE waiting for the
    // previous kernel (under programmatic dependent launch - small batches - this CTA is often resident while the previous,
    // smaller kernel is still running, and the weight stream of a few-crop GEMM is most of its traffic).  The A tiles of
    // the same stages follow after the wait; both complete on the stage's full barrier.
    int pre_stages = 0;
    if (warp == 0 && lane == 0 && worker < num_items) {
        const GemmItem wi = gemm_item(worker, full_tiles, ksplit, num_k);
        const int n_blk = wi.tile % n_blocks;
        const int n0 = n_blk * BN + (int)cta_rank * Cfg::B_ROWS;
        pre_stages = wi.kb1 - wi.kb0 < STAGES ? wi.kb1 - wi.kb0 : STAGES;
        for (int i = 0; i < pre_stages; ++i) {           // every slot is free at kernel start
            if constexpr (CTA_GROUP == 1) {
                ptx::mbar_arrive_expect_tx(&full_bar[i], Cfg::STAGE_BYTES);
                ptx::tma_load_2d(smem_b + i * Cfg::B_BYTES, &tmap_b, &full_bar[i], (wi.kb0 + i) * GEMM_BK, n0);
            } else {
                if (is_leader) ptx::mbar_arrive_expect_tx(&full_bar[i], 2 * Cfg::STAGE_BYTES);
                ptx::tma_load_2d_2sm(smem_b + i * Cfg::B_BYTES, &tmap_b, &full_bar[i], (wi.kb0 + i) * GEMM_BK, n0);
            }
        }
    }
    ptx::griddep_wait();                // barriers / TMEM / tensor maps are set up; now wait for the previous kernel's data

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = worker; item < num_items; item += num_workers) {
                const GemmItem wi = gemm_item(item, full_tiles, ksplit, num_k);
                const int m_blk = wi.tile / n_blocks, n_blk = wi.tile % n_blocks;
                const int m0 = m_blk * tile_m + (int)cta_rank * GEMM_BM;
                const int n0 = n_blk * BN + (int)cta_rank * Cfg::B_ROWS;
                for (int kb = wi.kb0; kb < wi.kb1; ++kb) {
                    const int ka = (p.a_kwrap > 0 && kb >= p.a_kwrap) ? kb - p.a_kwrap : kb;
                    if (item == worker && kb - wi.kb0 < pre_stages) {
                        // the stage's barrier is armed and its W tile already on the way: only the A tile is missing
                        if constexpr (CTA_GROUP == 1)
                            ptx::tma_load_2d(smem_a + stage * Cfg::A_BYTES, &tmap_a, &full_bar[stage], ka * GEMM_BK, m0);
                        else
                            ptx::tma_load_2d_2sm(smem_a + stage * Cfg::A_BYTES, &tmap_a, &full_bar[stage], ka * GEMM_BK, m0);
                        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                        continue;
                    }
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
                    if constexpr (CTA_GROUP == 1) {
                        ptx::mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                        ptx::tma_load_2d(smem_a + stage * Cfg::A_BYTES, &tmap_a, &full_bar[stage],
                                         ka * GEMM_BK, m0);
                        ptx::tma_load_2d(smem_b + stage * Cfg::B_BYTES, &tmap_b, &full_bar[stage],
                                         kb * GEMM_BK, n0);
                    } else {
                        if (is_leader)
                            ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
                        ptx::tma_load_2d_2sm(smem_a + stage * Cfg::A_BYTES, &tmap_a,
                                             &full_bar[stage], ka * GEMM_BK, m0);
                        ptx::tma_load_2d_2sm(smem_b + stage * Cfg::B_BYTES, &tmap_b,
                                             &full_bar[stage], kb * GEMM_BK, n0);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA, one lane) =====================
        if (is_leader && lane == 0) {
            const uint32_t idesc =
                ptx::make_idesc_f16kind_f32((int)p.b_format, (int)p.b_format, GEMM_BM * CTA_GROUP, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int item = worker; item < num_items; item += num_workers) {
                const GemmItem wi = gemm_item(item, full_tiles, ksplit, num_k);
                ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kb = wi.kb0; kb < wi.kb1; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint64_t a_desc =
                        ptx::make_smem_desc_k_sw128(ptx::smem_u32(smem_a + stage * Cfg::A_BYTES));
                    const uint64_t b_desc =
                        ptx::make_smem_desc_k_sw128(ptx::smem_u32(smem_b + stage * Cfg::B_BYTES));
#pragma unroll
                    for (int k = 0; k < GEMM_BK / GEMM_UMMA_K; ++k) {
                        // advance 16 elements (32 B) along K inside the 128 B swizzle atom
                        const uint64_t koff = (uint64_t)((k * GEMM_UMMA_K * 2) >> 4);
                        ptx::mma_f16<CTA_GROUP>(d_tmem, a_desc + koff, b_desc + koff, idesc,
                                                (uint32_t)(kb != wi.kb0 || k != 0));
                    }
                    if constexpr (CTA_GROUP == 1) ptx::mma_commit_1sm(&empty_bar[stage]);
                    else ptx::mma_commit_2sm(&empty_bar[stage], 0x3);
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                if constexpr (CTA_GROUP == 1) ptx::mma_commit_1sm(&tmem_full[acc]);
                else ptx::mma_commit_2sm(&tmem_full[acc], 0x3);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
            }
        }
    } else {
        // ===================== epilogue warps =====================
        const int ew = warp - 2;
        const int quad = warp & 3;              // TMEM lane quadrant this warp may access
        const int part = ew >> 2;               // which slice of the BN columns
        constexpr int COLS_PER_WARP = Cfg::COLS_PER_WARP;
        float* st = staging + ew * 2048;        // this warp's two 32x32 fp32 tiles (F32_OUT only)
        uint32_t chunk_no = 0;                  // running count of staged chunks (double buffering)
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int item = worker; item < num_items; item += num_workers) {
            const GemmItem wi = gemm_item(item, full_tiles, ksplit, num_k);
            const bool with_bias = wi.kb0 == 0;                       // K-slices after the first add only their product
            const int m_blk = wi.tile / n_blocks, n_blk = wi.tile % n_blocks;
            const int row0 = m_blk * tile_m + (int)cta_rank * GEMM_BM + quad * 32;
            const int row = row0 + lane;
            const int col0 = n_blk * BN + part * COLS_PER_WARP;
            // the bias of the warp's columns into L1 while the MMAs of the tile are still running: read behind
            // tcgen05.wait::ld it is an exposed trip to L2 per chunk (one tile per CTA at small batches: nothing hides it)
            if (lane < COLS_PER_WARP / 32) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.bias + col0 + lane * 32));
            ptx::mbar_wait(&tmem_full[acc], acc_phase);
            ptx::tc_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) +
                                    (uint32_t)(acc * BN + part * COLS_PER_WARP);
#pragma unroll 1
            for (int c = 0; c < COLS_PER_WARP / 32; ++c) {
                uint32_t r[32];
                ptx::tmem_ld_32x32(t_addr + (uint32_t)(c * 32), r);
                ptx::tmem_ld_wait();
                const int col = col0 + c * 32;
                if constexpr (!Cfg::F32_OUT) {
                    if (row < p.M) {
                        const float4* bias4 = reinterpret_cast<const float4*>(p.bias + col);
                        __half* out = reinterpret_cast<__half*>(p.out) + (size_t)row * p.ldo + col;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float2 v[4];                        // packed adds: the accumulator registers are adjacent pairs
#pragma unroll
                            for (int q = 0; q < 2; ++q) {
                                const float4 b = __ldg(bias4 + j * 2 + q);
                                v[q * 2 + 0] = __fadd2_rn(make_float2(__uint_as_float(r[j * 8 + q * 4 + 0]), __uint_as_float(r[j * 8 + q * 4 + 1])),
                                                          make_float2(b.x, b.y));
                                v[q * 2 + 1] = __fadd2_rn(make_float2(__uint_as_float(r[j * 8 + q * 4 + 2]), __uint_as_float(r[j * 8 + q * 4 + 3])),
                                                          make_float2(b.z, b.w));
                            }
                            if constexpr (EPI == EPI_BIAS_GELU_BF16) {
#pragma unroll
                                for (int e = 0; e < 4; ++e) v[e] = gelu_erf2(v[e]);
                            }
                            if constexpr (EPI == EPI_BIAS_RESID_RELU_F16) {
                                const uint4 rq = *reinterpret_cast<const uint4*>(
                                    reinterpret_cast<const __half*>(p.resid) + (size_t)row * p.ldo + col + j * 8);
                                const uint32_t rw[4] = {rq.x, rq.y, rq.z, rq.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float2 rf = __half22float2(*reinterpret_cast<const __half2*>(&rw[e]));
                                    v[e] = __fadd2_rn(v[e], rf);
                                }
                            }
                            if constexpr (EPI == EPI_BIAS_RELU_F16 || EPI == EPI_BIAS_RESID_RELU_F16) {
#pragma unroll
                                for (int e = 0; e < 4; ++e) v[e] = make_float2(fmaxf(v[e].x, 0.f), fmaxf(v[e].y, 0.f));
                            }
                            uint4 o;
                            o.x = pack_act2(v[0].x, v[0].y);
                            o.y = pack_act2(v[1].x, v[1].y);
                            o.z = pack_act2(v[2].x, v[2].y);
                            o.w = pack_act2(v[3].x, v[3].y);
                            *reinterpret_cast<uint4*>(out + j * 8) = o;
                        }
                    }
                } else if constexpr (EPI == EPI_BIAS_RESID_F32) {
                    // (acc + bias) -> 128B-swizzled smem tile -> TMA reduce-add into the stream
                    float* buf = st + (chunk_no & 1u) * 1024;
                    if (chunk_no >= 2) {                 // the TMA that read this buffer has drained
                        if (lane == 0) ptx::bulk_wait_read<1>();
                        __syncwarp();
                    }
                    const float4* bias4 = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 b = with_bias ? __ldg(bias4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                        const float2 lo = __fadd2_rn(make_float2(__uint_as_float(r[j * 4 + 0]), __uint_as_float(r[j * 4 + 1])), make_float2(b.x, b.y));
                        const float2 hi = __fadd2_rn(make_float2(__uint_as_float(r[j * 4 + 2]), __uint_as_float(r[j * 4 + 3])), make_float2(b.z, b.w));
                        const float4 v = make_float4(lo.x, lo.y, hi.x, hi.y);
                        *reinterpret_cast<float4*>(buf + lane * 32 + ((j ^ (lane & 7)) << 2)) = v;
                    }
                    ptx::fence_proxy_async();            // generic-proxy writes -> visible to TMA
                    __syncwarp();
                    if (lane == 0) {
                        if (row0 < p.M) ptx::tma_reduce_add_2d(&tmap_out, buf, col, row0);
                        ptx::bulk_commit();
                    }
                    ++chunk_no;
                } else {
                    // thread = accumulator row  ->  XOR-swizzled smem tile  ->  lane = 4 columns
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 v;
                        v.x = __uint_as_float(r[j * 4 + 0]); v.y = __uint_as_float(r[j * 4 + 1]);
                        v.z = __uint_as_float(r[j * 4 + 2]); v.w = __uint_as_float(r[j * 4 + 3]);
                        *reinterpret_cast<float4*>(st + lane * 32 + ((j ^ (lane & 7)) << 2)) = v;
                    }
                    __syncwarp();
                    const int c4 = lane & 7;            // float4 column inside the 32-column chunk
                    const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col) + c4);
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int rr = it * 4 + (lane >> 3);
                        const int grow = row0 + rr;
                        if (grow < p.M) {
                            const float4 a = *reinterpret_cast<const float4*>(st + rr * 32 + ((c4 ^ (rr & 7)) << 2));
                            float* out = reinterpret_cast<float*>(p.out) + (size_t)grow * p.ldo + col + c4 * 4;
                            float4 x = __ldg(reinterpret_cast<const float4*>(
                                p.pos + (size_t)(grow % p.pos_rows) * p.N + col + c4 * 4));
                            x.x += a.x + b.x; x.y += a.y + b.y; x.z += a.z + b.z; x.w += a.w + b.w;
                            *reinterpret_cast<float4*>(out) = x;
                        }
                    }
                }
            }
            // this warp has finished reading the accumulator stage
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (CTA_GROUP == 1) ptx::mbar_arrive(&tmem_empty[acc]);
                else ptx::mbar_arrive_cluster(&tmem_empty[acc], 0);
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
        if constexpr (EPI == EPI_BIAS_RESID_F32) {
            if (lane == 0) ptx::bulk_wait_all();        // all reduce-adds have landed in global memory
        }
    }

    // ===================== teardown =====================
    __syncwarp();  // reconverge the single-lane roles before the aligned barriers below
    ptx::tc_fence_before();
    if constexpr (CTA_GROUP == 2) ptx::cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<CTA_GROUP>(tmem_base, Cfg::TMEM_COLS);
    }
}

}  // namespace hm
