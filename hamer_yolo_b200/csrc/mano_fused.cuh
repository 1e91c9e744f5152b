// MANO layer as ONE kernel: kinematic chain prologue, blend-shape GEMM on tcgen05, linear blend skinning as the
// GEMM's epilogue.
//
// Replaces MANO.forward (hamer/hamer/models/mano_wrapper.py:32-44 -> smplx 0.1.28 lbs(): blend_shapes,
// vertices2joints, batch_rigid_transform, skinning) and perspective_projection
// (hamer/hamer/utils/geometry.py:72-109) as called from HAMER.forward_step (hamer/hamer/models/hamer.py:143-152).
// The only HBM traffic is the algorithmic one: 616 B of betas / rotations in, 9,336 B of vertices and 420 B of
// joints / 2-D keypoints out per hand.  v_posed, the blend coefficients and the skinning transforms never leave
// the SM (round 1 had a separate prep kernel that round-tripped 1,088 B per hand and cost 18 % of the time).
//
//   work unit  = 256 hands (a CTA pair, UMMA M = 256) x a range of the 25 coordinate tiles (N = 96 = 32 vertices).
//   inputs     : the betas and rotations of a unit's 128 hands per CTA (79 KB) are fetched by one bulk copy
//                (cp.async.bulk, issued by an otherwise idle warp) while the previous unit is still being skinned.
//   prologue   : per unit, thread = (hand, part): the thread reads its hand's betas and the rotations it needs from
//                shared memory, evaluates the rest joints from the folded joint regressor (constant-bank operands),
//                walks the MANO kinematic tree for its own joints (ancestors recomputed - the tree is 3 levels
//                deep) and writes the skinning transforms A_j = [G_j | t_j - G_j J_j] and its share of the fp16
//                blend coefficients straight into tensor memory (tcgen05.st); chain joints (+ projection) are
//                transposed through the warp's staging tile so that each hand's floats leave as contiguous runs.
//                The tree is the standard MANO one, fixed at compile time so that every index is static (other
//                trees take the three-kernel path of mano_lbs.cuh).
//   tensor     : 512 columns per CTA, lane = hand:  [0,192) two fp32 accumulators, [192,384) the 16 skinning
//   memory       transforms of the hand (12 columns per joint), [384,464) its 160 fp16 blend coefficients.
//   MMA        : D[256 x 96] = coef[256 x 160] x basis^T as 10 k-steps, A from tensor memory (TS form), B = the
//                constant fp16 basis streamed by TMA, one ring stage per coordinate tile.
//                K layout: [shape hi (10) | pose (135) | shape lo (10) | template hi | template lo | 0 0 0] with
//                coefficients [beta | vec(R - I) | beta | 1 | 1 | 0 0 0]: the accumulator IS v_posed (template
//                included, exact to 2^-22 like the shape basis; the pose basis is rounded to fp16, <= 0.004 mm
//                on MANO).
//   epilogue   : thread = hand (TMEM lane), MF_VPT vertices per thread per tile.  out_v = sum_u w_vu (R_u p_v + t_u)
//                over the union of joints of the vertex group - weights and joint ids are warp-uniform (broadcast
//                reads of a resident table), vertex pairs without a weight for the joint are skipped, arithmetic
//                is packed FFMA2 over vertex pairs (basis rows are ordered x0 x1 y0 y1 z0 z1 per pair, so the
//                accumulator registers already are the pairs).
//   stores     : what bounds the store side is the number of L1 -> L2 requests, i.e. how many contiguous bytes one
//                warp instruction covers (tests/cuda/tc_microbench.cu: 612 MB of vertex rows take 0.68 ms as
//                8-byte thread-per-hand stores, 0.45 ms as 16-byte ones, 0.18 ms in 96-byte runs, 0.15 ms in
//                256-byte runs; cp.async.bulk.tensor cannot address the odd rows - the 9,336-byte row pitch is
//                8 mod 16).  Every warp therefore transposes its [32 hands x MF_GCOLS floats] tile through a
//                private shared-memory tile (row pitch + 16 B: the thread-per-row STS.128 are conflict free) and
//                stores 8 bytes per lane with the lanes running along the hand's row.
#pragma once
#include "mano_lbs.cuh"
#include "ptx.cuh"

namespace hm {

constexpr int MF_BM = 128;                 // hands per CTA
constexpr int MF_BN = 96;                  // coordinates per tile (32 vertices)
constexpr int MF_NT = 25;                  // 25 x 96 = 2400 >= 2334
#ifndef HM_MF_VPT
#define HM_MF_VPT 8
#endif
constexpr int MF_VPT = HM_MF_VPT;          // vertices per thread per tile (8: 16 skinning warps, 16: 8 skinning warps)
constexpr int MF_GCOLS = MF_VPT * 3;       // accumulator columns per thread per tile
constexpr int MF_PARTS = MF_BN / MF_GCOLS; // column slices per tile = warps per TMEM lane quadrant
constexpr int MF_NP = MF_VPT / 2;          // vertex pairs per thread
constexpr int MF_GROUPS = MF_NT * MF_PARTS;// vertex groups (the trailing ones are padding)
constexpr int MF_LIVE_GROUPS = (MANO_V + MF_VPT - 1) / MF_VPT;
constexpr int MF_JPP = MANO_J / MF_PARTS;  // joints whose transforms a thread of part p produces: [p JPP, (p+1) JPP)
constexpr int MF_CPP = (MANO_KB / 2) / MF_PARTS;   // packed coefficient columns per part
constexpr int MF_KSTEPS = MANO_KB / 16;    // 10 UMMA k-steps
constexpr int MF_KBLOCKS = (MANO_KB + 63) / 64; // 3 basis boxes of 64 columns per tile (the last one half used)
constexpr int MF_BSTAGES = 2;              // one stage = one whole coordinate tile (one barrier wait + one commit per 10 MMAs)
constexpr int MF_EPI_WARPS = 4 * MF_PARTS;
constexpr int MF_THREADS = 128 + 32 * MF_EPI_WARPS;   // warpgroup 0: basis producer, MMA issuer, input producer (+1 idle); then the skinning warps
constexpr int MF_MAX_UNION = 12;           // union slots per vertex group (MANO needs 10; larger unions take the three-kernel path)
constexpr uint32_t MF_B_BLOCK = (MF_BN / 2) * 128;            // 48 basis rows x 64 columns
constexpr uint32_t MF_B_STAGE = MF_KBLOCKS * MF_B_BLOCK;      // 18 KB per CTA per coordinate tile
constexpr uint32_t MF_PITCH = MF_GCOLS * 4 + 16;              // staging row pitch: an odd number of 16-byte chunks
constexpr uint32_t MF_WARP_STAGE = 32 * MF_PITCH;             // per skinning warp: [32 hands][MF_GCOLS floats (+ 16 B)]
constexpr uint32_t MF_OFF_STAGING = MF_BSTAGES * MF_B_STAGE;
// per vertex-group table: int [0] union size, [1] fingertip mask, [4, 4+12) per union slot (12 x joint) |
// (mask of the vertex pairs with a non-zero weight) << 16, then float weights [12 slots][MF_VPT]
constexpr uint32_t MF_TAB_STRIDE = 16 + 4 * MF_MAX_UNION + 4 * MF_MAX_UNION * MF_VPT;
constexpr uint32_t MF_TAB_BYTES = MF_LIVE_GROUPS * MF_TAB_STRIDE;
constexpr uint32_t MF_OFF_TAB = MF_OFF_STAGING + MF_EPI_WARPS * MF_WARP_STAGE;
// raw inputs of the CTA's 128 hands: rotations [128][16 B of padding + 144 floats] then betas [128][10].  The 592-byte row
// pitch makes the thread-per-hand LDS.128 conflict free (37 16-byte chunks: odd), the padding lets every part read its
// coefficient range with aligned float4 loads
constexpr uint32_t MF_IN_PITCH = 16 + 144 * 4;
constexpr uint32_t MF_IN_ROT = MF_BM * MF_IN_PITCH, MF_IN_BETA = MF_BM * 10 * 4;
constexpr uint32_t MF_OFF_IN = MF_OFF_TAB + ((MF_TAB_BYTES + 127u) & ~127u);
constexpr uint32_t MF_OFF_BARS = MF_OFF_IN + MF_IN_ROT + MF_IN_BETA;
constexpr uint32_t MF_SMEM_BYTES = 1024 + MF_OFF_BARS + 256;
constexpr uint32_t MF_TMEM_COLS = 512;
constexpr uint32_t MF_AU_COL = 2 * MF_BN;                     // transforms: joint j at MF_AU_COL + 12 j
constexpr uint32_t MF_COEF_COL = MF_AU_COL + MANO_J * 12;     // coefficients: k-step s at MF_COEF_COL + 8 s
constexpr int MF_K_TMPL = 155;                                // coefficient / basis column of the template (hi; lo at + 1)
// setmaxnreg moves registers inside the CTA's launch allocation only (MF_THREADS x the compiled register count): the
// re-balanced total must fit or the .inc blocks forever
constexpr int MF_LAUNCH_REGS = (65536 / MF_THREADS) / 8 * 8;
constexpr int MF_SKIN_REGS_RAW = ((MF_THREADS * MF_LAUNCH_REGS - 128 * 40) / (32 * MF_EPI_WARPS)) / 8 * 8;
constexpr int MF_SKIN_REGS = MF_SKIN_REGS_RAW > 232 ? 232 : MF_SKIN_REGS_RAW;
static_assert(128 * 40 + 32 * MF_EPI_WARPS * MF_SKIN_REGS <= MF_THREADS * MF_LAUNCH_REGS, "register pool");
static_assert(MF_SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(MF_COEF_COL + MANO_KB / 2 <= MF_TMEM_COLS, "tensor memory budget");
static_assert(MF_TAB_STRIDE % 16 == 0 && MF_OFF_TAB % 128 == 0 && MF_OFF_IN % 128 == 0 && (MF_PITCH / 16) % 2 == 1, "alignment");
static_assert(MF_JPP * 12 % 16 == 0 && (MF_CPP % 16 == 4 || MF_CPP % 16 == 8), "tensor-memory store chunking");
static_assert(MF_JPP * 5 * 4 <= MF_GCOLS * 4, "chain joints are transposed through the staging tile");
static_assert(96 % (MF_GCOLS / 2) == 0, "store read-out period");

// The standard MANO kinematic tree (smplx / manopth): wrist + 5 chains of 3.  The prologue is specialised for it.
__host__ __device__ constexpr int mf_parent(int j) { return j == 0 ? -1 : ((j - 1) % 3 == 0 ? 0 : j - 1); }
// is joint a an ancestor of (or equal to) joint j?
__host__ __device__ constexpr bool mf_on_chain(int a, int j) {
    while (j >= 0) { if (j == a) return true; j = mf_parent(j); }
    return false;
}
// Own joints of part p: the threads (hand, p) produce the skinning transforms of these joints.  Finger chains stay
// within one part where possible so that the fewest ancestors are recomputed: 4 parts -> {0,1,2,3} {4,5,6,13} {7,8,9,14}
// {10,11,12,15}; 2 parts -> {0..7} {8..15}.
__host__ __device__ constexpr int mf_own_joint(int p, int jj) {
    if (MF_PARTS == 4) return jj < 3 ? 3 * p + jj + (p == 0 ? 0 : 1) : (p == 0 ? 3 : 12 + p);
    return p * 8 + jj;
}
__host__ __device__ constexpr bool mf_owns(int p, int j) {
    for (int jj = 0; jj < MF_JPP; ++jj) if (mf_own_joint(p, jj) == j) return true;
    return false;
}
// does part p walk through joint a (own joints and their ancestors)?
__host__ __device__ constexpr bool mf_needs(int p, int a) {
    for (int jj = 0; jj < MF_JPP; ++jj) if (mf_on_chain(a, mf_own_joint(p, jj))) return true;
    return false;
}
// The walk of part p as a code word: 6 bits per step, lowest first - joint (4) | parent is the root, not the previous
// step (1) | own (1) - and the number of steps in bits 60..63.  Joints are visited in increasing order, which for the
// MANO tree means: every joint's parent is either the previous step or the root.
__host__ __device__ constexpr uint64_t mf_chain_code(int p) {
    uint64_t code = 0;
    int n = 0, prev = -1;
    for (int j = 0; j < MANO_J; ++j) {
        if (!mf_needs(p, j)) continue;
        const uint64_t from_root = (j != 0 && mf_parent(j) != prev) ? 1 : 0;
        code |= ((uint64_t)j | (from_root << 4) | ((uint64_t)(mf_owns(p, j) ? 1 : 0) << 5)) << (6 * n);
        prev = j;
        ++n;
    }
    return code | ((uint64_t)n << 60);
}
// bit il: coefficient k = p NC + il of part p is a diagonal entry of a rotation matrix (k in [10, 145), (k - 10) % 9 in {0,4,8})
__host__ __device__ constexpr uint64_t mf_diag_mask(int p) {
    uint64_t m = 0;
    for (int il = 0; il < 2 * MF_CPP && il < 64; ++il) {
        const int k = p * 2 * MF_CPP + il;
        if (k >= 10 && k < 145 && ((k - 10) % 9) % 4 == 0) m |= 1ull << il;
    }
    return m;
}
__host__ __device__ constexpr uint32_t mf_diag_mask_hi(int p) {
    uint32_t m = 0;
    for (int il = 64; il < 2 * MF_CPP; ++il) {
        const int k = p * 2 * MF_CPP + il;
        if (k >= 10 && k < 145 && ((k - 10) % 9) % 4 == 0) m |= 1u << (il - 64);
    }
    return m;
}
__host__ __device__ constexpr bool mf_chain_ok(int p) {
    int prev = -1;
    for (int j = 0; j < MANO_J; ++j) {
        if (!mf_needs(p, j)) continue;
        if (j != 0 && mf_parent(j) != prev && mf_parent(j) != 0) return false;
        prev = j;
    }
    return true;
}
static_assert(mf_chain_ok(0) && mf_chain_ok(1) && mf_chain_ok(MF_PARTS - 2) && mf_chain_ok(MF_PARTS - 1), "walk encoding");

// Work list: `whole` pair tiles are done in one piece each (all 25 coordinate tiles, one prologue); the remaining pair
// tiles - the partial last wave - are cut into n_split pieces so that the wave ends together (same idea as the K-split of
// the residual GEMMs' tail tiles).
struct ManoItem { int pt, q, nt0, nt1; };
__device__ __forceinline__ ManoItem mano_item(int unit, int whole, int n_split) {
    ManoItem it;
    if (unit < whole) { it.pt = unit; it.q = 0; it.nt0 = 0; it.nt1 = MF_NT; return it; }
    const int r = unit - whole;
    it.pt = whole + r / n_split;
    it.q = r - (r / n_split) * n_split;
    it.nt0 = it.q * MF_NT / n_split;
    it.nt1 = (it.q + 1) * MF_NT / n_split;
    return it;
}

struct ManoFusedArgs {
    const float* betas;     // [B,10], 16-byte aligned
    const float* rotmats;   // [B,16,3,3], 16-byte aligned
    const float* cam_t;     // [B,3] or null
    float* verts;           // [B,778,3], 8-byte aligned
    float* joints;          // [B,21,3]
    float* kp2d;            // [B,21,2] or null
    float focal_over_size;
    int B;
    int whole;              // pair tiles (256 hands) processed as one unit each
    int n_split;            // the remaining pair tiles are cut into n_split units (contiguous ranges of coordinate tiles)
    int dbg;                // diagnostics (HM_MANO_DBG): 1 = no vertex stores, 2 = no skinning loop, 4 = accumulator read only
    long long* clk;         // diagnostics: per CTA and unit (first 8), 6 phase clocks of skinning warp 0 (or null)
};

// constant-bank copy of the folded joint regressor: J[j][c] = jreg[j][c][0] + sum_l jreg[j][c][1 + l] beta_l
struct ManoChainConst {
    float jreg[MANO_J][3][11];
    int32_t chain_slot[MANO_J];
    int32_t tip_slot[5];
    int32_t tip_ids[5];
};

// ---- prologue: one thread = one hand x one part ------------------------------------------------------------
// s_rot: this hand's row of the prefetched rotations (16 bytes of padding, then 144 floats; shared-window address);
// s_beta: its 10 betas; stg: the thread's row of its warp's staging tile.  Deliberately compact, rolled code: the
// prologue runs once per unit, so a fully unrolled, per-part specialised version (6,000 instructions) was bound by
// instruction fetch - 20-40 k clocks per unit.  Returns with the chain joints of the part ([jj][x y z u v], 5 floats
// each, in own-joint order) in the thread's staging row.
__device__ __forceinline__ void mf_prologue(const ManoChainConst& cc, const ManoFusedArgs& a, int part, uint32_t s_rot,
                                            uint32_t s_beta, float ct0, float ct1, float ct2, uint32_t stg,
                                            uint32_t tmem_lane, uint64_t* in_empty, int lane) {
    float be[10];
#pragma unroll
    for (int i = 0; i < 5; ++i) {                   // 40-byte rows: 8-byte aligned
        const float2 v = ptx::lds64(s_beta + 8 * i);
        be[2 * i] = v.x; be[2 * i + 1] = v.y;
    }
    // ---- blend coefficients [beta | vec(R - I) | beta | 1 | 1 | 0 0 0]: this part's k in [K0, K0 + 2 CPP) -> tensor memory.
    // The rotation float behind k in [10, 145) is k - 1, i.e. byte 16 + 4 (k - 1) of the padded row: with K0 a multiple of 8
    // the part reads the aligned float4s from byte 4 K0 on and coefficient i of the part is element i + 3 of what it read.
    {
        constexpr int NC = 2 * MF_CPP;               // coefficients per part
        const int K0 = part * NC;
        float L[NC + 4];
#pragma unroll
        for (int q = 0; q < (NC + 4) / 4; ++q) {
            const float4 v = ptx::lds128(s_rot + (uint32_t)(4 * K0 + 16 * q));
            L[4 * q] = v.x; L[4 * q + 1] = v.y; L[4 * q + 2] = v.z; L[4 * q + 3] = v.w;
        }
        // diagonal entries of the rotations ((k - 10) % 9 in {0, 4, 8}, k in [10, 145)) as one bit mask per part
        constexpr uint64_t D0 = mf_diag_mask(0), D1 = mf_diag_mask(1 % MF_PARTS), D2 = mf_diag_mask(2 % MF_PARTS), D3 = mf_diag_mask(3 % MF_PARTS);
        const uint64_t dm = part == 0 ? D0 : part == 1 ? D1 : part == 2 ? D2 : D3;
        const uint32_t dlo = (uint32_t)dm, dhi = (uint32_t)(dm >> 32);
        float v[NC];
#pragma unroll
        for (int il = 0; il < NC; ++il) {
            static_assert(NC <= 64 || MF_PARTS == 2, "mask width");
            const uint32_t bit = il < 32 ? (dlo >> il) & 1u : il < 64 ? (dhi >> (il - 32)) & 1u : 0u;
            v[il] = L[il + 3] - (bit ? 1.0f : 0.0f);
        }
        if constexpr (NC > 64) {           // 2 parts: 80 coefficients each, the mask of elements 64..79 separately
            constexpr uint32_t E0 = mf_diag_mask_hi(0), E1 = mf_diag_mask_hi(1 % MF_PARTS);
            const uint32_t de = part == 0 ? E0 : E1;
#pragma unroll
            for (int il = 64; il < NC; ++il) v[il] = L[il + 3] - (((de >> (il - 64)) & 1u) ? 1.0f : 0.0f);
        }
        if (part == 0) {                    // k < 10: betas
#pragma unroll
            for (int il = 0; il < 10; ++il) v[il] = be[il];
        }
        if (part == MF_PARTS - 1) {         // k >= 145: betas again (shape lo columns), 1 1 (template hi / lo), zeros
            constexpr int OFF = 145 - (MF_PARTS - 1) * NC;
#pragma unroll
            for (int il = OFF; il < NC; ++il) v[il] = il < OFF + 10 ? be[il - OFF] : (il < OFF + 12 ? 1.0f : 0.0f);
        }
        uint32_t cp[MF_CPP];
#pragma unroll
        for (int i = 0; i < MF_CPP; ++i) cp[i] = ptx::pack_half2_sat(v[2 * i], v[2 * i + 1]);
        uint32_t c16[16];
#pragma unroll
        for (int ch = 0; ch < MF_CPP / 16; ++ch) {
#pragma unroll
            for (int i = 0; i < 16; ++i) c16[i] = cp[ch * 16 + i];
            ptx::tmem_st_32x16(tmem_lane + MF_COEF_COL + (uint32_t)(part * MF_CPP + ch * 16), c16);
        }
        constexpr int REM = MF_CPP % 16, RB = MF_CPP - REM;
        if constexpr (REM == 4) {
            uint32_t c4[4] = {cp[RB], cp[RB + 1], cp[RB + 2], cp[RB + 3]};
            ptx::tmem_st_32x4(tmem_lane + MF_COEF_COL + (uint32_t)(part * MF_CPP + RB), c4);
        } else if constexpr (REM == 8) {
            uint32_t c8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) c8[i] = cp[RB + i];
            ptx::tmem_st_32x8(tmem_lane + MF_COEF_COL + (uint32_t)(part * MF_CPP + RB), c8);
        }
    }
    // ---- kinematic chain: walk this part's joints (code word, 6 bits per step) --------------------------------------
    constexpr uint64_t C0 = mf_chain_code(0), C1 = mf_chain_code(1 % MF_PARTS), C2 = mf_chain_code(2 % MF_PARTS),
                       C3 = mf_chain_code(3 % MF_PARTS);
    uint64_t code = part == 0 ? C0 : part == 1 ? C1 : part == 2 ? C2 : C3;
    const int nsteps = (int)(code >> 60);
    float Gp[12], Jp[3], Gr[12], Jroot[3];
#pragma unroll
    for (int e = 0; e < 12; ++e) { Gp[e] = 0.f; Gr[e] = 0.f; }
    Jp[0] = Jp[1] = Jp[2] = Jroot[0] = Jroot[1] = Jroot[2] = 0.f;
    uint32_t sj = stg;                                                       // staging slot of the next own joint
#pragma unroll 1
    for (int st = 0; st < nsteps; ++st, code >>= 6) {
        const int j = (int)(code & 15u);
        const bool from_root = (code >> 4) & 1u, own = (code >> 5) & 1u;
        // rest joint from the folded regressor (constant bank, warp-uniform index)
        float Jj[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v = cc.jreg[j][c][0];
#pragma unroll
            for (int l = 0; l < 10; ++l) v = fmaf(cc.jreg[j][c][1 + l], be[l], v);
            Jj[c] = v;
        }
        // R_j: 9 floats from byte 16 + 36 j of the row = three aligned float4s, shifted by j mod 4 floats
        float Rj[9];
        {
            const uint32_t a0 = s_rot + 16u + ((36u * (uint32_t)j) & ~15u);
            const float4 q0 = ptx::lds128(a0), q1 = ptx::lds128(a0 + 16), q2 = ptx::lds128(a0 + 32);
            const float Lq[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
            const int sh = j & 3;                                            // warp-uniform
            if (sh == 0) {
#pragma unroll
                for (int e = 0; e < 9; ++e) Rj[e] = Lq[e];
            } else if (sh == 1) {
#pragma unroll
                for (int e = 0; e < 9; ++e) Rj[e] = Lq[e + 1];
            } else if (sh == 2) {
#pragma unroll
                for (int e = 0; e < 9; ++e) Rj[e] = Lq[e + 2];
            } else {
#pragma unroll
                for (int e = 0; e < 9; ++e) Rj[e] = Lq[e + 3];
            }
        }
        float G[12];
        if (st == 0) {                                                       // the root: G = [R | J]
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                G[r * 4 + 0] = Rj[r * 3 + 0]; G[r * 4 + 1] = Rj[r * 3 + 1]; G[r * 4 + 2] = Rj[r * 3 + 2];
                G[r * 4 + 3] = Jj[r];
            }
#pragma unroll
            for (int e = 0; e < 12; ++e) Gr[e] = G[e];
            Jroot[0] = Jj[0]; Jroot[1] = Jj[1]; Jroot[2] = Jj[2];
        } else {                                                             // G = G_parent o [R_j | J_j - J_parent]
            if (from_root) {
#pragma unroll
                for (int e = 0; e < 12; ++e) Gp[e] = Gr[e];
                Jp[0] = Jroot[0]; Jp[1] = Jroot[1]; Jp[2] = Jroot[2];
            }
            const float t0 = Jj[0] - Jp[0], t1 = Jj[1] - Jp[1], t2 = Jj[2] - Jp[2];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const float p0 = Gp[r * 4 + 0], p1 = Gp[r * 4 + 1], p2 = Gp[r * 4 + 2];
#pragma unroll
                for (int c = 0; c < 3; ++c) G[r * 4 + c] = p0 * Rj[c] + p1 * Rj[3 + c] + p2 * Rj[6 + c];
                G[r * 4 + 3] = p0 * t0 + p1 * t1 + p2 * t2 + Gp[r * 4 + 3];
            }
        }
        if (own) {       // skinning transform A_j = [G | t - G J_j] -> tensor memory; posed joint (+ projection) -> staging row
            uint32_t a8[8], a4[4];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const float tr = G[r * 4 + 3] - (G[r * 4 + 0] * Jj[0] + G[r * 4 + 1] * Jj[1] + G[r * 4 + 2] * Jj[2]);
                const float row[4] = {G[r * 4 + 0], G[r * 4 + 1], G[r * 4 + 2], tr};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (r < 2) a8[r * 4 + c] = __float_as_uint(row[c]); else a4[c] = __float_as_uint(row[c]);
                }
            }
            const uint32_t col = tmem_lane + MF_AU_COL + 12u * (uint32_t)j;
            ptx::tmem_st_32x8(col, a8);
            ptx::tmem_st_32x4(col + 8, a4);
            const float X = G[3] + ct0, Y = G[7] + ct1, Z = G[11] + ct2;
            const float fz = __fdividef(a.focal_over_size, Z);              // 2 ulp: 1e-5 px on the 256-pixel crop
            ptx::sts32(sj, G[3]);
            ptx::sts32(sj + 4, G[7]);
            ptx::sts32(sj + 8, G[11]);
            ptx::sts32(sj + 12, X * fz);
            ptx::sts32(sj + 16, Y * fz);
            sj += 20;
        }
#pragma unroll
        for (int e = 0; e < 12; ++e) Gp[e] = G[e];
        Jp[0] = Jj[0]; Jp[1] = Jj[1]; Jp[2] = Jj[2];
    }
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(in_empty);      // this warp has read its inputs: the next unit's may land
    ptx::tmem_st_wait();
}

__global__ void __launch_bounds__(MF_THREADS, 1)
mano_fused_kernel(const __grid_constant__ CUtensorMap tm_basis, const __grid_constant__ ManoChainConst cc,
                  const float4* __restrict__ fz_tab, const ManoFusedArgs a) {
    extern __shared__ uint8_t mf_smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(mf_smem_raw);
    uint8_t* smem = mf_smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* sB = smem;
    uint8_t* sStage = smem + MF_OFF_STAGING;
    uint8_t* sTab = smem + MF_OFF_TAB;
    uint8_t* sIn = smem + MF_OFF_IN;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MF_OFF_BARS);
    uint64_t* full_b = bars;                       // [MF_BSTAGES] basis tile landed (leader's)
    uint64_t* empty_b = bars + MF_BSTAGES;         // [MF_BSTAGES] MMAs that read the tile retired (both CTAs)
    uint64_t* tmem_full = bars + 2 * MF_BSTAGES;   // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2] (leader's)
    uint64_t* coef_full = tmem_empty + 2;          // both CTAs' coefficients are in tensor memory (leader's)
    uint64_t* in_full = coef_full + 1;             // this CTA's inputs of the unit have landed
    uint64_t* in_empty = in_full + 1;              // every skinning warp has read them
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(in_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const bool is_leader = rank == 0;
    const int worker = blockIdx.x >> 1, num_workers = gridDim.x >> 1;
    const int pair_tiles = (a.B + 2 * MF_BM - 1) / (2 * MF_BM);
    const int units = a.whole + (pair_tiles - a.whole) * a.n_split;
    ptx::griddep_launch_dependents();   // PDL: the next kernel may start its launch / set-up

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tm_basis);
        for (int s = 0; s < MF_BSTAGES; ++s) { ptx::mbar_init(&full_b[s], 1); ptx::mbar_init(&empty_b[s], 1); }
        for (int s = 0; s < 2; ++s) { ptx::mbar_init(&tmem_full[s], 1); ptx::mbar_init(&tmem_empty[s], MF_EPI_WARPS * 2); }
        ptx::mbar_init(coef_full, MF_EPI_WARPS * 2);
        ptx::mbar_init(in_full, 1);
        ptx::mbar_init(in_empty, MF_EPI_WARPS);
        ptx::fence_mbar_init();
    }
    if (warp == 1) ptx::tmem_alloc<2>(tmem_slot, MF_TMEM_COLS);
    ptx::tc_fence_before();
    ptx::cluster_sync_all();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::griddep_wait();                // betas / rotations / camera come from the previous kernel

    // Register re-balancing between warpgroups: the single-lane roles give registers back, the skinning warps take them.
    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == 0) {
            // ===================== producer: basis tiles =====================
            if (lane == 0) {
                int stage = 0;
                uint32_t phase = 0;
                for (int unit = worker; unit < units; unit += num_workers) {
                    const ManoItem wi = mano_item(unit, a.whole, a.n_split);
                    for (int nt = wi.nt0; nt < wi.nt1; ++nt) {
                        ptx::mbar_wait(&empty_b[stage], phase ^ 1u);
                        if (is_leader) ptx::mbar_arrive_expect_tx(&full_b[stage], 2 * MF_B_STAGE);
#pragma unroll
                        for (int kb = 0; kb < MF_KBLOCKS; ++kb)        // columns >= 160 of the last box are zero-filled
                            ptx::tma_load_2d_2sm(sB + stage * MF_B_STAGE + kb * MF_B_BLOCK, &tm_basis, &full_b[stage],
                                                 kb * 64, nt * MF_BN + (int)rank * (MF_BN / 2));
                        if (++stage == MF_BSTAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        } else if (warp == 1) {
            // ===================== MMA issuer (leader CTA, one lane) =====================
            if (is_leader && lane == 0) {
                constexpr uint32_t idesc = ptx::make_idesc_f16kind_f32(0, 0, 2 * MF_BM, MF_BN);
                int stage = 0, acc = 0;
                uint32_t phase = 0, acc_phase = 0, it = 0;
                for (int unit = worker; unit < units; unit += num_workers, ++it) {
                    const ManoItem wi = mano_item(unit, a.whole, a.n_split);
                    ptx::mbar_wait(coef_full, it & 1u);
                    ptx::tc_fence_after();
                    for (int nt = wi.nt0; nt < wi.nt1; ++nt) {
                        ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
                        ptx::mbar_wait(&full_b[stage], phase);
                        ptx::tc_fence_after();
                        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * MF_BN);
#pragma unroll
                        for (int kb = 0; kb < MF_KBLOCKS; ++kb) {
                            const uint64_t b_desc =
                                ptx::make_smem_desc_k_sw128(ptx::smem_u32(sB + stage * MF_B_STAGE + kb * MF_B_BLOCK));
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const int t = kb * 4 + k;                           // k-step
                                if (t < MF_KSTEPS)
                                    ptx::mma_f16_ts_2sm(d_tmem, tmem_base + MF_COEF_COL + (uint32_t)(t * 8),
                                                        b_desc + (uint64_t)(k * 2), idesc, (uint32_t)(t != 0));
                            }
                        }
                        ptx::mma_commit_2sm(&empty_b[stage], 0x3);
                        ptx::mma_commit_2sm(&tmem_full[acc], 0x3);
                        if (++stage == MF_BSTAGES) { stage = 0; phase ^= 1u; }
                        acc ^= 1;
                        if (acc == 0) acc_phase ^= 1u;
                    }
                }
            }
        } else if (warp == 2) {
            // ===================== producer: raw inputs of this CTA's 128 hands, one unit ahead =====================
            uint32_t it = 0;
            for (int unit = worker; unit < units; unit += num_workers, ++it) {
                const int pt = mano_item(unit, a.whole, a.n_split).pt;
                const int h0 = (pt * 2 + (int)rank) * MF_BM;                      // first hand of this CTA's tile
                ptx::mbar_wait(in_empty, (it & 1u) ^ 1u);                         // the previous unit's inputs have been read
                int n = a.B - h0;
                n = n < 0 ? 0 : (n > MF_BM ? MF_BM : n);
                // bulk-copy sizes are multiples of 16 B: a hand's 10 betas are 40 B, so an odd count is rounded down (the
                // last hand of an odd batch is fetched by its own thread)
                const int nb = n & ~1;
                if (lane == 0) {
                    ptx::mbar_arrive_expect_tx(in_full, (uint32_t)(n * 576 + nb * 40));
                    if (nb > 0) ptx::bulk_load_1d(sIn + MF_IN_ROT, a.betas + (size_t)h0 * 10, (uint32_t)(nb * 40), in_full);
                }
                __syncwarp();
                for (int h = lane; h < n; h += 32)                               // one 576-byte copy per hand into the padded rows
                    ptx::bulk_load_1d(sIn + (uint32_t)h * MF_IN_PITCH + 16, a.rotmats + (size_t)(h0 + h) * 144, 576u, in_full);
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(MF_SKIN_REGS));
        // ===================== skinning warps: thread = hand =====================
        const int ew = warp - 4, quad = warp & 3, part = ew >> 2;
        const int hl = quad * 32 + lane;                                   // hand inside this CTA's tile = TMEM lane
        const uint32_t tmem_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
        const uint32_t sbase = ptx::smem_u32(sStage) + (uint32_t)ew * MF_WARP_STAGE;   // this warp's staging tile
        const uint32_t srow = sbase + (uint32_t)lane * MF_PITCH;                        // the thread's own row
        const uint32_t tab0 = ptx::smem_u32(sTab);
        // read-out of the staging tile: element e = 32 t + lane of the [32 rows x HALF] matrix of 8-byte pairs; three
        // consecutive t advance RS rows, so the offsets are per-lane constants with period 3
        constexpr int HALF = MF_GCOLS / 2, RS = 96 / HALF;
        int rd_row[3], rd_pos[3];
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) { const int e = kk * 32 + lane; rd_row[kk] = e / HALF; rd_pos[kk] = e % HALF; }
        // chain-joint flush: lane l < 3 JPP writes coordinate l % 3 of own joint l / 3 (3-D joints), lane l < 2 JPP the
        // projection l % 2 of own joint l / 2; the per-lane offsets are fixed for the whole kernel
        int j3_dst = -1, j2_dst = -1;
#pragma unroll
        for (int pp = 0; pp < MF_PARTS; ++pp)
#pragma unroll
            for (int jj = 0; jj < MF_JPP; ++jj) {
                const int slot = cc.chain_slot[mf_own_joint(pp, jj)];
                if (pp == part && slot >= 0) {
                    if (lane / 3 == jj && lane < 3 * MF_JPP) j3_dst = slot * 3 + lane % 3;
                    if (lane / 2 == jj && lane < 2 * MF_JPP) j2_dst = slot * 2 + lane % 2;
                }
            }
        const uint32_t j3_src = (uint32_t)((lane / 3) * 20 + 4 * (lane % 3)), j2_src = (uint32_t)((lane / 2) * 20 + 12 + 4 * (lane % 2));
        // the constant group tables (joint unions, pair masks, weights) are loaded once
        for (int i = (int)threadIdx.x - 128; i < (int)(MF_TAB_BYTES / 16); i += 32 * MF_EPI_WARPS)
            ptx::sts128(tab0 + 16 * i, __ldg(fz_tab + i));
        int acc = 0;
        uint32_t acc_phase = 0, it = 0;
        for (int unit = worker; unit < units; unit += num_workers, ++it) {
            const ManoItem wi = mano_item(unit, a.whole, a.n_split);
            const int q = wi.q, nt0 = wi.nt0, nt1 = wi.nt1;
            const int tile128 = wi.pt * 2 + (int)rank;
            const int row0 = tile128 * MF_BM + quad * 32;                  // first hand of this warp
            const int hand = row0 + lane;
            const bool live = hand < a.B;
            float ct0 = 0.f, ct1 = 0.f, ct2 = 1.f;
            if (live && a.kp2d != nullptr) {
                const float* ct = a.cam_t + (size_t)hand * 3;
                ct0 = __ldg(ct); ct1 = __ldg(ct + 1); ct2 = __ldg(ct + 2);
            }
            long long* ck = (a.clk != nullptr && ew == 0 && lane == 0 && it < 8) ? a.clk + ((size_t)blockIdx.x * 8 + it) * 6 : nullptr;
            if (ck) ck[0] = clock64();
            // every skinning warp has consumed the previous unit's last accumulator (so its MMAs, the readers of the
            // coefficients, have retired) and is done with the previous transforms
            asm volatile("bar.sync 1, %0;" ::"n"(32 * MF_EPI_WARPS) : "memory");
            if (ck) ck[1] = clock64();
            ptx::mbar_wait(in_full, it & 1u);
            if (ck) ck[2] = clock64();
            // the last hand of an odd batch: its betas are not part of the (16-byte granular) bulk copy
            if (live && (a.B & 1) && hand == a.B - 1) {
                const float* bp = a.betas + (size_t)hand * 10;
#pragma unroll
                for (int i = 0; i < 10; ++i) ptx::sts32(ptx::smem_u32(sIn) + MF_IN_ROT + (uint32_t)hl * 40 + 4 * i, __ldg(bp + i));
            }
            __syncwarp();
            const uint32_t s_rot = ptx::smem_u32(sIn) + (uint32_t)hl * MF_IN_PITCH, s_beta = ptx::smem_u32(sIn) + MF_IN_ROT + (uint32_t)hl * 40;
            mf_prologue(cc, a, part, s_rot, s_beta, ct0, ct1, ct2, srow, tmem_lane, in_empty, lane);
            if (ck) ck[3] = clock64();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(coef_full, 0);          // 2 x MF_EPI_WARPS warps of the pair -> the MMA issuer may start
            // ---- chain joints of the warp's 32 hands: staging rows [jj][x y z u v] -> OpenPose slots.  One hand per
            // instruction, lane = one of the part's JPP x (3 + 2) values (per-lane offsets fixed for the whole kernel)
            if (q == 0) {
                const int nrow = a.B - row0 < 32 ? a.B - row0 : 32;
                if (j3_dst >= 0) {
                    float* dst = a.joints + (size_t)row0 * 63 + j3_dst;
                    for (int r = 0; r < nrow; ++r)
                        dst[(size_t)r * 63] = __uint_as_float(ptx::lds32(sbase + (uint32_t)r * MF_PITCH + j3_src));
                }
                if (j2_dst >= 0 && a.kp2d != nullptr) {
                    float* dst = a.kp2d + (size_t)row0 * 42 + j2_dst;
                    for (int r = 0; r < nrow; ++r)
                        dst[(size_t)r * 42] = __uint_as_float(ptx::lds32(sbase + (uint32_t)r * MF_PITCH + j2_src));
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(32 * MF_EPI_WARPS) : "memory");
            ptx::tc_fence_after();
            if (ck) ck[4] = clock64();

            for (int nt = nt0; nt < nt1; ++nt) {
                const int g = nt * MF_PARTS + part;                        // vertex group
                const bool active = g < MF_LIVE_GROUPS && !(a.dbg & 4);    // trailing groups are padding (warp-uniform)
                const uint32_t tab = tab0 + (uint32_t)g * MF_TAB_STRIDE;   // never read for padding groups
                ptx::mbar_wait(&tmem_full[acc], acc_phase);
                ptx::tc_fence_after();
                const uint32_t t_addr = tmem_lane + (uint32_t)(acc * MF_BN + part * MF_GCOLS);
                uint32_t ra[MF_GCOLS];
                if constexpr (MF_VPT == 8) {
                    ptx::tmem_ld_32x16(t_addr, *reinterpret_cast<uint32_t(*)[16]>(&ra[0]));
                    ptx::tmem_ld_32x8(t_addr + 16, *reinterpret_cast<uint32_t(*)[8]>(&ra[16]));
                } else {
                    ptx::tmem_ld_32x32(t_addr, *reinterpret_cast<uint32_t(*)[32]>(&ra[0]));
                    ptx::tmem_ld_32x16(t_addr + 32, *reinterpret_cast<uint32_t(*)[16]>(&ra[32]));
                }
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive_cluster(&tmem_empty[acc], 0);   // accumulator is in registers
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
                // (padding groups skip the arithmetic but still take part in the quadrant's barriers and row stores)
                // v_posed as packed pairs P[i][c] = (p_c(v_2i), p_c(v_2i+1)): the accumulator columns of pair i are
                // x0 x1 y0 y1 z0 z1 (basis row order), template included (K columns 155 / 156)
                float2 P[MF_NP][3], O[MF_NP][3];
#pragma unroll
                for (int i = 0; i < MF_NP; ++i)
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        P[i][c] = make_float2(__uint_as_float(ra[6 * i + 2 * c]), __uint_as_float(ra[6 * i + 2 * c + 1]));
                        O[i][c] = make_float2(0.f, 0.f);
                    }
                const int nu = (!active || (a.dbg & 2)) ? 0 : (int)ptx::lds32(tab);
                const uint32_t tipmask = active ? ptx::lds32(tab + 4) : 0u;
                for (int u = 0; u < nu; ++u) {
                    const uint32_t jp = ptx::lds32(tab + 16 + u * 4);
                    const uint32_t pm = jp >> 16;                           // vertex pairs with a non-zero weight
                    uint32_t ta[8], tc[4];
                    ptx::tmem_ld_32x8(tmem_lane + MF_AU_COL + (jp & 0xffffu), ta);
                    ptx::tmem_ld_32x4(tmem_lane + MF_AU_COL + (jp & 0xffffu) + 8, tc);
                    float2 w2[MF_NP];
#pragma unroll
                    for (int i4 = 0; i4 < MF_NP / 2; ++i4) {
                        const float4 W = ptx::lds128(tab + 16 + 4 * MF_MAX_UNION + u * (4 * MF_VPT) + i4 * 16);
                        w2[2 * i4] = make_float2(W.x, W.y);
                        w2[2 * i4 + 1] = make_float2(W.z, W.w);
                    }
                    ptx::tmem_ld_wait();
                    const float4 Rc[3] = {
                        make_float4(__uint_as_float(ta[0]), __uint_as_float(ta[1]), __uint_as_float(ta[2]), __uint_as_float(ta[3])),
                        make_float4(__uint_as_float(ta[4]), __uint_as_float(ta[5]), __uint_as_float(ta[6]), __uint_as_float(ta[7])),
                        make_float4(__uint_as_float(tc[0]), __uint_as_float(tc[1]), __uint_as_float(tc[2]), __uint_as_float(tc[3]))};
#pragma unroll
                    for (int i = 0; i < MF_NP; ++i) {
                        if (pm & (1u << i)) {                               // warp-uniform: the weights do not depend on the hand
#pragma unroll
                            for (int c = 0; c < 3; ++c) {
                                float2 t = __ffma2_rn(make_float2(Rc[c].z, Rc[c].z), P[i][2], make_float2(Rc[c].w, Rc[c].w));
                                t = __ffma2_rn(make_float2(Rc[c].y, Rc[c].y), P[i][1], t);
                                t = __ffma2_rn(make_float2(Rc[c].x, Rc[c].x), P[i][0], t);
                                O[i][c] = __ffma2_rn(w2[i], t, O[i][c]);
                            }
                        }
                    }
                }
                // ---- transpose through shared memory: thread = row  ->  lanes along the row ----
                __syncwarp();                                               // the previous read-out of the staging rows is done
                if (active) {
#pragma unroll
                    for (int b = 0; b < MF_VPT / 4; ++b) {                  // 4 vertices = 12 floats = x0 y0 z0 x1 | y1 z1 x2 y2 | z2 x3 y3 z3
                        const uint32_t dst = srow + (uint32_t)b * 48;
                        ptx::sts128(dst, make_float4(O[2 * b][0].x, O[2 * b][1].x, O[2 * b][2].x, O[2 * b][0].y));
                        ptx::sts128(dst + 16, make_float4(O[2 * b][1].y, O[2 * b][2].y, O[2 * b + 1][0].x, O[2 * b + 1][1].x));
                        ptx::sts128(dst + 32, make_float4(O[2 * b + 1][2].x, O[2 * b + 1][0].y, O[2 * b + 1][1].y, O[2 * b + 1][2].y));
                    }
                }
                // fingertips (5 of the 778 vertices): the lane reads its own values back, before the tile is handed over
                if (tipmask != 0 && live) {
#pragma unroll
                    for (int qd = 0; qd < 5; ++qd) {
                        if (tipmask & (1u << qd)) {
                            const uint32_t f0 = (uint32_t)(cc.tip_ids[qd] % MF_VPT) * 12u;   // byte offset inside the thread's columns
                            const float pv0 = __uint_as_float(ptx::lds32(srow + f0)), pv1 = __uint_as_float(ptx::lds32(srow + f0 + 4)),
                                        pv2 = __uint_as_float(ptx::lds32(srow + f0 + 8));
                            const int k = cc.tip_slot[qd];
                            float* jo = a.joints + ((size_t)hand * 21 + k) * 3;
                            jo[0] = pv0; jo[1] = pv1; jo[2] = pv2;
                            if (a.kp2d != nullptr) {
                                const float X = pv0 + ct0, Y = pv1 + ct1, Z = pv2 + ct2;
                                float* ko = a.kp2d + ((size_t)hand * 21 + k) * 2;
                                ko[0] = a.focal_over_size * (X / Z);
                                ko[1] = a.focal_over_size * (Y / Z);
                            }
                        }
                    }
                }
                __syncwarp();
                if (!active) continue;
                const int col0 = g * MF_GCOLS;                             // first float of the group inside a vertex row
                float* gbase = a.verts + (size_t)row0 * (MANO_V * 3) + col0;
                if (a.dbg & 1) {
                } else if (row0 + 32 <= a.B && col0 + MF_GCOLS <= MANO_V * 3) {
#pragma unroll
                    for (int t = 0; t < HALF; ++t) {
                        const int r = rd_row[t % 3] + (t / 3) * RS, pos = rd_pos[t % 3];
                        *reinterpret_cast<float2*>(gbase + (size_t)r * (MANO_V * 3) + 2 * pos) =
                            ptx::lds64(sbase + (uint32_t)r * MF_PITCH + 8u * (uint32_t)pos);
                    }
                } else {                                                    // ragged last rows / last vertex group
#pragma unroll
                    for (int t = 0; t < HALF; ++t) {
                        const int r = rd_row[t % 3] + (t / 3) * RS, pos = rd_pos[t % 3];
                        if (row0 + r < a.B && col0 + 2 * pos + 2 <= MANO_V * 3)
                            *reinterpret_cast<float2*>(gbase + (size_t)r * (MANO_V * 3) + 2 * pos) =
                                ptx::lds64(sbase + (uint32_t)r * MF_PITCH + 8u * (uint32_t)pos);
                    }
                }
            }
            if (ck) ck[5] = clock64();
        }
    }

    // ===================== teardown =====================
    __syncwarp();
    ptx::tc_fence_before();
    ptx::cluster_sync_all();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<2>(tmem_base, MF_TMEM_COLS);
    }
}

}  // namespace hm
