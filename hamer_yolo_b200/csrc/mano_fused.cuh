// MANO layer, fused: blend-shape GEMM on tcgen05 with linear blend skinning as its epilogue.
//
// Replaces the same reference code as mano_lbs.cuh (MANO.forward, hamer/hamer/models/mano_wrapper.py:32-44
// -> smplx 0.1.28 lbs(); perspective_projection, hamer/hamer/utils/geometry.py:72-109) but never
// materialises v_posed in global memory: the only HBM traffic of the kernel is the coefficient /
// transform tiles in (1,088 B per hand) and the vertices out (9,336 B per hand).
//
//   work unit  = 256 hands (a CTA pair, UMMA M = 256) x a range of the 25 coordinate tiles (N = 96 =
//                32 vertices).  The fp16 coefficient tile (128 x 160 per CTA) and the 16 skinning
//                transforms of the CTA's 128 hands stay resident in shared memory for the whole unit;
//                the constant basis [2400 x 320] ([hi | lo] fp16 split, L2-resident) streams through a
//                5-stage TMA ring, 48 rows per CTA per stage.
//   MMA        : D[256 x 96] = coef[256 x 160] x (B_hi + B_lo)^T as 20 k-steps (the A descriptor of
//                k-step t is the coefficient slice t mod 10, so the [a | a] duplication costs no
//                memory), fp32 accumulators double-buffered in TMEM.
//   epilogue   : 8 warps, thread = hand (TMEM lane), 16 vertices per thread per tile.  v_posed =
//                accumulator + template; out_v = sum_u w_vu (R_u p_v + t_u) over the union of joints of
//                the 16-vertex group - the weights are warp-uniform (every lane works on the same
//                vertices), the transforms are per-lane 16-byte shared-memory loads (48-byte lane
//                stride: conflict free), the arithmetic is packed FFMA2 over vertex pairs.  The tile is
//                transposed through a padded shared-memory buffer so that each warp store writes 256
//                contiguous bytes (the vertex rows have a 9,336-byte pitch, not a multiple of 16 B, which
//                rules out a TMA store).  Fingertip joints and their projection are written by the
//                thread that owns the vertex.
#pragma once
#include "mano_lbs.cuh"
#include "ptx.cuh"

namespace hm {

constexpr int MF_BM = 128;                 // hands per CTA
constexpr int MF_BN = 96;                  // coordinates per tile (32 vertices)
constexpr int MF_NT = 25;                  // 25 x 96 = 2400 >= 2334
constexpr int MF_GV = 16;                  // vertices per thread per tile
constexpr int MF_GCOLS = MF_GV * 3;        // 48 accumulator columns per thread per tile
constexpr int MF_GROUPS = MF_NT * 2;       // 16-vertex groups
constexpr int MF_KSTEPS = MANO_KB / 16;    // 10 UMMA k-steps per basis half
constexpr int MF_BSTAGES = 5;
constexpr int MF_EPI_WARPS = 8;
constexpr int MF_THREADS = 128 + 32 * MF_EPI_WARPS;   // warpgroup 0: producer, MMA issuer (+2 idle warps); 1-2: epilogue
constexpr uint32_t MF_A_BYTES = 3 * MF_BM * 128;              // three 64-column SW128 blocks (cols >= 160 zero)
constexpr uint32_t MF_B_STAGE = (MF_BN / 2) * 128;            // 48 basis rows x 64 columns
constexpr uint32_t MF_AU_BYTES = MANO_J * MF_BM * 48;         // [joint][hand][12] fp32
constexpr int MF_PITCH = 50;                                  // staging row pitch (floats): conflict-free STS.64
constexpr uint32_t MF_STAGING = MF_EPI_WARPS * 32 * MF_PITCH * 4;
constexpr uint32_t MF_OFF_B = MF_A_BYTES;
constexpr uint32_t MF_OFF_AU = MF_OFF_B + MF_BSTAGES * MF_B_STAGE;
constexpr uint32_t MF_OFF_STAGING = MF_OFF_AU + MF_AU_BYTES;
constexpr uint32_t MF_OFF_BARS = MF_OFF_STAGING + MF_STAGING;
constexpr uint32_t MF_SMEM_BYTES = 1024 + MF_OFF_BARS + 256;
constexpr uint32_t MF_TMEM_COLS = 256;
// per 16-vertex group table (float4 units): [0,12) template in accumulator column order, [12,16) union joint
// ids (int), [16] = {union size, fingertip mask}, [20,24) vertex-pair masks per union slot (int),
// [32,96) weights [16 slots][16 vertices]
constexpr int MF_TAB_F4 = 96;                        // 2 x 96 accumulator columns
static_assert(MF_SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct ManoFusedArgs {
    const float* Au;        // [ceil(B/128)][16][128][12] skinning transforms (written by the prep kernel)
    const float* cam_t;     // [B,3] or null
    float* verts;           // [B,778,3]
    float* joints;          // [B,21,3] (fingertip slots only; the chain joints come from the prep kernel)
    float* kp2d;            // [B,21,2] or null
    float focal_over_size;
    int B;
    int n_split;            // work units per 256-hand tile (each takes a contiguous range of coordinate tiles)
};

__global__ void __launch_bounds__(MF_THREADS, 1)
mano_fused_kernel(const __grid_constant__ CUtensorMap tm_coef, const __grid_constant__ CUtensorMap tm_basis,
                  const ManoDevice m, const ManoFusedArgs a) {
    extern __shared__ uint8_t mf_smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(mf_smem_raw);
    uint8_t* smem = mf_smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* sA = smem;
    uint8_t* sB = smem + MF_OFF_B;
    float* sAu = reinterpret_cast<float*>(smem + MF_OFF_AU);
    float* sStage = reinterpret_cast<float*>(smem + MF_OFF_STAGING);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MF_OFF_BARS);
    uint64_t* full_b = bars;                       // [5] basis stage landed (leader's)
    uint64_t* empty_b = bars + MF_BSTAGES;         // [5] MMAs that read the stage retired (both CTAs)
    uint64_t* tmem_full = bars + 2 * MF_BSTAGES;   // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2] (leader's)
    uint64_t* a_full = tmem_empty + 2;             // coefficient tiles of both CTAs landed (leader's)
    uint64_t* a_empty = a_full + 1;                // the unit's last MMA retired (both CTAs)
    uint64_t* au_full = a_empty + 1;               // this CTA's transforms landed
    uint64_t* au_empty = au_full + 1;              // this CTA's epilogue warps are done with them
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(au_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const bool is_leader = rank == 0;
    const int worker = blockIdx.x >> 1, num_workers = gridDim.x >> 1;
    const int pair_tiles = (a.B + 2 * MF_BM - 1) / (2 * MF_BM);
    const int units = pair_tiles * a.n_split;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tm_coef);
        ptx::prefetch_tmap(&tm_basis);
        for (int s = 0; s < MF_BSTAGES; ++s) { ptx::mbar_init(&full_b[s], 1); ptx::mbar_init(&empty_b[s], 1); }
        for (int s = 0; s < 2; ++s) { ptx::mbar_init(&tmem_full[s], 1); ptx::mbar_init(&tmem_empty[s], MF_EPI_WARPS * 2); }
        ptx::mbar_init(a_full, 1);
        ptx::mbar_init(a_empty, 1);
        ptx::mbar_init(au_full, 1);
        ptx::mbar_init(au_empty, MF_EPI_WARPS);
        ptx::fence_mbar_init();
    }
    if (warp == 1) ptx::tmem_alloc<2>(tmem_slot, MF_TMEM_COLS);
    ptx::tc_fence_before();
    ptx::cluster_sync_all();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Register re-balancing between warpgroups: the single-lane roles give registers back, the
    // skinning warps (96 live fp32 pairs of v_posed / output per thread) take them.
    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
        // ===================== producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, it = 0;
            for (int unit = worker; unit < units; unit += num_workers, ++it) {
                const int pt = unit / a.n_split, q = unit - pt * a.n_split;
                const int nt0 = q * MF_NT / a.n_split, nt1 = (q + 1) * MF_NT / a.n_split;
                const int tile128 = pt * 2 + (int)rank;
                ptx::mbar_wait(a_empty, (it & 1u) ^ 1u);
                if (is_leader) ptx::mbar_arrive_expect_tx(a_full, 2 * MF_A_BYTES);
#pragma unroll
                for (int kb = 0; kb < 3; ++kb)
                    ptx::tma_load_2d_2sm(sA + kb * (MF_BM * 128), &tm_coef, a_full, kb * 64, tile128 * MF_BM);
                for (int nt = nt0; nt < nt1; ++nt) {
                    if (nt == nt0 + 1 || nt1 - nt0 == 1) {
                        // the transforms are only needed by the epilogue: loading them after the first
                        // basis tile lets the MMAs of this unit start while the previous unit drains
                        ptx::mbar_wait(au_empty, (it & 1u) ^ 1u);
                        ptx::mbar_arrive_expect_tx(au_full, MF_AU_BYTES);
                        ptx::bulk_load_1d(sAu, a.Au + (size_t)tile128 * (MF_AU_BYTES / 4), MF_AU_BYTES, au_full);
                    }
                    for (int kb = 0; kb < MF_BSTAGES; ++kb) {      // 320 / 64 = 5 basis blocks per tile
                        ptx::mbar_wait(&empty_b[stage], phase ^ 1u);
                        if (is_leader) ptx::mbar_arrive_expect_tx(&full_b[stage], 2 * MF_B_STAGE);
                        ptx::tma_load_2d_2sm(sB + stage * MF_B_STAGE, &tm_basis, &full_b[stage], kb * 64,
                                             nt * MF_BN + (int)rank * (MF_BN / 2));
                        if (++stage == MF_BSTAGES) { stage = 0; phase ^= 1u; }
                    }
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
                const int pt = unit / a.n_split, q = unit - pt * a.n_split;
                const int nt0 = q * MF_NT / a.n_split, nt1 = (q + 1) * MF_NT / a.n_split;
                ptx::mbar_wait(a_full, it & 1u);
                ptx::tc_fence_after();
                for (int nt = nt0; nt < nt1; ++nt) {
                    ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
                    ptx::tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * MF_BN);
                    for (int kb = 0; kb < MF_BSTAGES; ++kb) {
                        ptx::mbar_wait(&full_b[stage], phase);
                        ptx::tc_fence_after();
                        const uint64_t b_desc = ptx::make_smem_desc_k_sw128(ptx::smem_u32(sB + stage * MF_B_STAGE));
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int t = kb * 4 + k;                 // k-step over [hi | lo]
                            const int s = t >= MF_KSTEPS ? t - MF_KSTEPS : t;   // coefficient slice
                            const uint64_t a_desc =
                                ptx::make_smem_desc_k_sw128(ptx::smem_u32(sA + (s >> 2) * (MF_BM * 128))) +
                                (uint64_t)((s & 3) * 2);
                            ptx::mma_f16<2>(d_tmem, a_desc, b_desc + (uint64_t)(k * 2), idesc, (uint32_t)(t != 0));
                        }
                        ptx::mma_commit_2sm(&empty_b[stage], 0x3);
                        if (++stage == MF_BSTAGES) { stage = 0; phase ^= 1u; }
                    }
                    ptx::mma_commit_2sm(&tmem_full[acc], 0x3);
                    acc ^= 1;
                    if (acc == 0) acc_phase ^= 1u;
                }
                ptx::mma_commit_2sm(a_empty, 0x3);
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        // ===================== epilogue: skinning, thread = hand =====================
        const int ew = warp - 4, quad = warp & 3, part = ew >> 2;
        const int hl = quad * 32 + lane;                      // hand inside this CTA's tile = TMEM lane
        float* st = sStage + ew * (32 * MF_PITCH);            // per-warp buffer: group table, then the output tile
        const float4* st4 = reinterpret_cast<const float4*>(st);
        int acc = 0;
        uint32_t acc_phase = 0, it = 0;
        int soff[3], goff[3], rr[3], cc[3];                    // store pattern of this lane (see below)
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {
            const int e = kk * 32 + lane;
            rr[kk] = e / 24; cc[kk] = 2 * (e % 24);
            soff[kk] = rr[kk] * MF_PITCH + cc[kk];
            goff[kk] = rr[kk] * (MANO_V * 3) + cc[kk];
        }
        float4 tb0, tb1, tb2;                                  // this lane's slice of the next group's table
        auto fetch_table = [&](int g) {
            const float4* t = m.fz_tab + (size_t)g * MF_TAB_F4;
            tb0 = __ldg(t + lane); tb1 = __ldg(t + 32 + lane); tb2 = __ldg(t + 64 + lane);
        };
        auto store_table = [&]() {
            float4* d = reinterpret_cast<float4*>(st);
            d[lane] = tb0; d[32 + lane] = tb1; d[64 + lane] = tb2;
        };
        for (int unit = worker; unit < units; unit += num_workers, ++it) {
            const int pt = unit / a.n_split, q = unit - pt * a.n_split;
            const int nt0 = q * MF_NT / a.n_split, nt1 = (q + 1) * MF_NT / a.n_split;
            const int row0 = (pt * 2 + (int)rank) * MF_BM + quad * 32;   // first hand of this warp
            const int hand = row0 + lane;
            fetch_table(nt0 * 2 + part);
            __syncwarp();                                       // previous unit's reads of `st` are done
            store_table();
            __syncwarp();
            ptx::mbar_wait(au_full, it & 1u);
            for (int nt = nt0; nt < nt1; ++nt) {
                const int g = nt * 2 + part;                   // 16-vertex group
                const bool active = g * MF_GCOLS < MANO_V * 3; // the last group is padding (warp-uniform)
                if (nt + 1 < nt1) fetch_table(g + 2);          // L2 latency hidden behind this tile's arithmetic
                ptx::mbar_wait(&tmem_full[acc], acc_phase);
                ptx::tc_fence_after();
                const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * MF_BN + part * MF_GCOLS);
                uint32_t ra[32], rb[16];
                ptx::tmem_ld_32x32(t_addr, ra);
                ptx::tmem_ld_32x16(t_addr + 32, rb);
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive_cluster(&tmem_empty[acc], 0);   // accumulator is in registers
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;

                if (active) {
                    // v_posed = accumulator + template.  The basis rows are ordered (x0 x1 y0 y1 z0 z1) per vertex
                    // pair, so the accumulator registers already are the packed pairs P[i][c] = (p_c(v_2i), p_c(v_2i+1)).
                    float2 P[8][3], O[8][3];
#pragma unroll
                    for (int j4 = 0; j4 < 12; ++j4) {
                        const float4 t4 = st4[j4];
                        const uint32_t* r = j4 < 8 ? &ra[j4 * 4] : &rb[(j4 - 8) * 4];
                        const float2 lo = __fadd2_rn(make_float2(__uint_as_float(r[0]), __uint_as_float(r[1])), make_float2(t4.x, t4.y));
                        const float2 hi = __fadd2_rn(make_float2(__uint_as_float(r[2]), __uint_as_float(r[3])), make_float2(t4.z, t4.w));
                        P[(j4 * 2) / 3][(j4 * 2) % 3] = lo;
                        P[(j4 * 2 + 1) / 3][(j4 * 2 + 1) % 3] = hi;
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int c = 0; c < 3; ++c) O[i][c] = make_float2(0.f, 0.f);
                    const int* sti = reinterpret_cast<const int*>(st);
                    const int nu = sti[64], tipmask = sti[65];
                    for (int u = 0; u < nu; ++u) {
                        const int j = sti[48 + u];
                        const uint32_t pm = (uint32_t)sti[80 + u];   // vertex pairs with a non-zero weight for this joint
                        const float4* Aj = reinterpret_cast<const float4*>(sAu + (size_t)(j * MF_BM + hl) * 12);
                        const float4 Rc[3] = {Aj[0], Aj[1], Aj[2]};
                        const float4 W0 = st4[32 + u * 4 + 0], W1 = st4[32 + u * 4 + 1];
                        const float4 W2 = st4[32 + u * 4 + 2], W3 = st4[32 + u * 4 + 3];
                        const float wv[16] = {W0.x, W0.y, W0.z, W0.w, W1.x, W1.y, W1.z, W1.w,
                                              W2.x, W2.y, W2.z, W2.w, W3.x, W3.y, W3.z, W3.w};
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            if (pm & (1u << i)) {                // warp-uniform: the weights do not depend on the hand
                                const float2 w2 = make_float2(wv[2 * i], wv[2 * i + 1]);
#pragma unroll
                                for (int c = 0; c < 3; ++c) {
                                    float2 t = __ffma2_rn(make_float2(Rc[c].z, Rc[c].z), P[i][2], make_float2(Rc[c].w, Rc[c].w));
                                    t = __ffma2_rn(make_float2(Rc[c].y, Rc[c].y), P[i][1], t);
                                    t = __ffma2_rn(make_float2(Rc[c].x, Rc[c].x), P[i][0], t);
                                    O[i][c] = __ffma2_rn(w2, t, O[i][c]);
                                }
                            }
                        }
                    }
                    // ---- transpose through shared memory: lane = hand  ->  lane = 8 contiguous bytes ----
                    __syncwarp();                               // every lane is done with the group table
                    {
                        float2* row = reinterpret_cast<float2*>(st + lane * MF_PITCH);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            row[3 * i + 0] = make_float2(O[i][0].x, O[i][1].x);
                            row[3 * i + 1] = make_float2(O[i][2].x, O[i][0].y);
                            row[3 * i + 2] = make_float2(O[i][1].y, O[i][2].y);
                        }
                    }
                    __syncwarp();
                    // 32 rows x 24 float2: element e = 32 k + lane -> row e / 24, column pair e % 24; three
                    // consecutive k advance exactly 4 rows, so all offsets are per-lane constants + immediates
                    float* gbase = a.verts + (size_t)row0 * (MANO_V * 3) + g * MF_GCOLS;
                    if (row0 + 32 <= a.B && g < MF_GROUPS - 2) {
#pragma unroll
                        for (int t = 0; t < 8; ++t)
#pragma unroll
                            for (int kk = 0; kk < 3; ++kk)
                                *reinterpret_cast<float2*>(gbase + goff[kk] + t * 4 * (MANO_V * 3)) =
                                    *reinterpret_cast<const float2*>(st + soff[kk] + t * 4 * MF_PITCH);
                    } else {                                    // ragged last rows / last vertex group
                        const int rows = a.B - row0, cols = MANO_V * 3 - g * MF_GCOLS;
#pragma unroll
                        for (int t = 0; t < 8; ++t)
#pragma unroll
                            for (int kk = 0; kk < 3; ++kk)
                                if (rr[kk] + 4 * t < rows && cc[kk] < cols)
                                    *reinterpret_cast<float2*>(gbase + goff[kk] + t * 4 * (MANO_V * 3)) =
                                        *reinterpret_cast<const float2*>(st + soff[kk] + t * 4 * MF_PITCH);
                    }
                    // ---- fingertips: the lane reads its own row back ----
                    if (tipmask != 0 && hand < a.B) {
#pragma unroll
                        for (int qd = 0; qd < 5; ++qd) {
                            if (tipmask & (1 << qd)) {
                                const float* pv = st + lane * MF_PITCH + (m.tip_ids[qd] & 15) * 3;
                                const float px = pv[0], py = pv[1], pz = pv[2];
                                const int k = m.tip_slot[qd];
                                float* jo = a.joints + ((size_t)hand * 21 + k) * 3;
                                jo[0] = px; jo[1] = py; jo[2] = pz;
                                if (a.kp2d != nullptr) {
                                    const float* ct = a.cam_t + (size_t)hand * 3;
                                    const float X = px + ct[0], Y = py + ct[1], Z = pz + ct[2];
                                    float* ko = a.kp2d + ((size_t)hand * 21 + k) * 2;
                                    ko[0] = a.focal_over_size * (X / Z);
                                    ko[1] = a.focal_over_size * (Y / Z);
                                }
                            }
                        }
                    }
                }
                if (nt + 1 < nt1) {                             // next group's table replaces the output tile
                    __syncwarp();
                    store_table();
                    __syncwarp();
                }
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(au_empty);
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
