// MANO layer, fused: blend-shape GEMM on tcgen05 with linear blend skinning as its epilogue.
//
// Replaces the same reference code as mano_lbs.cuh (MANO.forward, hamer/hamer/models/mano_wrapper.py:32-44
// -> smplx 0.1.28 lbs(); perspective_projection, hamer/hamer/utils/geometry.py:72-109) but never
// materialises v_posed in global memory: the only HBM traffic of the kernel is the coefficient /
// transform rows in (1,088 B per hand, written by mano_prep_kernel<true>) and the vertices out (9,336 B).
//
//   work unit  = 256 hands (a CTA pair, UMMA M = 256) x a range of the 25 coordinate tiles (N = 96 =
//                32 vertices).
//   tensor     : 512 columns per CTA, lane = hand:  [0,192) two fp32 accumulators, [192,384) the 16 skinning
//   memory       transforms of the hand (12 columns per joint), [384,464) its 160 fp16 blend coefficients.
//                Transforms and coefficients are written once per unit (global -> registers -> tcgen05.st)
//                and read by tcgen05.ld / as the MMA's A operand: they cost no shared-memory bandwidth,
//                which is what bounds this kernel.
//   MMA        : D[256 x 96] = coef[256 x 160] x basis^T as 10 k-steps, A from tensor memory, B = the constant
//                fp16 basis [2400 x 160] (L2-resident) streamed by TMA; one ring stage is one whole coordinate
//                tile (three 48-row x 64-column boxes per CTA), so the issuing thread pays one barrier wait and
//                one commit per 10 MMAs.  K layout: [shape hi (10) | pose (135) | shape lo (10) | 0 (5)] with the
//                betas repeated under "shape lo": the shape basis is exact to 2^-22, the pose basis is rounded to
//                fp16 (<= 0.004 mm on MANO, tests/tools note in DESIGN.md) - half the MMAs, basis traffic and
//                shared-memory reads of a full [hi | lo] split.
//   epilogue   : 16 warps, thread = hand (TMEM lane), 8 vertices per thread per tile.  v_posed =
//                accumulator + template; out_v = sum_u w_vu (R_u p_v + t_u) over the union of joints of
//                the 8-vertex group - weights and joint ids are warp-uniform (broadcast reads of a 768-byte
//                table), vertex pairs without a weight for the joint are skipped, arithmetic is packed
//                FFMA2 over vertex pairs (the basis rows are ordered x0 x1 y0 y1 z0 z1 so the accumulator
//                registers already are the pairs).  The tile is transposed through shared memory - dense
//                96-byte rows, each lane's pairs rotated by 3 (lane / 4 mod 4) positions so that both the
//                row-wise writes and the linear reads are bank-conflict free - and stored with 8-byte
//                stores, 256 contiguous bytes per warp instruction (the vertex rows have a 9,336-byte
//                pitch, not a multiple of 16 B, which rules out a TMA store).  Fingertip joints and their
//                projection are written by the thread that owns the vertex.
#pragma once
#include "mano_lbs.cuh"
#include "ptx.cuh"

namespace hm {

constexpr int MF_BM = 128;                 // hands per CTA
constexpr int MF_BN = 96;                  // coordinates per tile (32 vertices)
constexpr int MF_NT = 25;                  // 25 x 96 = 2400 >= 2334
constexpr int MF_GV = 8;                   // vertices per thread per tile
constexpr int MF_GCOLS = MF_GV * 3;        // 24 accumulator columns per thread per tile
constexpr int MF_PARTS = MF_BN / MF_GCOLS; // 4 column slices per tile
constexpr int MF_GROUPS = MF_NT * MF_PARTS;// 8-vertex groups (the last two are padding)
constexpr int MF_KSTEPS = MANO_KB / 16;    // 10 UMMA k-steps per basis half
constexpr int MF_KBLOCKS = (MANO_KB + 63) / 64; // 3 basis boxes of 64 columns per tile (the last one half used)
constexpr int MF_BSTAGES = 4;              // one stage = one whole coordinate tile: with one k-block per stage the issuing
                                           // thread, not the tensor pipe, bounded the kernel (450 cycles per 4 MMAs of 41)
constexpr int MF_EPI_WARPS = 4 * MF_PARTS; // 16: four TMEM lane quadrants x four column slices
constexpr int MF_THREADS = 128 + 32 * MF_EPI_WARPS;   // warpgroup 0: producer, MMA issuer (+2 idle warps); 1-4: epilogue
constexpr uint32_t MF_B_BLOCK = (MF_BN / 2) * 128;            // 48 basis rows x 64 columns
constexpr uint32_t MF_B_STAGE = MF_KBLOCKS * MF_B_BLOCK;      // 18 KB per CTA per coordinate tile
constexpr uint32_t MF_WARP_STAGE = 32 * MF_GCOLS * 4;         // 3,072 B per epilogue warp: 32 dense rows of 24 floats
constexpr uint32_t MF_STAGING = MF_EPI_WARPS * MF_WARP_STAGE;
constexpr uint32_t MF_OFF_STAGING = MF_BSTAGES * MF_B_STAGE;
constexpr uint32_t MF_TAB_BYTES = 98 * 48 * 16;                // group tables of the 98 live 8-vertex groups, resident for the whole kernel
constexpr uint32_t MF_OFF_TAB = MF_OFF_STAGING + MF_STAGING;
constexpr uint32_t MF_OFF_BARS = MF_OFF_TAB + MF_TAB_BYTES;
constexpr uint32_t MF_SMEM_BYTES = 1024 + MF_OFF_BARS + 256;
constexpr uint32_t MF_TMEM_COLS = 512;
constexpr uint32_t MF_AU_COL = 2 * MF_BN;                     // transforms: joint j at MF_AU_COL + 12 j
constexpr uint32_t MF_COEF_COL = MF_AU_COL + MANO_J * 12;     // coefficients: k-step s at MF_COEF_COL + 8 s
// per 8-vertex group table, 48 float4 = 768 B:
//   float4 [0,6)   template in accumulator column order (x0 x1 y0 y1 z0 z1 per vertex pair)
//   int    [24]    union size, [25] fingertip mask
//   int    [32,48) per union slot: (12 x joint) | (mask of the vertex pairs with a non-zero weight) << 16
//   float  [64,192) weights [16 slots][8 vertices]
constexpr int MF_TAB_F4 = 48;
static_assert(MF_SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(MF_COEF_COL + MANO_KB / 2 <= MF_TMEM_COLS, "tensor memory budget");
// setmaxnreg moves registers inside the CTA's launch allocation only (640 threads x 96 registers): the
// re-balanced total must fit or the .inc blocks forever
static_assert(128 * 40 + 32 * MF_EPI_WARPS * 104 <= MF_THREADS * 96, "register pool");

struct ManoFusedArgs {
    const __half* coef;     // [ceil(B/128)*128][160] fp16 blend coefficients (written by the prep kernel)
    const float* Au;        // [ceil(B/128)][16][128][12] skinning transforms (written by the prep kernel)
    const float* cam_t;     // [B,3] or null
    float* verts;           // [B,778,3]
    float* joints;          // [B,21,3] (fingertip slots only; the chain joints come from the prep kernel)
    float* kp2d;            // [B,21,2] or null
    float focal_over_size;
    int B;
    int n_split;            // work units per 256-hand tile (each takes a contiguous range of coordinate tiles)
    int dbg;                // diagnostics (HM_MANO_DBG): 1 = no vertex stores, 2 = no skinning loop, 4 = accumulator read only
};

__global__ void __launch_bounds__(MF_THREADS, 1)
mano_fused_kernel(const __grid_constant__ CUtensorMap tm_basis, const ManoDevice m, const ManoFusedArgs a) {
    extern __shared__ uint8_t mf_smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(mf_smem_raw);
    uint8_t* smem = mf_smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* sB = smem;
    uint8_t* sStage = smem + MF_OFF_STAGING;
    uint8_t* sTab = smem + MF_OFF_TAB;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MF_OFF_BARS);
    uint64_t* full_b = bars;                       // [MF_BSTAGES] basis tile landed (leader's)
    uint64_t* empty_b = bars + MF_BSTAGES;         // [MF_BSTAGES] MMAs that read the tile retired (both CTAs)
    uint64_t* tmem_full = bars + 2 * MF_BSTAGES;   // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2] (leader's)
    uint64_t* coef_full = tmem_empty + 2;          // both CTAs' coefficients are in tensor memory (leader's)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(coef_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const bool is_leader = rank == 0;
    const int worker = blockIdx.x >> 1, num_workers = gridDim.x >> 1;
    const int pair_tiles = (a.B + 2 * MF_BM - 1) / (2 * MF_BM);
    const int units = pair_tiles * a.n_split;
    ptx::griddep_launch_dependents();   // PDL: the next kernel may start its launch / set-up

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tm_basis);
        for (int s = 0; s < MF_BSTAGES; ++s) { ptx::mbar_init(&full_b[s], 1); ptx::mbar_init(&empty_b[s], 1); }
        for (int s = 0; s < 2; ++s) { ptx::mbar_init(&tmem_full[s], 1); ptx::mbar_init(&tmem_empty[s], MF_EPI_WARPS * 2); }
        ptx::mbar_init(coef_full, MF_EPI_WARPS * 2);
        ptx::fence_mbar_init();
    }
    if (warp == 1) ptx::tmem_alloc<2>(tmem_slot, MF_TMEM_COLS);
    ptx::tc_fence_before();
    ptx::cluster_sync_all();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::griddep_wait();                // coefficients / transforms come from the prep kernel

    // Register re-balancing between warpgroups: the single-lane roles give registers back, the
    // skinning warps take them.
    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == 0) {
            // ===================== producer: basis tiles =====================
            if (lane == 0) {
                int stage = 0;
                uint32_t phase = 0;
                for (int unit = worker; unit < units; unit += num_workers) {
                    const int pt = unit / a.n_split, q = unit - pt * a.n_split;
                    const int nt0 = q * MF_NT / a.n_split, nt1 = (q + 1) * MF_NT / a.n_split;
                    for (int nt = nt0; nt < nt1; ++nt) {
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
                    const int pt = unit / a.n_split, q = unit - pt * a.n_split;
                    const int nt0 = q * MF_NT / a.n_split, nt1 = (q + 1) * MF_NT / a.n_split;
                    ptx::mbar_wait(coef_full, it & 1u);
                    ptx::tc_fence_after();
                    for (int nt = nt0; nt < nt1; ++nt) {
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
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        // ===================== epilogue: skinning, thread = hand, 8 vertices per tile =====================
        const int ew = warp - 4, quad = warp & 3, part = ew >> 2;
        const int hl = quad * 32 + lane;                                   // hand inside this CTA's tile = TMEM lane
        const uint32_t tmem_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
        const uint32_t sbase = ptx::smem_u32(sStage) + (uint32_t)ew * MF_WARP_STAGE;   // this warp's output tile
        const uint32_t tab0 = ptx::smem_u32(sTab);                                      // + 768 * group
        // staging layout: row = lane (96 dense bytes), pair k of the row at position (k + rot) mod 12
        const int rot = 3 * ((lane >> 2) & 3);
        const uint32_t wbase = sbase + (uint32_t)lane * (MF_GCOLS * 4) + 8 * rot;
        const int wrap = 12 - rot;                                         // pairs k >= wrap sit 96 bytes lower
        // read-out: element e = 32 k + lane of the 32 x 12 pair matrix (linear in shared memory) is row e / 12,
        // position e % 12, i.e. logical pair (position - rot(row)) mod 12; three consecutive k advance 8 rows,
        // so the global offsets are per-lane constants (two variants: rot(row) alternates with the 8-row step)
        int goff[3][2];
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {
            const int e = kk * 32 + lane, r = e / 12, pos = e % 12;
#pragma unroll
            for (int par = 0; par < 2; ++par) {
                const int rr = 3 * (((r >> 2) + 2 * par) & 3);
                goff[kk][par] = r * (MANO_V * 3) + 2 * ((pos - rr + 12) % 12);
            }
        }
        // the constant group tables (template, joint unions, pair masks, weights; 74 KB) are loaded once: a per-tile
        // global load sat on the critical path of every tile while the L2 absorbs the kernel's 2 TB/s of stores
        for (int i = (int)threadIdx.x - 128; i < (int)(MF_TAB_BYTES / 16); i += 32 * MF_EPI_WARPS)
            ptx::sts128(tab0 + 16 * i, __ldg(m.fz_tab + i));
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int unit = worker; unit < units; unit += num_workers) {
            const int pt = unit / a.n_split, q = unit - pt * a.n_split;
            const int nt0 = q * MF_NT / a.n_split, nt1 = (q + 1) * MF_NT / a.n_split;
            const int tile128 = pt * 2 + (int)rank;
            const int row0 = tile128 * MF_BM + quad * 32;                  // first hand of this warp
            const int hand = row0 + lane;
            // ---- this unit's per-hand operands: global -> registers -> tensor memory.  The four warps of a lane
            // quadrant split the work: joints 4 part .. 4 part + 3 and coefficient columns 20 part .. 20 part + 19.
            {
                uint32_t v[3][16], c16[16], c4[4];
                const float4* src = reinterpret_cast<const float4*>(
                    a.Au + ((size_t)tile128 * MANO_J + part * 4) * (MF_BM * 12) + (size_t)hl * 12);
#pragma unroll
                for (int k = 0; k < 12; ++k) {
                    const float4 f = __ldg(src + (k / 3) * (MF_BM * 3) + (k % 3));
                    v[k / 4][(k % 4) * 4 + 0] = __float_as_uint(f.x); v[k / 4][(k % 4) * 4 + 1] = __float_as_uint(f.y);
                    v[k / 4][(k % 4) * 4 + 2] = __float_as_uint(f.z); v[k / 4][(k % 4) * 4 + 3] = __float_as_uint(f.w);
                }
                const uint4* csrc = reinterpret_cast<const uint4*>(a.coef + (size_t)(tile128 * MF_BM + hl) * MANO_KB + part * 40);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint4 f = __ldg(csrc + k);
                    c16[k * 4 + 0] = f.x; c16[k * 4 + 1] = f.y; c16[k * 4 + 2] = f.z; c16[k * 4 + 3] = f.w;
                }
                {
                    const uint4 f = __ldg(csrc + 4);
                    c4[0] = f.x; c4[1] = f.y; c4[2] = f.z; c4[3] = f.w;
                }
                // every epilogue warp has consumed the previous unit's last accumulator (so its MMAs, the readers
                // of the coefficients, have retired) and is done with the previous transforms
                asm volatile("bar.sync 1, %0;" ::"n"(32 * MF_EPI_WARPS) : "memory");
                ptx::tmem_st_32x16(tmem_lane + MF_AU_COL + (uint32_t)(part * 48), v[0]);
                ptx::tmem_st_32x16(tmem_lane + MF_AU_COL + (uint32_t)(part * 48) + 16, v[1]);
                ptx::tmem_st_32x16(tmem_lane + MF_AU_COL + (uint32_t)(part * 48) + 32, v[2]);
                ptx::tmem_st_32x16(tmem_lane + MF_COEF_COL + (uint32_t)(part * 20), c16);
                ptx::tmem_st_32x4(tmem_lane + MF_COEF_COL + (uint32_t)(part * 20) + 16, c4);
                ptx::tmem_st_wait();
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(coef_full, 0);          // 32 warps of the pair -> the MMA issuer may start
            asm volatile("bar.sync 1, %0;" ::"n"(32 * MF_EPI_WARPS) : "memory");
            ptx::tc_fence_after();

            for (int nt = nt0; nt < nt1; ++nt) {
                const int g = nt * MF_PARTS + part;                        // 8-vertex group
                const bool active = g * MF_GCOLS < MANO_V * 3 && !(a.dbg & 4);   // the last two groups are padding (warp-uniform)
                const uint32_t tab = tab0 + (uint32_t)g * (MF_TAB_F4 * 16);  // this group's table (groups >= 98 are padding, never read)
                ptx::mbar_wait(&tmem_full[acc], acc_phase);
                ptx::tc_fence_after();
                const uint32_t t_addr = tmem_lane + (uint32_t)(acc * MF_BN + part * MF_GCOLS);
                uint32_t ra[16], rb[8];
                ptx::tmem_ld_32x16(t_addr, ra);
                ptx::tmem_ld_32x8(t_addr + 16, rb);
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive_cluster(&tmem_empty[acc], 0);   // accumulator is in registers
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;

                if (active) {
                    // v_posed = accumulator + template as packed pairs P[i][c] = (p_c(v_2i), p_c(v_2i+1))
                    float2 P[4][3], O[4][3];
#pragma unroll
                    for (int j4 = 0; j4 < 6; ++j4) {
                        const float4 t4 = ptx::lds128(tab + j4 * 16);
                        const uint32_t* r = j4 < 4 ? &ra[j4 * 4] : &rb[(j4 - 4) * 4];
                        P[(j4 * 2) / 3][(j4 * 2) % 3] =
                            __fadd2_rn(make_float2(__uint_as_float(r[0]), __uint_as_float(r[1])), make_float2(t4.x, t4.y));
                        P[(j4 * 2 + 1) / 3][(j4 * 2 + 1) % 3] =
                            __fadd2_rn(make_float2(__uint_as_float(r[2]), __uint_as_float(r[3])), make_float2(t4.z, t4.w));
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int c = 0; c < 3; ++c) O[i][c] = make_float2(0.f, 0.f);
                    const int nu = (a.dbg & 2) ? 0 : (int)ptx::lds32(tab + 96);
                    const uint32_t tipmask = ptx::lds32(tab + 100);
                    for (int u = 0; u < nu; ++u) {
                        const uint32_t jp = ptx::lds32(tab + 128 + u * 4);
                        const uint32_t pm = jp >> 16;                       // vertex pairs with a non-zero weight
                        uint32_t ta[8], tc[4];
                        ptx::tmem_ld_32x8(tmem_lane + MF_AU_COL + (jp & 0xffffu), ta);
                        ptx::tmem_ld_32x4(tmem_lane + MF_AU_COL + (jp & 0xffffu) + 8, tc);
                        const float4 W0 = ptx::lds128(tab + 256 + u * 32), W1 = ptx::lds128(tab + 272 + u * 32);
                        ptx::tmem_ld_wait();
                        const float4 Rc[3] = {
                            make_float4(__uint_as_float(ta[0]), __uint_as_float(ta[1]), __uint_as_float(ta[2]), __uint_as_float(ta[3])),
                            make_float4(__uint_as_float(ta[4]), __uint_as_float(ta[5]), __uint_as_float(ta[6]), __uint_as_float(ta[7])),
                            make_float4(__uint_as_float(tc[0]), __uint_as_float(tc[1]), __uint_as_float(tc[2]), __uint_as_float(tc[3]))};
                        const float2 w2[4] = {make_float2(W0.x, W0.y), make_float2(W0.z, W0.w),
                                              make_float2(W1.x, W1.y), make_float2(W1.z, W1.w)};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            if (pm & (1u << i)) {                           // warp-uniform: the weights do not depend on the hand
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
                    // ---- transpose through shared memory: lane = hand  ->  lane = 8 contiguous bytes ----
                    __syncwarp();                                           // the previous tile's read-out of the staging rows is done
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        // natural order of the row: pair 3i = (x0 y0), 3i+1 = (z0 x1), 3i+2 = (y1 z1)
                        ptx::sts64(wbase + 8 * (3 * i + 0) - (3 * i + 0 >= wrap ? 96 : 0), make_float2(O[i][0].x, O[i][1].x));
                        ptx::sts64(wbase + 8 * (3 * i + 1) - (3 * i + 1 >= wrap ? 96 : 0), make_float2(O[i][2].x, O[i][0].y));
                        ptx::sts64(wbase + 8 * (3 * i + 2) - (3 * i + 2 >= wrap ? 96 : 0), make_float2(O[i][1].y, O[i][2].y));
                    }
                    __syncwarp();
                    float* gbase = a.verts + (size_t)row0 * (MANO_V * 3) + g * MF_GCOLS;
                    if (a.dbg & 1) {
                    } else if (row0 + 32 <= a.B && (g + 1) * MF_GCOLS <= MANO_V * 3) {
#pragma unroll
                        for (int t = 0; t < 4; ++t)
#pragma unroll
                            for (int kk = 0; kk < 3; ++kk)
                                *reinterpret_cast<float2*>(gbase + goff[kk][t & 1] + t * 8 * (MANO_V * 3)) =
                                    ptx::lds64(sbase + 8 * lane + 256 * (3 * t + kk));
                    } else {                                                // ragged last rows / last vertex group
                        const int rows = a.B - row0, cols = MANO_V * 3 - g * MF_GCOLS;
#pragma unroll
                        for (int t = 0; t < 4; ++t)
#pragma unroll
                            for (int kk = 0; kk < 3; ++kk) {
                                const int r0 = (kk * 32 + lane) / 12;
                                const int col = goff[kk][t & 1] - r0 * (MANO_V * 3);
                                if (r0 + 8 * t < rows && col < cols)
                                    *reinterpret_cast<float2*>(gbase + goff[kk][t & 1] + t * 8 * (MANO_V * 3)) =
                                        ptx::lds64(sbase + 8 * lane + 256 * (3 * t + kk));
                            }
                    }
                    // ---- fingertips: the lane reads its own row back ----
                    if (tipmask != 0 && hand < a.B) {
#pragma unroll
                        for (int qd = 0; qd < 5; ++qd) {
                            if (tipmask & (1u << qd)) {
                                float pv[3];
#pragma unroll
                                for (int c = 0; c < 3; ++c) {
                                    const int f = (m.tip_ids[qd] & (MF_GV - 1)) * 3 + c;   // float index inside the row
                                    const int pos = ((f >> 1) + rot) % 12;
                                    pv[c] = __uint_as_float(ptx::lds32(sbase + (uint32_t)lane * (MF_GCOLS * 4) + 8 * pos + 4 * (f & 1)));
                                }
                                const int k = m.tip_slot[qd];
                                float* jo = a.joints + ((size_t)hand * 21 + k) * 3;
                                jo[0] = pv[0]; jo[1] = pv[1]; jo[2] = pv[2];
                                if (a.kp2d != nullptr) {
                                    const float* ct = a.cam_t + (size_t)hand * 3;
                                    const float X = pv[0] + ct[0], Y = pv[1] + ct[1], Z = pv[2] + ct[2];
                                    float* ko = a.kp2d + ((size_t)hand * 21 + k) * 2;
                                    ko[0] = a.focal_over_size * (X / Z);
                                    ko[1] = a.focal_over_size * (Y / Z);
                                }
                            }
                        }
                    }
                }
            }
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
