// MANO layer: blend shapes + joint regression + kinematic chain + linear blend skinning,
// fingertip append / OpenPose re-ordering and the crop-camera projection.
//
// Replaces MANO.forward (hamer/hamer/models/mano_wrapper.py:32-44) -> smplx 0.1.28
// MANOLayer.forward -> lbs() (third-party; algorithm in SURVEY 8a row a16) and
// perspective_projection (hamer/hamer/utils/geometry.py:72-109) as called from
// HAMER.forward_step (hamer/hamer/models/hamer.py:143-152).
//
// Three launches per chunk of hands (the chunk is sized so that the intermediate stays in L2):
//   1. mano_prep_kernel   (16 lanes per hand, lane = joint): blend coefficients [beta | vec(R_j - I)] as an fp16 GEMM
//      operand, rest joints J = J_template + J_shapedirs * beta (joint regressor folded into the
//      basis at set-up), the 16-joint kinematic chain, the skinning transforms A_j and the posed joints.
//   2. the tcgen05 GEMM of gemm_tcgen05.cuh:  v_posed[hands, 2334] = coef[hands, 320] * basis^T + v_template
//      K = 2 x 160: the constant basis is split hi + lo in fp16 (exact to 2^-22), the coefficients are
//      fp16 (|error| <= 2^-12 |a| |b| per term: < 0.02 mm worst case), the template is added in fp32
//      by the epilogue.  This is the 0.73 MFLOP/hand that kept the fp32 SIMT version at 3 % of the HBM
//      roofline.
//   3. mano_skin_kernel   (thread = vertex, 16 hands per CTA): T_v = sum_j w_vj A_j over the warp's
//      union of joints (each A_j read is one shared-memory broadcast for the whole warp; absent joints
//      carry weight 0), v = T_v [v_posed; 1]; warps run independently (no per-hand barrier), each
//      reading / writing 384 contiguous bytes per hand; fingertips, OpenPose joint order, projection.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"

namespace hm {

constexpr int MANO_V = 778;
constexpr int MANO_J = 16;
constexpr int MANO_KB = 160;           // 10 shape + 135 pose coefficients, zero-padded to 160
constexpr int MANO_KTOT = 2 * MANO_KB; // [a | a] x [b_hi ; b_lo]
constexpr int MANO_NPAD = 2432;        // 2334 coordinates padded to 19 x 128 (GEMM N)
constexpr int MANO_VTHREADS = 800;     // 25 warps cover the 778 vertices
constexpr int MANO_WARPS = MANO_VTHREADS / 32;
constexpr int MANO_MAX_UNION = 16;
constexpr int SKIN_HANDS = 16;

struct ManoDevice {
    const __half* basis16;     // [2432][320] fp16: row = v*3 + c, cols [b_hi(160) | b_lo(160)]
    const __half* basis16f;    // [2432][160] fused kernel: rows in accumulator column order, cols [shape hi | pose | shape lo | 0]
    const float* template_n;   // [2432] v_template flattened (GEMM bias), zero padded
    const float* zero_row;     // [2432] zeros (GEMM "pos" operand)
    const float* j_template;   // [16][3]
    const float* j_shapedirs;  // [16][3][10]
    const int32_t* perm;       // [800] thread -> vertex id (natural order), -1 = idle lane
    const int32_t* warp_nu;    // [25] size of each warp's joint union
    const int32_t* warp_union; // [25][16] joint ids
    const float* lane_w;       // [800][16] skinning weight of the thread's vertex for each union slot
    // fused kernel (mano_fused.cuh): per vertex group, the union of joints, pair masks and weights per union slot
    const float4* fz_tab;      // [MF_LIVE_GROUPS][MF_TAB_STRIDE / 16]
    int32_t parents[16];
    int32_t tip_ids[5];
    int32_t joint_map[21];
    int32_t max_depth;         // depth of the kinematic tree (longest ancestor chain)
    int32_t chain_slot[16];    // output slot (0..20) of chain joint j, -1 = not in joint_map
    int32_t tip_slot[5];       // output slot of fingertip q
};

struct ManoPrepArgs {
    const float* betas;    // [B,10]
    const float* rotmats;  // [B,16,3,3]
    __half* coef;          // [B,320]
    float* A;              // [B,16,12] skinning transforms (3x4 row-major)
    float* jpos;           // [B,16,3] posed joints
    int B;
};

// 16 lanes per hand (lane = joint), 8 hands per 128-thread block.
__global__ void __launch_bounds__(128)
mano_prep_kernel(const ManoDevice m, const ManoPrepArgs a) {
    ptx::pdl_prologue();
    __shared__ float sJ[8][MANO_J][3];
    __shared__ float sG[8][MANO_J][12];
    const int slot = threadIdx.x >> 4, j = threadIdx.x & 15;
    const int h = blockIdx.x * 8 + slot;
    const bool live = h < a.B;
    const float* R = a.rotmats + (size_t)(live ? h : 0) * 144;
    const float* be = a.betas + (size_t)(live ? h : 0) * 10;
    // ---- blend coefficients (fp16, duplicated for the hi / lo halves of the basis): 10 per lane ----
    if (live) {
        __half* co = a.coef + (size_t)h * MANO_KTOT;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const int k = j * 10 + i * 2;
            float v[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int kk = k + q;
                float c = 0.f;
                if (kk < 10) c = be[kk];
                else if (kk < 145) {
                    // kk - 10 = (joint-1)*9 + row*3 + col; the diagonal entries ((kk - 10) % 9 in {0,4,8}) come from a
                    // 160-bit constant mask instead of an integer modulo per element
                    constexpr uint32_t DIAG[5] = {0x188c4400u, 0x88c46231u, 0x8c462311u, 0xc4623118u, 0x00011188u};
                    c = R[kk - 1] - (float)((DIAG[kk >> 5] >> (kk & 31)) & 1u);
                }
                v[q] = c;
            }
            const __half2 p = __floats2half2_rn(v[0], v[1]);
            *reinterpret_cast<__half2*>(co + k) = p;
            *reinterpret_cast<__half2*>(co + MANO_KB + k) = p;
        }
    }
    // ---- rest joint of this lane ------------------------------------------------------------------
    float Jj[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float v = m.j_template[j * 3 + c];
#pragma unroll
        for (int l = 0; l < 10; ++l) v = fmaf(m.j_shapedirs[(j * 3 + c) * 10 + l], be[l], v);
        Jj[c] = v;
        sJ[slot][j][c] = v;
    }
    __syncthreads();
    // ---- kinematic chain, level by level (a joint's depth = number of ancestors, <= 15) -----------
    const int pa = m.parents[j];
    int depth = 0;
    for (int q = pa; q >= 0; q = m.parents[q]) ++depth;
    float L[12], G[12];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) L[r * 4 + c] = R[j * 9 + r * 3 + c];
        L[r * 4 + 3] = Jj[r] - (pa >= 0 ? sJ[slot][pa][r] : 0.f);
    }
#pragma unroll
    for (int e = 0; e < 12; ++e) G[e] = L[e];
    for (int lev = 0; lev <= m.max_depth; ++lev) {        // host-computed tree depth (3 for MANO): one barrier per level
        if (depth == lev) {
            if (pa >= 0) {
                const float* P = sG[slot][pa];
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float v = P[r * 4 + 0] * L[0 * 4 + c] + P[r * 4 + 1] * L[1 * 4 + c] + P[r * 4 + 2] * L[2 * 4 + c];
                        if (c == 3) v += P[r * 4 + 3];
                        G[r * 4 + c] = v;
                    }
            }
#pragma unroll
            for (int e = 0; e < 12; ++e) sG[slot][j][e] = G[e];
        }
        __syncthreads();
    }
    if (!live) return;
    // ---- outputs: A = G with translation t - R_glob J_rest; posed joint = translation of G --------
    float* Ao = a.A + ((size_t)h * MANO_J + j) * 12;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float4 v;
        v.x = G[r * 4 + 0]; v.y = G[r * 4 + 1]; v.z = G[r * 4 + 2];
        v.w = G[r * 4 + 3] - (G[r * 4 + 0] * Jj[0] + G[r * 4 + 1] * Jj[1] + G[r * 4 + 2] * Jj[2]);
        reinterpret_cast<float4*>(Ao)[r] = v;
    }
    float* jo = a.jpos + ((size_t)h * MANO_J + j) * 3;
    jo[0] = G[3]; jo[1] = G[7]; jo[2] = G[11];
}

struct ManoSkinArgs {
    const float* vposed;    // [B, 2432] fp32 from the GEMM
    const float* A;         // [B,16,12]
    const float* jpos;      // [B,16,3]
    const float* cam_t;     // [B,3] or null (no projection)
    float* verts;           // [B,778,3]
    float* joints;          // [B,21,3]
    float* kp2d;            // [B,21,2] or null
    float focal_over_size;  // 5000/256
    int B;
};

struct SkinSmem {
    float A[SKIN_HANDS][MANO_J][12];
    float tips[SKIN_HANDS][5][3];
};

__global__ void __launch_bounds__(MANO_VTHREADS, 1)
mano_skin_kernel(const ManoDevice m, const ManoSkinArgs a) {
    ptx::pdl_prologue();
    extern __shared__ __align__(16) uint8_t skin_smem_raw[];
    SkinSmem& S = *reinterpret_cast<SkinSmem*>(skin_smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int hand0 = blockIdx.x * SKIN_HANDS;
    const int nh = min(SKIN_HANDS, a.B - hand0);

    {   // skinning transforms of this CTA's hands: contiguous block of nh * 192 floats
        const float4* src = reinterpret_cast<const float4*>(a.A + (size_t)hand0 * MANO_J * 12);
        float4* dst = reinterpret_cast<float4*>(&S.A[0][0][0]);
        for (int e = tid; e < nh * MANO_J * 3; e += MANO_VTHREADS) dst[e] = src[e];
    }
    const int v = m.perm[tid];                                   // vertex of this thread (or -1)
    const int nu = m.warp_nu[warp];
    int uj[MANO_MAX_UNION];
    float uw[MANO_MAX_UNION];
#pragma unroll
    for (int u = 0; u < MANO_MAX_UNION; ++u) {
        uj[u] = m.warp_union[warp * MANO_MAX_UNION + u];
        uw[u] = m.lane_w[(size_t)tid * MANO_MAX_UNION + u];
    }
    int tip_slot = -1;
#pragma unroll
    for (int q = 0; q < 5; ++q) if (m.tip_ids[q] == v) tip_slot = q;
    __syncthreads();

    // Warps are independent from here on (no staging, no per-hand barrier): lane = vertex in natural
    // order, so a warp reads / writes 384 contiguous bytes of v_posed / vertices per hand, and 25 warps
    // x 16 hands of independent work hide the L2 latency.
    float x = 0.f, y = 0.f, z = 0.f;
    if (v >= 0) {
        const float* vp = a.vposed + (size_t)hand0 * MANO_NPAD + v * 3;
        x = vp[0]; y = vp[1]; z = vp[2];
    }
    for (int h = 0; h < nh; ++h) {
        float nx = 0.f, ny = 0.f, nz = 0.f;
        if (v >= 0 && h + 1 < nh) {                                // prefetch the next hand's vertex
            const float* vp = a.vposed + (size_t)(hand0 + h + 1) * MANO_NPAD + v * 3;
            nx = vp[0]; ny = vp[1]; nz = vp[2];
        }
        // T = sum_u w_u A_u with packed fp32 FMAs (FFMA2: two lanes of fp32 per instruction on sm_100)
        float2 T[6];
#pragma unroll
        for (int e = 0; e < 6; ++e) T[e] = make_float2(0.f, 0.f);
#pragma unroll
        for (int u = 0; u < MANO_MAX_UNION; ++u) {
            if (u < nu) {                                          // warp-uniform
                const float4* Aj = reinterpret_cast<const float4*>(S.A[h][uj[u]]);   // broadcast reads
                const float4 r0 = Aj[0], r1 = Aj[1], r2 = Aj[2];
                const float2 w2 = make_float2(uw[u], uw[u]);
                T[0] = __ffma2_rn(w2, make_float2(r0.x, r0.y), T[0]);
                T[1] = __ffma2_rn(w2, make_float2(r0.z, r0.w), T[1]);
                T[2] = __ffma2_rn(w2, make_float2(r1.x, r1.y), T[2]);
                T[3] = __ffma2_rn(w2, make_float2(r1.z, r1.w), T[3]);
                T[4] = __ffma2_rn(w2, make_float2(r2.x, r2.y), T[4]);
                T[5] = __ffma2_rn(w2, make_float2(r2.z, r2.w), T[5]);
            }
        }
        const float ox = T[0].x * x + T[0].y * y + T[1].x * z + T[1].y;
        const float oy = T[2].x * x + T[2].y * y + T[3].x * z + T[3].y;
        const float oz = T[4].x * x + T[4].y * y + T[5].x * z + T[5].y;
        if (v >= 0) {
            float* o = a.verts + ((size_t)(hand0 + h) * MANO_V + v) * 3;
            o[0] = ox; o[1] = oy; o[2] = oz;
            if (tip_slot >= 0) { S.tips[h][tip_slot][0] = ox; S.tips[h][tip_slot][1] = oy; S.tips[h][tip_slot][2] = oz; }
        }
        x = nx; y = ny; z = nz;
    }
    __syncthreads();

    // ---- 21 joints in OpenPose order (+ projection) ---------------------------------------------
    for (int e = tid; e < nh * 21; e += MANO_VTHREADS) {
        const int h = e / 21, k = e % 21;
        const int srcj = m.joint_map[k];
        float px, py, pz;
        if (srcj < 16) {
            const float* jp = a.jpos + ((size_t)(hand0 + h) * MANO_J + srcj) * 3;
            px = jp[0]; py = jp[1]; pz = jp[2];
        } else {
            px = S.tips[h][srcj - 16][0]; py = S.tips[h][srcj - 16][1]; pz = S.tips[h][srcj - 16][2];
        }
        float* jo = a.joints + ((size_t)(hand0 + h) * 21 + k) * 3;
        jo[0] = px; jo[1] = py; jo[2] = pz;
        if (a.kp2d != nullptr) {
            const float* t = a.cam_t + (size_t)(hand0 + h) * 3;
            const float X = px + t[0], Y = py + t[1], Z = pz + t[2];
            float* ko = a.kp2d + ((size_t)(hand0 + h) * 21 + k) * 2;
            ko[0] = a.focal_over_size * (X / Z);
            ko[1] = a.focal_over_size * (Y / Z);
        }
    }
}

}  // namespace hm
