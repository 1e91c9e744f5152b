// MANO layer: blend shapes + joint regression + kinematic chain + linear blend skinning,
// fingertip append / OpenPose re-ordering and the crop-camera projection, fused in one kernel.
//
// Replaces MANO.forward (hamer/hamer/models/mano_wrapper.py:32-44) -> smplx 0.1.28
// MANOLayer.forward -> lbs() (third-party; algorithm in SURVEY 8a row a16) and
// perspective_projection (hamer/hamer/utils/geometry.py:72-109) as called from
// HAMER.forward_step (hamer/hamer/models/hamer.py:143-152).
//
// v1 (fp32 SIMT): one CTA handles LBS_HANDS hands; thread = vertex.  The 146 x 2334 blend basis
// ([v_template | shapedirs | posedirs], re-laid out as [k][coord][vertex] so that vertex-indexed
// reads coalesce) streams from L2 once per CTA and is reused across the CTA's hands from
// registers; the per-hand coefficients sit in shared memory.  Joints use the folded regressor
// J = J_template + J_shapedirs * beta (J_regressor applied to the basis once at set-up).
// Skinning uses each vertex's <= 8 non-zero weights held in registers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hm {

constexpr int MANO_V = 778;
constexpr int MANO_J = 16;
constexpr int MANO_K = 146;            // 1 (template) + 10 (shape) + 135 (pose)
constexpr int MANO_VPAD = 800;         // vertex pitch of the re-laid-out basis (25 warps)
constexpr int MANO_MAX_NNZ = 8;
constexpr int LBS_HANDS = 16;
constexpr int LBS_THREADS = 800;

struct ManoDevice {
    const float* basis;        // [146][3][800]  k=0 template, 1..10 shapedirs, 11..145 posedirs
    const float* j_template;   // [16][3]
    const float* j_shapedirs;  // [16][3][10]
    const int32_t* skin_idx;   // [800][8]
    const float* skin_w;       // [800][8]
    int32_t parents[16];
    int32_t tip_ids[5];
    int32_t joint_map[21];
};

struct LbsArgs {
    const float* betas;     // [B,10]
    const float* rotmats;   // [B,16,3,3]
    const float* cam_t;     // [B,3] or null (no projection)
    float* verts;           // [B,778,3]
    float* joints;          // [B,21,3]
    float* kp2d;            // [B,21,2] or null
    float focal_over_size;  // 5000/256
    int B;
};

struct LbsSmem {
    float coef[MANO_K][LBS_HANDS];        // [k][hand] blend coefficients
    float A[LBS_HANDS][MANO_J][12];       // skinning transforms (3x4, row-major)
    float G[LBS_HANDS][MANO_J][12];       // global joint transforms
    float Jrest[LBS_HANDS][MANO_J][3];
    float j21[LBS_HANDS][21][3];
    float stage[2][MANO_V * 3];           // double-buffered vertex staging for coalesced stores
};

__global__ void __launch_bounds__(LBS_THREADS, 1)
mano_lbs_kernel(const ManoDevice m, const LbsArgs a) {
    extern __shared__ __align__(16) uint8_t lbs_smem_raw[];
    LbsSmem& S = *reinterpret_cast<LbsSmem*>(lbs_smem_raw);
    float (*coef)[LBS_HANDS] = S.coef;
    float (*A)[MANO_J][12] = S.A;
    float (*G)[MANO_J][12] = S.G;
    float (*Jrest)[MANO_J][3] = S.Jrest;
    float (*j21)[21][3] = S.j21;
    float (*stage)[MANO_V * 3] = S.stage;

    const int tid = threadIdx.x;
    const int hand0 = blockIdx.x * LBS_HANDS;
    const int nh = min(LBS_HANDS, a.B - hand0);

    // ---- coefficients: [1 | beta | vec(R_j - I), j = 1..15] -------------------------------------
    for (int e = tid; e < MANO_K * LBS_HANDS; e += LBS_THREADS) {
        const int k = e / LBS_HANDS, h = e % LBS_HANDS;
        float c = 0.f;
        if (h < nh) {
            const size_t hb = (size_t)(hand0 + h);
            if (k == 0) c = 1.f;
            else if (k <= 10) c = a.betas[hb * 10 + (k - 1)];
            else {
                const int q = k - 11;                               // (joint-1)*9 + row*3 + col
                const int e9 = q % 9;
                c = a.rotmats[hb * 144 + 9 + q] - ((e9 == 0 || e9 == 4 || e9 == 8) ? 1.f : 0.f);
            }
        }
        coef[k][h] = c;
    }
    // ---- rest joints: J = J_template + J_shapedirs * beta ---------------------------------------
    for (int e = tid; e < LBS_HANDS * MANO_J * 3; e += LBS_THREADS) {
        const int h = e / (MANO_J * 3), jc = e % (MANO_J * 3);
        float v = m.j_template[jc];
        if (h < nh) {
            const float* be = a.betas + (size_t)(hand0 + h) * 10;
#pragma unroll
            for (int l = 0; l < 10; ++l) v = fmaf(m.j_shapedirs[jc * 10 + l], be[l], v);
        }
        Jrest[h][jc / 3][jc % 3] = v;
    }
    __syncthreads();

    // ---- kinematic chain: one thread per hand walks the 16 joints (parents precede children) ----
    if (tid < nh) {
        const int h = tid;
        const float* R = a.rotmats + (size_t)(hand0 + h) * 144;
        for (int j = 0; j < MANO_J; ++j) {
            const int pa = m.parents[j];
            float L[12];                                            // local [R | J_j - J_parent]
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
                for (int c = 0; c < 3; ++c) L[r * 4 + c] = R[j * 9 + r * 3 + c];
                L[r * 4 + 3] = Jrest[h][j][r] - (pa >= 0 ? Jrest[h][pa][r] : 0.f);
            }
            float Gj[12];
            if (pa < 0) {
#pragma unroll
                for (int e = 0; e < 12; ++e) Gj[e] = L[e];
            } else {
                const float* P = G[h][pa];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float v = P[r * 4 + 0] * L[0 * 4 + c] + P[r * 4 + 1] * L[1 * 4 + c] +
                                  P[r * 4 + 2] * L[2 * 4 + c];
                        if (c == 3) v += P[r * 4 + 3];
                        Gj[r * 4 + c] = v;
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < 12; ++e) G[h][j][e] = Gj[e];
            // A = G with translation t - R_glob * J_rest
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                A[h][j][r * 4 + 0] = Gj[r * 4 + 0];
                A[h][j][r * 4 + 1] = Gj[r * 4 + 1];
                A[h][j][r * 4 + 2] = Gj[r * 4 + 2];
                A[h][j][r * 4 + 3] = Gj[r * 4 + 3] - (Gj[r * 4 + 0] * Jrest[h][j][0] +
                                                     Gj[r * 4 + 1] * Jrest[h][j][1] +
                                                     Gj[r * 4 + 2] * Jrest[h][j][2]);
            }
        }
    }

    // ---- blend shapes: v_posed[h][v][:] = sum_k coef[k][h] * basis[k][:][v] ----------------------
    float vp[LBS_HANDS][3];
#pragma unroll
    for (int h = 0; h < LBS_HANDS; ++h) vp[h][0] = vp[h][1] = vp[h][2] = 0.f;
    const float* bp = m.basis + tid;
#pragma unroll 2
    for (int k = 0; k < MANO_K; ++k) {
        const float bx = __ldg(bp + (size_t)(k * 3 + 0) * MANO_VPAD);
        const float by = __ldg(bp + (size_t)(k * 3 + 1) * MANO_VPAD);
        const float bz = __ldg(bp + (size_t)(k * 3 + 2) * MANO_VPAD);
#pragma unroll
        for (int h4 = 0; h4 < LBS_HANDS / 4; ++h4) {
            const float4 c = *reinterpret_cast<const float4*>(&coef[k][h4 * 4]);
            vp[h4 * 4 + 0][0] = fmaf(c.x, bx, vp[h4 * 4 + 0][0]);
            vp[h4 * 4 + 0][1] = fmaf(c.x, by, vp[h4 * 4 + 0][1]);
            vp[h4 * 4 + 0][2] = fmaf(c.x, bz, vp[h4 * 4 + 0][2]);
            vp[h4 * 4 + 1][0] = fmaf(c.y, bx, vp[h4 * 4 + 1][0]);
            vp[h4 * 4 + 1][1] = fmaf(c.y, by, vp[h4 * 4 + 1][1]);
            vp[h4 * 4 + 1][2] = fmaf(c.y, bz, vp[h4 * 4 + 1][2]);
            vp[h4 * 4 + 2][0] = fmaf(c.z, bx, vp[h4 * 4 + 2][0]);
            vp[h4 * 4 + 2][1] = fmaf(c.z, by, vp[h4 * 4 + 2][1]);
            vp[h4 * 4 + 2][2] = fmaf(c.z, bz, vp[h4 * 4 + 2][2]);
            vp[h4 * 4 + 3][0] = fmaf(c.w, bx, vp[h4 * 4 + 3][0]);
            vp[h4 * 4 + 3][1] = fmaf(c.w, by, vp[h4 * 4 + 3][1]);
            vp[h4 * 4 + 3][2] = fmaf(c.w, bz, vp[h4 * 4 + 3][2]);
        }
    }
    __syncthreads();   // A / G ready

    // ---- skinning + output ---------------------------------------------------------------------
    int sidx[MANO_MAX_NNZ];
    float sw[MANO_MAX_NNZ];
#pragma unroll
    for (int q = 0; q < MANO_MAX_NNZ; ++q) {
        sidx[q] = m.skin_idx[tid * MANO_MAX_NNZ + q];
        sw[q] = m.skin_w[tid * MANO_MAX_NNZ + q];
    }
    int tip_slot = -1;
#pragma unroll
    for (int q = 0; q < 5; ++q) if (m.tip_ids[q] == tid) tip_slot = 16 + q;

#pragma unroll
    for (int h = 0; h < LBS_HANDS; ++h) {
        if (h < nh) {                                              // block-uniform
            float T[12];
#pragma unroll
            for (int e = 0; e < 12; ++e) T[e] = 0.f;
#pragma unroll
            for (int q = 0; q < MANO_MAX_NNZ; ++q) {
                if (sw[q] != 0.f) {
                    const float4* Aj = reinterpret_cast<const float4*>(A[h][sidx[q]]);
                    const float4 r0 = Aj[0], r1 = Aj[1], r2 = Aj[2];
                    T[0] = fmaf(sw[q], r0.x, T[0]); T[1] = fmaf(sw[q], r0.y, T[1]);
                    T[2] = fmaf(sw[q], r0.z, T[2]); T[3] = fmaf(sw[q], r0.w, T[3]);
                    T[4] = fmaf(sw[q], r1.x, T[4]); T[5] = fmaf(sw[q], r1.y, T[5]);
                    T[6] = fmaf(sw[q], r1.z, T[6]); T[7] = fmaf(sw[q], r1.w, T[7]);
                    T[8] = fmaf(sw[q], r2.x, T[8]); T[9] = fmaf(sw[q], r2.y, T[9]);
                    T[10] = fmaf(sw[q], r2.z, T[10]); T[11] = fmaf(sw[q], r2.w, T[11]);
                }
            }
            const float x = vp[h][0], y = vp[h][1], z = vp[h][2];
            const float ox = T[0] * x + T[1] * y + T[2] * z + T[3];
            const float oy = T[4] * x + T[5] * y + T[6] * z + T[7];
            const float oz = T[8] * x + T[9] * y + T[10] * z + T[11];
            float* st = stage[h & 1];
            if (tid < MANO_V) {
                st[tid * 3 + 0] = ox; st[tid * 3 + 1] = oy; st[tid * 3 + 2] = oz;
                if (tip_slot >= 0) { j21[h][tip_slot][0] = ox; j21[h][tip_slot][1] = oy; j21[h][tip_slot][2] = oz; }
            }
            __syncthreads();
            // coalesced 8-byte stores (a hand's 9336-byte block is only 8-byte aligned)
            float2* dst = reinterpret_cast<float2*>(a.verts + (size_t)(hand0 + h) * MANO_V * 3);
            const float2* src = reinterpret_cast<const float2*>(st);
            for (int e = tid; e < MANO_V * 3 / 2; e += LBS_THREADS) dst[e] = src[e];
        }
    }
    __syncthreads();

    // ---- 21 joints in OpenPose order (+ projection) ---------------------------------------------
    for (int e = tid; e < nh * 21; e += LBS_THREADS) {
        const int h = e / 21, k = e % 21;
        const int srcj = m.joint_map[k];
        float px, py, pz;
        if (srcj < 16) { px = G[h][srcj][3]; py = G[h][srcj][7]; pz = G[h][srcj][11]; }
        else { px = j21[h][srcj][0]; py = j21[h][srcj][1]; pz = j21[h][srcj][2]; }
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
