// Probe (test infrastructure, not shipped): how fast can the MANO vertex array [B][2334] floats (row pitch 9,336 B = 8 mod 16)
// be written from per-warp staging tiles [32 hands][RUN bytes] when every LANE issues one 1-D bulk copy
// (cp.async.bulk.global.shared::cta) of its own row piece?  Even hands' pieces are 16-byte aligned; odd hands' pieces start
// 8 bytes into a 16-byte line, so the odd lanes copy bytes [8, RUN - 8) in bulk and write the first / last 8 bytes with STG.64.
//   bulk_store_probe <mode 0 = LDS.64 + STG.64 (shipping pattern), 1 = per-lane bulk copies> <RUN bytes: 96 | 192 | 384>
//                    <warps per CTA> <tiles in flight per warp: 1 | 2> <CTAs per SM>
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/bulk_store_probe tests/cuda/bulk_store_probe.cu
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../hamer_yolo_b200/csrc/ptx.cuh"

using namespace hm;

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                   \
        }                                                                              \
    } while (0)

template <int N>
__device__ __forceinline__ void bulk_wait_read_n() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

constexpr int ROWF = 2334;

// value written at (hand, float column)
__host__ __device__ inline float want(int hand, int col) { return (float)((hand * 131 + col * 7) & 0xffff); }

template <int RUN>
__global__ void store_kernel(float* __restrict__ out, int hands, int mode, int inflight) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int RF = RUN / 4;                       // floats per row piece
    constexpr uint32_t PITCH = RUN + 48;              // staging row: 16 B of lead room (odd lanes stage at + 8), odd chunk count
    static_assert((PITCH / 16) % 2 == 1, "conflict-free STS.128");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const uint32_t sbase = ptx::smem_u32(smem_raw) + (uint32_t)warp * 2u * 32u * PITCH;
    const int cgroups = ROWF * 4 / RUN;               // whole column groups (the ragged tail is not part of the probe)
    const int rblocks = hands / 32;
    const long long tiles = (long long)cgroups * rblocks;
    const long long gw = (long long)blockIdx.x * nw + warp, gnw = (long long)gridDim.x * nw;
    int buf = 0;
    for (long long t = gw; t < tiles; t += gnw) {
        // consecutive warps of a CTA take consecutive column groups of the same 32 hands (like the 4 parts of a fused tile)
        const int rb = (int)(t / cgroups), cg = (int)(t - (long long)rb * cgroups);
        const int hand = rb * 32 + lane, col0 = cg * RF;
        const uint32_t tile = sbase + (uint32_t)buf * 32u * PITCH;
        const uint32_t srow = tile + (uint32_t)lane * PITCH;
        if (mode == 1) {
            if (inflight == 2) bulk_wait_read_n<1>(); else bulk_wait_read_n<0>();
        }
        __syncwarp();
        const bool odd = hand & 1;
        // stage: thread = row.  Bulk mode: odd lanes place their row 8 bytes late, so that byte 8 of the piece is 16-byte aligned
        // in shared memory exactly where it is in global memory
        const uint32_t s0 = srow + 16u - ((mode == 1 && odd) ? 8u : 0u);
        if (mode == 1 && odd) {
            ptx::sts64(s0, make_float2(want(hand, col0), want(hand, col0 + 1)));
#pragma unroll
            for (int q = 0; q < RF / 4 - 1; ++q)
                ptx::sts128(s0 + 8 + 16 * q, make_float4(want(hand, col0 + 2 + 4 * q), want(hand, col0 + 3 + 4 * q),
                                                         want(hand, col0 + 4 + 4 * q), want(hand, col0 + 5 + 4 * q)));
            ptx::sts64(s0 + RUN - 8, make_float2(want(hand, col0 + RF - 2), want(hand, col0 + RF - 1)));
        } else {
#pragma unroll
            for (int q = 0; q < RF / 4; ++q)
                ptx::sts128(s0 + 16 * q, make_float4(want(hand, col0 + 4 * q), want(hand, col0 + 1 + 4 * q),
                                                     want(hand, col0 + 2 + 4 * q), want(hand, col0 + 3 + 4 * q)));
        }
        float* g = out + (size_t)hand * ROWF + col0;
        if (mode == 1) {
            ptx::fence_proxy_async();
            if (!odd) {
                ptx::bulk_store_1d(g, s0, RUN);
            } else {
                ptx::bulk_store_1d(g + 2, s0 + 8, RUN - 16);
                *reinterpret_cast<float2*>(g) = make_float2(want(hand, col0), want(hand, col0 + 1));
                *reinterpret_cast<float2*>(g + RF - 2) = make_float2(want(hand, col0 + RF - 2), want(hand, col0 + RF - 1));
            }
            ptx::bulk_commit();
            if (inflight == 2) buf ^= 1;
        } else {
            __syncwarp();
            constexpr int HALF = RF / 2;                 // 8-byte pairs per row
            float* gb = out + (size_t)(rb * 32) * ROWF + col0;
#pragma unroll
            for (int k = 0; k < HALF; ++k) {
                const int e = k * 32 + lane, r = e / HALF, pos = e % HALF;
                *reinterpret_cast<float2*>(gb + (size_t)r * ROWF + 2 * pos) = ptx::lds64(tile + (uint32_t)r * PITCH + 16u + 8u * pos);
            }
        }
    }
    if (mode == 1) ptx::bulk_wait_all();
}

template <int RUN>
float run(float* d_out, int hands, int mode, int warps, int inflight, int ctas_per_sm, int sms) {
    const int smem = warps * 2 * 32 * (RUN + 48);
    CK(cudaFuncSetAttribute(store_kernel<RUN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float ms = 0, best = 1e9f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0));
        store_kernel<RUN><<<sms * ctas_per_sm, warps * 32, smem>>>(d_out, hands, mode, inflight);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    return best;
}

int main(int argc, char** argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 1, runb = argc > 2 ? atoi(argv[2]) : 96, warps = argc > 3 ? atoi(argv[3]) : 16;
    const int inflight = argc > 4 ? atoi(argv[4]) : 2, cps = argc > 5 ? atoi(argv[5]) : 1;
    CK(cudaSetDevice(0));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int hands = 65536;
    const size_t n = (size_t)hands * ROWF;
    float* d_out;
    CK(cudaMalloc(&d_out, n * 4 + 4096));
    CK(cudaMemset(d_out, 0xff, n * 4 + 4096));
    float ms = 0;
    if (runb == 96) ms = run<96>(d_out, hands, mode, warps, inflight, cps, prop.multiProcessorCount);
    else if (runb == 192) ms = run<192>(d_out, hands, mode, warps, inflight, cps, prop.multiProcessorCount);
    else if (runb == 384) ms = run<384>(d_out, hands, mode, warps, inflight, cps, prop.multiProcessorCount);
    else { printf("RUN must be 96, 192 or 384\n"); return 1; }
    CK(cudaDeviceSynchronize());
    std::vector<float> h(n);
    CK(cudaMemcpy(h.data(), d_out, n * 4, cudaMemcpyDeviceToHost));
    const int cols = (ROWF * 4 / runb) * (runb / 4);
    long long bad = 0;
    for (int hand = 0; hand < hands; hand += 61)
        for (int c = 0; c < cols; ++c)
            if (h[(size_t)hand * ROWF + c] != want(hand, c)) ++bad;
    const double bytes = (double)cols * 4 * hands;
    printf("mode %d run %3d B, %2d warps x %d CTA/SM, %d in flight: %.3f ms  %.0f GB/s  mismatches %lld\n", mode, runb, warps, cps,
           inflight, ms, bytes / ms / 1e6, bad);
    return 0;
}
