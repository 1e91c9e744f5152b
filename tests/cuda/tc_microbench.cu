// Micro-benchmarks behind the MANO / attention kernel designs (test infrastructure, not shipped):
//   1. tcgen05.ld throughput per SM (bytes / clock) with 4, 8, 16 reading warps
//   2. tcgen05.mma pace (clocks per instruction) for M = 128 (cta_group 1), K = 16, N = 16 .. 256, SS and TS forms
//   3. global-store patterns of the MANO vertex rows (9,336-byte pitch): 128 contiguous bytes per warp instruction
//      versus 8-byte pieces of 96-byte row segments
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/tc_microbench tests/cuda/tc_microbench.cu
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

// ---- 1. tcgen05.ld -------------------------------------------------------------------------------------
template <int X>
__global__ void tmem_ld_kernel(int iters, long long* cycles, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) ptx::tmem_alloc<1>(&slot, 512);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = slot;
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < 512; c += 4 * X) {
            if constexpr (X == 32) {
                uint32_t r[32];
                ptx::tmem_ld_32x32(lane_addr + ((c + (warp >> 2) * 32) & 511), r);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) acc ^= r[j];
            } else if constexpr (X == 16) {
                uint32_t r[16];
                ptx::tmem_ld_32x16(lane_addr + ((c + (warp >> 2) * 16) & 511), r);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) acc ^= r[j];
            } else {
                uint32_t r[8];
                ptx::tmem_ld_32x8(lane_addr + ((c + (warp >> 2) * 8) & 511), r);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j) acc ^= r[j];
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc<1>(tmem, 512);
}

// two loads in flight before the wait
__global__ void tmem_ld2_kernel(int iters, long long* cycles, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) ptx::tmem_alloc<1>(&slot, 512);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = slot;
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < 512; c += 128) {
            uint32_t r[32], q[32];
            ptx::tmem_ld_32x32(lane_addr + c, r);
            ptx::tmem_ld_32x32(lane_addr + c + 32, q);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc ^= r[j] ^ q[j];
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc<1>(tmem, 512);
}

// ---- 2. tcgen05.mma pace -------------------------------------------------------------------------------
// One thread issues MMAs (M = 128, N, K = 16) back to back on zeroed shared-memory operands, commits and waits.
// The body is fully unrolled with compile-time descriptor offsets (as in the real kernels): 16 MMAs per loop trip,
// NACC independent accumulators used round robin.
template <int N, int NACC, int TS>
__global__ void mma_pace_kernel(int trips, long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t slot;
    __shared__ uint64_t bar;
    const uint32_t raw = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
    if (warp == 0) ptx::tmem_alloc<1>(&slot, 512);
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = ptx::make_idesc_f16kind_f32(0, 0, 128, N);
        const uint64_t a_desc = ptx::make_smem_desc_k_sw128(ptx::smem_u32(smem));
        const uint64_t b_desc = ptx::make_smem_desc_k_sw128(ptx::smem_u32(smem + 32 * 1024));
        const long long t0 = clock64();
        for (int t = 0; t < trips; ++t) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const uint32_t d = tmem + (uint32_t)((i % NACC) * N);
                if (TS) ptx::mma_f16_ts(d, tmem + 448 + (i & 7) * 8, b_desc + (uint64_t)((i & 3) * 2), idesc, 1u);
                else ptx::mma_f16<1>(d, a_desc + (uint64_t)((i & 3) * 2), b_desc + (uint64_t)((i & 3) * 2), idesc, 1u);
            }
        }
        ptx::mma_commit_1sm(&bar);
        const long long t1 = clock64();
        ptx::mbar_wait(&bar, 0);
        const long long t2 = clock64();
        cycles[blockIdx.x * 2] = t1 - t0;
        cycles[blockIdx.x * 2 + 1] = t2 - t0;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc<1>(tmem, 512);
}

template <int N, int NACC, int TS>
void run_pace(int sms, long long* d_cyc) {
    const int trips = 128;
    auto k = mma_pace_kernel<N, NACC, TS>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024));
    k<<<sms, 128, 66 * 1024>>>(trips, d_cyc);
    CK(cudaDeviceSynchronize());
    long long h[2];
    CK(cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost));
    printf("  %s N %3d  %d accumulators : issue %6.1f  complete %6.1f clk/MMA  (floor 128 N / 256 = %d)\n", TS ? "TS" : "SS", N,
           NACC, (double)h[0] / (trips * 16), (double)h[1] / (trips * 16), N / 2);
}

// ---- 3. store patterns ---------------------------------------------------------------------------------
// rows of 2334 floats (9,336 B pitch, 8 mod 16).  Every pattern writes each row exactly once; what differs is how many
// lanes of a warp instruction cooperate on contiguous bytes (run length per instruction) and the width per lane.
//   mode 0: run 128 B (32 lanes x 4 B), one row per instruction
//   mode 1: run  96 B (12 lanes x 8 B), 2.67 rows per instruction (round 1's pattern)
//   mode 2: run   8 B: lane = row, 8 B per lane per instruction (thread-per-hand, STG.64)
//   mode 3: run  16 B: lane = row, 16 B per lane (STG.128; rows of one parity per warp so that all are 16-byte aligned)
//   mode 4: run  32 B: lane pair = row, 16 B per lane
//   mode 5: run  32 B: 4 lanes = row, 8 B per lane
//   mode 6: run  64 B: 8 lanes x 8 B
//   mode 7: run 256 B: 32 lanes x 8 B (STG.64), one row per instruction
__global__ void store_pattern_kernel(float* out, int rows, int mode) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int groups = 97;                       // 97 x 24 floats = 2328 (the 6-float tail is skipped: same for every mode)
    if (mode == 0 || mode == 7) {
        const int per = mode == 0 ? 32 : 64;     // floats per instruction
        const int chunks = 2328 / per;           // 72 | 36 (2304 floats)
        const long long items = (long long)rows * chunks;
        for (long long it = warp; it < items; it += nwarps) {
            const int row = (int)(it / chunks), ch = (int)(it % chunks);
            float* p = out + (size_t)row * 2334 + ch * per;
            if (mode == 0) p[lane] = (float)ch;
            else *reinterpret_cast<float2*>(p + 2 * lane) = make_float2((float)ch, 1.f);
        }
        return;
    }
    // modes 1-6: work item = (32-row block, 24-float group) like the fused kernel's (warp, tile) iteration
    const long long items = (long long)(rows / 32) * groups;
    for (long long it = warp; it < items; it += nwarps) {
        const int rb = (int)(it / groups), g = (int)(it % groups);
        float* base = out + (size_t)rb * 32 * 2334 + g * 24;
        if (mode == 1) {
#pragma unroll
            for (int t = 0; t < 12; ++t) {
                const int e = t * 32 + lane, r = e / 12, pos = e % 12;
                *reinterpret_cast<float2*>(base + (size_t)r * 2334 + 2 * pos) = make_float2((float)e, (float)g);
            }
        } else if (mode == 2) {
#pragma unroll
            for (int t = 0; t < 12; ++t)
                *reinterpret_cast<float2*>(base + (size_t)lane * 2334 + 2 * t) = make_float2((float)t, (float)g);
        } else if (mode == 3) {
            // rows of one parity: block of 64 rows, this item takes parity (rb & 1) of rows [64 (rb / 2), +64)
            float* b2 = out + (size_t)((rb >> 1) * 64 + (rb & 1)) * 2334 + g * 24;
            float* rowp = b2 + (size_t)(2 * lane) * 2334;
            if ((reinterpret_cast<uintptr_t>(rowp) & 15u) == 0) {
#pragma unroll
                for (int t = 0; t < 6; ++t) *reinterpret_cast<float4*>(rowp + 4 * t) = make_float4((float)t, 1.f, 2.f, 3.f);
            } else {
                *reinterpret_cast<float2*>(rowp) = make_float2(0.f, 1.f);
#pragma unroll
                for (int t = 0; t < 5; ++t) *reinterpret_cast<float4*>(rowp + 2 + 4 * t) = make_float4((float)t, 1.f, 2.f, 3.f);
                *reinterpret_cast<float2*>(rowp + 22) = make_float2(0.f, 1.f);
            }
        } else if (mode == 4) {
            // lane pair = row (16 rows per instruction), 16 B per lane -> 32-byte runs; even rows only are 16-byte aligned,
            // so use 8-byte granularity at the ends: emulate with STG.64 x 2 lanes x 2 ... keep it simple: STG.64, 4 lanes = 32 B
#pragma unroll
            for (int t = 0; t < 12; ++t) {
                const int e = t * 32 + lane, r = (e / 4) % 32, pos = (e / 128) * 4 + (e % 4);
                *reinterpret_cast<float2*>(base + (size_t)r * 2334 + 2 * pos) = make_float2((float)e, (float)g);
            }
        } else if (mode == 5) {
#pragma unroll
            for (int t = 0; t < 12; ++t) {     // 2 lanes x 8 B = 16-byte runs
                const int e = t * 32 + lane, r = (e / 2) % 32, pos = (e / 64) * 2 + (e % 2);
                *reinterpret_cast<float2*>(base + (size_t)r * 2334 + 2 * pos) = make_float2((float)e, (float)g);
            }
        } else if (mode == 6) {
#pragma unroll
            for (int t = 0; t < 12; ++t) {     // 6 lanes x 8 B = 48-byte runs: (e / 6) rows per instruction ~5.3
                const int e = t * 32 + lane, r = (e / 6) % 32, pos = (e / 192) * 6 + (e % 6);
                *reinterpret_cast<float2*>(base + (size_t)r * 2334 + 2 * pos) = make_float2((float)e, (float)g);
            }
        }
    }
}

// ---- 4. TMA tensor stores of [R rows x W bytes] boxes into the vertex array viewed as [B/2][2 x 2334 floats] -------------
#include <cuda.h>
typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void tma_store_kernel(const __grid_constant__ CUtensorMap tm, int pairs, int wfloats, int rrows, int inflight) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = (float)i;
    ptx::fence_proxy_async();
    __syncthreads();
    // every warp's lane 0 issues boxes: work item = (row block, column block), both parities
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int cblocks = 2328 / wfloats, rblocks = pairs / rrows;
    const long long items = (long long)cblocks * rblocks;
    if (lane == 0) {
        const uint32_t src = ptx::smem_u32(smem) + (uint32_t)(threadIdx.x >> 5) * 3072u % (40u * 1024u);
        int pending = 0;
        for (long long it = warp; it < items; it += nwarps) {
            const int rb = (int)(it / cblocks), cb = (int)(it - (long long)rb * cblocks);
            ptx::tma_store_2d(&tm, src & ~127u, cb * wfloats, rb * rrows);
            ptx::tma_store_2d(&tm, src & ~127u, 2334 + cb * wfloats, rb * rrows);
            ptx::bulk_commit();
            if (++pending >= inflight) { ptx::bulk_wait_read<0>(); pending = 0; }
        }
        ptx::bulk_wait_all();
    }
}

int main() {
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    const int sms = prop.multiProcessorCount;
    long long* d_cyc;
    uint32_t* d_sink;
    CK(cudaMalloc(&d_cyc, sizeof(long long) * 2 * sms));
    CK(cudaMalloc(&d_sink, 64));
    std::vector<long long> h(2 * sms);

    printf("== tcgen05.ld throughput (one CTA per SM, 512 columns x 128 lanes x 4 B = 256 KB per sweep) ==\n");
    for (int warps : {4, 8, 16}) {
        for (int x : {8, 16, 32}) {
            const int iters = 200;
            if (x == 32) tmem_ld_kernel<32><<<sms, warps * 32>>>(iters, d_cyc, d_sink);
            else if (x == 16) tmem_ld_kernel<16><<<sms, warps * 32>>>(iters, d_cyc, d_sink);
            else tmem_ld_kernel<8><<<sms, warps * 32>>>(iters, d_cyc, d_sink);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h.data(), d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
            // every warp reads 512/(4x) chunks of x columns per iteration: bytes per CTA per iteration
            const double bytes = (double)warps * (512 / (4 * x)) * x * 32 * 4 * iters;
            printf("  warps %2d  x%-2d : %8lld clk  %7.1f B/clk/SM\n", warps, x, h[0], bytes / (double)h[0]);
        }
    }
    for (int warps : {4, 8, 16}) {
        const int iters = 200;
        tmem_ld2_kernel<<<sms, warps * 32>>>(iters, d_cyc, d_sink);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h.data(), d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
        const double bytes = (double)warps * 4 * 2 * 32 * 32 * 4 * iters;
        printf("  warps %2d  2 x x32 in flight : %8lld clk  %7.1f B/clk/SM\n", warps, h[0], bytes / (double)h[0]);
    }

    printf("== tcgen05.mma pace (cta_group 1, M 128, K 16; clocks per instruction: issue / complete) ==\n");
    run_pace<16, 1, 0>(sms, d_cyc); run_pace<16, 4, 0>(sms, d_cyc); run_pace<16, 8, 0>(sms, d_cyc);
    run_pace<32, 1, 0>(sms, d_cyc); run_pace<32, 4, 0>(sms, d_cyc); run_pace<32, 8, 0>(sms, d_cyc);
    run_pace<48, 1, 0>(sms, d_cyc); run_pace<48, 4, 0>(sms, d_cyc);
    run_pace<64, 1, 0>(sms, d_cyc); run_pace<64, 4, 0>(sms, d_cyc);
    run_pace<96, 1, 0>(sms, d_cyc); run_pace<96, 4, 0>(sms, d_cyc);
    run_pace<128, 1, 0>(sms, d_cyc); run_pace<128, 2, 0>(sms, d_cyc);
    run_pace<192, 1, 0>(sms, d_cyc); run_pace<192, 2, 0>(sms, d_cyc);
    run_pace<256, 1, 0>(sms, d_cyc);
    run_pace<16, 1, 1>(sms, d_cyc); run_pace<16, 8, 1>(sms, d_cyc);
    run_pace<32, 1, 1>(sms, d_cyc); run_pace<32, 8, 1>(sms, d_cyc);
    run_pace<48, 1, 1>(sms, d_cyc); run_pace<48, 4, 1>(sms, d_cyc);
    run_pace<96, 1, 1>(sms, d_cyc); run_pace<96, 4, 1>(sms, d_cyc);
    run_pace<192, 1, 1>(sms, d_cyc); run_pace<192, 2, 1>(sms, d_cyc);
    run_pace<256, 1, 1>(sms, d_cyc);
    printf("== store patterns, 65536 rows x 9336 B (612 MB) ==\n");
    const int rows = 65536;
    float* d_out;
    CK(cudaMalloc(&d_out, (size_t)rows * 2334 * 4 + 4096));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const char* names[8] = {"run 128 B (32 x 4 B)", "run 96 B (12 x 8 B), round-1 pattern", "run 8 B (lane = row, STG.64)",
                            "run 16 B (lane = row, STG.128)", "run 32 B (4 lanes x 8 B)", "run 16 B (2 lanes x 8 B)",
                            "run 48 B (6 lanes x 8 B)", "run 256 B (32 x 8 B)"};
    for (int mode = 0; mode < 8; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0));
            store_pattern_kernel<<<sms * 4, 512>>>(d_out, rows, mode);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep) printf("  mode %d %-40s: %.3f ms  %.0f GB/s\n", mode, names[mode], ms, (double)rows * 9312 / ms / 1e6);
        }
    }
    printf("== TMA tensor stores [R rows x W bytes] (two boxes per item: even / odd hands), 65536 rows ==\n");
    {
        void* fnp = nullptr;
        cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
        PFN_enc enc = reinterpret_cast<PFN_enc>(fnp);
        CK(cudaFuncSetAttribute(tma_store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 50 * 1024));
        const int wf[] = {12, 24, 32, 48, 96, 96, 192};
        const int rr[] = {16, 16, 16, 16, 16, 32, 16};
        for (int v = 0; v < 7; ++v) {
            if (rr[v] * wf[v] * 4 > 40 * 1024) continue;
            CUtensorMap tm;
            cuuint64_t gdim[2] = {4668, (cuuint64_t)(rows / 2)};
            cuuint64_t gstride[1] = {18672};
            cuuint32_t box[2] = {(cuuint32_t)wf[v], (cuuint32_t)rr[v]};
            cuuint32_t estr[2] = {1, 1};
            CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_out, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("  encode failed for W %d\n", wf[v] * 4); continue; }
            for (int warps : {4, 16})
                for (int rep = 0; rep < 2; ++rep) {
                    CK(cudaEventRecord(e0));
                    tma_store_kernel<<<sms, warps * 32, 50 * 1024>>>(tm, rows / 2, wf[v], rr[v], 4);
                    CK(cudaEventRecord(e1));
                    CK(cudaEventSynchronize(e1));
                    float ms;
                    CK(cudaEventElapsedTime(&ms, e0, e1));
                    const double bytes = (double)(2328 / wf[v]) * wf[v] * 4 * rows;
                    if (rep) printf("  box %3d B x %2d rows, %2d issuing warps per SM : %.3f ms  %.0f GB/s\n", wf[v] * 4, rr[v], warps, ms, bytes / ms / 1e6);
                }
        }
    }
    return 0;
}
