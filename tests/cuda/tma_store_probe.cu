// Probe (test infrastructure, not shipped): can TMA tensor stores write the MANO vertex array [B][2334] floats (row pitch
// 9,336 B = 8 mod 16) when it is viewed as [B / 2][4668] (pitch 18,672 B, a multiple of 16)?  The even hand of a row pair
// starts at column 0 (16-byte aligned), the odd hand at column 2334 (8 mod 16).  One process per variant because a
// rejected instruction kills the context.
//   tma_store_probe <parity mask 1|2|3> <box floats> <box rows> <odd shift floats> <issuing warps> <in flight>
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/tma_store_probe tests/cuda/tma_store_probe.cu
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

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void tma_store_kernel(const __grid_constant__ CUtensorMap tm, int pairs, int wfloats, int rrows, int inflight,
                                 int mask, int oddshift) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    for (int i = threadIdx.x; i < rrows * wfloats; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = (float)i;
    ptx::fence_proxy_async();
    __syncthreads();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int cblocks = 2328 / wfloats, rblocks = pairs / rrows;
    const long long items = (long long)cblocks * rblocks;
    if (lane == 0) {
        const uint32_t src = ptx::smem_u32(smem);
        int pending = 0;
        for (long long it = warp; it < items; it += nwarps) {
            const int rb = (int)(it / cblocks), cb = (int)(it - (long long)rb * cblocks);
            if (mask & 1) ptx::tma_store_2d(&tm, src, cb * wfloats, rb * rrows);
            if (mask & 2) ptx::tma_store_2d(&tm, src, 2334 + oddshift + cb * wfloats, rb * rrows);
            ptx::bulk_commit();
            if (++pending >= inflight) { ptx::bulk_wait_read<0>(); pending = 0; }
        }
        ptx::bulk_wait_all();
    }
}

int main(int argc, char** argv) {
    const int mask = argc > 1 ? atoi(argv[1]) : 3, wf = argc > 2 ? atoi(argv[2]) : 96, rr = argc > 3 ? atoi(argv[3]) : 16;
    const int oddshift = argc > 4 ? atoi(argv[4]) : 0, warps = argc > 5 ? atoi(argv[5]) : 4, inflight = argc > 6 ? atoi(argv[6]) : 4;
    CK(cudaSetDevice(0));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount, rows = 65536;
    float* d_out;
    const size_t n = (size_t)rows * 2334;
    CK(cudaMalloc(&d_out, n * 4 + 4096));
    CK(cudaMemset(d_out, 0xff, n * 4 + 4096));
    void* fnp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
    PFN_enc enc = reinterpret_cast<PFN_enc>(fnp);
    const int smem_bytes = rr * wf * 4 + 2048;
    CK(cudaFuncSetAttribute(tma_store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    CUtensorMap tm;
    cuuint64_t gdim[2] = {4668, (cuuint64_t)(rows / 2)};
    cuuint64_t gstride[1] = {18672};
    cuuint32_t box[2] = {(cuuint32_t)wf, (cuuint32_t)rr};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_out, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed (%d)\n", (int)r); return 1; }
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float ms = 0;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        tma_store_kernel<<<sms, warps * 32, smem_bytes>>>(tm, rows / 2, wf, rr, inflight, mask, oddshift);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
    }
    const int nb = (mask & 1) + ((mask >> 1) & 1);
    const double bytes = (double)(2328 / wf) * wf * 4 * rows * nb / 2;
    // verify
    std::vector<float> h(n);
    CK(cudaMemcpy(h.data(), d_out, n * 4, cudaMemcpyDeviceToHost));
    long long bad = 0;
    const int cblocks = 2328 / wf;
    for (int hand = 0; hand < rows; hand += 97) {
        const int par = hand & 1, pr = hand >> 1, rin = pr % rr;
        if (!((mask >> par) & 1)) continue;
        const int sh = par ? oddshift : 0;
        for (int c = 0; c < cblocks * wf; ++c) {
            const int cc = c + sh;
            if (cc >= 2334) continue;
            const float want = (float)(rin * wf + (c % wf));
            if (h[(size_t)hand * 2334 + cc] != want) ++bad;
        }
    }
    printf("mask %d box %3d B x %2d rows, odd shift %d, %2d warps, %d in flight: %.3f ms  %.0f GB/s  mismatches %lld\n", mask, wf * 4,
           rr, oddshift, warps, inflight, ms, bytes / ms / 1e6, bad);
    return 0;
}
