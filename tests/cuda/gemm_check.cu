// Stand-alone check of the tcgen05 GEMM against a naive CUDA-core GEMM on the same bf16 inputs.
// Test infrastructure only (not part of the shipped library).
//   gemm_check <cta_group> <bn> <epi> <M> <N> <K> [iters] [w_format: 1 bf16 (default) | 0 fp16]
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include <cuda_fp16.h>

#include "../../hamer_yolo_b200/csrc/gemm_host.cuh"

using namespace hm;

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                   \
        }                                                                              \
    } while (0)

__device__ float w2f(uint16_t v, int fmt) {
    if (fmt == 1) return __bfloat162float(*reinterpret_cast<__nv_bfloat16*>(&v));
    return __half2float(*reinterpret_cast<__half*>(&v));
}

__global__ void ref_gemm(const __half* A, const uint16_t* W, int w_fmt, const float* bias,
                         const float* pos, int pos_rows, const float* resid, float* out, int M,
                         int N, int K, int epi) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    int m = blockIdx.y;
    if (n >= N || m >= M) return;
    float acc = 0.f;
    for (int k = 0; k < K; ++k)
        acc += __half2float(A[(size_t)m * K + k]) * w2f(W[(size_t)n * K + k], w_fmt);
    float v = acc + bias[n];
    if (epi == EPI_BIAS_GELU_BF16) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
    if (epi == EPI_BIAS_RESID_F32) v += resid[(size_t)m * N + n];
    if (epi == EPI_BIAS_POS_F32) v += pos[(size_t)(m % pos_rows) * N + n];
    out[(size_t)m * N + n] = v;
}

static uint32_t rng_state = 12345u;
static float frand() {
    rng_state = rng_state * 1664525u + 1013904223u;
    return ((rng_state >> 8) & 0xFFFF) / 65536.0f - 0.5f;
}

int main(int argc, char** argv) {
    if (argc < 7) {
        printf("usage: gemm_check cta_group bn epi M N K [iters]\n");
        return 1;
    }
    int cg = atoi(argv[1]), bn = atoi(argv[2]), epi = atoi(argv[3]);
    int M = atoi(argv[4]), N = atoi(argv[5]), K = atoi(argv[6]);
    int iters = argc > 7 ? atoi(argv[7]) : 10;
    int w_fmt = argc > 8 ? atoi(argv[8]) : 0;
    int dev = 0;
    cudaDeviceProp prop;
    CK(cudaSetDevice(dev));
    CK(cudaGetDeviceProperties(&prop, dev));
    int num_sms = prop.multiProcessorCount;

    std::vector<__half> hA((size_t)M * K);
    std::vector<uint16_t> hW((size_t)N * K);
    std::vector<float> hbias(N), hpos((size_t)192 * N), hres((size_t)M * N);
    for (auto& v : hA) v = __float2half(frand());
    for (auto& v : hW) {
        const float f = frand() * 0.1f;
        if (w_fmt == 1) { __nv_bfloat16 b = __float2bfloat16(f); v = *reinterpret_cast<uint16_t*>(&b); }
        else { __half b = __float2half(f); v = *reinterpret_cast<uint16_t*>(&b); }
    }
    for (auto& v : hbias) v = frand();
    for (auto& v : hpos) v = frand();
    for (auto& v : hres) v = frand();

    __half* dA;
    uint16_t* dW;
    float *dbias, *dpos, *dres, *dref, *dout_f;
    __half* dout_h;
    CK(cudaMalloc(&dA, hA.size() * 2));
    CK(cudaMalloc(&dW, hW.size() * 2));
    CK(cudaMalloc(&dbias, N * 4));
    CK(cudaMalloc(&dpos, hpos.size() * 4));
    CK(cudaMalloc(&dres, hres.size() * 4));
    CK(cudaMalloc(&dref, (size_t)M * N * 4));
    CK(cudaMalloc(&dout_f, (size_t)M * N * 4));
    CK(cudaMalloc(&dout_h, (size_t)M * N * 2));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dW, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dbias, hbias.data(), N * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dpos, hpos.data(), hpos.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dres, hres.data(), hres.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dout_f, hres.data(), hres.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dout_h, 0, (size_t)M * N * 2));

    ref_gemm<<<dim3((N + 127) / 128, M), 128>>>(dA, dW, w_fmt, dbias, dpos, 192, dres, dref, M, N, K, epi);
    CK(cudaDeviceSynchronize());

    CUtensorMap ta, tb, tout;
    if (!make_tmap_bf16_2d(&ta, dA, M, K, K, GEMM_BM) ||
        !make_tmap_bf16_2d(&tb, dW, N, K, K, bn / cg) || !make_tmap_f32_out(&tout, dout_f, M, N, N)) {
        printf("tensor map encode failed\n");
        return 2;
    }
    const bool f32_out = (epi == EPI_BIAS_RESID_F32 || epi == EPI_BIAS_POS_F32);
    GemmParams p{};
    p.M = M; p.N = N; p.K = K;
    p.bias = dbias;
    p.out = f32_out ? (void*)dout_f : (void*)dout_h;
    p.ldo = N;
    p.pos = dpos;
    p.pos_rows = 192;
    p.b_format = (uint32_t)w_fmt;
    p.ksplit = getenv("HM_NO_KSPLIT") ? -1 : 0;   // -1: keep whole tiles in the tail wave (A/B timing)
    GemmPlan plan = {cg, bn};
    CK(gemm_launch(plan, epi, ta, tb, tout, p, num_sms, 0));
    CK(cudaDeviceSynchronize());

    // compare
    std::vector<float> href((size_t)M * N), hout((size_t)M * N);
    CK(cudaMemcpy(href.data(), dref, href.size() * 4, cudaMemcpyDeviceToHost));
    if (f32_out) {
        CK(cudaMemcpy(hout.data(), dout_f, hout.size() * 4, cudaMemcpyDeviceToHost));
    } else {
        std::vector<__half> tmp((size_t)M * N);
        CK(cudaMemcpy(tmp.data(), dout_h, tmp.size() * 2, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < tmp.size(); ++i) hout[i] = __half2float(tmp[i]);
    }
    double max_err = 0, max_ref = 0;
    size_t bad = 0, first_bad = (size_t)-1;
    for (size_t i = 0; i < href.size(); ++i) {
        double e = fabs((double)href[i] - (double)hout[i]);
        double tol = f32_out ? 2e-3 : (2e-3 * fabs(href[i]) + 1e-3);
        if (e > tol || hout[i] != hout[i]) {
            ++bad;
            if (first_bad == (size_t)-1) first_bad = i;
        }
        if (e > max_err) max_err = e;
        if (fabs(href[i]) > max_ref) max_ref = fabs(href[i]);
    }
    if (bad) {
        printf("first mismatch at m=%zu n=%zu ref=%f got=%f\n", first_bad / N, first_bad % N,
               href[first_bad], hout[first_bad]);
        // print a coarse map of which 32x32 blocks are wrong (first 256 rows/cols)
        for (int bm = 0; bm < (M < 256 ? M : 256); bm += 32) {
            for (int bnn = 0; bnn < (N < 512 ? N : 512); bnn += 32) {
                int cnt = 0;
                for (int i = 0; i < 32 && bm + i < M; ++i)
                    for (int j = 0; j < 32; ++j) {
                        size_t idx = (size_t)(bm + i) * N + bnn + j;
                        double e = fabs((double)href[idx] - (double)hout[idx]);
                        if (e > 1e-2 * fabs(href[idx]) + 2e-3) ++cnt;
                    }
                printf("%c", cnt == 0 ? '.' : (cnt == 1024 ? '#' : 'x'));
            }
            printf("\n");
        }
    }

    // timing (residual epilogue keeps accumulating; values are irrelevant here)
    float ms = 0;
    if (iters > 0 && (!bad || getenv("HM_TIME_ANYWAY"))) {   // (HM_TIME_ANYWAY: timing experiments with deliberately wrong epilogues)
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        for (int i = 0; i < 3; ++i) CK(gemm_launch(plan, epi, ta, tb, tout, p, num_sms, 0));
        CK(cudaEventRecord(e0));
        for (int i = 0; i < iters; ++i) CK(gemm_launch(plan, epi, ta, tb, tout, p, num_sms, 0));
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= iters;
    }
    double tflops = ms > 0 ? 2.0 * M * N * K / (ms * 1e-3) / 1e12 : 0;
    printf("{\"test\":\"gemm\",\"cta_group\":%d,\"bn\":%d,\"epi\":%d,\"M\":%d,\"N\":%d,\"K\":%d,"
           "\"w_fmt\":%d,\"max_err\":%.3e,\"max_ref\":%.3e,\"bad\":%zu,\"ms\":%.4f,\"tflops\":%.1f,\"ok\":%s}\n",
           cg, bn, epi, M, N, K, w_fmt, max_err, max_ref, bad, ms, tflops, bad ? "false" : "true");
    return bad ? 1 : 0;
}
