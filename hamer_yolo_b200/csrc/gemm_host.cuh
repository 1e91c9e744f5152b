// Host side of the tcgen05 GEMM: TMA tensor-map creation and kernel launch.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <utility>

#include "gemm_tcgen05.cuh"

namespace hm {

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                        const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// libcuda is resolved at run time (the build box has no driver; the .so must still link).
inline PFN_tmapEncodeTiled get_tmap_encode() {
    static PFN_tmapEncodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
                cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
    }
    return fn;
}

// Row-major 16-bit (fp16 or bf16 - TMA only moves bytes) matrix [rows, cols] (cols contiguous, row
// pitch = ld elements); box = [box_rows, 64 cols] with the 128-byte swizzle.
inline bool make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                              uint64_t ld, uint32_t box_rows) {
    PFN_tmapEncodeTiled enc = get_tmap_encode();
    if (!enc) return false;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)GEMM_BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim,
                     gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// fp32 output matrix [rows, cols] for the TMA reduce-add epilogue: box = 32 x 32 floats
// (128-byte rows, 128B swizzle to match the epilogue's XOR-swizzled staging tile).
inline bool make_tmap_f32_out(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                              uint64_t ld) {
    PFN_tmapEncodeTiled enc = get_tmap_encode();
    if (!enc) return false;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld * 4};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// 2-D maps over the fp16 qkv buffer [rows, 3840] for the attention kernel: a 64-column box with the
// 128-byte swizzle and a 16-column box with the 32-byte swizzle (head_dim 80 = 64 + 16), 192 rows each.
inline bool make_tmap_attn(CUtensorMap* t128, CUtensorMap* t32, const void* base, uint64_t rows) {
    PFN_tmapEncodeTiled enc = get_tmap_encode();
    if (!enc) return false;
    cuuint64_t gdim[2] = {3840, rows};
    cuuint64_t gstride[1] = {3840 * 2};
    cuuint32_t estr[2] = {1, 1};
    cuuint32_t b128[2] = {64, 192}, b32[2] = {16, 192};
    return enc(t128, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, b128, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS &&
           enc(t32, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, b32, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Kernel launch, optionally with the programmatic-stream-serialization attribute (PDL): the launch and set-up of
// kernel n + 1 then overlap the tail of kernel n (every kernel waits on griddepcontrol.wait before touching
// dependent memory).  Measured inside the replayed CUDA graph (profiles/r02_notes.md): -3 to -5 % at 2-8 crops (a forward
// is ~280 short kernels there), nothing at 64 crops.  HM_PDL=1 / 0 forces it on / off; unset = automatic: on for calls
// of at most HM_PDL_MAX_B crops (default 24: -1.8 % there, neutral from 32 on), which the entry points announce through PdlScope.
inline int pdl_mode() {
    static const int m = [] { const char* e = getenv("HM_PDL"); return !e ? -1 : (e[0] == '1' ? 1 : 0); }();
    return m;
}
inline int pdl_max_batch() {
    static const int b = [] { const char* e = getenv("HM_PDL_MAX_B"); return e ? atoi(e) : 24; }();
    return b;
}
inline bool& pdl_auto_flag() { static thread_local bool f = false; return f; }
inline bool pdl_enabled() { const int m = pdl_mode(); return m == 1 || (m < 0 && pdl_auto_flag()); }
struct PdlScope {      // marks the launches of one API call as small-batch (automatic PDL)
    bool prev;
    explicit PdlScope(int batch) : prev(pdl_auto_flag()) { pdl_auto_flag() = batch <= pdl_max_batch(); }
    ~PdlScope() { pdl_auto_flag() = prev; }
};
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

struct GemmPlan {
    int cta_group;  // 1 or 2
    int bn;         // 128 or 256
};

template <int CTA_GROUP, int BN, int EPI>
inline cudaError_t gemm_launch_t(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tout,
                                 const GemmParams& p,
                                 int num_sms, cudaStream_t stream) {
    using Cfg = GemmCfg<CTA_GROUP, BN, EPI>;
    auto kern = gemm_bf16_tn_kernel<CTA_GROUP, BN, EPI>;
    // the attribute is per device: one flag per ordinal (several contexts on different GPUs in one process)
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    const int tile_m = GEMM_BM * CTA_GROUP;
    const int tiles = ((p.M + tile_m - 1) / tile_m) * (p.N / BN);
    const int max_workers = num_sms / CTA_GROUP;
    int workers = tiles < max_workers ? tiles : max_workers;
    GemmParams pp = p;
    pp.ksplit = 1;
    pp.ksplit_workers = max_workers;
    if (EPI == EPI_BIAS_RESID_F32 && p.ksplit >= 0 && tiles % max_workers != 0) {
        // K-slices for the tiles of the partial wave (all tiles when there are fewer tiles than workers): as many as
        // there are workers per such tile, at least 4 k-blocks each
        const int tail = tiles % max_workers;
        int ks = max_workers / tail;
        const int num_k = p.K / GEMM_BK;
        if (ks > num_k / 4) ks = num_k / 4;
        if (ks > 8) ks = 8;
        if (ks >= 2) {
            pp.ksplit = ks;
            const int items = (tiles - tail) + tail * ks;
            workers = items < max_workers ? items : max_workers;
        }
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(workers * CTA_GROUP);
    cfg.blockDim = dim3(Cfg::THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CTA_GROUP;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kern, ta, tb, tout, pp);
}

template <int CTA_GROUP, int BN>
inline cudaError_t gemm_launch_e(int epi, const CUtensorMap& ta, const CUtensorMap& tb,
                                 const CUtensorMap& tout, const GemmParams& p, int num_sms,
                                 cudaStream_t s) {
    switch (epi) {
        case EPI_BIAS_BF16: return gemm_launch_t<CTA_GROUP, BN, EPI_BIAS_BF16>(ta, tb, tout, p, num_sms, s);
        case EPI_BIAS_GELU_BF16: return gemm_launch_t<CTA_GROUP, BN, EPI_BIAS_GELU_BF16>(ta, tb, tout, p, num_sms, s);
        case EPI_BIAS_RESID_F32: return gemm_launch_t<CTA_GROUP, BN, EPI_BIAS_RESID_F32>(ta, tb, tout, p, num_sms, s);
        case EPI_BIAS_POS_F32: return gemm_launch_t<CTA_GROUP, BN, EPI_BIAS_POS_F32>(ta, tb, tout, p, num_sms, s);
    }
    if constexpr (CTA_GROUP == 1 && BN == 128) {      // the RootNet convolutions only use single-CTA 128 x 128 tiles
        if (epi == EPI_BIAS_RELU_F16) return gemm_launch_t<1, 128, EPI_BIAS_RELU_F16>(ta, tb, tout, p, num_sms, s);
        if (epi == EPI_BIAS_RESID_RELU_F16) return gemm_launch_t<1, 128, EPI_BIAS_RESID_RELU_F16>(ta, tb, tout, p, num_sms, s);
    }
    return cudaErrorInvalidValue;
}

// The W tensor map's box rows must equal BN / cta_group (see GemmCfg::B_ROWS).  `tout` is the fp32
// output map of the residual epilogue (make_tmap_f32_out); other epilogues ignore it (pass ta).
inline cudaError_t gemm_launch(GemmPlan plan, int epi, const CUtensorMap& ta, const CUtensorMap& tb,
                               const CUtensorMap& tout, const GemmParams& p, int num_sms,
                               cudaStream_t s) {
    if (p.K % GEMM_BK != 0 || p.N % plan.bn != 0) return cudaErrorInvalidValue;
    if (plan.cta_group == 1 && plan.bn == 256) return gemm_launch_e<1, 256>(epi, ta, tb, tout, p, num_sms, s);
    if (plan.cta_group == 1 && plan.bn == 128) return gemm_launch_e<1, 128>(epi, ta, tb, tout, p, num_sms, s);
    if (plan.cta_group == 2 && plan.bn == 256) return gemm_launch_e<2, 256>(epi, ta, tb, tout, p, num_sms, s);
    if (plan.cta_group == 2 && plan.bn == 128) return gemm_launch_e<2, 128>(epi, ta, tb, tout, p, num_sms, s);
    return cudaErrorInvalidValue;
}

}  // namespace hm
