// libhamer_b200.so - C ABI implementation (see include/hamer_b200.h for the contract and the
// reference interfaces each entry point replaces).
#include "../../include/hamer_b200.h"

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <array>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "comm.cuh"
#include "attention.cuh"
#include "attention_tc5.cuh"
#include "crop_warp.cuh"
#include "decoder_head.cuh"
#include "elementwise.cuh"
#include "gemm_host.cuh"
#include "letterbox.cuh"
#include "mano_lbs.cuh"
#include "mano_fused.cuh"
#include "rootnet.cuh"

using namespace hm;

namespace {

constexpr int TOK = 192;       // tokens per crop (16 x 12 patches)
constexpr int DM = 1280;       // backbone width
constexpr int DFF = 5120;
constexpr int DEPTH = 32;
constexpr int KPATCH = 1536;   // 2 x 768: [hi | lo] split of the fp32 pixels
constexpr int HD = 1024;       // decoder width
constexpr int HDEPTH = 6;
constexpr int NREAD = 109;     // 96 pose6d + 10 betas + 3 cam
constexpr int MANO_CHUNK = 8192;  // legacy 3-kernel path: hands per pass (80 MB of fp32 v_posed stays inside the 126 MB L2)

std::string g_create_error;

struct DevTensor {
    float* f32 = nullptr;
    __nv_bfloat16* bf16 = nullptr;
    std::vector<int64_t> shape;
    size_t numel = 0;
};

// raw pointers cached by hm_finalize_weights so the launch path does no name lookups
struct BlockW {
    const float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *b_qkv, *b_proj, *b_fc1, *b_fc2;
    CUtensorMap tm_qkv, tm_proj, tm_fc1, tm_fc2;
};
struct HeadLayerW {
    const float *n0g, *n0b, *n1g, *n1b, *n2g, *n2b;
    const float *w_sa, *b_sa_o;                       // single-token self-attention: to_out . W_v merged (1024 x 1024), to_out bias
    const float *w_q, *w_k, *w_v, *w_ca_o, *b_ca_o;   // to_q, to_kv rows 0..511 (k) and 512.. (v), to_out
    const float *w1, *b1, *w2, *b2;                   // feed-forward
};

// fp32 -> 16-bit weight operand (fmt 1 = bf16, 0 = fp16), round to nearest even; fp16 saturates at +-65,504
// instead of overflowing to inf.
// split != 0 (hm_config.split_weights, fp16 only): in is [rows, K], out is [rows, 2K] = [hi | lo] with hi = fp16(w),
// lo = fp16(w - hi): hi + lo carries ~21 mantissa bits of w, the GEMM runs over K' = 2K re-reading A
// (GemmParams::a_kwrap).
__global__ void f32_to_w16_kernel(const float* __restrict__ in, uint16_t* __restrict__ out, size_t n,
                                  int fmt, int split, int K) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (fmt == 1) {
        const __nv_bfloat16 v = __float2bfloat16_rn(in[i]);
        out[i] = *reinterpret_cast<const uint16_t*>(&v);
    } else if (!split) {
        out[i] = ptx::half_bits_sat(in[i]);
    } else {
        const size_t r = i / (size_t)K, k = i - r * (size_t)K;
        const uint16_t hi = ptx::half_bits_sat(in[i]);
        const float hf = __half2float(*reinterpret_cast<const __half*>(&hi));
        out[r * 2 * (size_t)K + k] = hi;
        out[r * 2 * (size_t)K + K + k] = ptx::half_bits_sat(in[i] - hf);
    }
}

}  // namespace

struct hm_ctx {
    int device = 0;
    int num_sms = 0;
    hm_config cfg{};
    std::string err;
    GemmPlan plan{2, 256};
    bool deterministic = false;
    bool split_w = false;        // hm_config.split_weights: backbone weights as fp16 hi + lo pairs, K doubled
    int small_m = 384;           // rows at or below which the backbone GEMMs use single-CTA 128 x 128 tiles (HM_SMALL_M)

    std::map<std::string, DevTensor> w;   // every state-dict entry, on the device
    bool finalized = false;
    bool mano_set = false;

    // fused tables built by hm_finalize_weights
    __nv_bfloat16* w_patch2 = nullptr;    // [1280, 1536] = [W | W]
    float* pos_tab = nullptr;             // [192, 1280] = pos[1:] + pos[0]
    float* tok0 = nullptr;                // [1024] layer-0 decoder token
    unsigned int* hticket = nullptr;      // per-crop arrival counter of head_xattn_kernel (zero between launches)
    float* w_sa[HDEPTH] = {};             // [1024][1024] = to_out.weight (1024 x 512) . to_qkv.weight[1024:1536] (512 x 1024)
    float* w_read = nullptr;              // [109, 1024]
    float* b_read = nullptr;              // [109] bias + init_*
    CUtensorMap tm_wpatch;
    BlockW blk[DEPTH];
    HeadLayerW hl[HDEPTH];
    const float *b_patch = nullptr, *lnf_g = nullptr, *lnf_b = nullptr;

    // workspace (max_batch)
    __half *patches = nullptr, *xn = nullptr, *qkv = nullptr, *att = nullptr, *hbuf = nullptr;  // fp16 activations
    uint32_t b_format = 0;   // operand format of the backbone GEMMs (A and B): 0 = fp16, 1 = bf16
    bool att_legacy = false; // true: mma.sync attention kernel instead of the tcgen05 one
    float *x = nullptr, *feats = nullptr;
    CUtensorMap tm_patches, tm_xn, tm_att, tm_h;
    CUtensorMap tm_x_out;   // fp32 residual stream as the target of the TMA reduce-add epilogue
    CUtensorMap tm_att128, tm_att32;   // qkv buffer: 64- and 16-column boxes for the attention kernel
    // decoder workspace
    float *hx = nullptr, *hx2 = nullptr, *hq = nullptr, *hu = nullptr, *hpart = nullptr, *hml = nullptr,
          *hpool = nullptr, *ho = nullptr, *hh = nullptr, *hraw = nullptr, *hrot = nullptr,
          *hcamt = nullptr, *hbetas = nullptr, *hscratch = nullptr;

    // MANO
    ManoDevice mano{};
    std::vector<void*> mano_allocs;
    __half* m_coef = nullptr;     // [MANO_CHUNK, 320]
    float *m_vposed = nullptr, *m_A = nullptr, *m_jpos = nullptr;
    CUtensorMap tm_mcoef, tm_mbasis;
    CUtensorMap tm_fbasis;
    ManoChainConst mano_chain{};  // folded joint regressor + output slots: constant-bank parameter of the fused kernel
    bool mano_fused_ok = false;   // the fused kernel applies: joint_map is a permutation (every source joint has exactly one
                                  // slot), the kinematic tree is the standard MANO one, no vertex group needs > 16 joints
    bool mano_legacy = false;
    long long* mano_clk = nullptr;   // phase-clock buffer of the fused kernel (hm_test_mano_clocks), null in normal use

    // preprocess
    uint8_t* frame_dev = nullptr;
    std::vector<std::array<double, 4>> crop_win;   // per crop: source window x0, x1, y0, y1 (host frames: what gets uploaded)
    std::vector<std::array<long, 4>> crop_rect;
    CropParams* crop_host = nullptr;   // pinned
    CropParams* crop_dev = nullptr;
    BlurParams* blur_host = nullptr;   // pinned; anti-aliased crops (hm_preprocess_blur)
    BlurParams* blur_dev = nullptr;
    float lut_mean[3] = {}, lut_sd[3] = {};   // float32 normalisation constants (RGB), the ones the LUT is built from
    float* lut = nullptr;              // [3][256]
    cudaEvent_t crop_event = nullptr;
    bool crop_event_pending = false;
    cudaEvent_t crop_kernel_event = nullptr;   // the last crop kernel has finished reading crop_dev / frame_dev
    bool crop_kernel_recorded = false;

    std::vector<void*> allocs;

    // RootNet / SAR depth (hm_rootnet_*): BatchNorm-folded fp16 convolution weights, workspace for RN_CHUNK hands
    struct RnConv { int cin, cout, k, stride, pad, kpad, npad, hin, hout; __half* w; float* bias; CUtensorMap tm_w, tm_a; };
    RnConv rn[RN_NCONV];
    bool rn_ready = false;
    std::vector<void*> rn_allocs;
    float *rn_img = nullptr, *rn_lut = nullptr, *rn_wd = nullptr, *rn_bd = nullptr, *rn_kval = nullptr;
    __half *rn_col = nullptr, *rn_act[4] = {};

    // the context's own NCCL communicator (hm_comm_init), one rank per context / GPU
    ncclComm_t comm = nullptr;
    int comm_ranks = 0;

    // optional per-category timing (hm_profile_*): CUDA events recorded on the launch stream
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_pool;
    size_t prof_used = 0;
    struct ProfSpan { int cat; cudaEvent_t a, b; };
    std::vector<ProfSpan> prof_spans;
};

// RAII bracket around one launch (or one group of launches) of category `cat`
struct ProfScope {
    hm_ctx* c; cudaStream_t s; int cat; cudaEvent_t a = nullptr, b = nullptr;
    static cudaEvent_t take(hm_ctx* c) {
        if (c->prof_used == c->prof_pool.size()) {
            cudaEvent_t e;
            if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
            c->prof_pool.push_back(e);
        }
        return c->prof_pool[c->prof_used++];
    }
    ProfScope(hm_ctx* c_, cudaStream_t s_, int cat_) : c(c_), s(s_), cat(cat_) {
        if (!c->prof_on) return;
        a = take(c); b = take(c);
        if (a) cudaEventRecord(a, s);
    }
    ~ProfScope() {
        if (!c->prof_on || !a || !b) return;
        cudaEventRecord(b, s);
        c->prof_spans.push_back({cat, a, b});
    }
};

namespace {

// profiling categories (order documented in include/hamer_b200.h: hm_profile_collect)
enum ProfCat : int { PC_IM2COL = 0, PC_GEMM_PATCH, PC_LAYERNORM, PC_GEMM_QKV, PC_ATTENTION, PC_GEMM_PROJ,
                     PC_GEMM_FC1, PC_GEMM_FC2, PC_HEAD, PC_MANO_LBS, PC_CROP, PC_COUNT };

int fail(hm_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}

#define HM_CUDA(ctx, call)                                                                   \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess)                                                               \
            return fail(ctx, HM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

template <typename T>
cudaError_t dev_alloc(hm_ctx* c, T** p, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T));
    if (e == cudaSuccess) {
        c->allocs.push_back(q);
        *p = reinterpret_cast<T*>(q);
    }
    return e;
}

const DevTensor* find(const hm_ctx* c, const std::string& name) {
    auto it = c->w.find(name);
    return it == c->w.end() ? nullptr : &it->second;
}

bool shape_is(const DevTensor* t, std::initializer_list<int64_t> s) {
    if (!t || t->shape.size() != s.size()) return false;
    size_t i = 0;
    for (int64_t v : s) if (t->shape[i++] != v) return false;
    return true;
}

std::string blk_name(int i, const char* tail) {
    return "backbone.blocks." + std::to_string(i) + "." + tail;
}
std::string dec_name(int l, int sub, const char* tail) {
    return "mano_head.transformer.transformer.layers." + std::to_string(l) + "." +
           std::to_string(sub) + "." + tail;
}

// which state-dict matrices feed the bf16 tensor-core GEMMs
bool is_backbone_matrix(const std::string& n) {
    if (n.rfind("backbone.blocks.", 0) != 0) return false;
    const char* tails[] = {"attn.qkv.weight", "attn.proj.weight", "mlp.fc1.weight", "mlp.fc2.weight"};
    for (const char* t : tails) {
        size_t L = strlen(t);
        if (n.size() > L && n.compare(n.size() - L, L, t) == 0) return true;
    }
    return false;
}

// ---- cv2.getAffineTransform for the reference's three axis-aligned point pairs ------------------
// gen_trans_from_patch_cv (hamer/hamer/datasets/utils.py:82-129) with rot = 0, scale = 1; the 6x6
// system is solved exactly like OpenCV's cv::solve(DECOMP_LU) (partial pivoting, plain double ops,
// this file is compiled with -ffp-contract=off) so that M matches cv2 bit for bit.
void lu_solve6(double A[6][6], double b[6]) {
    const int m = 6;
    for (int i = 0; i < m; ++i) {
        int k = i;
        for (int j = i + 1; j < m; ++j)
            if (fabs(A[j][i]) > fabs(A[k][i])) k = j;
        if (k != i) {
            for (int j = i; j < m; ++j) { double t = A[i][j]; A[i][j] = A[k][j]; A[k][j] = t; }
            double t = b[i]; b[i] = b[k]; b[k] = t;
        }
        const double d = -1 / A[i][i];
        for (int j = i + 1; j < m; ++j) {
            const double alpha = A[j][i] * d;
            for (int c = i + 1; c < m; ++c) A[j][c] += alpha * A[i][c];
            b[j] += alpha * b[i];
        }
    }
    for (int i = m - 1; i >= 0; --i) {
        double s = b[i];
        for (int c = i + 1; c < m; ++c) s -= A[i][c] * b[c];
        b[i] = s / A[i][i];
    }
}

// gen_trans_from_patch_cv for a square window of `size` pixels around (cx, cy): 256 x 256 patch, rot = 0, scale = 1
void center_to_affine(double cx, double cy, double size, double M[6]) {
    const float half = (float)(size * 0.5);
    const float s0x = (float)cx, s0y = (float)cy;
    const float s1x = (float)(cx + 0.0), s1y = (float)(cy + (double)half);
    const float s2x = (float)(cx + (double)half), s2y = (float)(cy + 0.0);
    const double sx[3] = {s0x, s1x, s2x}, sy[3] = {s0y, s1y, s2y};
    const double dx[3] = {128, 128, 256}, dy[3] = {128, 256, 128};
    double A[6][6] = {}, b[6];
    for (int i = 0; i < 3; ++i) {
        A[2 * i][0] = sx[i]; A[2 * i][1] = sy[i]; A[2 * i][2] = 1;
        A[2 * i + 1][3] = sx[i]; A[2 * i + 1][4] = sy[i]; A[2 * i + 1][5] = 1;
        b[2 * i] = dx[i]; b[2 * i + 1] = dy[i];
    }
    lu_solve6(A, b);
    for (int i = 0; i < 6; ++i) M[i] = b[i];
}

// box -> centre and side of the square crop window (infer.py:182-199, datasets/utils.py:15-34 expand_to_aspect_ratio)
void box_to_center_size(const double* box, double* cx_out, double* cy_out, double* size_out) {
    const double x1 = box[0], y1 = box[1], x2 = box[2], y2 = box[3];
    const double cx = (x1 + x2) / 2.0, cy = (y1 + y2) / 2.0;               // infer.py:182-183
    const double w = (2.5 * (x2 - x1) / 200.0) * 200, h = (2.5 * (y2 - y1) / 200.0) * 200;  // :193-199
    double w_new, h_new;                                                    // expand_to_aspect_ratio
    if (h / w < 256.0 / 192.0) { h_new = fmax(w * 256 / 192, h); w_new = w; }
    else { h_new = h; w_new = fmax(h * 192 / 256, w); }
    *cx_out = cx; *cy_out = cy; *size_out = fmax(w_new, h_new);
}

// cv::warpAffine's inversion of M (no WARP_INVERSE_MAP)
void invert_affine(const double M[6], double out[6]) {
    double m[6];
    for (int i = 0; i < 6; ++i) m[i] = M[i];
    double D = m[0] * m[4] - m[1] * m[3];
    D = D != 0 ? 1. / D : 0;
    const double A11 = m[4] * D, A22 = m[0] * D;
    m[0] = A11; m[1] *= -D; m[3] *= -D; m[4] = A22;
    const double b1 = -m[0] * m[2] - m[1] * m[5];
    const double b2 = -m[3] * m[2] - m[4] * m[5];
    m[2] = b1; m[5] = b2;
    for (int i = 0; i < 6; ++i) out[i] = m[i];
}

// ---- launch helpers ---------------------------------------------------------------------------
template <int V, typename OutT>
void launch_ln(const float* x, const float* g, const float* b, OutT* y, int rows, float eps,
               cudaStream_t s) {
    if (rows <= 24 * TOK) launch_pdl(layernorm_kernel<V, OutT, true>, dim3((rows + 7) / 8), dim3(256), 0, s, x, g, b, y, rows, eps);
    else launch_pdl(layernorm_kernel<V, OutT>, dim3((rows + 7) / 8), dim3(256), 0, s, x, g, b, y, rows, eps);
}

template <int CPW, bool LNF, int RT = SK_ROWS>
void launch_skinny_epi(int epi, dim3 grid, size_t smem, cudaStream_t s, const SkinnyParams& p) {
    if (epi == SK_NONE) launch_pdl(skinny_linear_kernel<CPW, SK_NONE, LNF, RT>, grid, dim3(SK_WARPS * 32), smem, s, p);
    else if (epi == SK_GELU) launch_pdl(skinny_linear_kernel<CPW, SK_GELU, LNF, RT>, grid, dim3(SK_WARPS * 32), smem, s, p);
    else launch_pdl(skinny_linear_kernel<CPW, SK_RESID, LNF, RT>, grid, dim3(SK_WARPS * 32), smem, s, p);
}

// ln_g != nullptr: the input rows are LayerNorm-ed (eps 1e-5, the decoder's PreNorm) inside the kernel
cudaError_t launch_skinny(int num_sms, int epi, const float* X, int ldx, long long x_bs, const float* W,
                          long long w_bs, const float* bias, int bias_bs, float* Y, int ldy,
                          long long y_bs, int B, int N, int K, int batch, float scale,
                          const float* ln_g, const float* ln_b, const float* resid, cudaStream_t s) {
    SkinnyParams p;
    p.X = X; p.ldx = ldx; p.x_bs = x_bs; p.W = W; p.w_bs = w_bs; p.bias = bias; p.bias_bs = bias_bs;
    p.Y = Y; p.ldy = ldy; p.y_bs = y_bs; p.B = B; p.N = N; p.K = K; p.scale = scale;
    p.ln_g = ln_g; p.ln_b = ln_b; p.ln_eps = 1e-5f;
    p.R = resid; p.ldr = ldy;
    if (epi == SK_RESID && (!resid || batch != 1)) return cudaErrorInvalidValue;
    if (K % 128 != 0 || K > SK_KMAX || ldx % 4 != 0) return cudaErrorInvalidValue;
    // columns per warp: the widest tile that still gives (nearly) every SM a block
    const int row_tiles = (B + SK_ROWS - 1) / SK_ROWS;
    int cpw = 4;
    while (cpw > 1 && (((N + SK_WARPS * cpw - 1) / (SK_WARPS * cpw)) * row_tiles * batch < (num_sms * 3) / 4 ||
                       sk_smem_bytes(K, cpw) > SK_SMEM_MAX)) cpw >>= 1;
    dim3 grid((N + SK_WARPS * cpw - 1) / (SK_WARPS * cpw), row_tiles, batch);
    const size_t smem = sk_smem_bytes(K, cpw);
    const bool lnf = ln_g != nullptr;
    if (cpw == 4) { if (lnf) launch_skinny_epi<4, true>(epi, grid, smem, s, p); else launch_skinny_epi<4, false>(epi, grid, smem, s, p); }
    else if (cpw == 2) { if (lnf) launch_skinny_epi<2, true>(epi, grid, smem, s, p); else launch_skinny_epi<2, false>(epi, grid, smem, s, p); }
    else if (B <= 4) { if (lnf) launch_skinny_epi<1, true, 4>(epi, grid, smem, s, p); else launch_skinny_epi<1, false, 4>(epi, grid, smem, s, p); }
    else if (B <= 8) { if (lnf) launch_skinny_epi<1, true, 8>(epi, grid, smem, s, p); else launch_skinny_epi<1, false, 8>(epi, grid, smem, s, p); }
    else { if (lnf) launch_skinny_epi<1, true>(epi, grid, smem, s, p); else launch_skinny_epi<1, false>(epi, grid, smem, s, p); }
    return cudaGetLastError();
}

int run_gemm(hm_ctx* c, int epi, const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K,
             const float* bias, void* out, const float* pos, cudaStream_t s) {
    GemmParams p{};
    p.M = M; p.N = N; p.K = K; p.bias = bias; p.out = out; p.ldo = N; p.pos = pos; p.pos_rows = TOK;
    if (c->split_w) { p.K = 2 * K; p.a_kwrap = K / GEMM_BK; }   // W = [hi | lo]: second pass over A against the low halves
    p.b_format = c->b_format;
    p.ksplit = c->deterministic ? -1 : 0;   // -1: never split K (fixed fp32 summation order, batch-invariant results)
    // Few rows (video regime, a handful of hands per frame): 256 x 256 CTA-pair tiles leave most SMs idle and every
    // tile still pays 256-row MMAs.  Single-CTA 128 x 128 tiles give 4x the tiles at a quarter of the latency; the
    // weight tensor maps are shared (128-row boxes in both plans).
    GemmPlan plan = c->plan;
    if (plan.cta_group == 2 && plan.bn == 256 && M <= c->small_m) plan = GemmPlan{1, 128};
    HM_CUDA(c, gemm_launch(plan, epi, ta, tb, c->tm_x_out, p, c->num_sms, s));
    return HM_OK;
}

int backbone_impl(hm_ctx* c, const float* img, int B, float* feats_out, cudaStream_t s) {
    const int M = B * TOK;
    const float att_scale = 0.11180339887498948f * 1.4426950408889634f;   // 80^-0.5 * log2(e)
    int rc;
    { ProfScope ps(c, s, PC_IM2COL); launch_pdl(im2col_patch16_kernel, dim3(16, B), dim3(256), 0, s, img, c->patches, B); }
    { ProfScope ps(c, s, PC_GEMM_PATCH);
      if ((rc = run_gemm(c, EPI_BIAS_POS_F32, c->tm_patches, c->tm_wpatch, M, DM, KPATCH, c->b_patch, c->x, c->pos_tab, s))) return rc; }
    for (int i = 0; i < DEPTH; ++i) {
        const BlockW& w = c->blk[i];
        { ProfScope ps(c, s, PC_LAYERNORM); launch_ln<10, __half>(c->x, w.ln1_g, w.ln1_b, c->xn, M, 1e-6f, s); }
        { ProfScope ps(c, s, PC_GEMM_QKV);
          if ((rc = run_gemm(c, EPI_BIAS_BF16, c->tm_xn, w.tm_qkv, M, 3 * DM, DM, w.b_qkv, c->qkv, nullptr, s))) return rc; }
        { ProfScope ps(c, s, PC_ATTENTION);
          if (c->att_legacy) launch_pdl(vit_attention_kernel, dim3(16, B), dim3(ATT_THREADS), ATT_SMEM_BYTES, s, c->qkv, c->att, att_scale);
          else launch_pdl(vit_attention_tc5_kernel<false>, dim3(16, B, (32 * B <= 2 * c->num_sms) ? 2 : 1), dim3(AT5_THREADS),
                          AT5_SMEM_BYTES, s, c->tm_att128, c->tm_att32, c->att, att_scale, nullptr); }
        { ProfScope ps(c, s, PC_GEMM_PROJ);
          if ((rc = run_gemm(c, EPI_BIAS_RESID_F32, c->tm_att, w.tm_proj, M, DM, DM, w.b_proj, c->x, nullptr, s))) return rc; }
        { ProfScope ps(c, s, PC_LAYERNORM); launch_ln<10, __half>(c->x, w.ln2_g, w.ln2_b, c->xn, M, 1e-6f, s); }
        { ProfScope ps(c, s, PC_GEMM_FC1);
          if ((rc = run_gemm(c, EPI_BIAS_GELU_BF16, c->tm_xn, w.tm_fc1, M, DFF, DM, w.b_fc1, c->hbuf, nullptr, s))) return rc; }
        { ProfScope ps(c, s, PC_GEMM_FC2);
          if ((rc = run_gemm(c, EPI_BIAS_RESID_F32, c->tm_h, w.tm_fc2, M, DM, DFF, w.b_fc2, c->x, nullptr, s))) return rc; }
    }
    { ProfScope ps(c, s, PC_LAYERNORM); launch_ln<10, float>(c->x, c->lnf_g, c->lnf_b, feats_out, M, 1e-6f, s); }
    HM_CUDA(c, cudaGetLastError());
    return HM_OK;
}

int head_impl(hm_ctx* c, const float* feats, int B, const hm_outputs* out, cudaStream_t s) {
    ProfScope ps(c, s, PC_HEAD);
    launch_pdl(head_init_token_kernel, dim3((B * HD + 255) / 256), dim3(256), 0, s, c->tok0, c->hx, B);
    const float scale = 0.125f;   // dim_head ** -0.5 = 64^-0.5
    const int nsm = c->num_sms;
    const int row_tiles = (B + SK_ROWS - 1) / SK_ROWS;
    float *xa = c->hx, *xb = c->hx2;   // token before / after the layer's self-attention (swapped per layer)
    for (int l = 0; l < HDEPTH; ++l) {
        const HeadLayerW& w = c->hl[l];
        // self-attention on one token: xb = xa + (W_o W_v) LN(xa) + b_o  (out of place: every block normalises whole rows of xa)
        HM_CUDA(c, launch_skinny(nsm, SK_RESID, xa, HD, 0, w.w_sa, 0, w.b_sa_o, 0, xb, HD, 0, B, HD, HD, 1, 1.f, w.n0g, w.n0b, xa, s));
        // cross-attention against the 192 context tokens: q = W_q LN(x), u[b,h,:] = scale * W_k,h^T q[b,h,:]
        HM_CUDA(c, launch_skinny(nsm, SK_NONE, xb, HD, 0, w.w_q, 0, nullptr, 0, c->hq, 512, 0, B, 512, HD, 1, 1.f, w.n1g, w.n1b, nullptr, s));
        launch_pdl(head_u_kernel, dim3(DM / 128, 8, row_tiles), dim3(128), 0, s, (const float*)c->hq, w.w_k, c->hu, B, scale);
        {
            const int xg = std::min(nsm, XA_NCH * B);
            if (c->deterministic) launch_pdl(head_xattn_kernel<true>, dim3(xg), dim3(XA_THREADS), XA_SMEM_BYTES, s, feats, (const float*)c->hu, c->hpart, c->hml, c->hticket, c->hpool, B);
            else launch_pdl(head_xattn_kernel<false>, dim3(xg), dim3(XA_THREADS), XA_SMEM_BYTES, s, feats, (const float*)c->hu, c->hpart, c->hml, c->hticket, c->hpool, B);
        }
        // o[b, h*64 + d] = W_v,h pooled[b,h,:]
        HM_CUDA(c, launch_skinny(nsm, SK_NONE, c->hpool, 8 * DM, DM, w.w_v, (long long)64 * DM, nullptr, 0,
                                 c->ho, 512, 64, B, 64, DM, 8, 1.f, nullptr, nullptr, nullptr, s));
        HM_CUDA(c, launch_skinny(nsm, SK_RESID, c->ho, 512, 0, w.w_ca_o, 0, w.b_ca_o, 0, xb, HD, 0, B, HD, 512, 1, 1.f, nullptr, nullptr, xb, s));
        // feed-forward
        HM_CUDA(c, launch_skinny(nsm, SK_GELU, xb, HD, 0, w.w1, 0, w.b1, 0, c->hh, HD, 0, B, HD, HD, 1, 1.f, w.n2g, w.n2b, nullptr, s));
        HM_CUDA(c, launch_skinny(nsm, SK_RESID, c->hh, HD, 0, w.w2, 0, w.b2, 0, xb, HD, 0, B, HD, HD, 1, 1.f, nullptr, nullptr, xb, s));
        std::swap(xa, xb);
    }
    HM_CUDA(c, launch_skinny(nsm, SK_NONE, xa, HD, 0, c->w_read, 0, c->b_read, 0, c->hraw, NREAD, 0, B,
                             NREAD, HD, 1, 1.f, nullptr, nullptr, nullptr, s));
    HeadOut ho;
    // outputs the caller skipped are routed into scratch so the kernel stays branch-free
    float* scratch = c->hscratch;
    ho.global_orient = out && out->global_orient ? out->global_orient : scratch;
    ho.hand_pose = out && out->hand_pose ? out->hand_pose : scratch + (size_t)B * 9;
    ho.betas = c->hbetas;
    ho.pred_cam = out && out->pred_cam ? out->pred_cam : scratch + (size_t)B * 144;
    ho.pred_cam_t = c->hcamt;
    ho.focal_length = out && out->focal_length ? out->focal_length : scratch + (size_t)B * 147;
    ho.rotmats = c->hrot;
    launch_pdl(head_finalize_kernel, dim3((B * 16 + 127) / 128), dim3(128), 0, s, c->hraw, ho, B, 5000.0f, 256.0f);
    HM_CUDA(c, cudaGetLastError());
    if (out && out->betas)
        HM_CUDA(c, cudaMemcpyAsync(out->betas, c->hbetas, (size_t)B * 10 * 4, cudaMemcpyDeviceToDevice, s));
    if (out && out->pred_cam_t)
        HM_CUDA(c, cudaMemcpyAsync(out->pred_cam_t, c->hcamt, (size_t)B * 3 * 4, cudaMemcpyDeviceToDevice, s));
    return HM_OK;
}

// Work list of the fused kernel: the pair tiles (256 hands) of the full waves are one unit each; the tiles of the partial
// last wave are cut into n_split pieces (each repeats the tile's prologue, cost ~4 coordinate tiles) so that the wave
// ends together.  With fewer tiles than workers every tile is split (small batches).
void mano_plan(int pair_tiles, int workers, int* whole, int* n_split) {
    static const int forced = [] { const char* e = getenv("HM_MANO_SPLIT"); return e ? atoi(e) : 0; }();
    const int tail = pair_tiles % workers;
    *whole = pair_tiles - tail;
    *n_split = 1;
    if (tail == 0) return;
    if (forced >= 1 && forced <= MF_NT) { *n_split = forced; return; }
    double best_cost = 1e30;
    for (int ns = 1; ns <= MF_NT; ++ns) {
        const int waves = (tail * ns + workers - 1) / workers;
        const double cost = waves * ((MF_NT + ns - 1) / ns + 4.0);
        if (cost < best_cost - 1e-9) { best_cost = cost; *n_split = ns; }
    }
}

int lbs_fused_impl(hm_ctx* c, const float* betas, const float* rotmats, const float* cam_t, int B,
                   float* verts, float* joints, float* kp2d, cudaStream_t s) {
    const bool proj = cam_t != nullptr && kp2d != nullptr;
    ManoFusedArgs a{};
    a.betas = betas; a.rotmats = rotmats; a.cam_t = proj ? cam_t : nullptr;
    a.verts = verts; a.joints = joints; a.kp2d = proj ? kp2d : nullptr;
    a.focal_over_size = 5000.0f / 256.0f;
    a.B = B;
    const int pair_tiles = (B + 2 * MF_BM - 1) / (2 * MF_BM);
    const int workers_max = c->num_sms / 2;
    mano_plan(pair_tiles, workers_max, &a.whole, &a.n_split);
    static const int mano_dbg = [] { const char* e = getenv("HM_MANO_DBG"); return e ? atoi(e) : 0; }();
    a.dbg = mano_dbg;   // kernel ablation switches for profiling (see ManoFusedArgs::dbg); 0 in normal use
    a.clk = c->mano_clk;
    const int units = a.whole + (pair_tiles - a.whole) * a.n_split;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * (units < workers_max ? units : workers_max));
    cfg.blockDim = dim3(MF_THREADS);
    cfg.dynamicSmemBytes = MF_SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    HM_CUDA(c, cudaLaunchKernelEx(&cfg, mano_fused_kernel, c->tm_fbasis, c->mano_chain, c->mano.fz_tab, a));
    HM_CUDA(c, cudaGetLastError());
    return HM_OK;
}

int lbs_impl(hm_ctx* c, const float* betas, const float* rotmats, const float* cam_t, int B,
             float* verts, float* joints, float* kp2d, cudaStream_t s) {
    ProfScope ps(c, s, PC_MANO_LBS);
    // the fused kernel fetches betas / rotations with 16-byte granular bulk copies and stores vertices 8 bytes per lane
    if (c->mano_fused_ok && !c->mano_legacy && (reinterpret_cast<uintptr_t>(verts) & 7u) == 0 &&
        (reinterpret_cast<uintptr_t>(rotmats) & 15u) == 0 && (reinterpret_cast<uintptr_t>(betas) & 15u) == 0)
        return lbs_fused_impl(c, betas, rotmats, cam_t, B, verts, joints, kp2d, s);
    for (int lo = 0; lo < B; lo += MANO_CHUNK) {                 // chunk: v_posed stays L2-resident
        const int n = (B - lo < MANO_CHUNK) ? (B - lo) : MANO_CHUNK;
        ManoPrepArgs pa{};
        pa.betas = betas + (size_t)lo * 10; pa.rotmats = rotmats + (size_t)lo * 144;
        pa.coef = c->m_coef; pa.A = c->m_A; pa.jpos = c->m_jpos; pa.B = n;
        launch_pdl(mano_prep_kernel, dim3((n + 7) / 8), dim3(128), 0, s, c->mano, pa);
        GemmParams p{};
        p.M = n; p.N = MANO_NPAD; p.K = MANO_KTOT; p.bias = c->mano.template_n; p.out = c->m_vposed;
        p.ldo = MANO_NPAD; p.pos = c->mano.zero_row; p.pos_rows = 1; p.b_format = 0;
        GemmPlan plan{c->plan.cta_group, 128};
        HM_CUDA(c, gemm_launch(plan, EPI_BIAS_POS_F32, c->tm_mcoef, c->tm_mbasis, c->tm_mcoef, p, c->num_sms, s));
        ManoSkinArgs a;
        a.vposed = c->m_vposed; a.A = c->m_A; a.jpos = c->m_jpos;
        a.cam_t = cam_t ? cam_t + (size_t)lo * 3 : nullptr;
        a.verts = verts + (size_t)lo * MANO_V * 3; a.joints = joints + (size_t)lo * 63;
        a.kp2d = (cam_t != nullptr && kp2d != nullptr) ? kp2d + (size_t)lo * 42 : nullptr;
        a.focal_over_size = 5000.0f / 256.0f;
        a.B = n;
        launch_pdl(mano_skin_kernel, dim3((n + SKIN_HANDS - 1) / SKIN_HANDS), dim3(MANO_VTHREADS), sizeof(SkinSmem), s, c->mano, a);
    }
    HM_CUDA(c, cudaGetLastError());
    return HM_OK;
}

int check_ready(hm_ctx* c, int B, bool need_weights, bool need_mano) {
    if (!c) return HM_ERR_INVALID;
    if (B <= 0) return fail(c, HM_ERR_INVALID, "batch size must be positive");
    if (need_weights && !c->finalized) return fail(c, HM_ERR_STATE, "weights not finalized (hm_finalize_weights)");
    if (need_weights && B > c->cfg.max_batch)
        return fail(c, HM_ERR_INVALID, "B = " + std::to_string(B) + " exceeds hm_config.max_batch = " +
                                           std::to_string(c->cfg.max_batch));
    if (need_mano && !c->mano_set) return fail(c, HM_ERR_STATE, "MANO tables not set (hm_set_mano)");
    return HM_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

const char* hm_version(void) { return "hamer_b200 0.1 (sm_100a; tcgen05/TMA GEMM, fp32 head, fused MANO LBS)"; }

const char* hm_last_error(const hm_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int hm_create(hm_ctx** out, int device_ordinal, const hm_config* cfg) {
    if (!out || !cfg || cfg->max_batch <= 0) return fail(nullptr, HM_ERR_INVALID, "hm_create: bad arguments");
    *out = nullptr;
    hm_ctx* c = new hm_ctx();
    c->device = device_ordinal;
    c->cfg = *cfg;
    auto bail = [&](int code, const std::string& m) { g_create_error = m; hm_destroy(c); return code; };
    cudaError_t e = cudaSetDevice(device_ordinal);
    if (e != cudaSuccess) return bail(HM_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device_ordinal);
    if (e != cudaSuccess) return bail(HM_ERR_CUDA, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e));
    if (prop.major != 10) return bail(HM_ERR_UNSUPPORTED, "hamer_b200 needs an sm_100 (B200) device, found sm_" +
                                                              std::to_string(prop.major) + std::to_string(prop.minor));
    c->num_sms = prop.multiProcessorCount;
    c->plan.cta_group = (cfg->gemm_cta_group == 1) ? 1 : 2;
    c->plan.bn = 256;
    c->b_format = 0u;   // fp16 weight operands (kind::f16 requires A and B of the same 16-bit type)
    c->att_legacy = cfg->attention_legacy != 0;
    c->mano_legacy = cfg->mano_legacy != 0;
    c->deterministic = cfg->deterministic != 0;
    c->split_w = cfg->split_weights != 0;
    if (const char* e = getenv("HM_SMALL_M")) c->small_m = atoi(e);
    if (!get_tmap_encode()) return bail(HM_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");

    const size_t Mmax = (size_t)cfg->max_batch * TOK, Bm = (size_t)cfg->max_batch;
#define CA(call) do { if ((call) != cudaSuccess) return bail(HM_ERR_CUDA, "workspace allocation failed: " #call); } while (0)
    CA(dev_alloc(c, &c->patches, Mmax * KPATCH));
    CA(dev_alloc(c, &c->xn, Mmax * DM));
    CA(dev_alloc(c, &c->qkv, Mmax * 3 * DM));
    CA(dev_alloc(c, &c->att, Mmax * DM));
    CA(dev_alloc(c, &c->hbuf, Mmax * DFF));
    CA(dev_alloc(c, &c->x, Mmax * DM));
    CA(dev_alloc(c, &c->feats, Mmax * DM));
    CA(dev_alloc(c, &c->hx, Bm * HD));
    CA(dev_alloc(c, &c->hx2, Bm * HD));
    CA(dev_alloc(c, &c->hq, Bm * 512));
    CA(dev_alloc(c, &c->hu, Bm * 8 * DM));
    {   // per-segment partial pooled contexts of the one-pass cross-attention (at most one per 16-token chunk)
        const size_t chunks = Bm * XA_NCH;
        CA(dev_alloc(c, &c->hpart, chunks * 8 * DM));
        CA(dev_alloc(c, &c->hml, chunks * 16));
        CA(dev_alloc(c, &c->hticket, Bm));
        CA(cudaMemset(c->hticket, 0, Bm * sizeof(unsigned int)));
    }
    CA(dev_alloc(c, &c->hpool, Bm * 8 * DM));
    CA(dev_alloc(c, &c->ho, Bm * 512));
    CA(dev_alloc(c, &c->hh, Bm * HD));
    CA(dev_alloc(c, &c->hraw, Bm * NREAD));
    CA(dev_alloc(c, &c->hrot, Bm * 144));
    CA(dev_alloc(c, &c->hcamt, Bm * 3));
    CA(dev_alloc(c, &c->hbetas, Bm * 10));
    CA(dev_alloc(c, &c->hscratch, Bm * 160));
    CA(dev_alloc(c, &c->crop_dev, Bm));
    CA(dev_alloc(c, &c->blur_dev, Bm));
    CA(cudaMallocHost(reinterpret_cast<void**>(&c->blur_host), Bm * sizeof(BlurParams)));
    CA(dev_alloc(c, &c->lut, (size_t)768));
    if (cfg->max_frame_bytes > 0) CA(dev_alloc(c, &c->frame_dev, (size_t)cfg->max_frame_bytes));
    CA(cudaMallocHost(reinterpret_cast<void**>(&c->crop_host), Bm * sizeof(CropParams)));
    c->crop_win.resize(Bm);
    c->crop_rect.resize(Bm);
    CA(cudaEventCreateWithFlags(&c->crop_event, cudaEventDisableTiming));
    CA(cudaEventCreateWithFlags(&c->crop_kernel_event, cudaEventDisableTiming));
    // the workspace is read by TMA beyond the live rows of a partial tile: keep it finite
    CA(cudaMemset(c->patches, 0, Mmax * KPATCH * 2));
    CA(cudaMemset(c->xn, 0, Mmax * DM * 2));
    CA(cudaMemset(c->att, 0, Mmax * DM * 2));
    CA(cudaMemset(c->hbuf, 0, Mmax * DFF * 2));
#undef CA
    // normalisation LUT with the reference's float32 arithmetic (infer.py:145-146, 235-238)
    {
        const double mean[3] = {0.485, 0.456, 0.406}, stdv[3] = {0.229, 0.224, 0.225};
        float lut[768];
        for (int ch = 0; ch < 3; ++ch) {
            const float m = (float)(255. * mean[ch]), sd = (float)(255. * stdv[ch]);
            c->lut_mean[ch] = m; c->lut_sd[ch] = sd;
            for (int v = 0; v < 256; ++v) lut[ch * 256 + v] = ((float)v - m) / sd;
        }
        if (cudaMemcpy(c->lut, lut, sizeof(lut), cudaMemcpyHostToDevice) != cudaSuccess)
            return bail(HM_ERR_CUDA, "LUT upload failed");
    }
    bool ok = make_tmap_bf16_2d(&c->tm_patches, c->patches, Mmax, KPATCH, KPATCH, GEMM_BM) &&
              make_tmap_bf16_2d(&c->tm_xn, c->xn, Mmax, DM, DM, GEMM_BM) &&
              make_tmap_bf16_2d(&c->tm_att, c->att, Mmax, DM, DM, GEMM_BM) &&
              make_tmap_bf16_2d(&c->tm_h, c->hbuf, Mmax, DFF, DFF, GEMM_BM) &&
              make_tmap_f32_out(&c->tm_x_out, c->x, Mmax, DM, DM) &&
              make_tmap_attn(&c->tm_att128, &c->tm_att32, c->qkv, Mmax);
    if (!ok) return bail(HM_ERR_CUDA, "activation tensor-map encoding failed");
    const int sk_smem = SK_SMEM_MAX;
    bool sk_ok = true;
    auto sk_attr = [&](auto kern) { sk_ok = sk_ok && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, sk_smem) == cudaSuccess; };
#define SK_ATTR_ALL(CPW)                                                                                        \
    sk_attr(skinny_linear_kernel<CPW, SK_NONE, false>); sk_attr(skinny_linear_kernel<CPW, SK_NONE, true>);     \
    sk_attr(skinny_linear_kernel<CPW, SK_GELU, false>); sk_attr(skinny_linear_kernel<CPW, SK_GELU, true>);     \
    sk_attr(skinny_linear_kernel<CPW, SK_RESID, false>); sk_attr(skinny_linear_kernel<CPW, SK_RESID, true>);
    SK_ATTR_ALL(1) SK_ATTR_ALL(2) SK_ATTR_ALL(4)
#undef SK_ATTR_ALL
#define SK_ATTR_RT(RT)                                                                                                \
    sk_attr(skinny_linear_kernel<1, SK_NONE, false, RT>); sk_attr(skinny_linear_kernel<1, SK_NONE, true, RT>);       \
    sk_attr(skinny_linear_kernel<1, SK_GELU, false, RT>); sk_attr(skinny_linear_kernel<1, SK_GELU, true, RT>);       \
    sk_attr(skinny_linear_kernel<1, SK_RESID, false, RT>); sk_attr(skinny_linear_kernel<1, SK_RESID, true, RT>);
    SK_ATTR_RT(4) SK_ATTR_RT(8)
#undef SK_ATTR_RT
    sk_ok = sk_ok && cudaFuncSetAttribute(head_xattn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XA_SMEM_BYTES) == cudaSuccess &&
            cudaFuncSetAttribute(head_xattn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XA_SMEM_BYTES) == cudaSuccess;
    if (!sk_ok ||
        cudaFuncSetAttribute(vit_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             ATT_SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(vit_attention_tc5_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)AT5_SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(vit_attention_tc5_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)AT5_SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(mano_skin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(SkinSmem)) != cudaSuccess ||
        cudaFuncSetAttribute(mano_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)MF_SMEM_BYTES) != cudaSuccess)
        return bail(HM_ERR_CUDA, "cudaFuncSetAttribute(max dynamic smem) failed");
    *out = c;
    return HM_OK;
}

void hm_destroy(hm_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    if (c->comm) { nccl_api().CommDestroy(c->comm); c->comm = nullptr; }
    for (void* p : c->allocs) cudaFree(p);
    for (void* p : c->mano_allocs) cudaFree(p);
    for (void* p : c->rn_allocs) cudaFree(p);
    for (auto& kv : c->w) { cudaFree(kv.second.f32); cudaFree(kv.second.bf16); }
    if (c->crop_host) cudaFreeHost(c->crop_host);
    if (c->blur_host) cudaFreeHost(c->blur_host);
    if (c->crop_event) cudaEventDestroy(c->crop_event);
    if (c->crop_kernel_event) cudaEventDestroy(c->crop_kernel_event);
    for (cudaEvent_t e : c->prof_pool) cudaEventDestroy(e);
    delete c;
}

int hm_set_weight(hm_ctx* c, const char* name, const float* data, const int64_t* shape, int ndim) {
    if (!c || !name || !data || !shape || ndim < 1 || ndim > 4) return fail(c, HM_ERR_INVALID, "hm_set_weight: bad arguments");
    HM_CUDA(c, cudaSetDevice(c->device));
    DevTensor t;
    t.numel = 1;
    for (int i = 0; i < ndim; ++i) { t.shape.push_back(shape[i]); t.numel *= (size_t)shape[i]; }
    const std::string key(name);
    auto old = c->w.find(key);
    if (old != c->w.end()) { cudaFree(old->second.f32); cudaFree(old->second.bf16); c->w.erase(old); }
    if (key.rfind("rootnet.", 0) == 0) c->rn_ready = false; else c->finalized = false;
    HM_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&t.f32), t.numel * 4));
    cudaError_t e = cudaMemcpy(t.f32, data, t.numel * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && is_backbone_matrix(key) && ndim == 2) {
        // tensor-core operand: keep only the 16-bit (fp16) copy ([hi | lo] pairs with split_weights)
        const int split = c->split_w ? 1 : 0;
        e = cudaMalloc(reinterpret_cast<void**>(&t.bf16), t.numel * 2 * (split + 1));
        if (e == cudaSuccess) {
            f32_to_w16_kernel<<<(unsigned)((t.numel + 255) / 256), 256>>>(
                t.f32, reinterpret_cast<uint16_t*>(t.bf16), t.numel, (int)c->b_format, split, (int)shape[1]);
            e = cudaDeviceSynchronize();
        }
        cudaFree(t.f32);
        t.f32 = nullptr;
    }
    if (e != cudaSuccess) {
        cudaFree(t.f32);
        cudaFree(t.bf16);
        return fail(c, HM_ERR_CUDA, std::string("hm_set_weight(") + name + "): " + cudaGetErrorString(e));
    }
    c->w[key] = t;
    return HM_OK;
}

int hm_finalize_weights(hm_ctx* c) {
    if (!c) return HM_ERR_INVALID;
    HM_CUDA(c, cudaSetDevice(c->device));
    // ---- presence / shape checks -----------------------------------------------------------
    auto need = [&](const std::string& n, std::initializer_list<int64_t> s) -> bool {
        if (!shape_is(find(c, n), s)) { c->err = "missing or mis-shaped weight: " + n; return false; }
        return true;
    };
    if (!need("backbone.patch_embed.proj.weight", {DM, 3, 16, 16}) || !need("backbone.patch_embed.proj.bias", {DM}) ||
        !need("backbone.pos_embed", {1, TOK + 1, DM}) || !need("backbone.last_norm.weight", {DM}) ||
        !need("backbone.last_norm.bias", {DM}))
        return HM_ERR_INVALID;
    for (int i = 0; i < DEPTH; ++i) {
        if (!need(blk_name(i, "norm1.weight"), {DM}) || !need(blk_name(i, "norm1.bias"), {DM}) ||
            !need(blk_name(i, "norm2.weight"), {DM}) || !need(blk_name(i, "norm2.bias"), {DM}) ||
            !need(blk_name(i, "attn.qkv.weight"), {3 * DM, DM}) || !need(blk_name(i, "attn.qkv.bias"), {3 * DM}) ||
            !need(blk_name(i, "attn.proj.weight"), {DM, DM}) || !need(blk_name(i, "attn.proj.bias"), {DM}) ||
            !need(blk_name(i, "mlp.fc1.weight"), {DFF, DM}) || !need(blk_name(i, "mlp.fc1.bias"), {DFF}) ||
            !need(blk_name(i, "mlp.fc2.weight"), {DM, DFF}) || !need(blk_name(i, "mlp.fc2.bias"), {DM}))
            return HM_ERR_INVALID;
    }
    if (!need("mano_head.transformer.pos_embedding", {1, 1, HD}) ||
        !need("mano_head.transformer.to_token_embedding.weight", {HD, 1}) ||
        !need("mano_head.transformer.to_token_embedding.bias", {HD}) ||
        !need("mano_head.decpose.weight", {96, HD}) || !need("mano_head.decpose.bias", {96}) ||
        !need("mano_head.decshape.weight", {10, HD}) || !need("mano_head.decshape.bias", {10}) ||
        !need("mano_head.deccam.weight", {3, HD}) || !need("mano_head.deccam.bias", {3}) ||
        !need("mano_head.init_hand_pose", {1, 96}) || !need("mano_head.init_betas", {1, 10}) ||
        !need("mano_head.init_cam", {1, 3}))
        return HM_ERR_INVALID;
    for (int l = 0; l < HDEPTH; ++l) {
        for (int sub = 0; sub < 3; ++sub)
            if (!need(dec_name(l, sub, "norm.weight"), {HD}) || !need(dec_name(l, sub, "norm.bias"), {HD}))
                return HM_ERR_INVALID;
        if (!need(dec_name(l, 0, "fn.to_qkv.weight"), {1536, HD}) ||
            !need(dec_name(l, 0, "fn.to_out.0.weight"), {HD, 512}) || !need(dec_name(l, 0, "fn.to_out.0.bias"), {HD}) ||
            !need(dec_name(l, 1, "fn.to_kv.weight"), {1024, DM}) || !need(dec_name(l, 1, "fn.to_q.weight"), {512, HD}) ||
            !need(dec_name(l, 1, "fn.to_out.0.weight"), {HD, 512}) || !need(dec_name(l, 1, "fn.to_out.0.bias"), {HD}) ||
            !need(dec_name(l, 2, "fn.net.0.weight"), {HD, HD}) || !need(dec_name(l, 2, "fn.net.0.bias"), {HD}) ||
            !need(dec_name(l, 2, "fn.net.3.weight"), {HD, HD}) || !need(dec_name(l, 2, "fn.net.3.bias"), {HD}))
            return HM_ERR_INVALID;
    }
    // ---- fused tables (built on the host from device copies; one-time cost) -----------------
    auto fetch = [&](const std::string& n, std::vector<float>& h) -> cudaError_t {
        const DevTensor* t = find(c, n);
        h.resize(t->numel);
        return cudaMemcpy(h.data(), t->f32, t->numel * 4, cudaMemcpyDeviceToHost);
    };
    std::vector<float> h, h2, h3;
    const int wmul = c->split_w ? 2 : 1;     // split_weights: every 16-bit weight matrix is [hi | lo], K doubled
    if (!c->w_patch2) {
        HM_CUDA(c, dev_alloc(c, &c->w_patch2, (size_t)DM * KPATCH * wmul));
        HM_CUDA(c, dev_alloc(c, &c->pos_tab, (size_t)TOK * DM));
        HM_CUDA(c, dev_alloc(c, &c->tok0, (size_t)HD));
        HM_CUDA(c, dev_alloc(c, &c->w_read, (size_t)NREAD * HD));
        HM_CUDA(c, dev_alloc(c, &c->b_read, (size_t)NREAD));
        for (int l = 0; l < HDEPTH; ++l) HM_CUDA(c, dev_alloc(c, &c->w_sa[l], (size_t)HD * HD));
    }
    {   // patch weight [1280,768] -> bf16 [1280, 1536] = [W | W]
        HM_CUDA(c, fetch("backbone.patch_embed.proj.weight", h));
        std::vector<float> w2((size_t)DM * KPATCH);
        for (int n = 0; n < DM; ++n)
            for (int k = 0; k < 768; ++k) {
                w2[(size_t)n * KPATCH + k] = h[(size_t)n * 768 + k];
                w2[(size_t)n * KPATCH + 768 + k] = h[(size_t)n * 768 + k];
            }
        float* tmp = nullptr;
        HM_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&tmp), w2.size() * 4));
        cudaError_t e1 = cudaMemcpy(tmp, w2.data(), w2.size() * 4, cudaMemcpyHostToDevice);
        f32_to_w16_kernel<<<(unsigned)((w2.size() + 255) / 256), 256>>>(
            tmp, reinterpret_cast<uint16_t*>(c->w_patch2), w2.size(), (int)c->b_format, c->split_w ? 1 : 0, KPATCH);
        cudaError_t e2 = cudaDeviceSynchronize();
        cudaFree(tmp);
        HM_CUDA(c, e1);
        HM_CUDA(c, e2);
    }
    {   // pos table = pos[:,1:] + pos[:,:1]  (vit.py:327 adds them to x one after the other)
        HM_CUDA(c, fetch("backbone.pos_embed", h));
        std::vector<float> tab((size_t)TOK * DM);
        for (int t = 0; t < TOK; ++t)
            for (int d = 0; d < DM; ++d) tab[(size_t)t * DM + d] = h[(size_t)(t + 1) * DM + d] + h[d];
        HM_CUDA(c, cudaMemcpy(c->pos_tab, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
    }
    {   // layer-0 token: Linear(1,1024)(0) + pos_embedding = bias + pos  (mano_head.py:86, pose_transformer.py:350-354)
        HM_CUDA(c, fetch("mano_head.transformer.to_token_embedding.bias", h));
        HM_CUDA(c, fetch("mano_head.transformer.pos_embedding", h2));
        for (int i = 0; i < HD; ++i) h[i] += h2[i];
        HM_CUDA(c, cudaMemcpy(c->tok0, h.data(), HD * 4, cudaMemcpyHostToDevice));
    }
    for (int l = 0; l < HDEPTH; ++l) {
        // Self-attention over ONE token: softmax of a single logit is 1, so SA(y) = to_out(W_v y) = (W_o W_v) y + b_o
        // (pose_transformer.py:55-86).  The product is formed once here in double and rounded to fp32: one 1024 x 1024
        // matrix instead of a 512 x 1024 and a 1024 x 512 launch per layer (the rounding of the merged matrix, 6e-8
        // relative, is below the fp32 noise of evaluating the two factors one after the other).
        HM_CUDA(c, fetch(dec_name(l, 0, "fn.to_qkv.weight"), h));       // [1536][1024], rows 1024.. = W_v
        HM_CUDA(c, fetch(dec_name(l, 0, "fn.to_out.0.weight"), h2));    // [1024][512]
        std::vector<float> m((size_t)HD * HD);
        std::vector<double> row(HD);
        for (int n = 0; n < HD; ++n) {
            std::fill(row.begin(), row.end(), 0.0);
            for (int j = 0; j < 512; ++j) {
                const double wo = h2[(size_t)n * 512 + j];
                const float* wv = &h[(size_t)(1024 + j) * HD];
                for (int k = 0; k < HD; ++k) row[k] += wo * (double)wv[k];
            }
            for (int k = 0; k < HD; ++k) m[(size_t)n * HD + k] = (float)row[k];
        }
        HM_CUDA(c, cudaMemcpy(c->w_sa[l], m.data(), m.size() * 4, cudaMemcpyHostToDevice));
    }
    {   // read-out: [decpose; decshape; deccam], bias + init_* (mano_head.py:93-95)
        std::vector<float> wr((size_t)NREAD * HD), br(NREAD);
        const char* wn[3] = {"mano_head.decpose", "mano_head.decshape", "mano_head.deccam"};
        const char* in[3] = {"mano_head.init_hand_pose", "mano_head.init_betas", "mano_head.init_cam"};
        const int rows[3] = {96, 10, 3};
        int r0 = 0;
        for (int q = 0; q < 3; ++q) {
            HM_CUDA(c, fetch(std::string(wn[q]) + ".weight", h));
            HM_CUDA(c, fetch(std::string(wn[q]) + ".bias", h2));
            HM_CUDA(c, fetch(in[q], h3));
            memcpy(&wr[(size_t)r0 * HD], h.data(), (size_t)rows[q] * HD * 4);
            for (int r = 0; r < rows[q]; ++r) br[r0 + r] = h2[r] + h3[r];
            r0 += rows[q];
        }
        HM_CUDA(c, cudaMemcpy(c->w_read, wr.data(), wr.size() * 4, cudaMemcpyHostToDevice));
        HM_CUDA(c, cudaMemcpy(c->b_read, br.data(), br.size() * 4, cudaMemcpyHostToDevice));
    }
    // ---- weight tensor maps ------------------------------------------------------------------
    const uint32_t brow = (uint32_t)(c->plan.bn / c->plan.cta_group);
    bool ok = make_tmap_bf16_2d(&c->tm_wpatch, c->w_patch2, DM, KPATCH * wmul, KPATCH * wmul, brow);
    for (int i = 0; i < DEPTH && ok; ++i) {
        ok = make_tmap_bf16_2d(&c->blk[i].tm_qkv, find(c, blk_name(i, "attn.qkv.weight"))->bf16, 3 * DM, DM * wmul, DM * wmul, brow) &&
             make_tmap_bf16_2d(&c->blk[i].tm_proj, find(c, blk_name(i, "attn.proj.weight"))->bf16, DM, DM * wmul, DM * wmul, brow) &&
             make_tmap_bf16_2d(&c->blk[i].tm_fc1, find(c, blk_name(i, "mlp.fc1.weight"))->bf16, DFF, DM * wmul, DM * wmul, brow) &&
             make_tmap_bf16_2d(&c->blk[i].tm_fc2, find(c, blk_name(i, "mlp.fc2.weight"))->bf16, DM, DFF * wmul, DFF * wmul, brow);
    }
    if (!ok) return fail(c, HM_ERR_CUDA, "weight tensor-map encoding failed");
    // ---- cached raw pointers -----------------------------------------------------------------
    auto F = [&](const std::string& n) { return (const float*)find(c, n)->f32; };
    c->b_patch = F("backbone.patch_embed.proj.bias");
    c->lnf_g = F("backbone.last_norm.weight");
    c->lnf_b = F("backbone.last_norm.bias");
    for (int i = 0; i < DEPTH; ++i) {
        BlockW& w = c->blk[i];
        w.ln1_g = F(blk_name(i, "norm1.weight")); w.ln1_b = F(blk_name(i, "norm1.bias"));
        w.ln2_g = F(blk_name(i, "norm2.weight")); w.ln2_b = F(blk_name(i, "norm2.bias"));
        w.b_qkv = F(blk_name(i, "attn.qkv.bias")); w.b_proj = F(blk_name(i, "attn.proj.bias"));
        w.b_fc1 = F(blk_name(i, "mlp.fc1.bias")); w.b_fc2 = F(blk_name(i, "mlp.fc2.bias"));
    }
    for (int l = 0; l < HDEPTH; ++l) {
        HeadLayerW& w = c->hl[l];
        w.n0g = F(dec_name(l, 0, "norm.weight")); w.n0b = F(dec_name(l, 0, "norm.bias"));
        w.n1g = F(dec_name(l, 1, "norm.weight")); w.n1b = F(dec_name(l, 1, "norm.bias"));
        w.n2g = F(dec_name(l, 2, "norm.weight")); w.n2b = F(dec_name(l, 2, "norm.bias"));
        w.w_sa = c->w_sa[l]; w.b_sa_o = F(dec_name(l, 0, "fn.to_out.0.bias"));
        w.w_q = F(dec_name(l, 1, "fn.to_q.weight"));
        w.w_k = F(dec_name(l, 1, "fn.to_kv.weight"));
        w.w_v = F(dec_name(l, 1, "fn.to_kv.weight")) + (size_t)512 * DM;
        w.w_ca_o = F(dec_name(l, 1, "fn.to_out.0.weight")); w.b_ca_o = F(dec_name(l, 1, "fn.to_out.0.bias"));
        w.w1 = F(dec_name(l, 2, "fn.net.0.weight")); w.b1 = F(dec_name(l, 2, "fn.net.0.bias"));
        w.w2 = F(dec_name(l, 2, "fn.net.3.weight")); w.b2 = F(dec_name(l, 2, "fn.net.3.bias"));
    }
    c->finalized = true;
    return HM_OK;
}

int hm_set_mano(hm_ctx* c, const float* v_template, const float* shapedirs, const float* posedirs,
                const float* J_regressor, const float* lbs_weights, const int32_t* parents,
                const int32_t* tip_ids, const int32_t* joint_map) {
    if (!c || !v_template || !shapedirs || !posedirs || !J_regressor || !lbs_weights || !parents ||
        !tip_ids || !joint_map)
        return fail(c, HM_ERR_INVALID, "hm_set_mano: null argument");
    HM_CUDA(c, cudaSetDevice(c->device));
    for (int j = 0; j < 16; ++j)
        if (parents[j] >= j || parents[j] < -1) return fail(c, HM_ERR_INVALID, "hm_set_mano: parents must precede children");
    for (int q = 0; q < 5; ++q)
        if (tip_ids[q] < 0 || tip_ids[q] >= MANO_V) return fail(c, HM_ERR_INVALID, "hm_set_mano: tip id out of range");
    for (int q = 0; q < 21; ++q)
        if (joint_map[q] < 0 || joint_map[q] >= 21) return fail(c, HM_ERR_INVALID, "hm_set_mano: joint_map out of range");
    // GEMM weight: basis16[n = v*3+c][k] with k = [hi(160) | lo(160)], k < 10 shapedirs, 10..144 posedirs
    std::vector<__half> basis16((size_t)MANO_NPAD * MANO_KTOT, __float2half(0.f));
    std::vector<float> tmpl(MANO_NPAD, 0.f), zrow(MANO_NPAD, 0.f);
    for (int v = 0; v < MANO_V; ++v)
        for (int cdim = 0; cdim < 3; ++cdim) {
            const int n = v * 3 + cdim;
            tmpl[n] = v_template[n];
            for (int k = 0; k < 145; ++k) {
                const float b = (k < 10) ? shapedirs[n * 10 + k] : posedirs[(size_t)(k - 10) * 2334 + n];
                const __half hi = __float2half_rn(b);
                const __half lo = __float2half_rn(b - __half2float(hi));
                basis16[(size_t)n * MANO_KTOT + k] = hi;
                basis16[(size_t)n * MANO_KTOT + MANO_KB + k] = lo;
            }
        }
    // fused kernel: same basis with the rows of each vertex group ordered (x0 x1 y0 y1 z0 z1) per vertex pair, so that the
    // accumulator columns a thread reads are packed pairs for FFMA2; K = [shape hi | pose | shape lo | template hi | lo | 0 0 0]
    std::vector<__half> basis16f((size_t)MANO_NPAD * MANO_KB, __float2half(0.f));
    for (int v = 0; v < MANO_V; ++v)
        for (int cdim = 0; cdim < 3; ++cdim) {
            const int i = v % MF_VPT;
            const size_t nf = (size_t)(v / MF_VPT) * MF_GCOLS + 6 * (i >> 1) + 2 * cdim + (i & 1);
            const __half* src = &basis16[(size_t)(v * 3 + cdim) * MANO_KTOT];
            __half* dst = &basis16f[nf * MANO_KB];
            for (int k = 0; k < 145; ++k) dst[k] = src[k];                     // shape hi | pose (fp16-rounded)
            for (int k = 0; k < 10; ++k) dst[145 + k] = src[MANO_KB + k];      // shape lo, multiplied by the repeated betas
            const float t = v_template[v * 3 + cdim];
            dst[MF_K_TMPL] = __float2half_rn(t);                               // template hi / lo, coefficient 1
            dst[MF_K_TMPL + 1] = __float2half_rn(t - __half2float(dst[MF_K_TMPL]));
        }
    // folded joint regressor (double accumulation; one-time)
    std::vector<float> jt(48), js(480);
    for (int j = 0; j < 16; ++j)
        for (int cdim = 0; cdim < 3; ++cdim) {
            double a = 0;
            for (int v = 0; v < MANO_V; ++v) a += (double)J_regressor[j * MANO_V + v] * v_template[v * 3 + cdim];
            jt[j * 3 + cdim] = (float)a;
            for (int l = 0; l < 10; ++l) {
                double s = 0;
                for (int v = 0; v < MANO_V; ++v)
                    s += (double)J_regressor[j * MANO_V + v] * shapedirs[(v * 3 + cdim) * 10 + l];
                js[(j * 3 + cdim) * 10 + l] = (float)s;
            }
        }
    // skinning tables: the kernel walks each warp's union of joints with broadcast shared-memory reads;
    // lanes whose vertex does not use a joint carry weight 0 for it
    std::vector<int> order(MANO_V);
    std::vector<uint32_t> mask(MANO_V, 0);
    for (int v = 0; v < MANO_V; ++v) {
        order[v] = v;
        for (int j = 0; j < 16; ++j) if (lbs_weights[v * 16 + j] != 0.f) mask[v] |= 1u << j;
    }
    // natural vertex order: a warp's 32 vertices are contiguous in memory (coalesced I/O without staging);
    // sorting by joint set would shrink the unions (4.4 vs 6.6 joints on MANO) but scatter the stores
    std::vector<int32_t> perm(MANO_VTHREADS, -1), wnu(MANO_WARPS, 0), wun((size_t)MANO_WARPS * MANO_MAX_UNION, 0);
    std::vector<float> lw((size_t)MANO_VTHREADS * MANO_MAX_UNION, 0.f);
    for (int t = 0; t < MANO_V; ++t) perm[t] = order[t];
    for (int w = 0; w < MANO_WARPS; ++w) {
        uint32_t um = 0;
        for (int l = 0; l < 32; ++l) if (perm[w * 32 + l] >= 0) um |= mask[perm[w * 32 + l]];
        int nu = 0;
        for (int j = 0; j < 16; ++j)
            if (um & (1u << j)) {
                wun[(size_t)w * MANO_MAX_UNION + nu] = j;
                for (int l = 0; l < 32; ++l) {
                    const int v = perm[w * 32 + l];
                    if (v >= 0) lw[(size_t)(w * 32 + l) * MANO_MAX_UNION + nu] = lbs_weights[v * 16 + j];
                }
                ++nu;
            }
        wnu[w] = nu;
    }
    // fused kernel: union of joints + weights per vertex group (warp-uniform in the epilogue); layout at MF_TAB_STRIDE
    std::vector<float> ftab((size_t)MF_LIVE_GROUPS * MF_TAB_STRIDE / 4, 0.f);
    bool union_fits = true;
    for (int g = 0; g < MF_LIVE_GROUPS; ++g) {
        float* tab = ftab.data() + (size_t)g * MF_TAB_STRIDE / 4;
        int32_t* tabi = reinterpret_cast<int32_t*>(tab);
        uint32_t um = 0;
        for (int i = 0; i < MF_VPT; ++i) {
            const int v = g * MF_VPT + i;
            if (v < MANO_V) um |= mask[v];
        }
        for (int q = 0; q < 5; ++q) if (tip_ids[q] / MF_VPT == g) tabi[1] |= 1 << q;
        int nu = 0;
        for (int j = 0; j < 16; ++j)
            if (um & (1u << j)) {
                if (nu == MF_MAX_UNION) { union_fits = false; break; }
                tabi[4 + nu] = 12 * j;
                for (int i = 0; i < MF_VPT; ++i) {
                    const int v = g * MF_VPT + i;
                    if (v >= MANO_V || lbs_weights[v * 16 + j] == 0.f) continue;
                    tab[4 + MF_MAX_UNION + nu * MF_VPT + i] = lbs_weights[v * 16 + j];
                    tabi[4 + nu] |= 1 << (16 + (i >> 1));
                }
                ++nu;
            }
        tabi[0] = nu;
    }
    {   // output slots: joint_map[k] = source (0..15 chain joint, 16..20 fingertip)
        int count[21] = {0};
        for (int j = 0; j < 16; ++j) c->mano.chain_slot[j] = -1;
        for (int q = 0; q < 5; ++q) c->mano.tip_slot[q] = -1;
        for (int k = 0; k < 21; ++k) {
            const int src = joint_map[k];
            ++count[src];
            if (src < 16) c->mano.chain_slot[src] = k; else c->mano.tip_slot[src - 16] = k;
        }
        c->mano_fused_ok = union_fits;
        for (int q = 0; q < 21; ++q) if (count[q] != 1) c->mano_fused_ok = false;   // not a permutation: 3-kernel path
        for (int j = 0; j < 16; ++j) if (parents[j] != mf_parent(j)) c->mano_fused_ok = false;   // prologue is specialised for the MANO tree
        for (int j = 0; j < 16; ++j) {
            c->mano_chain.chain_slot[j] = c->mano.chain_slot[j];
            for (int cdim = 0; cdim < 3; ++cdim) {
                c->mano_chain.jreg[j][cdim][0] = jt[j * 3 + cdim];
                for (int l = 0; l < 10; ++l) c->mano_chain.jreg[j][cdim][1 + l] = js[(j * 3 + cdim) * 10 + l];
            }
        }
        for (int q = 0; q < 5; ++q) { c->mano_chain.tip_slot[q] = c->mano.tip_slot[q]; c->mano_chain.tip_ids[q] = tip_ids[q]; }
    }
    for (void* p : c->mano_allocs) cudaFree(p);
    for (void* p : c->rn_allocs) cudaFree(p);
    c->mano_allocs.clear();
    c->mano_set = false;
    auto up = [&](const void* src, size_t bytes, const void** dst) -> cudaError_t {
        void* d = nullptr;
        cudaError_t e = cudaMalloc(&d, bytes);
        if (e != cudaSuccess) return e;
        c->mano_allocs.push_back(d);
        *dst = d;
        return cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice);
    };
    HM_CUDA(c, up(basis16.data(), basis16.size() * 2, reinterpret_cast<const void**>(&c->mano.basis16)));
    HM_CUDA(c, up(basis16f.data(), basis16f.size() * 2, reinterpret_cast<const void**>(&c->mano.basis16f)));
    HM_CUDA(c, up(tmpl.data(), tmpl.size() * 4, reinterpret_cast<const void**>(&c->mano.template_n)));
    HM_CUDA(c, up(zrow.data(), zrow.size() * 4, reinterpret_cast<const void**>(&c->mano.zero_row)));
    HM_CUDA(c, up(jt.data(), jt.size() * 4, reinterpret_cast<const void**>(&c->mano.j_template)));
    HM_CUDA(c, up(js.data(), js.size() * 4, reinterpret_cast<const void**>(&c->mano.j_shapedirs)));
    HM_CUDA(c, up(perm.data(), perm.size() * 4, reinterpret_cast<const void**>(&c->mano.perm)));
    HM_CUDA(c, up(wnu.data(), wnu.size() * 4, reinterpret_cast<const void**>(&c->mano.warp_nu)));
    HM_CUDA(c, up(wun.data(), wun.size() * 4, reinterpret_cast<const void**>(&c->mano.warp_union)));
    HM_CUDA(c, up(lw.data(), lw.size() * 4, reinterpret_cast<const void**>(&c->mano.lane_w)));
    HM_CUDA(c, up(ftab.data(), ftab.size() * 4, reinterpret_cast<const void**>(&c->mano.fz_tab)));
    // per-chunk workspace (coefficients, v_posed, skinning transforms, posed joints) + tensor maps
    {
        void* d = nullptr;
        const size_t ch = MANO_CHUNK;
        HM_CUDA(c, cudaMalloc(&d, ch * MANO_KTOT * 2)); c->mano_allocs.push_back(d); c->m_coef = reinterpret_cast<__half*>(d);
        HM_CUDA(c, cudaMalloc(&d, ch * MANO_NPAD * 4)); c->mano_allocs.push_back(d); c->m_vposed = reinterpret_cast<float*>(d);
        HM_CUDA(c, cudaMalloc(&d, ch * MANO_J * 12 * 4)); c->mano_allocs.push_back(d); c->m_A = reinterpret_cast<float*>(d);
        HM_CUDA(c, cudaMalloc(&d, ch * MANO_J * 3 * 4)); c->mano_allocs.push_back(d); c->m_jpos = reinterpret_cast<float*>(d);
        HM_CUDA(c, cudaMemset(c->m_coef, 0, ch * MANO_KTOT * 2));
        if (!make_tmap_bf16_2d(&c->tm_mcoef, c->m_coef, ch, MANO_KTOT, MANO_KTOT, GEMM_BM) ||
            !make_tmap_bf16_2d(&c->tm_mbasis, c->mano.basis16, MANO_NPAD, MANO_KTOT, MANO_KTOT, 128 / c->plan.cta_group))
            return fail(c, HM_ERR_CUDA, "hm_set_mano: tensor-map encoding failed");
        if (!make_tmap_bf16_2d(&c->tm_fbasis, c->mano.basis16f, MANO_NPAD, MANO_KB, MANO_KB, MF_BN / 2))
            return fail(c, HM_ERR_CUDA, "hm_set_mano: tensor-map encoding failed (fused)");
    }
    for (int j = 0; j < 16; ++j) c->mano.parents[j] = parents[j];
    c->mano.max_depth = 0;
    for (int j = 0; j < 16; ++j) {
        int d = 0;
        for (int q = parents[j]; q >= 0; q = parents[q]) ++d;
        if (d > c->mano.max_depth) c->mano.max_depth = d;
    }
    for (int q = 0; q < 5; ++q) c->mano.tip_ids[q] = tip_ids[q];
    for (int q = 0; q < 21; ++q) c->mano.joint_map[q] = joint_map[q];
    c->mano_set = true;
    return HM_OK;
}

// Common body of hm_preprocess / hm_preprocess_centers: ccs = [n][cx, cy, size] of the square source windows (for mode 2
// cx is given in the mirrored image), mode 0 = plain, 1 = mirror the patch after the warp, 2 = sample the mirrored image.
static int preprocess_impl(hm_ctx* c, const uint8_t* frame_bgr, int H, int W, int64_t row_stride, int frame_on_device,
                           const double* ccs, const uint8_t* mode, int n, float* img_out, float* trans, cudaStream_t s,
                           const int32_t* blur_radius = nullptr, const double* blur_weights = nullptr) {
    HM_CUDA(c, cudaSetDevice(c->device));
    if (c->crop_event_pending) {   // pinned staging still owned by the previous call's copy?
        HM_CUDA(c, cudaEventSynchronize(c->crop_event));
        c->crop_event_pending = false;
    }
    // crop_dev / frame_dev are single context-owned buffers: a call on another stream (pipelined frames) must not
    // overwrite them while the previous call's crop kernel still reads them (no-op on the same stream)
    if (c->crop_kernel_recorded) HM_CUDA(c, cudaStreamWaitEvent(s, c->crop_kernel_event, 0));
    double band_lo = (double)H, band_hi = -1.0;   // source rows any crop can read (host frames: only these are uploaded)
    bool any_blur = false;
    for (int i = 0; i < n; ++i) {
        const double cx = ccs[3 * i], cy = ccs[3 * i + 1], size = ccs[3 * i + 2];
        if (!(size > 0)) return fail(c, HM_ERR_INVALID, "hm_preprocess: empty box");
        if (mode[i] > 2) return fail(c, HM_ERR_INVALID, "hm_preprocess: bad crop mode");
        double M[6];
        center_to_affine(cx, cy, size, M);
        // anti-aliased crops read the blur's support around every tap
        const int br = blur_radius ? blur_radius[i] : 0;
        if (br < 0 || br > CROP_BLUR_MAX_R) return fail(c, HM_ERR_INVALID, "hm_preprocess_blur: radius outside [0, 32]");
        c->blur_host[i].radius = br;
        c->blur_host[i].pad = 0;
        if (br > 0) {
            any_blur = true;
            for (int k = 0; k <= CROP_BLUR_MAX_R; ++k) c->blur_host[i].w[k] = k <= br ? blur_weights[(size_t)i * (CROP_BLUR_MAX_R + 1) + k] : 0.0;
        }
        const double half = 0.5 * size + (double)br;
        if (cy - half < band_lo) band_lo = cy - half;
        if (cy + half > band_hi) band_hi = cy + half;
        // columns of the ORIGINAL frame the crop reads (mode 2 samples column W - 1 - x)
        const double xa = cx - half, xb = cx + half;
        c->crop_win[i][0] = mode[i] == 2 ? (double)(W - 1) - xb : xa;
        c->crop_win[i][1] = mode[i] == 2 ? (double)(W - 1) - xa : xb;
        c->crop_win[i][2] = cy - half; c->crop_win[i][3] = cy + half;
        invert_affine(M, c->crop_host[i].m);
        c->crop_host[i].flip = mode[i];
        c->crop_host[i].pad = 0;
        if (trans) for (int q = 0; q < 6; ++q) trans[6 * i + q] = (float)M[q];
    }
    const uint8_t* frame = frame_bgr;
    if (!frame_on_device) {
        const size_t bytes = (size_t)H * (size_t)row_stride;
        if (!c->frame_dev || bytes - (size_t)(row_stride - (int64_t)W * 3) > (size_t)c->cfg.max_frame_bytes)
            return fail(c, HM_ERR_INVALID, "hm_preprocess: frame larger than hm_config.max_frame_bytes");
        // the warp maps the 256 x 256 patch onto the square [c - size/2, c + size/2]^2 (no rotation) and reads
        // bilinear neighbours: rows outside the union of those windows (+ margin) are never touched
        long r0 = (long)floor(band_lo) - 4, r1 = (long)ceil(band_hi) + 4;
        if (r0 < 0) r0 = 0;
        if (r1 > H - 1) r1 = H - 1;
        // per-crop rectangles when together they are much smaller than the band (hands scattered over the frame)
        size_t rect_bytes = 0;
        for (int i = 0; i < n; ++i) {
            long x0 = (long)floor(c->crop_win[i][0]) - 4, x1 = (long)ceil(c->crop_win[i][1]) + 4;
            long y0 = (long)floor(c->crop_win[i][2]) - 4, y1 = (long)ceil(c->crop_win[i][3]) + 4;
            if (x0 < 0) x0 = 0; if (y0 < 0) y0 = 0; if (x1 > W - 1) x1 = W - 1; if (y1 > H - 1) y1 = H - 1;
            c->crop_rect[i][0] = x0; c->crop_rect[i][1] = x1; c->crop_rect[i][2] = y0; c->crop_rect[i][3] = y1;
            if (x1 >= x0 && y1 >= y0) rect_bytes += (size_t)(x1 - x0 + 1) * 3 * (size_t)(y1 - y0 + 1);
        }
        // whole rows except the last one, which ends with its last pixel (the caller's buffer need not hold row padding there)
        const size_t band_bytes = r1 >= r0 ? (size_t)(r1 - r0) * row_stride + (size_t)W * 3 : 0;
        if (rect_bytes + (size_t)n * 65536 < band_bytes) {       // 64 KB per copy ~ the fixed cost of a pageable 2-D copy
            for (int i = 0; i < n; ++i) {
                const long x0 = c->crop_rect[i][0], x1 = c->crop_rect[i][1], y0 = c->crop_rect[i][2], y1 = c->crop_rect[i][3];
                if (x1 < x0 || y1 < y0) continue;               // window entirely outside the frame: constant border only
                const size_t off = (size_t)y0 * row_stride + (size_t)x0 * 3;
                HM_CUDA(c, cudaMemcpy2DAsync(c->frame_dev + off, (size_t)row_stride, frame_bgr + off, (size_t)row_stride,
                                             (size_t)(x1 - x0 + 1) * 3, (size_t)(y1 - y0 + 1), cudaMemcpyHostToDevice, s));
            }
        } else if (r1 >= r0) {
            HM_CUDA(c, cudaMemcpyAsync(c->frame_dev + (size_t)r0 * row_stride, frame_bgr + (size_t)r0 * row_stride,
                                       band_bytes, cudaMemcpyHostToDevice, s));
        }
        frame = c->frame_dev;
    }
    HM_CUDA(c, cudaMemcpyAsync(c->crop_dev, c->crop_host, (size_t)n * sizeof(CropParams), cudaMemcpyHostToDevice, s));
    if (!any_blur) {       // (with blurred crops the event follows the second staging copy, below)
        HM_CUDA(c, cudaEventRecord(c->crop_event, s));
        c->crop_event_pending = true;
    }
    {
        ProfScope ps(c, s, PC_CROP);
        launch_pdl(crop_warp_normalize_kernel, dim3(256, n), dim3(256), 0, s, frame, H, W, row_stride, c->crop_dev, c->lut, img_out);
        if (any_blur) {    // the anti-aliased crops are overwritten by the float path
            HM_CUDA(c, cudaMemcpyAsync(c->blur_dev, c->blur_host, (size_t)n * sizeof(BlurParams), cudaMemcpyHostToDevice, s));
            launch_pdl(crop_warp_blur_kernel, dim3(256, n), dim3(256), 0, s, frame, H, W, row_stride, (const CropParams*)c->crop_dev,
                       (const BlurParams*)c->blur_dev, make_float3(c->lut_mean[0], c->lut_mean[1], c->lut_mean[2]),
                       make_float3(c->lut_sd[0], c->lut_sd[1], c->lut_sd[2]), img_out);
        }
    }
    HM_CUDA(c, cudaGetLastError());
    if (any_blur) {
        HM_CUDA(c, cudaEventRecord(c->crop_event, s));
        c->crop_event_pending = true;
    }
    HM_CUDA(c, cudaEventRecord(c->crop_kernel_event, s));
    c->crop_kernel_recorded = true;
    return HM_OK;
}

int hm_preprocess(hm_ctx* c, const uint8_t* frame_bgr, int H, int W, int64_t row_stride,
                  int frame_on_device, const double* boxes_xyxy, const uint8_t* is_right, int n,
                  float* img_out, float* box_center, float* box_size, float* trans, void* stream) {
    if (!c) return HM_ERR_INVALID;
    if (!frame_bgr || !boxes_xyxy || !is_right || !img_out || H <= 0 || W <= 0 || row_stride < (int64_t)W * 3)
        return fail(c, HM_ERR_INVALID, "hm_preprocess: bad arguments");
    if (n <= 0) return fail(c, HM_ERR_INVALID, "Invalid detections format");   // infer.py:360-361
    if (n > c->cfg.max_batch) return fail(c, HM_ERR_INVALID, "hm_preprocess: n exceeds max_batch");
    std::vector<double> ccs((size_t)n * 3);
    std::vector<uint8_t> mode((size_t)n);
    for (int i = 0; i < n; ++i) {
        const double* box = boxes_xyxy + 4 * i;
        if (!(box[2] > box[0]) || !(box[3] > box[1])) return fail(c, HM_ERR_INVALID, "hm_preprocess: empty box");
        box_to_center_size(box, &ccs[3 * i], &ccs[3 * i + 1], &ccs[3 * i + 2]);
        mode[i] = is_right[i] ? 0 : 1;                                          // left hands: cv2.flip(patch, 1) (infer.py:229-230)
        if (box_center) { box_center[2 * i] = (float)ccs[3 * i]; box_center[2 * i + 1] = (float)ccs[3 * i + 1]; }
        if (box_size) box_size[i] = (float)ccs[3 * i + 2];
    }
    return preprocess_impl(c, frame_bgr, H, W, row_stride, frame_on_device, ccs.data(), mode.data(), n, img_out, trans,
                           reinterpret_cast<cudaStream_t>(stream));
}

int hm_preprocess_centers(hm_ctx* c, const uint8_t* frame_bgr, int H, int W, int64_t row_stride, int frame_on_device,
                          const double* center_size, const uint8_t* mode, int n, float* img_out, float* trans, void* stream) {
    if (!c) return HM_ERR_INVALID;
    if (!frame_bgr || !center_size || !mode || !img_out || H <= 0 || W <= 0 || row_stride < (int64_t)W * 3)
        return fail(c, HM_ERR_INVALID, "hm_preprocess_centers: bad arguments");
    if (n <= 0 || n > c->cfg.max_batch) return fail(c, HM_ERR_INVALID, "hm_preprocess_centers: n out of range");
    return preprocess_impl(c, frame_bgr, H, W, row_stride, frame_on_device, center_size, mode, n, img_out, trans,
                           reinterpret_cast<cudaStream_t>(stream));
}

int hm_preprocess_blur(hm_ctx* c, const uint8_t* frame_bgr, int H, int W, int64_t row_stride, int frame_on_device,
                       const double* center_size, const uint8_t* mode, const int32_t* blur_radius, const double* blur_weights,
                       int n, float* img_out, float* trans, void* stream) {
    if (!c) return HM_ERR_INVALID;
    if (!frame_bgr || !center_size || !mode || !blur_radius || !blur_weights || !img_out || H <= 0 || W <= 0 ||
        row_stride < (int64_t)W * 3)
        return fail(c, HM_ERR_INVALID, "hm_preprocess_blur: bad arguments");
    if (n <= 0 || n > c->cfg.max_batch) return fail(c, HM_ERR_INVALID, "hm_preprocess_blur: n out of range");
    return preprocess_impl(c, frame_bgr, H, W, row_stride, frame_on_device, center_size, mode, n, img_out, trans,
                           reinterpret_cast<cudaStream_t>(stream), blur_radius, blur_weights);
}

// ---- detector hand-off --------------------------------------------------------------------------------
namespace {
long py_round(double v) { return (long)nearbyint(v); }   // Python's round(): half to even (default FE_TONEAREST)

// letterbox(img, new_shape, auto=True, scaleFill=False, scaleup=True, stride) geometry, yolov7/utils/datasets.py:999-1028
bool letterbox_geom(int H, int W, int new_size, int stride, LetterboxGeom* g) {
    if (H <= 0 || W <= 0 || new_size <= 0 || stride <= 0) return false;
    const double r = std::min((double)new_size / H, (double)new_size / W);
    const long uw = py_round(W * r), uh = py_round(H * r);          // new_unpad = (w, h)
    if (uw <= 0 || uh <= 0) return false;
    double dw = (double)(new_size - uw), dh = (double)(new_size - uh);
    dw = fmod(dw, (double)stride); dh = fmod(dh, (double)stride);   // np.mod of non-negative values
    dw /= 2; dh /= 2;
    const long top = py_round(dh - 0.1), bottom = py_round(dh + 0.1), left = py_round(dw - 0.1), right = py_round(dw + 0.1);
    g->H = H; g->W = W; g->rh = (int)uh; g->rw = (int)uw; g->top = (int)top; g->left = (int)left;
    g->oh = (int)(uh + top + bottom); g->ow = (int)(uw + left + right);
    g->resize = (uw != W || uh != H) ? 1 : 0;
    g->scale_x = (double)W / (double)uw; g->scale_y = (double)H / (double)uh;   // cv::resize: inv_scale = 1 / (dsize / ssize)
    return true;
}
}  // namespace

int hm_letterbox_shape(int H, int W, int new_size, int stride, int32_t* out_h, int32_t* out_w) {
    LetterboxGeom g;
    if (!out_h || !out_w || !letterbox_geom(H, W, new_size, stride, &g)) return HM_ERR_INVALID;
    *out_h = g.oh; *out_w = g.ow;
    return HM_OK;
}

int hm_letterbox(hm_ctx* c, const uint8_t* frame_bgr, int H, int W, int64_t row_stride, int frame_on_device, int new_size,
                 int stride, int out_half, void* img_out, const uint8_t** frame_dev_out, void* stream) {
    if (!c) return HM_ERR_INVALID;
    LetterboxGeom g;
    if (!frame_bgr || !img_out || row_stride < (int64_t)W * 3 || !letterbox_geom(H, W, new_size, stride, &g))
        return fail(c, HM_ERR_INVALID, "hm_letterbox: bad arguments");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    HM_CUDA(c, cudaSetDevice(c->device));
    const uint8_t* frame = frame_bgr;
    if (!frame_on_device) {
        const size_t bytes = (size_t)(H - 1) * (size_t)row_stride + (size_t)W * 3;
        if (!c->frame_dev || bytes > (size_t)c->cfg.max_frame_bytes)
            return fail(c, HM_ERR_INVALID, "hm_letterbox: frame larger than hm_config.max_frame_bytes");
        if (c->crop_kernel_recorded) HM_CUDA(c, cudaStreamWaitEvent(s, c->crop_kernel_event, 0));   // frame_dev may still be read
        HM_CUDA(c, cudaMemcpyAsync(c->frame_dev, frame_bgr, bytes, cudaMemcpyHostToDevice, s));
        frame = c->frame_dev;
    }
    if (frame_dev_out) *frame_dev_out = frame;
    const dim3 grid((g.ow + 127) / 128, g.oh);
    if (out_half) launch_pdl(letterbox_kernel<__half>, grid, dim3(128), 0, s, frame, row_stride, g, reinterpret_cast<__half*>(img_out));
    else launch_pdl(letterbox_kernel<float>, grid, dim3(128), 0, s, frame, row_stride, g, reinterpret_cast<float*>(img_out));
    HM_CUDA(c, cudaGetLastError());
    if (!frame_on_device) {
        HM_CUDA(c, cudaEventRecord(c->crop_kernel_event, s));
        c->crop_kernel_recorded = true;
    }
    return HM_OK;
}

int hm_scale_boxes(hm_ctx* c, const void* det, int det_half, int n, int det_stride, int img1_h, int img1_w, int img0_h,
                   int img0_w, double* boxes_xyxy, uint8_t* is_right, void* stream) {
    if (!c) return HM_ERR_INVALID;
    if (n == 0) return HM_OK;
    if (!det || !boxes_xyxy || !is_right || n < 0 || det_stride < 6 || img1_h <= 0 || img1_w <= 0 || img0_h <= 0 || img0_w <= 0)
        return fail(c, HM_ERR_INVALID, "hm_scale_boxes: bad arguments");
    // scale_coords, ratio_pad=None (general.py:325-327): Python double arithmetic, then used as scalars of tensor ops
    const double gain = std::min((double)img1_h / img0_h, (double)img1_w / img0_w);
    const double pad_x = (img1_w - img0_w * gain) / 2, pad_y = (img1_h - img0_h * gain) / 2;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    HM_CUDA(c, cudaSetDevice(c->device));
    const dim3 grid((n + 127) / 128);
    if (det_half)
        launch_pdl(scale_boxes_kernel<__half>, grid, dim3(128), 0, s, reinterpret_cast<const __half*>(det), n, det_stride,
                   (float)pad_x, (float)pad_y, (float)gain, (float)img0_w, (float)img0_h, boxes_xyxy, is_right);
    else
        launch_pdl(scale_boxes_kernel<float>, grid, dim3(128), 0, s, reinterpret_cast<const float*>(det), n, det_stride,
                   (float)pad_x, (float)pad_y, (float)gain, (float)img0_w, (float)img0_h, boxes_xyxy, is_right);
    HM_CUDA(c, cudaGetLastError());
    return HM_OK;
}

int hm_profile_enable(hm_ctx* c, int on) {
    if (!c) return HM_ERR_INVALID;
    c->prof_on = on != 0;
    c->prof_used = 0;
    c->prof_spans.clear();
    return HM_OK;
}

int hm_profile_collect(hm_ctx* c, double* ms_per_cat, int32_t* spans_per_cat, int ncat) {
    if (!c || !ms_per_cat || !spans_per_cat || ncat < PC_COUNT) return fail(c, HM_ERR_INVALID, "hm_profile_collect: need 11 categories");
    HM_CUDA(c, cudaSetDevice(c->device));
    HM_CUDA(c, cudaDeviceSynchronize());
    for (int i = 0; i < ncat; ++i) { ms_per_cat[i] = 0; spans_per_cat[i] = 0; }
    for (const auto& sp : c->prof_spans) {
        float ms = 0.f;
        HM_CUDA(c, cudaEventElapsedTime(&ms, sp.a, sp.b));
        ms_per_cat[sp.cat] += ms;
        spans_per_cat[sp.cat] += 1;
    }
    c->prof_used = 0;
    c->prof_spans.clear();
    return HM_OK;
}

// Camera post-processing of estimate_from_rgb (infer.py:382-528 -> utils/renderer.py:24-74 cam_crop_to_full and the
// full-image projection), one thread per (hand, keypoint).  Every operation is a separately rounded fp32 op in the
// reference's order, so the result is bit-identical to the PyTorch expressions.
__global__ void camera_postprocess_kernel(const float* __restrict__ pred_cam, float* __restrict__ kp3d,
                                          const float* __restrict__ meta, int meta_stride, int fxfy_differ,
                                          int add_center, int n, float* __restrict__ cam_t_full,
                                          float* __restrict__ kp2d_full) {
    ptx::pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int h = i / 21, k = i - h * 21;
    if (h >= n) return;
    const float* mt = meta + (size_t)h * meta_stride;
    const float flip = mt[5], fx = mt[12], fy = mt[13], cx = mt[14], cy = mt[15];
    const float c0 = pred_cam[h * 3], c1 = __fmul_rn(pred_cam[h * 3 + 1], __fsub_rn(1.0f, __fmul_rn(2.0f, flip))),
                c2 = pred_cam[h * 3 + 2];
    const float bs = __fadd_rn(__fmul_rn(mt[2], c0), 1e-9f);
    const float tz = __fdiv_rn(__fmul_rn(2.0f, fx), bs);
    const float tx = __fadd_rn(__fdiv_rn(__fmul_rn(2.0f, __fsub_rn(mt[0], cx)), bs), c1);
    float ty = __fadd_rn(__fdiv_rn(__fmul_rn(2.0f, __fsub_rn(mt[1], cy)), bs), c2);
    if (fxfy_differ) ty = __fmul_rn(ty, __fdiv_rn(fx, fy));
    if (k == 0) { cam_t_full[h * 3] = tx; cam_t_full[h * 3 + 1] = ty; cam_t_full[h * 3 + 2] = tz; }
    float* p = kp3d + ((size_t)h * 21 + k) * 3;
    const float x = __fmul_rn(p[0], flip), y = p[1], z = p[2];
    p[0] = x;                                                                   // infer.py:392 (in place)
    const float X = __fadd_rn(x, tx), Y = __fadd_rn(y, ty), Z = __fadd_rn(z, tz);
    float u, v;
    if (add_center) {                                                           // intrinsics given (infer.py:474-486)
        const float depth = __fadd_rn(Z, 1e-9f);
        u = __fadd_rn(__fmul_rn(__fdiv_rn(X, depth), fx), cx);
        v = __fadd_rn(__fmul_rn(__fdiv_rn(Y, depth), fy), cy);
    } else {                                                                    // perspective_projection, zero centre
        u = __fmul_rn(__fdiv_rn(X, Z), fx);
        v = __fmul_rn(__fdiv_rn(Y, Z), fy);
    }
    kp2d_full[((size_t)h * 21 + k) * 2] = u;
    kp2d_full[((size_t)h * 21 + k) * 2 + 1] = v;
}

// Rotation matrix -> axis-angle (the log map), one thread per matrix, double arithmetic.  Replaces the matrix -> vector
// branch of cv2.Rodrigues, which the reference calls once per 3 x 3 on the host (matrix_to_axis_angle, hamer/infer.py:1082-1096,
// 16 calls per hand).  cv2 takes r = (R32 - R23, R13 - R31, R21 - R12), s = |r| / 2, theta = acos((tr R - 1) / 2) and returns
// r theta / (2 s): that divides the float32 rounding of R (6e-8) by sin(theta), i.e. it loses the axis near theta = pi.  Here the
// rotation goes through its unit quaternion (largest-pivot branch, Shepperd), theta = 2 atan2(|v|, w): the same vector as cv2
// to ~1e-7 away from pi and a vector that still reproduces R to 1e-7 next to pi (tests/test_gpu_api.py).  OpenCV's SVD
// re-orthogonalisation of the input is not needed: the decoder's matrices are orthonormal by construction (Gram-Schmidt in
// head_finalize_kernel).
__global__ void rotmat_log_kernel(const float* __restrict__ R, int n, float* __restrict__ out) {
    ptx::pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* m = R + (size_t)i * 9;
    const double r00 = m[0], r01 = m[1], r02 = m[2], r10 = m[3], r11 = m[4], r12 = m[5], r20 = m[6], r21 = m[7], r22 = m[8];
    const double tr = r00 + r11 + r22;
    double w, x, y, z;
    if (tr > 0.0) {
        const double S = 2.0 * sqrt(tr + 1.0);
        w = 0.25 * S; x = (r21 - r12) / S; y = (r02 - r20) / S; z = (r10 - r01) / S;
    } else if (r00 > r11 && r00 > r22) {
        const double S = 2.0 * sqrt(fmax(1.0 + r00 - r11 - r22, 0.0));
        w = (r21 - r12) / S; x = 0.25 * S; y = (r01 + r10) / S; z = (r02 + r20) / S;
    } else if (r11 > r22) {
        const double S = 2.0 * sqrt(fmax(1.0 + r11 - r00 - r22, 0.0));
        w = (r02 - r20) / S; x = (r01 + r10) / S; y = 0.25 * S; z = (r12 + r21) / S;
    } else {
        const double S = 2.0 * sqrt(fmax(1.0 + r22 - r00 - r11, 0.0));
        w = (r10 - r01) / S; x = (r02 + r20) / S; y = (r12 + r21) / S; z = 0.25 * S;
    }
    if (w < 0.0) { w = -w; x = -x; y = -y; z = -z; }              // theta in [0, pi]
    const double nv = sqrt(x * x + y * y + z * z);
    const double k = nv > 1e-300 ? 2.0 * atan2(nv, w) / nv : 2.0;  // theta / sin(theta / 2) -> 2 as theta -> 0
    out[(size_t)i * 3] = (float)(x * k); out[(size_t)i * 3 + 1] = (float)(y * k); out[(size_t)i * 3 + 2] = (float)(z * k);
}

int hm_rotmat_to_axis_angle(hm_ctx* c, const float* rotmats, int n, float* axis_angle, void* stream) {
    if (!c) return HM_ERR_INVALID;
    if (!rotmats || !axis_angle || n <= 0) return fail(c, HM_ERR_INVALID, "hm_rotmat_to_axis_angle: bad arguments");
    HM_CUDA(c, cudaSetDevice(c->device));
    HM_CUDA(c, launch_pdl(rotmat_log_kernel, dim3((n + 127) / 128), dim3(128), 0, reinterpret_cast<cudaStream_t>(stream),
                          rotmats, n, axis_angle));
    return HM_OK;
}

int hm_camera_postprocess(hm_ctx* c, const float* pred_cam, float* kp3d, const float* meta, int meta_stride,
                          int fxfy_differ, int add_center, int n, float* cam_t_full, float* kp2d_full, void* stream) {
    if (!c) return HM_ERR_INVALID;
    if (!pred_cam || !kp3d || !meta || !cam_t_full || !kp2d_full || n <= 0 || meta_stride < 16)
        return fail(c, HM_ERR_INVALID, "hm_camera_postprocess: bad arguments");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    HM_CUDA(c, cudaSetDevice(c->device));
    HM_CUDA(c, launch_pdl(camera_postprocess_kernel, dim3((n * 21 + 127) / 128), dim3(128), 0, s, pred_cam, kp3d, meta,
                          meta_stride, fxfy_differ, add_center, n, cam_t_full, kp2d_full));
    return HM_OK;
}

int hm_backbone(hm_ctx* c, const float* img, int B, float* feats, void* stream) {
    int rc = check_ready(c, B, true, false);
    if (rc) return rc;
    if (!img || !feats) return fail(c, HM_ERR_INVALID, "hm_backbone: null pointer");
    HM_CUDA(c, cudaSetDevice(c->device));
    PdlScope pdl(B);
    return backbone_impl(c, img, B, feats, reinterpret_cast<cudaStream_t>(stream));
}

int hm_mano_head(hm_ctx* c, const float* feats, int B, const hm_outputs* out, void* stream) {
    int rc = check_ready(c, B, true, false);
    if (rc) return rc;
    if (!feats || !out) return fail(c, HM_ERR_INVALID, "hm_mano_head: null pointer");
    HM_CUDA(c, cudaSetDevice(c->device));
    PdlScope pdl(B);
    return head_impl(c, feats, B, out, reinterpret_cast<cudaStream_t>(stream));
}

int hm_mano_lbs(hm_ctx* c, const float* betas, const float* rotmats, const float* cam_t, int B,
                float* verts, float* joints21, float* kp2d, void* stream) {
    int rc = check_ready(c, B, false, true);
    if (rc) return rc;
    if (!betas || !rotmats || !verts || !joints21) return fail(c, HM_ERR_INVALID, "hm_mano_lbs: null pointer");
    HM_CUDA(c, cudaSetDevice(c->device));
    return lbs_impl(c, betas, rotmats, cam_t, B, verts, joints21, kp2d, reinterpret_cast<cudaStream_t>(stream));
}

int hm_forward(hm_ctx* c, const float* img, int B, const hm_outputs* out, void* stream) {
    int rc = check_ready(c, B, true, true);
    if (rc) return rc;
    if (!img || !out) return fail(c, HM_ERR_INVALID, "hm_forward: null pointer");
    if (!out->pred_vertices || !out->pred_keypoints_3d || !out->pred_keypoints_2d)
        return fail(c, HM_ERR_INVALID, "hm_forward: pred_vertices / pred_keypoints_3d / pred_keypoints_2d are required");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    HM_CUDA(c, cudaSetDevice(c->device));
    float* feats = out->conditioning_feats ? out->conditioning_feats : c->feats;
    PdlScope pdl(B);
    if ((rc = backbone_impl(c, img, B, feats, s))) return rc;
    if ((rc = head_impl(c, feats, B, out, s))) return rc;
    return lbs_impl(c, c->hbetas, c->hrot, c->hcamt, B, out->pred_vertices, out->pred_keypoints_3d,
                    out->pred_keypoints_2d, s);
}

// ---- multi-GPU: one context per rank / GPU, crops sharded by the caller, one all-gather of the per-hand records --------
#define HM_NCCL(ctx, call)                                                                                    \
    do {                                                                                                      \
        ncclResult_t r_ = (call);                                                                             \
        if (r_ != ncclSuccess)                                                                                \
            return fail(ctx, HM_ERR_CUDA, std::string(#call) + ": " + nccl_api().GetErrorString(r_));         \
    } while (0)

int hm_comm_unique_id(uint8_t* id_out) {
    if (!id_out) return fail(nullptr, HM_ERR_INVALID, "hm_comm_unique_id: null pointer");
    NcclApi& api = nccl_api();
    if (!api.ok) return fail(nullptr, HM_ERR_UNSUPPORTED, api.error);
    static_assert(sizeof(ncclUniqueId) == HM_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    ncclResult_t r = api.GetUniqueId(&id);
    if (r != ncclSuccess) return fail(nullptr, HM_ERR_CUDA, std::string("ncclGetUniqueId: ") + api.GetErrorString(r));
    memcpy(id_out, &id, sizeof(id));
    return HM_OK;
}

int hm_comm_init(hm_ctx* c, const uint8_t* id, int nranks, int rank) {
    if (!c) return HM_ERR_INVALID;
    if (!id || nranks < 1 || rank < 0 || rank >= nranks) return fail(c, HM_ERR_INVALID, "hm_comm_init: bad arguments");
    NcclApi& api = nccl_api();
    if (!api.ok) return fail(c, HM_ERR_UNSUPPORTED, api.error);
    HM_CUDA(c, cudaSetDevice(c->device));
    if (c->comm) { api.CommDestroy(c->comm); c->comm = nullptr; }
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    HM_NCCL(c, api.CommInitRank(&c->comm, nranks, uid, rank));
    c->comm_ranks = nranks;
    return HM_OK;
}

int hm_comm_destroy(hm_ctx* c) {
    if (!c) return HM_ERR_INVALID;
    if (c->comm) { nccl_api().CommDestroy(c->comm); c->comm = nullptr; c->comm_ranks = 0; }
    return HM_OK;
}

int hm_allgather(hm_ctx* c, void* nccl_comm, const float* local_records, float* global_records, int B_local, void* stream) {
    if (!c) return HM_ERR_INVALID;
    if (!local_records || !global_records || B_local <= 0) return fail(c, HM_ERR_INVALID, "hm_allgather: bad arguments");
    NcclApi& api = nccl_api();
    if (!api.ok) return fail(c, HM_ERR_UNSUPPORTED, api.error);
    ncclComm_t comm = nccl_comm ? reinterpret_cast<ncclComm_t>(nccl_comm) : c->comm;
    if (!comm) return fail(c, HM_ERR_STATE, "hm_allgather: no communicator (hm_comm_init, or pass an ncclComm_t)");
    HM_CUDA(c, cudaSetDevice(c->device));
    HM_NCCL(c, api.AllGather(local_records, global_records, (size_t)B_local * HM_RECORD_FLOATS, ncclFloat, comm,
                             reinterpret_cast<cudaStream_t>(stream)));
    return HM_OK;
}

int hm_forward_launch_count(const hm_ctx* ctx) {
    // im2col + patch GEMM + 32 x (LN, qkv, attention, proj, LN, fc1, fc2) + last LN
    // + decoder (token init + 6 x 8 + read-out + finalize) + MANO (one fused kernel; 3 on the three-kernel path)
    return 2 + DEPTH * 7 + 1 + (1 + HDEPTH * 8 + 2) + ((ctx && (ctx->mano_legacy || !ctx->mano_fused_ok)) ? 3 : 1);
}

// ---- RootNet / SAR root depth ---------------------------------------------------------------------------
namespace {

// the 36 convolutions of torchvision's ResNet-34 in execution order, with their state-dict prefixes inside SARresnet34
// (rootnet/Model_RGB.py:179-197: extract_mid = [conv1, bn1, relu, maxpool, layer1, layer2], extract_high[0] = [layer3, layer4])
struct RnSpec { std::string conv, bn; int cin, cout, k, stride, pad, hin; };
std::vector<RnSpec> rn_specs() {
    std::vector<RnSpec> v;
    v.push_back({"extract_mid.0", "extract_mid.1", 3, 64, 7, 2, 3, 256});
    const char* layer_prefix[4] = {"extract_mid.4", "extract_mid.5", "extract_high.0.0", "extract_high.0.1"};
    const int blocks[4] = {3, 4, 6, 3}, width[4] = {64, 128, 256, 512};
    int cin = 64, h = 64;
    for (int l = 0; l < 4; ++l)
        for (int b = 0; b < blocks[l]; ++b) {
            const std::string p = std::string(layer_prefix[l]) + "." + std::to_string(b);
            const int stride = (b == 0 && l > 0) ? 2 : 1, cout = width[l];
            v.push_back({p + ".conv1", p + ".bn1", cin, cout, 3, stride, 1, h});
            v.push_back({p + ".conv2", p + ".bn2", cout, cout, 3, 1, 1, h / stride});
            if (stride == 2) v.push_back({p + ".downsample.0", p + ".downsample.1", cin, cout, 1, 2, 0, h});
            cin = cout;
            h /= stride;
        }
    return v;
}

cudaError_t rn_alloc_bytes(hm_ctx* c, void** p, size_t bytes) {
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaSuccess) c->rn_allocs.push_back(*p);
    return e;
}
#define rn_alloc(ctx, pp, count) rn_alloc_bytes(ctx, reinterpret_cast<void**>(pp), (size_t)(count) * sizeof(**(pp)))

}  // namespace

int hm_rootnet_finalize(hm_ctx* c) {
    if (!c) return HM_ERR_INVALID;
    HM_CUDA(c, cudaSetDevice(c->device));
    const std::vector<RnSpec> specs = rn_specs();
    if ((int)specs.size() != RN_NCONV) return fail(c, HM_ERR_INVALID, "hm_rootnet_finalize: internal layer table");
    auto get = [&](const std::string& n, std::initializer_list<int64_t> shape, std::vector<float>& h) -> int {
        const DevTensor* t = find(c, "rootnet.backbone." + n);
        if (!shape_is(t, shape)) return fail(c, HM_ERR_INVALID, "missing or mis-shaped weight: rootnet.backbone." + n);
        h.resize(t->numel);
        HM_CUDA(c, cudaMemcpy(h.data(), t->f32, t->numel * 4, cudaMemcpyDeviceToHost));
        return HM_OK;
    };
    for (void* p : c->rn_allocs) cudaFree(p);
    c->rn_allocs.clear();
    c->rn_ready = false;
    // workspace for RN_CHUNK hands: crops, im2col rows (largest: conv1, 16,384 x 192 per hand), four NHWC activation buffers
    // (largest: the stem output, 16,384 x 128 per hand)
    HM_CUDA(c, rn_alloc(c, &c->rn_img, (size_t)RN_CHUNK * 3 * 65536));
    HM_CUDA(c, rn_alloc(c, &c->rn_col, (size_t)RN_CHUNK * 16384 * 192));
    HM_CUDA(c, cudaMemset(c->rn_col, 0, (size_t)RN_CHUNK * 16384 * 192 * 2));
    for (int i = 0; i < 4; ++i) {
        HM_CUDA(c, rn_alloc(c, &c->rn_act[i], (size_t)RN_CHUNK * 16384 * 128));
        HM_CUDA(c, cudaMemset(c->rn_act[i], 0, (size_t)RN_CHUNK * 16384 * 128 * 2));
    }
    HM_CUDA(c, rn_alloc(c, &c->rn_lut, (size_t)768));
    HM_CUDA(c, rn_alloc(c, &c->rn_wd, (size_t)512));
    HM_CUDA(c, rn_alloc(c, &c->rn_bd, (size_t)1));
    HM_CUDA(c, rn_alloc(c, &c->rn_kval, (size_t)RN_CHUNK));
    {   // ToTensor + Normalize (Model_RGB.py:317-318) as a table: (v / 255 - mean) / std, torchvision's float32 operations
        const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
        float lut[768];
        for (int ch = 0; ch < 3; ++ch)
            for (int v = 0; v < 256; ++v) lut[ch * 256 + v] = ((float)v / 255.0f - mean[ch]) / stdv[ch];
        HM_CUDA(c, cudaMemcpy(c->rn_lut, lut, sizeof(lut), cudaMemcpyHostToDevice));
    }
    std::vector<float> w, g, b, mu, var;
    for (int i = 0; i < RN_NCONV; ++i) {
        const RnSpec& sp = specs[i];
        hm_ctx::RnConv& cv = c->rn[i];
        cv.cin = sp.cin; cv.cout = sp.cout; cv.k = sp.k; cv.stride = sp.stride; cv.pad = sp.pad;
        cv.hin = sp.hin; cv.hout = (sp.hin + 2 * sp.pad - sp.k) / sp.stride + 1;
        const int K = sp.k * sp.k * sp.cin;
        cv.kpad = (K + 63) / 64 * 64;
        cv.npad = (sp.cout + 127) / 128 * 128;
        int rc;
        if ((rc = get(sp.conv + ".weight", {sp.cout, sp.cin, sp.k, sp.k}, w)) || (rc = get(sp.bn + ".weight", {sp.cout}, g)) ||
            (rc = get(sp.bn + ".bias", {sp.cout}, b)) || (rc = get(sp.bn + ".running_mean", {sp.cout}, mu)) ||
            (rc = get(sp.bn + ".running_var", {sp.cout}, var)))
            return rc;
        // BatchNorm2d (eval, eps 1e-5) folded: w' = w gamma / sqrt(var + eps), b' = beta - mean gamma / sqrt(var + eps);
        // K order (ky, kx, c) to match the NHWC im2col; fp16 operands, padded rows / columns are zero
        std::vector<uint16_t> w16((size_t)cv.npad * cv.kpad, 0);
        std::vector<float> bias(cv.npad, 0.f);
        for (int o = 0; o < sp.cout; ++o) {
            const double sc = (double)g[o] / sqrt((double)var[o] + 1e-5);
            bias[o] = (float)((double)b[o] - (double)mu[o] * sc);
            for (int ci = 0; ci < sp.cin; ++ci)
                for (int ky = 0; ky < sp.k; ++ky)
                    for (int kx = 0; kx < sp.k; ++kx) {
                        const float v = (float)((double)w[(((size_t)o * sp.cin + ci) * sp.k + ky) * sp.k + kx] * sc);
                        const __half hv = __float2half_rn(v);
                        w16[(size_t)o * cv.kpad + (size_t)(ky * sp.k + kx) * sp.cin + ci] = *reinterpret_cast<const uint16_t*>(&hv);
                    }
        }
        HM_CUDA(c, rn_alloc(c, &cv.w, w16.size()));
        HM_CUDA(c, rn_alloc(c, &cv.bias, bias.size()));
        HM_CUDA(c, cudaMemcpy(cv.w, w16.data(), w16.size() * 2, cudaMemcpyHostToDevice));
        HM_CUDA(c, cudaMemcpy(cv.bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice));
        const uint64_t rows = (uint64_t)RN_CHUNK * cv.hout * cv.hout;
        if (!make_tmap_bf16_2d(&cv.tm_w, cv.w, cv.npad, cv.kpad, cv.kpad, 128) ||
            !make_tmap_bf16_2d(&cv.tm_a, c->rn_col, rows, cv.kpad, cv.kpad, GEMM_BM))
            return fail(c, HM_ERR_CUDA, "hm_rootnet_finalize: tensor-map encoding failed");
    }
    {   // ResRootNet.depth_layer (Model_RGB.py:256-262): Conv2d(512, 1, 1)
        const DevTensor* wd = find(c, "rootnet.rootnet.depth_layer.weight");
        const DevTensor* bd = find(c, "rootnet.rootnet.depth_layer.bias");
        if (!shape_is(wd, {1, 512, 1, 1}) || !shape_is(bd, {1})) return fail(c, HM_ERR_INVALID, "missing or mis-shaped weight: rootnet.rootnet.depth_layer");
        HM_CUDA(c, cudaMemcpy(c->rn_wd, wd->f32, 512 * 4, cudaMemcpyDeviceToDevice));
        HM_CUDA(c, cudaMemcpy(c->rn_bd, bd->f32, 4, cudaMemcpyDeviceToDevice));
    }
    c->rn_ready = true;
    return HM_OK;
}

namespace {

// one convolution of the chunk: im2col of `in` (NHWC, pitch ld_in) into the shared column buffer, then the GEMM
int rn_conv(hm_ctx* c, int idx, const __half* in, int ld_in, int n, int epi, __half* out, const __half* resid, cudaStream_t s) {
    const hm_ctx::RnConv& cv = c->rn[idx];
    const int M = n * cv.hout * cv.hout;
    launch_pdl(rn_im2col_kernel, dim3((M + 7) / 8), dim3(256), 0, s, in, c->rn_col, n, cv.hin, cv.hin, ld_in, cv.cin, cv.k, cv.stride,
               cv.pad, cv.hout, cv.hout);
    GemmParams p{};
    p.M = M; p.N = cv.npad; p.K = cv.kpad; p.bias = cv.bias; p.out = out; p.ldo = cv.npad; p.resid = resid;
    p.b_format = 0; p.ksplit = -1;
    HM_CUDA(c, gemm_launch(GemmPlan{1, 128}, epi, cv.tm_a, cv.tm_w, cv.tm_a, p, c->num_sms, s));
    return HM_OK;
}

}  // namespace

int hm_rootnet_depth(hm_ctx* c, const uint8_t* frame_bgr, int H, int W, int64_t row_stride, int frame_on_device,
                     const float* bbox_xywh, const float* k_value, int n, float* depth, void* stream) {
    if (!c) return HM_ERR_INVALID;
    if (!c->rn_ready) return fail(c, HM_ERR_STATE, "hm_rootnet_depth: call hm_rootnet_finalize after setting the rootnet.* weights");
    if (!frame_bgr || !bbox_xywh || !k_value || !depth || H <= 0 || W <= 0 || row_stride < (int64_t)W * 3 || n <= 0)
        return fail(c, HM_ERR_INVALID, "hm_rootnet_depth: bad arguments");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    HM_CUDA(c, cudaSetDevice(c->device));
    const uint8_t* frame = frame_bgr;
    if (!frame_on_device) {
        const size_t bytes = (size_t)(H - 1) * (size_t)row_stride + (size_t)W * 3;
        if (!c->frame_dev || bytes > (size_t)c->cfg.max_frame_bytes)
            return fail(c, HM_ERR_INVALID, "hm_rootnet_depth: frame larger than hm_config.max_frame_bytes");
        if (c->crop_kernel_recorded) HM_CUDA(c, cudaStreamWaitEvent(s, c->crop_kernel_event, 0));
        HM_CUDA(c, cudaMemcpyAsync(c->frame_dev, frame_bgr, bytes, cudaMemcpyHostToDevice, s));
        frame = c->frame_dev;
    }
    int rc;
    for (int lo = 0; lo < n; lo += RN_CHUNK) {
        const int m = std::min(RN_CHUNK, n - lo);
        if (m > c->cfg.max_batch) return fail(c, HM_ERR_INVALID, "hm_rootnet_depth: max_batch too small");
        if (c->crop_event_pending) {          // the pinned crop staging is shared with hm_preprocess
            HM_CUDA(c, cudaEventSynchronize(c->crop_event));
            c->crop_event_pending = false;
        }
        for (int i = 0; i < m; ++i) {
            // generate_patch_image (rootnet/preprocessing.py:62-73): float32 source points, LU solve in double, the matrix cast
            // to float32 before cv2.warpAffine inverts it in double
            const float* bb = bbox_xywh + 4 * (lo + i);
            const float cx = bb[0] + 0.5f * bb[2], cy = bb[1] + 0.5f * bb[3];      // numpy float32 arithmetic
            const float hw = (float)((double)bb[2] * 0.5), hh = (float)((double)bb[3] * 0.5);
            if (!(bb[2] > 0.f) || !(bb[3] > 0.f)) return fail(c, HM_ERR_INVALID, "hm_rootnet_depth: empty box");
            const double sx[3] = {cx, cx, (double)(cx + hw)}, sy[3] = {cy, (double)(cy + hh), cy};
            const double dx[3] = {128, 128, 256}, dy[3] = {128, 256, 128};
            double A[6][6] = {}, bv[6];
            for (int q = 0; q < 3; ++q) {
                A[2 * q][0] = sx[q]; A[2 * q][1] = sy[q]; A[2 * q][2] = 1;
                A[2 * q + 1][3] = sx[q]; A[2 * q + 1][4] = sy[q]; A[2 * q + 1][5] = 1;
                bv[2 * q] = dx[q]; bv[2 * q + 1] = dy[q];
            }
            lu_solve6(A, bv);
            double M32[6];
            for (int q = 0; q < 6; ++q) M32[q] = (double)(float)bv[q];
            invert_affine(M32, c->crop_host[i].m);
            c->crop_host[i].flip = 0;
            c->crop_host[i].pad = 0;
        }
        HM_CUDA(c, cudaMemcpyAsync(c->crop_dev, c->crop_host, (size_t)m * sizeof(CropParams), cudaMemcpyHostToDevice, s));
        HM_CUDA(c, cudaMemcpyAsync(c->rn_kval, k_value + lo, (size_t)m * 4, cudaMemcpyHostToDevice, s));
        HM_CUDA(c, cudaEventRecord(c->crop_event, s));
        c->crop_event_pending = true;
        launch_pdl(crop_warp_normalize_kernel, dim3(256, m), dim3(256), 0, s, frame, H, W, row_stride, (const CropParams*)c->crop_dev,
                   (const float*)c->rn_lut, c->rn_img);
        // stem: conv1 + bn1 + relu, maxpool
        launch_pdl(rn_im2col_conv1_kernel, dim3(128, m), dim3(128), 0, s, (const float*)c->rn_img, c->rn_col);
        {
            const hm_ctx::RnConv& cv = c->rn[0];
            GemmParams p{};
            p.M = m * 16384; p.N = cv.npad; p.K = cv.kpad; p.bias = cv.bias; p.out = c->rn_act[0]; p.ldo = cv.npad; p.ksplit = -1;
            HM_CUDA(c, gemm_launch(GemmPlan{1, 128}, EPI_BIAS_RELU_F16, cv.tm_a, cv.tm_w, cv.tm_a, p, c->num_sms, s));
        }
        {
            const size_t total = (size_t)m * 64 * 64 * (128 / 8);
            launch_pdl(rn_maxpool_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, s, (const __half*)c->rn_act[0], c->rn_act[1], m, 128, 128, 128);
        }
        // BasicBlocks: x -> relu(bn1(conv1 x)) -> bn2(conv2 .) + identity -> relu; identity = x or bn(conv1x1 stride 2 (x))
        __half *x = c->rn_act[1], *t = c->rn_act[2], *y = c->rn_act[3], *ds = c->rn_act[0];
        int ld = 128, idx = 1;
        const int blocks[4] = {3, 4, 6, 3};
        for (int l = 0; l < 4; ++l)
            for (int b = 0; b < blocks[l]; ++b) {
                const bool down = b == 0 && l > 0;
                const int ld_out = c->rn[idx].npad;
                if ((rc = rn_conv(c, idx, x, ld, m, EPI_BIAS_RELU_F16, t, nullptr, s))) return rc;
                const __half* identity = x;
                if (down) {
                    if ((rc = rn_conv(c, idx + 2, x, ld, m, EPI_BIAS_BF16, ds, nullptr, s))) return rc;
                    identity = ds;
                }
                if ((rc = rn_conv(c, idx + 1, t, ld_out, m, EPI_BIAS_RESID_RELU_F16, y, identity, s))) return rc;
                std::swap(x, y);
                ld = ld_out;
                idx += down ? 3 : 2;
            }
        launch_pdl(rn_head_kernel, dim3(m), dim3(512), 0, s, (const __half*)x, (const float*)c->rn_wd, (const float*)c->rn_bd,
                   (const float*)c->rn_kval, depth + lo);
        HM_CUDA(c, cudaGetLastError());
        HM_CUDA(c, cudaEventRecord(c->crop_kernel_event, s));
        c->crop_kernel_recorded = true;
    }
    return HM_OK;
}

// test hook: the normalised 256 x 256 patches of the last hm_rootnet_depth chunk (DEVICE out [n,3,256,256], n <= 8)
int hm_test_rootnet_patches(hm_ctx* c, int n, float* out, void* stream) {
    if (!c || !c->rn_ready || !out || n <= 0 || n > RN_CHUNK) return fail(c, HM_ERR_INVALID, "hm_test_rootnet_patches: bad arguments");
    HM_CUDA(c, cudaMemcpyAsync(out, c->rn_img, (size_t)n * 3 * 65536 * 4, cudaMemcpyDeviceToDevice, reinterpret_cast<cudaStream_t>(stream)));
    return HM_OK;
}

// ---- test hooks ---------------------------------------------------------------------------------
int hm_test_gemm(hm_ctx* c, int cta_group, int bn, int epi, int w_format, const uint16_t* A,
                 const uint16_t* W, const float* bias, const float* pos, void* out, int M, int N,
                 int K, void* stream) {
    if (!c || !A || !W || !bias || !out) return fail(c, HM_ERR_INVALID, "hm_test_gemm: null pointer");
    HM_CUDA(c, cudaSetDevice(c->device));
    CUtensorMap ta, tb, tout;
    if (!make_tmap_bf16_2d(&ta, A, M, K, K, GEMM_BM) || !make_tmap_bf16_2d(&tb, W, N, K, K, bn / (cta_group == 1 ? 1 : 2)))
        return fail(c, HM_ERR_CUDA, "hm_test_gemm: tensor map encoding failed");
    tout = ta;
    if (epi == EPI_BIAS_RESID_F32 && !make_tmap_f32_out(&tout, out, M, N, N))
        return fail(c, HM_ERR_CUDA, "hm_test_gemm: output tensor map encoding failed");
    GemmParams p{};
    p.M = M; p.N = N; p.K = K; p.bias = bias; p.out = out; p.ldo = N; p.pos = pos; p.pos_rows = TOK;
    p.b_format = w_format ? 1u : 0u;
    GemmPlan plan{cta_group == 1 ? 1 : 2, bn};
    HM_CUDA(c, gemm_launch(plan, epi, ta, tb, tout, p, c->num_sms, reinterpret_cast<cudaStream_t>(stream)));
    return HM_OK;
}

// diagnostics: run hm_mano_lbs once with the fused kernel's phase clocks enabled and copy them to `clocks_out`
// (HOST [n_ctas * 8 * 6] int64; n_ctas = 148).  Layout per CTA and unit (first 8 units): unit start, previous unit done on
// all warps, inputs landed, prologue done, MMA may start, tiles done.
int hm_test_mano_clocks(hm_ctx* c, const float* betas, const float* rotmats, int B, float* verts, float* joints21,
                        long long* clocks_out, int n_ctas, void* stream) {
    int rc = check_ready(c, B, false, true);
    if (rc) return rc;
    if (!clocks_out || n_ctas <= 0) return fail(c, HM_ERR_INVALID, "hm_test_mano_clocks: bad arguments");
    HM_CUDA(c, cudaSetDevice(c->device));
    const size_t bytes = (size_t)n_ctas * 8 * 6 * sizeof(long long);
    HM_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->mano_clk), bytes));
    cudaMemset(c->mano_clk, 0, bytes);
    rc = lbs_impl(c, betas, rotmats, nullptr, B, verts, joints21, nullptr, reinterpret_cast<cudaStream_t>(stream));
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(clocks_out, c->mano_clk, bytes, cudaMemcpyDeviceToHost);
    cudaFree(c->mano_clk);
    c->mano_clk = nullptr;
    if (rc) return rc;
    HM_CUDA(c, e);
    return HM_OK;
}

int hm_test_layernorm(hm_ctx* c, const float* x, const float* gamma, const float* beta, void* y,
                      int rows, float eps, int out_bf16, void* stream) {
    if (!c || !x || !gamma || !beta || !y) return fail(c, HM_ERR_INVALID, "hm_test_layernorm: null pointer");
    HM_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (out_bf16) launch_ln<10, __half>(x, gamma, beta, reinterpret_cast<__half*>(y), rows, eps, s);
    else launch_ln<10, float>(x, gamma, beta, reinterpret_cast<float*>(y), rows, eps, s);
    HM_CUDA(c, cudaGetLastError());
    return HM_OK;
}

int hm_test_attention(hm_ctx* c, int variant, const uint16_t* qkv, uint16_t* out, int B, void* stream) {
    if (!c || !qkv || !out || B <= 0) return fail(c, HM_ERR_INVALID, "hm_test_attention: bad arguments");
    HM_CUDA(c, cudaSetDevice(c->device));
    const float sc = 0.11180339887498948f * 1.4426950408889634f;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (variant == 0)
        vit_attention_kernel<<<dim3(16, B), ATT_THREADS, ATT_SMEM_BYTES, s>>>(
            reinterpret_cast<const __half*>(qkv), reinterpret_cast<__half*>(out), sc);
    else {   // 1: tcgen05 kernel, 2: same + phase clocks of every CTA (16 x int64 each) copied into `out`
        CUtensorMap t128, t32;
        if (!make_tmap_attn(&t128, &t32, qkv, (uint64_t)B * AT5_T))
            return fail(c, HM_ERR_CUDA, "hm_test_attention: tensor map encoding failed");
        __half* o = reinterpret_cast<__half*>(out);
        if (variant == 1)
            launch_pdl(vit_attention_tc5_kernel<false>, dim3(16, B), dim3(AT5_THREADS), AT5_SMEM_BYTES, s, t128, t32, o, sc, nullptr);
        else {
            long long* dbg = reinterpret_cast<long long*>(c->hbuf);
            launch_pdl(vit_attention_tc5_kernel<true>, dim3(16, B), dim3(AT5_THREADS), AT5_SMEM_BYTES, s, t128, t32, o, sc, dbg);
            HM_CUDA(c, cudaMemcpyAsync(out, c->hbuf, (size_t)16 * B * 16 * sizeof(long long), cudaMemcpyDeviceToDevice, s));
        }
    }
    HM_CUDA(c, cudaGetLastError());
    return HM_OK;
}

}  // extern "C"
