/*
 * hamer_b200 - C ABI of the B200-native HaMeR hot path (libhamer_b200.so).
 *
 * This is the drop-in boundary: every entry point replaces a Python/PyTorch interface of the
 * reference (2646207530/hamer-yolo, paths below are relative to its root).  The reference is pure
 * Python, so its "FFI" for this path is a ctypes binding; signatures use plain
 * pointers and sizes only - no torch / C++ types.
 *
 * Conventions
 *   - every function returns 0 on success and a negative hm_status on failure; the message is
 *     available from hm_last_error(ctx).  No C++ exception crosses the ABI.
 *   - the caller owns every tensor; the library owns only packed weights, MANO tables and a
 *     workspace sized at hm_create() for max_batch crops.
 *   - all device work is enqueued on the cudaStream_t passed as `stream` (a void*; 0 = default
 *     stream) without host synchronisation unless stated; hm_forward() performs no allocation and
 *     is CUDA-graph capturable.
 *   - a context is bound to one device and is not thread-safe (one context per rank / stream).
 *   - "device" pointers must be valid on the context's device; "host" pointers are ordinary memory.
 */
#ifndef HAMER_B200_H
#define HAMER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hm_ctx hm_ctx;

typedef enum hm_status {
    HM_OK = 0,
    HM_ERR_INVALID = -1,      /* bad argument / shape / missing weight */
    HM_ERR_CUDA = -2,         /* a CUDA runtime / driver call failed */
    HM_ERR_STATE = -3,        /* weights or MANO tables not set */
    HM_ERR_UNSUPPORTED = -4   /* not an sm_100 device */
} hm_status;

typedef struct hm_config {
    int32_t max_batch;        /* largest B of hm_forward / n of hm_preprocess */
    int32_t max_frame_bytes;  /* largest H*row_stride accepted by hm_preprocess (host frames) */
    int32_t gemm_cta_group;   /* 0 = default (2: CTA-pair UMMA M=256), 1 = single-CTA tiles */
    int32_t attention_legacy; /* 0 = tcgen05/TMEM attention kernel (default), 1 = mma.sync kernel */
    int32_t mano_legacy;      /* 0 = fused tcgen05 blend + skinning kernel (default), 1 = prep / GEMM / skinning kernels */
    int32_t deterministic;    /* 1 = never split the K dimension of the residual GEMMs: fixed fp32 summation order, results
                                 bit-identical across runs and batch sizes (default 0: the partial-wave tiles are split
                                 into K-slices that are reduce-added in arrival order - last-bit differences) */
    int32_t split_weights;    /* 1 = the backbone weight matrices are kept as W_hi + W_lo (two fp16 matrices, ~21 mantissa
                                 bits) and every backbone GEMM runs over K' = 2K: an arbitrary fp32 checkpoint then meets the
                                 1e-3 parameter bar with the margin of fp16-representable weights, at twice the tensor-core
                                 work (default 0: weights rounded once to fp16).  Must be chosen before
                                 the first hm_set_weight. */
    int32_t reserved[1];
} hm_config;

/* Device pointers of the tensors HAMER.forward_step returns (hamer/hamer/models/hamer.py:124-156).
 * All float32, contiguous, leading dimension B.  Any pointer may be NULL to skip that output,
 * except that pred_vertices / pred_keypoints_3d / pred_keypoints_2d need nothing else. */
typedef struct hm_outputs {
    float* pred_cam;           /* [B,3]        output['pred_cam'] */
    float* global_orient;      /* [B,1,3,3]    pred_mano_params['global_orient'] */
    float* hand_pose;          /* [B,15,3,3]   pred_mano_params['hand_pose'] */
    float* betas;              /* [B,10]       pred_mano_params['betas'] */
    float* pred_cam_t;         /* [B,3]        output['pred_cam_t'] (= params['trans']) */
    float* focal_length;       /* [B,2]        output['focal_length'] (5000) */
    float* pred_keypoints_3d;  /* [B,21,3] */
    float* pred_vertices;      /* [B,778,3] */
    float* pred_keypoints_2d;  /* [B,21,2] */
    float* conditioning_feats; /* [B,192,1280] backbone output, optional (NULL to skip) */
} hm_outputs;

/* ---- life cycle ------------------------------------------------------------------------------
 * replaces load_hamer() / HAMER.__init__ (hamer/hamer/models/__init__.py:32-47, models/hamer.py:21-72) */
int hm_create(hm_ctx** out, int device_ordinal, const hm_config* cfg);
void hm_destroy(hm_ctx* ctx);
const char* hm_last_error(const hm_ctx* ctx);   /* ctx may be NULL: last hm_create error */

/* ---- weights ---------------------------------------------------------------------------------
 * One call per state-dict entry of the reference checkpoint (key layout: SURVEY.md section 5;
 * e.g. "backbone.blocks.0.attn.qkv.weight", "mano_head.transformer.transformer.layers.0.1.fn.to_kv.weight",
 * "mano_head.init_hand_pose").  `data` is a HOST pointer to float32 values in the reference's own
 * layout; matrices used by the backbone GEMMs are converted to fp16 once (tensor-core operands,
 * fp32 accumulation; precision budget in DESIGN.md).  Replaces
 * HAMER.load_from_checkpoint(..., strict=False) (models/__init__.py:46). */
int hm_set_weight(hm_ctx* ctx, const char* name, const float* data, const int64_t* shape, int ndim);
/* Checks that every tensor is present and builds the fused tables (positional table, decoder
 * constants).  Must be called once after the last hm_set_weight. */
int hm_finalize_weights(hm_ctx* ctx);

/* MANO model data in the layout smplx.MANOLayer builds (HOST pointers, float32 / int32):
 * v_template[778,3] shapedirs[778,3,10] posedirs[135,2334] J_regressor[16,778] lbs_weights[778,16]
 * parents[16] (root -1) tip_ids[5] joint_map[21].  Replaces MANO.__init__
 * (hamer/hamer/models/mano_wrapper.py:12-30). */
int hm_set_mano(hm_ctx* ctx, const float* v_template, const float* shapedirs, const float* posedirs,
                const float* J_regressor, const float* lbs_weights, const int32_t* parents,
                const int32_t* tip_ids, const int32_t* joint_map);

/* ---- the hot path ----------------------------------------------------------------------------
 * hm_preprocess replaces hamer_inference.prepare_batch_bbox (hamer/infer.py:154-259):
 *   frame_bgr : H x W x 3 uint8, BGR, `row_stride` bytes per row; host memory unless
 *               frame_on_device != 0
 *   boxes_xyxy: HOST [n,4] float64 (x1,y1,x2,y2), is_right: HOST [n] (1 = 'right', 0 = flip)
 *   img_out   : DEVICE [n,3,256,256] float32, normalised RGB        (batch['img'])
 *   box_center: HOST [n,2] f32, box_size: HOST [n] f32, trans: HOST [n,2,3] f32 (batch['trans'],
 *               also returned by the reference as 'inv_trans'); any of the three may be NULL.
 * Synchronises `stream` only if it has to recycle its pinned staging buffer. */
int hm_preprocess(hm_ctx* ctx, const uint8_t* frame_bgr, int H, int W, int64_t row_stride,
                  int frame_on_device, const double* boxes_xyxy, const uint8_t* is_right, int n,
                  float* img_out, float* box_center, float* box_size, float* trans, void* stream);

/* hm_preprocess_centers: the same crop / warp / normalise kernel for callers that derive the crop windows themselves
 * - ViTDetDataset.__getitem__ (hamer/hamer/datasets/vitdet_dataset.py:48-95) computes centre and size in float32 and
 * flips INSIDE the warp (generate_image_patch_cv2 do_flip, hamer/hamer/datasets/utils.py:347-349).
 *   center_size: HOST [n,3] float64 (c_x, c_y, bbox_size) of the square source window; mode: HOST [n]
 *   0 = plain crop, 1 = mirror the patch after the warp (infer.py:229-230), 2 = warp the horizontally mirrored frame
 *   (c_x is then the centre in the mirrored frame, W - c_x - 1).  trans: HOST [n,2,3] f32 or NULL. */
int hm_preprocess_centers(hm_ctx* ctx, const uint8_t* frame_bgr, int H, int W, int64_t row_stride,
                          int frame_on_device, const double* center_size, const uint8_t* mode, int n,
                          float* img_out, float* trans, void* stream);

/* hm_forward replaces HAMER.forward(batch) (hamer/hamer/models/hamer.py:269-277 -> forward_step
 * :99-156): img = batch['img'], DEVICE [B,3,256,256] float32. */
int hm_forward(hm_ctx* ctx, const float* img, int B, const hm_outputs* out, void* stream);

/* Stages of hm_forward, individually callable (same kernels):
 *   hm_backbone  = self.backbone(x[:,:,:,32:-32])          (hamer.py:119; backbones/vit.py:320-339)
 *   hm_mano_head = self.mano_head(feats) + pred_cam_t      (hamer.py:121-134; heads/mano_head.py:61-115)
 *   hm_mano_lbs  = self.mano(...) + perspective_projection (hamer.py:143-152; mano_wrapper.py:32-44)
 * hm_mano_lbs: betas DEVICE [B,10], rotmats DEVICE [B,16,3,3] (global_orient then hand_pose),
 * cam_t DEVICE [B,3] or NULL, verts DEVICE [B,778,3], joints21 DEVICE [B,21,3], kp2d DEVICE
 * [B,21,2] or NULL.  B is not limited by max_batch. */
int hm_backbone(hm_ctx* ctx, const float* img, int B, float* feats, void* stream);
int hm_mano_head(hm_ctx* ctx, const float* feats, int B, const hm_outputs* out, void* stream);
int hm_mano_lbs(hm_ctx* ctx, const float* betas, const float* rotmats, const float* cam_t, int B,
                float* verts, float* joints21, float* kp2d, void* stream);

/* hm_camera_postprocess replaces the camera post-processing of hamer_inference.estimate_from_rgb
 * (hamer/infer.py:382-528: kp3d x-flip :392, cam y-flip :402-409, cam_crop_to_full = hamer/hamer/utils/renderer.py:24-74,
 * projection of the 21 joints into the full image :474-486 / utils/geometry.py:72-109).  All pointers are DEVICE:
 *   pred_cam [n,3]; kp3d [n,21,3] (x is multiplied by do_flip in place, as the reference does);
 *   meta [n, meta_stride] float32 per hand: 0,1 box_center  2 box_size  5 do_flip (1 = left hand)  12 fx  13 fy  14 cx  15 cy
 *   (other columns free for the caller: the Python mirror keeps img_size and the crop transform there);
 *   fxfy_differ: the reference's batch-level `not torch.allclose(fx, fy)`; add_center: 1 = intrinsics given
 *   (u = X/(Z+1e-9)*fx + cx), 0 = default focal (u = X/Z*f, zero camera centre);
 *   outputs cam_t_full [n,3], kp2d_full [n,21,2].  fp32 results are bit-identical to the reference's PyTorch expressions. */
int hm_camera_postprocess(hm_ctx* ctx, const float* pred_cam, float* kp3d, const float* meta, int meta_stride,
                          int fxfy_differ, int add_center, int n, float* cam_t_full, float* kp2d_full, void* stream);

/* hm_preprocess_blur: hm_preprocess_centers with the anti-alias pre-blur of ViTDetDataset.__getitem__
 * (hamer/hamer/datasets/vitdet_dataset.py:62-71) and prepare_item (hamer/infer.py:312-316): when the crop window is more
 * than 2.2 x the patch the reference blurs the frame with skimage.filters.gaussian(sigma = (bbox_size / 256 / 2 - 1) / 2)
 * (= scipy.ndimage.gaussian_filter, mode 'nearest', truncate 4) and warps the float64 result.
 *   blur_radius : HOST [n] int32, int(4 sigma + 0.5) in [0, 32]; 0 = no blur (the crop is the 8-bit one of hm_preprocess)
 *   blur_weights: HOST [n, 33] float64, [i, k] = normalised gaussian weight at offset +-k (k <= radius) - computed by the
 *                 caller with numpy like scipy's _gaussian_kernel1d, so the crop matches the reference bit for bit. */
#define HM_BLUR_MAX_RADIUS 32
int hm_preprocess_blur(hm_ctx* ctx, const uint8_t* frame_bgr, int H, int W, int64_t row_stride, int frame_on_device,
                       const double* center_size, const uint8_t* mode, const int32_t* blur_radius, const double* blur_weights,
                       int n, float* img_out, float* trans, void* stream);

/* ---- detector hand-off on the device (the YOLOv7 detector itself stays upstream) --------------------------------------
 * hm_letterbox replaces LoadImage.process_img (yolo/yolov7/utils/datasets.py:137-141: letterbox(img0, new_size, stride,
 * auto=True), datasets.py:999-1030 = cv2.resize INTER_LINEAR + cv2.copyMakeBorder 114; BGR -> RGB; HWC -> CHW) and the
 * uint8 -> float / 255 of Detector.detect (yolo/detector.py:121-123), bit for bit:
 *   frame_bgr : H x W x 3 uint8 BGR, host unless frame_on_device != 0 (a host frame is uploaded whole into the context's
 *               frame buffer - hm_config.max_frame_bytes - and *frame_dev_out, if not NULL, receives the device copy so
 *               that the following hm_preprocess can run with frame_on_device = 1: the frame crosses PCIe once)
 *   img_out   : DEVICE [3, out_h, out_w] float32 (out_half = 0) or float16 (out_half = 1), RGB, 0..1;
 *               hm_letterbox_shape gives (out_h, out_w) for H, W, new_size (640), stride (32).
 * hm_scale_boxes replaces scale_coords(img.shape[2:], det[:, :4], im0.shape).round() and the right / left labelling
 * (yolo/detector.py:139-145, yolov7/utils/general.py:323-345) in the detections' own precision:
 *   det       : DEVICE [n, det_stride] float32 / float16 rows x1 y1 x2 y2 conf cls (non_max_suppression output)
 *   boxes_xyxy: DEVICE [n,4] float64, is_right: DEVICE [n] (cls == 1) - after one small copy, the arguments of hm_preprocess. */
int hm_letterbox_shape(int H, int W, int new_size, int stride, int32_t* out_h, int32_t* out_w);
int hm_letterbox(hm_ctx* ctx, const uint8_t* frame_bgr, int H, int W, int64_t row_stride, int frame_on_device, int new_size,
                 int stride, int out_half, void* img_out, const uint8_t** frame_dev_out, void* stream);
int hm_scale_boxes(hm_ctx* ctx, const void* det, int det_half, int n, int det_stride, int img1_h, int img1_w, int img0_h,
                   int img0_w, double* boxes_xyxy, uint8_t* is_right, void* stream);

/* hm_rotmat_to_axis_angle replaces matrix_to_axis_angle (hamer/infer.py:1082-1096: cv2.Rodrigues per 3 x 3 on the host, 16
 * calls per hand) for the npy writer's 'theta' / 'pose_hand' / 'pose_global' (infer.py:1279-1303): rotmats DEVICE [n,3,3]
 * float32 -> axis_angle DEVICE [n,3] float32, double arithmetic inside (through the unit quaternion: well conditioned up to
 * theta = pi, where cv2's r theta / (2 sin theta) amplifies the float32 rounding of R); agrees with cv2.Rodrigues to ~1e-6 rad
 * away from pi (cv2's SVD re-orthogonalisation of the input is skipped - the decoder's matrices are orthonormal by
 * construction). */
int hm_rotmat_to_axis_angle(hm_ctx* ctx, const float* rotmats, int n, float* axis_angle, void* stream);

/* ---- RootNet / SAR root depth --------------------------------------------------------------------------------------
 * replaces EstimateRGB.estimate_root_depth_custom (rootnet/Model_RGB.py:572-638), which hamer/infer.py:1275 calls per hand
 * right after estimate_from_rgb: 256 x 256 crop of the process_bbox'ed window (rootnet/preprocessing.py:39-79, :167-188),
 * ToTensor + Normalize, SARresnet34 (torchvision ResNet-34 trunk, Model_RGB.py:179-197), ResRootNet (global average pool,
 * 1 x 1 conv, Model_RGB.py:240-287), depth = gamma * k_value.
 * Weights: hm_set_weight with the reference's state-dict keys under the prefix "rootnet.": "rootnet.backbone.<key>" for the
 * entries of checkpoint['model'] that start with "backbone." (convolution weights and the BatchNorm weight / bias /
 * running_mean / running_var) and "rootnet.rootnet.depth_layer.{weight,bias}" for checkpoint['rootnet'].
 * hm_rootnet_finalize folds the BatchNorms into fp16 convolution operands (tensor cores, fp32 accumulation) and sizes the
 * workspace; call it once after the last rootnet.* weight.
 *   bbox_xywh : HOST [n,4] float32, the PROCESSED boxes (process_bbox(bbox, W, H, (256, 256), 1.5): square windows)
 *   k_value   : HOST [n] float32, calculate_k (Model_RGB.py:494-498) = sqrt(0.3 * 0.3 * fx * fy / (w * h))
 *   depth     : DEVICE [n] float32 */
int hm_rootnet_finalize(hm_ctx* ctx);
int hm_rootnet_depth(hm_ctx* ctx, const uint8_t* frame_bgr, int H, int W, int64_t row_stride, int frame_on_device,
                     const float* bbox_xywh, const float* k_value, int n, float* depth, void* stream);

/* ---- multi-GPU --------------------------------------------------------------------------------
 * The reference runs one process and one hand at a time (hamer/infer.py:1268-1274); crops are independent, so the path
 * shards data-parallel: one context per rank / GPU, each rank runs hm_forward on its crops, and ONE collective gathers the
 * per-hand records (north_star).  A record is HM_RECORD_FLOATS float32 values per hand:
 *   verts 778*3 | kp3d 21*3 | kp2d 21*2 | pred_cam 3 | pred_cam_t 3 | rotmats 16*9 (global_orient, hand_pose) | betas 10
 * = 2599 floats = 10,396 bytes (hamer_yolo_b200/dist.py packs / unpacks it).
 *   hm_comm_unique_id : rank 0 creates the 128-byte NCCL id and hands it to the other ranks by any out-of-band channel
 *   hm_comm_init      : every rank joins (ncclCommInitRank on the context's device); the context owns the communicator
 *   hm_allgather      : ncclAllGather of B_local records per rank into global_records [nranks * B_local][2599] on
 *                       `stream` (no host synchronisation; equal counts on every rank - callers pad ragged shards).
 *                       nccl_comm: an existing ncclComm_t to use instead of the context's own, or NULL.
 * NCCL is resolved at run time (libnccl.so.2 already loaded in the process, else the system library). */
#define HM_COMM_ID_BYTES 128
#define HM_RECORD_FLOATS 2599
int hm_comm_unique_id(uint8_t* id_out /* HOST [HM_COMM_ID_BYTES] */);
int hm_comm_init(hm_ctx* ctx, const uint8_t* id, int nranks, int rank);
int hm_comm_destroy(hm_ctx* ctx);
int hm_allgather(hm_ctx* ctx, void* nccl_comm, const float* local_records, float* global_records, int B_local,
                 void* stream);

/* ---- introspection ---------------------------------------------------------------------------*/
/* Number of kernels hm_forward(B) launches (for bench.py's gpu_launches). */
int hm_forward_launch_count(const hm_ctx* ctx);
/* Library / build identification string. */
const char* hm_version(void);
/* Optional per-category device timing.  While enabled, every kernel launch of the hot path is
 * bracketed by CUDA events recorded on the launch stream (no synchronisation, a few microseconds
 * of host time per launch).  hm_profile_collect synchronises the device, returns the summed
 * elapsed milliseconds and the number of bracketed spans per category, and resets the recorder.
 * Categories (ncat >= 11): 0 im2col, 1 gemm_patch, 2 layernorm, 3 gemm_qkv, 4 attention,
 * 5 gemm_proj, 6 gemm_fc1, 7 gemm_fc2, 8 decoder head (one span per call), 9 mano_lbs, 10 crop. */
int hm_profile_enable(hm_ctx* ctx, int on);
int hm_profile_collect(hm_ctx* ctx, double* ms_per_cat, int32_t* spans_per_cat, int ncat);

/* ---- test hooks (used by tests/ only; same kernels as the hot path) ----------------------------
 * C[M,N] = A[M,K] * W[N,K]^T with the fused epilogue `epi` (0 bias->fp16, 1 bias+GELU->fp16,
 * 2 bias + residual add into f32 out, 3 bias + pos[(m % 192), n] -> f32).  A: DEVICE fp16 bits,
 * W: DEVICE fp16 bits (w_format 0; 1 = both operands bf16), bias/pos: DEVICE f32,
 * out: DEVICE fp16 (epi 0,1) or f32 (epi 2,3). */
int hm_test_rootnet_patches(hm_ctx* ctx, int n, float* out, void* stream);   /* crops of the last hm_rootnet_depth chunk */
int hm_test_gemm(hm_ctx* ctx, int cta_group, int bn, int epi, int w_format, const uint16_t* A,
                 const uint16_t* W, const float* bias, const float* pos, void* out, int M, int N,
                 int K, void* stream);
/* hm_mano_lbs once with the fused kernel's per-unit phase clocks recorded (diagnostics; clocks_out HOST [n_ctas*8*6] int64) */
int hm_test_mano_clocks(hm_ctx* ctx, const float* betas, const float* rotmats, int B, float* verts, float* joints21,
                        long long* clocks_out, int n_ctas, void* stream);
/* y = LayerNorm(x) over rows of 1280 (out_f16 != 0: fp16 output, else f32) */
int hm_test_layernorm(hm_ctx* ctx, const float* x, const float* gamma, const float* beta, void* y,
                      int rows, float eps, int out_f16, void* stream);
/* attention core on a [B*192,3840] fp16 qkv buffer -> [B*192,1280] fp16
 * (variant 0 = mma.sync kernel, 1 = tcgen05/TMEM kernel, 2 = tcgen05 kernel that then overwrites the
 * head of `out` with per-CTA phase clocks: 16 x int64 per (head, crop), diagnostics only) */
int hm_test_attention(hm_ctx* ctx, int variant, const uint16_t* qkv, uint16_t* out, int B, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HAMER_B200_H */
