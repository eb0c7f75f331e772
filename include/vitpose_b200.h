/* vitpose_b200.h -- C ABI of the B200-native ViTPose crop engine (libvitpose_b200.so).
 *
 * One data-parallel hot path of JunkyByte/easy_ViTPose, rebuilt for sm_100a:
 *     crops f32 [B,3,256,192] -> ViT backbone -> TopdownHeatmapSimpleHead -> heatmaps f32 [B,K,64,48]
 *     -> argmax + DARK/UDP refine -> keypoints f32 [B,K,3] rows (y, x, score)
 * Plain pointers and sizes only; no torch types.  Device pointers are CUDA device addresses on the
 * engine's device, `stream` is a cudaStream_t passed as void* (NULL = default stream).  All calls are
 * asynchronous on `stream` unless stated; none of them frees or keeps caller memory.
 * There is no CPU fallback: every entry point fails (non-zero + vpb_last_error()) without an sm_100 GPU.
 *
 * Concurrency contract.  An engine owns ONE activation workspace, two host-staging slots and its CUDA graphs:
 *   - calls on one engine must come from one host thread at a time (the handle holds no lock);
 *   - calls on the same engine from DIFFERENT streams are safe and are executed one after the other: every entry point
 *     waits (on the device, cudaStreamWaitEvent) for the engine's previous enqueue when the stream changes -- there is no
 *     overlap between two calls of one engine; use one engine per stream (or per GPU) for concurrency;
 *   - the synchronous host calls (vpb_infer_host, vpb_infer_frame_host) use staging slot 0, the same buffers as
 *     vpb_submit_host / vpb_submit_frame_host with slot 0; they are ordered after that slot's last submit and the next
 *     submit(0) is ordered after them;
 *   - if `stream` is being captured by the caller, the engine launches its kernels eagerly into that capture (no nested
 *     graph) and leaves its cross-stream ordering to the caller;
 *   - several engines may share one GPU.  Calls that launch the chained persistent GEMM kernels (batch >= the "chain_min_batch"
 *     option, 48 by default) are serialised per device across engines and streams (a device-side event wait plus a host mutex
 *     around the enqueue): such a kernel needs all of its thread-block clusters resident at once and must not share the SMs
 *     with a second one.  Another PROCESS running chained launches on the same GPU (MPS) is outside that gate: give chained
 *     engines the GPU to themselves or set option "chain" = 0;
 *   - engines on different devices may live in one process: each entry point makes its engine's device current for the
 *     duration of the call and restores the caller's; `stream` must belong to the engine's device.
 *
 * Each entry point names the reference interface it replaces (paths relative to the reference repo).
 */
#ifndef VITPOSE_B200_H
#define VITPOSE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VPB_OK 0
#define VPB_ERR_ARG 1      /* bad argument / unsupported configuration / missing weights */
#define VPB_ERR_CUDA 2     /* CUDA runtime or driver error (message in vpb_last_error) */
#define VPB_ERR_STATE 3    /* call order violated (e.g. forward before finalize) */

typedef struct vpb_engine vpb_engine;

/* Model hyper-parameters: easy_ViTPose/configs/ViTPose_common.py:65-195 (embed_dim, depth, num_heads;
 * mlp_ratio 4, qkv_bias, patch 16, img 256x192, 2 deconv layers of 256 filters, 1x1 final conv are fixed
 * on this path) and the per-dataset out_channels (e.g. configs/ViTPose_coco.py:16-18). */
typedef struct vpb_config {
  int32_t embed_dim;      /* 384 / 768 / 1024 / 1280 */
  int32_t depth;          /* 12 / 12 / 24 / 32 */
  int32_t num_heads;      /* 12 / 12 / 16 / 16 */
  int32_t num_keypoints;  /* K = keypoint_head.out_channels, 1..144 */
  int32_t max_batch;      /* workspace is sized for this many crops per call */
  int32_t device;         /* CUDA device ordinal */
} vpb_config;

/* Thread-local description of the last failure in this thread ("" if none). */
const char* vpb_last_error(void);

/* Replaces ViTPose(cfg) construction (easy_ViTPose/vit_models/model.py:10-18; VitInference.__init__,
 * easy_ViTPose/inference.py:156-157): allocates weights arena + workspace on cfg->device. */
int vpb_create(const vpb_config* cfg, vpb_engine** out);
void vpb_destroy(vpb_engine* e);

/* Replaces nn.Module.load_state_dict (easy_ViTPose/inference.py:162-166), one tensor at a time, under the
 * reference's own key names ("backbone.blocks.3.attn.qkv.weight", ...).  `data` is HOST
 * float32, C-contiguous, `numel` elements; "...num_batches_tracked" keys are accepted and ignored. */
int vpb_load_tensor(vpb_engine* e, const char* key, const float* data, int64_t numel);
/* Strict check (every key of the contract loaded exactly once) + one-time packing on the GPU: bf16
 * conversion, q-scale fold into attn.qkv, pos_embed+conv-bias fold, BatchNorm fold into the deconv phases. */
int vpb_finalize(vpb_engine* e);

/* Replaces ViTPose.forward (vit_models/model.py:23-24): d_crops f32 [batch,3,256,192] ->
 * d_heatmaps f32 [batch,K,64,48]. */
int vpb_forward(vpb_engine* e, const float* d_crops, int32_t batch, float* d_heatmaps, void* stream);
/* Replaces ViTPose.forward_features (vit_models/model.py:20-21): -> d_features f32 [batch,D,16,12]. */
int vpb_forward_features(vpb_engine* e, const float* d_crops, int32_t batch, float* d_features, void* stream);

/* Replaces keypoints_from_heatmaps(unbiased=True, use_udp=True) + the (y,x,score) packing of
 * VitInference.postprocess (vit_utils/top_down_eval.py:493-641 branch :586-589; easy_ViTPose/inference.py:187-205).
 * d_heatmaps f32 [n,k,64,48]; d_org_wh i32 [n,2] crop (width,height); d_kpts f32 [n,k,3]; d_idx i32 [n,k] flat
 * argmax or NULL.  wrap_batch selects the reference's "previous map" for max<=0 maps: 0 = one reference call per
 * crop (VitInference), 1 = one reference call on the whole [n,k,H,W] array.  Does not modify d_heatmaps. */
int vpb_decode(const float* d_heatmaps, int32_t n, int32_t k, const int32_t* d_org_wh, float* d_kpts, int32_t* d_idx,
               int32_t wrap_batch, void* stream);

/* Replaces TopdownHeatmapSimpleHead.forward (vit_models/head/topdown_heatmap_simple_head.py:188-193) on its own:
 * d_features f32 [batch,D,16,12] (what vpb_forward_features returns) -> d_heatmaps f32 [batch,K,64,48]. */
int vpb_head(vpb_engine* e, const float* d_features, int32_t batch, float* d_heatmaps, void* stream);
/* Replaces flip_back (vit_utils/post_processing/post_transforms.py:110-147, GaussianHeatmap) plus the optional one-pixel
 * shift of TopdownHeatmapSimpleHead.inference_model (:210-212): d_in / d_out f32 [n,k,64,48] (distinct buffers), d_perm i32 [k]
 * = the keypoint permutation the flip pairs induce (perm[left] = right, perm[right] = left, identity elsewhere). */
int vpb_flip_back(const float* d_in, int32_t n, int32_t k, const int32_t* d_perm, int32_t shift, float* d_out, void* stream);

/* The other modes of keypoints_from_heatmaps (vit_utils/top_down_eval.py:493-641), with the
 * general transform_preds (post_processing/post_transforms.py:150-194).  mode: 0 post_process=None (:598), 1 'default' (+-0.25 px,
 * :617-631), 2 'unbiased' (Gaussian modulation + log + _taylor, :600-607), 3 'megvii' (:573-574,:629-639) -- all use_udp=False --
 * and 4 = use_udp=True DARK (:576-579) with arbitrary centre / scale.  Exactly one of d_cs32 (f32 [n,4]) / d_cs64 (f64 [n,4]) holds
 * (centre_x, centre_y, scale_x, scale_y) per crop: float32 arrays keep numpy's arithmetic in float32, int64 / float64 arrays
 * promote it to float64.  Output layout as vpb_decode: d_kpts f32 [n,k,3] (y, x, score), d_idx i32 [n,k] or NULL.
 * vpb_decode_modes is the kernel = 11 form (every reference config: modulate_kernel=11).  vpb_decode_modes_ex adds
 *   kernel        the `kernel` argument (:499): odd, 1..35 (17 for sigma = 3; 1 only for modes 4-5); used by modes 2-5;
 *   mode 5        use_udp=True with target_type='CombinedTarget' (:580-593): d_heatmaps is f32 [n,3k,64,48], triples of
 *                 (response, offset x, offset y); 2*kernel+1 <= 35; valid_radius = (float)(valid_radius_factor * 64) (:586);
 *                 the (-1,-1) sentinel reads its offsets one row and one pixel before the keypoint's plane, wrapping to the
 *                 last plane of the call for the first keypoint, as numpy's flat index does.  (The reference's own index
 *                 arithmetic (:589) only broadcasts for n = 1; n > 1 here is that formula with the intended shape.) */
#define VPB_DECODE_NONE 0
#define VPB_DECODE_DEFAULT 1
#define VPB_DECODE_UNBIASED 2
#define VPB_DECODE_MEGVII 3
#define VPB_DECODE_DARK_UDP 4
#define VPB_DECODE_COMBINED 5
int vpb_decode_modes(const float* d_heatmaps, int32_t n, int32_t k, int32_t mode, const float* d_cs32, const double* d_cs64,
                     float* d_kpts, int32_t* d_idx, void* stream);
int vpb_decode_modes_ex(const float* d_heatmaps, int32_t n, int32_t k, int32_t mode, int32_t kernel, float valid_radius,
                        const float* d_cs32, const double* d_cs64, float* d_kpts, int32_t* d_idx, void* stream);

/* Replaces the model + postprocess part of VitInference._inference_torch (easy_ViTPose/inference.py:320-328)
 * for a whole batch of crops resident on the device.  d_heatmaps may be NULL. */
int vpb_infer(vpb_engine* e, const float* d_crops, const int32_t* d_org_wh, int32_t batch, float* d_kpts,
              int32_t* d_idx, float* d_heatmaps, void* stream);
/* Same with HOST buffers: H2D of crops/org_wh, the path, D2H of keypoints (+idx), then a stream sync --
 * the reference's `.to(device)` ... `.cpu().numpy()` bracket (inference.py:324,327).  Pinned host memory
 * (vpb_host_alloc) makes the copies asynchronous DMA. */
int vpb_infer_host(vpb_engine* e, const float* h_crops, const int32_t* h_org_wh, int32_t batch, float* h_kpts,
                   int32_t* h_idx, void* stream);
/* Pipelined form: vpb_submit_host(slot 0|1) enqueues H2D (engine copy stream), the path and D2H (engine compute stream)
 * and returns; vpb_wait_host(slot) blocks until that slot's keypoints are in h_kpts.  Keeping two slots in flight hides
 * the H2D of batch i+1 under the compute of batch i.  Host buffers must stay valid until the wait (pinned: real overlap). */
int vpb_submit_host(vpb_engine* e, const float* h_crops, const int32_t* h_org_wh, int32_t batch, float* h_kpts,
                    int32_t* h_idx, int32_t slot);
int vpb_wait_host(vpb_engine* e, int32_t slot);
void* vpb_host_alloc(int64_t bytes);   /* cudaHostAlloc; NULL on failure */
void vpb_host_free(void* p);

/* ---- frame-level entry points: the per-person loop of VitInference.inference
 * (easy_ViTPose/inference.py:258-272) as one batched call.
 *
 * vpb_preprocess replaces, for all n boxes of a frame at once: the +-pad_bbox box padding and clipping (:259-261), the
 * crop (:264), pad_image(crop, 3/4) (vit_utils/inference.py:41-70) and VitInference.pre_img (:314-318: cv2 uint8
 * INTER_LINEAR resize to 192x256, /255, (x-MEAN)/STD in float64, CHW float32).  Bit-exact with cv2 4.13.
 *   d_frame u8 [frame_h, frame_w, 3] RGB, row pitch pitch_bytes (0 = packed);  d_bboxes i32 [n,4] (x0,y0,x1,y1), already
 *   rounded as `res_pd[:, :4].round().astype(int)` (:253);  d_crops f32 [n,3,256,192];  d_org_wh i32 [n,2] padded canvas
 *   (w,h) = what pre_img returns as (org_w, org_h);  d_offs_yx i32 [n,2] = (y0 - top_pad, x0 - left_pad), the offset :270
 *   adds;  d_status i32 [1] or NULL: bit 0 is OR-ed in when a box is empty after clipping (the reference raises there;
 *   the kernel emits a black crop with org_wh = 0). */
int vpb_preprocess(const uint8_t* d_frame, int32_t frame_h, int32_t frame_w, int64_t pitch_bytes, const int32_t* d_bboxes,
                   int32_t n, int32_t pad_bbox, float* d_crops, int32_t* d_org_wh, int32_t* d_offs_yx, int32_t* d_status,
                   void* stream);
/* vpb_decode with the :270 offsets applied in the same kernel: keypoints come out in FRAME pixels. */
int vpb_decode_frame(const float* d_heatmaps, int32_t n, int32_t k, const int32_t* d_org_wh, const int32_t* d_offs_yx,
                     float* d_kpts, int32_t* d_idx, int32_t wrap_batch, void* stream);
/* frame + boxes on the device -> d_kpts f32 [n,K,3] (y, x, score) in frame pixels, d_idx i32 [n,K] or NULL; n <= max_batch.
 * pad_bbox is the reference's 10. */
int vpb_infer_frame(vpb_engine* e, const uint8_t* d_frame, int32_t frame_h, int32_t frame_w, const int32_t* d_bboxes,
                    int32_t n, float* d_kpts, int32_t* d_idx, void* stream);
/* Status of the device-side frame calls since the last query: bit 0 = some box was empty after padding and clipping
 * (the reference raises there: ZeroDivisionError in pad_image / cv2.resize, easy_ViTPose/inference.py:259-265; the
 * device path cannot raise without a sync and decodes such a box from a black crop).  Synchronises the device, clears
 * the word. */
int vpb_frame_status(vpb_engine* e, int32_t* h_status);
/* Same with HOST buffers (H2D of the packed uint8 frame + 16 B per box, D2H of the keypoints, stream sync); empty boxes
 * return VPB_ERR_ARG where the reference raises.  The pipelined form shares its slots and vpb_wait_host with vpb_submit_host. */
int vpb_infer_frame_host(vpb_engine* e, const uint8_t* h_frame, int32_t frame_h, int32_t frame_w, const int32_t* h_bboxes,
                         int32_t n, float* h_kpts, int32_t* h_idx, void* stream);
int vpb_submit_frame_host(vpb_engine* e, const uint8_t* h_frame, int32_t frame_h, int32_t frame_w, const int32_t* h_bboxes,
                          int32_t n, float* h_kpts, int32_t* h_idx, int32_t slot);

/* Introspection used by bench.py / tests. */
int vpb_kernel_launches(const vpb_engine* e, int32_t batch);          /* kernels one vpb_infer enqueues */
/* Options (all keep the results bit-identical unless noted): "stop_after", "profile", "pdl", "graph", "ln_fused",
 * "chain" (chained persistent launches, default 1), "chain_min_batch" (smallest batch that takes them, default 48),
 * "ln_in_gemm" (LayerNorm + its consumer GEMM as one launch on the unchained path, default 0), "gelu_erf" (fc1 epilogue with
 * erf instead of the fitted tanh form: rounding-level differences), "ln_ctl" (chained launches: counter polls / publishes of
 * the LayerNorm jobs on a control warp, default 1; VPB_LN_CTL), "ln_job_rows" (8 | 16 rows per LayerNorm job, default 16),
 * "resid_rmw" (residual epilogues as load + add + store instead of TMA reduce-add: measured slower, default 0; VPB_RESID_RMW). */
int vpb_set_option(vpb_engine* e, const char* name, int32_t value);
/* With option "profile"=1 every launch is bracketed by a CUDA-event pair on its stream; collect() synchronises,
 * sums elapsed ms and launch counts per kernel class (arrays of vpb_profile_classes() entries) and resets. */
int vpb_profile_classes(void);
const char* vpb_profile_class_name(int32_t cls);
int vpb_profile_collect(vpb_engine* e, float* ms_per_class, int32_t* launches_per_class);
int vpb_read_buffer(vpb_engine* e, const char* name, void* host_dst, int64_t bytes);  /* synchronous debug read */

/* Kernel-level entry points (unit tests / profiling).  All pointers are device pointers.
 * vpb_gemm: out = epilogue(A[M,K] bf16 * W[N,K]^T bf16 + bias);  epilogue ids as in csrc/gemm.cuh (0 bias->bf16, 1 bias+GELU->bf16,
 * 2 implicit-GEMM deconv: aux = H, W, tile rows, (tile cols << 16) | Cin, 4 bias->f32 NCHW: aux0 = channels, aux1 = pixels,
 * 5 f32 out += ...).  d_resid / resid_mod are reserved (pass NULL / 0). */
int vpb_gemm(const void* d_a, const void* d_w, const float* d_bias, void* d_out, int32_t m, int32_t n, int32_t k,
             int32_t epilogue, const float* d_resid, int32_t resid_mod, int32_t aux0, int32_t aux1, int32_t aux2,
             int32_t aux3, void* stream);
/* Debug: limit the GEMM smem ring depth and/or collect per-CTA cycle counters (int64 [grid*8]) for following GEMM launches. */
int vpb_debug_gemm(int32_t stages_limit, void* d_counters);
int vpb_attention(const void* d_qkv, int32_t batch, int32_t heads, int32_t head_dim, void* d_out, void* stream);
/* Debug / measurement switch (process-wide) for the attention variants.  flags < 0: the defaults (packed half tiles for
 * head_dim 32 / 64 with every 4th softmax exponential evaluated by a polynomial on the FMA pipe; VPB_ATT_PACK / VPB_ATT_POLY = 0 / 1
 * in the environment change them); otherwise bit 0 = polynomial exponentials, bit 1 = packed half tiles, bits 8.. = cap on the
 * number of CTAs (0 = one per SM; tests use it to move the boundaries of the per-CTA step ranges). */
int vpb_debug_attention(int32_t flags);
int vpb_layernorm(const float* d_x, const float* d_gamma, const float* d_beta, void* d_y, int32_t rows, int32_t dim,
                  float eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VITPOSE_B200_H */
