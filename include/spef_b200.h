/*
 * spef_b200.h -- C ABI of libspef_b200.so: the B200 (sm_100a) implementation of SPEF's batched
 * pose-inference hot path (Mobile-URSONet forward -> softmax -> soft-classification decode ->
 * pose error / ESA score -> temporal pdf filter).
 *
 * The reference (possoj/Spacecraft-Pose-Estimation-Framework) is 100 % Python and has no FFI; its
 * "back-end" boundary is a duck-typed object with predict(images) (src/spe/spe_torch.py:41-76,
 * called from src/tools/evaluation.py:71 and src/temporal/inference.py:132).  Each entry point
 * below names the reference function(s) it replaces; INTEGRATION.md shows the ctypes binding a
 * reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 (SPEF_OK) or a spef_status code; nothing throws across the ABI;
 *     spef_last_error(ctx) returns a human-readable message for the last failure on that ctx.
 *   - *_dev pointers are caller-owned device memory on the ctx's device (e.g. torch
 *     Tensor.data_ptr()); *_host pointers are caller-owned host memory (pinned memory makes the
 *     copies asynchronous).  The library owns only weights, tables, workspaces and stream state
 *     inside the ctx.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All work is
 *     stream-ordered; *_host variants synchronise the stream before returning.
 *   - one ctx per device, not thread-safe, no global state.
 *   - tensors: images [B,3,H,W] float32 NCHW in [0,1] (the reference's contract,
 *     src/data/datasets/speed.py:66-69); quaternions scalar-first [B,4] float32; positions [B,3]
 *     float32; logits / pdfs [B,n] float32 row-major.
 */
#ifndef SPEF_B200_H
#define SPEF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPEF_ABI_VERSION 1

typedef enum spef_status {
  SPEF_OK = 0,
  SPEF_ERR_INVALID = 1,     /* bad argument / shape / key */
  SPEF_ERR_CUDA = 2,        /* CUDA runtime or driver error (message has the CUDA string) */
  SPEF_ERR_STATE = 3,       /* call order violated (weights not finalised, histogram missing, ...) */
  SPEF_ERR_UNSUPPORTED = 4, /* configuration the kernels do not implement */
  SPEF_ERR_NUMERIC = 5      /* a reference ValueError condition (NaN in decode, zero pdf sum) */
} spef_status;

enum { SPEF_FP32 = 0, SPEF_BF16 = 1 };

/* flag bits written per image by the decode / score kernels */
enum {
  SPEF_FLAG_ORI_NAN = 1u,      /* NaN in A: classification_utils.py:134-135 raises ValueError */
  SPEF_FLAG_POS_ZERO_SUM = 2u, /* classification_utils.py:253-254 */
  SPEF_FLAG_POS_NAN = 4u,      /* classification_utils.py:262-263 */
  SPEF_FLAG_DOT_GT_1_01 = 8u   /* spe_utils.py:137-138 (dead code in the reference: diagnostic only) */
};

typedef struct spef_ctx spef_ctx;

typedef struct spef_config {
  int32_t struct_size; /* = sizeof(spef_config) */
  int32_t device;      /* CUDA device ordinal */
  int32_t img_h;       /* 240  (src/config/train/config.py:27) */
  int32_t img_w;       /* 384 */
  int32_t n_ori;       /* orientation head outputs (bins), e.g. 1728 */
  int32_t n_pos;       /* position head outputs: 3 (regression) or bins (classification) */
  int32_t pos_classification; /* 0: 'regression', 1: 'classification' */
  int32_t precision;   /* SPEF_FP32 | SPEF_BF16: storage type of activations / GEMM operands */
  int32_t max_batch;   /* largest B passed to forward/predict; sizes the workspaces */
  int32_t pw_impl;     /* 0: default (tcgen05 GEMM for BF16, SIMT FP32 for FP32); 1: force SIMT (debug cross-check) */
} spef_config;

/* ---- lifecycle ------------------------------------------------------------------------------ */
int spef_abi_version(void);
/* replaces: SPETorch.__init__ (src/spe/spe_torch.py:24-39) + import_model's module construction
 * (src/modeling/model.py:185-257) */
int spef_create(spef_ctx** out, const spef_config* cfg);
void spef_destroy(spef_ctx* ctx);
const char* spef_last_error(const spef_ctx* ctx); /* ctx may be NULL: message of the last failed spef_create */

/* ---- weights: replaces model.load_state_dict (src/modeling/model.py:261-266) --------------------
 * `key` is a reference state_dict key (features.features.{i}..., head.ori.1.*,
 * head.pos.0.*); data is the raw float32 tensor in PyTorch layout.  num_batches_tracked is ignored.
 * spef_finalize_weights folds BatchNorm (eps 1e-5, src/modeling/common/pytorch_layers.py:55-56),
 * repacks to kernel layouts, casts and uploads; it fails if any of the 315 float tensors is missing. */
int spef_load_tensor(spef_ctx* ctx, const char* key, const float* data_host, const int64_t* shape, int32_t ndim);
int spef_finalize_weights(spef_ctx* ctx);

/* ---- histograms: replaces OrientationSoftClassification.histogram / .b and
 * PositionSoftClassification.histogram (src/spe/classification_utils.py:39-83,168-176,201-216) -- */
int spef_set_ori_histogram(spef_ctx* ctx, const double* quat_bins_host /*[n,4]*/, int32_t n);
int spef_set_pos_histogram(spef_ctx* ctx, const double* pos_bins_host /*[n,3]*/, int32_t n);

/* ---- image element type (input side of the path, src/data/utils.py:212-249): SPEF_IMG_F32 (default) is the reference's
 * tensor contract, float32 in [0,1]; SPEF_IMG_U8 takes the 8-bit pixels that ToTensor divides by 255 -- the stem computes
 * float(u8) / 255.0f itself, bit-identical to feeding the float tensor, with 4x less host-to-device and HBM traffic.
 * Applies to every images_* argument of this context (BF16 engine, tcgen05 path only). */
enum { SPEF_IMG_F32 = 0, SPEF_IMG_U8 = 1 };
int spef_set_image_dtype(spef_ctx* ctx, int32_t image_dtype);

/* ---- camera frames (the step before the tensor contract) -----------------------------
 * replaces, per frame, SPEDataset.__getitem__'s  self.transform(Image.open(path).convert("RGB"))  (src/data/utils.py:215-226)
 * with transform = Compose([Resize(img_size), ToTensor()]) (src/data/datasets/speed.py:59-62): Pillow's antialiased BILINEAR
 * resample (separable triangle filter, 22-bit fixed point on 8-bit pixels, horizontal pass first, each pass rounded and
 * clipped) followed by float32(u8) / 255.  frames_dev: decoded 8-bit frames [B, src_h, src_w, channels], channels = 1
 * (SPEED's greyscale JPEGs; replicated to three planes like convert("RGB")) or 3 (RGB, HWC as np.array(image)).
 * images_out_dev: [B, 3, img_h, img_w] of out_dtype: SPEF_IMG_U8 (feed it to a context set to SPEF_IMG_U8) or SPEF_IMG_F32
 * (the reference tensor).  Bit-exact against torchvision + Pillow for any source size (up- or down-scaling).  JPEG
 * decoding stays with the caller. */
int spef_resize_frames(spef_ctx* ctx, const uint8_t* frames_dev, int32_t batch, int32_t src_h, int32_t src_w, int32_t channels,
                       void* images_out_dev, int32_t out_dtype, void* stream);

/* ---- network: replaces ModelWrapper.forward (src/modeling/common/pytorch_layers.py:29-32) ------ */
int spef_forward(spef_ctx* ctx, const float* images_dev, int32_t batch,
                 float* ori_out_dev /*[B,n_ori] logits*/, float* pos_out_dev /*[B,n_pos]*/, void* stream);

/* teacher-forced single layer (parity harness; mirrors hooking one ConvBnAct of the reference):
 * layer 0 = stem (input NCHW f32), 1..51 = the following conv layers in execution order
 * (NHWC, ctx precision), 52 = global mean over H,W, 53 = head GEMM (in [B,1280], out f32 [B,n_ori+n_pos pad]).
 * residual_dev may be NULL. */
int spef_num_layers(const spef_ctx* ctx);
int spef_layer_info(const spef_ctx* ctx, int32_t layer, int32_t* kind /*0 stem,1 pw,2 dw,3 pool,4 head*/,
                    int32_t* cin, int32_t* cout, int32_t* hin, int32_t* win, int32_t* hout, int32_t* wout,
                    int32_t* stride, int32_t* relu, int32_t* has_residual);
int spef_layer_forward(spef_ctx* ctx, int32_t layer, const void* in_dev, const void* residual_dev,
                       void* out_dev, int32_t batch, void* stream);

/* InvertedResidual blocks (src/modeling/common/pytorch_layers.py:65-98; 17 of them, mobilenet_v2.py:240-262).
 * On the BF16 tcgen05 path a block runs as ONE kernel (expand 1x1 -> depthwise 3x3 -> project 1x1 [+ x]) whose
 * hidden tensor stays in shared memory / TMEM.  spef_block_info reports which layers [first_layer, first_layer +
 * n_layers) a block covers and whether it is fused in the current configuration (*fused = 0 per-layer kernels, 1 staged
 * fused kernel, 2 channel-lane fused kernel; plus its tile plan);
 * spef_block_forward is the teacher-forced single block (in [B,H,W,Cin] -> out [B,Ho,Wo,Cout], NHWC bf16) and fails with
 * SPEF_ERR_UNSUPPORTED for a block that is not fused.  spef_set_fusion(0) makes spef_forward run the per-layer kernels
 * (the parity cross-check of the fused path); default on, or SPEF_FUSE=0 in the environment. */
int spef_num_blocks(const spef_ctx* ctx);
int spef_block_info(const spef_ctx* ctx, int32_t block, int32_t* first_layer, int32_t* n_layers, int32_t* fused,
                    int32_t* tile_h, int32_t* tile_w, int32_t* groups, int32_t* w_stages, int32_t* resident);
int spef_set_fusion(spef_ctx* ctx, int32_t on);
int spef_block_forward(spef_ctx* ctx, int32_t block, const void* in_dev, void* out_dev, int32_t batch, void* stream);

/* Stem fused into the first block (features[0] ConvBnAct(3 -> 32, k3, s2) + features[1], src/modeling/backbone/mobilenet_v2.py:252-262;
 * BF16 tcgen05 path, both image dtypes): ONE kernel reads the image and writes the first block's output [B,H/2,W/2,16]; the stem's output
 * (the largest tensor of the network) never reaches HBM and stays FP32 between the stem GEMM and the depthwise taps.  Default on
 * (SPEF_STEM_FUSE=0 in the environment or spef_set_stem_fusion(ctx, 0): separate stem launch, the parity cross-check);
 * spef_stem_fusion_active reports whether spef_forward takes that route in the current configuration (it needs the channel-lane
 * kernel of block 0 and spef_set_fusion on); spef_stem_block_forward is the teacher-forced launch (images_dev as for spef_forward,
 * out_dev NHWC bf16) and fails with SPEF_ERR_UNSUPPORTED when the route is not active. */
int spef_set_stem_fusion(spef_ctx* ctx, int32_t on);
int spef_stem_fusion_active(const spef_ctx* ctx);
/* Last 1x1 conv of the backbone (mobilenet_v2.py:264) + the global average pool in front of the heads (`x.mean([2, 3])`) as ONE
 * kernel on the BF16 tcgen05 path (csrc/conv_pool.cuh; bit-identical to the two launches; SPEF_POOL_FUSE=0 or spef_set_fusion(ctx, 0):
 * two launches).  Reports whether spef_forward takes that route in the current configuration. */
int spef_pool_fusion_active(const spef_ctx* ctx);
int spef_stem_block_forward(spef_ctx* ctx, const void* images_dev, void* out_dev, int32_t batch, void* stream);

/* ---- encode (label side) and error statistics (8f #3) ----------------------------
 * spef_encode_ori replaces OrientationSoftClassification.encode for a batch of labels
 * (src/spe/classification_utils.py:85-111): k_b = exp(-((2 acos(min(1, |q . h_b|)) / pi)^2 / (2 variance))), bins with
 * masked_dev[b] != 0 (the reference's redundant_flags when delete_unused_bins is False; may be NULL) set to 0,
 * pdf = k / sum k, float64 arithmetic, float32 result [B,n].  spef_encode_pos replaces PositionSoftClassification.encode
 * (:218-240).  variance = (smooth_factor / n_bins_per_dim)^2 / 12.  flags_dev (nullable, zero it first): SPEF_FLAG_ENC_NAN
 * where the reference raises ValueError('NaN found in encoded ...').
 * spef_error_stats replaces np.mean / np.std / np.median / mad() over the per-image error lists of evaluation()
 * (src/tools/evaluation.py:16-32, 95-99): x_dev[i * stride], i < n (e.g. a column of per_image) ->
 * out_host[4] = {mean, population std, median, median absolute deviation}; synchronises the stream. */
enum { SPEF_FLAG_ENC_NAN = 16u };
int spef_encode_ori(spef_ctx* ctx, const double* quat_dev /*[B,4] float64 labels*/, int32_t batch, int32_t n, double variance,
                    const uint8_t* masked_dev /*[n] or NULL*/, float* pdf_out_dev /*[B,n]*/, uint32_t* flags_dev, void* stream);
int spef_encode_pos(spef_ctx* ctx, const double* pos_dev /*[B,3] float64 labels*/, int32_t batch, int32_t n, double variance,
                    float* pdf_out_dev /*[B,n]*/, uint32_t* flags_dev, void* stream);
int spef_error_stats(spef_ctx* ctx, const float* x_dev, int32_t stride, int32_t n, double* out_host /*[4]*/, void* stream);

/* ---- post-processing -------------------------------------------------------------------------
 * spef_decode_ori replaces SPEUtils.last_activ (softmax, src/spe/spe_utils.py:75-76) when
 * is_logits != 0, and OrientationSoftClassification.decode_batch
 * (src/spe/classification_utils.py:113-166).  n must equal the histogram size.
 * soft_out_dev [B,n], hinv_out_dev [B,16], argmax_out_dev [B] may be NULL.  flags_dev [B] is OR-ed. */
int spef_decode_ori(spef_ctx* ctx, const float* in_dev, int32_t batch, int32_t n, int32_t is_logits,
                    float* soft_out_dev, float* quat_out_dev, float* hinv_out_dev, int32_t* argmax_out_dev,
                    uint32_t* flags_dev, void* stream);
/* replaces softmax (spe_utils.py:77-79) + PositionSoftClassification.decode_batch
 * (classification_utils.py:242-285) */
int spef_decode_pos(spef_ctx* ctx, const float* in_dev, int32_t batch, int32_t n, int32_t is_logits,
                    float* soft_out_dev, float* pos_out_dev, uint32_t* flags_dev, void* stream);
/* replaces SPEUtils.get_score (src/spe/spe_utils.py:104-159) and the per-image error lists of
 * evaluation() (src/tools/evaluation.py:82-85).  sums_dev[8] (double) is ACCUMULATED into:
 *   [0] sum e_q (rad)  [1] sum e_t/|t|  [2] sum e_t (m)  [3] image count  [4] #images with |q.q^|>1.01
 *   [5] #images with NaN error  [6] #images whose orientation decode raised its NaN guard  [7] #images whose position decode
 *   raised a guard (zero pdf sum / NaN) -- [6],[7] are only counted by the fused spef_eval_batch* route, which knows the decode
 *   flags of the batch; the reference raises ValueError from decode() for these (classification_utils.py:134, 253, 262), and so
 *   does evaluation() after spef_eval_read.  per_image_dev [B,2] = (e_q in degrees, e_t) may be NULL. */
int spef_score(spef_ctx* ctx, const float* quat_pred_dev, const float* pos_pred_dev,
               const float* quat_true_dev, const float* pos_true_dev, int32_t batch,
               double* sums_dev, float* per_image_dev, void* stream);

/* ---- fused predict: replaces SPETorch.predict (src/spe/spe_torch.py:41-76) ---------------------
 * forward + softmax + decode on device.  ori_soft_out / pos_soft_out may be NULL.  When the ctx
 * has pos_classification = 0, pos_out is the regressed position and pos_soft_out must be NULL. */
int spef_predict(spef_ctx* ctx, const float* images_dev, int32_t batch,
                 float* ori_soft_out_dev, float* quat_out_dev, float* pos_soft_out_dev, float* pos_out_dev,
                 int32_t* argmax_out_dev, uint32_t* flags_dev, void* stream);
/* same with HOST buffers: H2D of the images, predict, D2H of the results, stream synchronised. */
int spef_predict_host(spef_ctx* ctx, const float* images_host, int32_t batch,
                      float* ori_soft_out_host, float* quat_out_host, float* pos_soft_out_host, float* pos_out_host,
                      int32_t* argmax_out_host, uint32_t* flags_out_host, void* stream);
/* predict + score in one call (the body of evaluation()'s batch loop, src/tools/evaluation.py:69-85):
 * targets come from the host, sums are accumulated in the ctx (spef_eval_reset / spef_eval_read). */
int spef_eval_reset(spef_ctx* ctx, void* stream);
int spef_eval_batch_host(spef_ctx* ctx, const float* images_host, const float* quat_true_host,
                         const float* pos_true_host, int32_t batch, float* per_image_out_host /*[B,2] or NULL*/,
                         void* stream);
/* pipelined variant of spef_eval_batch_host for loaders that keep several batches in flight: the H2D copy of this
 * batch runs on an internal copy stream into one of two staging buffers while the previous batch is still computing on
 * `stream`; the call returns without synchronising.  per_image_out_host (pinned, or NULL) is filled asynchronously and is
 * valid after spef_eval_wait.  images / targets must stay valid until spef_eval_wait (or two further submits). */
int spef_eval_submit_host(spef_ctx* ctx, const float* images_host, const float* quat_true_host,
                          const float* pos_true_host, int32_t batch, float* per_image_out_host, void* stream);
int spef_eval_wait(spef_ctx* ctx, void* stream);
/* Packed upload for spef_eval_submit_host (float images, BF16 engine; replaces the `.to(device)` of the float batch in
 * SPETorch.predict, src/spe/spe_torch.py:57-61): the library's host threads round the pixels to BF16 -- the stem's own first step,
 * same round-to-nearest-even -- chunk by chunk into pinned staging while the DMA engine moves the previous chunk, and a widening
 * kernel restores the float tensor on the device: half the bytes on the bus, bit-identical results.  Default on for the BF16 engine
 * (SPEF_HOST_PACK=0 or spef_set_host_pack(ctx, 0): plain copy; SPEF_PACK_THREADS: host threads, default min(16, cores / LOCAL_WORLD_SIZE)).
 * A submit splits its batch: the tail crosses as float (the DMA engine starts on it at once), the head is packed; the split
 * balances the host threads against the bus from rates the context measures while it runs (float bytes / s converted, bytes / s
 * of the plain slice's copy); when the conversion rate is below half the copy rate (many ranks sharing a small host: the
 * copies are then limited by the host's memory system, which the conversion would load further) the batch goes as it is and only a
 * small probe slice is packed now and then.  spef_host_pack_info: whether packing is enabled, with how many host threads, and
 * stats[3] = {packed fraction of the recent submits, measured conversion rate, measured copy rate} (bytes / s). */
int spef_set_host_pack(spef_ctx* ctx, int32_t on);
/* the host half of the packed upload on its own (no GPU involved): dst[i] = bf16 bits of src[i], round to nearest even */
int spef_pack_bf16_host(const float* src_host, uint16_t* dst_host, int64_t n);
int spef_host_pack_info(const spef_ctx* ctx, int32_t* active, int32_t* threads, double* stats /*[3] or NULL*/);
int spef_eval_batch(spef_ctx* ctx, const float* images_dev, const float* quat_true_dev,
                    const float* pos_true_dev, int32_t batch, float* per_image_out_dev, void* stream);
int spef_eval_read(spef_ctx* ctx, double* sums_host /*[8]*/, void* stream);
double* spef_eval_sums_dev(spef_ctx* ctx); /* device pointer of the 8 accumulators (for an NCCL all-reduce) */

/* host-buffer variants of the post-processing entry points (what SPEUtils.decode / get_score bind to) */
int spef_decode_ori_host(spef_ctx* ctx, const float* in_host, int32_t batch, int32_t n, int32_t is_logits,
                         float* soft_out_host, float* quat_out_host, float* hinv_out_host,
                         int32_t* argmax_out_host, uint32_t* flags_out_host, void* stream);
int spef_decode_pos_host(spef_ctx* ctx, const float* in_host, int32_t batch, int32_t n, int32_t is_logits,
                         float* soft_out_host, float* pos_out_host, uint32_t* flags_out_host, void* stream);
int spef_score_host(spef_ctx* ctx, const float* quat_pred_host, const float* pos_pred_host,
                    const float* quat_true_host, const float* pos_true_host, int32_t batch,
                    double* sums_out_host /*[8], overwritten*/, float* per_image_out_host, void* stream);

/* ---- temporal: replaces Inference.predict(..., 'Adaptative') / Inference.reset
 * (src/temporal/inference.py:92-99, 131-180) and TemporalPDF.update_pdf
 * (src/temporal/pdf_compare.py:94-133) for n_streams independent videos (one frame each per call).
 * Requires pos_classification = 1.  Outputs are [n_streams, ...]; any *_out pointer may be NULL. */
typedef struct spef_temporal_out {
  float* still_ori_soft;  /* [S,n_ori] */
  float* still_pos_soft;  /* [S,n_pos] */
  float* still_quat;      /* [S,4]  sign-continuous */
  float* still_pos;       /* [S,3] */
  float* video_ori_soft;  /* [S,n_ori] filtered pdf */
  float* video_pos_soft;  /* [S,n_pos] */
  float* video_quat;      /* [S,4] */
  float* video_pos;       /* [S,3] */
  float* ori_distance;    /* [S] */
  float* pos_distance;    /* [S] */
  uint32_t* flags;        /* [S] */
} spef_temporal_out;
int spef_temporal_reset(spef_ctx* ctx, int32_t n_streams, void* stream);
/* from logits already on the device (filter + decode only) */
int spef_temporal_step_logits(spef_ctx* ctx, const float* ori_logits_dev, const float* pos_logits_dev,
                              int32_t n_streams, const spef_temporal_out* out_dev, void* stream);
/* from images on the device: forward + the above.  apply_filter = 0 reproduces Inference.predict(image, None):
 * still pose + sign continuity only, the filter state is left untouched and video_* outputs are not written. */
int spef_temporal_step(spef_ctx* ctx, const float* images_dev, int32_t n_streams, int32_t apply_filter,
                       const spef_temporal_out* out_dev, void* stream);
/* TemporalPDF.update_pdf (src/temporal/pdf_compare.py:94-133, 'l2' metric) on caller-owned state:
 * state_dev [S,n] previous filtered pdf, has_state_dev [S] (0 = first frame; set to 1 by the call). */
int spef_pdf_filter(spef_ctx* ctx, const float* cur_dev /*[S,n]*/, int32_t n_streams, int32_t n, float* state_dev,
                    int32_t* has_state_dev, float n_coef, float alpha, float* out_dev /*[S,n]*/,
                    float* distance_dev /*[S]*/, void* stream);

/* ---- introspection (bench / roofline bookkeeping) --------------------------------------------- */
/* number of kernel launches this ctx has issued since creation */
int64_t spef_launch_count(const spef_ctx* ctx);
/* algorithmic bytes / flops of one forward at batch B under the ctx precision */
int spef_forward_cost(const spef_ctx* ctx, int32_t batch, double* bytes_out, double* flops_out);
/* per-layer device time of the most recent spef_forward_timed call, milliseconds [spef_num_layers] */
int spef_forward_timed(spef_ctx* ctx, const float* images_dev, int32_t batch, float* ori_out_dev,
                       float* pos_out_dev, float* layer_ms_host, void* stream);

/* debug: the device-side 4x4 Jacobi eigen-solver compiled for the host (no GPU needed), so that the CPU
 * test-suite can pin it against LAPACK.  a_in row-major symmetric; evecs row-major with eigenvectors as columns. */
int spef_debug_jacobi4_host(const double* a_in, double* evals, double* evecs);
/* the tap tables spef_resize_frames builds on the host for one axis (first input sample, tap count, 22-bit fixed-point coefficients
 * [out_size][*ksize_out]); coef_capacity = number of int32 the caller allocated for coef_out */
int spef_debug_resize_taps_host(int32_t in_size, int32_t out_size, int32_t* first_out, int32_t* count_out, int32_t* coef_out,
                                int32_t coef_capacity, int32_t* ksize_out);
/* the eigen-solve of the large-batch decode kernel on the host: sums = {S, a00 a01 a02 a03 a11 a12 a13 a22 a23 a33} */
int spef_debug_decode_solve_host(const double* sums, int32_t is_logits, float* quat, float* hinv);

#ifdef __cplusplus
}
#endif
#endif /* SPEF_B200_H */
