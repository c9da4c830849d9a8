/*
 * lcm_unet.h — C ABI of the B200 (sm_100a) LCM denoising hot path.
 *
 * The reference (zamazincode/cv-diffusion-model) is pure Python and has no FFI of its own; the
 * drop-in boundary is its Python class surface.  Every entry point below replaces the
 * *execution* of one reference function and is what a binding for that function would call.
 * Reference citations are relative to /root/reference.
 *
 * Conventions
 *   - every function returns 0 on success, a negative lcm_status otherwise; the message of the
 *     last failure on the calling thread is returned by lcm_last_error().
 *   - all pointers named *dev* / workspace are DEVICE pointers owned by the caller (PyTorch in the
 *     Python binding); the library owns only the opaque plan and its packed-weight copies.
 *   - calls enqueue work on `stream` (a cudaStream_t passed as void*) and never synchronise.
 *   - a plan is bound to (device, config, batch, height, width, precision); it is not thread-safe.
 *   - images/latents cross the boundary as fp32 NCHW (the reference's tensors); activations inside
 *     the plan are NHWC in the plan's precision.
 *   - there is no CPU path: every call fails with LCM_ERR_CUDA when no sm_100 device is present.
 */
#ifndef LCM_UNET_H_
#define LCM_UNET_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LCM_MAX_LEVELS 8

typedef enum {
  LCM_OK = 0,
  LCM_ERR_INVALID = -1,     /* bad argument / unsupported configuration (Python: ValueError)  */
  LCM_ERR_CUDA = -2,        /* CUDA runtime failure (Python: RuntimeError)                    */
  LCM_ERR_UNKNOWN_WEIGHT = -3,
  LCM_ERR_MISSING_WEIGHT = -4,
  LCM_ERR_WORKSPACE = -5
} lcm_status;

typedef enum { LCM_PREC_FP32 = 0, LCM_PREC_BF16 = 1 } lcm_precision;
/* Activation element type accepted by the single-kernel entry points only: the bf16 tensor-core plan stores the
 * two hidden tensors of every inverted-residual block (expand output, depthwise output) as fp16. */
#define LCM_ACT_F16 2

/* plan flags */
#define LCM_FLAG_SIMT_GEMM 1u     /* bf16 plans: use the CUDA-core GEMM/conv kernels instead of tcgen05 (cross-check) */
#define LCM_FLAG_TAPS 2u          /* keep every intermediate alive (no buffer reuse) so lcm_plan_read_tap works */
#define LCM_FLAG_DRY 8u           /* describe only: op lists, workspace and gradient layout, no device needed; cannot run */
#define LCM_FLAG_TRAIN 4u         /* training plan: forward keeps every activation, the plan also holds the backward op list,
                                     gradient buffers and a flat fp32 weight-gradient buffer inside the workspace */

/* Mirrors EfficientUNetConfig (src/models/efficient_unet.py:24-57). */
typedef struct {
  int32_t in_channels;
  int32_t out_channels;
  int32_t base_channels;
  int32_t num_levels;
  int32_t channel_multipliers[LCM_MAX_LEVELS];
  int32_t num_attention_resolutions;
  int32_t attention_resolutions[LCM_MAX_LEVELS];
  int32_t num_attention_heads;
  int32_t num_res_blocks;
  int32_t expansion_ratio;
  float se_ratio;
  int32_t time_embed_dim;
  int32_t image_size;       /* config.image_size: decides where attention modules exist (:426,447) */
  int32_t groupnorm_gcd;    /* 0: min(32,C) groups like the reference; 1: gcd(32,C) (tiny/base patch) */
  int32_t standard_attention; /* 0: LinearAttention (every preset); 1: StandardAttention, use_linear_attention=False
                                 (efficient_unet.py:311-357, 448-454, 473-474); inference plans only */
} lcm_unet_config;

typedef struct lcm_plan lcm_plan;

const char* lcm_last_error(void);
int lcm_version(void);

/* ---- plan life cycle ------------------------------------------------------------------------
 * Replaces EfficientUNet.__init__'s role of fixing the op sequence (efficient_unet.py:403-530). */
int lcm_plan_create(const lcm_unet_config* cfg, int batch, int height, int width, int precision,
                    uint32_t flags, int device, lcm_plan** out);
void lcm_plan_destroy(lcm_plan* plan);
size_t lcm_plan_workspace_bytes(const lcm_plan* plan);

/* Weights: state_dict entries without the "unet." prefix, fp32, device memory.
 * Replaces model.load_state_dict() for the native path; packs into the plan's own layouts. */
int lcm_plan_num_weights(const lcm_plan* plan);
int lcm_plan_weight_info(const lcm_plan* plan, int index, const char** name, int64_t* numel);
int lcm_plan_set_weight(lcm_plan* plan, const char* name, const float* dev_values, int64_t numel, void* stream);

/* ---- EfficientUNet.forward (efficient_unet.py:532-606) ----------------------------------------
 * x = cat([xa, xb], dim=1): xa has ca channels, xb has cb channels (cb may be 0), ca+cb == in_channels;
 * each is fp32 NCHW with the given batch stride in elements.  t_dev: int64[batch].  eps_dev: fp32 NCHW. */
int lcm_unet_forward(lcm_plan* plan, const float* xa_dev, int ca, int64_t xa_batch_stride,
                     const float* xb_dev, int cb, int64_t xb_batch_stride, const int64_t* t_dev,
                     float* eps_dev, void* workspace, void* stream);

/* ---- LowLightDiffusion.enhance loop (low_light_diffusion.py:204-240) with LCMScheduler.step
 * (lcm_scheduler.py:204-242) fused into the last kernel of every forward.
 *   cond_dev      fp32 [B,3,H,W]   low-light image in [-1,1]
 *   latents_dev   fp32 [B,3,H,W]   in: initial noise; out: pre-clamp final latents (updated in place)
 *   noises_dev    fp32 [steps-1,B,3,H,W] per-step noise (the reference draws randn_like, :237)
 *   timesteps     host int64[steps]  (e.g. 739,499,259,19)
 *   coef          host float[steps*4]: sqrt(1-abar_t), sqrt(abar_t), sqrt(abar_prev), sqrt(1-abar_prev)
 *   out_dev       fp32 [B,3,H,W]   clamp(latents,-1,1)
 *   trace_dev     optional fp32 [steps,B,3,H,W]: post-step latents (return_intermediate), may be NULL */
int lcm_enhance(lcm_plan* plan, const float* cond_dev, float* latents_dev, const float* noises_dev,
                int steps, const int64_t* timesteps, const float* coef, float* out_dev, float* trace_dev,
                void* workspace, void* stream);

/* ---- condition_mode="add" (low_light_diffusion.py:108-113,158-160,223-225): the UNet has in_channels == out_channels, its
 * input is latents + condition_encoder(low_light).  lcm_condition_encode runs the encoder (Conv3x3(3->hidden) -> SiLU ->
 * Conv3x3(hidden->3), weights in the reference's layouts, fp32) once; lcm_enhance_add is lcm_enhance with the per-step add. */
size_t lcm_condition_encode_scratch_bytes(int batch, int height, int width, int hidden);
int lcm_condition_encode(const float* low_dev, const float* w1_dev, const float* b1_dev, const float* w2_dev, const float* b2_dev,
                         float* out_dev, int batch, int height, int width, int hidden, void* scratch_dev, void* stream);
int lcm_enhance_add(lcm_plan* plan, const float* cond_feat_dev, float* latents_dev, const float* noises_dev, int steps,
                    const int64_t* timesteps, const float* coef, float* out_dev, float* trace_dev, void* workspace, void* stream);

/* ---- LCMScheduler.step (lcm_scheduler.py:204-242), stand-alone.  prediction: 0 epsilon, 1 v_prediction.
 * noise_dev == NULL means last step (prev = x0). */
int lcm_scheduler_step(const float* model_out_dev, const float* sample_dev, const float* noise_dev,
                       float* prev_dev, float* x0_dev, int64_t numel, int prediction, float sqrt_beta_t,
                       float sqrt_alpha_t, float sqrt_alpha_prev, float sqrt_beta_prev, void* stream);

/* ---- LCMScheduler.add_noise / get_velocity (lcm_scheduler.py:255-305).
 * out = sa[t]*a + sb[t]*b (add_noise: a=x0,b=noise) or sa[t]*b - sb[t]*a (velocity: a=sample,b=noise). */
int lcm_scheduler_mix(const float* a_dev, const float* b_dev, const int64_t* t_dev, const float* abar_dev,
                      float* out_dev, int batch, int64_t per_sample, int velocity, void* stream);

/* ---- consistency distillation (LowLightLCMDistillation.consistency_distillation_loss, low_light_diffusion.py:325-408):
 * the element-wise steps around its three UNet forwards; per-sample timesteps index abar_dev.
 * lcm_ddim_step (:372-381): x0 = (x_t - sqrt(1-a_t) eps) / sqrt(a_t); x_next = sqrt(a_next) x0 + sqrt(1-a_next) eps.
 * lcm_consistency_loss (:399-406): *loss_dev = mean huber(student_x0, target_x0), d_eps_student = d loss / d eps_student. */
int lcm_ddim_step(const float* x_t_dev, const float* eps_dev, const int64_t* t_dev, const int64_t* t_next_dev, const float* abar_dev,
                  float* x_next_dev, int batch, int64_t per_sample, void* stream);
int lcm_consistency_loss(const float* x_t_dev, const float* eps_student_dev, const int64_t* t_dev, const float* x_next_dev,
                         const float* eps_target_dev, const int64_t* t_next_dev, const float* abar_dev, double* loss_dev,
                         float* d_eps_student_dev, int batch, int64_t per_sample, void* stream);

/* ---- image formats either side of the path -----------------------------------
 * preprocess (scripts/inference.py:111-116): uint8 RGB HWC [N][H][W][3] -> fp32 NCHW, x / 127.5 - 1 (fp32 division, then
 * fp32 subtraction: bit-identical to the reference's numpy expression).
 * postprocess (scripts/inference.py:121-127): fp32 NCHW -> uint8 RGB HWC, (y + 1) * 127.5, clip to [0, 255], truncate.
 * Resizing (cv2.resize, scripts/inference.py:109,130) is lcm_image_resize_u8 below. */
int lcm_image_preprocess_u8(const uint8_t* hwc_dev, float* nchw_dev, int batch, int height, int width, void* stream);
int lcm_image_postprocess_u8(const float* nchw_dev, uint8_t* hwc_dev, int batch, int height, int width, void* stream);
/* cv2.resize(img, (dst_w, dst_h)) with the default INTER_LINEAR on 8-bit 3-channel images (scripts/inference.py:109,130):
 * OpenCV's fixed-point bilinear (11-bit coefficients, x taps clamped with the weight reset, y taps clamped by row index),
 * bit-identical to OpenCV 4.x including the exact-2x shrink (where OpenCV switches to area averaging: same values). */
int lcm_image_resize_u8(const uint8_t* src_hwc_dev, int batch, int src_h, int src_w, uint8_t* dst_hwc_dev, int dst_h,
                        int dst_w, void* stream);

/* ---- training step (BASELINE config 5) -----------------------------------------------------------------
 * Replaces, for a plan created with LCM_FLAG_TRAIN (precision fp32, or bf16 on the tensor-core path):
 *   loss.backward() of LowLightDiffusion.compute_loss (src/models/low_light_diffusion.py:250-277, the UNet part of
 *   forward :143-163) and, optionally, clip_grad_norm_ + AdamW.step + EMAModel.update (src/training/trainer.py:296-322,
 *   :98-104, :152-156).
 * Protocol per step:  lcm_unet_forward (x = cat([noisy, low_light]), per-sample t)  ->  lcm_train_loss  ->
 *   lcm_train_backward  ->  [all-reduce of the flat gradient buffer by the caller]  ->  lcm_grad_sumsq +
 *   lcm_adamw_ema_step on flat fp32 buffers  ->  lcm_plan_set_weights_flat.
 * The flat gradient buffer lives inside the workspace at lcm_train_grad_offset_bytes(); it holds every weight's gradient
 * in the reference's state_dict layout at the element offset lcm_train_grad_info reports (16-byte aligned slices, zero
 * padding), in lcm_plan_weight_info order.  A flat PARAMETER buffer with the same offsets feeds lcm_plan_set_weights_flat. */
int lcm_train_num_backward_ops(const lcm_plan* plan);
int lcm_train_backward_op_info(const lcm_plan* plan, int index, const char** name, const char** kernel);
int64_t lcm_train_grad_elems(const lcm_plan* plan);
size_t lcm_train_grad_offset_bytes(const lcm_plan* plan);
/* ready_after_op: index of the last backward op that writes this gradient (it is complete once ops [0, index] have
 * run) — lets the caller all-reduce buckets while the rest of the backward pass executes. */
int lcm_train_grad_info(const lcm_plan* plan, int index, const char** name, int64_t* offset, int64_t* numel, int* ready_after_op);
/* loss (mean over all elements; loss_type 0 mse, 1 l1, 2 huber(delta 1): low_light_diffusion.py:268-275) -> *loss_dev */
int lcm_train_loss(const float* eps_dev, const float* target_dev, int64_t numel, int loss_type, double* loss_dev, void* stream);
/* Backward ops [op_begin, op_end) of the most recent lcm_unet_forward on this plan/workspace (op_end < 0: all).  Inputs
 * are the forward's (xa/xb/t) plus its output eps and the regression target (the noise).  d loss/d eps is
 * grad_scale * (*grad_scale_dev, if given) * loss'(eps - target) / numel.  loss_type 3: target_dev IS the upstream gradient
 * d loss / d eps (autograd through EfficientUNet.forward; no division by numel).  op_begin == 0 zeroes the gradient buffer. */
int lcm_train_backward(lcm_plan* plan, const float* xa_dev, int ca, int64_t xa_batch_stride, const float* xb_dev, int cb,
                       int64_t xb_batch_stride, const int64_t* t_dev, const float* eps_dev, const float* target_dev, int loss_type,
                       float grad_scale, const float* grad_scale_dev, int op_begin, int op_end, void* workspace, void* stream);
/* re-pack every weight from a flat fp32 parameter buffer laid out like the gradient buffer */
int lcm_plan_set_weights_flat(lcm_plan* plan, const float* flat_dev, void* stream);
/* sum of squares of a flat gradient buffer -> *out_dev (fp64) */
int lcm_grad_sumsq(const float* grads_dev, int64_t n, double* out_dev, void* stream);
/* clip_grad_norm_(max_norm) + AdamW + EMA in one pass: g' = g / grad_div ; g' *= min(1, max_norm / (||g'|| + 1e-6)) ;
 * p *= 1 - lr wd ; m, v updates ; p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps) ; ema = d ema + (1 - d) p.
 * ema_dev may be NULL; max_norm <= 0 disables clipping. */
int lcm_adamw_ema_step(float* params_dev, const float* grads_dev, float* exp_avg_dev, float* exp_avg_sq_dev, float* ema_dev,
                       int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay, int step, float ema_decay,
                       const double* grad_sumsq_dev, float grad_div, float max_norm, void* stream);
/* gradient of a named forward intermediate after a full backward pass (LCM_FLAG_TRAIN | LCM_FLAG_TAPS), fp32 NCHW */
int lcm_train_read_grad_tap(lcm_plan* plan, const char* name, float* out_nchw_dev, void* workspace, void* stream);

/* ---- debugging / unit parity --------------------------------------------------------------------
 * Copy a named intermediate of the most recent forward (e.g. "encoder_blocks.0.0.expand",
 * "mid_attn.out", see lcm_plan_tap_info) into fp32 NCHW.  Only valid when the plan was created with
 * LCM_FLAG_TAPS (which disables buffer reuse). */
int lcm_plan_num_taps(const lcm_plan* plan);
int lcm_plan_tap_info(const lcm_plan* plan, int index, const char** name, int* channels, int* height, int* width);
int lcm_plan_read_tap(lcm_plan* plan, const char* name, float* out_nchw_dev, void* workspace, void* stream);

/* Number of kernels the plan launches per UNet forward (for bench.py's gpu_launches). */
int lcm_plan_launches_per_forward(const lcm_plan* plan);
/* Algorithmic HBM bytes / FLOPs of one UNet forward under SURVEY §8(d) accounting (every op of the reference's sequence reads
 * its inputs and writes its output once).  lcm_plan_fused_bytes: the same sum for the kernels THIS plan launches — smaller where
 * a fusion keeps a tensor on chip (the expand -> depthwise kernel never writes the 4x-wide hidden tensor). */
double lcm_plan_algorithmic_bytes(const lcm_plan* plan);
double lcm_plan_fused_bytes(const lcm_plan* plan);
double lcm_plan_algorithmic_flops(const lcm_plan* plan);

/* Per-kernel profile of one forward: runs the forward with a cudaEvent pair around every op and
 * fills up to `cap` records.  Returns the number of ops. */
typedef struct {
  char name[96];
  char kernel[32];
  float ms;
  double bytes;   /* algorithmic bytes of the kernels this op launches */
  double flops;
  double ref_bytes; /* the same under SURVEY 8(d) accounting: what the reference's op sequence moves for this op (>= bytes where a
                     * fusion keeps a tensor on chip: xstats carries the expand op's bytes, xdw_fused the depthwise op's) */
} lcm_op_profile;
int lcm_plan_profile_forward(lcm_plan* plan, const float* xa_dev, int ca, int64_t xa_batch_stride,
                             const float* xb_dev, int cb, int64_t xb_batch_stride, const int64_t* t_dev,
                             float* eps_dev, void* workspace, void* stream, lcm_op_profile* records, int cap);

/* ---- single-kernel entry points (unit parity tests, micro-benchmarks, ncu) --------------------------
 * Raw device pointers; activations NHWC in `precision`; weights are given in the reference's fp32
 * layouts and packed internally.  impl: 0 = CUDA-core kernel, 1 = tcgen05 / tuned kernel (bf16 only).
 * The kernel is launched `repeat` times (>=1); when ms_out != NULL the average device time per launch
 * (CUDA events on `stream`) is stored there. */
typedef struct {
  const void* A;      /* [M][K] activation slice source, channel stride == K                      */
  const void* coef;   /* float2 [images][K] prologue coefficients for this segment, or NULL        */
  int32_t K;
  int32_t mode;       /* 0 none, 1 a*x+b, 2 relu6(a*x+b), 3 silu(a*x+b), 4 a*x (SE gate; b must be 0) */
  int32_t f16;        /* tcgen05 kernel only: this segment is fp16 (a block's hidden tensor), not bf16        */
  int32_t reserved;
} lcm_gemm_seg;
/* out[m][n] = sum_s sum_k xform_s(A_s[m][k]) * W[n][koff_s + k]   (1x1 convs, efficient_unet.py:174,186,199,265,267)
 * impl bit 8 (0x100, tcgen05 kernel only): store the output as fp16 (the expand GEMM's hidden tensor) */
int lcm_op_gemm(const lcm_gemm_seg* segs, int nseg, const float* w_dev, void* out_dev, double* stats_dev, int64_t M,
                int P, int Nc, int precision, int impl, int repeat, float* ms_out, void* stream);
/* dense 3x3 conv, mode 0: stride 1, 1: stride 2, 2: bilinear x2 then stride 1 (efficient_unet.py:367,380-384);
 * w_dev [Co][Ci][3][3], bias [Co] */
int lcm_op_conv3x3(const void* in_dev, const float* w_dev, const float* bias_dev, void* out_dev, double* stats_dev,
                   int N, int Hin, int Win, int Ci, int Co, int mode, int precision, int impl, int repeat,
                   float* ms_out, void* stream);
/* debug aid: clock64 stamps of the tcgen05 GEMM pipeline roles (block 0, first 64 tiles; LCM_TC_DEBUG=64) */
int lcm_debug_timeline(long long* host, int n);
/* depthwise 3x3 with relu6(a*x+b) prologue and pooled-sum epilogue (efficient_unet.py:212-223); w_dev [C][1][3][3];
 * precision LCM_ACT_F16 selects the TMA-streamed fp16 kernel of the tensor-core plan (impl ignored) */
int lcm_op_dwconv(const void* in_dev, const void* coef_dev, const float* w_dev, void* out_dev, double* pool_dev, int N,
                  int H, int W, int C, int precision, int impl, int repeat, float* ms_out, void* stream);
/* fused expand -> GroupNorm2 / FiLM / ReLU6 -> depthwise 3x3 + SE pool of an inverted-residual block (efficient_unet.py:207-220,
 * :97) in three launches, none of which writes the 4x-wide hidden tensor: xstats (t = relu6(GN1(x)) / 6 -> t_dev [N][H][W][Kt]
 * bf16, Gram / column sums), statistics finalisation (stats_dev [N][Nc][2] of the expand output), fused kernel (out_dev fp16
 * [N][H][W][Nc], pool_dev [N][Nc]).  segs: 1-2 bf16 input parts with their GroupNorm1 (a, b) coefficients, mode 2;
 * w_dev [Nc][Kt] fp32; coef2_dev [N][Nc][2] fp32; wdw_dev [Nc][1][3][3].  Note: the CALLER supplies coef2 (in the plan it is
 * derived from stats_dev by the GroupNorm finalisation between the second and the third launch). */
int lcm_op_xdw(const lcm_gemm_seg* segs, int nseg, const float* w_dev, const void* coef2_dev, const float* wdw_dev, void* t_dev,
               void* out_dev, double* pool_dev, double* stats_dev, int N, int H, int W, int Nc, int repeat, float* ms_out,
               void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LCM_UNET_H_ */
