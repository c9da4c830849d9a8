// Launch wrappers of every kernel of the LCM-UNet hot path (host-callable, enqueue only).
// `bf16act` selects the activation storage type: 0 = fp32 (verification mode), 1 = bf16.
#pragma once
#include "common.cuh"

namespace lcm {

// ---- a2: SinusoidalPosEmb + time_mlp (efficient_unet.py:60-76, 412-417) ------------------------
// t_dev (int64[N]) or, when null, the scalar t for every sample.  Writes silu(t_emb) [N][ted]
// (every consumer applies SiLU first, efficient_unet.py:189-190) and optionally t_emb itself.
void launch_time_embed(const long long* t_dev, long long t_scalar, int N, int base, int ted, const float* w1,
                       const float* b1, const float* w3, const float* b3, float* temb, float* silu_temb,
                       cudaStream_t st);

// ---- a4.3: all blocks' FiLM linears in one launch: out[n][r] = b[r] + W[r][:] . silu_temb[n][:] ---
void launch_film(const float* silu_temb, const float* W, const float* b, float* out, int N, int rows, int ted,
                 cudaStream_t st);

// ---- a4.1: GroupNorm finalise: channel statistics -> per-(image, channel) affine (a, b) ------------
// channel c < C0 comes from stats0, the rest from stats1 (concat inputs).  film (optional) points at
// this block's [scale | shift] rows: a' = a(1+scale), b' = b(1+scale)+shift.
void launch_gn_coef(const double* stats0, int C0, const double* stats1, int C1, int groups, double count,
                    const float* gamma, const float* beta, const float* film, int film_ld, float2* coef, int N,
                    cudaStream_t st);

// ---- a4.2: 1x1 convs as GEMM, CUDA-core version (fp32 mode and cross-check) -----------------------
// W is row-major [Nc][Ktot] in the activation storage type.
void launch_gemm_simt(const GemmParams& p, int bf16act, cudaStream_t st);

// ---- a4.4 + SE pool: depthwise 3x3 with affine+ReLU6 prologue, pooled sums epilogue ---------------
// in/out NHWC [N][H][W][C]; coef [N][C]; w [9][C] fp32; pool [N][C] fp64 sums (atomically accumulated).
void launch_dwconv(const void* in, const float2* coef, const float* w, void* out, double* pool, int N, int H, int W,
                   int C, int bf16act, int fast, cudaStream_t st);

// tcgen05 product path: fp16 NHWC in and out (the block's hidden tensors), TMA-streamed rows (dwconv_stream.cu).
// Returns non-zero when the shape is not supported.
int launch_dwconv_f16(const void* in, const float2* coef, const float* w, void* out, double* pool, int N, int H, int W,
                      int C, int num_sms, cudaStream_t st);

// ---- a4.5: SE gate: sigmoid(fc2(relu6(fc1(mean)))) -> coef (gate, 0); one launch (8-CTA clusters, 8 images each).
// Returns non-zero when the layer does not fit (shared memory).
int launch_se_gate(const double* pool, float inv_count, const float* w1, const float* b1, const float* w2,
                   const float* b2, float2* coef, int N, int C, int SQ, cudaStream_t st);

// ---- a3/a6/a7: dense 3x3 convs ------------------------------------------------------------------
enum Conv3Mode : int { CONV_S1 = 0, CONV_S2 = 1, CONV_UP2 = 2 };
// generic NHWC->NHWC, CUDA-core implicit GEMM; W row-major [Co][9*Ci] (k = tap*Ci + ci), bias fp32.
// Hin,Win are the stored input size; output is Hin x Win (S1), Hin/2 (S2, pad 1) or 2*Hin (UP2: bilinear x2 first).
void launch_conv3x3_simt(const void* in, const void* Wt, const float* bias, void* out, double* stats, int N,
                         int Hin, int Win, int Ci, int Co, int mode, int bf16act, cudaStream_t st);
// bf16 NHWC bilinear x2 (align_corners=False) materialisation for the tensor-core up-convolution
void launch_upsample2x(const void* in, void* out, int N, int H, int W, int C, cudaStream_t st);
// init_conv: x = cat([xa, xb]) fp32 NCHW -> NHWC; w fp32 [9*Cin][Co] (k = tap*Cin + ci), bias.
void launch_init_conv(const float* xa, int ca, long long sa, const float* xb, int cb, long long sb, const float* w,
                      const float* bias, void* out, double* stats, int N, int H, int W, int Co, int bf16act,
                      cudaStream_t st);
// packed-fp16 versions for the bf16 tensor-core plan (conv_edge.cu); false = shape not covered, caller falls back
bool launch_init_conv_h2(const float* xa, int ca, long long sa, const float* xb, int cb, long long sb, const float* w,
                         const float* bias, void* out, double* stats, int N, int H, int W, int Co, cudaStream_t st);
// final: GN+SiLU prologue, 3x3 conv to Cout (<=4), fp32 NCHW output; optional fused LCMScheduler.step.
struct FinalStep {
  int enabled;           // 0: write eps only
  const float* noise;    // null on the last step
  float* latents;        // x_t in, x_prev out (in place)
  float* clamped;        // optional clamp(x_prev,-1,1) output
  float* trace;          // optional copy of x_prev
  float sb_t, sa_t, sa_p, sb_p;
};
void launch_final_conv(const void* in, const float2* coef, const float* w, const float* bias, float* eps,
                       const FinalStep& step, int N, int H, int W, int Ci, int Co, int bf16act, cudaStream_t st);

bool launch_final_conv_h2(const void* in, const float2* coef, const float* w, const float* bias, float* eps,
                          const FinalStep& step, int N, int H, int W, int Ci, int Co, cudaStream_t st);

// ---- a5: linear attention --------------------------------------------------------------------------
// qkv NHWC [N][P][3*inner] (q | k | v, each heads*32); state [N][heads][32][33] fp64 (col 32 = k_sum), zeroed.
// (fp64 accumulators make the atomic reductions order-insensitive to ~1e-16, i.e. run-to-run reproducible)
void launch_attn_kv(const void* qkv, double* state, int N, int P, int heads, int bf16act, cudaStream_t st);
void launch_attn_apply(const void* qkv, const double* state, void* out, int N, int P, int heads, int bf16act,
                       cudaStream_t st);
// StandardAttention core (efficient_unet.py:344-349): out = softmax(q k^T d^-0.5) v per head (d = 32), qkv as above
void launch_attn_softmax(const void* qkv, void* out, int N, int P, int heads, int bf16act, cudaStream_t st);
// y = a*u + b + x (to_out GroupNorm + residual), channel statistics of y.
void launch_affine_residual(const void* u, const float2* coef, const void* x, void* y, double* stats, int N, int P,
                            int C, int bf16act, cudaStream_t st);

// ---- a11: LCMScheduler.step / add_noise stand-alone ---------------------------------------------------
void launch_lcm_step(const float* eps, const float* sample, const float* noise, float* prev, float* x0,
                     long long numel, int prediction, float sb_t, float sa_t, float sa_p, float sb_p, cudaStream_t st);
// image formats either side of the path (scripts/inference.py:111-116, 121-127)
void launch_image_pre_u8(const uint8_t* hwc, float* nchw, int N, int H, int W, cudaStream_t st);
void launch_image_post_u8(const float* nchw, uint8_t* hwc, int N, int H, int W, cudaStream_t st);
void launch_image_resize_u8(const uint8_t* src, int N, int sh, int sw, uint8_t* dst, int dh, int dw, cudaStream_t st);
void launch_add_f32(const float* a, const float* b, float* out, long long n, cudaStream_t st);
void launch_fill_float2(float2* p, float2 v, long long n, cudaStream_t st);
void launch_ddim_step(const float* x_t, const float* eps, const long long* t, const long long* t_next, const float* abar,
                      float* x_next, int batch, long long per_sample, cudaStream_t st);
void launch_consistency_loss(const float* x_t, const float* eps_s, const long long* t, const float* x_next, const float* eps_tgt,
                             const long long* t_next, const float* abar, double* loss, float* d_eps, int batch, long long per_sample,
                             cudaStream_t st);
void launch_lcm_mix(const float* a, const float* b, const long long* t, const float* abar, float* out, int batch,
                    long long per_sample, int velocity, cudaStream_t st);

// ---- weight packing ---------------------------------------------------------------------------------
enum PackKind : int {
  PACK_MAT_T = 6,    // src [Cc][src_ld] -> logical W(r, off + c) = src[c][src_col0 + r]   (transposed copy: dgrad weights)
  PACK_CONV3_T = 7,  // src [Co][Ci][3][3] -> logical W(ci, off + (8 - tap)*tap_stride + co): the transposed conv as a
                     // conv over dY with flipped taps and swapped channel roles (R = Ci rows, Cc = Co)
  PACK_COPY = 0,     // dst_f32[i] = src[i]
  PACK_MAT = 1,      // src [R][Cc] -> logical W(r, off + c)
  PACK_CONV3 = 2,    // src [Co][Ci][3][3] -> logical W(co, off + tap*Ci + ci)
  PACK_DW = 3,       // src [C][1][3][3] -> dst_f32[tap*C + c]
  PACK_CONV3_KN = 4, // src [Co][Ci][3][3] -> dst_f32[(tap*Ci+ci)*Co + co]
  PACK_IDENTITY = 5  // logical W(r, off + c) = (r == c)
};
enum WLayout : int { WL_ROWMAJOR = 0, WL_UMMA = 1 };
struct PackJob {
  int kind;
  int layout;      // WLayout (for logical-matrix kinds)
  int bf16;        // destination element type for logical-matrix kinds: 0 fp32, 1 bf16, 2 fp16 (fp16 segments of a tcgen05 GEMM)
  void* dst;       // plan-time: byte offset in the weight arena; resolved to a pointer before launch
  int R, Cc, Ci;   // rows (out channels), columns taken from this source, conv input channels
  int src_ld;      // PACK_MAT: source row stride (elements)
  int src_col0;    // PACK_MAT: first source column of the slice
  int tap_stride;  // PACK_CONV3: logical columns per tap (Ci, or Ci rounded up to 64 for WL_UMMA)
  int ld;          // K extent of the logical matrix (padded for WL_UMMA)
  int off;         // column offset of this source in the logical matrix
  int block_n;     // WL_UMMA: rows per N tile
  float scale;     // logical-matrix kinds: values are multiplied by this (0 = 1; the expand kernel folds relu6's 6 here)
};
void launch_pack(const PackJob& job, const float* src, cudaStream_t st);
// element offset of logical (n, k) in the tcgen05 weight image (shared with the kernel's consumer side)
__host__ __device__ inline long long umma_weight_offset(int n, int k, int Ktot, int block_n) {
  // [n tile][k chunk of 64][row within tile: 128 B, 16-byte units XOR-swizzled by (row & 7)]
  int nchunks = (Ktot + 63) / 64;
  int nt = n / block_n, r = n % block_n, kc = k / 64, kk = k % 64;
  long long base = ((long long)nt * nchunks + kc) * (long long)block_n * 64;
  int unit = (kk / 8) ^ (r & 7);
  return base + (long long)r * 64 + unit * 8 + (kk % 8);
}

// NHWC (storage type: bf16act 0 fp32, 1 bf16, 2 fp16) -> fp32 NCHW, for taps
void launch_nhwc_to_nchw(const void* in, float* out, int N, int H, int W, int C, int bf16act, cudaStream_t st);

// ---- tcgen05 kernels (gemm_tcgen05.cu) ------------------------------------------------------------------
// 1x1 GEMM and dense 3x3 implicit GEMM on the 5th-gen tensor cores.  W in WL_UMMA layout (bf16).
struct ConvGeom {   // conv3x3 producer geometry (mode < 0: plain 1x1 GEMM)
  int mode;         // Conv3Mode or -1
  int Hin, Win, Hout, Wout, Ci;
  const float* bias;
};
int launch_gemm_tc(const GemmParams& p, const ConvGeom& g, int block_n, int num_sms, cudaStream_t st);
// Expand GEMM specialisation (gemm_expand.cu): bf16 segments with relu6(a x + b) prologue, fp16 output + statistics,
// weights packed with block_n = 128 and scale = 6.  `supported` is a pure shape test (plan time).
bool gemm_expand_supported(int nseg, const int* segK, int Nc, int P);
size_t gemm_expand_scratch_bytes(int images);   // Gram / column-sum scratch; must be zero at launch (zero_scratch: cleared on the stream)
// fused expand -> GN2 / FiLM / ReLU6 -> depthwise 3x3 + SE pool (xdw_fused.cu); the GroupNorm2 coefficients come from the
// statistics-only pass of the expand kernel (launch_gemm_expand(..., stats_only = true))
// (xstats.cu: t = relu6(GN1(x)) / 6 as one dense bf16 tensor + its per-image Gram matrix / column sums;
//  launch_expand_stats_finalize turns those into the statistics of the expand output h1, which is never materialised)
bool xstats_supported(int Ktot, int P);
int launch_xstats(const GemmParams& g, void* t, void* scratch, int num_sms, cudaStream_t st);
int launch_expand_stats_finalize(void* scratch, const void* W, double* stats, int images, int Nc, int nchunks, cudaStream_t st, int Ktot = 0);
bool xdw_fused_supported(int nseg, const int* segK, int Nc, int H, int W);
int launch_xdw_fused(const void* t, int Kt, const void* Wp, int Nc, const float2* coef2, const float* wdw, void* out, double* pool,
                     int N, int H, int W, int num_sms, cudaStream_t st);
int launch_gemm_expand(const GemmParams& p, void* scratch, bool zero_scratch, int num_sms, cudaStream_t st, bool stats_only = false);
// proj_stream.cu: the level-0 project GEMM (K = 128 fp16 SE-gated + 32 bf16 -> N = 32) as a barrier-free streaming kernel on mma.sync;
// W row-major [Nc][160] 16-bit (fp16 | bf16 columns)
bool proj_stream_supported(int nseg, const int* segK, const int* seg_f16, const int* seg_mode, int Nc, int P);
int launch_proj_stream(const GemmParams& g, int num_sms, cudaStream_t st);
// gemm_wide.cu: the same operation for 128 <= K <= 448 (activation tile stationary, weights streamed; statistics from the
// epilogue, accumulated into p.stats).  Weights packed like gemm_expand's (block_n = 128, x6).
bool gemm_wide_supported(int nseg, const int* segK, int Nc, int P);
int launch_gemm_wide(const GemmParams& p, int num_sms, cudaStream_t st);
int gemm_wide_read_profile(long long* host8);   // LCM_W_DEBUG & 16
int gemm_expand_read_timeline(long long* host, int n);   // debug: LCM_X_TIMELINE=1
int gemm_tc_pick_block_n(int Nc);
int gemm_tc_read_timeline(long long* host, int n);   // debug: per-tile clock stamps of block 0 (LCM_TC_DEBUG & 64)

// =====================================================================================================
// Training step (BASELINE config 5): backward of every op above + loss + optimizer (train_kernels.cu).
// Element types are run-time codes: one backward op touches the residual stream (bf16), hidden tensors (fp16) and
// gradients (bf16) on the tensor-core plan, fp32 everywhere on the fp32 plan.
enum DType : int { DT_F32 = 0, DT_BF16 = 1, DT_F16 = 2 };
inline size_t dtype_size(int dt) { return dt == DT_F32 ? 4 : 2; }

// g <- g * [0 < a x + b < 6] (mode 1, in place) and T1 += sum_p g, T2 += sum_p g x per (image, channel) (mode 0: no write)
void launch_bwd_mask_reduce(void* g, int dtg, int ldg, int goff, const void* x, int dtx, int ldx, int xoff, const float2* coef,
                            int coef_ld, double* t12, int t_ld, int N, int P, int Cs, int mode, cudaStream_t st);
// dst (=|+=) A g + B x + C (+ r) with (A, B, C) = coef4[n][coff + c]
void launch_bwd_affine3(const void* g, int dtg, int ldg, int goff, const void* x, int dtx, int ldx, int xoff,
                        const float4* coef4, int coef_ld, int coff, const void* r, int dtr, int ldr, int roff, void* dst,
                        int dtd, int ldd, int doff, int accumulate, int N, int P, int Cs, cudaStream_t st);
void launch_bwd_add(const void* src, int dts, int lds, int soff, void* dst, int dtd, int ldd, int doff, int accumulate,
                    long long rows, int Cs, cudaStream_t st);
// GroupNorm (+FiLM) backward finalise: (T1, T2) + forward statistics -> (A, B, C), d gamma, d beta, d FiLM rows
void launch_gn_bwd_coef(const double* t12, const double* stats0, int C0, const double* stats1, int C1, int groups, double count,
                        const float* gamma, const float* beta, const float* film, float* dfilm, int film_ld, float4* coef4,
                        float* dgamma, float* dbeta, int N, cudaStream_t st, const float* ascale = nullptr, int ascale_stride = 0);
// coef_se[n][c] = (gate s, dpm/P s, 1/s, 0); s = 1 unless dq_stats (sum, sum^2 of dq per (image, channel)) is given
int launch_se_bwd_vec(const double* pool, float inv_count, const float* w1, const float* b1, const float* w2, const float2* gate,
                      const double* t12, const double* dq_stats, float4* coef_se, float* v_pm, float* v_z, float* v_ds2,
                      float* v_dz1, int N, int C, int SQ, cudaStream_t st);
void launch_outer_sum(const float* A, int lda, const float* B, int ldb, float* dW, float* dbias, int N, int R, int Cc,
                      cudaStream_t st);
void launch_dwconv_bwd(const void* dq, int dtg, const float4* coef_se, const void* h1, int dth, const float2* coef2, const float* w,
                       void* du, double* t12, float* dW, int N, int H, int W, int C, int num_sms, cudaStream_t st);
// streaming packed-fp16 version (dwconv_bwd_stream.cu): dq bf16, h1 fp16, du fp16 SCALED by the s of coef_se; C % 64 == 0
int launch_dwconv_bwd_stream(const void* dq, const float4* cse, const void* h1, const float2* coef2, const float* w, void* du,
                             double* t12, float* dW, int N, int H, int W, int C, int num_sms, cudaStream_t st);
void launch_wgrad_1x1(const GemmParams& p, const int* seg_dt, const void* dY, int dty, float* const* dst, const int* dst_ld,
                      int num_sms, cudaStream_t st);
// the same on the tensor cores (wgrad_tc.cu, bf16 plan); non-zero = shape not covered, use the CUDA-core kernel
int launch_wgrad_tc(const GemmParams& p, const int* seg_dt, const void* dY, int dty, float* const* dst, const int* dst_ld,
                    int num_sms, cudaStream_t st, float* img_dst = nullptr);
// project weight gradient + SE gate gradient from the per-image products R[n][k][o] = sum_p [h2 | x][p][k] dY[p][o]
// (launch_wgrad_tc, per-image mode, h2 NOT gated):  dgate[n][c] = sum_o Wp[o][c] R[n][c][o] -> t12[(n Ch + c) 2 + 1] (+=),
// dWp[o][c] += sum_n gate[n][c] R[n][c][o],  dWskip[o][ci] += sum_n R[n][Ch + ci][o] (dWs may be null: identity residual)
void launch_se_project_combine(const float* R, const float* Wp, const float2* gate, double* t12, float* dWp, float* dWs, int N,
                               int Ch, int Ci, int Co, cudaStream_t st);
// dense 3x3 conv (pad 1, stride 1 or 2) on the tensor cores (bf16); H, W = INPUT size; bias via launch_colsum
int launch_wgrad_conv3_tc(const void* in, const void* dY, float* dW, int N, int H, int W, int Ci, int Co, int stride, int num_sms,
                          cudaStream_t st);
void launch_zero_insert2x(const void* in, void* out, int N, int H, int W, int C, cudaStream_t st);   // 16-bit, C % 8 == 0
void launch_colsum(const void* g, int dt, long long rows, int C, float* out, cudaStream_t st);
void launch_wgrad_conv3(const void* in, int dti, const void* dY, int dty, float* dW, float* dbias, int N, int Hin, int Win, int Ci,
                        int Co, int mode, int num_sms, cudaStream_t st);
void launch_upsample2x_any(const void* in, void* out, int dt, int N, int H, int W, int C, cudaStream_t st);
void launch_upsample2x_bwd(const void* dout, void* din, int dt, int N, int H, int W, int C, int accumulate, cudaStream_t st);
void launch_attn_bwd(const void* qkv, int dtq, const double* state, const void* dO, int dtg, void* dqkv, double* dstate, int N, int P,
                     int heads, cudaStream_t st);
void launch_film_bwd_input(const float* dfilm, const float* W, float* dst, int N, int rows, int ted, cudaStream_t st);
void launch_time_mlp_bwd(const long long* t_dev, long long t_scalar, int N, int base, int ted, const float* w1, const float* b1,
                         const float* w3, const float* b3, const float* dst, float* dw1, float* db1, float* dw3, float* db3,
                         cudaStream_t st);
void launch_loss(const float* eps, const float* target, long long numel, int type, double* out, cudaStream_t st);
int launch_final_conv_bwd(const void* h, int dth, const float2* coef, const float* w, const float* eps, const float* target,
                          int loss_type, float gscale, const float* gscale_dev, void* dpre, int dtg, double* t12, float* dW,
                          float* dbias, int N, int H, int W, int Ci, int Co, int num_sms, cudaStream_t st);
int launch_init_conv_wgrad(const float* xa, int ca, long long sa, const float* xb, int cb, long long sb, const void* dY, int dtg,
                           float* dW, float* dbias, int N, int H, int W, int Co, int num_sms, cudaStream_t st);
void launch_sumsq(const float* g, long long n, double* out, cudaStream_t st);
void launch_adamw_ema(float* p, const float* g, float* m, float* v, float* ema, long long n, float lr, float beta1, float beta2,
                      float eps, float wd, int step, float ema_decay, const double* sumsq, float grad_div, float max_norm,
                      cudaStream_t st);
// transposed dense 3x3 conv on the CUDA cores (input gradient): dX [N][Hin][Win][Ci] from dY [N][Hout][Wout][Co];
// Wt row-major [Ci][9*Co] packed with PACK_CONV3_T.  mode CONV_S1 or CONV_S2 (forward stride).
void launch_conv3x3_dgrad_simt(const void* dY, const void* Wt, void* dX, int N, int Hin, int Win, int Ci, int Co, int mode,
                               int bf16act, cudaStream_t st);

}  // namespace lcm
