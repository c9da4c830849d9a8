// Single-kernel C entry points (include/lcm_unet.h, "single-kernel entry points"): unit parity, ncu.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/lcm_unet.h"
#include "kernels.h"

using namespace lcm;

namespace {
struct Timer {
  cudaEvent_t a = nullptr, b = nullptr;
  cudaStream_t st;
  float* out;
  int repeat;
  Timer(cudaStream_t s, float* o, int r) : st(s), out(o), repeat(r) {
    if (out) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, st); }
  }
  void stop() {
    if (out) {
      cudaEventRecord(b, st); cudaEventSynchronize(b);
      float ms = 0; cudaEventElapsedTime(&ms, a, b); *out = ms / repeat;
      cudaEventDestroy(a); cudaEventDestroy(b);
    }
  }
};
int finish(cudaStream_t st) {
  cudaError_t e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) e = cudaGetLastError();
  return e == cudaSuccess ? 0 : LCM_ERR_CUDA;
}
}  // namespace

extern "C" {

int lcm_debug_timeline(long long* host, int n) {
  if (n == 8) return gemm_wide_read_profile(host);   // MMA-warp cycle breakdown of gemm_wide.cu
  return n < 0 ? gemm_expand_read_timeline(host, -n) : gemm_tc_read_timeline(host, n);
}

int lcm_op_gemm(const lcm_gemm_seg* segs, int nseg, const float* w_dev, void* out_dev, double* stats_dev, int64_t M,
                int P, int Nc, int precision, int impl, int repeat, float* ms_out, void* stream) {
  if (!segs || nseg < 1 || nseg > LCM_MAX_SEGS || !w_dev || !out_dev || repeat < 1) return LCM_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  const bool out16 = (impl & 0x100) != 0;
  impl &= 0xff;
  const bool bf = precision == LCM_PREC_BF16, tc = bf && impl == 1;
  if (!tc && out16) return LCM_ERR_INVALID;
  GemmParams gp{};
  gp.out_f16 = out16 ? 1 : 0;
  int Ktot = 0, Kpad = 0;
  std::vector<int> off, poff;
  for (int i = 0; i < nseg; ++i) { off.push_back(Ktot); poff.push_back(Kpad); Ktot += segs[i].K; Kpad += (segs[i].K + 63) / 64 * 64; }
  bool expand = tc && out16 && stats_dev && nseg <= 2;
  if (expand) {
    int segK[2] = {0, 0};
    for (int i = 0; i < nseg; ++i) { segK[i] = segs[i].K; expand = expand && segs[i].coef && segs[i].mode == XF_AFFINE_RELU6 && !segs[i].f16; }
    expand = expand && gemm_expand_supported(nseg, segK, Nc, P) && M % 128 == 0;
  }
  bool wide = tc && out16 && stats_dev && nseg <= 2 && !expand;
  if (wide) {
    int segK[2] = {0, 0};
    for (int i = 0; i < nseg; ++i) { segK[i] = segs[i].K; wide = wide && segs[i].coef && segs[i].mode == XF_AFFINE_RELU6 && !segs[i].f16; }
    wide = wide && gemm_wide_supported(nseg, segK, Nc, P) && M % 128 == 0;
  }
  bool pstream = false;   // level-0 project shape: the streaming kernel (proj_stream.cu), row-major 16-bit weights
  if (tc && !out16 && stats_dev && nseg == 2 && segs[0].coef && M % P == 0) {
    const int sk[2] = {segs[0].K, segs[1].K}, sf[2] = {segs[0].f16 ? 1 : 0, segs[1].f16 ? 1 : 0};
    const int sm[2] = {segs[0].coef ? segs[0].mode : XF_NONE, segs[1].coef ? segs[1].mode : XF_NONE};
    pstream = proj_stream_supported(2, sk, sf, sm, Nc, P);
  }
  const int block_n = tc ? ((expand || wide) ? 128 : gemm_tc_pick_block_n(Nc)) : 0;
  const size_t wbytes = tc ? (size_t)Nc * Kpad * 2 : (size_t)Nc * Ktot * (bf ? 2 : 4);
  void* wbuf = nullptr;
  if (cudaMalloc(&wbuf, wbytes) != cudaSuccess) return LCM_ERR_CUDA;
  cudaMemsetAsync(wbuf, 0, wbytes, st);
  for (int i = 0; i < nseg; ++i) {
    PackJob j{};
    if (segs[i].f16 && !tc) { cudaFree(wbuf); return LCM_ERR_INVALID; }
    j.kind = PACK_MAT; j.layout = (tc && !pstream) ? WL_UMMA : WL_ROWMAJOR; j.bf16 = segs[i].f16 ? 2 : (bf ? 1 : 0); j.dst = wbuf; j.R = Nc; j.Cc = segs[i].K;
    j.src_ld = Ktot; j.src_col0 = off[i]; j.ld = (tc && !pstream) ? Kpad : Ktot; j.off = (tc && !pstream) ? poff[i] : off[i]; j.block_n = block_n;
    j.scale = (tc && segs[i].coef && segs[i].mode == XF_AFFINE_RELU6) ? 6.f : 0.f;   // relu6 prologue = 6 sat(.) on the tcgen05 path
    launch_pack(j, w_dev, st);
    gp.seg[i].A = segs[i].A; gp.seg[i].K = segs[i].K; gp.seg[i].ld = segs[i].K;
    gp.seg[i].coef = (const float2*)segs[i].coef; gp.seg[i].coef_ld = segs[i].K; gp.seg[i].coef_off = 0;
    gp.seg[i].mode = segs[i].coef ? segs[i].mode : XF_NONE;
    gp.seg[i].f16 = segs[i].f16 ? 1 : 0;
  }
  gp.nseg = nseg; gp.Ktot = Ktot; gp.W = wbuf; gp.out = out_dev; gp.stats = stats_dev; gp.M = M; gp.P = P; gp.Nc = Nc;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int rc = 0;
  void* xscratch = nullptr;
  if (expand && cudaMalloc(&xscratch, gemm_expand_scratch_bytes((int)(M / P))) != cudaSuccess) { cudaFree(wbuf); return LCM_ERR_CUDA; }
  Timer t(st, ms_out, repeat);
  for (int r = 0; r < repeat && rc == 0; ++r) {
    if ((expand || wide) && repeat > 1 && stats_dev) cudaMemsetAsync(stats_dev, 0, (size_t)(M / P) * Nc * 2 * sizeof(double), st);   // the finalisation kernel owns its entries
    if (pstream) rc = launch_proj_stream(gp, sms, st);
    else if (expand) rc = launch_gemm_expand(gp, xscratch, true, sms, st);
    else if (wide) rc = launch_gemm_wide(gp, sms, st);
    else if (tc) { ConvGeom g{}; g.mode = -1; rc = launch_gemm_tc(gp, g, block_n, sms, st); }
    else launch_gemm_simt(gp, bf, st);
  }
  t.stop();
  int rc2 = finish(st);
  cudaFree(wbuf);
  if (xscratch) cudaFree(xscratch);
  return rc ? LCM_ERR_INVALID : rc2;
}

int lcm_op_conv3x3(const void* in_dev, const float* w_dev, const float* bias_dev, void* out_dev, double* stats_dev,
                   int N, int Hin, int Win, int Ci, int Co, int mode, int precision, int impl, int repeat,
                   float* ms_out, void* stream) {
  if (!in_dev || !w_dev || !bias_dev || !out_dev || repeat < 1 || mode < 0 || mode > 2) return LCM_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  const bool bf = precision == LCM_PREC_BF16, tc = bf && impl == 1;
  const int Cpad = (Ci + 63) / 64 * 64;
  const int block_n = tc ? gemm_tc_pick_block_n(Co) : 0;
  const size_t wbytes = tc ? (size_t)Co * 9 * Cpad * 2 : (size_t)Co * 9 * Ci * (bf ? 2 : 4);
  void* wbuf = nullptr;
  if (cudaMalloc(&wbuf, wbytes) != cudaSuccess) return LCM_ERR_CUDA;
  cudaMemsetAsync(wbuf, 0, wbytes, st);
  PackJob j{};
  j.kind = PACK_CONV3; j.bf16 = bf; j.R = Co; j.Ci = Ci; j.dst = wbuf;
  if (tc) { j.layout = WL_UMMA; j.ld = 9 * Cpad; j.tap_stride = Cpad; j.block_n = block_n; }
  else { j.layout = WL_ROWMAJOR; j.ld = 9 * Ci; j.tap_stride = Ci; }
  launch_pack(j, w_dev, st);
  const int Ho = mode == CONV_S2 ? Hin / 2 : (mode == CONV_UP2 ? Hin * 2 : Hin);
  const int Wo = mode == CONV_S2 ? Win / 2 : (mode == CONV_UP2 ? Win * 2 : Win);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int rc = 0;
  Timer t(st, ms_out, repeat);
  for (int r = 0; r < repeat && rc == 0; ++r) {
    if (tc) {
      GemmParams gp{};
      gp.nseg = 1; gp.seg[0].A = in_dev; gp.seg[0].K = 9 * Ci; gp.seg[0].ld = Ci; gp.seg[0].mode = XF_NONE;
      gp.Ktot = 9 * Ci; gp.W = wbuf; gp.out = out_dev; gp.stats = stats_dev; gp.P = Ho * Wo; gp.M = (long long)N * Ho * Wo; gp.Nc = Co;
      ConvGeom cg{mode, Hin, Win, Ho, Wo, Ci, bias_dev};
      rc = launch_gemm_tc(gp, cg, block_n, sms, st);
    } else {
      launch_conv3x3_simt(in_dev, wbuf, bias_dev, out_dev, stats_dev, N, Hin, Win, Ci, Co, mode, bf, st);
    }
  }
  t.stop();
  int rc2 = finish(st);
  cudaFree(wbuf);
  return rc ? LCM_ERR_INVALID : rc2;
}

int lcm_op_dwconv(const void* in_dev, const void* coef_dev, const float* w_dev, void* out_dev, double* pool_dev, int N,
                  int H, int W, int C, int precision, int impl, int repeat, float* ms_out, void* stream) {
  if (!in_dev || !coef_dev || !w_dev || !out_dev || !pool_dev || repeat < 1 || C % 32) return LCM_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  const bool bf = precision == LCM_PREC_BF16;
  float* wbuf = nullptr;
  if (cudaMalloc((void**)&wbuf, (size_t)9 * C * 4) != cudaSuccess) return LCM_ERR_CUDA;
  PackJob j{}; j.kind = PACK_DW; j.dst = wbuf; j.R = C;
  launch_pack(j, w_dev, st);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int bad = 0;
  Timer t(st, ms_out, repeat);
  for (int r = 0; r < repeat && !bad; ++r) {
    if (precision == LCM_ACT_F16) bad = launch_dwconv_f16(in_dev, (const float2*)coef_dev, wbuf, out_dev, pool_dev, N, H, W, C, sms, st);
    else launch_dwconv(in_dev, (const float2*)coef_dev, wbuf, out_dev, pool_dev, N, H, W, C, bf, impl, st);
  }
  t.stop();
  int rc = finish(st);
  cudaFree(wbuf);
  return bad ? LCM_ERR_INVALID : rc;
}

int lcm_op_xdw(const lcm_gemm_seg* segs, int nseg, const float* w_dev, const void* coef2_dev, const float* wdw_dev, void* t_dev,
               void* out_dev, double* pool_dev, double* stats_dev, int N, int H, int W, int Nc, int repeat, float* ms_out,
               void* stream) {
  if (!segs || nseg < 1 || nseg > 2 || !w_dev || !coef2_dev || !wdw_dev || !t_dev || !out_dev || !pool_dev || !stats_dev || repeat < 1)
    return LCM_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  int Kt = 0;
  GemmParams gp{};
  for (int i = 0; i < nseg; ++i) {
    if (!segs[i].coef || segs[i].mode != XF_AFFINE_RELU6 || segs[i].f16) return LCM_ERR_INVALID;
    gp.seg[i].A = segs[i].A; gp.seg[i].K = segs[i].K; gp.seg[i].ld = segs[i].K;
    gp.seg[i].coef = (const float2*)segs[i].coef; gp.seg[i].coef_ld = segs[i].K; gp.seg[i].coef_off = 0;
    gp.seg[i].mode = XF_AFFINE_RELU6; gp.seg[i].f16 = 0;
    Kt += segs[i].K;
  }
  gp.nseg = nseg; gp.P = H * W; gp.M = (long long)N * H * W;
  if (Kt % 16 || Kt > 128 || !xstats_supported(Kt, H * W) || Nc % 64 || Nc < 128 || W % 64 || H % 2) return LCM_ERR_INVALID;
  const int Kpad = (Kt + 63) / 64 * 64;
  const int Npad = (Nc + 127) / 128 * 128;   // n-blocks of 128 rows, the last one zero-padded
  void* wbuf = nullptr; float* dwbuf = nullptr; void* scratch = nullptr;
  if (cudaMalloc(&wbuf, (size_t)Npad * Kpad * 2) != cudaSuccess || cudaMalloc((void**)&dwbuf, (size_t)9 * Nc * 4) != cudaSuccess ||
      cudaMalloc(&scratch, gemm_expand_scratch_bytes(N)) != cudaSuccess) {
    cudaFree(wbuf); cudaFree(dwbuf); cudaFree(scratch);
    return LCM_ERR_CUDA;
  }
  cudaMemsetAsync(wbuf, 0, (size_t)Npad * Kpad * 2, st);
  {
    PackJob j{};
    j.kind = PACK_MAT; j.layout = WL_UMMA; j.bf16 = 1; j.dst = wbuf; j.R = Nc; j.Cc = Kt; j.src_ld = Kt; j.src_col0 = 0; j.ld = Kpad; j.off = 0;
    j.block_n = 128; j.scale = 6.f;
    launch_pack(j, w_dev, st);
    PackJob d{}; d.kind = PACK_DW; d.dst = dwbuf; d.R = Nc;
    launch_pack(d, wdw_dev, st);
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int rc = 0;
  Timer t(st, ms_out, repeat);
  for (int r = 0; r < repeat && rc == 0; ++r) {
    cudaMemsetAsync(scratch, 0, gemm_expand_scratch_bytes(N), st);
    cudaMemsetAsync(stats_dev, 0, (size_t)N * Nc * 2 * sizeof(double), st);
    cudaMemsetAsync(pool_dev, 0, (size_t)N * Nc * sizeof(double), st);
    const bool dbg = getenv("LCM_XDW_SYNC") != nullptr;   // debugging aid: locate a failing launch
    auto stage = [&](const char* what) {
      if (!dbg) return;
      cudaError_t e = cudaStreamSynchronize(st);
      if (e == cudaSuccess) e = cudaGetLastError();
      fprintf(stderr, "lcm_op_xdw: %s -> rc %d, %s\n", what, rc, cudaGetErrorString(e));
    };
    rc = launch_xstats(gp, t_dev, scratch, sms, st); stage("xstats");
    if (!rc) { rc = launch_expand_stats_finalize(scratch, wbuf, stats_dev, N, Nc, Kpad / 64, st, Kt); stage("finalize"); }
    if (!rc) { rc = launch_xdw_fused(t_dev, Kt, wbuf, Nc, (const float2*)coef2_dev, dwbuf, out_dev, pool_dev, N, H, W, sms, st); stage("fused"); }
  }
  t.stop();
  int rc2 = finish(st);
  cudaFree(wbuf); cudaFree(dwbuf); cudaFree(scratch);
  return rc ? LCM_ERR_INVALID : rc2;
}

}  // extern "C"
