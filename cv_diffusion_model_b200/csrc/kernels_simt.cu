// CUDA-core kernels of the hot path, templated on the activation storage type.
//   * gemm_simt / conv3x3_simt: fp32-mode GEMMs and the independent cross-check of the tcgen05 kernels
//   * dwconv (generic version), init conv, final conv + fused LCM step, linear attention, GN+residual
#include "kernels.h"

namespace lcm {

// =================================================================================================
// Channel statistics helper: a block whose rows all belong to image `img` adds its per-column
// partial sums (already reduced into shared memory) to the global double accumulators.
__device__ __forceinline__ void flush_stats(double* stats, int img, int Nc, int col0, int ncols, const float* s_sum,
                                            const float* s_sq) {
  for (int c = threadIdx.x; c < ncols; c += blockDim.x) {
    if (col0 + c < Nc) {
      double* p = stats + ((size_t)img * Nc + col0 + c) * 2;
      atomicAdd(p, (double)s_sum[c]);
      atomicAdd(p + 1, (double)s_sq[c]);
    }
  }
}

// =================================================================================================
// A-operand loaders for the CUDA-core implicit GEMM.  load4 returns 4 consecutive k of row m.
template <typename T>
struct Loader1x1 {
  GemmSeg seg[LCM_MAX_SEGS];
  int nseg;
  int P;
  __device__ __forceinline__ void load4(long long m, int k, float (&v)[4]) const {
    int s = 0, koff = 0;
    while (s + 1 < nseg && k >= koff + seg[s].K) { koff += seg[s].K; ++s; }
    const GemmSeg& g = seg[s];
    const int kk = k - koff;
    const T* p = reinterpret_cast<const T*>(g.A) + m * g.ld + kk;
    const int img = (int)(m / P);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float x = to_f<T>(p[j]);
      if (g.mode != XF_NONE) x = xform(x, g.coef[(size_t)img * g.coef_ld + g.coef_off + kk + j], g.mode);
      v[j] = x;
    }
  }
};

template <typename T>
struct LoaderConv3 {
  const T* in;
  int Hin, Win, Hout, Wout, Ci, mode;
  __device__ __forceinline__ void load4(long long m, int k, float (&v)[4]) const {
    const int tap = k / Ci, ci = k - tap * Ci;
    const int ky = tap / 3, kx = tap - ky * 3;
    const int x = (int)(m % Wout);
    const long long q = m / Wout;
    const int y = (int)(q % Hout);
    const long long n = q / Hout;
    const T* base = in + n * (long long)Hin * Win * Ci + ci;
    if (mode == CONV_UP2) {
      const int uy = y + ky - 1, ux = x + kx - 1;   // coordinates in the (virtual) upsampled image
      if (uy < 0 || uy >= Hout || ux < 0 || ux >= Wout) { v[0] = v[1] = v[2] = v[3] = 0.f; return; }
      // F.interpolate(scale_factor=2, bilinear, align_corners=False): src = max(dst/2 - 0.25, 0)
      const float sy = fmaxf(uy * 0.5f - 0.25f, 0.f), sx = fmaxf(ux * 0.5f - 0.25f, 0.f);
      const int y0 = (int)sy, x0 = (int)sx;
      const int y1 = min(y0 + 1, Hin - 1), x1 = min(x0 + 1, Win - 1);
      const float ly = sy - y0, lx = sx - x0;
      const T* p00 = base + ((long long)y0 * Win + x0) * Ci;
      const T* p01 = base + ((long long)y0 * Win + x1) * Ci;
      const T* p10 = base + ((long long)y1 * Win + x0) * Ci;
      const T* p11 = base + ((long long)y1 * Win + x1) * Ci;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        v[j] = (1.f - ly) * ((1.f - lx) * to_f<T>(p00[j]) + lx * to_f<T>(p01[j])) +
               ly * ((1.f - lx) * to_f<T>(p10[j]) + lx * to_f<T>(p11[j]));
      return;
    }
    const int st = (mode == CONV_S2) ? 2 : 1;
    const int iy = y * st + ky - 1, ix = x * st + kx - 1;
    if (iy < 0 || iy >= Hin || ix < 0 || ix >= Win) { v[0] = v[1] = v[2] = v[3] = 0.f; return; }
    const T* p = base + ((long long)iy * Win + ix) * Ci;
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = to_f<T>(p[j]);
  }
};

// Transposed conv (input gradient) as an implicit GEMM over dY: row m = input pixel (n, yi, xi), k = tap' * Co + co with
// FLIPPED taps (weights packed by PACK_CONV3_T): yy = yi + ky' - 1 is the position in stride-1 output coordinates; for the
// stride-2 forward only even yy < 2 Hout hit an output pixel (yo = yy / 2).
template <typename T>
struct LoaderConv3T {
  const T* dy;
  int Hin, Win, Hout, Wout, Co, mode;
  __device__ __forceinline__ void load4(long long m, int k, float (&v)[4]) const {
    const int tap = k / Co, co = k - tap * Co;
    const int ky = tap / 3, kx = tap - ky * 3;
    const int x = (int)(m % Win);
    const long long q = m / Win;
    const int y = (int)(q % Hin);
    const long long n = q / Hin;
    int yy = y + ky - 1, xx = x + kx - 1;
    bool ok = yy >= 0 && xx >= 0;
    if (mode == CONV_S2) { ok = ok && !(yy & 1) && !(xx & 1); yy >>= 1; xx >>= 1; }
    if (!ok || yy >= Hout || xx >= Wout) { v[0] = v[1] = v[2] = v[3] = 0.f; return; }
    const T* p = dy + ((n * Hout + yy) * (long long)Wout + xx) * Co + co;
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = to_f<T>(p[j]);
  }
};

// 64x64 output tile, BK = 16, 256 threads, 4x4 outputs per thread.  W row-major [Nc][Ktot].
template <typename T, typename Loader>
__global__ void __launch_bounds__(256) gemm_simt_kernel(Loader ld, const T* __restrict__ W, T* __restrict__ out,
                                                        const float* __restrict__ bias, double* __restrict__ stats,
                                                        long long M, int P, int Nc, int Ktot) {
  __shared__ float As[16][68];
  __shared__ float Bs[16][68];
  __shared__ float s_sum[64], s_sq[64];
  __shared__ float s_ps[16][64], s_pq[16][64];   // per-row-group partial column sums (fixed-order fold)
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.x * 64;
  const int n0 = blockIdx.y * 64;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < Ktot; k0 += 16) {
    float a4[4] = {0.f, 0.f, 0.f, 0.f}, b4[4] = {0.f, 0.f, 0.f, 0.f};
    if (m0 + lrow < M) ld.load4(m0 + lrow, k0 + lk, a4);
    if (n0 + lrow < Nc) {
      const T* wp = W + (size_t)(n0 + lrow) * Ktot + k0 + lk;
#pragma unroll
      for (int j = 0; j < 4; ++j) b4[j] = to_f<T>(wp[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { As[lk + j][lrow] = a4[j]; Bs[lk + j][lrow] = b4[j]; }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // epilogue: bias, store, channel statistics of the stored values
  const long long last = (m0 + 63 < M ? m0 + 63 : M - 1);
  const bool one_image = (m0 / P) == (last / P);
  float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= Nc) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      T o = from_f<T>(v);
      out[m * Nc + n] = o;
      v = to_f<T>(o);
      if (stats) {
        if (one_image) { cs[j] += v; cq[j] += v * v; }
        else {
          double* p = stats + ((size_t)(m / P) * Nc + n) * 2;
          atomicAdd(p, (double)v);
          atomicAdd(p + 1, (double)v * v);
        }
      }
    }
  }
  if (stats && one_image) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { s_ps[ty][tx * 4 + j] = cs[j]; s_pq[ty][tx * 4 + j] = cq[j]; }
    __syncthreads();
    if (tid < 64) {
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int g = 0; g < 16; ++g) { a += s_ps[g][tid]; b += s_pq[g][tid]; }
      s_sum[tid] = a; s_sq[tid] = b;
    }
    __syncthreads();
    flush_stats(stats, (int)(m0 / P), Nc, n0, 64, s_sum, s_sq);
  }
}

void launch_gemm_simt(const GemmParams& p, int bf16act, cudaStream_t st) {
  dim3 grid((unsigned)((p.M + 63) / 64), (unsigned)((p.Nc + 63) / 64));
  if (bf16act) {
    Loader1x1<bf16> ld; for (int i = 0; i < LCM_MAX_SEGS; ++i) ld.seg[i] = p.seg[i]; ld.nseg = p.nseg; ld.P = p.P;
    gemm_simt_kernel<bf16, Loader1x1<bf16>><<<grid, 256, 0, st>>>(ld, (const bf16*)p.W, (bf16*)p.out, nullptr, p.stats,
                                                                p.M, p.P, p.Nc, p.Ktot);
  } else {
    Loader1x1<float> ld; for (int i = 0; i < LCM_MAX_SEGS; ++i) ld.seg[i] = p.seg[i]; ld.nseg = p.nseg; ld.P = p.P;
    gemm_simt_kernel<float, Loader1x1<float>><<<grid, 256, 0, st>>>(ld, (const float*)p.W, (float*)p.out, nullptr,
                                                                  p.stats, p.M, p.P, p.Nc, p.Ktot);
  }
}

void launch_conv3x3_simt(const void* in, const void* Wt, const float* bias, void* out, double* stats, int N, int Hin,
                         int Win, int Ci, int Co, int mode, int bf16act, cudaStream_t st) {
  const int Hout = mode == CONV_S2 ? Hin / 2 : (mode == CONV_UP2 ? Hin * 2 : Hin);
  const int Wout = mode == CONV_S2 ? Win / 2 : (mode == CONV_UP2 ? Win * 2 : Win);
  const long long M = (long long)N * Hout * Wout;
  dim3 grid((unsigned)((M + 63) / 64), (unsigned)((Co + 63) / 64));
  if (bf16act) {
    LoaderConv3<bf16> ld{(const bf16*)in, Hin, Win, Hout, Wout, Ci, mode};
    gemm_simt_kernel<bf16, LoaderConv3<bf16>><<<grid, 256, 0, st>>>(ld, (const bf16*)Wt, (bf16*)out, bias, stats, M,
                                                                  Hout * Wout, Co, 9 * Ci);
  } else {
    LoaderConv3<float> ld{(const float*)in, Hin, Win, Hout, Wout, Ci, mode};
    gemm_simt_kernel<float, LoaderConv3<float>><<<grid, 256, 0, st>>>(ld, (const float*)Wt, (float*)out, bias, stats,
                                                                    M, Hout * Wout, Co, 9 * Ci);
  }
}

void launch_conv3x3_dgrad_simt(const void* dY, const void* Wt, void* dX, int N, int Hin, int Win, int Ci, int Co, int mode,
                               int bf16act, cudaStream_t st) {
  const int Hout = mode == CONV_S2 ? Hin / 2 : Hin, Wout = mode == CONV_S2 ? Win / 2 : Win;
  const long long M = (long long)N * Hin * Win;
  dim3 grid((unsigned)((M + 63) / 64), (unsigned)((Ci + 63) / 64));
  if (bf16act) {
    LoaderConv3T<bf16> ld{(const bf16*)dY, Hin, Win, Hout, Wout, Co, mode};
    gemm_simt_kernel<bf16, LoaderConv3T<bf16>><<<grid, 256, 0, st>>>(ld, (const bf16*)Wt, (bf16*)dX, nullptr, nullptr, M,
                                                                   Hin * Win, Ci, 9 * Co);
  } else {
    LoaderConv3T<float> ld{(const float*)dY, Hin, Win, Hout, Wout, Co, mode};
    gemm_simt_kernel<float, LoaderConv3T<float>><<<grid, 256, 0, st>>>(ld, (const float*)Wt, (float*)dX, nullptr, nullptr, M,
                                                                     Hin * Win, Ci, 9 * Co);
  }
}

// =================================================================================================
// Depthwise 3x3 (efficient_unet.py:177-180,220), generic version: 16x16 pixel tile x 32 channels per
// block, transformed halo tile staged once in shared memory (fp32), 4 pixels x 8 channels per thread.
// Prologue: act(a*x+b) per (image, channel) = GroupNorm2 + FiLM + ReLU6 (:212-219).  Zero padding is
// applied AFTER the activation (the conv pads its own input).  Epilogue: SE pooled sums (:97).
template <typename T>
__global__ void __launch_bounds__(256) dwconv_kernel(const T* __restrict__ in, const float2* __restrict__ coef,
                                                     const float* __restrict__ w, T* __restrict__ out,
                                                     double* __restrict__ pool, int H, int W, int C, int tilesX) {
  constexpr int TS = 16, HS = TS + 2, CB = 32;
  __shared__ __align__(16) float tile[HS * HS * CB];
  __shared__ float s_part[8][CB];   // per-warp partial pooled sums (fixed-order reduction: reproducible)
  const int tid = threadIdx.x;
  const int n = blockIdx.z, c0 = blockIdx.y * CB;
  const int ty0 = (blockIdx.x / tilesX) * TS, tx0 = (blockIdx.x % tilesX) * TS;
  const int cg = tid & 3;  // this thread's 8-channel group (fixed for loads and compute: 256 % 4 == 0)

  float2 ab[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) ab[j] = coef[(size_t)n * C + c0 + cg * 8 + j];

  const T* img = in + (size_t)n * H * W * C;
  for (int i = tid; i < HS * HS * 4; i += 256) {
    const int px = i >> 2;
    const int yy = px / HS, xx = px - yy * HS;
    const int gy = ty0 + yy - 1, gx = tx0 + xx - 1;
    float v[8];
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
      Vec8<T>::load(img + ((size_t)gy * W + gx) * C + c0 + cg * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fminf(fmaxf(fmaf(ab[j].x, v[j], ab[j].y), 0.f), 6.f);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
    }
    float4* d = reinterpret_cast<float4*>(tile + px * CB + cg * 8);
    d[0] = make_float4(v[0], v[1], v[2], v[3]);
    d[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
  float wt[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) wt[t][j] = w[(size_t)t * C + c0 + cg * 8 + j];
  __syncthreads();

  const int strip = tid >> 2;           // 64 strips of 4 pixels
  const int row = strip >> 2, xs = (strip & 3) * 4;
  float psum[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) psum[j] = 0.f;
  const int gy = ty0 + row;
#pragma unroll
  for (int px = 0; px < 4; ++px) {
    const int gx = tx0 + xs + px;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float4* s = reinterpret_cast<const float4*>(tile + ((row + ky) * HS + xs + px + kx) * CB + cg * 8);
        const float4 lo = s[0], hi = s[1];
        const float* wv = wt[ky * 3 + kx];
        acc[0] = fmaf(lo.x, wv[0], acc[0]); acc[1] = fmaf(lo.y, wv[1], acc[1]);
        acc[2] = fmaf(lo.z, wv[2], acc[2]); acc[3] = fmaf(lo.w, wv[3], acc[3]);
        acc[4] = fmaf(hi.x, wv[4], acc[4]); acc[5] = fmaf(hi.y, wv[5], acc[5]);
        acc[6] = fmaf(hi.z, wv[6], acc[6]); acc[7] = fmaf(hi.w, wv[7], acc[7]);
      }
    if (gy < H && gx < W) {
      Vec8<T>::store(out + ((size_t)n * H * W + (size_t)gy * W + gx) * C + c0 + cg * 8, acc);
#pragma unroll
      for (int j = 0; j < 8; ++j) psum[j] += acc[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {  // lanes l, l^4, l^8, l^16 share the channel group
    float v = psum[j];
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    if ((tid & 31) < 4) s_part[tid >> 5][cg * 8 + j] = v;
  }
  __syncthreads();
  if (tid < CB) {
    float s = 0.f;
#pragma unroll
    for (int wi = 0; wi < 8; ++wi) s += s_part[wi][tid];
    atomicAdd(&pool[(size_t)n * C + c0 + tid], (double)s);
  }
}

void launch_dwconv_fast(const void* in, const float2* coef, const float* w, void* out, double* pool, int N, int H, int W,
                        int C, cudaStream_t st);   // dwconv_fast.cu (bf16 only)

void launch_dwconv(const void* in, const float2* coef, const float* w, void* out, double* pool, int N, int H, int W,
                   int C, int bf16act, int fast, cudaStream_t st) {
  if (fast && bf16act) { launch_dwconv_fast(in, coef, w, out, pool, N, H, W, C, st); return; }
  const int tilesX = (W + 15) / 16, tilesY = (H + 15) / 16;
  dim3 grid(tilesX * tilesY, C / 32, N);
  if (bf16act) dwconv_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)in, coef, w, (bf16*)out, pool, H, W, C, tilesX);
  else dwconv_kernel<float><<<grid, 256, 0, st>>>((const float*)in, coef, w, (float*)out, pool, H, W, C, tilesX);
}

// =================================================================================================
// init_conv (efficient_unet.py:420,553) fused with the conditioning concat (low_light_diffusion.py:222):
// reads the two fp32 NCHW tensors directly, one thread per output pixel, all Co channels in registers.
template <typename T, int CO_MAX>
__global__ void __launch_bounds__(128) init_conv_kernel(const float* __restrict__ xa, int ca, long long sa,
                                                        const float* __restrict__ xb, int cb, long long sb,
                                                        const float* __restrict__ w, const float* __restrict__ bias,
                                                        T* __restrict__ out, double* __restrict__ stats, int H, int W,
                                                        int Co) {
  extern __shared__ float sw[];  // [9*Cin][Co], then per-warp partials [4][2*Co], then s_sum[Co], s_sq[Co]
  const int Cin = ca + cb, K = 9 * Cin;
  float* s_part = sw + K * Co;
  float* s_sum = s_part + 4 * 2 * Co;
  float* s_sq = s_sum + Co;
  for (int i = threadIdx.x; i < K * Co; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int n = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = p < H * W;
  const int y = valid ? p / W : 0, x = valid ? p - (p / W) * W : 0;
  float acc[CO_MAX];
#pragma unroll
  for (int c = 0; c < CO_MAX; ++c) acc[c] = (c < Co) ? bias[c] : 0.f;
  if (valid) {
    for (int tap = 0; tap < 9; ++tap) {
      const int iy = y + tap / 3 - 1, ix = x + tap % 3 - 1;
      if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
      for (int ci = 0; ci < Cin; ++ci) {
        const float v = ci < ca ? xa[n * sa + ((long long)ci * H + iy) * W + ix]
                                : xb[n * sb + ((long long)(ci - ca) * H + iy) * W + ix];
        const float* wr = sw + (tap * Cin + ci) * Co;
#pragma unroll
        for (int c = 0; c < CO_MAX; ++c)
          if (c < Co) acc[c] = fmaf(v, wr[c], acc[c]);
      }
    }
    T* o = out + ((size_t)n * H * W + p) * Co;
#pragma unroll
    for (int c = 0; c < CO_MAX; c += 8)
      if (c < Co) {
        float v8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v8[j] = acc[c + j];
        Vec8<T>::store(o + c, v8);
      }
  }
#pragma unroll
  for (int c = 0; c < CO_MAX; ++c) {
    if (c < Co) {
      float v = valid ? rt<T>(acc[c]) : 0.f, q = v * v;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
      if ((threadIdx.x & 31) == 0) { s_part[(threadIdx.x >> 5) * 2 * Co + c] = v; s_part[(threadIdx.x >> 5) * 2 * Co + Co + c] = q; }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * Co; i += blockDim.x)   // fixed-order reduction over the 4 warps
    s_sum[i] = (s_part[i] + s_part[2 * Co + i]) + (s_part[4 * Co + i] + s_part[6 * Co + i]);
  __syncthreads();
  flush_stats(stats, n, Co, 0, Co, s_sum, s_sq);
}

void launch_init_conv(const float* xa, int ca, long long sa, const float* xb, int cb, long long sb, const float* w,
                      const float* bias, void* out, double* stats, int N, int H, int W, int Co, int bf16act,
                      cudaStream_t st) {
  dim3 grid((H * W + 127) / 128, N);
  size_t smem = ((size_t)9 * (ca + cb) * Co + 10 * Co) * sizeof(float);
  if (bf16act)
    init_conv_kernel<bf16, 64><<<grid, 128, smem, st>>>(xa, ca, sa, xb, cb, sb, w, bias, (bf16*)out, stats, H, W, Co);
  else
    init_conv_kernel<float, 64><<<grid, 128, smem, st>>>(xa, ca, sa, xb, cb, sb, w, bias, (float*)out, stats, H, W, Co);
}

// =================================================================================================
// final_norm -> SiLU -> final_conv (efficient_unet.py:528-530,600-602) with LCMScheduler.step
// (lcm_scheduler.py:214-242) fused into the epilogue.  16x16 pixel tile, transformed halo in smem.
template <typename T>
__global__ void __launch_bounds__(256) final_conv_kernel(const T* __restrict__ in, const float2* __restrict__ coef,
                                                         const float* __restrict__ w, const float* __restrict__ bias,
                                                         float* __restrict__ eps, FinalStep step, int H, int W, int Ci,
                                                         int Co, int tilesX) {
  constexpr int TS = 16, HS = TS + 2;
  extern __shared__ __align__(16) float fsm[];
  const int PS = Ci + 4;                 // padded pixel stride (floats): conflict-free float4 reads
  float* tile = fsm;                     // [HS*HS][PS]
  float* sw = fsm + HS * HS * PS;        // [9][Ci][4]
  const int tid = threadIdx.x, n = blockIdx.y;
  const int ty0 = (blockIdx.x / tilesX) * TS, tx0 = (blockIdx.x % tilesX) * TS;
  for (int i = tid; i < 9 * Ci * 4; i += 256) {
    const int co = i & 3, k = i >> 2;   // k = tap*Ci + ci ; source w [9*Ci][Co]
    sw[i] = co < Co ? w[(size_t)k * Co + co] : 0.f;
  }
  const int vecs = Ci / 8;
  const T* img = in + (size_t)n * H * W * Ci;
  for (int i = tid; i < HS * HS * vecs; i += 256) {
    const int px = i / vecs, cv = i - px * vecs;
    const int yy = px / HS, xx = px - yy * HS;
    const int gy = ty0 + yy - 1, gx = tx0 + xx - 1;
    float v[8];
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
      Vec8<T>::load(img + ((size_t)gy * W + gx) * Ci + cv * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = xform(v[j], coef[(size_t)n * Ci + cv * 8 + j], XF_AFFINE_SILU);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
    }
    float4* d = reinterpret_cast<float4*>(tile + px * PS + cv * 8);
    d[0] = make_float4(v[0], v[1], v[2], v[3]);
    d[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
  __syncthreads();
  const int ly = tid >> 4, lx = tid & 15;
  const int gy = ty0 + ly, gx = tx0 + lx;
  if (gy >= H || gx >= W) return;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int tap = 0; tap < 9; ++tap) {
    const float* s = tile + ((ly + tap / 3) * HS + lx + tap % 3) * PS;
    const float4* wv = reinterpret_cast<const float4*>(sw + tap * Ci * 4);
    for (int c = 0; c < Ci; c += 4) {
      const float4 a = *reinterpret_cast<const float4*>(s + c);
      const float4 w0 = wv[c], w1 = wv[c + 1], w2 = wv[c + 2], w3 = wv[c + 3];
      acc[0] = fmaf(a.x, w0.x, acc[0]); acc[1] = fmaf(a.x, w0.y, acc[1]); acc[2] = fmaf(a.x, w0.z, acc[2]); acc[3] = fmaf(a.x, w0.w, acc[3]);
      acc[0] = fmaf(a.y, w1.x, acc[0]); acc[1] = fmaf(a.y, w1.y, acc[1]); acc[2] = fmaf(a.y, w1.z, acc[2]); acc[3] = fmaf(a.y, w1.w, acc[3]);
      acc[0] = fmaf(a.z, w2.x, acc[0]); acc[1] = fmaf(a.z, w2.y, acc[1]); acc[2] = fmaf(a.z, w2.z, acc[2]); acc[3] = fmaf(a.z, w2.w, acc[3]);
      acc[0] = fmaf(a.w, w3.x, acc[0]); acc[1] = fmaf(a.w, w3.y, acc[1]); acc[2] = fmaf(a.w, w3.z, acc[2]); acc[3] = fmaf(a.w, w3.w, acc[3]);
    }
  }
  for (int co = 0; co < Co; ++co) {
    const size_t o = (((size_t)n * Co + co) * H + gy) * W + gx;
    const float e = acc[co] + bias[co];
    if (eps) eps[o] = e;
    if (step.enabled) {
      // same fp32 operation order as the reference, no FMA contraction
      const float x = step.latents[o];
      const float x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(step.sb_t, e)), step.sa_t);
      const float prev = step.noise ? __fadd_rn(__fmul_rn(step.sa_p, x0), __fmul_rn(step.sb_p, step.noise[o])) : x0;
      step.latents[o] = prev;
      if (step.trace) step.trace[o] = prev;
      if (step.clamped) step.clamped[o] = fminf(fmaxf(prev, -1.f), 1.f);
    }
  }
}

void launch_final_conv(const void* in, const float2* coef, const float* w, const float* bias, float* eps,
                       const FinalStep& step, int N, int H, int W, int Ci, int Co, int bf16act, cudaStream_t st) {
  const int tilesX = (W + 15) / 16, tilesY = (H + 15) / 16;
  dim3 grid(tilesX * tilesY, N);
  size_t smem = ((size_t)18 * 18 * (Ci + 4) + 9 * Ci * 4) * sizeof(float);
  if (bf16act) {
    cudaFuncSetAttribute(final_conv_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    final_conv_kernel<bf16><<<grid, 256, smem, st>>>((const bf16*)in, coef, w, bias, eps, step, H, W, Ci, Co, tilesX);
  } else {
    cudaFuncSetAttribute(final_conv_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    final_conv_kernel<float><<<grid, 256, smem, st>>>((const float*)in, coef, w, bias, eps, step, H, W, Ci, Co, tilesX);
  }
}

// =================================================================================================
// Linear attention (efficient_unet.py:289-302), d = 32 per head.
// Pass 1: KV[d][e] = sum_p phi(k[p][d]) v[p][e], ksum[d] = sum_p phi(k[p][d]); phi = elu + 1.
// Split over positions: each block reduces a chunk of 64 positions and adds to the fp32 state.
__device__ __forceinline__ float phi(float x) { return x > 0.f ? x + 1.f : expf(x); }  // elu(x)+1

template <typename T>
__global__ void __launch_bounds__(256) attn_kv_kernel(const T* __restrict__ qkv, double* __restrict__ state, int P,
                                                      int heads) {
  pdl_wait();
  pdl_trigger();
  // a block reduces up to 256 positions (4 tiles of 64) in registers before touching the fp64 state: 4x fewer
  // atomics than one tile per block (1024 positions per block left too few blocks in flight: measured slower)
  __shared__ float ks[64][33], vs[64][33];
  const int n = blockIdx.z, h = blockIdx.y;
  const int inner = heads * 32, ld = 3 * inner;
  const int tid = threadIdx.x;
  const int d = tid >> 3, e0 = (tid & 7) * 4;
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, ksum = 0.f;
  const int pend = min(P, (int)(blockIdx.x + 1) * 256);
  for (int p0 = blockIdx.x * 256; p0 < pend; p0 += 64) {
    __syncthreads();
    for (int i = tid; i < 64 * 32; i += 256) {
      const int pp = i >> 5, dd = i & 31;
      const int p = p0 + pp;
      float k = 0.f, v = 0.f;
      if (p < P) {
        const T* row = qkv + ((size_t)n * P + p) * ld;
        k = phi(to_f<T>(row[inner + h * 32 + dd]));
        v = to_f<T>(row[2 * inner + h * 32 + dd]);
      }
      ks[pp][dd] = k;   // zero (not phi(0)=1) beyond P so that padding does not contribute
      vs[pp][dd] = v;
    }
    __syncthreads();
    for (int pp = 0; pp < 64; ++pp) {
      const float k = ks[pp][d];
      ksum += k;
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = fmaf(k, vs[pp][e0 + j], acc[j]);
    }
  }
  double* s = state + (((size_t)n * heads + h) * 32 + d) * 33;
#pragma unroll
  for (int j = 0; j < 4; ++j) atomicAdd(&s[e0 + j], (double)acc[j]);
  if (e0 == 0) atomicAdd(&s[32], (double)ksum);
}

// Pass 2: out[p][e] = sum_d phi(q[p][d]) KV[d][e] / (sum_d phi(q[p][d]) ksum[d] + 1e-6)
template <typename T>
__global__ void __launch_bounds__(256) attn_apply_kernel(const T* __restrict__ qkv, const double* __restrict__ state,
                                                         T* __restrict__ out, int P, int heads) {
  pdl_wait();
  pdl_trigger();
  __shared__ float kv[32][33];
  __shared__ float qs[64][33];
  const int n = blockIdx.z, h = blockIdx.y, p0 = blockIdx.x * 64;
  const int inner = heads * 32, ld = 3 * inner;
  const int tid = threadIdx.x;
  const double* s = state + ((size_t)n * heads + h) * 32 * 33;
  for (int i = tid; i < 32 * 33; i += 256) kv[i / 33][i % 33] = (float)s[i];
  for (int i = tid; i < 64 * 32; i += 256) {
    const int pp = i >> 5, d = i & 31, p = p0 + pp;
    qs[pp][d] = p < P ? phi(to_f<T>(qkv[((size_t)n * P + p) * ld + h * 32 + d])) : 0.f;
  }
  __syncthreads();
  for (int i = tid; i < 64 * 32; i += 256) {
    const int pp = i >> 5, e = i & 31, p = p0 + pp;
    if (p >= P) continue;
    float num = 0.f, den = 0.f;
#pragma unroll
    for (int d = 0; d < 32; ++d) {
      const float q = qs[pp][d];
      num = fmaf(q, kv[d][e], num);
      den = fmaf(q, kv[d][32], den);
    }
    out[((size_t)n * P + p) * inner + h * 32 + e] = from_f<T>(num / (den + 1e-6f));
  }
}

void launch_attn_kv(const void* qkv, double* state, int N, int P, int heads, int bf16act, cudaStream_t st) {
  dim3 grid((P + 255) / 256, heads, N);
  if (bf16act) launch_pdl(attn_kv_kernel<bf16>, grid, dim3(256), 0, st, (const bf16*)qkv, state, P, heads);
  else launch_pdl(attn_kv_kernel<float>, grid, dim3(256), 0, st, (const float*)qkv, state, P, heads);
}
void launch_attn_apply(const void* qkv, const double* state, void* out, int N, int P, int heads, int bf16act,
                       cudaStream_t st) {
  dim3 grid((P + 63) / 64, heads, N);
  if (bf16act) launch_pdl(attn_apply_kernel<bf16>, grid, dim3(256), 0, st, (const bf16*)qkv, state, (bf16*)out, P, heads);
  else launch_pdl(attn_apply_kernel<float>, grid, dim3(256), 0, st, (const float*)qkv, state, (float*)out, P, heads);
}

// =================================================================================================
// StandardAttention (efficient_unet.py:311-357, use_linear_attention=False): softmax(q k^T / sqrt(d)) v per head, d = 32.
// One block = 64 queries of one (image, head); keys / values stream through shared memory in tiles of 64 with the
// online-softmax recurrence (running max and sum per query), fp32 throughout.  A thread owns one query and 8 of the 32
// output channels (4 threads per query).
template <typename T>
__global__ void __launch_bounds__(256) attn_softmax_kernel(const T* __restrict__ qkv, T* __restrict__ out, int P, int heads) {
  pdl_wait();
  pdl_trigger();
  __shared__ float ks[64][33], vs[64][33];
  const int n = blockIdx.z, h = blockIdx.y, p0 = blockIdx.x * 64;
  const int inner = heads * 32, ld = 3 * inner;
  const int tid = threadIdx.x;
  const int qi = tid >> 2, part = tid & 3;       // query, channel quarter
  const int p = p0 + qi;
  float q[32];
#pragma unroll
  for (int d = 0; d < 32; ++d) q[d] = p < P ? to_f<T>(qkv[((size_t)n * P + p) * ld + h * 32 + d]) * 0.17677669529663687f : 0.f;   // 32^-0.5
  float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int k0 = 0; k0 < P; k0 += 64) {
    __syncthreads();
    for (int i = tid; i < 64 * 32; i += 256) {
      const int kk = i >> 5, d = i & 31, kp = k0 + kk;
      float kv = 0.f, vv = 0.f;
      if (kp < P) {
        const T* row = qkv + ((size_t)n * P + kp) * ld;
        kv = to_f<T>(row[inner + h * 32 + d]);
        vv = to_f<T>(row[2 * inner + h * 32 + d]);
      }
      ks[kk][d] = kv;
      vs[kk][d] = vv;
    }
    __syncthreads();
    const int kend = min(64, P - k0);
    // the 4 threads of a query each score 16 of the 64 keys, then exchange the tile maximum and rescale once per tile
    float s[16];
    float tmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int kk = part * 16 + j;
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < 32; ++d) a = fmaf(q[d], ks[kk][d], a);
      s[j] = kk < kend ? a : -INFINITY;
      tmax = fmaxf(tmax, s[j]);
    }
    tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 1));
    tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 2));
    const float mn = fmaxf(m, tmax);
    const float corr = __expf(m - mn);          // exp(-inf) = 0 on the first tile
    float lsum = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) { s[j] = __expf(s[j] - mn); lsum += s[j]; }
    lsum += __shfl_xor_sync(0xffffffffu, lsum, 1);
    lsum += __shfl_xor_sync(0xffffffffu, lsum, 2);
    l = l * corr + lsum;
    m = mn;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= corr;
    // acc[e] += sum_k p[k] v[k][e] for this thread's 8 channels: the probabilities of the other three threads come by shuffle
#pragma unroll
    for (int src = 0; src < 4; ++src) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float pj = __shfl_sync(0xffffffffu, s[j], (tid & 31 & ~3) | src);
        const int kk = src * 16 + j;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(pj, vs[kk][part * 8 + e], acc[e]);
      }
    }
  }
  if (p < P) {
    const float inv = 1.f / l;
#pragma unroll
    for (int e = 0; e < 8; ++e) out[((size_t)n * P + p) * inner + h * 32 + part * 8 + e] = from_f<T>(acc[e] * inv);
  }
}

void launch_attn_softmax(const void* qkv, void* out, int N, int P, int heads, int bf16act, cudaStream_t st) {
  dim3 grid((P + 63) / 64, heads, N);
  if (bf16act) launch_pdl(attn_softmax_kernel<bf16>, grid, dim3(256), 0, st, (const bf16*)qkv, (bf16*)out, P, heads);
  else launch_pdl(attn_softmax_kernel<float>, grid, dim3(256), 0, st, (const float*)qkv, (float*)out, P, heads);
}

// =================================================================================================
// y = a*u + b + x : the GroupNorm after to_out plus the attention residual (efficient_unet.py:266-269,
// 306-308), with channel statistics of y for the next GroupNorm.  Block = 64 pixels of one image.
template <typename T>
__global__ void __launch_bounds__(256) affine_residual_kernel(const T* __restrict__ u, const float2* __restrict__ coef,
                                                              const T* __restrict__ x, T* __restrict__ y,
                                                              double* __restrict__ stats, int P, int C) {
  pdl_wait();
  pdl_trigger();
  const int n = blockIdx.y, p0 = blockIdx.x * 64;
  const int cvecs = C / 8;
  for (int cv = threadIdx.x; cv < cvecs; cv += blockDim.x) {
    float2 ab[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) ab[j] = coef[(size_t)n * C + cv * 8 + j];
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
    for (int p = p0; p < min(p0 + 64, P); ++p) {
      const size_t o = ((size_t)n * P + p) * C + cv * 8;
      float a[8], r[8];
      Vec8<T>::load(u + o, a);
      Vec8<T>::load(x + o, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a[j] = fmaf(ab[j].x, a[j], ab[j].y) + r[j];
        const float v = rt<T>(a[j]);
        s[j] += v; q[j] += v * v;
      }
      Vec8<T>::store(y + o, a);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      double* d = stats + ((size_t)n * C + cv * 8 + j) * 2;
      atomicAdd(d, (double)s[j]);
      atomicAdd(d + 1, (double)q[j]);
    }
  }
}

void launch_affine_residual(const void* u, const float2* coef, const void* x, void* y, double* stats, int N, int P,
                            int C, int bf16act, cudaStream_t st) {
  dim3 grid((P + 63) / 64, N);
  int threads = C / 8 < 32 ? 32 : (C / 8 > 256 ? 256 : ((C / 8 + 31) / 32) * 32);
  if (bf16act) launch_pdl(affine_residual_kernel<bf16>, dim3(grid), dim3(threads), 0, st, (const bf16*)u, coef, (const bf16*)x, (bf16*)y, stats, P, C);
  else launch_pdl(affine_residual_kernel<float>, dim3(grid), dim3(threads), 0, st, (const float*)u, coef, (const float*)x, (float*)y, stats, P, C);
}

// =================================================================================================
// Bilinear x2 upsampling (F.interpolate(scale_factor=2, mode='bilinear', align_corners=False),
// efficient_unet.py:383) materialised in bf16 NHWC.  Used by the tensor-core path: the up-convolution then
// becomes a stride-1 conv whose operand tiles are plain TMA boxes (zero fill = padding); the blend is done once
// per output pixel instead of once per tap.  Memory-bound.
// One thread = one INPUT pixel x 8 channels: its 3x3 neighbourhood (9 loads, mostly L1 hits) yields the 2x2 output
// quad.  With scale 2 and align_corners=False every output is a fixed (1/4, 3/4) blend: rows {y-1, y} for the even
// output row, {y, y+1} for the odd one (indices clamped at the border, which reproduces PyTorch's max(src, 0) / min
// clamping exactly), same along x.  Horizontal blends are shared by the two output rows.
__global__ void __launch_bounds__(256) upsample2x_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int H, int W,
                                                         int C, long long total) {
  pdl_wait();
  pdl_trigger();
  const int cvecs = C >> 3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvecs);
    long long q = i / cvecs;
    const int x = (int)(q % W);
    q /= W;
    const int y = (int)(q % H);
    const long long n = q / H;
    const int ym = max(y - 1, 0), yp = min(y + 1, H - 1), xm = max(x - 1, 0), xp = min(x + 1, W - 1);
    const bf16* b = in + n * (long long)H * W * C + cv * 8;
    float l[3][8], r[3][8];   // per input row: left output column (0.25 x-1 + 0.75 x), right (0.75 x + 0.25 x+1)
    const int ys[3] = {ym, y, yp};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float a[8], c[8], d[8];
      Vec8<bf16>::load(b + ((long long)ys[k] * W + xm) * C, a);
      Vec8<bf16>::load(b + ((long long)ys[k] * W + x) * C, c);
      Vec8<bf16>::load(b + ((long long)ys[k] * W + xp) * C, d);
#pragma unroll
      for (int j = 0; j < 8; ++j) { l[k][j] = 0.25f * a[j] + 0.75f * c[j]; r[k][j] = 0.75f * c[j] + 0.25f * d[j]; }
    }
    bf16* o = out + ((n * 2 * H + 2 * y) * 2 * W + 2 * x) * (long long)C + cv * 8;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.25f * l[0][j] + 0.75f * l[1][j];
    Vec8<bf16>::store(o, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.25f * r[0][j] + 0.75f * r[1][j];
    Vec8<bf16>::store(o + C, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.75f * l[1][j] + 0.25f * l[2][j];
    Vec8<bf16>::store(o + (long long)2 * W * C, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.75f * r[1][j] + 0.25f * r[2][j];
    Vec8<bf16>::store(o + (long long)2 * W * C + C, v);
  }
}

void launch_upsample2x(const void* in, void* out, int N, int H, int W, int C, cudaStream_t st) {
  const long long total = (long long)N * H * W * (C / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 64) blocks = 148LL * 64;
  launch_pdl(upsample2x_kernel, dim3((int)blocks), dim3(256), 0, st, (const bf16*)in, (bf16*)out, H, W, C, total);
}

}  // namespace lcm
