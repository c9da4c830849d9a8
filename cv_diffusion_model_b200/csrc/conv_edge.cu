// The two dense 3x3 convolutions at the edges of the UNet on the bf16 tensor-core plan, as CUDA-core kernels built
// around packed fp16 math (HFMA2: two MACs per instruction).  Both are tiny in FLOPs but move full-resolution
// tensors, so the goal is simply to stay out of the way of the HBM stream:
//
//   init_conv  (efficient_unet.py:420,553 fused with the conditioning concat, low_light_diffusion.py:222):
//              fp32 NCHW latents + condition (6 channels) -> bf16 NHWC C0 channels + GroupNorm statistics
//   final_conv (efficient_unet.py:528-530,600-602 fused with LCMScheduler.step, lcm_scheduler.py:214-242):
//              GN + SiLU prologue on bf16 NHWC C0 channels -> 3 fp32 NCHW channels -> x0 / x_prev update
//
// Inputs to the MACs are O(1) (images in [-1,1], latents a few sigma, SiLU of normalised activations), so fp16
// operands lose nothing against the bf16 tensors around them; accumulation is fp16 only over short runs (see each
// kernel) and fp32 beyond.  The fp32 plan and the SIMT cross-check keep the generic kernels in kernels_simt.cu.
#include <cuda_fp16.h>

#include "kernels.h"

namespace lcm {

namespace {

__device__ __forceinline__ __half2 h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t u32(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

// ---------------------------------------------------------------------------------------------------------------
// init_conv: block = 128 threads = 32 x 8 pixel tile, two horizontally adjacent pixels per thread, all CO outputs
// in registers as half2 pairs; the block walks 64 rows so that weights / statistics are set up once per 2048 pixels.
// Per tap and input channel: CO/8 broadcast LDS.128 of weights feed CO HFMA2 (both pixels).  54 terms per output are
// accumulated in fp16: the result is rounded to bf16 (8 bits) anyway.
template <int CO>
__global__ void __launch_bounds__(128) init_conv_h2_kernel(const float* __restrict__ xa, int ca, long long sa,
                                                           const float* __restrict__ xb, int cb, long long sb,
                                                           const float* __restrict__ w, const float* __restrict__ bias,
                                                           bf16* __restrict__ out, double* __restrict__ stats, int H, int W,
                                                           int CoT, int nblk) {
  // CoT = total output channels; blockIdx.z = image * nblk + channel block (CO channels each: 48 = 3 x 16, 64 = 2 x 32)
  constexpr int TW = 32, TH = 8, ROWS = 64, SW_ = TW + 4;   // smem row: 34 used columns, padded to 36
  __shared__ __align__(16) float in_s[2][8][TH + 2][SW_];    // double-buffered fp32 halo tile, filled with cp.async
  __shared__ __align__(16) __half w_s[72][CO];
  __shared__ float red[4][2 * CO];
  const int Cin = ca + cb;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int n = blockIdx.z / nblk, co0 = (blockIdx.z % nblk) * CO, x0 = blockIdx.x * TW, yb = blockIdx.y * ROWS;
  const int yend = min(yb + ROWS, H);
  // stage the halo tile of rows [y0-1, y0+TH] asynchronously: one warp per (channel, row) line, 4-byte cp.async with
  // zero fill outside the image — nothing is held in registers while the previous tile is being computed
  auto stage = [&](int y0, int buf) {
    for (int line = tid >> 5; line < Cin * (TH + 2); line += 4) {
      const int ci = line / (TH + 2), r = line - ci * (TH + 2);
      const int gy = y0 + r - 1;
      const bool yok = gy >= 0 && gy < H;
      const float* src = ci < ca ? xa + n * sa + ((long long)ci * H + (yok ? gy : 0)) * W
                                 : xb + n * sb + ((long long)(ci - ca) * H + (yok ? gy : 0)) * W;
      for (int c = tid & 31; c < TW + 2; c += 32) {
        const int gx = x0 + c - 1;
        const bool ok = yok && gx >= 0 && gx < W;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&in_s[buf][ci][r][c]);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src + (ok ? gx : 0)), "r"(ok ? 4 : 0) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage(yb, 0);
  for (int i = tid; i < 9 * Cin * CO; i += 128) w_s[i / CO][i % CO] = __float2half_rn(w[(size_t)(i / CO) * CoT + co0 + i % CO]);
  __half2 bias2[CO / 2];
#pragma unroll
  for (int j = 0; j < CO / 2; ++j) bias2[j] = __floats2half2_rn(bias[co0 + 2 * j], bias[co0 + 2 * j + 1]);
  float s[CO], q[CO];
#pragma unroll
  for (int c = 0; c < CO; ++c) { s[c] = 0.f; q[c] = 0.f; }

  int buf = 0;
  for (int y0 = yb; y0 < yend; y0 += TH, buf ^= 1) {
    if (y0 + TH < yend) {
      stage(y0 + TH, buf ^ 1);   // that buffer was last read two tiles ago (barrier at the end of the previous tile)
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();   // this tile's data (and, first time round, the weights) are visible to every thread
    __half2 a0[CO / 2], a1[CO / 2];
#pragma unroll
    for (int j = 0; j < CO / 2; ++j) { a0[j] = bias2[j]; a1[j] = bias2[j]; }
    for (int dy = 0; dy < 3; ++dy) {
      for (int ci = 0; ci < Cin; ++ci) {
        const float2 f01 = *reinterpret_cast<const float2*>(&in_s[buf][ci][ty + dy][2 * tx]);
        const float2 f23 = *reinterpret_cast<const float2*>(&in_s[buf][ci][ty + dy][2 * tx + 2]);
        const __half2 c01 = __floats2half2_rn(f01.x, f01.y), c23 = __floats2half2_rn(f23.x, f23.y);
        const __half2 col[4] = {__low2half2(c01), __high2half2(c01), __low2half2(c23), __high2half2(c23)};
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const uint4* wr = reinterpret_cast<const uint4*>(&w_s[(dy * 3 + dx) * Cin + ci][0]);
#pragma unroll
          for (int v = 0; v < CO / 8; ++v) {
            const uint4 w4 = wr[v];
            a0[4 * v + 0] = __hfma2(h2(w4.x), col[dx], a0[4 * v + 0]); a1[4 * v + 0] = __hfma2(h2(w4.x), col[dx + 1], a1[4 * v + 0]);
            a0[4 * v + 1] = __hfma2(h2(w4.y), col[dx], a0[4 * v + 1]); a1[4 * v + 1] = __hfma2(h2(w4.y), col[dx + 1], a1[4 * v + 1]);
            a0[4 * v + 2] = __hfma2(h2(w4.z), col[dx], a0[4 * v + 2]); a1[4 * v + 2] = __hfma2(h2(w4.z), col[dx + 1], a1[4 * v + 2]);
            a0[4 * v + 3] = __hfma2(h2(w4.w), col[dx], a0[4 * v + 3]); a1[4 * v + 3] = __hfma2(h2(w4.w), col[dx + 1], a1[4 * v + 3]);
          }
        }
      }
    }
    const int gy = y0 + ty, gx = x0 + 2 * tx;
    if (gy < H) {
#pragma unroll
      for (int px = 0; px < 2; ++px) {
        if (gx + px < W) {
          bf16* o = out + (((size_t)n * H + gy) * W + gx + px) * CoT + co0;
#pragma unroll
          for (int v = 0; v < CO / 8; ++v) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __half22float2(px ? a1[4 * v + e] : a0[4 * v + e]);
              pk[e] = pack_bf16(f.x, f.y);
              const float r0 = bf16lo(pk[e]), r1 = bf16hi(pk[e]);   // statistics of the stored values
              s[8 * v + 2 * e] += r0; q[8 * v + 2 * e] = fmaf(r0, r0, q[8 * v + 2 * e]);
              s[8 * v + 2 * e + 1] += r1; q[8 * v + 2 * e + 1] = fmaf(r1, r1, q[8 * v + 2 * e + 1]);
            }
            *reinterpret_cast<uint4*>(o + 8 * v) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
      }
    }
    __syncthreads();   // everyone is done reading in_s[buf] before the next-but-one stage overwrites it
  }
  // ---- statistics: warp shuffle tree, fixed-order sum over the 4 warps, one fp64 atomic per (block, channel, moment)
#pragma unroll
  for (int c = 0; c < CO; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s[c] += __shfl_xor_sync(0xffffffffu, s[c], o); q[c] += __shfl_xor_sync(0xffffffffu, q[c], o); }
    if ((tid & 31) == 0) { red[tid >> 5][c] = s[c]; red[tid >> 5][CO + c] = q[c]; }
  }
  __syncthreads();
  if (tid < 2 * CO) {
    const float v = (red[0][tid] + red[1][tid]) + (red[2][tid] + red[3][tid]);
    const int c = tid < CO ? tid : tid - CO;
    atomicAdd(&stats[((size_t)n * CoT + co0 + c) * 2 + (tid < CO ? 0 : 1)], (double)v);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// final_conv: block = 128 threads = 64 x 8 pixel tile, four pixels per thread (x = tx + 16 j, so that the 8 threads of
// a quarter warp read 8 consecutive 16-byte vectors: conflict-free LDS.128).  The GN + SiLU prologue is applied once
// per halo pixel while staging the tile in shared memory as fp16 [row][8-channel chunk][pixel].  Per tile row and
// chunk a thread loads 12 pixel vectors + 9 weight vectors (LDS.128) for 144 HFMA2; the 24 products of a
// (row, chunk, output) are accumulated in fp16, everything beyond that in fp32.
template <int CI>
__global__ void __launch_bounds__(128) final_conv_h2_kernel(const bf16* __restrict__ in, const float2* __restrict__ coef,
                                                            const float* __restrict__ w, const float* __restrict__ bias,
                                                            float* __restrict__ eps, FinalStep step, int H, int W, int Co) {
  constexpr int TW = 64, TH = 8, PW = TW + 2, PH = TH + 2;
  extern __shared__ __align__(16) uint8_t fsm_raw[];
  constexpr int NCH = CI / 8;
  __half* tile = reinterpret_cast<__half*>(fsm_raw);                       // [PH][NCH][PW][8]
  __half* w_s = tile + PH * PW * CI;                                       // [9][CI/8][3 co][8 ch]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int n = blockIdx.z, x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  // weights: source w [tap*CI + ci][Co] fp32 -> fp16 grouped so that one LDS.128 = 8 channels of one (tap, co)
  for (int i = tid; i < 9 * CI * 3; i += 128) {
    const int ch = i % 8, co = (i / 8) % 3, chunk = (i / 24) % (CI / 8), tap = i / (24 * (CI / 8));
    w_s[i] = __float2half_rn(co < Co ? w[(size_t)(tap * CI + chunk * 8 + ch) * Co + co] : 0.f);
  }
  // transformed halo tile: silu(a x + b), zero outside the image (the conv's padding applies after the activation)
  // silu(v) = h tanh(h) + h with h = v / 2: fp32 affine (the 1/2 folded into the coefficients), one packed
  // tanh.approx.f16x2 (MUFU) and one HFMA2 per two channels
  const bf16* img = in + (size_t)n * H * W * CI;
  for (int r = 0; r < PH; ++r) {
    const int gy = y0 + r - 1;
    for (int i = tid; i < PW * NCH; i += 128) {
      const int cv = i % NCH, c = i / NCH;   // NCH is a power of two or 6: cheap
      const int gx = x0 + c - 1;
      uint4 o = make_uint4(0u, 0u, 0u, 0u);
      if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
        const uint4 u = *reinterpret_cast<const uint4*>(img + ((size_t)gy * W + gx) * CI + cv * 8);
        const float4* c4 = reinterpret_cast<const float4*>(coef + (size_t)n * CI + cv * 8);
        const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float4 ab = c4[e];
          const __half2 h = __floats2half2_rn(0.5f * fmaf(ab.x, bf16lo(uu[e]), ab.y), 0.5f * fmaf(ab.z, bf16hi(uu[e]), ab.w));
          uint32_t t;
          asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(u32(h)));
          pk[e] = u32(__hfma2(h, h2(t), h));
        }
        o = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      *reinterpret_cast<uint4*>(tile + ((size_t)(r * NCH + cv) * PW + c) * 8) = o;
    }
  }
  __syncthreads();
  float acc[4][3];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int co = 0; co < 3; ++co) acc[p][co] = 0.f;
  for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
    for (int chunk = 0; chunk < NCH; ++chunk) {
      const __half* trow = tile + ((size_t)((ty + dy) * NCH + chunk) * PW + tx) * 8;
      uint4 a[4][3];
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) a[p][dx] = *reinterpret_cast<const uint4*>(trow + (16 * p + dx) * 8);
      __half2 part[4][3];
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const uint4* wv = reinterpret_cast<const uint4*>(w_s + ((size_t)((dy * 3 + dx) * (CI / 8) + chunk) * 3) * 8);
#pragma unroll
        for (int co = 0; co < 3; ++co) {
          const uint4 w4 = wv[co];
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const uint4 av = a[p][dx];
            __half2 t = dx == 0 ? __hmul2(h2(av.x), h2(w4.x)) : __hfma2(h2(av.x), h2(w4.x), part[p][co]);
            t = __hfma2(h2(av.y), h2(w4.y), t);
            t = __hfma2(h2(av.z), h2(w4.z), t);
            part[p][co] = __hfma2(h2(av.w), h2(w4.w), t);
          }
        }
      }
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int co = 0; co < 3; ++co) {
          const float2 f = __half22float2(part[p][co]);
          acc[p][co] += f.x + f.y;
        }
    }
  }
  const int gy = y0 + ty;
  if (gy >= H) return;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int gx = x0 + tx + 16 * p;
    if (gx >= W) continue;
    for (int co = 0; co < Co; ++co) {
      const size_t o = (((size_t)n * Co + co) * H + gy) * W + gx;
      const float e = acc[p][co] + bias[co];
      if (eps) eps[o] = e;
      if (step.enabled) {
        // same fp32 operation order as the reference, no FMA contraction
        const float x = step.latents[o];
        const float x0v = __fdiv_rn(__fsub_rn(x, __fmul_rn(step.sb_t, e)), step.sa_t);
        const float prev = step.noise ? __fadd_rn(__fmul_rn(step.sa_p, x0v), __fmul_rn(step.sb_p, step.noise[o])) : x0v;
        step.latents[o] = prev;
        if (step.trace) step.trace[o] = prev;
        if (step.clamped) step.clamped[o] = fminf(fmaxf(prev, -1.f), 1.f);
      }
    }
  }
}

}  // namespace

bool launch_init_conv_h2(const float* xa, int ca, long long sa, const float* xb, int cb, long long sb, const float* w,
                         const float* bias, void* out, double* stats, int N, int H, int W, int Co, cudaStream_t st) {
  if (ca + cb > 8 || (Co != 16 && Co != 32 && Co != 48 && Co != 64)) return false;
  const int CO = Co % 32 == 0 ? 32 : 16, nblk = Co / CO;
  dim3 grid((W + 31) / 32, (H + 63) / 64, N * nblk);
  if (CO == 32) init_conv_h2_kernel<32><<<grid, 128, 0, st>>>(xa, ca, sa, xb, cb, sb, w, bias, (bf16*)out, stats, H, W, Co, nblk);
  else init_conv_h2_kernel<16><<<grid, 128, 0, st>>>(xa, ca, sa, xb, cb, sb, w, bias, (bf16*)out, stats, H, W, Co, nblk);
  return true;
}

bool launch_final_conv_h2(const void* in, const float2* coef, const float* w, const float* bias, float* eps,
                          const FinalStep& step, int N, int H, int W, int Ci, int Co, cudaStream_t st) {
  if (Co > 3 || (Ci != 16 && Ci != 32 && Ci != 48 && Ci != 64)) return false;
  dim3 grid((W + 63) / 64, (H + 7) / 8, N);
  const size_t smem = ((size_t)10 * 66 * Ci + 9 * Ci * 3) * sizeof(__half);
  switch (Ci) {
    case 16: final_conv_h2_kernel<16><<<grid, 128, smem, st>>>((const bf16*)in, coef, w, bias, eps, step, H, W, Co); break;
    case 32: final_conv_h2_kernel<32><<<grid, 128, smem, st>>>((const bf16*)in, coef, w, bias, eps, step, H, W, Co); break;
    case 48: {
      static bool done = false;
      if (!done) { cudaFuncSetAttribute(final_conv_h2_kernel<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); done = true; }
      final_conv_h2_kernel<48><<<grid, 128, smem, st>>>((const bf16*)in, coef, w, bias, eps, step, H, W, Co); break;
    }
    default: {
      static bool done = false;
      if (!done) { cudaFuncSetAttribute(final_conv_h2_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); done = true; }
      final_conv_h2_kernel<64><<<grid, 128, smem, st>>>((const bf16*)in, coef, w, bias, eps, step, H, W, Co); break;
    }
  }
  return true;
}

}  // namespace lcm
