// The two dense 3x3 convolutions at the edges of the UNet on the bf16 tensor-core plan.  Both are tiny in FLOPs but move
// full-resolution tensors, so the goal is simply to stay out of the way of the HBM stream.  init_conv is an implicit GEMM
// on mma.sync (legacy HMMA path: ~0.15 ms of tensor time for its 17 GFLOP, measured), final_conv packed-fp16 CUDA-core
// math (its N = 3 wastes 5/8 of an m16n8k16 tile: the mma.sync version measured slower, 0.37 vs 0.32 ms):
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
// init_conv as an implicit GEMM on mma.sync (m16n8k16, fp16 operands, fp32 accumulation): K = 9 * Cin <= 80, N = CO.
// The packed-HFMA2 version was instruction-bound (ncu: 57 % issue utilisation at 3 blocks per SM, ~2000 thread
// instructions per pixel); here a warp spends 8 LDS.32 + 4 cvt + CO/8 HMMA per 16 pixels and K step.  tcgen05 is the
// wrong tool for a 54-deep contraction over fp32 NCHW inputs: the operand would have to be im2col'ed through shared
// memory in the UMMA layout first.
//   block = 128 threads = 32 x 8 pixel tile (warp = 2 rows = four 16-pixel m-tiles), walks 64 rows; the fp32 NCHW halo
//   tile is staged with 4-byte cp.async (double-buffered, zero fill outside the image); A fragments are gathered from it
//   with per-thread precomputed (channel, tap) offsets; B fragments (weights) live in registers for the whole block.
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int CO>
__global__ void __launch_bounds__(128) init_conv_mma_kernel(const float* __restrict__ xa, int ca, long long sa,
                                                            const float* __restrict__ xb, int cb, long long sb,
                                                            const float* __restrict__ w, const float* __restrict__ bias,
                                                            bf16* __restrict__ out, double* __restrict__ stats, int H, int W,
                                                            int CoT, int nblk) {
  pdl_wait();
  pdl_trigger();
  // CoT = total output channels; blockIdx.z = image * nblk + channel block (CO channels each: 48 = 3 x 16, 64 = 2 x 32)
  // smem row: pixel x0 + j sits at column j + 4 (the 32 interior floats are 16-byte aligned: one 16-byte cp.async per four
  // pixels), left / right halo at columns 3 / 36
  constexpr int TW = 32, TH = 8, ROWS = 64, SW_ = TW + 8;
  constexpr int CH_PITCH = (TH + 2) * SW_;                  // floats per input channel plane
  constexpr int NT = CO / 8, KS = 5, PPITCH = CO * 2 + 16;  // n-tiles, max K steps, output patch row pitch (bytes)
  __shared__ __align__(16) float in_s[2][8 * CH_PITCH];     // double-buffered fp32 halo tile, filled with cp.async
  __shared__ __align__(16) uint8_t patch[4][16 * PPITCH];   // per warp: 16 pixels x CO bf16, transposed for 16-byte stores
  __shared__ float red[4][2 * CO];
  const int Cin = ca + cb, K = 9 * Cin, ksteps = (K + 15) / 16;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int n = blockIdx.z / nblk, co0 = (blockIdx.z % nblk) * CO, x0 = blockIdx.x * TW, yb = blockIdx.y * ROWS;
  const int yend = min(yb + ROWS, H);
  const bool wide = (W % 32 == 0) && ((sa | sb) % 4 == 0) &&   // full-width tiles, 16-byte aligned rows
                    ((reinterpret_cast<uintptr_t>(xa) | reinterpret_cast<uintptr_t>(xb)) % 16 == 0);
  auto stage = [&](int y0, int buf) {
    if (wide) {
      // 10 copies per tile line: 8 x 16 bytes (interior) + the two halo pixels.  The element-wise version below issued 34
      // 4-byte copies per line and its address arithmetic was 45 % of the kernel's instructions (ncu source view).
      for (int u = tid; u < Cin * (TH + 2) * 10; u += 128) {
        const int line = u / 10, part = u - line * 10;
        const int ci = line / (TH + 2), r = line - ci * (TH + 2);
        const int gy = y0 + r - 1;
        const bool yok = gy >= 0 && gy < H;
        const float* row = (ci < ca ? xa + n * sa + ((long long)ci * H + (yok ? gy : 0)) * W
                                    : xb + n * sb + ((long long)(ci - ca) * H + (yok ? gy : 0)) * W);
        float* drow = &in_s[buf][ci * CH_PITCH + r * SW_];
        if (part < 8) {
          const uint32_t dst = (uint32_t)__cvta_generic_to_shared(drow + 4 + part * 4);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(row + x0 + part * 4), "r"(yok ? 16 : 0) : "memory");
        } else {
          const int gx = part == 8 ? x0 - 1 : x0 + TW;
          const bool ok = yok && gx >= 0 && gx < W;
          const uint32_t dst = (uint32_t)__cvta_generic_to_shared(drow + (part == 8 ? 3 : TW + 4));
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(row + (ok ? gx : 0)), "r"(ok ? 4 : 0) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      return;
    }
    for (int line = tid >> 5; line < Cin * (TH + 2); line += 4) {
      const int ci = line / (TH + 2), r = line - ci * (TH + 2);
      const int gy = y0 + r - 1;
      const bool yok = gy >= 0 && gy < H;
      const float* src = ci < ca ? xa + n * sa + ((long long)ci * H + (yok ? gy : 0)) * W
                                 : xb + n * sb + ((long long)(ci - ca) * H + (yok ? gy : 0)) * W;
      for (int c = tid & 31; c < TW + 2; c += 32) {
        const int gx = x0 + c - 1;
        const bool ok = yok && gx >= 0 && gx < W;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&in_s[buf][ci * CH_PITCH + r * SW_ + c + 3]);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src + (ok ? gx : 0)), "r"(ok ? 4 : 0) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage(yb, 0);
  // k = tap * Cin + ci (the packed weight order).  This thread's fragment columns: k0 = 16 ks + 2t + {0, 1, 8, 9}
  int koff[KS][4];
  uint32_t bfr[KS][NT][2];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = ks * 16 + 2 * t + (q & 1) + (q >> 1) * 8;
      const int tap = k / Cin, ci = k - tap * Cin;
      koff[ks][q] = k < K ? ci * CH_PITCH + (tap / 3) * SW_ + tap % 3 + 3 : 3;
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int col = co0 + nt * 8 + g;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int k = ks * 16 + 2 * t + hh * 8;
        const float w0 = k < K ? w[(size_t)k * CoT + col] : 0.f, w1 = k + 1 < K ? w[(size_t)(k + 1) * CoT + col] : 0.f;
        const __half2 hw = __floats2half2_rn(w0, w1);
        bfr[ks][nt][hh] = u32(hw);
      }
    }
  }
  float bia[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) { bia[nt][0] = bias[co0 + nt * 8 + 2 * t]; bia[nt][1] = bias[co0 + nt * 8 + 2 * t + 1]; }
  float s[NT][2], q2[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) { s[nt][0] = s[nt][1] = q2[nt][0] = q2[nt][1] = 0.f; }

  int buf = 0;
  for (int y0 = yb; y0 < yend; y0 += TH, buf ^= 1) {
    if (y0 + TH < yend) {
      stage(y0 + TH, buf ^ 1);   // that buffer was last read two tiles ago (barrier at the end of the previous tile)
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const float* tile = in_s[buf];
#pragma unroll 1
    for (int mt = 0; mt < 4; ++mt) {
      const int row = 2 * warp + (mt >> 1), xh = (mt & 1) * 16;
      const int base = row * SW_ + xh + g;
      float c[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) { c[nt][0] = c[nt][2] = bia[nt][0]; c[nt][1] = c[nt][3] = bia[nt][1]; }
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        if (ks < ksteps) {
          uint32_t a[4];
          a[0] = u32(__floats2half2_rn(tile[base + koff[ks][0]], tile[base + koff[ks][1]]));
          a[1] = u32(__floats2half2_rn(tile[base + 8 + koff[ks][0]], tile[base + 8 + koff[ks][1]]));
          a[2] = u32(__floats2half2_rn(tile[base + koff[ks][2]], tile[base + koff[ks][3]]));
          a[3] = u32(__floats2half2_rn(tile[base + 8 + koff[ks][2]], tile[base + 8 + koff[ks][3]]));
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) mma_16816(c[nt], a, bfr[ks][nt][0], bfr[ks][nt][1]);
        }
      }
      // ---- epilogue: bf16, statistics of the stored values, transpose through the warp's patch, 16-byte stores ----
      const int gy = y0 + row;
      const bool v0 = gy < H && x0 + xh + g < W, v1 = gy < H && x0 + xh + g + 8 < W;
      uint8_t* pw = patch[warp];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const uint32_t p0 = pack_bf16(c[nt][0], c[nt][1]), p1 = pack_bf16(c[nt][2], c[nt][3]);
        *reinterpret_cast<uint32_t*>(pw + g * PPITCH + nt * 16 + t * 4) = p0;
        *reinterpret_cast<uint32_t*>(pw + (g + 8) * PPITCH + nt * 16 + t * 4) = p1;
        if (v0) {
          const float r0 = bf16lo(p0), r1 = bf16hi(p0);
          s[nt][0] += r0; q2[nt][0] = fmaf(r0, r0, q2[nt][0]); s[nt][1] += r1; q2[nt][1] = fmaf(r1, r1, q2[nt][1]);
        }
        if (v1) {
          const float r0 = bf16lo(p1), r1 = bf16hi(p1);
          s[nt][0] += r0; q2[nt][0] = fmaf(r0, r0, q2[nt][0]); s[nt][1] += r1; q2[nt][1] = fmaf(r1, r1, q2[nt][1]);
        }
      }
      __syncwarp();
      constexpr int UPP = CO / 8;   // 16-byte units per pixel
#pragma unroll
      for (int i = lane; i < 16 * UPP; i += 32) {
        const int px = i / UPP, unit = i % UPP;
        if (gy < H && x0 + xh + px < W) {
          const uint4 v = *reinterpret_cast<const uint4*>(pw + px * PPITCH + unit * 16);
          *reinterpret_cast<uint4*>(out + (((size_t)n * H + gy) * W + x0 + xh + px) * CoT + co0 + unit * 8) = v;
        }
      }
      __syncwarp();
    }
    __syncthreads();   // everyone is done reading in_s[buf] before the next-but-one stage overwrites it
  }
  // ---- statistics: reduce over the 8 row lanes (g), fixed-order sum over the 4 warps, one fp64 atomic per channel ----
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) { s[nt][j] += __shfl_xor_sync(0xffffffffu, s[nt][j], o); q2[nt][j] += __shfl_xor_sync(0xffffffffu, q2[nt][j], o); }
      if (g == 0) { red[warp][nt * 8 + 2 * t + j] = s[nt][j]; red[warp][CO + nt * 8 + 2 * t + j] = q2[nt][j]; }
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
  pdl_wait();
  pdl_trigger();
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
  if (128 % NCH == 0) {
    // a thread owns ONE 8-channel unit column (128 % NCH == 0): its GroupNorm coefficients are loaded once, and the halo tile is
    // staged five vectors at a time, all loads issued before the first use (the one-vector loop with its four coefficient loads
    // per vector kept 10 KB in flight per SM: 1.4 TB/s)
    const int cv = tid % NCH;
    float2 ab[8];
    {
      const float4* c4 = reinterpret_cast<const float4*>(coef + (size_t)n * CI + cv * 8);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float4 q = c4[e];
        ab[2 * e] = make_float2(q.x, q.y);
        ab[2 * e + 1] = make_float2(q.z, q.w);
      }
    }
    constexpr int TOTAL = PH * PW * NCH, UB = 5;
    for (int base = tid; base < TOTAL; base += 128 * UB) {
      uint4 raw[UB];
      bool ok[UB];
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int i = base + u * 128;
        const int px = i / NCH, r = px / PW, c = px - r * PW;
        const int gy = y0 + r - 1, gx = x0 + c - 1;
        ok[u] = i < TOTAL && gy >= 0 && gy < H && gx >= 0 && gx < W;
        if (ok[u]) raw[u] = *reinterpret_cast<const uint4*>(img + ((size_t)gy * W + gx) * CI + cv * 8);
      }
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int i = base + u * 128;
        if (i >= TOTAL) break;
        const int px = i / NCH, r = px / PW, c = px - r * PW;
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (ok[u]) {
          const uint32_t uu[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            // identical arithmetic to the generic loop below
            const __half2 h = __floats2half2_rn(0.5f * fmaf(ab[2 * e].x, bf16lo(uu[e]), ab[2 * e].y),
                                                0.5f * fmaf(ab[2 * e + 1].x, bf16hi(uu[e]), ab[2 * e + 1].y));
            uint32_t t;
            asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(u32(h)));
            pk[e] = u32(__hfma2(h, h2(t), h));
          }
          o = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
        *reinterpret_cast<uint4*>(tile + ((size_t)(r * NCH + cv) * PW + c) * 8) = o;
      }
    }
  } else
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
#pragma unroll
    for (int co = 0; co < 3; ++co) {   // constant trip count: a runtime bound would index acc[] dynamically and put it in local memory
      if (co >= Co) break;
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
  if (CO == 32) launch_pdl(init_conv_mma_kernel<32>, grid, dim3(128), 0, st, xa, ca, sa, xb, cb, sb, w, bias, (bf16*)out, stats, H, W, Co, nblk);
  else launch_pdl(init_conv_mma_kernel<16>, grid, dim3(128), 0, st, xa, ca, sa, xb, cb, sb, w, bias, (bf16*)out, stats, H, W, Co, nblk);
  return true;
}

bool launch_final_conv_h2(const void* in, const float2* coef, const float* w, const float* bias, float* eps,
                          const FinalStep& step, int N, int H, int W, int Ci, int Co, cudaStream_t st) {
  if (Co > 3 || (Ci != 16 && Ci != 32 && Ci != 48 && Ci != 64)) return false;
  dim3 grid((W + 63) / 64, (H + 7) / 8, N);
  const size_t smem = ((size_t)10 * 66 * Ci + 9 * Ci * 3) * sizeof(__half);
  switch (Ci) {
    case 16: launch_pdl(final_conv_h2_kernel<16>, grid, dim3(128), smem, st, (const bf16*)in, coef, w, bias, eps, step, H, W, Co); break;
    case 32: launch_pdl(final_conv_h2_kernel<32>, grid, dim3(128), smem, st, (const bf16*)in, coef, w, bias, eps, step, H, W, Co); break;
    case 48: {
      if (ensure_dyn_smem_fn(final_conv_h2_kernel<48>, smem)) return false;
      launch_pdl(final_conv_h2_kernel<48>, grid, dim3(128), smem, st, (const bf16*)in, coef, w, bias, eps, step, H, W, Co); break;
    }
    default: {
      if (ensure_dyn_smem_fn(final_conv_h2_kernel<64>, smem)) return false;
      launch_pdl(final_conv_h2_kernel<64>, grid, dim3(128), smem, st, (const bf16*)in, coef, w, bias, eps, step, H, W, Co); break;
    }
  }
  return true;
}

}  // namespace lcm
