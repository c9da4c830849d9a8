// Backward kernels of the data-parallel training step (BASELINE config 5): the reverse of every forward op of
// plan.cu, plus loss, global-norm clip + AdamW + EMA.  Reference semantics: src/training/trainer.py:269-338
// (train_epoch), :86-111 (EMA), :152-156 (AdamW); src/models/low_light_diffusion.py:115-175,250-277;
// the forward ops being differentiated: src/models/efficient_unet.py:203-236 (block), :273-308 (attention),
// :360-384 (down / up), :528-530,600-602 (final norm / SiLU / conv), :60-76,412-417 (time embedding).
//
// Conventions
//   * activations and their gradients are NHWC; the element type of every tensor is a run-time code (DT_F32 / DT_BF16 /
//     DT_F16) because one backward op touches up to three storage types on the tensor-core plan: the residual stream
//     (bf16), a block's hidden tensors (fp16) and gradients (bf16 — fp32's exponent range: d(loss)/d(eps) is
//     O(1/numel) and would flush to zero in fp16).  The fp32 plan uses DT_F32 everywhere.
//   * weight gradients are fp32 in the reference's state_dict layouts, accumulated with atomics into a flat buffer
//     that is zeroed at the start of every backward pass.
//   * per-(image, channel) reductions go to fp64 accumulators (zeroed with the backward scratch region), the same
//     discipline as the forward statistics.
//   * GroupNorm backward never gets its own pass over a full tensor: the producer of du emits
//     T1 = sum_p du, T2 = sum_p du * x per (image, channel); gn_bwd_coef turns them into (A, B, C) with
//     dx = A * du + B * x + C  (plus d gamma, d beta, d FiLM), applied by the next elementwise pass.
#include <cuda_fp16.h>

#include <cmath>
#include <cstdio>

#include "kernels.h"

namespace lcm {

// ---- run-time typed vector access ------------------------------------------------------------------
__device__ __forceinline__ void ld8(const void* base, int dt, size_t i, float (&v)[8]) {
  if (dt == DT_F32) {
    Vec8<float>::load(reinterpret_cast<const float*>(base) + i, v);
  } else if (dt == DT_BF16) {
    Vec8<bf16>::load(reinterpret_cast<const bf16*>(base) + i, v);
  } else {
    const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(base) + i);
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h[j]); v[2 * j] = f.x; v[2 * j + 1] = f.y; }
  }
}
__device__ __forceinline__ void st8(void* base, int dt, size_t i, const float (&v)[8]) {
  if (dt == DT_F32) {
    Vec8<float>::store(reinterpret_cast<float*>(base) + i, v);
  } else if (dt == DT_BF16) {
    Vec8<bf16>::store(reinterpret_cast<bf16*>(base) + i, v);
  } else {
    uint4 u;
    __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(base) + i) = u;
  }
}
// 16-bit storage types: the raw 16-byte vector first (several in flight per thread), converted later
__device__ __forceinline__ uint4 ldraw16(const void* base, size_t i) {
  return *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(base) + i);
}
__device__ __forceinline__ void cvt8(const uint4& u, int dt, float (&v)[8]) {
  if (dt == DT_BF16) {
    v[0] = bf16lo(u.x); v[1] = bf16hi(u.x); v[2] = bf16lo(u.y); v[3] = bf16hi(u.y);
    v[4] = bf16lo(u.z); v[5] = bf16hi(u.z); v[6] = bf16lo(u.w); v[7] = bf16hi(u.w);
  } else {
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h[j]); v[2 * j] = f.x; v[2 * j + 1] = f.y; }
  }
}
__device__ __forceinline__ float ld1(const void* base, int dt, size_t i) {
  if (dt == DT_F32) return reinterpret_cast<const float*>(base)[i];
  if (dt == DT_BF16) return __bfloat162float(reinterpret_cast<const bf16*>(base)[i]);
  return __half2float(reinterpret_cast<const __half*>(base)[i]);
}
__device__ __forceinline__ void st1(void* base, int dt, size_t i, float v) {
  if (dt == DT_F32) reinterpret_cast<float*>(base)[i] = v;
  else if (dt == DT_BF16) reinterpret_cast<bf16*>(base)[i] = __float2bfloat16_rn(v);
  else reinterpret_cast<__half*>(base)[i] = __float2half_rn(v);
}

// ---- geometry of the elementwise (+ per-channel reduction) kernels ------------------------------------
// grid (pixel chunks, images); 256 threads = `pl` pixel lanes x `cvp` channel vectors (8 channels each).
struct RowGeom {
  int cvp, pl, cv, lane;
  bool active;
  __device__ __forceinline__ RowGeom(int cvecs) {
    cvp = cvecs < 256 ? cvecs : 256;
    pl = 256 / cvp;
    cv = threadIdx.x % cvp;
    lane = threadIdx.x / cvp;
    active = lane < pl;
  }
};
constexpr int kRowChunk = 2048;   // pixels per block at the high-resolution levels
// enough blocks to fill the machine at the low-resolution levels too (a block reduces `chunk` pixels before its atomics)
static inline int row_chunk(long long P, int N) {
  int c = kRowChunk;
  while (c > 128 && ((P + c - 1) / c) * N < 148 * 6) c >>= 1;
  return c;
}

// sum v[8] over the pixel lanes of the block and add to dst[(cbase + cv*8 + j) * stride] in fp64.  red: 256*8 floats.
// ncv: channel vectors of this pass that exist (<= cvp); `act` = this thread holds a valid partial.
__device__ __forceinline__ void lanes_reduce_add(const float (&v)[8], float* red, const RowGeom& g, bool act, int ncv, double* dst,
                                                 int stride) {
  __syncthreads();
  if (g.active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) red[(g.lane * g.cvp + g.cv) * 8 + j] = act ? v[j] : 0.f;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < ncv * 8; t += 256) {
    float s = 0.f;
    for (int l = 0; l < g.pl; ++l) s += red[l * g.cvp * 8 + t];
    atomicAdd(dst + (size_t)t * stride, (double)s);
  }
}

// =================================================================================================
// (1) gradient through clamp + per-channel products:  for a channel slice of width Cs
//       mode 0: T1 += sum_p g, T2 += sum_p g * x                               (no write)
//       mode 1: g <- g * [0 < a x + b < 6]  (ReLU6 backward, in place), then T1, T2 of the masked g
//     g: [N][P][ldg] slice starting at channel goff; x: [N][P][ldx] starting at xoff; coef / t12 indexed with goff.
__global__ void __launch_bounds__(256) bwd_mask_reduce_kernel(void* g, int dtg, int ldg, int goff, const void* x, int dtx,
                                                              int ldx, int xoff, const float2* __restrict__ coef, int coef_ld,
                                                              double* __restrict__ t12, int t_ld, int P, int Cs, int mode, int chunk) {
  __shared__ float red[256 * 8];
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * chunk, p1 = min(P, p0 + chunk);
  const int cvecs = Cs / 8;
  RowGeom G(cvecs);
  for (int cvb = 0; cvb < cvecs; cvb += G.cvp) {
    const int cv = cvb + G.cv;
    const bool act = G.active && cv < cvecs;
    float s1[8], s2[8];
    float2 ab[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; ab[j] = make_float2(0.f, 0.f); }
    if (act && mode == 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) ab[j] = coef[(size_t)n * coef_ld + goff + cv * 8 + j];
    }
    if (act) {
      int p = p0 + G.lane;
      if (dtg != DT_F32 && dtx != DT_F32) {
        constexpr int U = 4;   // four pixels per iteration, loads first (see bwd_affine3_kernel)
        for (; p + (U - 1) * G.pl < p1; p += U * G.pl) {
          uint4 g4[U], x4[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const size_t row = (size_t)n * P + p + u * G.pl;
            g4[u] = ldraw16(g, row * ldg + goff + cv * 8);
            x4[u] = ldraw16(x, row * ldx + xoff + cv * 8);
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const size_t row = (size_t)n * P + p + u * G.pl;
            float gv[8], xv[8];
            cvt8(g4[u], dtg, gv);
            cvt8(x4[u], dtx, xv);
            if (mode == 1) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float uu = fmaf(ab[j].x, xv[j], ab[j].y);
                gv[j] = (uu > 0.f && uu < 6.f) ? gv[j] : 0.f;
              }
              st8(g, dtg, row * ldg + goff + cv * 8, gv);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) { s1[j] += gv[j]; s2[j] = fmaf(gv[j], xv[j], s2[j]); }
          }
        }
      }
      for (; p < p1; p += G.pl) {
        const size_t row = (size_t)n * P + p;
        float gv[8], xv[8];
        ld8(g, dtg, row * ldg + goff + cv * 8, gv);
        ld8(x, dtx, row * ldx + xoff + cv * 8, xv);
        if (mode == 1) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float u = fmaf(ab[j].x, xv[j], ab[j].y);
            gv[j] = (u > 0.f && u < 6.f) ? gv[j] : 0.f;
          }
          st8(g, dtg, row * ldg + goff + cv * 8, gv);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { s1[j] += gv[j]; s2[j] = fmaf(gv[j], xv[j], s2[j]); }
      }
    }
    const int ncv = min(G.cvp, cvecs - cvb);
    double* dst = t12 + ((size_t)n * t_ld + goff + cvb * 8) * 2;
    lanes_reduce_add(s1, red, G, act, ncv, dst, 2);
    lanes_reduce_add(s2, red, G, act, ncv, dst + 1, 2);
  }
}

void launch_bwd_mask_reduce(void* g, int dtg, int ldg, int goff, const void* x, int dtx, int ldx, int xoff, const float2* coef,
                            int coef_ld, double* t12, int t_ld, int N, int P, int Cs, int mode, cudaStream_t st) {
  const int chunk = row_chunk(P, N);
  dim3 grid((P + chunk - 1) / chunk, N);
  bwd_mask_reduce_kernel<<<grid, 256, 0, st>>>(g, dtg, ldg, goff, x, dtx, ldx, xoff, coef, coef_ld, t12, t_ld, P, Cs, mode, chunk);
}

// =================================================================================================
// (2) dst (=|+=) A * g + B * x + C (+ r)   per (image, channel) coefficients (A, B, C) = coef4[n][coff + c].
//     Applies GroupNorm backward (see gn_bwd_coef) and adds the residual-path gradient r.
__global__ void __launch_bounds__(256) bwd_affine3_kernel(const void* g, int dtg, int ldg, int goff, const void* x, int dtx,
                                                          int ldx, int xoff, const float4* __restrict__ coef4, int coef_ld,
                                                          int coff, const void* r, int dtr, int ldr, int roff, void* dst,
                                                          int dtd, int ldd, int doff, int accumulate, int P, int Cs, int chunk) {
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * chunk, p1 = min(P, p0 + chunk);
  const int cvecs = Cs / 8;
  RowGeom G(cvecs);
  if (!G.active) return;
  for (int cv = G.cv; cv < cvecs; cv += G.cvp) {
    float4 k[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) k[j] = coef4[(size_t)n * coef_ld + coff + cv * 8 + j];
    int p = p0 + G.lane;
    if (dtg != DT_F32 && dtx != DT_F32 && dtd != DT_F32 && (!r || dtr != DT_F32)) {
      // 16-bit tensors: four pixels per iteration, all their loads issued before the first use (the one-pixel loop kept two
      // 16-byte loads in flight per thread: 4.3 TB/s)
      constexpr int U = 4;
      for (; p + (U - 1) * G.pl < p1; p += U * G.pl) {
        uint4 g4[U], x4[U], r4[U], d4[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const size_t row = (size_t)n * P + p + u * G.pl;
          g4[u] = ldraw16(g, row * ldg + goff + cv * 8);
          x4[u] = ldraw16(x, row * ldx + xoff + cv * 8);
          if (r) r4[u] = ldraw16(r, row * ldr + roff + cv * 8);
          if (accumulate) d4[u] = ldraw16(dst, row * ldd + doff + cv * 8);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const size_t row = (size_t)n * P + p + u * G.pl;
          float gv[8], xv[8], o[8];
          cvt8(g4[u], dtg, gv);
          cvt8(x4[u], dtx, xv);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaf(k[j].x, gv[j], fmaf(k[j].y, xv[j], k[j].z));
          if (r) {
            float rv[8];
            cvt8(r4[u], dtr, rv);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] += rv[j];
          }
          if (accumulate) {
            float dv[8];
            cvt8(d4[u], dtd, dv);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] += dv[j];
          }
          st8(dst, dtd, row * ldd + doff + cv * 8, o);
        }
      }
    }
    for (; p < p1; p += G.pl) {
      const size_t row = (size_t)n * P + p;
      float gv[8], xv[8], o[8];
      ld8(g, dtg, row * ldg + goff + cv * 8, gv);
      ld8(x, dtx, row * ldx + xoff + cv * 8, xv);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(k[j].x, gv[j], fmaf(k[j].y, xv[j], k[j].z));
      if (r) {
        float rv[8];
        ld8(r, dtr, row * ldr + roff + cv * 8, rv);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += rv[j];
      }
      if (accumulate) {
        float dv[8];
        ld8(dst, dtd, row * ldd + doff + cv * 8, dv);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += dv[j];
      }
      st8(dst, dtd, row * ldd + doff + cv * 8, o);
    }
  }
}

void launch_bwd_affine3(const void* g, int dtg, int ldg, int goff, const void* x, int dtx, int ldx, int xoff,
                        const float4* coef4, int coef_ld, int coff, const void* r, int dtr, int ldr, int roff, void* dst,
                        int dtd, int ldd, int doff, int accumulate, int N, int P, int Cs, cudaStream_t st) {
  const int chunk = row_chunk(P, N);
  dim3 grid((P + chunk - 1) / chunk, N);
  bwd_affine3_kernel<<<grid, 256, 0, st>>>(g, dtg, ldg, goff, x, dtx, ldx, xoff, coef4, coef_ld, coff, r, dtr, ldr, roff, dst,
                                           dtd, ldd, doff, accumulate, P, Cs, chunk);
}

// dst (=|+=) src on channel slices (gradient of an identity edge, e.g. the attention residual)
__global__ void __launch_bounds__(256) bwd_add_kernel(const void* src, int dts, int lds, int soff, void* dst, int dtd, int ldd,
                                                      int doff, int accumulate, long long rows, int Cs) {
  const int cvecs = Cs / 8;
  const long long total = rows * cvecs;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long row = i / cvecs;
    const int cv = (int)(i % cvecs);
    float v[8];
    ld8(src, dts, (size_t)row * lds + soff + cv * 8, v);
    if (accumulate) {
      float d[8];
      ld8(dst, dtd, (size_t)row * ldd + doff + cv * 8, d);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += d[j];
    }
    st8(dst, dtd, (size_t)row * ldd + doff + cv * 8, v);
  }
}
void launch_bwd_add(const void* src, int dts, int lds, int soff, void* dst, int dtd, int ldd, int doff, int accumulate,
                    long long rows, int Cs, cudaStream_t st) {
  long long blocks = (rows * (Cs / 8) + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  bwd_add_kernel<<<(int)blocks, 256, 0, st>>>(src, dts, lds, soff, dst, dtd, ldd, doff, accumulate, rows, Cs);
}

// =================================================================================================
// (3) GroupNorm (+FiLM) backward finalise, one block per image (mirror of gn_coef_kernel).
//   forward: xh = (x - mu_g) r_g ; y = gamma xh + beta ; u = y (1 + s) + sh      (s = sh = 0 without FiLM)
//   given T1_c = sum_p du, T2_c = sum_p du x:
//     d sh_c = T1 ; d s_c = gamma r (T2 - mu T1) + beta T1
//     d beta_c += (1 + s) T1 ; d gamma_c += (1 + s) r (T2 - mu T1)
//     k_c = gamma (1 + s) ; m1_g = sum_{c in g} k T1 / cnt ; m2_g = sum_{c in g} k r (T2 - mu T1) / cnt
//     dx = A du + B x + C with A = r k, B = -r^2 m2, C = -r m1 + r^2 m2 mu
__global__ void __launch_bounds__(256) gn_bwd_coef_kernel(const double* __restrict__ t12, const double* __restrict__ s0, int C0,
                                                          const double* __restrict__ s1, int C1, int groups, double count,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          const float* __restrict__ film, float* __restrict__ dfilm, int film_ld,
                                                          float4* __restrict__ coef4, float* __restrict__ dgamma,
                                                          float* __restrict__ dbeta, const float* __restrict__ ascale,
                                                          int ascale_stride) {
  __shared__ float s_mean[64], s_rstd[64], s_m1[64], s_m2[64];
  const int n = blockIdx.x;
  const int C = C0 + C1;
  const int cpg = C / groups;
  const double inv_count = 1.0 / count;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int g = warp; g < groups; g += (int)(blockDim.x >> 5)) {
    double sum = 0.0, sq = 0.0;
    for (int c = g * cpg + lane; c < (g + 1) * cpg; c += 32) {
      const double2 v = *reinterpret_cast<const double2*>((c < C0) ? s0 + ((size_t)n * C0 + c) * 2 : s1 + ((size_t)n * C1 + (c - C0)) * 2);
      sum += v.x;
      sq += v.y;
    }
    for (int o = 16; o > 0; o >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, o); sq += __shfl_xor_sync(0xffffffffu, sq, o); }
    const double mean = sum * inv_count;
    double var = sq * inv_count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float mu = (float)mean, r = rsqrtf((float)var + 1e-5f);   // identical to the forward's (gn_coef_kernel)
    double a1 = 0.0, a2 = 0.0;
    for (int c = g * cpg + lane; c < (g + 1) * cpg; c += 32) {
      const double T1 = t12[((size_t)n * C + c) * 2], T2 = t12[((size_t)n * C + c) * 2 + 1];
      const double sc = film ? 1.0 + (double)film[(size_t)n * film_ld + c] : 1.0;
      const double k = (double)gamma[c] * sc;
      a1 += k * T1;
      a2 += k * (double)r * (T2 - (double)mu * T1);
    }
    for (int o = 16; o > 0; o >>= 1) { a1 += __shfl_xor_sync(0xffffffffu, a1, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o); }
    if (lane == 0) {
      s_mean[g] = mu;
      s_rstd[g] = r;
      s_m1[g] = (float)(a1 * inv_count);
      s_m2[g] = (float)(a2 * inv_count);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float mu = s_mean[g], r = s_rstd[g], m1 = s_m1[g], m2 = s_m2[g];
    const float T1 = (float)t12[((size_t)n * C + c) * 2];
    const float T2c = (float)(t12[((size_t)n * C + c) * 2 + 1] - (double)mu * t12[((size_t)n * C + c) * 2]);   // sum du (x - mu)
    const float sc = film ? 1.f + film[(size_t)n * film_ld + c] : 1.f;
    const float gm = gamma[c], bt = beta[c];
    const float dxhat_y = r * T2c;            // sum_p du * xhat
    if (dfilm) {
      dfilm[(size_t)n * film_ld + c] = gm * dxhat_y + bt * T1;      // d scale
      dfilm[(size_t)n * film_ld + C + c] = T1;                      // d shift
    }
    atomicAdd(dgamma + c, sc * dxhat_y);
    atomicAdd(dbeta + c, sc * T1);
    const float k = gm * sc;
    // ascale: the gradient tensor this A multiplies is stored scaled (du * s of dwconv_bwd_stream.cu): A carries the 1/s
    const float as = ascale ? ascale[((size_t)n * C + c) * ascale_stride] : 1.f;
    coef4[(size_t)n * C + c] = make_float4(r * k * as, -r * r * m2, -r * m1 + r * r * m2 * mu, 0.f);
  }
}

void launch_gn_bwd_coef(const double* t12, const double* stats0, int C0, const double* stats1, int C1, int groups, double count,
                        const float* gamma, const float* beta, const float* film, float* dfilm, int film_ld, float4* coef4,
                        float* dgamma, float* dbeta, int N, cudaStream_t st, const float* ascale, int ascale_stride) {
  gn_bwd_coef_kernel<<<N, 256, 0, st>>>(t12, stats0, C0, stats1, C1, groups, count, gamma, beta, film, dfilm, film_ld, coef4,
                                        dgamma, dbeta, ascale, ascale_stride);
}

// =================================================================================================
// (4) SE backward, vector part (efficient_unet.py:96-100).  One block per image.
//   forward: pm = pool / P ; z = relu6(w1 pm + b1) ; gate = sigmoid(w2 z + b2) ; out = h2 * gate
//   given dgate_c = sum_p dq h2 (t12[..][1] of bwd_mask_reduce mode 0):
//     ds2 = dgate gate (1 - gate) ; dz = w2^T ds2 ; dz1 = dz [0 < zpre < 6] ; dpm = w1^T dz1
//   outputs: coef_se[n][c] = (gate, dpm / P)  so that  d h2 = gate * dq + dpm / P  (prologue of the depthwise backward),
//            and the vectors pm, z, ds2, dz1 for the weight-gradient outer products (outer_sum_kernel).
__global__ void __launch_bounds__(256) se_bwd_vec_kernel(const double* __restrict__ pool, float inv_count,
                                                         const float* __restrict__ w1, const float* __restrict__ b1,
                                                         const float* __restrict__ w2, const float2* __restrict__ gate,
                                                         const double* __restrict__ t12, const double* __restrict__ dq_stats,
                                                         float4* __restrict__ coef_se, float* __restrict__ v_pm,
                                                         float* __restrict__ v_z, float* __restrict__ v_ds2,
                                                         float* __restrict__ v_dz1, int C, int SQ) {
  extern __shared__ float sm[];
  float* pm = sm;            // [C]
  float* ds2 = pm + C;       // [C]
  float* z = ds2 + C;        // [SQ]
  float* dz1 = z + SQ;       // [SQ]
  const int n = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // coef_se[n][c] = (gate s, dpm/P s, 1/s, 0).  s = 1 unless the streaming depthwise backward runs its packed-fp16
  // arithmetic on this block (dq_stats = per-(image, channel) sum / sum^2 of dq from the project-dgrad GEMM): then s is the
  // power of two with (gate rms(dq) + |dpm/P|) s in (0.5, 1] — see dwconv_bwd_stream.cu.
  auto write_cse = [&](int c, float g, float dpm_p) {
    float sc = 1.f;
    if (dq_stats) {
      const float ms = fmaxf((float)dq_stats[((size_t)n * C + c) * 2 + 1], 0.f) * inv_count;
      const float bound = g * sqrtf(ms) + fabsf(dpm_p);
      if (bound > 0.f && bound < 1e30f) {
        int e;
        frexpf(1.f / bound, &e);                        // 1 / bound = m 2^e, m in [0.5, 1)
        e = max(-40, min(60, e - 1));
        sc = ldexpf(1.f, e);
      }
    }
    coef_se[(size_t)n * C + c] = make_float4(g * sc, dpm_p * sc, 1.f / sc, 0.f);
  };
  for (int c = tid; c < C; c += 256) {
    const float m = (float)pool[(size_t)n * C + c] * inv_count;
    const float g = gate[(size_t)n * C + c].x;
    pm[c] = m;
    ds2[c] = (float)t12[((size_t)n * C + c) * 2 + 1] * g * (1.f - g);
    v_pm[(size_t)n * C + c] = m;
    v_ds2[(size_t)n * C + c] = ds2[c];
  }
  __syncthreads();
  // The three FC products are latency problems (64 blocks, weights of up to 2 x 1 MB out of L2): every stage keeps
  // several independent 16-byte loads in flight per thread and spreads its reduction dimension over the whole block.
  const bool vec = (C % 128 == 0) && (SQ % 4 == 0) && SQ <= 1024;
  float* scr = dz1 + SQ;     // [1024] partial sums of stages 2 and 3
  if (vec) {
    for (int j = warp; j < SQ; j += 8) {          // z_pre[j] = b1[j] + w1[j][:] . pm
      const float4* wr = reinterpret_cast<const float4*>(w1 + (size_t)j * C);
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      for (int c4 = lane; c4 < C / 4; c4 += 32) {
        const float4 w = wr[c4];
        const float4 m = *reinterpret_cast<const float4*>(pm + c4 * 4);
        a0 = fmaf(w.x, m.x, a0); a1 = fmaf(w.y, m.y, a1); a2 = fmaf(w.z, m.z, a2); a3 = fmaf(w.w, m.w, a3);
      }
      float acc = (a0 + a1) + (a2 + a3);
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) {
        const float zp = acc + b1[j];
        z[j] = fminf(fmaxf(zp, 0.f), 6.f);
        dz1[j] = (zp > 0.f && zp < 6.f) ? 1.f : 0.f;   // mask for now
      }
    }
    __syncthreads();
    {                                             // dz[j] = sum_c ds2[c] w2[c][j]: thread = (4 columns j, slice of c)
      const int nj = SQ / 4;                      // threads across j
      const int parts = nj >= 256 ? 1 : 256 / nj; // slices of c (nj is a power of two or a divisor of 256 for the presets)
      for (int j4 = tid % nj; j4 < nj; j4 += (nj >= 256 ? 256 : nj)) {
        const int part = nj >= 256 ? 0 : tid / nj;
        if (part < parts) {
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
          int c = part;
          for (; c + parts < C; c += 2 * parts) {
            const float4 w0 = *reinterpret_cast<const float4*>(w2 + (size_t)c * SQ + j4 * 4);
            const float4 w1v = *reinterpret_cast<const float4*>(w2 + (size_t)(c + parts) * SQ + j4 * 4);
            const float d0 = ds2[c], d1 = ds2[c + parts];
            a.x = fmaf(d0, w0.x, a.x); a.y = fmaf(d0, w0.y, a.y); a.z = fmaf(d0, w0.z, a.z); a.w = fmaf(d0, w0.w, a.w);
            b.x = fmaf(d1, w1v.x, b.x); b.y = fmaf(d1, w1v.y, b.y); b.z = fmaf(d1, w1v.z, b.z); b.w = fmaf(d1, w1v.w, b.w);
          }
          for (; c < C; c += parts) {
            const float4 w0 = *reinterpret_cast<const float4*>(w2 + (size_t)c * SQ + j4 * 4);
            const float d0 = ds2[c];
            a.x = fmaf(d0, w0.x, a.x); a.y = fmaf(d0, w0.y, a.y); a.z = fmaf(d0, w0.z, a.z); a.w = fmaf(d0, w0.w, a.w);
          }
          a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
          if (parts == 1) {
            dz1[j4 * 4 + 0] *= a.x; dz1[j4 * 4 + 1] *= a.y; dz1[j4 * 4 + 2] *= a.z; dz1[j4 * 4 + 3] *= a.w;
          } else {
            *reinterpret_cast<float4*>(scr + (part * nj + j4) * 4) = a;   // parts * SQ <= 1024 floats
          }
        }
        if (nj < 256) break;
      }
      if (parts > 1) {
        __syncthreads();
        for (int j = tid; j < SQ; j += 256) {
          float acc = 0.f;
          for (int q = 0; q < parts; ++q) acc += scr[q * SQ + j];
          dz1[j] *= acc;
        }
      }
      __syncthreads();
      for (int j = tid; j < SQ; j += 256) {
        v_z[(size_t)n * SQ + j] = z[j];
        v_dz1[(size_t)n * SQ + j] = dz1[j];
      }
    }
    {                                             // dpm[c] = sum_j dz1[j] w1[j][c]: thread = (4 channels c, slice of j)
      const int nc = C / 4;
      if (nc >= 256) {
        for (int c4 = tid; c4 < nc; c4 += 256) {
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
          int j = 0;
          for (; j + 1 < SQ; j += 2) {
            const float4 w0 = *reinterpret_cast<const float4*>(w1 + (size_t)j * C + c4 * 4);
            const float4 w1v = *reinterpret_cast<const float4*>(w1 + (size_t)(j + 1) * C + c4 * 4);
            const float d0 = dz1[j], d1 = dz1[j + 1];
            a.x = fmaf(d0, w0.x, a.x); a.y = fmaf(d0, w0.y, a.y); a.z = fmaf(d0, w0.z, a.z); a.w = fmaf(d0, w0.w, a.w);
            b.x = fmaf(d1, w1v.x, b.x); b.y = fmaf(d1, w1v.y, b.y); b.z = fmaf(d1, w1v.z, b.z); b.w = fmaf(d1, w1v.w, b.w);
          }
          for (; j < SQ; ++j) {
            const float4 w0 = *reinterpret_cast<const float4*>(w1 + (size_t)j * C + c4 * 4);
            const float d0 = dz1[j];
            a.x = fmaf(d0, w0.x, a.x); a.y = fmaf(d0, w0.y, a.y); a.z = fmaf(d0, w0.z, a.z); a.w = fmaf(d0, w0.w, a.w);
          }
          const float r[4] = {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w};
#pragma unroll
          for (int e = 0; e < 4; ++e)
            write_cse(c4 * 4 + e, gate[(size_t)n * C + c4 * 4 + e].x, r[e] * inv_count);
        }
      } else {
        const int parts = 256 / nc, c4 = tid % nc, part = tid / nc;   // nc in {32, 64, 96 -> parts 2 (64 threads idle), 128, 192}
        if (part < parts) {
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int j = part; j < SQ; j += parts) {
            const float4 w0 = *reinterpret_cast<const float4*>(w1 + (size_t)j * C + c4 * 4);
            const float d0 = dz1[j];
            a.x = fmaf(d0, w0.x, a.x); a.y = fmaf(d0, w0.y, a.y); a.z = fmaf(d0, w0.z, a.z); a.w = fmaf(d0, w0.w, a.w);
          }
          *reinterpret_cast<float4*>(scr + (part * nc + c4) * 4) = a;   // parts * C <= 1024 floats
        }
        __syncthreads();
        for (int c = tid; c < C; c += 256) {
          float acc = 0.f;
          for (int q = 0; q < parts; ++q) acc += scr[q * C + c];
          write_cse(c, gate[(size_t)n * C + c].x, acc * inv_count);
        }
      }
    }
    return;
  }
  for (int j = warp; j < SQ; j += 8) {          // z_pre[j] = b1[j] + w1[j][:] . pm
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) acc = fmaf(w1[(size_t)j * C + c], pm[c], acc);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      const float zp = acc + b1[j];
      z[j] = fminf(fmaxf(zp, 0.f), 6.f);
      dz1[j] = (zp > 0.f && zp < 6.f) ? 1.f : 0.f;   // mask for now
    }
  }
  __syncthreads();
  for (int j = tid; j < SQ; j += 256) {           // dz[j] = sum_c ds2[c] w2[c][j]
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc = fmaf(ds2[c], w2[(size_t)c * SQ + j], acc);
    dz1[j] *= acc;
    v_z[(size_t)n * SQ + j] = z[j];
    v_dz1[(size_t)n * SQ + j] = dz1[j];
  }
  __syncthreads();
  for (int c = tid; c < C; c += 256) {            // dpm[c] = sum_j dz1[j] w1[j][c]
    float acc = 0.f;
    for (int j = 0; j < SQ; ++j) acc = fmaf(dz1[j], w1[(size_t)j * C + c], acc);
    write_cse(c, gate[(size_t)n * C + c].x, acc * inv_count);
  }
}

int launch_se_bwd_vec(const double* pool, float inv_count, const float* w1, const float* b1, const float* w2, const float2* gate,
                      const double* t12, const double* dq_stats, float4* coef_se, float* v_pm, float* v_z, float* v_ds2,
                      float* v_dz1, int N, int C, int SQ, cudaStream_t st) {
  const size_t smem = (size_t)(2 * C + 2 * SQ + 1024) * sizeof(float);
  if (ensure_dyn_smem_fn(se_bwd_vec_kernel, smem)) return 1;
  se_bwd_vec_kernel<<<N, 256, smem, st>>>(pool, inv_count, w1, b1, w2, gate, t12, dq_stats, coef_se, v_pm, v_z, v_ds2, v_dz1, C, SQ);
  return 0;
}

// SE gate gradient without a pass over dq and h2: with the per-image products R[n][k][o] = sum_p [h2 | x][p][k] dY[p][o] of
// the project weight-gradient GEMM (h2 not gated) and dq = dY Wp,
//     dgate[n][c] = sum_p dq[p][c] h2[p][c] = sum_o Wp[o][c] R[n][c][o]
//     dWp[o][c]   = sum_n gate[n][c] R[n][c][o] ;  dWskip[o][ci] = sum_n R[n][Ch + ci][o]
// One block per k (128 threads = Co-wide rows x image slices); R is a few MB.
__global__ void __launch_bounds__(128) se_project_combine_kernel(const float* __restrict__ R, const float* __restrict__ Wp,
                                                                 const float2* __restrict__ gate, double* __restrict__ t12,
                                                                 float* __restrict__ dWp, float* __restrict__ dWs, int N, int Ch,
                                                                 int Ci, int Co) {
  const int k = blockIdx.x, Ktot = Ch + Ci;
  const int lanes_o = Co < 128 ? Co : 128;             // threads across o (Co is a multiple of 16; 32 | lanes_o or lanes_o = 16 / 48 ...)
  const int slices = 128 / lanes_o;
  const int ot = threadIdx.x % lanes_o, sl = threadIdx.x / lanes_o;
  const bool warp_rows = lanes_o % 32 == 0 && Co % 32 == 0;
  if (sl >= slices) return;
  for (int o = ot; o < Co; o += lanes_o) {
    float wacc = 0.f;
    const float wp = k < Ch ? Wp[(size_t)o * Ch + k] : 0.f;
    for (int n = sl; n < N; n += slices) {
      const float r = R[((size_t)n * Ktot + k) * Co + o];
      if (k < Ch) {
        wacc = fmaf(gate[(size_t)n * Ch + k].x, r, wacc);
        float d = wp * r;
        if (warp_rows) {      // the 32 lanes of a warp hold 32 output channels of the same (image, k)
          for (int sft = 16; sft > 0; sft >>= 1) d += __shfl_xor_sync(0xffffffffu, d, sft);
          if ((threadIdx.x & 31) == 0) atomicAdd(t12 + ((size_t)n * Ch + k) * 2 + 1, (double)d);
        } else {
          atomicAdd(t12 + ((size_t)n * Ch + k) * 2 + 1, (double)d);
        }
      } else {
        wacc += r;
      }
    }
    if (k < Ch) atomicAdd(dWp + (size_t)o * Ch + k, wacc);
    else if (dWs) atomicAdd(dWs + (size_t)o * Ci + (k - Ch), wacc);
  }
}
void launch_se_project_combine(const float* R, const float* Wp, const float2* gate, double* t12, float* dWp, float* dWs, int N,
                               int Ch, int Ci, int Co, cudaStream_t st) {
  se_project_combine_kernel<<<Ch + Ci, 128, 0, st>>>(R, Wp, gate, t12, dWp, dWs, N, Ch, Ci, Co);
}

// dW[r][c] += sum_n A[n][r] * B[n][c] ; dbias[r] += sum_n A[n][r]   (weight gradients of the tiny FCs: SE, FiLM, time MLP)
// 64 x 64 output tile per block, both operand tiles staged in shared memory, 4 x 4 outputs per thread: the FiLM table
// (R = all blocks' scale / shift rows, Cc = time_embed_dim) took 0.49 ms with one output per thread reading A and B from L2.
__global__ void __launch_bounds__(256) outer_sum_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                                        float* __restrict__ dW, float* __restrict__ dbias, int N, int R, int Cc) {
  __shared__ float sa[32][64 + 4], sb[32][64 + 4];
  const int tid = threadIdx.x, tr = tid >> 4, tc = tid & 15;
  const int tiles_c = (Cc + 63) / 64;
  const int r0 = (blockIdx.x / tiles_c) * 64, c0 = (blockIdx.x % tiles_c) * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
  for (int n0 = 0; n0 < N; n0 += 32) {
    __syncthreads();
    for (int i = tid; i < 32 * 64; i += 256) {
      const int n = n0 + (i >> 6), k = i & 63;
      sa[i >> 6][k] = (n < N && r0 + k < R) ? A[(size_t)n * lda + r0 + k] : 0.f;
      sb[i >> 6][k] = (n < N && c0 + k < Cc) ? B[(size_t)n * ldb + c0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int n = 0; n < 32; ++n) {
      const float4 a = *reinterpret_cast<const float4*>(&sa[n][tr * 4]);
      const float4 bq = *reinterpret_cast<const float4*>(&sb[n][tc * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        bsum[i] += av[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + tr * 4 + i;
    if (r >= R) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + tc * 4 + j;
      if (c < Cc) dW[(size_t)r * Cc + c] += acc[i][j];   // exclusive owner of the entry
    }
    if (dbias && c0 == 0 && tc == 0) dbias[r] += bsum[i];
  }
}
void launch_outer_sum(const float* A, int lda, const float* B, int ldb, float* dW, float* dbias, int N, int R, int Cc,
                      cudaStream_t st) {
  const int blocks = ((R + 63) / 64) * ((Cc + 63) / 64);
  outer_sum_kernel<<<blocks, 256, 0, st>>>(A, lda, B, ldb, dW, dbias, N, R, Cc);
}

// =================================================================================================
// (5) depthwise 3x3 backward (efficient_unet.py:212-223 reversed), fused with the SE-scale backward in its prologue and
//     the ReLU6 backward + GroupNorm2 reductions in its epilogue.
//   dh2 = gate * dq + dpm/P                       (coef_se, zero outside the image)
//   v   = relu6(a2 h1 + b2)                       (coef2, zero outside the image: the forward conv pads v)
//   dv[p] = sum_tap w[tap] dh2[p - off(tap)] ;  du = dv [0 < a2 h1 + b2 < 6]
//   dW[c][tap] += sum_p dh2[p] v[p + off(tap)] ;  S1 += sum_p du ; S2 += sum_p du h1
// A block owns a 32-channel slice and a contiguous range of (image, 8x16 tile) items; weight gradients stay in
// registers over the whole range, the per-image sums are flushed when the image changes.  Two fp32 halo tiles of
// 10 x 18 x 32 = 46 KB: four blocks per SM overlap one another's load and compute phases.
template <bool F32>   // F32: the fp32 plan (all tensors fp32); else 16-bit tensors (gradients bf16, h1 bf16 or fp16)
__global__ void __launch_bounds__(256, 2) dwconv_bwd_kernel(const void* __restrict__ dq, int dtg, const float4* __restrict__ coef_se,
                                                            const void* __restrict__ h1, int dth, const float2* __restrict__ coef2,
                                                            const float* __restrict__ w, void* __restrict__ du,
                                                            double* __restrict__ t12, float* __restrict__ dW, int N, int H, int W,
                                                            int C, int tilesX, int tilesY, int items_per_block) {
  constexpr int TSY = 8, TSX = 16, HSY = TSY + 2, HS = TSX + 2, CB = 32;
  extern __shared__ __align__(16) float dsm[];
  float* tg = dsm;                      // [HSY*HS][CB] dh2 halo tile
  float* tv = dsm + HSY * HS * CB;      // [HSY*HS][CB] v halo tile
  __shared__ __align__(16) float s_w[9 * CB];
  __shared__ float s_red[8][2 * CB];
  __shared__ __align__(16) float2 s_cse[CB], s_c2[CB];   // this image's prologue coefficients (reloaded when the image changes)
  const int tid = threadIdx.x;
  const int c0 = blockIdx.y * CB;
  const int cg = tid & 3;               // fill phase: 8-channel group
  // compute phase: a thread owns 4 channels x 4 consecutive pixels of one row.  The 8 lanes of a quarter warp read the 32
  // channels (128 contiguous bytes) of ONE pixel: every LDS.128 is conflict-free; a halo row is loaded once per thread and
  // serves all taps of that row (36 LDS.128 + 9 weight loads per 16 outputs).
  const int q = tid & 7, pg = tid >> 3;
  const int row = pg >> 2, xs = (pg & 3) * 4;
  const int tiles = tilesX * tilesY;
  const long long total = (long long)N * tiles;
  const long long i0 = (long long)blockIdx.x * items_per_block;
  const long long i1 = i0 + items_per_block < total ? i0 + items_per_block : total;
  for (int i = tid; i < 9 * CB; i += 256) s_w[i] = w[(size_t)(i / CB) * C + c0 + (i % CB)];
  float4 dw[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) dw[t] = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  int cur_n = -1;

  auto flush = [&](int n) {
    // s2 holds sum du * v (v = a2 h1 + b2 wherever du != 0): sum du * h1 = (s2 - b2 s1) / a2
    {
      const float s1v[4] = {s1.x, s1.y, s1.z, s1.w};
      float s2v[4] = {s2.x, s2.y, s2.z, s2.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 ab = s_c2[q * 4 + j];
        s2v[j] = fabsf(ab.x) > 1e-30f ? (s2v[j] - ab.y * s1v[j]) / ab.x : 0.f;
      }
      s2 = make_float4(s2v[0], s2v[1], s2v[2], s2v[3]);
    }
    // lanes with equal q (l, l^8, l^16) hold the same channels: reduce them, then the 8 warps through shared memory
    float a[4] = {s1.x, s1.y, s1.z, s1.w}, b[4] = {s2.x, s2.y, s2.z, s2.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      a[j] += __shfl_xor_sync(0xffffffffu, a[j], 8); b[j] += __shfl_xor_sync(0xffffffffu, b[j], 8);
      a[j] += __shfl_xor_sync(0xffffffffu, a[j], 16); b[j] += __shfl_xor_sync(0xffffffffu, b[j], 16);
      if ((tid & 31) < 8) { s_red[tid >> 5][q * 4 + j] = a[j]; s_red[tid >> 5][CB + q * 4 + j] = b[j]; }
    }
    s1 = make_float4(0.f, 0.f, 0.f, 0.f); s2 = s1;
    __syncthreads();
    if (tid < 2 * CB) {
      float s = 0.f;
#pragma unroll
      for (int wi = 0; wi < 8; ++wi) s += s_red[wi][tid];
      const int c = tid % CB, which = tid / CB;
      atomicAdd(t12 + ((size_t)n * C + c0 + c) * 2 + which, (double)s);
    }
    __syncthreads();
  };

  for (long long it = i0; it < i1; ++it) {
    const int n = (int)(it / tiles), tile = (int)(it % tiles);
    if (n != cur_n) {
      if (cur_n >= 0) flush(cur_n);     // (ends with a block-wide barrier: nobody still reads the old coefficients)
      cur_n = n;
      if (tid < CB) { const float4 e = coef_se[(size_t)n * C + c0 + tid]; s_cse[tid] = make_float2(e.x, e.y); }   // (s = 1 on this path)
      else if (tid < 2 * CB) s_c2[tid - CB] = coef2[(size_t)n * C + c0 + tid - CB];
    }
    const int ty0 = (tile / tilesX) * TSY, tx0 = (tile % tilesX) * TSX;
    __syncthreads();   // previous tile fully consumed; coefficients visible
    {
      float2 cse[8], c2[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        cse[j] = s_cse[cg * 8 + j];
        c2[j] = s_c2[cg * 8 + j];
      }
      // 720 (pixel, channel-group) items over 256 threads: three fully unrolled rounds, every global load issued before the
      // first use (the fill is latency-bound otherwise: one DRAM round trip per round)
      constexpr int ROUNDS = (HSY * HS * 4 + 255) / 256;
      uint4 ra[ROUNDS][F32 ? 2 : 1], rb[ROUNDS][F32 ? 2 : 1];
      bool inside[ROUNDS];
#pragma unroll
      for (int rd = 0; rd < ROUNDS; ++rd) {
        const int i = tid + rd * 256;
        const int px = i >> 2;
        const int yy = px / HS, xx = px - yy * HS;
        const int gy = ty0 + yy - 1, gx = tx0 + xx - 1;
        inside[rd] = i < HSY * HS * 4 && gy >= 0 && gy < H && gx >= 0 && gx < W;
        if (inside[rd]) {
          const size_t o = (((size_t)n * H + gy) * W + gx) * C + c0 + cg * 8;
          if (F32) {
            ra[rd][0] = *reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(dq) + o);
            ra[rd][F32 ? 1 : 0] = *reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(dq) + o + 4);
            rb[rd][0] = *reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(h1) + o);
            rb[rd][F32 ? 1 : 0] = *reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(h1) + o + 4);
          } else {
            ra[rd][0] = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(dq) + o);
            rb[rd][0] = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(h1) + o);   // bf16 or fp16: 16 bytes
          }
        }
      }
#pragma unroll
      for (int rd = 0; rd < ROUNDS; ++rd) {
        const int i = tid + rd * 256;
        if (i >= HSY * HS * 4) continue;
        const int px = i >> 2;
        float a[8], b[8];
        if (inside[rd]) {
          if (F32) {
            const uint4 u0 = ra[rd][0], u1 = ra[rd][F32 ? 1 : 0];
            a[0] = __uint_as_float(u0.x); a[1] = __uint_as_float(u0.y); a[2] = __uint_as_float(u0.z); a[3] = __uint_as_float(u0.w);
            a[4] = __uint_as_float(u1.x); a[5] = __uint_as_float(u1.y); a[6] = __uint_as_float(u1.z); a[7] = __uint_as_float(u1.w);
          } else {
            const uint4 u = ra[rd][0];
            a[0] = bf16lo(u.x); a[1] = bf16hi(u.x); a[2] = bf16lo(u.y); a[3] = bf16hi(u.y);
            a[4] = bf16lo(u.z); a[5] = bf16hi(u.z); a[6] = bf16lo(u.w); a[7] = bf16hi(u.w);
          }
          if (F32) {
            const uint4 u0 = rb[rd][0], u1 = rb[rd][F32 ? 1 : 0];
            b[0] = __uint_as_float(u0.x); b[1] = __uint_as_float(u0.y); b[2] = __uint_as_float(u0.z); b[3] = __uint_as_float(u0.w);
            b[4] = __uint_as_float(u1.x); b[5] = __uint_as_float(u1.y); b[6] = __uint_as_float(u1.z); b[7] = __uint_as_float(u1.w);
          } else if (dth == DT_BF16) {
            const uint4 u = rb[rd][0];
            b[0] = bf16lo(u.x); b[1] = bf16hi(u.x); b[2] = bf16lo(u.y); b[3] = bf16hi(u.y);
            b[4] = bf16lo(u.z); b[5] = bf16hi(u.z); b[6] = bf16lo(u.w); b[7] = bf16hi(u.w);
          } else {
            const uint4 u = rb[rd][0];
            const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
            for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h[j]); b[2 * j] = f.x; b[2 * j + 1] = f.y; }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            a[j] = fmaf(cse[j].x, a[j], cse[j].y);
            b[j] = fminf(fmaxf(fmaf(c2[j].x, b[j], c2[j].y), 0.f), 6.f);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) { a[j] = 0.f; b[j] = 0.f; }
        }
        float4* d = reinterpret_cast<float4*>(tg + px * CB + cg * 8);
        d[0] = make_float4(a[0], a[1], a[2], a[3]);
        d[1] = make_float4(a[4], a[5], a[6], a[7]);
        float4* e = reinterpret_cast<float4*>(tv + px * CB + cg * 8);
        e[0] = make_float4(b[0], b[1], b[2], b[3]);
        e[1] = make_float4(b[4], b[5], b[6], b[7]);
      }
    }
    __syncthreads();
    const int gy = ty0 + row;
    // centre dh2 of the 4 pixels (halo coordinates (row + 1, xs + 1 + p))
    float4 gc[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) gc[p] = *reinterpret_cast<const float4*>(tg + ((row + 1) * HS + xs + 1 + p) * CB + q * 4);
    float4 dv[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) dv[p] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int hr = 0; hr < 3; ++hr) {
      // halo row (row + hr), columns xs .. xs + 5
      float4 v6[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) v6[i] = *reinterpret_cast<const float4*>(tg + ((row + hr) * HS + xs + i) * CB + q * 4);
      // dv[p] += w[ky][kx] * dh2[p - off(tap)] : halo (row + 2 - ky, x + 2 - kx)  =>  ky = 2 - hr, column index i = p + 2 - kx
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float4 wv = *reinterpret_cast<const float4*>(s_w + ((2 - hr) * 3 + kx) * CB + q * 4);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const float4 g4 = v6[p + 2 - kx];
          dv[p].x = fmaf(g4.x, wv.x, dv[p].x); dv[p].y = fmaf(g4.y, wv.y, dv[p].y);
          dv[p].z = fmaf(g4.z, wv.z, dv[p].z); dv[p].w = fmaf(g4.w, wv.w, dv[p].w);
        }
      }
      // dW[ky][kx] += dh2[p] * v[p + off(tap)] : halo (row + ky, x + kx)  =>  ky = hr, column index i = p + kx
#pragma unroll
      for (int i = 0; i < 6; ++i) v6[i] = *reinterpret_cast<const float4*>(tv + ((row + hr) * HS + xs + i) * CB + q * 4);
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        float4 acc = dw[hr * 3 + kx];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const float4 v4 = v6[p + kx];
          acc.x = fmaf(gc[p].x, v4.x, acc.x); acc.y = fmaf(gc[p].y, v4.y, acc.y);
          acc.z = fmaf(gc[p].z, v4.z, acc.z); acc.w = fmaf(gc[p].w, v4.w, acc.w);
        }
        dw[hr * 3 + kx] = acc;
      }
    }
    // note: pixels of the tile that lie outside the image have dh2 = 0 in the halo tile, so their gc is 0 and they add
    // nothing to dW; their dv is simply not stored
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int gx = tx0 + xs + p;
      if (gy >= H || gx >= W) continue;
      const size_t o = (((size_t)n * H + gy) * W + gx) * C + c0 + q * 4;
      // ReLU6 backward from the activation itself: v = clamp(u, 0, 6), so 0 < u < 6  <=>  0 < v < 6 (and then v == u)
      const float4 vc = *reinterpret_cast<const float4*>(tv + ((row + 1) * HS + xs + 1 + p) * CB + q * 4);
      const float vv[4] = {vc.x, vc.y, vc.z, vc.w};
      float d4[4] = {dv[p].x, dv[p].y, dv[p].z, dv[p].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) d4[j] = (vv[j] > 0.f && vv[j] < 6.f) ? d4[j] : 0.f;
      if (F32) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(du) + o) = make_float4(d4[0], d4[1], d4[2], d4[3]);
      } else {
        uint2 pk;
        pk.x = pack_bf16(d4[0], d4[1]); pk.y = pack_bf16(d4[2], d4[3]);
        *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(du) + o) = pk;
        // statistics of the STORED gradient (what the next pass reads)
        d4[0] = bf16lo(pk.x); d4[1] = bf16hi(pk.x); d4[2] = bf16lo(pk.y); d4[3] = bf16hi(pk.y);
      }
      s1.x += d4[0]; s1.y += d4[1]; s1.z += d4[2]; s1.w += d4[3];
      s2.x = fmaf(d4[0], vv[0], s2.x); s2.y = fmaf(d4[1], vv[1], s2.y); s2.z = fmaf(d4[2], vv[2], s2.z); s2.w = fmaf(d4[3], vv[3], s2.w);
    }
  }
  if (cur_n >= 0) flush(cur_n);
  // weight gradients: lanes with equal q, then the 8 warps; one atomic per (tap, channel)
  __syncthreads();
  float* red = tg;   // reuse: [8 warps][9][CB]
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    float a[4] = {dw[t].x, dw[t].y, dw[t].z, dw[t].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      a[j] += __shfl_xor_sync(0xffffffffu, a[j], 8);
      a[j] += __shfl_xor_sync(0xffffffffu, a[j], 16);
      if ((tid & 31) < 8) red[((tid >> 5) * 9 + t) * CB + q * 4 + j] = a[j];
    }
  }
  __syncthreads();
  for (int i = tid; i < 9 * CB; i += 256) {
    float s = 0.f;
#pragma unroll
    for (int wi = 0; wi < 8; ++wi) s += red[wi * 9 * CB + i];
    const int t = i / CB, c = i % CB;
    atomicAdd(dW + (size_t)(c0 + c) * 9 + t, s);   // depthwise.weight [C][1][3][3]
  }
}

void launch_dwconv_bwd(const void* dq, int dtg, const float4* coef_se, const void* h1, int dth, const float2* coef2, const float* w,
                       void* du, double* t12, float* dW, int N, int H, int W, int C, int num_sms, cudaStream_t st) {
  const int tilesX = (W + 15) / 16, tilesY = (H + 7) / 8;
  const long long total = (long long)N * tilesX * tilesY;
  const int cblocks = C / 32;
  long long bx = ((long long)num_sms * 8 + cblocks - 1) / cblocks;   // ~8 blocks per SM in total (4 resident)
  if (bx < 1) bx = 1;
  if (bx > total) bx = total;
  const int per = (int)((total + bx - 1) / bx);
  bx = (total + per - 1) / per;
  const size_t smem = (size_t)2 * 10 * 18 * 32 * sizeof(float);
  // static (3.2 KB) + dynamic (45 KB) shared memory together exceed the 48 KB default: opt in explicitly
  if (dtg == DT_F32) {
    if (ensure_dyn_smem_fn(dwconv_bwd_kernel<true>, 64 * 1024)) return;
    dwconv_bwd_kernel<true><<<dim3((unsigned)bx, cblocks), 256, smem, st>>>(dq, dtg, coef_se, h1, dth, coef2, w, du, t12, dW, N, H, W, C,
                                                                      tilesX, tilesY, per);
  } else {
    if (ensure_dyn_smem_fn(dwconv_bwd_kernel<false>, 64 * 1024)) return;
    dwconv_bwd_kernel<false><<<dim3((unsigned)bx, cblocks), 256, smem, st>>>(dq, dtg, coef_se, h1, dth, coef2, w, du, t12, dW, N, H, W, C,
                                                                       tilesX, tilesY, per);
  }
}

// =================================================================================================
// (6) weight gradient of a 1x1 conv / dense 3x3 conv as a split-M GEMM on the CUDA cores:
//       dW[n][k] += sum_m dY[m][n] * xform(A[m][k])
//     64(n) x 64(k) tile per block, the m range of a block is one of `splits` slices; fp32 atomics into the flat buffer.
struct WgradDst {       // where logical column k of the gradient matrix lives (1x1: per K-segment; conv: [Co][Ci][3][3])
  float* seg_ptr[LCM_MAX_SEGS];
  int seg_ld[LCM_MAX_SEGS];
  int seg_k0[LCM_MAX_SEGS + 1];
  int nseg;
  int conv_ci;        // > 0: conv layout, k = tap * Ci + ci -> seg_ptr[0][(n * Ci + ci) * 9 + tap]
  float* dbias;       // optional: sum_m dY[m][n]
};

struct WLoader1x1 {
  GemmSeg seg[LCM_MAX_SEGS];
  int dt[LCM_MAX_SEGS];
  int nseg, P;
  __device__ __forceinline__ void load4(long long m, int k, float (&v)[4]) const {
    int s = 0, koff = 0;
    while (s + 1 < nseg && k >= koff + seg[s].K) { koff += seg[s].K; ++s; }
    const GemmSeg& g = seg[s];
    const int kk = k - koff;
    const int img = (int)(m / P);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float x = ld1(g.A, dt[s], (size_t)m * g.ld + kk + j);
      if (g.mode != XF_NONE) x = xform(x, g.coef[(size_t)img * g.coef_ld + g.coef_off + kk + j], g.mode);
      v[j] = x;
    }
  }
};

struct WLoaderConv3 {
  const void* in;
  int dt, Hin, Win, Hout, Wout, Ci, mode;
  __device__ __forceinline__ void load4(long long m, int k, float (&v)[4]) const {
    const int tap = k / Ci, ci = k - tap * Ci;
    const int ky = tap / 3, kx = tap - ky * 3;
    const int x = (int)(m % Wout);
    const long long q = m / Wout;
    const int y = (int)(q % Hout);
    const long long n = q / Hout;
    const size_t base = (size_t)n * Hin * Win * Ci + ci;
    if (mode == CONV_UP2) {
      const int uy = y + ky - 1, ux = x + kx - 1;
      if (uy < 0 || uy >= Hout || ux < 0 || ux >= Wout) { v[0] = v[1] = v[2] = v[3] = 0.f; return; }
      const float sy = fmaxf(uy * 0.5f - 0.25f, 0.f), sx = fmaxf(ux * 0.5f - 0.25f, 0.f);
      const int y0 = (int)sy, x0 = (int)sx;
      const int y1 = min(y0 + 1, Hin - 1), x1 = min(x0 + 1, Win - 1);
      const float ly = sy - y0, lx = sx - x0;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        v[j] = (1.f - ly) * ((1.f - lx) * ld1(in, dt, base + ((size_t)y0 * Win + x0) * Ci + j) + lx * ld1(in, dt, base + ((size_t)y0 * Win + x1) * Ci + j)) +
               ly * ((1.f - lx) * ld1(in, dt, base + ((size_t)y1 * Win + x0) * Ci + j) + lx * ld1(in, dt, base + ((size_t)y1 * Win + x1) * Ci + j));
      return;
    }
    const int st = (mode == CONV_S2) ? 2 : 1;
    const int iy = y * st + ky - 1, ix = x * st + kx - 1;
    if (iy < 0 || iy >= Hin || ix < 0 || ix >= Win) { v[0] = v[1] = v[2] = v[3] = 0.f; return; }
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = ld1(in, dt, base + ((size_t)iy * Win + ix) * Ci + j);
  }
};

template <typename Loader>
__global__ void __launch_bounds__(256) wgrad_simt_kernel(Loader ld, const void* __restrict__ dY, int dty, WgradDst dst, long long M,
                                                         int Nc, int Ktot, long long rows_per_split) {
  __shared__ float Ys[16][68];   // [m][n]
  __shared__ float As[16][68];   // [m][k]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int n0 = blockIdx.x * 64, k0 = blockIdx.y * 64;
  const long long m_begin = (long long)blockIdx.z * rows_per_split;
  const long long m_end = m_begin + rows_per_split < M ? m_begin + rows_per_split : M;
  const int lrow = tid >> 4, lc = (tid & 15) * 4;   // loads: 16 rows x 16 groups of 4 columns
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
  const bool want_bias = dst.dbias != nullptr && blockIdx.y == 0;
  for (long long m0 = m_begin; m0 < m_end; m0 += 16) {
    float y4[4] = {0.f, 0.f, 0.f, 0.f}, a4[4] = {0.f, 0.f, 0.f, 0.f};
    const long long m = m0 + lrow;
    if (m < m_end) {
      if (n0 + lc < Nc) {
#pragma unroll
        for (int j = 0; j < 4; ++j) y4[j] = (n0 + lc + j < Nc) ? ld1(dY, dty, (size_t)m * Nc + n0 + lc + j) : 0.f;
      }
      if (k0 + lc < Ktot) ld.load4(m, k0 + lc, a4);   // Ktot, segment widths and Ci are multiples of 4
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) { Ys[lrow][lc + j] = y4[j]; As[lrow][lc + j] = a4[j]; }
    __syncthreads();
#pragma unroll
    for (int mm = 0; mm < 16; ++mm) {
      float y[4], a[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { y[i] = Ys[mm][ty * 4 + i]; a[i] = As[mm][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(y[i], a[j], acc[i][j]);
      if (want_bias && tx == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) bsum[i] += y[i];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= Nc) continue;
    if (want_bias && tx == 0) atomicAdd(dst.dbias + n, bsum[i]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k >= Ktot) continue;
      if (dst.conv_ci > 0) {
        const int tap = k / dst.conv_ci, ci = k - tap * dst.conv_ci;
        atomicAdd(dst.seg_ptr[0] + ((size_t)n * dst.conv_ci + ci) * 9 + tap, acc[i][j]);
      } else {
        int s = 0;
        while (s + 1 < dst.nseg && k >= dst.seg_k0[s + 1]) ++s;
        if (dst.seg_ptr[s]) atomicAdd(dst.seg_ptr[s] + (size_t)n * dst.seg_ld[s] + (k - dst.seg_k0[s]), acc[i][j]);
      }
    }
  }
}

static inline void wgrad_grid(long long M, int Nc, int Ktot, int num_sms, dim3& grid, long long& rows) {
  const int tn = (Nc + 63) / 64, tk = (Ktot + 63) / 64;
  long long splits = ((long long)num_sms * 4 + (long long)tn * tk - 1) / ((long long)tn * tk);
  if (splits < 1) splits = 1;
  rows = (M + splits - 1) / splits;
  rows = (rows + 15) / 16 * 16;
  if (rows < 16) rows = 16;
  splits = (M + rows - 1) / rows;
  if (splits > 65535) { splits = 65535; rows = ((M + splits - 1) / splits + 15) / 16 * 16; splits = (M + rows - 1) / rows; }
  grid = dim3(tn, tk, (unsigned)splits);
}

// 1x1: segments as in the forward GEMM (GemmParams::seg, element types in seg_dt); dY [M][Nc]; dst[s] = gradient tensor
// of segment s ([Nc][K_s], row stride dst_ld[s]) or null (identity segment).
void launch_wgrad_1x1(const GemmParams& p, const int* seg_dt, const void* dY, int dty, float* const* dst, const int* dst_ld,
                      int num_sms, cudaStream_t st) {
  WLoader1x1 ld;
  WgradDst d{};
  int k = 0;
  for (int i = 0; i < LCM_MAX_SEGS; ++i) {
    ld.seg[i] = p.seg[i];
    ld.dt[i] = i < p.nseg ? seg_dt[i] : 0;
    d.seg_ptr[i] = i < p.nseg ? dst[i] : nullptr;
    d.seg_ld[i] = i < p.nseg ? dst_ld[i] : 0;
    d.seg_k0[i] = k;
    if (i < p.nseg) k += p.seg[i].K;
  }
  d.seg_k0[LCM_MAX_SEGS] = k;
  for (int i = p.nseg; i <= LCM_MAX_SEGS; ++i) d.seg_k0[i] = k;
  ld.nseg = p.nseg; ld.P = p.P;
  d.nseg = p.nseg; d.conv_ci = 0; d.dbias = nullptr;
  dim3 grid; long long rows;
  wgrad_grid(p.M, p.Nc, p.Ktot, num_sms, grid, rows);
  wgrad_simt_kernel<WLoader1x1><<<grid, 256, 0, st>>>(ld, dY, dty, d, p.M, p.Nc, p.Ktot, rows);
}

// dense 3x3: in [N][Hin][Win][Ci] (mode as the forward conv), dY [N][Hout][Wout][Co]; dW [Co][Ci][3][3], dbias [Co]
void launch_wgrad_conv3(const void* in, int dti, const void* dY, int dty, float* dW, float* dbias, int N, int Hin, int Win, int Ci,
                        int Co, int mode, int num_sms, cudaStream_t st) {
  const int Hout = mode == CONV_S2 ? Hin / 2 : (mode == CONV_UP2 ? Hin * 2 : Hin);
  const int Wout = mode == CONV_S2 ? Win / 2 : (mode == CONV_UP2 ? Win * 2 : Win);
  WLoaderConv3 ld{in, dti, Hin, Win, Hout, Wout, Ci, mode};
  WgradDst d{};
  d.seg_ptr[0] = dW; d.nseg = 1; d.conv_ci = Ci; d.dbias = dbias;
  const long long M = (long long)N * Hout * Wout;
  dim3 grid; long long rows;
  wgrad_grid(M, Co, 9 * Ci, num_sms, grid, rows);
  wgrad_simt_kernel<WLoaderConv3><<<grid, 256, 0, st>>>(ld, dY, dty, d, M, Co, 9 * Ci, rows);
}

// bias gradient of a conv: out[c] += sum over rows of g[row][c]
__global__ void __launch_bounds__(256) colsum_kernel(const void* __restrict__ g, int dt, long long rows, int C, float* __restrict__ out) {
  __shared__ float red[256 * 8];
  const int cvecs = C / 8;
  RowGeom G(cvecs);
  const long long r0 = (long long)blockIdx.x * kRowChunk, r1 = r0 + kRowChunk < rows ? r0 + kRowChunk : rows;
  for (int cvb = 0; cvb < cvecs; cvb += G.cvp) {
    const int cv = cvb + G.cv;
    const bool act = G.active && cv < cvecs;
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = 0.f;
    if (act) {
      for (long long r = r0 + G.lane; r < r1; r += G.pl) {
        float v[8];
        ld8(g, dt, (size_t)r * C + cv * 8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] += v[j];
      }
    }
    const int ncv = min(G.cvp, cvecs - cvb);
    __syncthreads();
    if (G.active) {
#pragma unroll
      for (int j = 0; j < 8; ++j) red[(G.lane * G.cvp + G.cv) * 8 + j] = act ? s[j] : 0.f;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < ncv * 8; t += 256) {
      float a = 0.f;
      for (int l = 0; l < G.pl; ++l) a += red[l * G.cvp * 8 + t];
      atomicAdd(out + cvb * 8 + t, a);
    }
  }
}
void launch_colsum(const void* g, int dt, long long rows, int C, float* out, cudaStream_t st) {
  colsum_kernel<<<(unsigned)((rows + kRowChunk - 1) / kRowChunk), 256, 0, st>>>(g, dt, rows, C, out);
}

// =================================================================================================
// (7) bilinear x2 (align_corners=False): forward for any storage type (the fp32 plan materialises it only for training),
//     and its transpose.  Forward rows: out[2y] = .25 in[max(y-1,0)] + .75 in[y] ; out[2y+1] = .75 in[y] + .25 in[min(y+1,H-1)].
__global__ void __launch_bounds__(256) upsample2x_any_kernel(const void* in, void* out, int dt, int H, int W, int C, long long total) {
  const int cvecs = C >> 3;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int cv = (int)(i % cvecs);
    long long q = i / cvecs;
    const int X = (int)(q % (2 * W));
    q /= 2 * W;
    const int Y = (int)(q % (2 * H));
    const long long n = q / (2 * H);
    const int y = Y >> 1, x = X >> 1;
    const int ya = (Y & 1) ? y : max(y - 1, 0), yb = (Y & 1) ? min(y + 1, H - 1) : y;
    const float wya = (Y & 1) ? 0.75f : 0.25f, wyb = 1.f - wya;
    const int xa = (X & 1) ? x : max(x - 1, 0), xb = (X & 1) ? min(x + 1, W - 1) : x;
    const float wxa = (X & 1) ? 0.75f : 0.25f, wxb = 1.f - wxa;
    const size_t b = (size_t)n * H * W * C + cv * 8;
    float p00[8], p01[8], p10[8], p11[8], o[8];
    ld8(in, dt, b + ((size_t)ya * W + xa) * C, p00);
    ld8(in, dt, b + ((size_t)ya * W + xb) * C, p01);
    ld8(in, dt, b + ((size_t)yb * W + xa) * C, p10);
    ld8(in, dt, b + ((size_t)yb * W + xb) * C, p11);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = wya * (wxa * p00[j] + wxb * p01[j]) + wyb * (wxa * p10[j] + wxb * p11[j]);
    st8(out, dt, (((size_t)n * 2 * H + Y) * 2 * W + X) * C + cv * 8, o);
  }
}
void launch_upsample2x_any(const void* in, void* out, int dt, int N, int H, int W, int C, cudaStream_t st) {
  const long long total = (long long)N * 4 * H * W * (C / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 64) blocks = 148LL * 64;
  upsample2x_any_kernel<<<(int)blocks, 256, 0, st>>>(in, out, dt, H, W, C, total);
}

// zero insertion: out[n][2y][2x][:] = in[n][y][x][:], every other element 0 (16-bit tensors).  The input gradient of a
// stride-2 3x3 conv is the STRIDE-1 transposed conv of this tensor, which runs on the tensor cores (conv3x3_tc, halo mode).
__global__ void __launch_bounds__(256) zero_insert2x_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int H, int W, int cvecs,
                                                            long long total) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int cv = (int)(i % cvecs);
    long long q = i / cvecs;
    const int X = (int)(q % (2 * W));
    q /= 2 * W;
    const int Y = (int)(q % (2 * H));
    const long long n = q / (2 * H);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (!((X | Y) & 1)) v = in[(((size_t)n * H + (Y >> 1)) * W + (X >> 1)) * cvecs + cv];
    out[i] = v;
  }
}
void launch_zero_insert2x(const void* in, void* out, int N, int H, int W, int C, cudaStream_t st) {
  const long long total = (long long)N * 4 * H * W * (C / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 64) blocks = 148LL * 64;
  zero_insert2x_kernel<<<(int)blocks, 256, 0, st>>>(reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), H, W, C / 8, total);
}

// transpose: d in[y][x] = sum_{i,j} wy[i] wx[j] d out[Y_i][X_j],  Y_i in {2y-1, 2y, 2y+1, 2y+2}:
//   weights {.25 (y >= 1), .75 (+.25 at y = 0: the clamped tap of row 0), .75 (+.25 at y = H-1), .25 (y <= H-2)}
__global__ void __launch_bounds__(256) upsample2x_bwd_kernel(const void* dout, void* din, int dt, int H, int W, int C, int accumulate,
                                                             long long total) {
  const int cvecs = C >> 3;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int cv = (int)(i % cvecs);
    long long q = i / cvecs;
    const int x = (int)(q % W);
    q /= W;
    const int y = (int)(q % H);
    const long long n = q / H;
    float wy[4], wx[4];
    wy[0] = y >= 1 ? 0.25f : 0.f;
    wy[1] = 0.75f + (y == 0 ? 0.25f : 0.f);
    wy[2] = 0.75f + (y == H - 1 ? 0.25f : 0.f);
    wy[3] = y <= H - 2 ? 0.25f : 0.f;
    wx[0] = x >= 1 ? 0.25f : 0.f;
    wx[1] = 0.75f + (x == 0 ? 0.25f : 0.f);
    wx[2] = 0.75f + (x == W - 1 ? 0.25f : 0.f);
    wx[3] = x <= W - 2 ? 0.25f : 0.f;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int a = 0; a < 4; ++a) {
      if (wy[a] == 0.f) continue;
      const int Y = 2 * y - 1 + a;
      for (int b = 0; b < 4; ++b) {
        if (wx[b] == 0.f) continue;
        const int X = 2 * x - 1 + b;
        float v[8];
        ld8(dout, dt, (((size_t)n * 2 * H + Y) * 2 * W + X) * C + cv * 8, v);
        const float wgt = wy[a] * wx[b];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(wgt, v[j], acc[j]);
      }
    }
    const size_t o = (((size_t)n * H + y) * W + x) * C + cv * 8;
    if (accumulate) {
      float d[8];
      ld8(din, dt, o, d);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += d[j];
    }
    st8(din, dt, o, acc);
  }
}
void launch_upsample2x_bwd(const void* dout, void* din, int dt, int N, int H, int W, int C, int accumulate, cudaStream_t st) {
  const long long total = (long long)N * H * W * (C / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 64) blocks = 148LL * 64;
  upsample2x_bwd_kernel<<<(int)blocks, 256, 0, st>>>(dout, din, dt, H, W, C, accumulate, total);
}

// =================================================================================================
// (8) linear attention backward (efficient_unet.py:289-302), d = 32 per head.  qkv [N][P][3 inner] (q | k | v),
//     state [N][heads][32][33] fp64 (KV, column 32 = ksum) from the forward, dO [N][P][inner].
//   o = num / den, num = q' KV, den = q' . ksum + 1e-6, q' = phi(q), k' = phi(k), phi = elu + 1, phi' = (x > 0 ? 1 : phi)
//   pass 1 (per position): dnum = dO / den ; dden = -(dO . o) / den ; dq' = KV dnum + ksum dden ; dq = dq' phi'(q)
//                          dKV += q'^T dnum ; dksum += q' dden                    (fp64 atomics into dstate)
//   pass 2 (per position): dk' = dKV v + dksum ; dk = dk' phi'(k) ; dv = dKV^T k'
__device__ __forceinline__ float phi_f(float x) { return x > 0.f ? x + 1.f : expf(x); }

__global__ void __launch_bounds__(256) attn_bwd1_kernel(const void* __restrict__ qkv, int dtq, const double* __restrict__ state,
                                                        const void* __restrict__ dO, int dtg, void* __restrict__ dqkv,
                                                        double* __restrict__ dstate, int P, int heads) {
  __shared__ float kv[32][33];
  __shared__ float qs[64][33];     // phi(q)
  __shared__ float qr[64][33];     // raw q (for phi')
  __shared__ float gs[64][33];     // dO, then dnum
  __shared__ float s_den[64], s_dden[64];
  const int n = blockIdx.z, h = blockIdx.y;
  const int inner = heads * 32, ld = 3 * inner;
  const int tid = threadIdx.x;
  const double* s = state + ((size_t)n * heads + h) * 32 * 33;
  for (int i = tid; i < 32 * 33; i += 256) kv[i / 33][i % 33] = (float)s[i];
  const int d = tid >> 3, e0 = (tid & 7) * 4;    // accumulators: dKV[d][e0..e0+3], dksum[d] (e0 == 0)
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, accs = 0.f;
  const int pend = min(P, (int)(blockIdx.x + 1) * 256);
  for (int p0 = blockIdx.x * 256; p0 < pend; p0 += 64) {
    __syncthreads();
    for (int i = tid; i < 64 * 32; i += 256) {
      const int pp = i >> 5, dd = i & 31, p = p0 + pp;
      float q = 0.f, g = 0.f;
      if (p < P) {
        q = ld1(qkv, dtq, ((size_t)n * P + p) * ld + h * 32 + dd);
        g = ld1(dO, dtg, ((size_t)n * P + p) * inner + h * 32 + dd);
      }
      qr[pp][dd] = q;
      qs[pp][dd] = p < P ? phi_f(q) : 0.f;
      gs[pp][dd] = g;
    }
    __syncthreads();
    // per position: den, and dO . num  (4 threads per position, 8 e each)
    {
      const int pp = tid >> 2, part = tid & 3;
      float den = 0.f, gn = 0.f;
      if (part == 0) {
#pragma unroll
        for (int dd = 0; dd < 32; ++dd) den = fmaf(qs[pp][dd], kv[dd][32], den);
      }
      for (int e = part * 8; e < part * 8 + 8; ++e) {
        float num = 0.f;
#pragma unroll
        for (int dd = 0; dd < 32; ++dd) num = fmaf(qs[pp][dd], kv[dd][e], num);
        gn = fmaf(gs[pp][e], num, gn);
      }
      gn += __shfl_xor_sync(0xffffffffu, gn, 1);
      gn += __shfl_xor_sync(0xffffffffu, gn, 2);
      den = __shfl_sync(0xffffffffu, den, (tid & 31) & ~3);
      if (part == 0) {
        den += 1e-6f;
        s_den[pp] = den;
        s_dden[pp] = -gn / (den * den);    // dden = -(dO . o) / den with o = num / den
      }
    }
    __syncthreads();
    for (int i = tid; i < 64 * 32; i += 256) {   // dnum = dO / den
      const int pp = i >> 5, e = i & 31;
      gs[pp][e] = gs[pp][e] / s_den[pp];
    }
    __syncthreads();
    // dq[p][d] = (sum_e dnum[p][e] KV[d][e] + dden[p] ksum[d]) * phi'(q)
    for (int i = tid; i < 64 * 32; i += 256) {
      const int pp = i >> 5, dd = i & 31, p = p0 + pp;
      if (p >= P) continue;
      float a = s_dden[pp] * kv[dd][32];
#pragma unroll
      for (int e = 0; e < 32; ++e) a = fmaf(gs[pp][e], kv[dd][e], a);
      const float q = qr[pp][dd];
      a *= q > 0.f ? 1.f : qs[pp][dd];
      st1(dqkv, dtg, ((size_t)n * P + p) * ld + h * 32 + dd, a);
    }
    // dKV[d][e] += sum_p q'[p][d] dnum[p][e] ; dksum[d] += sum_p q'[p][d] dden[p]   (padding rows have q' = 0)
    for (int pp = 0; pp < 64; ++pp) {
      const float q = qs[pp][d];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = fmaf(q, gs[pp][e0 + j], acc[j]);
      if (e0 == 0) accs = fmaf(q, s_dden[pp], accs);
    }
  }
  double* ds = dstate + (((size_t)n * heads + h) * 32 + d) * 33;
#pragma unroll
  for (int j = 0; j < 4; ++j) atomicAdd(&ds[e0 + j], (double)acc[j]);
  if (e0 == 0) atomicAdd(&ds[32], (double)accs);
}

__global__ void __launch_bounds__(256) attn_bwd2_kernel(const void* __restrict__ qkv, int dtq, const double* __restrict__ dstate,
                                                        void* __restrict__ dqkv, int dtg, int P, int heads) {
  __shared__ float dkv[32][33];
  __shared__ float ks[64][33], kr[64][33], vs[64][33];
  const int n = blockIdx.z, h = blockIdx.y, p0 = blockIdx.x * 64;
  const int inner = heads * 32, ld = 3 * inner;
  const int tid = threadIdx.x;
  const double* s = dstate + ((size_t)n * heads + h) * 32 * 33;
  for (int i = tid; i < 32 * 33; i += 256) dkv[i / 33][i % 33] = (float)s[i];
  for (int i = tid; i < 64 * 32; i += 256) {
    const int pp = i >> 5, dd = i & 31, p = p0 + pp;
    float k = 0.f, v = 0.f;
    if (p < P) {
      k = ld1(qkv, dtq, ((size_t)n * P + p) * ld + inner + h * 32 + dd);
      v = ld1(qkv, dtq, ((size_t)n * P + p) * ld + 2 * inner + h * 32 + dd);
    }
    kr[pp][dd] = k;
    ks[pp][dd] = phi_f(k);
    vs[pp][dd] = v;
  }
  __syncthreads();
  for (int i = tid; i < 64 * 32; i += 256) {
    const int pp = i >> 5, dd = i & 31, p = p0 + pp;
    if (p >= P) continue;
    // dk[p][dd] = (sum_e v[p][e] dKV[dd][e] + dksum[dd]) * phi'(k) ; dv[p][dd] = sum_d k'[p][d] dKV[d][dd]
    float a = dkv[dd][32], b = 0.f;
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      a = fmaf(vs[pp][e], dkv[dd][e], a);
      b = fmaf(ks[pp][e], dkv[e][dd], b);
    }
    a *= kr[pp][dd] > 0.f ? 1.f : ks[pp][dd];
    st1(dqkv, dtg, ((size_t)n * P + p) * ld + inner + h * 32 + dd, a);
    st1(dqkv, dtg, ((size_t)n * P + p) * ld + 2 * inner + h * 32 + dd, b);
  }
}

void launch_attn_bwd(const void* qkv, int dtq, const double* state, const void* dO, int dtg, void* dqkv, double* dstate, int N, int P,
                     int heads, cudaStream_t st) {
  attn_bwd1_kernel<<<dim3((P + 255) / 256, heads, N), 256, 0, st>>>(qkv, dtq, state, dO, dtg, dqkv, dstate, P, heads);
  attn_bwd2_kernel<<<dim3((P + 63) / 64, heads, N), 256, 0, st>>>(qkv, dtq, dstate, dqkv, dtg, P, heads);
}

// =================================================================================================
// (9) time path backward (efficient_unet.py:60-76,412-417,189-192).
//   film: f[n][r] = b[r] + W[r][:] . st[n][:], st = silu(temb) ; d st[n][j] = sum_r df[n][r] W[r][j]
__global__ void __launch_bounds__(256) film_bwd_input_kernel(const float* __restrict__ dfilm, const float* __restrict__ W,
                                                             float* __restrict__ dst, int rows, int ted, int rows_per_block) {
  // grid (row slices, N): partial sums over a slice of rows, atomically added (dst zeroed with the backward scratch)
  extern __shared__ float s_df[];   // [rows_per_block]
  const int n = blockIdx.y;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  for (int r = r0 + threadIdx.x; r < r1; r += 256) s_df[r - r0] = dfilm[(size_t)n * rows + r];
  __syncthreads();
  for (int j = threadIdx.x; j < ted; j += 256) {
    float acc = 0.f;
    for (int r = r0; r < r1; ++r) acc = fmaf(s_df[r - r0], W[(size_t)r * ted + j], acc);
    atomicAdd(dst + (size_t)n * ted + j, acc);
  }
}
void launch_film_bwd_input(const float* dfilm, const float* W, float* dst, int N, int rows, int ted, cudaStream_t st) {
  const int rpb = 1024;
  film_bwd_input_kernel<<<dim3((rows + rpb - 1) / rpb, N), 256, rpb * sizeof(float), st>>>(dfilm, W, dst, rows, ted, rpb);
}

__device__ __forceinline__ float silu_grad(float x) {
  const float s = 1.f / (1.f + expf(-x));
  return s * (1.f + x * (1.f - s));
}
//   e = [cos | sin](t f) ; p1 = W1 e + b1 ; hm = silu(p1) ; temb = W3 hm + b3 ; st = silu(temb)
//   one block per image: recompute the forward, then d temb = d st silu'(temb), dW3 += d temb hm^T, db3 += d temb,
//   d hm = W3^T d temb, d p1 = d hm silu'(p1), dW1 += d p1 e^T, db1 += d p1
__global__ void __launch_bounds__(128) time_mlp_bwd_kernel(const long long* __restrict__ t_dev, long long t_scalar, int base, int ted,
                                                           const float* __restrict__ w1, const float* __restrict__ b1,
                                                           const float* __restrict__ w3, const float* __restrict__ b3,
                                                           const float* __restrict__ dst, float* __restrict__ dw1,
                                                           float* __restrict__ db1, float* __restrict__ dw3, float* __restrict__ db3) {
  extern __shared__ float sm[];
  float* emb = sm;            // [base]
  float* p1 = emb + base;     // [ted]
  float* hm = p1 + ted;       // [ted]
  float* dte = hm + ted;      // [ted]  d temb
  float* dp1 = dte + ted;     // [ted]
  const int n = blockIdx.x;
  const float t = (float)(t_dev ? t_dev[n] : t_scalar);
  const int half = base / 2;
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float f = expf((-9.210340371976184f * (float)i) / (float)half);
    const float a = t * f;
    emb[i] = cosf(a);
    emb[half + i] = sinf(a);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < ted; j += blockDim.x) {
    float acc = b1[j];
    for (int i = 0; i < base; ++i) acc = fmaf(w1[j * base + i], emb[i], acc);
    p1[j] = acc;
    hm[j] = acc / (1.f + expf(-acc));
  }
  __syncthreads();
  for (int j = threadIdx.x; j < ted; j += blockDim.x) {
    float acc = b3[j];
    for (int i = 0; i < ted; ++i) acc = fmaf(w3[j * ted + i], hm[i], acc);
    const float d = dst[(size_t)n * ted + j] * silu_grad(acc);
    dte[j] = d;
    atomicAdd(db3 + j, d);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ted * ted; i += blockDim.x) atomicAdd(dw3 + i, dte[i / ted] * hm[i % ted]);
  for (int i = threadIdx.x; i < ted; i += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < ted; ++j) acc = fmaf(w3[j * ted + i], dte[j], acc);
    const float d = acc * silu_grad(p1[i]);
    dp1[i] = d;
    atomicAdd(db1 + i, d);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ted * base; i += blockDim.x) atomicAdd(dw1 + i, dp1[i / base] * emb[i % base]);
}
void launch_time_mlp_bwd(const long long* t_dev, long long t_scalar, int N, int base, int ted, const float* w1, const float* b1,
                         const float* w3, const float* b3, const float* dst, float* dw1, float* db1, float* dw3, float* db3,
                         cudaStream_t st) {
  time_mlp_bwd_kernel<<<N, 128, (base + 4 * ted) * sizeof(float), st>>>(t_dev, t_scalar, base, ted, w1, b1, w3, b3, dst, dw1, db1,
                                                                       dw3, db3);
}

// =================================================================================================
// (10) loss (low_light_diffusion.py:269-273) and the first backward op: final_conv / SiLU backward.
//   loss types: 0 mse, 1 l1, 2 huber (delta = 1), all mean-reduced over N*Co*H*W.
__device__ __forceinline__ float loss_val(float d, int type) {
  if (type == 0) return d * d;
  const float a = fabsf(d);
  if (type == 1) return a;
  return a < 1.f ? 0.5f * d * d : a - 0.5f;
}
__device__ __forceinline__ float loss_grad(float d, int type) {   // d loss_val / d d
  if (type == 0) return 2.f * d;
  if (type == 1) return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
  return fabsf(d) < 1.f ? d : (d > 0.f ? 1.f : -1.f);
}
__global__ void __launch_bounds__(256) loss_kernel(const float* __restrict__ eps, const float* __restrict__ target, long long numel,
                                                   int type, double inv_numel, double* __restrict__ out) {
  __shared__ double s_part[8];
  double acc = 0.0;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < numel; i += (long long)gridDim.x * 256)
    acc += (double)loss_val(eps[i] - target[i], type);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += s_part[i];
    atomicAdd(out, s * inv_numel);
  }
}
void launch_loss(const float* eps, const float* target, long long numel, int type, double* out, cudaStream_t st) {
  cudaMemsetAsync(out, 0, sizeof(double), st);
  long long blocks = (numel + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  loss_kernel<<<(int)blocks, 256, 0, st>>>(eps, target, numel, type, 1.0 / (double)numel, out);
}

// final_conv backward (efficient_unet.py:600-602 reversed): eps = conv3x3(s) + bias, s = silu(a h + b).
//   dE[p][co] = scale * loss'(eps - target) / numel          (computed on the fly from the fp32 NCHW tensors)
//   ds[p][ci] = sum_{tap,co} dE[p - off(tap)][co] w[tap][ci][co] ; dpre = ds silu'(a h + b)  -> written as the gradient
//   T1 += sum_p dpre, T2 += sum_p dpre h ; dW[co][ci][tap] += sum_p dE[p][co] s[p + off(tap)][ci] ; dbias[co] += sum_p dE[p][co]
// One block = one image row band of 16x16 tiles (grid-stride over tiles), Ci <= 64, Co <= 4.
__global__ void __launch_bounds__(256) final_conv_bwd_kernel(const void* __restrict__ h, int dth, const float2* __restrict__ coef,
                                                             const float* __restrict__ w /*[9*Ci][Co]*/, const float* __restrict__ eps,
                                                             const float* __restrict__ target, int loss_type, float gscale,
                                                             const float* __restrict__ gscale_dev, void* __restrict__ dpre, int dtg,
                                                             double* __restrict__ t12, float* __restrict__ dW, float* __restrict__ dbias,
                                                             int N, int H, int W, int Ci, int Co, int tilesX, int tilesY) {
  constexpr int TS = 16, HS = TS + 2;
  extern __shared__ __align__(16) float fsm[];
  float* ts = fsm;                          // [HS*HS][Ci]  s = silu(pre), zero outside
  float* te = ts + HS * HS * Ci;            // [HS*HS][4]   dE, zero outside
  float* sw = te + HS * HS * 4;             // [9][Ci][4]
  float* s_t = sw + 9 * Ci * 4;             // [2][Ci] per-block T1/T2 partials
  const int tid = threadIdx.x;
  for (int i = tid; i < 9 * Ci * 4; i += 256) {
    const int co = i & 3, k = i >> 2;
    sw[i] = co < Co ? w[(size_t)k * Co + co] : 0.f;
  }
  // loss_type 3: `target` already holds the upstream gradient d loss / d eps (autograd through eps)
  const float sc_ext = gscale * (gscale_dev ? *gscale_dev : 1.f);
  const float sc = sc_ext / (float)((double)N * Co * H * W);
  const int tiles = tilesX * tilesY;
  // weight gradient: thread = (input channel ci, row group g); it owns dW[tap][ci][co] partial sums for ALL taps and
  // output channels (36 registers) over the rows g, g + G, ... of every tile: per pixel 3 new s values (the window slides
  // along x), one broadcast dE and 27 FMAs — the previous entry-per-thread loop issued two LDS per FMA (4.4 ms per step).
  const int wci = tid % Ci, wg = tid / Ci, wG = 256 / Ci;
  float wacc[9][4];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j) wacc[t][j] = 0.f;
  float bacc = 0.f;   // thread co < Co accumulates dbias[co] (tid < Co)
  for (long long it = blockIdx.x; it < (long long)N * tiles; it += gridDim.x) {
    const int n = (int)(it / tiles), tile = (int)(it % tiles);
    const int ty0 = (tile / tilesX) * TS, tx0 = (tile % tilesX) * TS;
    __syncthreads();
    for (int i = tid; i < 2 * Ci; i += 256) s_t[i] = 0.f;
    const int vecs = Ci / 8;
    for (int i = tid; i < HS * HS * vecs; i += 256) {
      const int px = i / vecs, cv = i - px * vecs;
      const int yy = px / HS, xx = px - yy * HS;
      const int gy = ty0 + yy - 1, gx = tx0 + xx - 1;
      float v[8];
      if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
        ld8(h, dth, (((size_t)n * H + gy) * W + gx) * Ci + cv * 8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = xform(v[j], coef[(size_t)n * Ci + cv * 8 + j], XF_AFFINE_SILU);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) ts[px * Ci + cv * 8 + j] = v[j];
    }
    for (int i = tid; i < HS * HS; i += 256) {
      const int yy = i / HS, xx = i - yy * HS;
      const int gy = ty0 + yy - 1, gx = tx0 + xx - 1;
      float e[4] = {0.f, 0.f, 0.f, 0.f};
      if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
        for (int co = 0; co < Co; ++co) {
          const size_t o = (((size_t)n * Co + co) * H + gy) * W + gx;
          e[co] = loss_type == 3 ? sc_ext * target[o] : sc * loss_grad(eps[o] - target[o], loss_type);
        }
      }
      *reinterpret_cast<float4*>(te + i * 4) = make_float4(e[0], e[1], e[2], e[3]);
    }
    __syncthreads();
    // ---- ds / dpre for this thread's pixel --------------------------------------------------------
    const int ly = tid >> 4, lx = tid & 15;
    const int gy = ty0 + ly, gx = tx0 + lx;
    const bool inside = gy < H && gx < W;
    {
      const size_t o = (((size_t)n * H + (inside ? gy : 0)) * W + (inside ? gx : 0)) * Ci;
      for (int c0 = 0; c0 < Ci; c0 += 8) {
        float ds[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) ds[j] = 0.f;
        for (int tap = 0; tap < 9 && inside; ++tap) {
          const int ky = tap / 3, kx = tap % 3;
          // forward: eps[q] += w[tap] s[q + off(tap)]  =>  ds[p] += w[tap] dE[p - off(tap)]
          const float4 e = *reinterpret_cast<const float4*>(te + ((ly + 2 - ky) * HS + lx + 2 - kx) * 4);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 wv = *reinterpret_cast<const float4*>(sw + (tap * Ci + c0 + j) * 4);
            ds[j] = fmaf(e.x, wv.x, fmaf(e.y, wv.y, fmaf(e.z, wv.z, fmaf(e.w, wv.w, ds[j]))));
          }
        }
        float hv[8];
        ld8(h, dth, o + c0, hv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 ab = coef[(size_t)n * Ci + c0 + j];
          ds[j] *= silu_grad(fmaf(ab.x, hv[j], ab.y));
        }
        if (inside) st8(dpre, dtg, o + c0, ds);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          // statistics of the STORED gradient; warp reduction first (one shared-memory atomic per warp and channel)
          float d = (dtg == DT_BF16) ? __bfloat162float(__float2bfloat16_rn(ds[j])) : ds[j];
          if (!inside) d = 0.f;
          float dh = d * hv[j];
          for (int sft = 16; sft > 0; sft >>= 1) { d += __shfl_xor_sync(0xffffffffu, d, sft); dh += __shfl_xor_sync(0xffffffffu, dh, sft); }
          if ((tid & 31) == 0) { atomicAdd(&s_t[c0 + j], d); atomicAdd(&s_t[Ci + c0 + j], dh); }
        }
      }
    }
    // ---- weight gradient (dE is zero outside the image, so out-of-image pixels of a ragged tile add nothing) ----
    if (wg < wG) {
      for (int py = wg; py < TS; py += wG) {
        float sv[3][3];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          sv[ky][1] = ts[((py + ky) * HS + 0) * Ci + wci];
          sv[ky][2] = ts[((py + ky) * HS + 1) * Ci + wci];
        }
#pragma unroll 4
        for (int pxx = 0; pxx < TS; ++pxx) {
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            sv[ky][0] = sv[ky][1]; sv[ky][1] = sv[ky][2];
            sv[ky][2] = ts[((py + ky) * HS + pxx + 2) * Ci + wci];
          }
          const float4 e = *reinterpret_cast<const float4*>(te + ((py + 1) * HS + pxx + 1) * 4);
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const float sx = sv[ky][kx];
              wacc[ky * 3 + kx][0] = fmaf(e.x, sx, wacc[ky * 3 + kx][0]);
              wacc[ky * 3 + kx][1] = fmaf(e.y, sx, wacc[ky * 3 + kx][1]);
              wacc[ky * 3 + kx][2] = fmaf(e.z, sx, wacc[ky * 3 + kx][2]);
              wacc[ky * 3 + kx][3] = fmaf(e.w, sx, wacc[ky * 3 + kx][3]);
            }
        }
      }
    }
    if (tid < Co) {
      float a = 0.f;
      for (int py = 0; py < TS && ty0 + py < H; ++py)
        for (int pxx = 0; pxx < TS && tx0 + pxx < W; ++pxx) a += te[((py + 1) * HS + pxx + 1) * 4 + tid];
      bacc += a;
    }
    __syncthreads();
    for (int i = tid; i < 2 * Ci; i += 256) {
      const int c = i % Ci, which = i / Ci;
      atomicAdd(t12 + ((size_t)n * Ci + c) * 2 + which, (double)s_t[i]);
    }
  }
  // row groups -> one partial per (tap, ci, co) and block (shared-memory adds), then one global atomic each
  __syncthreads();
  float* wred = ts;   // [9][Ci][4]
  for (int i = tid; i < 9 * Ci * 4; i += 256) wred[i] = 0.f;
  __syncthreads();
  if (wg < wG) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(&wred[(t * Ci + wci) * 4 + j], wacc[t][j]);
  }
  __syncthreads();
  for (int i = tid; i < 9 * Ci * 4; i += 256) {
    const int co = i & 3, k = i >> 2;
    const int ci = k % Ci, tap = k / Ci;
    if (co < Co) atomicAdd(dW + ((size_t)co * Ci + ci) * 9 + tap, wred[i]);   // final_conv.weight [Co][Ci][3][3]
  }
  if (tid < Co) atomicAdd(dbias + tid, bacc);
}

int launch_final_conv_bwd(const void* h, int dth, const float2* coef, const float* w, const float* eps, const float* target,
                          int loss_type, float gscale, const float* gscale_dev, void* dpre, int dtg, double* t12, float* dW,
                          float* dbias, int N, int H, int W, int Ci, int Co, int num_sms, cudaStream_t st) {
  if (Ci > 64 || Ci % 8 || Co > 4) return 1;
  const int tilesX = (W + 15) / 16, tilesY = (H + 15) / 16;
  const size_t smem = ((size_t)18 * 18 * (Ci + 4) + 9 * Ci * 4 + 2 * Ci) * sizeof(float);
  if (ensure_dyn_smem_fn(final_conv_bwd_kernel, smem)) return 1;
  long long blocks = (long long)N * tilesX * tilesY;
  if (blocks > (long long)num_sms * 2) blocks = (long long)num_sms * 2;
  final_conv_bwd_kernel<<<(int)blocks, 256, smem, st>>>(h, dth, coef, w, eps, target, loss_type, gscale, gscale_dev, dpre, dtg, t12,
                                                       dW, dbias, N, H, W, Ci, Co, tilesX, tilesY);
  return 0;
}

// init_conv weight gradient (efficient_unet.py:553 reversed; the input needs no gradient):
//   dW[co][ci][tap] += sum_p dY[p][co] x[ci][p + off(tap)] ; dbias[co] += sum_p dY[p][co],  x = cat(xa, xb) fp32 NCHW
__global__ void __launch_bounds__(256) init_conv_wgrad_kernel(const float* __restrict__ xa, int ca, long long sa,
                                                              const float* __restrict__ xb, int cb, long long sb,
                                                              const void* __restrict__ dY, int dtg, float* __restrict__ dW,
                                                              float* __restrict__ dbias, int N, int H, int W, int Co, int tilesX,
                                                              int tilesY) {
  constexpr int TS = 16, HS = TS + 2;
  extern __shared__ __align__(16) float ism[];
  const int Cin = ca + cb;
  float* tx = ism;                      // [Cin][HS*HS]
  float* ty = tx + Cin * HS * HS;       // [TS*TS][Co]
  const int tid = threadIdx.x;
  const int nent = 9 * Cin * Co;        // entry e = (tap * Cin + ci) * Co + co
  float wacc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) wacc[j] = 0.f;
  float bacc = 0.f;
  const int tiles = tilesX * tilesY;
  for (long long it = blockIdx.x; it < (long long)N * tiles; it += gridDim.x) {
    const int n = (int)(it / tiles), tile = (int)(it % tiles);
    const int ty0 = (tile / tilesX) * TS, tx0 = (tile % tilesX) * TS;
    __syncthreads();
    for (int i = tid; i < Cin * HS * HS; i += 256) {
      const int ci = i / (HS * HS), px = i % (HS * HS);
      const int yy = px / HS, xx = px % HS;
      const int gy = ty0 + yy - 1, gx = tx0 + xx - 1;
      float v = 0.f;
      if (gy >= 0 && gy < H && gx >= 0 && gx < W)
        v = ci < ca ? xa[n * sa + ((long long)ci * H + gy) * W + gx] : xb[n * sb + ((long long)(ci - ca) * H + gy) * W + gx];
      tx[i] = v;
    }
    for (int i = tid; i < TS * TS * (Co / 8); i += 256) {
      const int px = i / (Co / 8), cv = i % (Co / 8);
      const int gy = ty0 + px / TS, gx = tx0 + px % TS;
      float v[8];
      if (gy < H && gx < W) ld8(dY, dtg, (((size_t)n * H + gy) * W + gx) * Co + cv * 8, v);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) ty[px * Co + cv * 8 + j] = v[j];
    }
    __syncthreads();
#pragma unroll
    for (int slot = 0; slot < 16; ++slot) {
      const int ent = tid + slot * 256;
      if (ent >= nent) continue;
      const int co = ent % Co, k = ent / Co;
      const int ci = k % Cin, tap = k / Cin;
      const int ky = tap / 3, kx = tap % 3;
      float a = 0.f;
      for (int py = 0; py < TS; ++py)
        for (int pxx = 0; pxx < TS; ++pxx)
          a = fmaf(ty[(py * TS + pxx) * Co + co], tx[ci * HS * HS + (py + ky) * HS + pxx + kx], a);
      wacc[slot] += a;
    }
    if (tid < Co) {
      float a = 0.f;
      for (int px = 0; px < TS * TS; ++px) a += ty[px * Co + tid];
      bacc += a;
    }
  }
#pragma unroll
  for (int slot = 0; slot < 16; ++slot) {
    const int ent = tid + slot * 256;
    if (ent >= nent) continue;
    const int co = ent % Co, k = ent / Co;
    const int ci = k % Cin, tap = k / Cin;
    atomicAdd(dW + ((size_t)co * Cin + ci) * 9 + tap, wacc[slot]);   // init_conv.weight [Co][Cin][3][3]
  }
  if (tid < Co) atomicAdd(dbias + tid, bacc);
}

// register-blocked version for the model's shape (CIN = 6): thread = (output channel co, row group g) owns all 9 x CIN
// partial sums of its co (54 registers); per pixel one dY value and, per input channel, 3 new broadcast x values (the 3x3
// window slides along x) feed 9 CIN FMAs — the entry-per-thread kernel above issues two LDS per FMA (2.0 ms per step).
template <int CIN>
__global__ void __launch_bounds__(256) init_conv_wgrad_rb_kernel(const float* __restrict__ xa, int ca, long long sa,
                                                                 const float* __restrict__ xb, long long sb,
                                                                 const void* __restrict__ dY, int dtg, float* __restrict__ dW,
                                                                 float* __restrict__ dbias, int N, int H, int W, int Co, int tilesX,
                                                                 int tilesY) {
  constexpr int TS = 16, HS = TS + 2;
  extern __shared__ __align__(16) float ism[];
  float* tx = ism;                      // [CIN][HS*HS]
  float* ty = tx + CIN * HS * HS;       // [TS*TS][Co]
  const int tid = threadIdx.x;
  const int wco = tid % Co, wg = tid / Co, wG = 256 / Co;
  float wacc[CIN][9];
#pragma unroll
  for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
    for (int t = 0; t < 9; ++t) wacc[ci][t] = 0.f;
  float bacc = 0.f;
  const int tiles = tilesX * tilesY;
  for (long long it = blockIdx.x; it < (long long)N * tiles; it += gridDim.x) {
    const int n = (int)(it / tiles), tile = (int)(it % tiles);
    const int ty0 = (tile / tilesX) * TS, tx0 = (tile % tilesX) * TS;
    __syncthreads();
    for (int i = tid; i < CIN * HS * HS; i += 256) {
      const int ci = i / (HS * HS), px = i % (HS * HS);
      const int yy = px / HS, xx = px % HS;
      const int gy = ty0 + yy - 1, gx = tx0 + xx - 1;
      float v = 0.f;
      if (gy >= 0 && gy < H && gx >= 0 && gx < W)
        v = ci < ca ? xa[n * sa + ((long long)ci * H + gy) * W + gx] : xb[n * sb + ((long long)(ci - ca) * H + gy) * W + gx];
      tx[i] = v;
    }
    for (int i = tid; i < TS * TS * (Co / 8); i += 256) {
      const int px = i / (Co / 8), cv = i % (Co / 8);
      const int gy = ty0 + px / TS, gx = tx0 + px % TS;
      float v[8];
      if (gy < H && gx < W) ld8(dY, dtg, (((size_t)n * H + gy) * W + gx) * Co + cv * 8, v);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) ty[px * Co + cv * 8 + j] = v[j];
    }
    __syncthreads();
    if (wg < wG) {
      for (int py = wg; py < TS; py += wG) {
        float sv[CIN][3][3];
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            sv[ci][ky][1] = tx[ci * HS * HS + (py + ky) * HS + 0];
            sv[ci][ky][2] = tx[ci * HS * HS + (py + ky) * HS + 1];
          }
#pragma unroll 2
        for (int pxx = 0; pxx < TS; ++pxx) {
          const float g = ty[(py * TS + pxx) * Co + wco];     // 0 beyond a ragged edge
          bacc += g;
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              sv[ci][ky][0] = sv[ci][ky][1]; sv[ci][ky][1] = sv[ci][ky][2];
              sv[ci][ky][2] = tx[ci * HS * HS + (py + ky) * HS + pxx + 2];
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) wacc[ci][ky * 3 + kx] = fmaf(g, sv[ci][ky][kx], wacc[ci][ky * 3 + kx]);
            }
        }
      }
    }
  }
  __syncthreads();
  float* wred = tx;     // [Co][CIN][9] + [Co]
  for (int i = tid; i < Co * CIN * 9 + Co; i += 256) wred[i] = 0.f;
  __syncthreads();
  if (wg < wG) {
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
      for (int t = 0; t < 9; ++t) atomicAdd(&wred[(wco * CIN + ci) * 9 + t], wacc[ci][t]);
    atomicAdd(&wred[Co * CIN * 9 + wco], bacc);
  }
  __syncthreads();
  for (int i = tid; i < Co * CIN * 9; i += 256) atomicAdd(dW + i, wred[i]);   // init_conv.weight [Co][CIN][3][3]
  for (int i = tid; i < Co; i += 256) atomicAdd(dbias + i, wred[Co * CIN * 9 + i]);
}

int launch_init_conv_wgrad(const float* xa, int ca, long long sa, const float* xb, int cb, long long sb, const void* dY, int dtg,
                           float* dW, float* dbias, int N, int H, int W, int Co, int num_sms, cudaStream_t st) {
  const int Cin = ca + cb;
  const int tilesX = (W + 15) / 16, tilesY = (H + 15) / 16;
  const size_t smem = ((size_t)Cin * 18 * 18 + 256 * Co) * sizeof(float);
  if (Cin == 6 && Co % 8 == 0 && Co <= 256 && (size_t)(Co * 6 * 9 + Co) <= (size_t)6 * 18 * 18) {
    if (ensure_dyn_smem_fn(init_conv_wgrad_rb_kernel<6>, smem)) return 1;
    long long blocks = (long long)N * tilesX * tilesY;
    if (blocks > (long long)num_sms * 2) blocks = (long long)num_sms * 2;
    init_conv_wgrad_rb_kernel<6><<<(int)blocks, 256, smem, st>>>(xa, ca, sa, xb, sb, dY, dtg, dW, dbias, N, H, W, Co, tilesX, tilesY);
    return 0;
  }
  if (9 * Cin * Co > 16 * 256 || Co % 8) return 1;   // 16 register slots of 256 threads hold every (tap, ci, co) entry
  if (ensure_dyn_smem_fn(init_conv_wgrad_kernel, smem)) return 1;
  long long blocks = (long long)N * tilesX * tilesY;
  if (blocks > (long long)num_sms * 2) blocks = (long long)num_sms * 2;
  init_conv_wgrad_kernel<<<(int)blocks, 256, smem, st>>>(xa, ca, sa, xb, cb, sb, dY, dtg, dW, dbias, N, H, W, Co, tilesX, tilesY);
  return 0;
}

// =================================================================================================
// (11) optimizer: clip_grad_norm_(max_norm) + AdamW + EMA in one pass over flat fp32 buffers
//      (trainer.py:296-302 -> torch.nn.utils.clip_grad_norm_, torch.optim.AdamW defaults; EMA :98-104).
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, double* __restrict__ out) {
  __shared__ double s_part[8];
  double acc = 0.0;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) acc += (double)g[i] * (double)g[i];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += s_part[i];
    atomicAdd(out, s);
  }
}
void launch_sumsq(const float* g, long long n, double* out, cudaStream_t st) {
  cudaMemsetAsync(out, 0, sizeof(double), st);
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  sumsq_kernel<<<(int)blocks, 256, 0, st>>>(g, n, out);
}

// grad_div: gradients are divided by it first (world size when the all-reduce summed them; the GradScaler's scale).
// clip: coef = min(1, max_norm / (||g|| + 1e-6)) with ||g|| = sqrt(*sumsq) / grad_div ; max_norm <= 0 disables clipping.
__global__ void __launch_bounds__(256) adamw_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, float* __restrict__ ema, long long n, float lr,
                                                        float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt,
                                                        float ema_decay, const double* __restrict__ sumsq, float grad_div,
                                                        float max_norm) {
  float coef = 1.f / grad_div;
  if (max_norm > 0.f && sumsq) {
    const float norm = (float)sqrt(*sumsq) / grad_div;
    const float c = max_norm / (norm + 1e-6f);
    if (c < 1.f) coef *= c;
  }
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float gi = g[i] * coef;
    float pi = p[i];
    pi *= 1.f - lr * wd;                             // decoupled weight decay (torch.optim.AdamW)
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= (lr / bc1) * (mi / denom);
    p[i] = pi;
    if (ema) ema[i] = ema_decay * ema[i] + (1.f - ema_decay) * pi;
  }
}
void launch_adamw_ema(float* p, const float* g, float* m, float* v, float* ema, long long n, float lr, float beta1, float beta2,
                      float eps, float wd, int step, float ema_decay, const double* sumsq, float grad_div, float max_norm,
                      cudaStream_t st) {
  const float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  adamw_ema_kernel<<<(int)blocks, 256, 0, st>>>(p, g, m, v, ema, n, lr, beta1, beta2, eps, wd, bc1, bc2_sqrt, ema_decay, sumsq,
                                                grad_div, max_norm);
}

}  // namespace lcm
