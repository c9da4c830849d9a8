// Small kernels of the hot path: time embedding, FiLM, GroupNorm finalise, SE gate, scheduler step,
// weight packing, layout conversion.  None of these moves a full activation tensor; they exist so
// that the heavy kernels' prologues are a single per-(image, channel) affine.
#include <cuda_fp16.h>

#include <cstdlib>

#include "kernels.h"

namespace lcm {

// ------------------------------------------------------------------------------------------------
// SinusoidalPosEmb -> Linear -> SiLU -> Linear  (efficient_unet.py:68-76, 412-417).
// One block per sample.  freqs follow the reference's fp32 op order:
//   (float(-ln 1e4) * float(i)) / float(half) -> expf ; args = float(t) * freq ; [cos | sin].
__global__ void time_embed_kernel(const long long* __restrict__ t_dev, long long t_scalar, int base, int ted,
                                  const float* __restrict__ w1, const float* __restrict__ b1,
                                  const float* __restrict__ w3, const float* __restrict__ b3,
                                  float* __restrict__ temb, float* __restrict__ silu_temb) {
  extern __shared__ float sm[];
  float* emb = sm;          // [base]
  float* hid = sm + base;   // [ted]
  const int n = blockIdx.x;
  const float t = (float)(t_dev ? t_dev[n] : t_scalar);
  const int half = base / 2;
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    float f = expf((-9.210340371976184f * (float)i) / (float)half);
    float a = t * f;
    emb[i] = cosf(a);
    emb[half + i] = sinf(a);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < ted; j += blockDim.x) {
    float acc = b1[j];
    for (int i = 0; i < base; ++i) acc = fmaf(w1[j * base + i], emb[i], acc);
    hid[j] = acc / (1.f + expf(-acc));
  }
  __syncthreads();
  for (int j = threadIdx.x; j < ted; j += blockDim.x) {
    float acc = b3[j];
    for (int i = 0; i < ted; ++i) acc = fmaf(w3[j * ted + i], hid[i], acc);
    if (temb) temb[(size_t)n * ted + j] = acc;
    silu_temb[(size_t)n * ted + j] = acc / (1.f + expf(-acc));
  }
}

void launch_time_embed(const long long* t_dev, long long t_scalar, int N, int base, int ted, const float* w1,
                       const float* b1, const float* w3, const float* b3, float* temb, float* silu_temb,
                       cudaStream_t st) {
  time_embed_kernel<<<N, 128, (base + ted) * sizeof(float), st>>>(t_dev, t_scalar, base, ted, w1, b1, w3, b3, temb,
                                                                  silu_temb);
}

// ------------------------------------------------------------------------------------------------
// All blocks' FiLM projections (efficient_unet.py:189-192,215) as one small GEMM out[n][r] = b[r] + W[r][:] . s[n][:]
// (rows = 2 x sum of hidden widths ~ 30 000, ted = 128, N = 64).  Block = 64 rows x up to 64 images: the activations
// (transposed, [k][image]) and the weight tile are staged in shared memory per 128-wide k slab; a warp owns 8 rows,
// lane = image (and image + 32), four rows at a time: 4 broadcast + 2 strided LDS feed 8 FMAs, no shuffles.
__global__ void __launch_bounds__(256) film_kernel(const float* __restrict__ s, const float* __restrict__ W, const float* __restrict__ b,
                                                   float* __restrict__ out, int N, int rows, int ted) {
  pdl_wait();
  pdl_trigger();
  constexpr int KS = 64;                   // k slab
  __shared__ float st[KS][64 + 1];         // [k][image]; reused as the [image][row] output tile
  __shared__ float ws[64][KS + 4];         // [row][k]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  float acc[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = 0.f;
  for (int k0 = 0; k0 < ted; k0 += KS) {
    const int kw = min(KS, ted - k0);
    __syncthreads();
    for (int i = tid; i < 64 * KS; i += 256) {
      const int n = i / KS, k = i % KS;
      st[k][n] = (n0 + n < N && k < kw) ? s[(size_t)(n0 + n) * ted + k0 + k] : 0.f;
      ws[n][k] = (r0 + n < rows && k < kw) ? W[(size_t)(r0 + n) * ted + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const float* w0 = ws[warp * 8 + half * 4];
#pragma unroll 4
      for (int k = 0; k < KS; ++k) {
        const float x0 = st[k][lane], x1 = st[k][lane + 32];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float wv = w0[i * (KS + 4) + k];
          acc[half * 4 + i][0] = fmaf(wv, x0, acc[half * 4 + i][0]);
          acc[half * 4 + i][1] = fmaf(wv, x1, acc[half * 4 + i][1]);
        }
      }
    }
  }
  __syncthreads();
  float (*ot)[64 + 1] = st;                // [image][row] (64 x 65 floats fit in st)
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = warp * 8 + i;
    const float bias = r0 + r < rows ? b[r0 + r] : 0.f;
    ot[lane][r] = acc[i][0] + bias;
    ot[lane + 32][r] = acc[i][1] + bias;
  }
  __syncthreads();
  for (int i = tid; i < 64 * 64; i += 256) {
    const int n = i >> 6, r = i & 63;
    if (n0 + n < N && r0 + r < rows) out[(size_t)(n0 + n) * rows + r0 + r] = ot[n][r];
  }
}

void launch_film(const float* silu_temb, const float* W, const float* b, float* out, int N, int rows, int ted,
                 cudaStream_t st) {
  launch_pdl(film_kernel, dim3((rows + 63) / 64, (N + 63) / 64), dim3(256), 0, st, silu_temb, W, b, out, N, rows, ted);
}

// ------------------------------------------------------------------------------------------------
// GroupNorm finalise (efficient_unet.py:170-171,207,212 ; F.group_norm semantics: biased variance,
// eps = 1e-5 inside the sqrt).  One block per image.
__global__ void __launch_bounds__(1024) gn_coef_kernel(const double* __restrict__ s0, int C0, const double* __restrict__ s1, int C1,
                               int groups, double count, const float* __restrict__ gamma,
                               const float* __restrict__ beta, const float* __restrict__ film, int film_ld,
                               float2* __restrict__ coef) {
  __shared__ float s_mean[64], s_rstd[64];
  const int n = blockIdx.x;
  const int C = C0 + C1;
  // gamma / beta / FiLM do not depend on the producer of the statistics: start pulling them in before the dependency
  // wait, so that their (cold) latency overlaps the statistics loads instead of following them
#pragma unroll 1
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    asm volatile("prefetch.global.L1 [%0];" ::"l"(gamma + c));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(beta + c));
    if (film) {
      asm volatile("prefetch.global.L1 [%0];" ::"l"(film + (size_t)n * film_ld + c));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(film + (size_t)n * film_ld + C + c));
    }
  }
  pdl_wait();
  pdl_trigger();
  const int cpg = C / groups;
  const double inv_count = 1.0 / count;   // one fp64 division per thread
  // one warp per group (fixed summation order: lanes stride over the group's channels, then a shuffle tree)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll 1
  for (int g = warp; g < groups; g += (int)(blockDim.x >> 5)) {
    double sum = 0.0, sq = 0.0;
#pragma unroll 1
    for (int c = g * cpg + lane; c < (g + 1) * cpg; c += 32) {
      const double2 v = *reinterpret_cast<const double2*>((c < C0) ? s0 + ((size_t)n * C0 + c) * 2 : s1 + ((size_t)n * C1 + (c - C0)) * 2);
      sum += v.x;
      sq += v.y;
    }
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, o); sq += __shfl_xor_sync(0xffffffffu, sq, o); }
    // sums in fp64 (cancellation in E[x^2] - mean^2), the rest in fp32: fp64 division and sqrt are long software
    // routines, and this kernel runs cold 47 times per forward — its cost is its code footprint
    const double mean = sum * inv_count;
    double var = sq * inv_count - mean * mean;
    if (var < 0.0) var = 0.0;
    if (lane == 0) {
      s_mean[g] = (float)mean;
      s_rstd[g] = rsqrtf((float)var + 1e-5f);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    float a = gamma[c] * s_rstd[g];
    float b = beta[c] - s_mean[g] * a;
    if (film) {
      const float sc = 1.f + film[(size_t)n * film_ld + c];
      const float sh = film[(size_t)n * film_ld + C + c];
      a *= sc;
      b = fmaf(b, sc, sh);
    }
    coef[(size_t)n * C + c] = make_float2(a, b);
  }
}

void launch_gn_coef(const double* stats0, int C0, const double* stats1, int C1, int groups, double count,
                    const float* gamma, const float* beta, const float* film, int film_ld, float2* coef, int N,
                    cudaStream_t st) {
  // one warp per group in a single pass (32 groups -> 1024 threads): the group loop is a chain of dependent L2 loads and
  // fp64 shuffles, and this kernel sits on the critical path 47 times per forward
  static int thr = -1;
  if (thr < 0) { const char* e = getenv("LCM_GN_THREADS"); thr = e ? atoi(e) : 1024; if (thr < 32 || thr > 1024 || thr % 32) thr = 256; }
  int block = groups * 32 < thr ? groups * 32 : thr;
  if (block < 64) block = 64;
  launch_pdl(gn_coef_kernel, dim3(N), dim3(block), 0, st, stats0, C0, stats1, C1, groups, count, gamma, beta, film, film_ld, coef);
}

// ------------------------------------------------------------------------------------------------
// Squeeze-and-Excitation gate (efficient_unet.py:96-100) as two batched row-GEMVs in ONE launch:
//   hid[n][j]  = relu6(b1[j] + sum_c w1[j][c] * mean[n][c])          (mean = pooled sum * inv_count)
//   gate[n][c] = sigmoid(b2[c] + sum_j w2[c][j] * hid[n][j])         -> prologue coefficient (gate, 0)
// The gate sits on the serial chain depthwise -> gate -> project of every block, so what it costs is latency, not
// bandwidth.  Measured inside the graph (skip / no-op experiments, profiles/r01_micro_kernels.txt): two dependent
// launches cost ~32 us per block; one launch with a fully unrolled 74 KB body still ~20 us although it executes only
// ~6000 warp instructions per CTA — between two uses GBs of activations pass through L2 and the instruction caches,
// so a cold kernel pays for its code FOOTPRINT.  Hence: one launch, and a deliberately compact body (one shared,
// non-inlined row routine, modest unrolling).
// A thread-block CLUSTER of 8 CTAs owns 4 images: each CTA stages the pooled vectors, computes 1/8 of the hidden rows,
// the slices are exchanged through distributed shared memory, then each CTA computes 1/8 of the gate rows.
// A weight row is split over just enough lanes (float4 per lane), several rows per warp when rows are short.
constexpr int SE_IMG = 4, SE_CL = 8, SE_THREADS = 512;

// out_mode 0: relu6 -> dst_s[i * dst_ld + (r - r0)] (shared);  1: sigmoid -> dst_g[(n0 + i) * dst_ld + r] = (gate, 0)
__device__ __noinline__ void se_fc_rows(const float* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ xs,
                                        int K, int r0, int r1, int out_mode, float* dst_s, float2* dst_g, int dst_ld, int nimg) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = SE_THREADS / 32;
  const bool vec = (K & 3) == 0;
  const int units = vec ? K >> 2 : K;          // float4 units (or scalars when K is not a multiple of 4)
  int lpr = 32;
  while (lpr > 1 && (lpr >> 1) >= units) lpr >>= 1;
  const int rpp = 32 / lpr, sub = lane / lpr, sl = lane - sub * lpr;
  constexpr int RB = 2;   // row groups per pass: their weight loads are in flight together
  for (int rb = r0 + warp * rpp * RB; rb < r1; rb += nw * rpp * RB) {
    float acc[RB][SE_IMG];
#pragma unroll
    for (int q = 0; q < RB; ++q)
#pragma unroll
      for (int i = 0; i < SE_IMG; ++i) acc[q][i] = 0.f;
    if (vec) {
#pragma unroll 2
      for (int u = sl; u < units; u += lpr) {
        float4 w[RB];
#pragma unroll
        for (int q = 0; q < RB; ++q) {
          const int r = rb + q * rpp + sub;
          w[q] = r < r1 ? __ldg(reinterpret_cast<const float4*>(W + (size_t)r * K) + u) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < SE_IMG; ++i) {
          const float4 x = *reinterpret_cast<const float4*>(xs + i * K + u * 4);
#pragma unroll
          for (int q = 0; q < RB; ++q) {
            acc[q][i] = fmaf(w[q].x, x.x, acc[q][i]); acc[q][i] = fmaf(w[q].y, x.y, acc[q][i]);
            acc[q][i] = fmaf(w[q].z, x.z, acc[q][i]); acc[q][i] = fmaf(w[q].w, x.w, acc[q][i]);
          }
        }
      }
    } else {
#pragma unroll 1
      for (int u = sl; u < units; u += lpr) {
#pragma unroll
        for (int q = 0; q < RB; ++q) {
          const int r = rb + q * rpp + sub;
          const float w = r < r1 ? __ldg(W + (size_t)r * K + u) : 0.f;
#pragma unroll
          for (int i = 0; i < SE_IMG; ++i) acc[q][i] = fmaf(w, xs[i * K + u], acc[q][i]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < RB; ++q) {
      for (int o = lpr >> 1; o > 0; o >>= 1) {
#pragma unroll
        for (int i = 0; i < SE_IMG; ++i) acc[q][i] += __shfl_xor_sync(0xffffffffu, acc[q][i], o);
      }
      const int r = rb + q * rpp + sub;
      if (r < r1 && sl == 0) {
        const float b = bias[r];
#pragma unroll
        for (int i = 0; i < SE_IMG; ++i) {
          const float v = acc[q][i] + b;
          if (out_mode == 0) dst_s[i * dst_ld + (r - r0)] = fminf(fmaxf(v, 0.f), 6.f);
          else if (i < nimg) dst_g[(size_t)i * dst_ld + r] = make_float2(1.f / (1.f + expf(-v)), 0.f);
        }
      }
    }
  }
}

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem(const float* local, uint32_t rank) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(local), ra;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
  return v;
}

__global__ void __cluster_dims__(SE_CL, 1, 1) __launch_bounds__(SE_THREADS)
se_gate_kernel(const double* __restrict__ pool, float inv_count, const float* __restrict__ w1, const float* __restrict__ b1,
               const float* __restrict__ w2, const float* __restrict__ b2, float2* __restrict__ coef, int N, int C, int SQ, int dbg) {
  extern __shared__ __align__(16) float se_sm[];
  if (dbg == 1) { pdl_wait(); pdl_trigger(); return; }   // LCM_SE_DEBUG: launch-only timing experiment
  const int sqr = (SQ + SE_CL - 1) / SE_CL;         // hidden rows per CTA
  float* xs = se_sm;                                // [SE_IMG][C] pooled means
  float* hid = xs + SE_IMG * C;                     // [SE_IMG][SQ] all hidden rows (after the exchange)
  float* mine = hid + SE_IMG * SQ;                  // [SE_IMG][sqr] this CTA's slice
  const uint32_t rank = cluster_ctarank();
  const int n0 = (blockIdx.x / SE_CL) * SE_IMG, nimg = min(SE_IMG, N - n0);
  const int h0 = (int)rank * sqr, h1 = min(SQ, h0 + sqr);
  const int cr = (C + SE_CL - 1) / SE_CL;
  const int c0 = (int)rank * cr, c1 = min(C, c0 + cr);
  // the FC weights do not depend on the previous kernel: pull this CTA's rows towards L2 before the dependency wait
  {
    const char* a = reinterpret_cast<const char*>(w1 + (size_t)h0 * C);
    const long long na = (long long)max(h1 - h0, 0) * C * 4;
#pragma unroll 1
    for (long long o = (long long)threadIdx.x * 128; o < na; o += SE_THREADS * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + o));
    const char* b = reinterpret_cast<const char*>(w2 + (size_t)c0 * SQ);
    const long long nb = (long long)max(c1 - c0, 0) * SQ * 4;
#pragma unroll 1
    for (long long o = (long long)threadIdx.x * 128; o < nb; o += SE_THREADS * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(b + o));
  }
  pdl_wait();
  pdl_trigger();
#pragma unroll 2
  for (int idx = threadIdx.x; idx < SE_IMG * C; idx += SE_THREADS)
    xs[idx] = idx < nimg * C ? (float)pool[(size_t)n0 * C + idx] * inv_count : 0.f;
  __syncthreads();
  se_fc_rows(w1, b1, xs, C, h0, h1, 0, mine, nullptr, sqr, nimg);
  cluster_sync_all();
#pragma unroll 1
  for (int idx = threadIdx.x; idx < SE_IMG * SQ; idx += SE_THREADS) {
    const int i = idx / SQ, j = idx - i * SQ;
    const int src = j / sqr;
    hid[idx] = ld_dsmem(mine + i * sqr + (j - src * sqr), (uint32_t)src);
  }
  cluster_sync_all();   // every peer has finished reading this CTA's slice (and hid is complete CTA-wide)
  se_fc_rows(w2, b2, hid, SQ, c0, c1, 1, nullptr, coef + (size_t)n0 * C, C, nimg);
}

int launch_se_gate(const double* pool, float inv_count, const float* w1, const float* b1, const float* w2,
                   const float* b2, float2* coef, int N, int C, int SQ, cudaStream_t st) {
  static int dbg = -1;
  if (dbg < 0) { const char* e = getenv("LCM_SE_DEBUG"); dbg = e ? atoi(e) : 0; }
  const int sqr = (SQ + SE_CL - 1) / SE_CL;
  const size_t sm = ((size_t)SE_IMG * C + (size_t)SE_IMG * SQ + (size_t)SE_IMG * sqr) * sizeof(float);
  if (sm > 200 * 1024) return 1;
  if (ensure_dyn_smem_fn(se_gate_kernel, sm)) return 1;
  const int groups = (N + SE_IMG - 1) / SE_IMG;
  launch_pdl(se_gate_kernel, dim3(groups * SE_CL), dim3(SE_THREADS), sm, st, pool, inv_count, w1, b1, w2, b2, coef, N, C, SQ, dbg);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// LCMScheduler.step (lcm_scheduler.py:214-242) and add_noise/get_velocity (:255-305), stand-alone.
__global__ void lcm_step_kernel(const float* __restrict__ eps, const float* __restrict__ sample,
                                const float* __restrict__ noise, float* __restrict__ prev, float* __restrict__ x0o,
                                long long numel, int prediction, float sb_t, float sa_t, float sa_p, float sb_p) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < numel;
       i += (long long)gridDim.x * blockDim.x) {
    const float e = eps[i], x = sample[i];
    // same fp32 operation order as the reference (mul, sub, div / mul, mul, add), no FMA contraction
    float x0 = prediction == 0 ? __fdiv_rn(__fsub_rn(x, __fmul_rn(sb_t, e)), sa_t)
                               : __fsub_rn(__fmul_rn(sa_t, x), __fmul_rn(sb_t, e));
    if (x0o) x0o[i] = x0;
    prev[i] = noise ? __fadd_rn(__fmul_rn(sa_p, x0), __fmul_rn(sb_p, noise[i])) : x0;
  }
}

void launch_lcm_step(const float* eps, const float* sample, const float* noise, float* prev, float* x0,
                     long long numel, int prediction, float sb_t, float sa_t, float sa_p, float sb_p,
                     cudaStream_t st) {
  int blocks = (int)((numel + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  lcm_step_kernel<<<blocks, 256, 0, st>>>(eps, sample, noise, prev, x0, numel, prediction, sb_t, sa_t, sa_p, sb_p);
}

__global__ void lcm_mix_kernel(const float* __restrict__ a, const float* __restrict__ b,
                               const long long* __restrict__ t, const float* __restrict__ abar,
                               float* __restrict__ out, long long per_sample, int velocity) {
  const int n = blockIdx.y;
  const float ab = abar[t[n]];
  const float sa = sqrtf(ab), sb = sqrtf(1.f - ab);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_sample;
       i += (long long)gridDim.x * blockDim.x) {
    const long long j = (long long)n * per_sample + i;
    out[j] = velocity ? sa * b[j] - sb * a[j] : sa * a[j] + sb * b[j];
  }
}

void launch_lcm_mix(const float* a, const float* b, const long long* t, const float* abar, float* out, int batch,
                    long long per_sample, int velocity, cudaStream_t st) {
  int bx = (int)((per_sample + 255) / 256);
  if (bx > 1024) bx = 1024;
  lcm_mix_kernel<<<dim3(bx, batch), 256, 0, st>>>(a, b, t, abar, out, per_sample, velocity);
}

// ------------------------------------------------------------------------------------------------
// condition_mode="add" (low_light_diffusion.py:108-113,158-160,223-225): model_input = latents + condition_encoder(low_light)
__global__ void add_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = a[i] + b[i];
}
void launch_add_f32(const float* a, const float* b, float* out, long long n, cudaStream_t st) {
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  add_f32_kernel<<<(int)blocks, 256, 0, st>>>(a, b, out, n);
}
__global__ void fill_float2_kernel(float2* __restrict__ p, float2 v, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}
void launch_fill_float2(float2* p, float2 v, long long n, cudaStream_t st) {
  fill_float2_kernel<<<(int)((n + 255) / 256 > 1024 ? 1024 : (n + 255) / 256), 256, 0, st>>>(p, v, n);
}

// ------------------------------------------------------------------------------------------------
// Consistency distillation (low_light_diffusion.py:325-408), the element-wise parts around the three UNet forwards.
// DDIM step of the teacher (:372-381): x0 = (x_t - sqrt(1 - a_t) eps) / sqrt(a_t) ; x_next = sqrt(a_n) x0 + sqrt(1 - a_n) eps,
// per-sample timesteps t, t_next.
__global__ void ddim_step_kernel(const float* __restrict__ x_t, const float* __restrict__ eps, const long long* __restrict__ t,
                                 const long long* __restrict__ t_next, const float* __restrict__ abar, float* __restrict__ x_next,
                                 long long per_sample) {
  const int n = blockIdx.y;
  const float at = abar[t[n]], an = abar[t_next[n]];
  const float sa = sqrtf(at), sb = sqrtf(1.f - at), na = sqrtf(an), nb = sqrtf(1.f - an);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_sample; i += (long long)gridDim.x * blockDim.x) {
    const long long j = (long long)n * per_sample + i;
    const float e = eps[j];
    const float x0 = (x_t[j] - sb * e) / sa;
    x_next[j] = na * x0 + nb * e;
  }
}
void launch_ddim_step(const float* x_t, const float* eps, const long long* t, const long long* t_next, const float* abar,
                      float* x_next, int batch, long long per_sample, cudaStream_t st) {
  int bx = (int)((per_sample + 255) / 256);
  if (bx > 1024) bx = 1024;
  ddim_step_kernel<<<dim3(bx, batch), 256, 0, st>>>(x_t, eps, t, t_next, abar, x_next, per_sample);
}
// loss = huber(student_x0, target_x0) (mean, delta 1; :399-406) and d loss / d eps_student:
//   student_x0 = (x_t - sqrt(1 - a_t) eps_s) / sqrt(a_t) ; target_x0 = (x_next - sqrt(1 - a_n) eps_tgt) / sqrt(a_n)
__global__ void consistency_loss_kernel(const float* __restrict__ x_t, const float* __restrict__ eps_s, const long long* __restrict__ t,
                                        const float* __restrict__ x_next, const float* __restrict__ eps_tgt,
                                        const long long* __restrict__ t_next, const float* __restrict__ abar, double inv_numel,
                                        double* __restrict__ loss, float* __restrict__ d_eps, long long per_sample) {
  __shared__ double s_part[8];
  const int n = blockIdx.y;
  const float at = abar[t[n]], an = abar[t_next[n]];
  const float sa = sqrtf(at), sb = sqrtf(1.f - at), na = sqrtf(an), nb = sqrtf(1.f - an);
  double acc = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_sample; i += (long long)gridDim.x * blockDim.x) {
    const long long j = (long long)n * per_sample + i;
    const float xs = (x_t[j] - sb * eps_s[j]) / sa;
    const float xt = (x_next[j] - nb * eps_tgt[j]) / na;
    const float d = xs - xt, a = fabsf(d);
    acc += a < 1.f ? 0.5 * (double)d * d : (double)a - 0.5;
    const float g = a < 1.f ? d : (d > 0.f ? 1.f : -1.f);
    d_eps[j] = g * (-sb / sa) * (float)inv_numel;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += s_part[i];
    atomicAdd(loss, s * inv_numel);
  }
}
void launch_consistency_loss(const float* x_t, const float* eps_s, const long long* t, const float* x_next, const float* eps_tgt,
                             const long long* t_next, const float* abar, double* loss, float* d_eps, int batch, long long per_sample,
                             cudaStream_t st) {
  cudaMemsetAsync(loss, 0, sizeof(double), st);
  int bx = (int)((per_sample + 255) / 256);
  if (bx > 256) bx = 256;
  consistency_loss_kernel<<<dim3(bx, batch), 256, 0, st>>>(x_t, eps_s, t, x_next, eps_tgt, t_next, abar,
                                                          1.0 / ((double)batch * (double)per_sample), loss, d_eps, per_sample);
}

// ------------------------------------------------------------------------------------------------
// Image formats either side of the path (scripts/inference.py:111-116, 121-127): byte work, HBM-bound.
// One thread per pixel: 3 interleaved bytes <-> one element of each of the three planes (plane accesses coalesced,
// the 3-byte pixel accesses are consecutive across the warp: 96 contiguous bytes).
__global__ void image_pre_u8_kernel(const uint8_t* __restrict__ hwc, float* __restrict__ nchw, long long hw, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / hw, p = i - n * hw;
    const uint8_t* src = hwc + i * 3;
    float* dst = nchw + n * 3 * hw + p;
#pragma unroll
    for (int c = 0; c < 3; ++c) dst[c * hw] = __fsub_rn(__fdiv_rn((float)src[c], 127.5f), 1.0f);   // numpy: x.astype(f32) / 127.5 - 1.0
  }
}
__global__ void image_post_u8_kernel(const float* __restrict__ nchw, uint8_t* __restrict__ hwc, long long hw, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / hw, p = i - n * hw;
    const float* src = nchw + n * 3 * hw + p;
    uint8_t* dst = hwc + i * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = __fmul_rn(__fadd_rn(src[c * hw], 1.0f), 127.5f);   // numpy: (y + 1.0) * 127.5
      dst[c] = (uint8_t)(int)fminf(fmaxf(v, 0.f), 255.f);                 // np.clip(., 0, 255).astype(np.uint8): truncation
    }
  }
}
// cv2.resize INTER_LINEAR, CV_8UC3 (OpenCV resize.cpp: HResizeLinear<uchar,int,short,2048> + VResizeLinear with
// FixedPtCast<int,uchar,22>): coefficient tables rebuilt per thread with the same float / double operations.
__device__ __forceinline__ void resize_tap(int d, double scale, int src, bool is_x, int& s, int& w0, int& w1) {
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  s = (int)floorf(f);
  f -= (float)s;
  if (is_x) {                       // x: out-of-range taps are folded into the border pixel with weight 1
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= src - 1) { f = 0.f; s = src - 1; }
  }
  w0 = __float2int_rn((1.f - f) * 2048.f);   // saturate_cast<short>(cvRound(.)): round half to even, |.| <= 2048
  w1 = __float2int_rn(f * 2048.f);
}
__global__ void image_resize_u8_kernel(const uint8_t* __restrict__ src, int sh, int sw, uint8_t* __restrict__ dst, int dh, int dw,
                                       double scale_x, double scale_y, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int dx = (int)(i % dw);
    const long long q = i / dw;
    const int dy = (int)(q % dh), n = (int)(q / dh);
    int sx, ax0, ax1, sy, by0, by1;
    resize_tap(dx, scale_x, sw, true, sx, ax0, ax1);
    resize_tap(dy, scale_y, sh, false, sy, by0, by1);
    const int x1 = min(sx + 1, sw - 1);
    const int y0 = min(max(sy, 0), sh - 1), y1 = min(max(sy + 1, 0), sh - 1);   // y: rows clamped, weights kept
    const uint8_t* r0 = src + ((long long)n * sh + y0) * sw * 3;
    const uint8_t* r1 = src + ((long long)n * sh + y1) * sw * 3;
    uint8_t* o = dst + i * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int h0 = (int)r0[sx * 3 + c] * ax0 + (int)r0[x1 * 3 + c] * ax1;
      const int h1 = (int)r1[sx * 3 + c] * ax0 + (int)r1[x1 * 3 + c] * ax1;
      const int v = (((by0 * (h0 >> 4)) >> 16) + ((by1 * (h1 >> 4)) >> 16) + 2) >> 2;
      o[c] = (uint8_t)min(max(v, 0), 255);
    }
  }
}
void launch_image_resize_u8(const uint8_t* src, int N, int sh, int sw, uint8_t* dst, int dh, int dw, cudaStream_t st) {
  const long long total = (long long)N * dh * dw;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  // OpenCV: inv_scale = dsize / ssize (double), scale = 1. / inv_scale
  const double scale_x = 1.0 / ((double)dw / (double)sw), scale_y = 1.0 / ((double)dh / (double)sh);
  image_resize_u8_kernel<<<blocks, 256, 0, st>>>(src, sh, sw, dst, dh, dw, scale_x, scale_y, total);
}
void launch_image_pre_u8(const uint8_t* hwc, float* nchw, int N, int H, int W, cudaStream_t st) {
  const long long hw = (long long)H * W, total = hw * N;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  image_pre_u8_kernel<<<blocks, 256, 0, st>>>(hwc, nchw, hw, total);
}
void launch_image_post_u8(const float* nchw, uint8_t* hwc, int N, int H, int W, cudaStream_t st) {
  const long long hw = (long long)H * W, total = hw * N;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  image_post_u8_kernel<<<blocks, 256, 0, st>>>(nchw, hwc, hw, total);
}

// ------------------------------------------------------------------------------------------------
// Weight packing: fp32 state_dict tensors -> the layouts the kernels read.
// destination element type of a logical-matrix job: PackJob::bf16 = 0 fp32, 1 bf16, 2 fp16
__device__ __forceinline__ void put_any(const PackJob& j, int n, int k, float v) {
  if (j.scale != 0.f) v *= j.scale;
  long long o = (j.layout == WL_UMMA) ? umma_weight_offset(n, k, j.ld, j.block_n) : (long long)n * j.ld + k;
  if (j.bf16 == 1) reinterpret_cast<bf16*>(j.dst)[o] = __float2bfloat16_rn(v);
  else if (j.bf16 == 2) reinterpret_cast<__half*>(j.dst)[o] = __float2half_rn(v);
  else reinterpret_cast<float*>(j.dst)[o] = v;
}

__global__ void pack_kernel(PackJob j, const float* __restrict__ src, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    switch (j.kind) {
      case PACK_COPY:
        reinterpret_cast<float*>(j.dst)[i] = src[i];
        break;
      case PACK_MAT: {
        int r = (int)(i / j.Cc), c = (int)(i % j.Cc);
        const float v = src[(size_t)r * j.src_ld + j.src_col0 + c];
        put_any(j, r, j.off + c, v);
        break;
      }
      case PACK_MAT_T: {   // logical (r, c) = src[c][src_col0 + r]
        int r = (int)(i / j.Cc), c = (int)(i % j.Cc);
        const float v = src[(size_t)c * j.src_ld + j.src_col0 + r];
        put_any(j, r, j.off + c, v);
        break;
      }
      case PACK_CONV3_T: {  // src [Co][Ci][3][3]; R = Ci rows, Cc = Co
        int tap = (int)(i % 9);
        long long q = i / 9;
        int ci = (int)(q % j.R), co = (int)(q / j.R);
        int k = j.off + (8 - tap) * j.tap_stride + co;
        put_any(j, ci, k, src[i]);
        break;
      }
      case PACK_IDENTITY: {
        int r = (int)(i / j.Cc), c = (int)(i % j.Cc);
        float v = (r == c) ? 1.f : 0.f;
        put_any(j, r, j.off + c, v);
        break;
      }
      case PACK_CONV3: {  // src [Co][Ci][3][3]
        int tap = (int)(i % 9);
        long long q = i / 9;
        int ci = (int)(q % j.Ci), co = (int)(q / j.Ci);
        int k = j.off + tap * j.tap_stride + ci;
        put_any(j, co, k, src[i]);
        break;
      }
      case PACK_CONV3_KN: {
        int tap = (int)(i % 9);
        long long q = i / 9;
        int ci = (int)(q % j.Ci), co = (int)(q / j.Ci);
        reinterpret_cast<float*>(j.dst)[(size_t)(tap * j.Ci + ci) * j.R + co] = src[i];
        break;
      }
      case PACK_DW: {  // src [C][1][3][3]
        int tap = (int)(i % 9), c = (int)(i / 9);
        reinterpret_cast<float*>(j.dst)[(size_t)tap * j.R + c] = src[i];
        break;
      }
    }
  }
}

void launch_pack(const PackJob& job, const float* src, cudaStream_t st) {
  long long total = 0;
  switch (job.kind) {
    case PACK_COPY: total = (long long)job.R * job.Cc; break;
    case PACK_MAT: case PACK_IDENTITY: case PACK_MAT_T: total = (long long)job.R * job.Cc; break;
    case PACK_CONV3: case PACK_CONV3_KN: total = (long long)job.R * job.Ci * 9; break;
    case PACK_CONV3_T: total = (long long)job.R * job.Cc * 9; break;
    case PACK_DW: total = (long long)job.R * 9; break;
  }
  if (total == 0) return;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  pack_kernel<<<blocks, 256, 0, st>>>(job, src, total);
}

// ------------------------------------------------------------------------------------------------
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ in, float* __restrict__ out, int HW, int C, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long n = i / ((long long)HW * C);
    long long r = i % ((long long)HW * C);
    int c = (int)(r / HW), p = (int)(r % HW);
    out[i] = to_f<T>(in[(n * HW + p) * C + c]);
  }
}

void launch_nhwc_to_nchw(const void* in, float* out, int N, int H, int W, int C, int bf16act, cudaStream_t st) {
  long long total = (long long)N * H * W * C;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 32) blocks = 148 * 32;
  if (bf16act == 2) nhwc_to_nchw_kernel<__half><<<blocks, 256, 0, st>>>((const __half*)in, out, H * W, C, total);
  else if (bf16act) nhwc_to_nchw_kernel<bf16><<<blocks, 256, 0, st>>>((const bf16*)in, out, H * W, C, total);
  else nhwc_to_nchw_kernel<float><<<blocks, 256, 0, st>>>((const float*)in, out, H * W, C, total);
}

}  // namespace lcm
