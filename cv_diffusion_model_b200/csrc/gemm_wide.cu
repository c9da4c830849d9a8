// Expand GEMM of the wide blocks (InvertedResidualBlock.expand, efficient_unet.py:207-209, for 128 <= K <= 448):
//   h1[M][N] (fp16) = relu6(a x + b)[M][K] * W[N][K]^T,  N = 4 K,  + per-(image, channel) sum / sum^2 of h1.
//
// Why a third GEMM kernel.  gemm_expand.cu keeps ALL weights in shared memory (K <= 128, N <= 512).  Beyond that the
// general kernel (gemm_tc2.cu) streams a [128 x 64] activation chunk AND a [256 x 64] weight chunk per MMA group and
// re-transforms the activation tile once per n-tile (N / 256 = 3..6 times).  Measured at K = 384, N = 1536: 2080
// cycles per chunk against 512 cycles of MMA — 1400 of them the L2 -> shared-memory supply (48 KB per chunk per SM is
// ~34 B/clk/SM = 10 TB/s over the chip), the rest the repeated prologue.  Here the ACTIVATION tile is stationary:
//   * the 128-pixel tile (all K, <= 112 KB) is loaded and transformed ONCE and multiplied by every n-block of the
//     weights, which stream through a ring of [128 x 64] chunks (16 KB) out of L2 — per 128x128x64 MMA group the SM
//     pulls 16 KB instead of 24 KB and never repeats a prologue;
//   * a chunk slot of the tile is released by the MMAs of the LAST n-block, one by one, so the next tile's chunks
//     land and are transformed while the current tile is still being multiplied;
//   * statistics come from the epilogue (an input-side Gram matrix would be K x K > TMEM here): after the per-warp
//     transpose a lane holds 8 columns x 8 rows; sums are reduced over rows by shuffles, over the 4 row-quadrant warps
//     through a small shared-memory scratch in FIXED order, and accumulated per image by an exclusive owner thread —
//     no floating-point atomics inside the CTA, so the result is bitwise reproducible; fp64 atomics (a few fp32
//     partials per entry) merge the CTAs that share an image.
//
// Roles (512 threads = 16 warps): warps 0-3 / 4-7 E0 / E1 (each drains one 64-column half of every 128-column
// accumulator), warp 8 MMA issuer, warp 9 TMA (activations), warp 10 bulk copies (weights), warps 12-15 prologue.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "kernels.h"
#include "tc_common.cuh"
#include "tmap.h"

namespace lcm {

namespace {

using namespace tc;

constexpr int kThreadsW = 512;
constexpr int kWXfBase = 384, kWXfThreads = 128;
constexpr int kWMmaWarp = 8, kWTmaAWarp = 9, kWTmaBWarp = 10, kWXfWarp0 = 12;   // warp 11 idles
constexpr uint32_t kWChunk = 16384;         // 128 rows x 64 bf16 (activations) or 128 output channels x 64 k (weights)
constexpr uint32_t kWSmemLimit = 232448;
constexpr int kWMaxChunks = 8, kWMaxNB = 16, kWMaxStages = 8;
constexpr uint32_t kWPatchBytes = 32768;    // 8 epilogue warps x 4 KB transpose patch
constexpr uint32_t kWPartBytes = 8192;      // [group 2][buffer 2][warp 4][64 columns][sum, sum^2] fp32

struct WideParams {
  CUtensorMap tmap_in[2];
  const float2* coef[2];
  int coef_ld[2], coef_off[2], segK[2];
  int nseg, nchunks, NB, stages, aslots;   // stages: weight ring; aslots >= nchunks: activation chunk ring
  const bf16* W;            // packed [n-block][chunk][128 rows x 64 k] (128-byte swizzled rows), scaled by 6
  __half* out;
  double* stats;            // [images][Nc][2]
  int m_tiles, P, Nc;
  uint32_t chunk[kWMaxChunks];   // segment | (first channel / 8) << 8 | (coefficient base / 8) << 16
  uint32_t b_off, stg_off, part_off, stat_off, coef_smem_off, misc_off;
  int debug;   // LCM_W_DEBUG bit mask, timing experiments only (results are wrong): 1 no statistics, 2 no stores, 4 tiny weight copies, 8 no prologue
};

__device__ __forceinline__ void tma_load_2d_w(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}

__device__ long long g_wprof[8];   // LCM_W_DEBUG & 16: cycles the MMA warp of block 0 spent { waiting for an accumulator, for the
                                   // tile's chunk, for weights, issuing, total }
__global__ void __launch_bounds__(kThreadsW, 1) gemm_wide_kernel(const __grid_constant__ WideParams p) {
  extern __shared__ uint8_t wsm_raw[];
  const uint32_t sraw = smem_u32(wsm_raw);
  const uint32_t sbase = (sraw + 1023u) & ~1023u;
  uint8_t* smem = wsm_raw + (sbase - sraw);
  // the warp index is broadcast: the role branches below are then warp-uniform for the compiler, and the single-warp
  // roles keep their loop state in uniform registers (the MMA issue loop cost ~10 R2UR per step without this)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;

  const uint32_t bar0 = sbase + p.misc_off;
  auto a_raw = [&](int c) { return bar0 + 8u * c; };            // activation slot c: chunk has landed
  auto a_xf = [&](int c) { return bar0 + 8u * (16 + c); };      // ... and is transformed
  auto a_empty = [&](int c) { return bar0 + 8u * (32 + c); };   // ... and the last n-block of its tile has read it
  auto b_full = [&](int s) { return bar0 + 8u * (48 + s); };
  auto b_empty = [&](int s) { return bar0 + 8u * (56 + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (64 + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (66 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + p.misc_off + 640);
  float2* s_coef = reinterpret_cast<float2*>(smem + p.coef_smem_off);
  float* s_stat = reinterpret_cast<float*>(smem + p.stat_off);     // [Nc][2], entry owned by one epilogue thread
  float* s_part = reinterpret_cast<float*>(smem + p.part_off);

  if (warp == kWTmaAWarp && lane == 0) {
    for (int c = 0; c < p.aslots; ++c) { mbar_init(a_raw(c), 1); mbar_init(a_xf(c), kWXfThreads); mbar_init(a_empty(c), 1); }
    for (int s = 0; s < p.stages; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int s = 0; s < p.nseg; ++s)
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmap_in[s])) : "memory");
  }
  if (warp == kWMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  for (int i = tid; i < p.Nc * 2; i += kThreadsW) s_stat[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  const int t_begin = (int)((long long)p.m_tiles * blockIdx.x / gridDim.x);
  const int t_end = (int)((long long)p.m_tiles * (blockIdx.x + 1) / gridDim.x);
  const int tiles_per_img = p.P >> 7;

  if (warp >= kWXfWarp0) {
    // ================================ prologue: relu6(a x + b) / 6 in place, once per tile ==================
    const int xt = tid - kWXfBase;
    // unit u = xt + 128 i: row = u >> 3, slot = u & 7 holds channel unit cu = slot ^ (row & 7); (row & 7) does not
    // depend on i (128 = 16 rows), so a thread transforms the same 8 channels in all of its rows
    const int cu = (xt & 7) ^ ((xt >> 3) & 7);
    int img = t_begin / tiles_per_img, tin = t_begin - img * tiles_per_img;
    int coef_img = -1;
    int slot = 0; uint32_t sphase = 0;
    for (int t = t_begin; t < t_end; ++t) {
      if (img != coef_img) {
        bar_sync(1, kWXfThreads);   // everyone is done with the previous image's coefficients
        for (int s = 0, base = 0; s < p.nseg; base += p.segK[s], ++s) {
          const float2* src = p.coef[s] + (size_t)img * p.coef_ld[s] + p.coef_off[s];
          for (int k = xt; k < p.segK[s]; k += kWXfThreads) {
            const float2 c = src[k];
            s_coef[base + k] = make_float2(c.x * (1.f / 6.f), c.y * (1.f / 6.f));
          }
        }
        bar_sync(1, kWXfThreads);
        coef_img = img;
      }
      for (int ci = 0; ci < p.nchunks; ++ci) {
        const uint32_t a_smem = sbase + (uint32_t)slot * kWChunk;
        float2 ab[8];
        {
          const float4* c4 = reinterpret_cast<const float4*>(s_coef + ((p.chunk[ci] >> 16) & 0xff) * 8 + cu * 8);
#pragma unroll
          for (int j = 0; j < 4; ++j) { const float4 c = c4[j]; ab[2 * j] = make_float2(c.x, c.y); ab[2 * j + 1] = make_float2(c.z, c.w); }
        }
        mbar_wait(a_raw(slot), sphase);
        if (!(p.debug & 8)) {
        uint4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = lds128(a_smem + (uint32_t)(xt + i * kWXfThreads) * 16u);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float f[8];
          unpack8(v[i], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = __saturatef(fmaf(ab[j].x, f[j], ab[j].y));
          sts128(a_smem + (uint32_t)(xt + i * kWXfThreads) * 16u, pack8(f));
        }
        fence_proxy_async();
        }
        mbar_arrive(a_xf(slot));
        if (++slot == p.aslots) { slot = 0; sphase ^= 1u; }
      }
      if (++tin == tiles_per_img) { tin = 0; ++img; }
    }
  } else if (warp == kWTmaAWarp) {
    // ================================ TMA: the tile's activation chunks ====================================
    // the ring has more slots than a tile has chunks, so the first chunks of the NEXT tile(s) are already resident and
    // transformed when the MMA warp gets there (load + prologue latency is ~9000 cycles, a whole tile at K = 128)
    int slot = 0; uint32_t sphase = 0;
    for (int t = t_begin; t < t_end; ++t) {
      for (int ci = 0; ci < p.nchunks; ++ci) {
        const uint32_t cd = p.chunk[ci];
        mbar_wait_relaxed(a_empty(slot), sphase ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(a_raw(slot), kWChunk);
          tma_load_2d_w(sbase + (uint32_t)slot * kWChunk, &p.tmap_in[cd & 0xff], (int)((cd >> 8) & 0xff) * 8, t * 128, a_raw(slot));
        }
        __syncwarp();
        if (++slot == p.aslots) { slot = 0; sphase ^= 1u; }
      }
    }
  } else if (warp == kWTmaBWarp) {
    // ================================ weights: [n-block][chunk] stream, the same for every tile ==============
    int stage = 0; uint32_t phase = 0;
    const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.W);
    const int per_tile = p.NB * p.nchunks;
    for (int t = t_begin; t < t_end; ++t) {
      for (int q = 0; q < per_tile; ++q) {
        mbar_wait_relaxed(b_empty(stage), phase ^ 1u);
        if (elect_one()) {
          const uint32_t nb = (p.debug & 4) ? 128u : kWChunk;
          mbar_expect_tx(b_full(stage), nb);
          bulk_g2s(sbase + p.b_off + (uint32_t)stage * kWChunk, wsrc + (size_t)q * kWChunk, nb, b_full(stage));
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == kWMmaWarp) {
    // ================================ MMA issuer ============================================================
    // D = f32, A/B = bf16, both K-major; M = 128, N = 128, K = 16 per instruction, 4 per chunk
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | (8u << 24);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint64_t ad0 = umma_desc(sbase);
    const uint64_t bd0 = umma_desc(sbase + p.b_off);
    int stage = 0; uint32_t phase = 0;      // weight ring
    int slot0 = 0; uint32_t sphase0 = 0;     // activation ring position of the current tile's chunk 0
    uint32_t g = 0;
    const bool prof = (p.debug & 16) && blockIdx.x == 0;
    long long w_acc = 0, w_a = 0, w_b = 0, w_iss = 0, c0 = 0, c1 = 0;
    const long long c_start = prof ? clock64() : 0;
    for (int t = t_begin; t < t_end; ++t) {
      for (int j = 0; j < p.NB; ++j, ++g) {
        const uint32_t acc = g & 1u;
        if (prof) c0 = clock64();
        mbar_wait(tempty_bar(acc), ((g >> 1) & 1u) ^ 1u);
        if (prof) { c1 = clock64(); w_acc += c1 - c0; }
        const bool last = j == p.NB - 1;
        int slot = slot0; uint32_t sphase = sphase0;
        // two chunks (8 MMAs) per step: the fixed cost of a step (barrier polls, election, operands into uniform
        // registers, commits: ~300 cycles measured) is above the 256 cycles one chunk keeps the tensor core busy
        for (int ci = 0; ci < p.nchunks; ci += 2) {
          const bool two = ci + 1 < p.nchunks;
          int slot1 = slot + 1; uint32_t sphase1 = sphase;
          if (slot1 == p.aslots) { slot1 = 0; sphase1 ^= 1u; }
          int stage1 = stage + 1; uint32_t phase1 = phase;
          if (stage1 == p.stages) { stage1 = 0; phase1 ^= 1u; }
          if (prof) c0 = clock64();
          if (j == 0) mbar_wait(a_xf(slot), sphase);
          if (prof) { c1 = clock64(); w_a += c1 - c0; }
          mbar_wait(b_full(stage), phase);
          if (prof) { c0 = clock64(); w_b += c0 - c1; }
          tc_fence_after();
          if (elect_one()) {
            const uint32_t d = tmem_u + acc * 128u;
            {
              const uint64_t ad = ad0 + (uint64_t)slot * (kWChunk >> 4);
              const uint64_t bd = bd0 + (uint64_t)stage * (kWChunk >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_bf16(d, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (ci | k) != 0 ? 1u : 0u);
              umma_commit(b_empty(stage));
              if (last) umma_commit(a_empty(slot));          // the tile's chunk ci will not be read again
            }
            if (two) {
              // the elected thread alone polls for the second chunk: the first one is already on the tensor core
              if (j == 0) mbar_wait(a_xf(slot1), sphase1);
              mbar_wait(b_full(stage1), phase1);
              tc_fence_after();
              const uint64_t ad = ad0 + (uint64_t)slot1 * (kWChunk >> 4);
              const uint64_t bd = bd0 + (uint64_t)stage1 * (kWChunk >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_bf16(d, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, 1u);
              umma_commit(b_empty(stage1));
              if (last) umma_commit(a_empty(slot1));
            }
            if (ci + 2 >= p.nchunks) umma_commit(tfull_bar(acc));
          }
          __syncwarp();
          if (prof) { c1 = clock64(); w_iss += c1 - c0; }
          if (two) {
            slot = slot1 + 1; sphase = sphase1; if (slot == p.aslots) { slot = 0; sphase ^= 1u; }
            stage = stage1 + 1; phase = phase1; if (stage == p.stages) { stage = 0; phase ^= 1u; }
          } else {
            slot = slot1; sphase = sphase1;
            stage = stage1; phase = phase1;
          }
        }
        if (last) { slot0 = slot; sphase0 = sphase; }
      }
    }
    if (prof && lane == 0) { g_wprof[0] = w_acc; g_wprof[1] = w_a; g_wprof[2] = w_b; g_wprof[3] = w_iss; g_wprof[4] = clock64() - c_start; g_wprof[5] = g; }
  } else if (warp == 11) {
    // idle
  } else {
    // ================================ E0 / E1: accumulator -> fp16 -> global, statistics ======================
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    const int grp = warp_u >> 2;               // column half [64 grp, 64 grp + 64) of every accumulator
    const int ew = warp_u & 3;                 // TMEM lane quadrant = warp % 4
    const int et = ew * 32 + lane;             // thread index inside the group: owner of statistic entry `et` per n-block
    const uint32_t lane_base = tmem_base + ((uint32_t)(ew * 32) << 16);
    const int total_g = (t_end - t_begin) * p.NB;
    auto flush_stats = [&](int im) {
      for (int jj = 0; jj < p.NB; ++jj) {
        const int idx = (jj * 128 + grp * 64) * 2 + et;
        const float v = s_stat[idx];
        s_stat[idx] = 0.f;
        atomicAdd(p.stats + ((size_t)im * p.Nc + jj * 128 + grp * 64) * 2 + et, (double)v);
      }
    };
    int t = t_begin, j = 0;
    int img = t_begin / tiles_per_img, tin = t_begin - img * tiles_per_img;
    int cur_img = img;
    for (int g = 0; g < total_g; ++g) {
      if (img != cur_img) { flush_stats(cur_img); cur_img = img; }
      const int acc = g & 1;
      mbar_wait(tfull_bar(acc), (uint32_t)(g >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = lane_base + (uint32_t)acc * 128u + (uint32_t)grp * 64u;
      uint32_t r[4][16];
      tmem_ld16(taddr, r[0]); tmem_ld16(taddr + 16, r[1]); tmem_ld16(taddr + 32, r[2]); tmem_ld16(taddr + 48, r[3]);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));                    // this group's half of the accumulator is drained
      // lane = row after tcgen05.ld; transposed through a private 4 KB patch (XOR-swizzled 16-byte units) so that every
      // store instruction writes 4 complete 128-byte lines (gemm_expand.cu has the measurements)
      const uint32_t patch = sbase + p.stg_off + (uint32_t)warp_u * 4096u;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int q = u >> 1, h = (u & 1) * 8;
        sts128(patch + (uint32_t)lane * 128u + (uint32_t)((u ^ (lane & 7)) << 4),
               make_uint4(pack_f16(__uint_as_float(r[q][h + 0]), __uint_as_float(r[q][h + 1])),
                          pack_f16(__uint_as_float(r[q][h + 2]), __uint_as_float(r[q][h + 3])),
                          pack_f16(__uint_as_float(r[q][h + 4]), __uint_as_float(r[q][h + 5])),
                          pack_f16(__uint_as_float(r[q][h + 6]), __uint_as_float(r[q][h + 7]))));
      }
      __syncwarp();
      const int unit = lane & 7, rsub = lane >> 3;
      __half* obase = p.out + ((size_t)t * 128 + ew * 32 + rsub) * p.Nc + j * 128 + grp * 64 + unit * 8;
      uint4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = i * 4 + rsub;
        v[i] = lds128(patch + (uint32_t)row * 128u + (uint32_t)((unit ^ (row & 7)) << 4));
      }
#pragma unroll
      if (!(p.debug & 2))
#pragma unroll
      for (int i = 0; i < 8; ++i) *reinterpret_cast<uint4*>(obase + (size_t)(i * 4) * p.Nc) = v[i];
      __syncwarp();   // the patch is free again
      // ---- statistics of exactly what was stored (the fp16-rounded values) ----
      if (!(p.debug & 1)) {
        float s[8], q[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) { s[c] = 0.f; q[c] = 0.f; }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float f[8];
          unpack8h(v[i], f);
#pragma unroll
          for (int c = 0; c < 8; ++c) { s[c] += f[c]; q[c] = fmaf(f[c], f[c], q[c]); }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          s[c] += __shfl_xor_sync(0xffffffffu, s[c], 8);  q[c] += __shfl_xor_sync(0xffffffffu, q[c], 8);
          s[c] += __shfl_xor_sync(0xffffffffu, s[c], 16); q[c] += __shfl_xor_sync(0xffffffffu, q[c], 16);
        }
        float* part = s_part + ((grp * 2 + (g & 1)) * 4) * 128;     // [warp 4][64 columns][2]
        if (rsub == 0) {
          float4* dst = reinterpret_cast<float4*>(part + ew * 128 + unit * 16);
          dst[0] = make_float4(s[0], q[0], s[1], q[1]); dst[1] = make_float4(s[2], q[2], s[3], q[3]);
          dst[2] = make_float4(s[4], q[4], s[5], q[5]); dst[3] = make_float4(s[6], q[6], s[7], q[7]);
        }
        bar_sync(2 + grp, 128);
        // fixed order over the four row quadrants; the next write to this buffer is two n-blocks away, behind the
        // next barrier, which every warp reaches only after this read
        const float a = (part[et] + part[128 + et]) + (part[256 + et] + part[384 + et]);
        s_stat[(j * 128 + grp * 64) * 2 + et] += a;
      }
      if (++j == p.NB) { j = 0; ++t; if (++tin == tiles_per_img) { tin = 0; ++img; } }
    }
    if (total_g > 0) flush_stats(cur_img);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
  }
}

struct WideLayout { int nchunks, NB, stages, aslots; uint32_t b_off, stg_off, part_off, stat_off, coef_off, misc_off, total; };

bool plan_layout_w(int nseg, const int* segK, int Nc, WideLayout* L) {
  int nch = 0, K = 0;
  for (int s = 0; s < nseg; ++s) {
    if (segK[s] < 64 || segK[s] % 64) return false;
    nch += segK[s] / 64;
    K += segK[s];
  }
  if (nch < 2 || nch > kWMaxChunks || Nc % 128 || Nc < 128 || Nc / 128 > kWMaxNB) return false;
  L->nchunks = nch; L->NB = Nc / 128;
  const uint32_t stat = ((uint32_t)Nc * 8 + 1023) / 1024 * 1024, coef = ((uint32_t)K * 8 + 1023) / 1024 * 1024;
  const uint32_t fixed = kWPatchBytes + kWPartBytes + stat + coef + 1024 /* misc */ + 1024 /* align */;
  const int units = (int)((kWSmemLimit - fixed) / kWChunk);   // 16 KB units shared by the two rings
  if (units < nch + 3) return false;
  // activation ring: the tile plus up to two tiles of look-ahead, as long as the weight ring keeps 5 stages (sweep 3..8: flat from 4 on, 3 is 15 % slower at K = 384)
  static int min_b = -1;   // LCM_W_BSTAGES: weight-ring stages reserved before the activation ring gets look-ahead slots
  if (min_b < 0) { const char* e = getenv("LCM_W_BSTAGES"); min_b = e ? atoi(e) : 5; if (min_b < 3) min_b = 3; }
  int extra = units - nch - min_b;
  if (extra < 0) extra = 0;
  if (extra > 2 * nch) extra = 2 * nch;
  int aslots = nch + extra;
  if (aslots > 16) aslots = 16;
  int stages = units - aslots;
  if (stages > kWMaxStages) stages = kWMaxStages;
  L->stages = stages; L->aslots = aslots;
  uint32_t off = (uint32_t)aslots * kWChunk;
  L->b_off = off; off += (uint32_t)stages * kWChunk;
  L->stg_off = off; off += kWPatchBytes;
  L->part_off = off; off += kWPartBytes;
  L->stat_off = off; off += stat;
  L->coef_off = off; off += coef;
  L->misc_off = off; off += 1024;
  L->total = off + 1024;
  return true;
}

}  // namespace

int gemm_wide_read_profile(long long* host) { return cudaMemcpyFromSymbol(host, g_wprof, sizeof(long long) * 8) == cudaSuccess ? 0 : -1; }

bool gemm_wide_supported(int nseg, const int* segK, int Nc, int P) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("LCM_NO_WIDE_KERNEL"); off = (e && atoi(e)) ? 1 : 0; }
  if (off || nseg < 1 || nseg > 2 || P % 128) return false;
  WideLayout L;
  return plan_layout_w(nseg, segK, Nc, &L);
}

// W: bf16 image packed with block_n = 128 and the x6 scale (PackJob::scale) — see plan.cu / ops_api.cu.
// stats: [images][Nc][2] fp64, accumulated into (the plan zeroes the table once per forward).
int launch_gemm_wide(const GemmParams& g, int num_sms, cudaStream_t st) {
  if (g.nseg < 1 || g.nseg > 2 || !g.out_f16 || !g.stats || g.P % 128 || g.M % g.P || g.M <= 0 || g.M > 0x7fffff00LL) return -1;
  int segK[2] = {0, 0};
  for (int s = 0; s < g.nseg; ++s) {
    if (g.seg[s].mode != XF_AFFINE_RELU6 || g.seg[s].f16 || !g.seg[s].coef || g.seg[s].ld % 8) return -1;
    segK[s] = g.seg[s].K;
  }
  WideLayout L;
  if (!plan_layout_w(g.nseg, segK, g.Nc, &L)) return -1;
  WideParams p;
  memset(&p, 0, sizeof(p));
  p.nseg = g.nseg; p.nchunks = L.nchunks; p.NB = L.NB; p.stages = L.stages; p.aslots = L.aslots;
  p.W = reinterpret_cast<const bf16*>(g.W);
  p.out = reinterpret_cast<__half*>(g.out);
  p.stats = g.stats;
  p.m_tiles = (int)(g.M / 128); p.P = g.P; p.Nc = g.Nc;
  p.b_off = L.b_off; p.stg_off = L.stg_off; p.part_off = L.part_off; p.stat_off = L.stat_off;
  p.coef_smem_off = L.coef_off; p.misc_off = L.misc_off;
  int nch = 0, cbase = 0;
  for (int s = 0; s < g.nseg; ++s) {
    p.coef[s] = g.seg[s].coef; p.coef_ld[s] = g.seg[s].coef_ld; p.coef_off[s] = g.seg[s].coef_off; p.segK[s] = g.seg[s].K;
    if (!tmap_rows128(g.seg[s].A, g.M, g.seg[s].K, g.seg[s].ld, TMAP_BF16, &p.tmap_in[s])) return -3;
    for (int c0 = 0; c0 < g.seg[s].K; c0 += 64)
      p.chunk[nch++] = (uint32_t)s | ((uint32_t)(c0 / 8) << 8) | ((uint32_t)((cbase + c0) / 8) << 16);
    cbase += g.seg[s].K;
  }
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("LCM_W_DEBUG"); dbg = e ? atoi(e) : 0; } p.debug = dbg; }
  if (ensure_dyn_smem_fn(gemm_wide_kernel, kWSmemLimit)) return -2;
  const int grid = p.m_tiles < num_sms ? p.m_tiles : num_sms;
  launch_pdl(gemm_wide_kernel, dim3(grid), dim3(kThreadsW), (size_t)L.total, st, p);
  return 0;
}

}  // namespace lcm
