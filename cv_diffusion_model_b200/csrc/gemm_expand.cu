// Expand GEMM of the inverted-residual block (efficient_unet.py:174,207-209) on the tensor-core path:
//
//   h1[m][n] = sum_k relu6(a_k * x[m][k] + b_k) * W[n][k]        x: bf16 residual stream (1-2 concat segments, K <= 128)
//                                                                 h1: fp16 hidden tensor + per-(image, channel) sum / sum^2
//
// K is small and N = 4K, so this op is a pure streaming problem — (K + N) * 2 bytes per pixel, ~40 FLOP/B — and what can
// keep it off the HBM roofline is everything that is done per OUTPUT element.  The general kernel (gemm_tc2.cu) spends
// ~9000 warp instructions per 128x128 tile in its copy-out + column-statistics roles; the first version of this kernel
// (TMA-store staging + statistics MMAs over the staged tile and its square) was bound by shared-memory bandwidth.  This
// version does NOTHING per output element except convert and store:
//
//   * statistics from the INPUT side.  With A' = relu6(..)/6 (the MMA's A operand) and W6 = 6 W:
//         sum_m out[m][n]   = W6[n] . S          S[k]     = sum_m A'[m][k]          (prologue warps, registers)
//         sum_m out[m][n]^2 = W6[n]^T G W6[n]    G[k][k'] = sum_m A'[m][k] A'[m][k'] (one Gram MMA per 16 rows)
//     The Gram MMA reads the activation chunks that are already in shared memory as MN-major operands for BOTH A and B
//     (D[128][64 nch] += chunk^T chunk) and accumulates in TMEM over all tiles of an image; S and G are added to a small
//     per-image scratch at image boundaries and `expand_stats_kernel` turns them into the usual (sum, sum^2) table.
//     Cost per tile: 8 small MMAs and 2 x 16 KB of shared-memory reads per chunk, independent of N.
//   * the epilogue is tcgen05.ld -> cvt.f16x2 -> a per-warp 32 x 128 B transpose through a private shared-memory patch
//     (lane = row after tcgen05.ld; row-per-thread stores reach only 1-2.3 TB/s with 16-byte and 4.8 TB/s with 32-byte
//     vectors — tests/diag/rowstore_bw.cu) -> coalesced 16-byte stores.  No block barrier, no fence, no TMA store.
//   * prologue: relu6(a x + b) = 6 sat(a/6 x + b/6): one FFMA.SAT per element, the 6 is folded into the packed weights.
//   * an m-tile (128 pixels) is multiplied by ALL n-blocks of the resident weights, so x is read exactly once.
//
// Roles (512 threads = 16 warps, 4 per register-file partition; one persistent CTA per SM):
//   warps 0-3, 4-7  E0, E1  each drains one 64-column half of every 128-column accumulator; group 0 also moves the
//                           Gram matrix from TMEM to the scratch at image boundaries
//   warp  8         MMA     main MMAs (M=128, N=128, K=16) n-block after n-block
//   warp  9         TMA     weights once; activation chunks (128 rows x 64 channels, 128-byte swizzle) ring
//   warp  10        GRAM    the tile's Gram MMAs (a second issuer: the issue code runs on the uniform datapath at
//                           ~10 cycles per dependent instruction, one warp doing both was the critical path)
//   warps 12-15     XF      in-place prologue on the landed chunk + column sums
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

#include "kernels.h"
#include "tc_common.cuh"
#include "tmap.h"

namespace lcm {

namespace {

using namespace tc;

constexpr int kThreadsX = 512;
constexpr int kXfBaseX = 384, kXfThreadsX = 128;
constexpr int kMmaWarp = 8, kTmaWarp = 9, kGramWarp = 10, kXfWarp0 = 12;   // warp 11 idles
constexpr uint32_t kChunkBytes = 16384;     // 128 rows x 64 16-bit elements
constexpr uint32_t kWChunkBytes = 16384;    // 128 output channels x 64 k
constexpr uint32_t kSmemLimitX = 232448;
constexpr int kMaxNB = 4;                   // n-blocks of 128 output channels (each drained as two 64-column halves)
constexpr int kMaxChunksX = 2;              // 64-wide K chunks per tile (Gram matrix = one 128-lane TMEM block)
constexpr uint32_t kGramCol0 = 256;         // TMEM: columns 0-255 = two 128-column accumulators, 256-383 = Gram matrix
constexpr int kGramLd = 128;                // scratch: G[img][128][128], S[img][128] (fp64)

struct XParams {
  CUtensorMap tmap_in[2];
  const float2* coef[2];
  int coef_ld[2], coef_off[2], segK[2];
  int nseg, nchunks, NB, stages;
  const bf16* W;            // packed [n-block][chunk][128 rows x 64 k] (128-byte swizzled rows), scaled by 6
  __half* out;
  double* gram;             // [images][128][128]  fp64: a few fp32 partial sums per entry add exactly, in any order
  double* colsum;           // [images][128]
  int m_tiles, P, Nc;
  uint32_t chunk[kMaxChunksX];   // seg | kvalid << 8 | c0/8 << 16 | coef base << 24
  uint32_t w_off, stg_off, coef_smem_off, misc_off;
  int debug;
  int dbg2;   // LCM_X_DBG bottleneck experiments: 1 skip the XF math, 2 skip the Gram MMAs, 4 one prologue thread group only
};

__device__ __forceinline__ void tma_load_2d_x(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
// MN-major operand, 128-byte swizzle: 64 elements (128 B) contiguous along MN, 8-row groups along K 1024 B apart,
// 64-wide MN blocks `lbo` bytes apart
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t lbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void stg256(void* p, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]),
               "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

__device__ long long g_xtimeline[64 * 16];   // LCM_X_TIMELINE: clock64 stamps of block 0, first 64 tiles
#define XSTAMP(tile_idx, slot) do { if (p.debug && blockIdx.x == 0 && (tile_idx) < 64) g_xtimeline[(tile_idx) * 16 + (slot)] = clock64(); } while (0)

// kStatsOnly: the statistics half alone — S and G of the transformed input, no main MMAs, nothing stored.  It runs ahead
// of the fused expand -> depthwise kernel (xdw_fused.cu), which needs the GroupNorm2 coefficients of a tensor it never
// materialises; the pass reads only the (4x narrower) block input.
template <bool kStatsOnly>
__global__ void __launch_bounds__(kThreadsX, 1) gemm_expand_kernel(const __grid_constant__ XParams p) {
  extern __shared__ uint8_t xsm_raw[];
  const uint32_t sraw = smem_u32(xsm_raw);
  const uint32_t sbase = (sraw + 1023u) & ~1023u;
  uint8_t* smem = xsm_raw + (sbase - sraw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const uint32_t bar0 = sbase + p.misc_off;
  auto raw_bar = [&](int s) { return bar0 + 8u * s; };
  auto xf_bar = [&](int s) { return bar0 + 8u * (8 + s); };
  auto empty_bar = [&](int s) { return bar0 + 8u * (16 + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (24 + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (26 + a); };
  const uint32_t wres_bar = bar0 + 8u * 28;
  const uint32_t gdone_bar = bar0 + 8u * 29;   // Gram MMAs of a tile have completed (one phase per tile)
  const uint32_t sread_bar = bar0 + 8u * 30;   // Gram matrix read out of TMEM (image flush)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + p.misc_off + 320);
  float2* s_coef = reinterpret_cast<float2*>(smem + p.coef_smem_off);

  if (warp == kTmaWarp && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(raw_bar(s), 1); mbar_init(xf_bar(s), kXfThreadsX); mbar_init(empty_bar(s), 2); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 256); }
    mbar_init(wres_bar, 1);
    mbar_init(gdone_bar, 1);
    mbar_init(sread_bar, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int s = 0; s < p.nseg; ++s)
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmap_in[s])) : "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  const int t_begin = (int)((long long)p.m_tiles * blockIdx.x / gridDim.x);
  const int t_end = (int)((long long)p.m_tiles * (blockIdx.x + 1) / gridDim.x);
  const int tiles_per_img = p.P >> 7;
  const int kw = 64 * p.nchunks;   // Gram width

  if (warp >= kXfWarp0) {
    // ================================ XF: relu6(a x + b) / 6 in place, column sums ==========================
    const int xt = tid - kXfBaseX;
    // unit u = xt + 128 i: row = u >> 3, slot = u & 7 holds channel unit cu = slot ^ (row & 7); (row & 7) does not
    // depend on i (128 = 16 rows), so a thread transforms the same 8 channels in all of its rows
    const int cu = (xt & 7) ^ ((xt >> 3) & 7);
    int stage = 0; uint32_t phase = 0;
    int img = t_begin / tiles_per_img, tin = t_begin - img * tiles_per_img;
    int coef_img = -1;
    float cs[kMaxChunksX][8];
#pragma unroll
    for (int c = 0; c < kMaxChunksX; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j) cs[c][j] = 0.f;
    auto flush_cs = [&](int im) {
#pragma unroll
      for (int c = 0; c < kMaxChunksX; ++c) {
        if (c < p.nchunks && cu * 8 < (int)((p.chunk[c] >> 8) & 0xff)) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { atomicAdd(&p.colsum[(size_t)im * kGramLd + c * 64 + cu * 8 + j], (double)cs[c][j]); cs[c][j] = 0.f; }
        }
      }
    };
    for (int t = t_begin; t < t_end; ++t) {
      if (img != coef_img) {
        if (coef_img >= 0) flush_cs(coef_img);
        bar_sync(1, kXfThreadsX);   // everyone is done with the previous image's coefficients
        for (int s = 0, base = 0; s < p.nseg; base += p.segK[s], ++s) {
          const float2* src = p.coef[s] + (size_t)img * p.coef_ld[s] + p.coef_off[s];
          for (int k = xt; k < p.segK[s]; k += kXfThreadsX) {
            const float2 c = src[k];
            s_coef[base + k] = make_float2(c.x * (1.f / 6.f), c.y * (1.f / 6.f));
          }
        }
        bar_sync(1, kXfThreadsX);
        coef_img = img;
      }
#pragma unroll
      for (int ci = 0; ci < kMaxChunksX; ++ci) {
        if (ci >= p.nchunks) break;
        const uint32_t cd = p.chunk[ci];
        const int kvalid = (cd >> 8) & 0xff, cbase = cd >> 24;
        const uint32_t a_smem = sbase + (uint32_t)stage * kChunkBytes;
        const bool act = cu * 8 < kvalid;
        float2 ab[8];
        if (act) {
          const float4* c4 = reinterpret_cast<const float4*>(s_coef + cbase + cu * 8);
#pragma unroll
          for (int j = 0; j < 4; ++j) { const float4 c = c4[j]; ab[2 * j] = make_float2(c.x, c.y); ab[2 * j + 1] = make_float2(c.z, c.w); }
        }
        mbar_wait(raw_bar(stage), phase);
        if (ci == 0 && xt == 0) XSTAMP(t - t_begin, 1);
        if (act && !(p.dbg2 & 1)) {
          uint4 v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = lds128(a_smem + (uint32_t)(xt + i * kXfThreadsX) * 16u);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            {
              float f[8];
              unpack8(v[i], f);
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = __saturatef(fmaf(ab[j].x, f[j], ab[j].y));
              const uint4 pk = pack8(f);
              sts128(a_smem + (uint32_t)(xt + i * kXfThreadsX) * 16u, pk);
              // column sums of exactly what the tensor core multiplies (the bf16-rounded values), so that S and G
              // describe the same matrix
              cs[ci][0] += bf16lo(pk.x); cs[ci][1] += bf16hi(pk.x); cs[ci][2] += bf16lo(pk.y); cs[ci][3] += bf16hi(pk.y);
              cs[ci][4] += bf16lo(pk.z); cs[ci][5] += bf16hi(pk.z); cs[ci][6] += bf16lo(pk.w); cs[ci][7] += bf16hi(pk.w);
            }
          }
          fence_proxy_async();
        }
        mbar_arrive(xf_bar(stage));
        if (ci == p.nchunks - 1 && xt == 0) XSTAMP(t - t_begin, 2);
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
      if (++tin == tiles_per_img) { tin = 0; ++img; }
    }
    if (coef_img >= 0) flush_cs(coef_img);
  } else if (warp == kTmaWarp) {
    // ================================ TMA: weights once, activation chunks ================================
    if (!kStatsOnly && elect_one()) {
      const uint32_t wbytes = (uint32_t)p.NB * p.nchunks * kWChunkBytes;
      mbar_expect_tx(wres_bar, wbytes);
      for (uint32_t o = 0; o < wbytes; o += kWChunkBytes)
        bulk_g2s(sbase + p.w_off + o, reinterpret_cast<const uint8_t*>(p.W) + o, kWChunkBytes, wres_bar);
    }
    __syncwarp();
    int stage = 0; uint32_t phase = 0;
    for (int t = t_begin; t < t_end; ++t) {
      for (int ci = 0; ci < p.nchunks; ++ci) {
        const uint32_t cd = p.chunk[ci];
        mbar_wait_relaxed(empty_bar(stage), phase ^ 1u);
        if (ci == 0 && lane == 0) XSTAMP(t - t_begin, 0);
        if (elect_one()) {
          mbar_expect_tx(raw_bar(stage), kChunkBytes);
          tma_load_2d_x(sbase + (uint32_t)stage * kChunkBytes, &p.tmap_in[cd & 0xff], (int)((cd >> 16) & 0xff) * 8, t * 128, raw_bar(stage));
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == kMmaWarp) {
    // ================================ MMA issuer ============================================================
    // main: D = f32, A/B = bf16, both K-major; M = 128, N = 128
    const uint32_t idesc_main = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | (8u << 24);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const int ks0 = (int)((p.chunk[0] >> 8) & 0xff) >> 4;                          // K steps of 16 per chunk
    const int ks1 = p.nchunks > 1 ? (int)((p.chunk[1] >> 8) & 0xff) >> 4 : 0;
    const uint64_t wd0 = umma_desc(sbase + p.w_off);
    if (!kStatsOnly) mbar_wait(wres_bar, 0);
    int stage = 0; uint32_t phase = 0;      // activation ring position of the CURRENT tile's first chunk
    int g = 0;                               // n-blocks issued so far (accumulator = g & 1)
    for (int t = t_begin; t < t_end; ++t) {
      // chunks of this tile become ready one after the other; they are needed by every n-block, so wait for all
      {
        int s2 = stage; uint32_t ph2 = phase;
        for (int ci = 0; ci < p.nchunks; ++ci) {
          mbar_wait(xf_bar(s2), ph2);
          if (++s2 == p.stages) { s2 = 0; ph2 ^= 1u; }
        }
      }
      tc_fence_after();
      if (lane == 0) XSTAMP(t - t_begin, 3);
      // The issue code runs on the uniform datapath, where every dependent instruction costs several cycles: keep it
      // to constant-offset descriptor adds and predicated MMAs (the generic loop version took ~600 cycles per n-block).
      const uint64_t ad0 = umma_desc(sbase + (uint32_t)stage * kChunkBytes);   // chunk ci at + ci * (16384 >> 4)
      for (int j = 0; j < (kStatsOnly ? 0 : p.NB); ++j, ++g) {
        const int acc = g & 1;
        mbar_wait(tempty_bar(acc), ((uint32_t)(g >> 1) & 1u) ^ 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t bd0 = wd0 + (uint64_t)(j * p.nchunks) * (kWChunkBytes >> 4);
          const uint32_t d = tmem_u + (uint32_t)acc * 128u;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k < ks0) umma_bf16(d, ad0 + (uint64_t)(2 * k), bd0 + (uint64_t)(2 * k), idesc_main, k != 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k < ks1) umma_bf16(d, ad0 + (uint64_t)((kChunkBytes >> 4) + 2 * k), bd0 + (uint64_t)((kWChunkBytes >> 4) + 2 * k), idesc_main, 1u);
          umma_commit(tfull_bar(acc));
        }
        __syncwarp();
        if (lane == 0 && j < 2) XSTAMP(t - t_begin, 8 + j);
      }
      if (elect_one())
        for (int ci = 0; ci < p.nchunks; ++ci) umma_commit(empty_bar(stage + ci));   // one of the two readers of the tile's chunks
      __syncwarp();
      if (lane == 0) XSTAMP(t - t_begin, 4);
      stage += p.nchunks;
      if (stage == p.stages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == kGramWarp) {
    // ================================ Gram issuer =============================================================
    // A and B are the SAME activation chunk(s) read MN-major (bits 15, 16); M = 128, N = kw; the matrix accumulates in
    // TMEM over the tiles of one image
    const uint32_t idesc_gram = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(kw >> 3) << 17) | (8u << 24);
    const uint32_t gram_lbo = p.nchunks == 2 ? kChunkBytes : 0u;   // one chunk: MN block 1 aliases block 0 (lanes 64-127 unused)
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    int stage = 0; uint32_t phase = 0;
    int tin = t_begin % tiles_per_img;
    uint32_t sread_phase = 0;
    for (int t = t_begin; t < t_end; ++t) {
      {
        int s2 = stage; uint32_t ph2 = phase;
        for (int ci = 0; ci < p.nchunks; ++ci) {
          mbar_wait(xf_bar(s2), ph2);
          if (++s2 == p.stages) { s2 = 0; ph2 ^= 1u; }
        }
      }
      const bool first_tile = t == t_begin || tin == 0;
      if (first_tile && t != t_begin) { mbar_wait(sread_bar, sread_phase); sread_phase ^= 1u; }
      tc_fence_after();
      if (lane == 0) XSTAMP(t - t_begin, 13);
      if (elect_one()) {
        const uint64_t d0 = umma_desc_mn(sbase + (uint32_t)stage * kChunkBytes, gram_lbo);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (!(p.dbg2 & 2)) umma_bf16(tmem_u + kGramCol0, d0 + (uint64_t)(k * (2048 >> 4)), d0 + (uint64_t)(k * (2048 >> 4)), idesc_gram, (first_tile && k == 0) ? 0u : 1u);
        umma_commit(gdone_bar);
        for (int ci = 0; ci < p.nchunks; ++ci) umma_commit(empty_bar(stage + ci));   // the other reader of the tile's chunks
      }
      __syncwarp();
      if (lane == 0) XSTAMP(t - t_begin, 14);
      stage += p.nchunks;
      if (stage == p.stages) { stage = 0; phase ^= 1u; }
      if (++tin == tiles_per_img) tin = 0;
    }
  } else if (warp == 11) {
    // idle
  } else {
    // ================================ E0 / E1: accumulator -> fp16 -> global ================================
    // the warp index is broadcast so that the compiler keeps everything derived from it in uniform registers
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    const int grp = warp_u >> 2;               // owns n-blocks g with (g & 1) == grp: accumulator grp
    const int ew = warp_u & 3;                 // TMEM lane quadrant = warp % 4
    const int et = ew * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(ew * 32) << 16);
    const int total_g = (t_end - t_begin) * p.NB;
    int cur_img = -1;
    // image flush (group 0): Gram matrix TMEM -> scratch.  `lt` = CTA-local index of the image's last tile; gdone
    // completes one phase per tile and cannot run ahead of this wait (the next image's first Gram MMA waits for sread)
    auto flush = [&](int im, int lt) {
      if (!kStatsOnly) mbar_wait(gdone_bar, (uint32_t)lt & 1u);
      tc_fence_after();
      if (ew * 32 < kw) {
        double* grow = p.gram + ((size_t)im * kGramLd + et) * kGramLd;
        for (int c = 0; c < kw; c += 16) {
          uint32_t r[16];
          tmem_ld16(lane_base + kGramCol0 + c, r);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) atomicAdd(grow + c + i, (double)__uint_as_float(r[i]));
        }
      }
      tc_fence_before();
      mbar_arrive(sread_bar);
    };
    int t = t_begin, j = 0;
    int img = t_begin / tiles_per_img, tin = t_begin - img * tiles_per_img;   // image of tile t, tile index inside it
    if (kStatsOnly) {
      // only the image flushes of group 0 remain
      // (nothing else paces these warps, so they follow gdone tile by tile: a parity wait is only meaningful when the
      // waiter is less than one phase behind)
      if (grp == 0) {
        for (; t < t_end; ++t) {
          mbar_wait(gdone_bar, (uint32_t)(t - t_begin) & 1u);
          if (tin == tiles_per_img - 1 || t == t_end - 1) flush(img, t - t_begin);
          if (++tin == tiles_per_img) { tin = 0; ++img; }
        }
      }
    } else {
    // every 128-column accumulator is drained by both groups: group grp takes columns [64 grp, 64 grp + 64)
    for (int g = 0; g < total_g; ++g) {
      // (t - tin) is the first tile of the current image: the previous image ended one tile before it
      if (grp == 0 && img != cur_img) { if (cur_img >= 0) flush(cur_img, t - tin - 1 - t_begin); cur_img = img; }
      const int acc = g & 1;
      mbar_wait(tfull_bar(acc), (uint32_t)(g >> 1) & 1u);
      if (et == 0 && j < 2) XSTAMP(t - t_begin, 5 + 5 * j);
      tc_fence_after();
      const uint32_t taddr = lane_base + (uint32_t)acc * 128u + (uint32_t)grp * 64u;
      uint32_t r[4][16];
      tmem_ld16(taddr, r[0]); tmem_ld16(taddr + 16, r[1]); tmem_ld16(taddr + 32, r[2]); tmem_ld16(taddr + 48, r[3]);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));                    // this group's half of the accumulator is drained
      if (et == 0 && j < 2) XSTAMP(t - t_begin, 6 + 5 * j);
      // lane = row, so a thread holds 128 contiguous bytes of ITS row; stored like that a warp instruction touches 32
      // different lines (measured: ~1 TB/s with 16-byte, ~4.8 TB/s with 32-byte stores).  Each warp transposes its
      // 32 x 128 B block through a private 4 KB shared-memory patch (XOR-swizzled 16-byte units, conflict-free both
      // ways, warp-level sync only) so that every store instruction writes 4 complete 128-byte lines.
      {
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
        for (int i = 0; i < 8; ++i) *reinterpret_cast<uint4*>(obase + (size_t)(i * 4) * p.Nc) = v[i];
        __syncwarp();   // the patch is free again
      }
      if (et == 0 && j < 2) XSTAMP(t - t_begin, 7 + 5 * j);
      if (++j == p.NB) { j = 0; ++t; if (++tin == tiles_per_img) { tin = 0; ++img; } }
    }
    if (grp == 0 && cur_img >= 0) flush(cur_img, t_end - 1 - t_begin);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// (S, G, W6) -> per-(image, channel) sum and sum of squares of the expand output.  grid = (images, Nc / 32), 128 threads:
// thread = (output channel n, quarter of the k range).  The channel's weight row sits in registers (de-swizzled from the
// packed tcgen05 image), G in shared memory (broadcast LDS.128); products are summed in fp32 runs of 16, beyond in fp64.
template <int KW>
__global__ void __launch_bounds__(128) expand_stats_kernel(const double* __restrict__ gram, const double* __restrict__ colsum,
                                                           const bf16* __restrict__ W, double* __restrict__ stats, int Nc,
                                                           int nchunks) {
  extern __shared__ __align__(16) float gs[];   // G[KW][KW], S[KW], Wt[KW][32], partial[4][32][2] (double)
  float* ss = gs + KW * KW;
  float* wt = ss + KW;
  double* part = reinterpret_cast<double*>(wt + KW * 32);
  const int img = blockIdx.x, tid = threadIdx.x, nl = tid & 31, kq = tid >> 5, n = blockIdx.y * 32 + nl;
  const double* g = gram + (size_t)img * kGramLd * kGramLd;
  pdl_wait();
  pdl_trigger();
  for (int i = tid; i < KW * KW / 2; i += 128) {
    const int row = (i * 2) / KW, col = (i * 2) % KW;
    const double2 v = *reinterpret_cast<const double2*>(g + row * kGramLd + col);
    reinterpret_cast<float2*>(gs)[i] = make_float2((float)v.x, (float)v.y);
  }
  for (int i = tid; i < KW; i += 128) ss[i] = (float)colsum[(size_t)img * kGramLd + i];
  float w[KW];
  {
    const int jb = n >> 7, r = n & 127;
    constexpr int NCH = (KW + 63) / 64;   // KW = 32 / 48 / 96: only the first units of the last chunk
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const uint4* row = reinterpret_cast<const uint4*>(W + ((size_t)(jb * nchunks + c) * 128 + r) * 64);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (c * 64 + u * 8 >= KW) continue;
        const uint4 v = row[u ^ (r & 7)];   // logical unit u sits in slot u ^ (r & 7)
        float f[8];
        tc::unpack8(v, f);
#pragma unroll
        for (int e = 0; e < 8; ++e) w[c * 64 + u * 8 + e] = f[e];
      }
    }
  }
  if (kq == 0) {
#pragma unroll
    for (int k = 0; k < KW; ++k) wt[k * 32 + nl] = w[k];
  }
  __syncthreads();
  double s = 0.0, q = 0.0;
  for (int k = kq * (KW / 4); k < (kq + 1) * (KW / 4); ++k) {
    const float4* grow = reinterpret_cast<const float4*>(gs + k * KW);
    double t = 0.0;
#pragma unroll
    for (int b = 0; b < KW / 16; ++b) {
      float a = 0.f;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const float4 g4 = grow[b * 4 + v];
        a = fmaf(g4.x, w[b * 16 + v * 4 + 0], a); a = fmaf(g4.y, w[b * 16 + v * 4 + 1], a);
        a = fmaf(g4.z, w[b * 16 + v * 4 + 2], a); a = fmaf(g4.w, w[b * 16 + v * 4 + 3], a);
      }
      t += (double)a;
    }
    const double wk = (double)wt[k * 32 + nl];
    q += wk * t;
    s += wk * (double)ss[k];
  }
  part[(kq * 32 + nl) * 2] = s;
  part[(kq * 32 + nl) * 2 + 1] = q;
  __syncthreads();
  if (tid < 64) {
    const int c = tid >> 1, m = tid & 1;
    const double v = (part[(0 * 32 + c) * 2 + m] + part[(1 * 32 + c) * 2 + m]) + (part[(2 * 32 + c) * 2 + m] + part[(3 * 32 + c) * 2 + m]);
    stats[((size_t)img * Nc + blockIdx.y * 32 + c) * 2 + m] += v;   // exclusive owner of this entry (the table is zeroed per forward)
  }
}

std::mutex g_x_mu;
struct XKey {
  const void* ptr; long long M; int K, ld, kind;
  bool operator==(const XKey& o) const { return ptr == o.ptr && M == o.M && K == o.K && ld == o.ld && kind == o.kind; }
};
struct XKeyHash {
  size_t operator()(const XKey& k) const {
    return std::hash<const void*>()(k.ptr) ^ (std::hash<long long>()(k.M) * 1315423911u) ^ ((size_t)k.K << 20) ^ ((size_t)k.ld << 2) ^ (size_t)k.kind;
  }
};
std::unordered_map<XKey, CUtensorMap, XKeyHash> g_x_maps;

// [M][K] 16-bit matrix with row stride ld: box = 64 elements x 128 rows, 128-byte swizzle
bool matrix_map(const void* ptr, long long M, int K, int ld, int dtype, CUtensorMap* out) {
  XKey key{ptr, M, K, ld, dtype};
  std::lock_guard<std::mutex> lk(g_x_mu);
  auto it = g_x_maps.find(key);
  if (it != g_x_maps.end()) { *out = it->second; return true; }
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)M};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, 128};
  CUtensorMap m;
  if (!encode_tmap(&m, dtype, 2, ptr, gdim, gstride, box, true)) return false;
  if (g_x_maps.size() > 4096) g_x_maps.clear();
  g_x_maps[key] = m;
  *out = m;
  return true;
}

struct XLayout { int nchunks, NB, stages; uint32_t w_off, stg_off, coef_off, misc_off, total; };

bool plan_layout(int nseg, const int* segK, int Nc, XLayout* L) {
  int nch = 0, ncoef = 0;
  for (int s = 0; s < nseg; ++s) {
    if (segK[s] % 16 || segK[s] < 16 || segK[s] > 64 * kMaxChunksX) return false;
    nch += (segK[s] + 63) / 64;
    ncoef += segK[s];
  }
  if (nch < 1 || nch > kMaxChunksX || Nc % 128 || Nc / 128 > kMaxNB || Nc < 128) return false;
  L->nchunks = nch; L->NB = Nc / 128;
  const uint32_t wbytes = (uint32_t)L->NB * nch * kWChunkBytes;
  const uint32_t fixed = wbytes + 32768 /* epilogue transpose patches */ + 2048 /* coef */ + 1024 /* misc */ + 1024 /* align */;
  if (fixed + (uint32_t)(2 * nch) * kChunkBytes > kSmemLimitX) return false;
  int stages = (int)((kSmemLimitX - fixed) / kChunkBytes);
  if (stages > 8) stages = 8;
  stages -= stages % nch;   // a tile's chunks never wrap around the ring
  L->stages = stages;
  uint32_t off = (uint32_t)stages * kChunkBytes;
  L->w_off = off; off += wbytes;
  L->stg_off = off; off += 32768;
  L->coef_off = off; off += 2048;
  L->misc_off = off; off += 1024;
  L->total = off + 1024;
  return true;
}

}  // namespace

bool tmap_rows128(const void* ptr, long long M, int K, int ld, int dtype, CUtensorMap* out) { return matrix_map(ptr, M, K, ld, dtype, out); }

int gemm_expand_read_timeline(long long* host, int n) {
  return cudaMemcpyFromSymbol(host, g_xtimeline, sizeof(long long) * (n < 1024 ? n : 1024)) == cudaSuccess ? 0 : -1;
}

bool gemm_expand_supported(int nseg, const int* segK, int Nc, int P) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("LCM_NO_EXPAND_KERNEL"); off = (e && atoi(e)) ? 1 : 0; }
  if (off || nseg < 1 || nseg > 2 || P % 128) return false;
  XLayout L;
  return plan_layout(nseg, segK, Nc, &L);
}

size_t gemm_expand_scratch_bytes(int images) { return (size_t)images * (kGramLd * kGramLd + kGramLd) * sizeof(double); }

// (S, G) scratch -> (sum, sum^2) of the expand output; W packed as for gemm_expand with `nchunks` 64-wide K chunks per n-block
// (the statistics pass of the fused expand -> depthwise path, xstats.cu, ends with this)
int launch_expand_stats_finalize(void* scratch, const void* W, double* stats, int images, int Nc, int nchunks, cudaStream_t st, int Ktot) {
  if (nchunks < 1 || nchunks > 2 || Nc % 32) return -1;
  if (ensure_dyn_smem_fn(expand_stats_kernel<128>, (128 * 128 + 128 + 128 * 32 + 512) * 4) ||
      ensure_dyn_smem_fn(expand_stats_kernel<64>, (64 * 64 + 64 + 64 * 32 + 512) * 4)) return -2;
  const double* gram = reinterpret_cast<const double*>(scratch);
  const double* colsum = gram + (size_t)images * kGramLd * kGramLd;
  const dim3 sg(images, Nc / 32);
  // instantiations at the actual width (the 64- / 128-wide ones multiply the zero padding too: 4x the products at K = 32)
  if (nchunks == 1 && Ktot > 0 && Ktot <= 32)
    launch_pdl(expand_stats_kernel<32>, sg, dim3(128), (size_t)(32 * 32 + 32 + 32 * 32 + 512) * 4, st, gram, colsum, reinterpret_cast<const bf16*>(W), stats, Nc, nchunks);
  else if (nchunks == 1 && Ktot > 0 && Ktot <= 48)
    launch_pdl(expand_stats_kernel<48>, sg, dim3(128), (size_t)(48 * 48 + 48 + 48 * 32 + 512) * 4, st, gram, colsum, reinterpret_cast<const bf16*>(W), stats, Nc, nchunks);
  else if (nchunks == 2 && Ktot > 0 && Ktot <= 96) {
    if (ensure_dyn_smem_fn(expand_stats_kernel<96>, (96 * 96 + 96 + 96 * 32 + 512) * 4)) return -2;
    launch_pdl(expand_stats_kernel<96>, sg, dim3(128), (size_t)(96 * 96 + 96 + 96 * 32 + 512) * 4, st, gram, colsum, reinterpret_cast<const bf16*>(W), stats, Nc, nchunks);
  } else if (nchunks == 1)
    launch_pdl(expand_stats_kernel<64>, sg, dim3(128), (size_t)(64 * 64 + 64 + 64 * 32 + 512) * 4, st, gram, colsum, reinterpret_cast<const bf16*>(W), stats, Nc, nchunks);
  else
    launch_pdl(expand_stats_kernel<128>, sg, dim3(128), (size_t)(128 * 128 + 128 + 128 * 32 + 512) * 4, st, gram, colsum, reinterpret_cast<const bf16*>(W), stats, Nc, nchunks);
  return 0;
}

// W: bf16 image packed with block_n = 64 and the x6 scale (PackJob::scale) — see plan.cu / ops_api.cu.
// scratch: gemm_expand_scratch_bytes(images) bytes; zero_scratch: clear it here, on the stream (the plan hands in a slice
// of its zeroed region instead, which keeps a memset node out of every block of the graph).
int launch_gemm_expand(const GemmParams& g, void* scratch, bool zero_scratch, int num_sms, cudaStream_t st, bool stats_only) {
  if (g.nseg < 1 || g.nseg > 2 || (!g.out_f16 && !stats_only) || !g.stats || !scratch || g.P % 128 || g.M % g.P || g.M <= 0 || g.M > 0x7fffff00LL) return -1;
  int segK[2] = {0, 0};
  for (int s = 0; s < g.nseg; ++s) {
    if (g.seg[s].mode != XF_AFFINE_RELU6 || g.seg[s].f16 || !g.seg[s].coef || g.seg[s].ld % 8) return -1;
    segK[s] = g.seg[s].K;
  }
  XLayout L;
  if (!plan_layout(g.nseg, segK, g.Nc, &L)) return -1;
  const int images = (int)(g.M / g.P);
  XParams p;
  memset(&p, 0, sizeof(p));
  p.nseg = g.nseg; p.nchunks = L.nchunks; p.NB = L.NB; p.stages = L.stages;
  p.W = reinterpret_cast<const bf16*>(g.W);
  p.out = reinterpret_cast<__half*>(g.out);
  p.gram = reinterpret_cast<double*>(scratch);
  p.colsum = p.gram + (size_t)images * kGramLd * kGramLd;
  p.m_tiles = (int)(g.M / 128); p.P = g.P; p.Nc = g.Nc;
  p.w_off = L.w_off; p.stg_off = L.stg_off; p.coef_smem_off = L.coef_off; p.misc_off = L.misc_off;
  int nch = 0, cbase = 0;
  for (int s = 0; s < g.nseg; ++s) {
    p.coef[s] = g.seg[s].coef; p.coef_ld[s] = g.seg[s].coef_ld; p.coef_off[s] = g.seg[s].coef_off; p.segK[s] = g.seg[s].K;
    if (!matrix_map(g.seg[s].A, g.M, g.seg[s].K, g.seg[s].ld, TMAP_BF16, &p.tmap_in[s])) return -3;
    for (int c0 = 0; c0 < g.seg[s].K; c0 += 64) {
      const int kv = g.seg[s].K - c0 < 64 ? g.seg[s].K - c0 : 64;
      p.chunk[nch++] = (uint32_t)s | ((uint32_t)kv << 8) | ((uint32_t)(c0 / 8) << 16) | ((uint32_t)(cbase + c0) << 24);
    }
    cbase += g.seg[s].K;
  }
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("LCM_X_TIMELINE"); dbg = (e && atoi(e)) ? 1 : 0; } p.debug = dbg; }
  { static int d2 = -1; if (d2 < 0) { const char* e = getenv("LCM_X_DBG"); d2 = e ? atoi(e) : 0; } p.dbg2 = d2; }
  if (ensure_dyn_smem_fn(gemm_expand_kernel<false>, kSmemLimitX) || ensure_dyn_smem_fn(gemm_expand_kernel<true>, kSmemLimitX) ||
      ensure_dyn_smem_fn(expand_stats_kernel<128>, (128 * 128 + 128 + 128 * 32 + 512) * 4) ||
      ensure_dyn_smem_fn(expand_stats_kernel<64>, (64 * 64 + 64 + 64 * 32 + 512) * 4)) return -2;
  const int grid = p.m_tiles < num_sms ? p.m_tiles : num_sms;
  if (zero_scratch) {
    if (cudaMemsetAsync(scratch, 0, gemm_expand_scratch_bytes(images), st) != cudaSuccess) return -2;
    if (stats_only) gemm_expand_kernel<true><<<grid, kThreadsX, L.total, st>>>(p);
    else gemm_expand_kernel<false><<<grid, kThreadsX, L.total, st>>>(p);
  } else if (stats_only) {
    launch_pdl(gemm_expand_kernel<true>, dim3(grid), dim3(kThreadsX), L.total, st, p);
  } else {
    launch_pdl(gemm_expand_kernel<false>, dim3(grid), dim3(kThreadsX), L.total, st, p);
  }
  const dim3 sg(images, g.Nc / 32);
  if (L.nchunks == 1)
    launch_pdl(expand_stats_kernel<64>, sg, dim3(128), (size_t)(64 * 64 + 64 + 64 * 32 + 512) * 4, st, p.gram, p.colsum, p.W, g.stats, g.Nc, L.nchunks);
  else
    launch_pdl(expand_stats_kernel<128>, sg, dim3(128), (size_t)(128 * 128 + 128 + 128 * 32 + 512) * 4, st, p.gram, p.colsum, p.W, g.stats, g.Nc, L.nchunks);
  return 0;
}

}  // namespace lcm
