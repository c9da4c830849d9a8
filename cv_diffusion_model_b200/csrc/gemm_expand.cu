// Expand GEMM of the inverted-residual block (efficient_unet.py:174,207-209) on the tensor-core path:
//
//   h1[m][n] = sum_k relu6(a_k * x[m][k] + b_k) * W[n][k]        x: bf16 residual stream (1-2 concat segments)
//                                                                 h1: fp16 hidden tensor + per-(image, channel) sum / sum^2
//
// K is small (32...128) and N = 4K, so this op is a pure streaming problem — (K + N) * 2 bytes per pixel, ~40 FLOP/B —
// and the only thing that can keep it off the HBM roofline is the per-element instruction count of the epilogue.
// The general kernel (gemm_tc2.cu) spends ~9000 warp instructions per 128x128 tile, most of them in the CUDA-core
// copy-out + column-statistics role; this kernel removes that role entirely:
//
//   * statistics on the tensor core: the epilogue stages the fp16 tile AND its element-wise square in shared memory
//     (128-byte swizzled [128 rows][64 cols] chunks); one more tcgen05.mma per 16 rows, D_stat[128][16] +=
//     [x | x^2]^T (MN-major A operand straight from the staging chunks) * ones[16 rows][16], accumulates the 64 column
//     sums and 64 sums of squares of the n-block in TMEM lanes 0-63 / 64-127.  They stay in TMEM across all tiles of an
//     image and are flushed with ONE fp64 atomic per (CTA, image, channel, moment).
//   * the global store is a TMA store of the same staging chunk: no per-element store instructions.
//   * prologue: relu6(a x + b) = 6 sat(a/6 x + b/6): one FFMA.SAT per element, the 6 is folded into the packed weights.
//   * an m-tile (128 pixels) is multiplied by ALL n-blocks of the resident weights, so x is read exactly once.
//
// Roles (512 threads, one persistent CTA per SM):
//   warps 0-3, 4-7  E0, E1  two epilogue groups, n-block g belongs to group g & 1 (its own accumulator and staging
//                           buffer): tcgen05.ld 64 accumulator columns -> fp16 x, x^2 -> staging; the group's first
//                           thread issues the TMA store.  One group alone is latency-bound (ld -> cvt -> sts -> fence
//                           -> barrier chain), two overlap.
//                           thread issues the TMA store AND the statistics MMAs of its n-block, so nothing else ever
//                           waits for the staging tile.
//   warp  8    MMA main MMAs (M=128, N=64, K=16), n-block after n-block
//   warp  9    TMA weights once; activation chunks (128 rows x 64 channels, 128-byte swizzle) ring
//   warps 10-15 XF in-place prologue on the landed chunk
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstring>
#include <mutex>
#include <unordered_map>

#include "kernels.h"
#include "tc_common.cuh"
#include "tmap.h"

namespace lcm {

namespace {

using namespace tc;

constexpr int kThreadsX = 512;   // 16 warps: 4 per register-file partition, 128 registers each
constexpr int kXfBaseX = 320, kXfThreadsX = 192;
constexpr int kMmaWarp = 8, kTmaWarp = 9, kXfWarp0 = 10;
constexpr uint32_t kChunkBytes = 16384;     // 128 rows x 64 16-bit elements
constexpr uint32_t kWChunkBytes = 8192;     // 64 output channels x 64 k
constexpr uint32_t kSmemLimitX = 232448;
constexpr int kMaxNB = 8;                   // n-blocks of 64 output channels
constexpr int kMaxChunksX = 4;              // 64-wide K chunks per tile
constexpr uint32_t kStatCol0 = 128;         // TMEM: columns 0-127 = two 64-column accumulators, then 16 per n-block

struct XParams {
  CUtensorMap tmap_in[2];
  CUtensorMap tmap_out;
  const float2* coef[2];
  int coef_ld[2], coef_off[2], segK[2];
  int nseg, nchunks, NB, stages;
  const __half* W;          // packed [n-block][chunk][64 rows x 64 k] (128-byte swizzled rows), scaled by 6
  double* stats;            // [images][Nc][2]
  int m_tiles, P, Nc, ncoef;
  uint32_t chunk[kMaxChunksX];   // seg | kvalid << 8 | c0 << 16 | coef base << 24
  uint32_t w_off, stg_off, ones_off, coef_smem_off, misc_off;
  int debug;
};

__device__ __forceinline__ void tma_load_2d_x(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// MN-major operand, 128-byte swizzle: 64 elements (128 B) contiguous along MN, 8-row groups along K 1024 B apart,
// 64-wide MN blocks `lbo` bytes apart
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t lbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// K-major, no swizzle: 8-row x 16-byte core matrices, 128 B each, consecutive along K then along the rows
__device__ __forceinline__ uint64_t umma_desc_plain(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (8ull << 16) | (16ull << 32) | (1ull << 46);
}
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, uint32_t& r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr));
}

__device__ long long g_xtimeline[64 * 16];   // LCM_X_TIMELINE: clock64 stamps of block 0, first 64 tiles
#define XSTAMP(tile_idx, slot) do { if (p.debug && blockIdx.x == 0 && (tile_idx) < 64) g_xtimeline[(tile_idx) * 16 + (slot)] = clock64(); } while (0)

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
  auto sfree_bar = [&](int b) { return bar0 + 8u * (30 + b); };
  const uint32_t wres_bar = bar0 + 8u * 32;
  const uint32_t sread_bar = bar0 + 8u * 33;   // statistics read out of TMEM (image flush)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + p.misc_off + 320);
  float2* s_coef = reinterpret_cast<float2*>(smem + p.coef_smem_off);

  if (warp == kTmaWarp && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(raw_bar(s), 1); mbar_init(xf_bar(s), kXfThreadsX); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 128);
      mbar_init(sfree_bar(a), 1);
    }
    mbar_init(wres_bar, 1);
    mbar_init(sread_bar, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int s = 0; s < p.nseg; ++s)
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmap_in[s])) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmap_out)) : "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  // ones operand of the statistics MMA: 512 B of fp16 1.0
  if (tid < 128) reinterpret_cast<uint32_t*>(smem + p.ones_off)[tid] = 0x3C003C00u;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int t_begin = (int)((long long)p.m_tiles * blockIdx.x / gridDim.x);
  const int t_end = (int)((long long)p.m_tiles * (blockIdx.x + 1) / gridDim.x);
  const int tiles_per_img = p.P >> 7;

  if (warp >= kXfWarp0) {
    // ================================ XF: relu6(a x + b) / 6 in place =====================================
    const int xt = tid - kXfBaseX;
    int stage = 0; uint32_t phase = 0;
    int cur_img = -1;
    for (int t = t_begin; t < t_end; ++t) {
      const int img = t / tiles_per_img;
      if (img != cur_img) {
        bar_sync(1, kXfThreadsX);   // everyone is done with the previous image's coefficients
        for (int s = 0, base = 0; s < p.nseg; base += p.segK[s], ++s) {
          const float2* src = p.coef[s] + (size_t)img * p.coef_ld[s] + p.coef_off[s];
          for (int k = xt; k < p.segK[s]; k += kXfThreadsX) {
            const float2 c = src[k];
            s_coef[base + k] = make_float2(c.x * (1.f / 6.f), c.y * (1.f / 6.f));
          }
        }
        bar_sync(1, kXfThreadsX);
        cur_img = img;
      }
      for (int ci = 0; ci < p.nchunks; ++ci) {
        const uint32_t cd = p.chunk[ci];
        const int kvalid = (cd >> 8) & 0xff, cbase = cd >> 24;
        const uint32_t a_smem = sbase + (uint32_t)stage * kChunkBytes;
        // unit u = xt + 192 i: row = u >> 3, slot = u & 7 holds channel unit cu = slot ^ (row & 7); (row & 7) does not
        // depend on i (192 = 24 rows), so a thread transforms the same 8 channels in all of its rows
        const int cu = (xt & 7) ^ ((xt >> 3) & 7);
        const bool act = cu * 8 < kvalid;
        float2 ab[8];
        if (act) {
          const float4* c4 = reinterpret_cast<const float4*>(s_coef + cbase + cu * 8);
#pragma unroll
          for (int j = 0; j < 4; ++j) { const float4 c = c4[j]; ab[2 * j] = make_float2(c.x, c.y); ab[2 * j + 1] = make_float2(c.z, c.w); }
        }
        mbar_wait(raw_bar(stage), phase);
        if (ci == 0 && xt == 0) XSTAMP(t - t_begin, 1);
        if (act) {
          uint4 v[6];
#pragma unroll
          for (int i = 0; i < 6; ++i)
            if (xt + i * kXfThreadsX < 1024) v[i] = lds128(a_smem + (uint32_t)(xt + i * kXfThreadsX) * 16u);
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            if (xt + i * kXfThreadsX < 1024) {
              float f[8];
              unpack8(v[i], f);
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = __saturatef(fmaf(ab[j].x, f[j], ab[j].y));
              sts128(a_smem + (uint32_t)(xt + i * kXfThreadsX) * 16u, pack8(f));
            }
          }
          fence_proxy_async();
        }
        mbar_arrive(xf_bar(stage));
        if (ci == p.nchunks - 1 && xt == 0) XSTAMP(t - t_begin, 2);
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == kTmaWarp) {
    // ================================ TMA: weights once, activation chunks ================================
    if (elect_one()) {
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
        mbar_wait(empty_bar(stage), phase ^ 1u);
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
    // D = f32, A/B = bf16 (main) ; M = 128, N = 64
    const uint32_t idesc_main = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | (8u << 24);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    mbar_wait(wres_bar, 0);
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
      for (int j = 0; j < p.NB; ++j, ++g) {
        const int acc = g & 1;
        mbar_wait(tempty_bar(acc), ((uint32_t)(g >> 1) & 1u) ^ 1u);
        tc_fence_after();
        if (elect_one()) {
          int s2 = stage;
          for (int ci = 0; ci < p.nchunks; ++ci) {
            const int ksteps = (int)((p.chunk[ci] >> 8) & 0xff) >> 4;
            const uint64_t ad = umma_desc(sbase + (uint32_t)s2 * kChunkBytes);
            const uint64_t bd = umma_desc(sbase + p.w_off + (uint32_t)(j * p.nchunks + ci) * kWChunkBytes);
            for (int k = 0; k < ksteps; ++k)
              umma_bf16(tmem_u + (uint32_t)acc * 64u, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc_main, (ci | k) != 0 ? 1u : 0u);
            if (++s2 == p.stages) s2 = 0;
          }
          umma_commit(tfull_bar(acc));
          if (j == p.NB - 1) {   // last reader of this tile's activation chunks: hand the stages back
            int s3 = stage;
            for (int ci = 0; ci < p.nchunks; ++ci) { umma_commit(empty_bar(s3)); if (++s3 == p.stages) s3 = 0; }
          }
        }
        __syncwarp();
      }
      if (lane == 0) XSTAMP(t - t_begin, 4);
      for (int ci = 0; ci < p.nchunks; ++ci) if (++stage == p.stages) { stage = 0; phase ^= 1u; }
    }
  } else {
    // ================================ E0 / E1: accumulator -> fp16 x, x^2 -> staging -> TMA store ============
    // the warp index is broadcast so that the compiler keeps everything derived from it (group, loop counters, TMA /
    // MMA operands) in uniform registers: issuing tcgen05.mma from per-thread registers costs ~150 cycles each (R2UR)
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    const int grp = warp_u >> 2;               // owns n-blocks g with (g & 1) == grp: accumulator grp, staging buffer grp
    const int ew = warp_u & 3;                 // TMEM lane quadrant = warp % 4
    const int et = ew * 32 + lane;
    const int r7 = et & 7;
    const uint32_t row_off = (uint32_t)et * 128u + ((uint32_t)r7 << 4);   // XOR with (unit << 4) gives the swizzled slot
    const uint32_t lane_base = tmem_base + ((uint32_t)(ew * 32) << 16);
    const uint32_t xs0 = sbase + p.stg_off + (uint32_t)grp * 2u * kChunkBytes;
    const int total_g = (t_end - t_begin) * p.NB;
    int cur_img = -1;
    // statistics MMA (issued by the group's first thread right after the staging tile is complete):
    // A/B = f16, A is MN-major (bit 15), M = 128, N = 16
    const uint32_t idesc_stat = (1u << 4) | (1u << 15) | ((uint32_t)(16 >> 3) << 17) | (8u << 24);
    const uint64_t ones_desc = umma_desc_plain(sbase + p.ones_off);
    uint32_t sread_phase = 0;     // image flushes waited for so far (leaders only)
    int waited_t = -1;
    // image flush (group 0 only): each group's statistics MMAs complete in order, so all of them up to n-block `gl`
    // have completed once the last one of either group (gl and gl - 1) has
    auto flush = [&](int img, int gl) {
      mbar_wait(sfree_bar(gl & 1), (uint32_t)(gl >> 1) & 1u);
      if (gl > 0) mbar_wait(sfree_bar((gl - 1) & 1), (uint32_t)((gl - 1) >> 1) & 1u);
      tc_fence_after();
      const int which = et >> 6, col = et & 63;
      for (int j = 0; j < p.NB; ++j) {
        uint32_t v;
        tmem_ld1(lane_base + kStatCol0 + 16u * j, v);
        tmem_wait_ld();
        atomicAdd(&p.stats[((size_t)img * p.Nc + j * 64 + col) * 2 + which], (double)__uint_as_float(v));
      }
      tc_fence_before();
      mbar_arrive(sread_bar);
    };
    int t = t_begin, j = grp;
    int img = t_begin / tiles_per_img, tin = t_begin - img * tiles_per_img;   // image of tile t, tile index inside it
    auto advance = [&]() { while (j >= p.NB) { j -= p.NB; ++t; if (++tin == tiles_per_img) { tin = 0; ++img; } } };
    advance();
    for (int g = grp; g < total_g; g += 2) {
      if (grp == 0 && img != cur_img) { if (cur_img >= 0) flush(cur_img, (t - t_begin) * p.NB - 1); cur_img = img; }
      const uint32_t use = (uint32_t)(g >> 1) & 1u;
      mbar_wait(tfull_bar(grp), use);
      if (et == 0 && j < 2) XSTAMP(t - t_begin, 5 + 5 * j);
      tc_fence_after();
      const uint32_t taddr = lane_base + (uint32_t)grp * 64u;
      uint32_t r[4][16];
      tmem_ld16(taddr, r[0]); tmem_ld16(taddr + 16, r[1]); tmem_ld16(taddr + 32, r[2]); tmem_ld16(taddr + 48, r[3]);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(tempty_bar(grp));                    // accumulator drained
      if (et == 0 && j < 2) XSTAMP(t - t_begin, 6 + 5 * j);
      if (ew == 0 && elect_one()) bulk_wait_read<0>(); // this group's previous store has read the staging buffer
      mbar_wait(sfree_bar(grp), use ^ 1u);             // ... and so has its statistics MMA
      bar_sync(3 + grp, 128);
      if (et == 0 && j < 2) XSTAMP(t - t_begin, 7 + 5 * j);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t x[4], sq[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            x[e] = pack_f16(__uint_as_float(r[q][h * 8 + 2 * e]), __uint_as_float(r[q][h * 8 + 2 * e + 1]));
            const __half2 hx = *reinterpret_cast<__half2*>(&x[e]);
            const __half2 hs = __hmul2(hx, hx);
            sq[e] = *reinterpret_cast<const uint32_t*>(&hs);
          }
          const uint32_t a = (xs0 + row_off) ^ ((uint32_t)(q * 2 + h) << 4);
          sts128(a, make_uint4(x[0], x[1], x[2], x[3]));
          sts128(a + kChunkBytes, make_uint4(sq[0], sq[1], sq[2], sq[3]));
        }
      }
      fence_proxy_async();
      bar_sync(3 + grp, 128);
      if (et == 0 && j < 2) XSTAMP(t - t_begin, 8 + 5 * j);
      if (ew == 0) {   // the group's first warp, in warp-uniform control flow; one elected lane issues
        if (elect_one()) { tma_store_2d(&p.tmap_out, xs0, j * 64, t * 128); bulk_commit(); }
        // the statistics columns of an n-block accumulate over the tiles of one image: the first tile of an image
        // (in this CTA's range) overwrites them — after the previous image's sums have been read out of TMEM
        const bool first_tile = t == t_begin || tin == 0;
        if (first_tile && t != t_begin && t != waited_t) { mbar_wait(sread_bar, sread_phase); sread_phase ^= 1u; waited_t = t; }
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16(tmem_base + kStatCol0 + 16u * j, umma_desc_mn(xs0 + (uint32_t)k * 2048u, kChunkBytes), ones_desc, idesc_stat,
                      (first_tile && k == 0) ? 0u : 1u);
          umma_commit(sfree_bar(grp));
        }
        __syncwarp();
        if (lane == 0 && j < 2) XSTAMP(t - t_begin, 9 + 5 * j);
      }
      j += 2;
      advance();
    }
    if (grp == 0 && cur_img >= 0) flush(cur_img, total_g - 1);
    if (ew == 0 && elect_one()) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
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

struct XLayout { int nchunks, NB, stages; uint32_t w_off, stg_off, ones_off, coef_off, misc_off, total; };

bool plan_layout(int nseg, const int* segK, int Nc, XLayout* L) {
  int nch = 0, ncoef = 0;
  for (int s = 0; s < nseg; ++s) {
    if (segK[s] % 16 || segK[s] < 16 || segK[s] > 248) return false;   // kvalid is stored in 8 bits
    nch += (segK[s] + 63) / 64;
    ncoef += segK[s];
  }
  if (nch < 1 || nch > kMaxChunksX || Nc % 64 || Nc / 64 > kMaxNB || Nc < 64) return false;
  if (ncoef > 255) return false;   // coefficient base is stored in 8 bits
  L->nchunks = nch; L->NB = Nc / 64;
  const uint32_t wbytes = (uint32_t)L->NB * nch * kWChunkBytes;
  const uint32_t fixed = wbytes + 4u * kChunkBytes /* staging: 2 x (x, x^2) */ + 1024 /* ones */ + 2048 /* coef */ + 1024 /* misc */ + 1024 /* align */;
  if (fixed + (uint32_t)(nch + 1) * kChunkBytes > kSmemLimitX) return false;
  int stages = (int)((kSmemLimitX - fixed) / kChunkBytes);
  if (stages > 8) stages = 8;
  L->stages = stages;
  uint32_t off = (uint32_t)stages * kChunkBytes;
  L->w_off = off; off += wbytes;
  L->stg_off = off; off += 4u * kChunkBytes;
  L->ones_off = off; off += 1024;
  L->coef_off = off; off += 2048;
  L->misc_off = off; off += 1024;
  L->total = off + 1024;
  return true;
}

}  // namespace

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

// W: fp16... no: bf16 image packed with block_n = 64 and the x6 scale (PackJob::scale) — see plan.cu / ops_api.cu
int launch_gemm_expand(const GemmParams& g, int num_sms, cudaStream_t st) {
  if (g.nseg < 1 || g.nseg > 2 || !g.out_f16 || !g.stats || g.P % 128 || g.M % 128 || g.M <= 0 || g.M > 0x7fffff00LL) return -1;
  int segK[2] = {0, 0};
  for (int s = 0; s < g.nseg; ++s) {
    if (g.seg[s].mode != XF_AFFINE_RELU6 || g.seg[s].f16 || !g.seg[s].coef || g.seg[s].ld % 8) return -1;
    segK[s] = g.seg[s].K;
  }
  XLayout L;
  if (!plan_layout(g.nseg, segK, g.Nc, &L)) return -1;
  XParams p;
  memset(&p, 0, sizeof(p));
  p.nseg = g.nseg; p.nchunks = L.nchunks; p.NB = L.NB; p.stages = L.stages;
  p.W = reinterpret_cast<const __half*>(g.W);
  p.stats = g.stats;
  p.m_tiles = (int)(g.M / 128); p.P = g.P; p.Nc = g.Nc;
  p.w_off = L.w_off; p.stg_off = L.stg_off; p.ones_off = L.ones_off; p.coef_smem_off = L.coef_off; p.misc_off = L.misc_off;
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
  p.ncoef = cbase;
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("LCM_X_TIMELINE"); dbg = (e && atoi(e)) ? 1 : 0; } p.debug = dbg; }
  if (!matrix_map(g.out, g.M, g.Nc, g.Nc, TMAP_F16, &p.tmap_out)) return -3;
  static bool attr_done = false;
  {
    std::lock_guard<std::mutex> lk(g_x_mu);
    if (!attr_done) {
      if (cudaFuncSetAttribute(gemm_expand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimitX) != cudaSuccess) return -2;
      attr_done = true;
    }
  }
  const int grid = p.m_tiles < num_sms ? p.m_tiles : num_sms;
  gemm_expand_kernel<<<grid, kThreadsX, L.total, st>>>(p);
  return 0;
}

}  // namespace lcm
