// tcgen05 GEMM, second generation: the kernel behind every 1x1 convolution (and, through a gather
// producer, every dense 3x3 convolution) of the EfficientUNet in bf16 mode.
//
//   out[m][n] = sum_k xform(A[m][k]) * W[n][k] (+ bias[n]);  bf16 operands, fp32 accumulation in TMEM,
//   epilogue: bf16 store + per-(image, channel) sum / sum-of-squares for the consumer's GroupNorm.
//
// Most of these GEMMs are HBM-bound (K, N <= 384 at 256^2 / 128^2), so the kernel is organised as a
// deep streaming pipeline in which no role ever waits on global-memory latency with registers:
//
//   warp 5 (1 thread)  TMA      cp.async.bulk.tensor.2d (128-byte swizzle) brings 128 x 64 activation tiles of
//                               each operand segment straight into the UMMA stage layout, up to `stages`
//                               chunks ahead; weights arrive as pre-swizzled bulk copies (resident in smem
//                               when they fit, else streamed with the A chunk).  Rows >= M and channels
//                               >= K are zero-filled by the TMA unit.
//   warps 14-21        XF       in-place prologue on the landed tile (smem -> registers -> smem): GroupNorm /
//                               FiLM affine + ReLU6, or the SE gate; operands that need no transform (the
//                               residual / skip-conv input) skip this stage entirely.  For 3x3 convolutions
//                               these warps gather the taps instead (zero padding, stride 2, bilinear x2).
//   warp 4 (1 thread)  MMA      tcgen05.mma M=128, N=block_n, K=16 per 32 bytes of K; tcgen05.commit frees
//                               the stage and publishes the accumulator (two accumulators in TMEM).
//   warps 0-3          E1       tcgen05.ld -> (+bias) -> bf16 -> padded smem staging tile (double-buffered)
//   warps 6-13         E2       staging -> coalesced 16-byte global stores + column statistics; partial sums
//                               are folded without atomics into CTA-resident accumulators that are flushed
//                               with fp64 atomics only when the image (or N tile) changes.
// All loop bookkeeping is incremental 32-bit arithmetic; every wait is a bounded mbarrier wait.
#include <cuda.h>

#include <cstdio>
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

constexpr int kThreads2 = 704;
constexpr int kE2Base = 192, kE2Threads = 256, kXfBase = 448, kXfThreads = 256;
constexpr uint32_t kStageA2 = 16384;
constexpr uint32_t kHaloBytes = 130u * 3u * 128u, kHaloStage = 51200;   // [3 rows][130 px][64 ch] bf16, stage rounded to 1 KB
constexpr int kMaxChunks2 = 160;
constexpr uint32_t kSmemLimit2 = 232448;
constexpr uint32_t kMisc2 = 4096 + 4096;    // barriers + per-column accumulators | statistics flush scratch [1024] fp32

struct alignas(64) Tc2Params {
  CUtensorMap tmap[LCM_MAX_SEGS];
  GemmSeg seg[LCM_MAX_SEGS];
  int coef_base[LCM_MAX_SEGS];
  int nseg, ncoef;
  const bf16* W;
  bf16* out;
  double* stats;
  const float* bias;
  long long M, m_tiles;
  int P, Nc, block_n, n_tiles, nchunks;
  int resident, stages, fast, nbuf;
  int conv_mode, Hin, Win, Hout, Wout, Ci;
  int out_f16;           // store the output (and take its statistics) as fp16
  int wgate;             // SE gate folded into the smem-resident weights per image (XF_SCALE segments stay raw)
  int all_raw;           // no chunk needs the XF stage: the MMA warp consumes TMA tiles directly
  int conv_stride;       // conv_tma: 1, or 2 for the stride-2 convs (tensor map with element strides 2 along W and H)
  int conv_tma, box_w;   // stride-1 3x3 conv fed by 4-D TMA tiles (zero fill = padding); box_w = pixels per tile row
  int coef_compact;      // fp16 SE-gated segments keep only their packed half2 gate pairs in shared memory (4 B per two
                         // channels instead of 16): with streamed weights every KB decides the number of pipeline stages
  int bpair;             // streamed weights shared by a 2-CTA cluster: each CTA fetches half of every weight chunk and
                         // multicasts it to both (halves the L2 -> shared-memory weight traffic, the bound of the 32x32 / 64x64 levels)
  int conv_halo, achunks; // halo mode: ONE [64 ch][130 px][3 rows] load per tile and 64-channel chunk serves all 9 taps
                          // (achunks = activation stages per tile; the weight chunks stay per (tap, chunk))
  uint32_t stage_bytes, bres_off, stg_off, stg_stride, stg_bytes, coef_off, misc_off;
  int debug;
  int prefetch;   // L2 prefetch distance of the activation chunks (0 = off): set when the ring is shorter than a tile
  uint32_t chunk[kMaxChunks2];  // seg/tap (7 bits) | fp16 segment << 7 | kvalid << 8 | c0 << 16
  uint8_t lo_slot[kMaxChunks2]; // wgate: index (after the nchunks weight chunks) of the chunk's low-order weight image, or 0xff
};

__device__ long long g_timeline[64 * 16];   // LCM_TC_DEBUG & 64: per-tile clock64 stamps of block 0
#define TSTAMP(slot) do { if (kDebug && (p.debug & 64) && blockIdx.x == 0 && it < 64) g_timeline[it * 16 + (slot)] = clock64(); } while (0)

__device__ __forceinline__ __half2 h2bits(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t bits_of(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}

// kConv: 3x3 gather producer instead of TMA + in-place prologue; kFast: every 128-row tile lies inside one image
// (P % 128 == 0); kDebug: bottleneck-experiment switches and the timeline (never instantiated on the product path
// unless LCM_TC_DEBUG is set).  Specialising keeps each instantiation's code small: the five roles run different
// code concurrently and the instruction cache is a real constraint.
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}

template <bool kConv, bool kFast, bool kDebug>
__global__ void __launch_bounds__(kThreads2, 1) gemm_tc2_kernel(const __grid_constant__ Tc2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sraw = smem_u32(smem_raw);
  const uint32_t sbase = (sraw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (sbase - sraw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  uint8_t* misc = smem + p.misc_off;
  const uint32_t bar0 = sbase + p.misc_off;
  auto raw_bar = [&](int s) { return bar0 + 8u * s; };                 // up to kMaxStages2 = 16 pipeline stages
  auto xf_bar = [&](int s) { return bar0 + 8u * (16 + s); };
  auto empty_bar = [&](int s) { return bar0 + 8u * (32 + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (48 + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (50 + a); };
  auto sfull_bar = [&](int b) { return bar0 + 8u * (52 + b); };
  auto sempty_bar = [&](int b) { return bar0 + 8u * (54 + b); };
  const uint32_t bres_bar = bar0 + 8u * 56;
  const uint32_t bsc_bar = bar0 + 8u * 57;    // weights rescaled by the SE gate of the current image
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 480);
  float* s_sum = reinterpret_cast<float*>(misc + 512);
  float* s_sq = s_sum + 256;
  float* s_bias = s_sq + 256;
  float* s_scr = reinterpret_cast<float*>(misc + 4096);   // [1024]: statistics flush scratch
  float2* s_coef = reinterpret_cast<float2*>(smem + p.coef_off);

  auto ld_s = [&](uint32_t saddr) { return *reinterpret_cast<const uint4*>(smem + (saddr - sbase)); };
  auto st_s = [&](uint32_t saddr, uint4 v) { *reinterpret_cast<uint4*>(smem + (saddr - sbase)) = v; };

  constexpr bool conv = kConv;
  const int dbg = kDebug ? p.debug : 0;
  if (warp == 5 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(raw_bar(s), 1);
      mbar_init(xf_bar(s), conv ? 128 : kXfThreads);
      mbar_init(empty_bar(s), p.bpair ? 2 : 1);   // pair mode: this CTA's and the peer's MMAs release a stage
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 128);
      mbar_init(sfull_bar(a), 128);
      mbar_init(sempty_bar(a), kE2Threads);
    }
    mbar_init(bres_bar, 1);
    mbar_init(bsc_bar, kXfThreads);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (!conv)
      for (int s = 0; s < (p.conv_tma ? 1 : p.nseg); ++s)
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmap[s])) : "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid < 256) { s_sum[tid] = 0.f; s_sq[tid] = 0.f; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // pair mode: the peer multicasts into this CTA's stages and arrives on its barriers, so they must be initialised
  if (p.bpair) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  pdl_wait();      // everything above is CTA-local; from here on the previous kernel's output is read
  pdl_trigger();

  const int m_tiles = (int)p.m_tiles;
  const long long total_tiles = p.m_tiles * p.n_tiles;
  long long t_begin = total_tiles * blockIdx.x / gridDim.x;
  int my_tiles = (int)(total_tiles * (blockIdx.x + 1) / gridDim.x - t_begin);
  const int tstep = p.bpair ? 2 : 1;
  const uint32_t pair_rank = blockIdx.x & 1u;   // = %cluster_ctarank for a (2,1,1) cluster
  if (p.bpair) {
    // the two CTAs of a cluster take tiles 2u and 2u + 1 of the same pair-unit u: same n tile (m_tiles is even), same
    // chunk sequence, equal tile counts — they advance in lockstep through the shared weight stream
    const long long units = total_tiles >> 1, npairs = gridDim.x >> 1, pr = blockIdx.x >> 1;
    const long long u0 = units * pr / npairs, u1 = units * (pr + 1) / npairs;
    t_begin = 2 * u0 + pair_rank;
    my_tiles = (int)(u1 - u0);
  }
  const uint32_t b_chunk_bytes = (uint32_t)p.block_n * 128u;
  const int M = (int)p.M;

  if (warp >= 14) {
    // ================================ XF: prologue transform / conv gather ==========================
    const int ptid = tid - kXfBase;
    const int group = ptid >> 7, gt = ptid & 127;
    int cur_img = -1, gate_img = -1, gate_nt = -1;
    uint32_t gate_phase = 0;
    TileIter ti; ti.init(t_begin, m_tiles, p.P, tstep);
    Ring ring{0, 0u, p.stages};
    int par = 0;
    for (int it = 0; it < my_tiles; ++it, ti.next(m_tiles, p.P)) {
      const int m0 = ti.m0;
      if (p.ncoef > 0 && kFast && ti.img != cur_img) {
        bar_sync(1, kXfThreads);
        for (int s = 0; s < p.nseg; ++s) {
          if (p.seg[s].mode == XF_NONE) continue;
          const float2* src = p.seg[s].coef + (size_t)ti.img * p.seg[s].coef_ld + p.seg[s].coef_off;
          if (p.seg[s].f16 && p.seg[s].mode == XF_SCALE) {
            // fp16 operand with a pure scale: the A-side prologue is one HMUL2 per two channels; the packed gate
            // pair of channels (k, k+1) travels in the unused b slot of entry k (k even)
            for (int k = 2 * ptid; k < p.seg[s].K; k += 2 * kXfThreads) {
              const float2 c0v = src[k], c1v = src[k + 1];
              const __half2 gp = __floats2half2_rn(c0v.x, c1v.x);
              if (p.coef_compact) {
                reinterpret_cast<uint32_t*>(s_coef + p.coef_base[s])[k >> 1] = *reinterpret_cast<const uint32_t*>(&gp);
              } else {
                s_coef[p.coef_base[s] + k] = make_float2(c0v.x, __uint_as_float(*reinterpret_cast<const uint32_t*>(&gp)));
                s_coef[p.coef_base[s] + k + 1] = make_float2(c1v.x, 0.f);
              }
            }
          } else {
            const float sc = p.seg[s].mode == XF_AFFINE_RELU6 ? (1.f / 6.f) : 1.f;   // relu6 as 6 sat(.): see apply_xform
            for (int k = ptid; k < p.seg[s].K; k += kXfThreads) { const float2 c = src[k]; s_coef[p.coef_base[s] + k] = make_float2(c.x * sc, c.y * sc); }
          }
        }
        bar_sync(1, kXfThreads);
        cur_img = ti.img;
      }
      if (p.wgate && (ti.img != gate_img || ti.n_tile != gate_nt)) {
        // fold gate[img][k] into the freshly (re)loaded weight tile: W'[co][k] = W[co][k] * gate[k]
        mbar_wait(bres_bar, gate_phase);
        gate_phase ^= 1u;
        for (int ci = 0; ci < p.nchunks; ++ci) {
          const uint32_t cd = p.chunk[ci];
          const int sidx = cd & 0x7f, kvalid = (cd >> 8) & 0xff, c0 = cd >> 16;
          if (p.seg[sidx].mode != XF_SCALE) continue;
          const uint32_t b_smem = sbase + p.bres_off + ci * b_chunk_bytes;
          for (int u = ptid; u < p.block_n * 8; u += kXfThreads) {
            const int co = u >> 3, cu = (u & 7) ^ (co & 7);
            if (cu * 8 < kvalid) {
              // W' = W * gate is kept to ~16 bits as a bf16 pair (hi, lo): both images are multiplied by the
              // same A tile, so folding the gate into the weights costs no accuracy (it removes the rounding
              // of gate * h2 that the A-side prologue would add).
              float f[8], l[8];
              const bool h16 = p.seg[sidx].f16 != 0;   // weights of an fp16 segment are packed as fp16
              const uint4 wraw = ld_s(b_smem + (uint32_t)u * 16u);
              if (h16) unpack8h(wraw, f); else unpack8(wraw, f);
              const float2* gk = s_coef + p.coef_base[sidx] + c0 + cu * 8;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                f[j] *= gk[j].x;
                l[j] = f[j] - (h16 ? __half2float(__float2half_rn(f[j])) : __bfloat162float(__float2bfloat16_rn(f[j])));
              }
              st_s(b_smem + (uint32_t)u * 16u, h16 ? pack8h(f) : pack8(f));
              // bf16 weights carry a low-order image so that W * gate keeps ~16 bits; fp16 weights (11 bits, finer than
              // every bf16 operand around them) do not need one
              if (p.lo_slot[ci] != 0xff)
                st_s(sbase + p.bres_off + (uint32_t)(p.nchunks + p.lo_slot[ci]) * b_chunk_bytes + (uint32_t)u * 16u, pack8(l));
            }
          }
        }
        fence_proxy_async();
        mbar_arrive(bsc_bar);
        gate_img = ti.img;
        gate_nt = ti.n_tile;
      }
      if (p.all_raw) continue;   // nothing per chunk: the MMA warp reads the TMA tiles directly
      int cy = 0, cx = 0, cn = 0;
      const int cm = m0 + gt;
      if constexpr (conv) {
        const int q = cm / p.Wout;
        cx = cm - q * p.Wout;
        cn = q / p.Hout;
        cy = q - cn * p.Hout;
      }
      for (int ci = 0; ci < p.nchunks; ++ci, ring.advance()) {
        const uint32_t cd = p.chunk[ci];
        const int sidx = cd & 0x7f, kvalid = (cd >> 8) & 0xff, c0 = cd >> 16;
        const int upr = kvalid >> 3;
        const int stage = ring.stage;
        const uint32_t a_smem = sbase + stage * p.stage_bytes;
        if constexpr (!conv) {
          int mode = p.conv_tma ? (int)XF_NONE : p.seg[sidx].mode;
          if (p.wgate && mode == XF_SCALE) mode = XF_NONE;
          // all 256 prologue threads work on every chunk (4 units each): with few pipeline stages the LATENCY of a
          // chunk through this role matters, not just its throughput
          // Every chunk passes through this stage so that xf_bar completes exactly one phase per use of the
          // stage (an mbarrier must never run two phases ahead of its waiter); operands that need no
          // transform (residual / skip input) are only handed on.
          mbar_wait(raw_bar(stage), ring.phase);  // TMA bytes landed (swizzled: unit (row, cu) sits at slot cu ^ (row & 7))
          if (ci == 0 && ptid == 0) TSTAMP(1);
          if (mode != XF_NONE && !(dbg & 16)) {
            const GemmSeg& sg = p.seg[sidx];
            // unit u = ptid + 256 i: row = u >> 3, smem slot = u & 7 holds channel unit cu = slot ^ (row & 7).  256 units
            // are 32 rows, so (row & 7) — and with it cu — is the same for all 4 units of a thread: its 8 channels'
            // coefficients are loaded ONCE per chunk into registers (they used to be re-read from shared memory per
            // unit with 4-way bank conflicts: ~5000 cycles per chunk, the bottleneck of every K >= 128 expand GEMM).
            const int cu = (ptid & 7) ^ ((ptid >> 3) & 7);
            if (ci == 0 && ptid == 0) TSTAMP(11);
            if (cu < upr) {
              uint4 v[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) v[i] = ld_s(a_smem + (uint32_t)(ptid + i * 256) * 16u);
              if (kFast && sg.f16 && mode == XF_SCALE) {
                __half2 g0, g1, g2, g3;
                if (p.coef_compact) {
                  const uint4 gp4 = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint32_t*>(s_coef + p.coef_base[sidx]) + ((c0 + cu * 8) >> 1));
                  g0 = h2bits(gp4.x); g1 = h2bits(gp4.y); g2 = h2bits(gp4.z); g3 = h2bits(gp4.w);
                } else {
                  const float4* c4 = reinterpret_cast<const float4*>(s_coef + p.coef_base[sidx] + c0 + cu * 8);
                  g0 = h2bits(__float_as_uint(c4[0].y)); g1 = h2bits(__float_as_uint(c4[1].y));
                  g2 = h2bits(__float_as_uint(c4[2].y)); g3 = h2bits(__float_as_uint(c4[3].y));
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const __half2* hv = reinterpret_cast<const __half2*>(&v[i]);
                  st_s(a_smem + (uint32_t)(ptid + i * 256) * 16u,
                       make_uint4(bits_of(__hmul2(hv[0], g0)), bits_of(__hmul2(hv[1], g1)), bits_of(__hmul2(hv[2], g2)), bits_of(__hmul2(hv[3], g3))));
                }
              } else if (kFast) {
                __align__(16) float2 ab[8];
                const float4* c4 = reinterpret_cast<const float4*>(s_coef + p.coef_base[sidx] + c0 + cu * 8);
#pragma unroll
                for (int j = 0; j < 4; ++j) { const float4 c = c4[j]; ab[2 * j] = make_float2(c.x, c.y); ab[2 * j + 1] = make_float2(c.z, c.w); }
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  st_s(a_smem + (uint32_t)(ptid + i * 256) * 16u, (dbg & 256) ? v[i] : apply_xform(v[i], ab, mode, sg.f16 != 0, true));
              } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int row = (ptid + i * 256) >> 3;
                  const int mrow = min(m0 + row, M - 1);
                  const int img = mrow / p.P;
                  __align__(16) float2 ab[8];
                  const float2* src = sg.coef + (size_t)img * sg.coef_ld + sg.coef_off + c0 + cu * 8;
#pragma unroll
                  for (int j = 0; j < 8; ++j) ab[j] = src[j];
                  st_s(a_smem + (uint32_t)(ptid + i * 256) * 16u, apply_xform(v[i], ab, mode, sg.f16 != 0, false));
                }
              }
            }
            if (ci == 0 && ptid == 0) TSTAMP(12);
            if (!(dbg & 128)) fence_proxy_async();
            if (ci == 0 && ptid == 0) TSTAMP(13);
          }
          mbar_arrive(xf_bar(stage));
          if (ci == 0 && ptid == 0) TSTAMP(2);
        } else {
          par ^= 1;
          if (par != (group ^ 1)) continue;
          mbar_wait(empty_bar(stage), ring.phase ^ 1u);
          // ---- 3x3 tap gather: one output pixel (row) per thread -------------------------------------
          const int tap = sidx, ky = tap / 3, kx = tap - ky * 3;
          const bf16* in = reinterpret_cast<const bf16*>(p.seg[0].A);
          const uint32_t rbase = a_smem + gt * 128;
          const int sw = gt & 7;
          if (p.conv_mode == CONV_UP2) {
            const int uy = cy + ky - 1, ux = cx + kx - 1;
            const bool ok = cm < M && uy >= 0 && uy < p.Hout && ux >= 0 && ux < p.Wout;
            if (ok) {
              // F.interpolate(scale_factor=2, bilinear, align_corners=False): src = max(dst/2 - 0.25, 0)
              const float sy = fmaxf(uy * 0.5f - 0.25f, 0.f), sx = fmaxf(ux * 0.5f - 0.25f, 0.f);
              const int y0 = (int)sy, x0 = (int)sx;
              const int y1 = min(y0 + 1, p.Hin - 1), x1 = min(x0 + 1, p.Win - 1);
              const float ly = sy - y0, lx = sx - x0;
              const bf16* b0 = in + ((long long)(cn * p.Hin + y0) * p.Win) * p.Ci + c0;
              const bf16* b1 = in + ((long long)(cn * p.Hin + y1) * p.Win) * p.Ci + c0;
              for (int cu = 0; cu < upr; ++cu) {
                float a[8], b[8], c[8], d[8], o[8];
                unpack8(ldg_cached(b0 + (long long)x0 * p.Ci + cu * 8), a);
                unpack8(ldg_cached(b0 + (long long)x1 * p.Ci + cu * 8), b);
                unpack8(ldg_cached(b1 + (long long)x0 * p.Ci + cu * 8), c);
                unpack8(ldg_cached(b1 + (long long)x1 * p.Ci + cu * 8), d);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  o[j] = (1.f - ly) * ((1.f - lx) * a[j] + lx * b[j]) + ly * ((1.f - lx) * c[j] + lx * d[j]);
                sts128(rbase + ((cu ^ sw) << 4), pack8(o));
              }
            } else {
              for (int cu = 0; cu < upr; ++cu) sts128(rbase + ((cu ^ sw) << 4), make_uint4(0u, 0u, 0u, 0u));
            }
          } else {
            const int st = p.conv_mode == CONV_S2 ? 2 : 1;
            const int iy = cy * st + ky - 1, ix = cx * st + kx - 1;
            const bool ok = cm < M && iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win;
            const bf16* src = in + ((long long)(cn * p.Hin + iy) * p.Win + ix) * p.Ci + c0;
            uint4 v[8];
#pragma unroll
            for (int cu = 0; cu < 8; ++cu) v[cu] = (ok && cu < upr) ? ldg_cached(src + cu * 8) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int cu = 0; cu < 8; ++cu)
              if (cu < upr) sts128(rbase + ((cu ^ sw) << 4), v[cu]);
          }
          fence_proxy_async();
          mbar_arrive(xf_bar(stage));
        }
      }
    }
  } else if (warp == 5) {
    // ================================ TMA: activations and weights ====================================
    // The whole warp runs the (uniform) control flow so that addresses and barriers live in uniform
    // registers; one elected lane issues the copies.
    {
      int cur_nt = -1, cur_img = -1;
      TileIter ti; ti.init(t_begin, m_tiles, p.P, tstep);
      Ring ring{0, 0u, p.stages};
      int acc = 0; uint32_t aphase = 0;   // accumulator stage / phase of the PREVIOUS tile
      int last_stage = -1; uint32_t last_phase = 0;   // ring slot / phase of the last activation chunk issued
      for (int it = 0; it < my_tiles; ++it, ti.next(m_tiles, p.P)) {
        const bf16* wt = p.W + (size_t)ti.n_tile * p.nchunks * p.block_n * 64;
        if (p.resident) {
          // conv + resident weights: this warp issues nothing per tile, so it follows the MMA warp tile by tile (a
          // parity wait is only meaningful when the waiter is less than one phase behind)
          if (conv && it > 0) mbar_wait(tfull_bar(acc), aphase);
          if (ti.n_tile != cur_nt || (p.wgate && ti.img != cur_img)) {
            // The previous tile's MMAs must be done before its weights are overwritten.  This warp runs up to `stages`
            // chunks (several tiles when K is small) ahead of the MMA warp, so it must NOT wait on tfull with a parity:
            // two phases behind looks like "done" (that was a race on the K = 160 project GEMMs).  The empty barrier of
            // the last chunk it issued is exact: it is committed after that tile's last MMA, and this warp is never
            // more than one phase away from it.
            if (!conv && last_stage >= 0) mbar_wait(empty_bar(last_stage), last_phase);
            if (elect_one()) {
              mbar_expect_tx(bres_bar, (uint32_t)p.nchunks * b_chunk_bytes);
              for (int ci = 0; ci < p.nchunks; ++ci)
                bulk_g2s(sbase + p.bres_off + ci * b_chunk_bytes, wt + (size_t)ci * p.block_n * 64, b_chunk_bytes, bres_bar);
            }
            __syncwarp();
            cur_nt = ti.n_tile;
            cur_img = ti.img;
          }
          if (it > 0) { acc ^= 1; if (acc == 0) aphase ^= 1u; }
        }
        if (conv && p.resident) continue;   // nothing per chunk: the gather warps fill A themselves
        int ty0 = 0, tx0 = 0;               // conv_tma: first pixel of this tile inside its image
        if (p.conv_tma) { ty0 = ti.rem / p.Wout; tx0 = ti.rem - ty0 * p.Wout; }   // first OUTPUT pixel of the tile
        for (int ci = 0; ci < p.achunks; ++ci, ring.advance()) {
          const int stage = ring.stage;
          const uint32_t cd = p.chunk[ci];
          const uint32_t a_smem = sbase + stage * p.stage_bytes;
          mbar_wait_relaxed(empty_bar(stage), ring.phase ^ 1u);
          if (ci == 0) TSTAMP(0);
          const bool load_a = !conv && !(dbg & 1);
          if (elect_one()) {
            mbar_expect_tx(raw_bar(stage), (load_a ? (p.conv_halo ? kHaloBytes : kStageA2) : 0u) + (p.resident ? 0u : b_chunk_bytes));
            if (load_a) {
              if (p.conv_halo) {
                tma_load_4d(a_smem, &p.tmap[0], ci * 64, tx0 - 1, ty0 - 1, ti.img, raw_bar(stage));
              } else if (p.conv_tma) {
                const int tap = cd & 0x7f, ky = tap / 3, kx = tap - ky * 3;
                tma_load_4d(a_smem, &p.tmap[0], (int)(cd >> 16), p.conv_stride * tx0 + kx - 1, p.conv_stride * ty0 + ky - 1, ti.img, raw_bar(stage));
              } else {
                tma_load_2d(a_smem, &p.tmap[cd & 0x7f], (int)(cd >> 16), ti.m0, raw_bar(stage));
                if (p.prefetch) {
                  // rings of 2-3 stages (K >= 640) cannot hide the DRAM latency of a 16 KB activation chunk: ask L2 for the
                  // chunk `prefetch` positions ahead (same tile, or the next tile of this CTA in the same n row)
                  int pc = ci + p.prefetch, pm0 = ti.m0;
                  bool ok = true;
                  if (pc >= p.achunks) { pc -= p.achunks; pm0 += 128 * ti.step; ok = it + 1 < my_tiles && ti.m_tile + ti.step < m_tiles && pc < p.achunks; }
                  if (ok) {
                    const uint32_t pd = p.chunk[pc];
                    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(&p.tmap[pd & 0x7f])),
                                 "r"((int)(pd >> 16)), "r"(pm0)
                                 : "memory");
                  }
                }
              }
            }
            if (!p.resident) {
              if (p.bpair) {
                // this CTA's half of the chunk, to the same stage offset (and barrier) of both CTAs of the pair
                const uint32_t half = b_chunk_bytes >> 1;
                asm volatile(
                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
                        a_smem + kStageA2 + pair_rank * half),
                    "l"(reinterpret_cast<const uint8_t*>(wt + (size_t)ci * p.block_n * 64) + pair_rank * half), "r"(half), "r"(raw_bar(stage)),
                    "h"((uint16_t)3)
                    : "memory");
              } else {
                bulk_g2s(a_smem + kStageA2, wt + (size_t)ci * p.block_n * 64, b_chunk_bytes, raw_bar(stage));
              }
            }
          }
          __syncwarp();
          last_stage = stage; last_phase = ring.phase;
        }
      }
    }
  } else if (warp == 4) {
    // ================================ MMA issuer (warp-uniform control, one elected lane issues) =====
    {
      // D = f32; A/B element format (bits 7-9 / 10-12): 1 = bf16, 0 = f16 (chunks of an fp16 segment)
      const uint32_t idesc_h = (1u << 4) | ((uint32_t)(p.block_n >> 3) << 17) | (8u << 24);
      const uint32_t idesc_b = idesc_h | (1u << 7) | (1u << 10);
      int cur_nt = -1, cur_img = -1;
      uint32_t bres_phase = 0;
      TileIter ti; ti.init(t_begin, m_tiles, p.P, tstep);
      Ring ring{0, 0u, p.stages};
      int acc = 0; uint32_t aphase = 0;
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      for (int it = 0; it < my_tiles; ++it, ti.next(m_tiles, p.P)) {
        if (p.resident && (ti.n_tile != cur_nt || (p.wgate && ti.img != cur_img))) {
          mbar_wait(p.wgate ? bsc_bar : bres_bar, bres_phase);
          bres_phase ^= 1u;
          cur_nt = ti.n_tile;
          cur_img = ti.img;
        }
        mbar_wait(tempty_bar(acc), aphase ^ 1u);
        TSTAMP(3);
        tc_fence_after();
        const uint32_t d_tmem = tmem_u + (uint32_t)acc * 256u;
        if (p.conv_halo) {
          // one halo stage per 64-channel chunk; tap (ky, kx) is the same buffer read from pixel (ky * 130 + kx) on —
          // a shift by whole 128-byte rows inside the SWIZZLE_128B pattern the TMA wrote
          const int nchh = p.achunks;
          for (int ci = 0; ci < nchh; ++ci, ring.advance()) {
            const int stage = ring.stage;
            mbar_wait(raw_bar(stage), ring.phase);
            if (ci == 0) TSTAMP(4);
            tc_fence_after();
            const int ksteps = (int)((p.chunk[ci] >> 8) & 0xff) >> 4;
            const uint32_t a_base = sbase + stage * p.stage_bytes;
            if (elect_one()) {
              for (int tap = 0; tap < 9; ++tap) {
                const uint32_t a_addr = a_base + (uint32_t)((tap / 3) * 130 + tap % 3) * 128u;
                // start address shifted by whole 128-B rows: the swizzle phase follows the absolute smem
                // address (measured: base_offset must stay 0, setting (addr>>7)&7 double-counts the shift)
                const uint64_t ad = umma_desc(a_addr);
                const uint64_t bd = umma_desc(sbase + p.bres_off + (uint32_t)(tap * nchh + ci) * b_chunk_bytes);
                for (int k = 0; k < ksteps; ++k)
                  umma_bf16(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc_b, (ci | tap | k) != 0 ? 1u : 0u);
              }
              umma_commit(empty_bar(stage));
              if (ci == nchh - 1) umma_commit(tfull_bar(acc));
            }
            __syncwarp();
            if (ci == nchh - 1) TSTAMP(5);
          }
        } else
        for (int ci = 0; ci < p.nchunks; ++ci, ring.advance()) {
          const int stage = ring.stage;
          const uint32_t cd = p.chunk[ci];
          mbar_wait(p.all_raw ? raw_bar(stage) : xf_bar(stage), ring.phase);
          if (ci == 0) TSTAMP(4);
          if (conv && !p.resident) mbar_wait(raw_bar(stage), ring.phase);   // streamed weights of a 3x3 conv
          tc_fence_after();
          const int ksteps = (cd >> 12) & 0xf;
          const uint32_t a_addr = sbase + stage * p.stage_bytes;
          const uint32_t b_addr = p.resident ? sbase + p.bres_off + ci * b_chunk_bytes : a_addr + kStageA2;
          const uint64_t ad = umma_desc(a_addr), bd = umma_desc(b_addr);
          const bool gated = p.wgate && p.lo_slot[ci] != 0xff;
          const uint32_t idesc = (cd & 0x80u) ? idesc_h : idesc_b;
          const uint64_t bl = umma_desc(sbase + p.bres_off + (uint32_t)(p.nchunks + p.lo_slot[ci]) * b_chunk_bytes);
          if (elect_one()) {
            for (int k = 0; k < ksteps; ++k)
              umma_bf16(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (ci | k) != 0 ? 1u : 0u);
            if (gated)
              for (int k = 0; k < ksteps; ++k) umma_bf16(d_tmem, ad + (uint64_t)(2 * k), bl + (uint64_t)(2 * k), idesc, 1u);
            if (p.bpair)
              asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                               empty_bar(stage)), "h"((uint16_t)3)
                           : "memory");
            else
              umma_commit(empty_bar(stage));
            if (ci == p.nchunks - 1) umma_commit(tfull_bar(acc));
          }
          __syncwarp();
          if (ci == p.nchunks - 1) TSTAMP(5);
        }
        acc ^= 1;
        if (acc == 0) aphase ^= 1u;
      }
    }
  } else if (warp < 4) {
    // ================================ E1: TMEM -> bf16 staging ==========================================
    int cur_nt = -1;
    TileIter ti; ti.init(t_begin, m_tiles, p.P, tstep);
    int acc = 0; uint32_t aphase = 0;
    int sb = 0; uint32_t sphase = 0;
    for (int it = 0; it < my_tiles; ++it, ti.next(m_tiles, p.P)) {
      if (p.bias && ti.n_tile != cur_nt) {
        bar_sync(3, 128);   // previous tile's bias reads are done
        for (int c = tid; c < p.block_n; c += 128) s_bias[c] = p.bias[ti.n_tile * p.block_n + c];
        bar_sync(3, 128);
      }
      cur_nt = ti.n_tile;
      mbar_wait(tfull_bar(acc), aphase);
      if (tid == 0) TSTAMP(6);
      mbar_wait(sempty_bar(sb), sphase ^ 1u);
      if (tid == 0) TSTAMP(7);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)acc * 256u;
      const uint32_t my_row = sbase + p.stg_off + (uint32_t)sb * p.stg_bytes + (uint32_t)tid * p.stg_stride;
      for (int cb = 0; cb < ((dbg & 32) ? 0 : p.block_n); cb += 32) {
        uint32_t r0[16], r1[16];
        const bool two = cb + 16 < p.block_n;
        tmem_ld16(taddr + cb, r0);
        if (two) tmem_ld16(taddr + cb + 16, r1);
        tmem_wait_ld();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (h == 1 && !two) break;
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(h == 0 ? r0[j] : r1[j]);
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] += s_bias[cb + h * 16 + j];
          }
          const uint32_t dst = my_row + (cb + h * 16) * 2;
          if (p.out_f16) {
            st_s(dst, make_uint4(pack_f16(f[0], f[1]), pack_f16(f[2], f[3]), pack_f16(f[4], f[5]), pack_f16(f[6], f[7])));
            st_s(dst + 16, make_uint4(pack_f16(f[8], f[9]), pack_f16(f[10], f[11]), pack_f16(f[12], f[13]), pack_f16(f[14], f[15])));
          } else {
            st_s(dst, make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7])));
            st_s(dst + 16, make_uint4(pack_bf16(f[8], f[9]), pack_bf16(f[10], f[11]), pack_bf16(f[12], f[13]), pack_bf16(f[14], f[15])));
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));   // accumulator drained
      mbar_arrive(sfull_bar(sb));     // staging tile complete (release)
      if (tid == 0) TSTAMP(8);
      acc ^= 1; if (acc == 0) aphase ^= 1u;
      if (++sb == p.nbuf) { sb = 0; sphase ^= 1u; }
    }
  } else {
    // ================================ E2: staging -> global + statistics (warps 6-13) ===================
    const int et = tid - kE2Base;
    const int upr = p.block_n >> 3;
    const int RG = kE2Threads / upr;          // rows covered per pass
    const bool active = et < upr * RG;
    const int cu = et % upr, rg = et / upr;
    const bool do_stats = p.stats != nullptr && !(dbg & 8);
    int cur_img = -1, cur_nt = -1;
    // Column statistics: every thread keeps the partial sums of its 8 columns in registers ACROSS tiles (the
    // thread <-> column mapping is tile-invariant); they are combined over the row groups through shared
    // memory in a fixed order, and added to the fp64 accumulators, only when the image or N tile changes.
    float cs[8], cq[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { cs[j] = 0.f; cq[j] = 0.f; }
    auto flush = [&]() {   // uniform across the 256 E2 threads
      // four passes through one 4 KB scratch (sums / squares x lower / upper half of the columns): with streamed weights
      // every KB of shared memory decides the number of pipeline stages, and a flush happens once per image
      const int hb = p.block_n >> 1;              // columns per pass
      float acc4[4] = {0.f, 0.f, 0.f, 0.f};       // { sum lo, sum hi, sq lo, sq hi } of columns et and et + hb
#pragma unroll
      for (int pass = 0; pass < 4; ++pass) {
        const int half = pass & 1;
        const int col = cu * 8 - half * hb;
        if (active && col >= 0 && col < hb) {
          float4* ps = reinterpret_cast<float4*>(s_scr + rg * hb + col);
          if (pass < 2) { ps[0] = make_float4(cs[0], cs[1], cs[2], cs[3]); ps[1] = make_float4(cs[4], cs[5], cs[6], cs[7]); }
          else { ps[0] = make_float4(cq[0], cq[1], cq[2], cq[3]); ps[1] = make_float4(cq[4], cq[5], cq[6], cq[7]); }
        }
        bar_sync(2, kE2Threads);
        if (et < hb) {
          float a = 0.f;
#pragma unroll 4
          for (int g = 0; g < RG; ++g) a += s_scr[g * hb + et];
          acc4[pass] = a;
        }
        bar_sync(2, kE2Threads);
      }
      if (et < hb) {
        double* d = p.stats + ((size_t)cur_img * p.Nc + (size_t)cur_nt * p.block_n + et) * 2;
        atomicAdd(d, (double)acc4[0]);
        atomicAdd(d + 1, (double)acc4[2]);
        atomicAdd(d + 2 * hb, (double)acc4[1]);
        atomicAdd(d + 2 * hb + 1, (double)acc4[3]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) { cs[j] = 0.f; cq[j] = 0.f; }
    };
    TileIter ti; ti.init(t_begin, m_tiles, p.P, tstep);
    int sb = 0; uint32_t sphase = 0;
    for (int it = 0; it < my_tiles; ++it, ti.next(m_tiles, p.P)) {
      const int n_tile = ti.n_tile, m0 = ti.m0, n0 = n_tile * p.block_n;
      if (do_stats && kFast && (ti.img != cur_img || n_tile != cur_nt)) {
        if (cur_img >= 0) flush();
        cur_img = ti.img;
      }
      cur_nt = n_tile;
      mbar_wait(sfull_bar(sb), sphase);
      if (et == 0) TSTAMP(9);
      if (active && !(dbg & 4)) {
        const int rows_valid = min(128, M - m0);
        const uint8_t* src = smem + p.stg_off + (uint32_t)sb * p.stg_bytes + (uint32_t)rg * p.stg_stride + cu * 16;
        bf16* dst = p.out + (long long)(m0 + rg) * p.Nc + n0 + cu * 8;
        const uint32_t src_step = (uint32_t)RG * p.stg_stride;
        const long long dst_step = (long long)RG * p.Nc;
#pragma unroll 4
        for (int r = rg; r < rows_valid; r += RG, src += src_step, dst += dst_step) {
          const uint4 v = *reinterpret_cast<const uint4*>(src);
          if (!(dbg & 2)) *reinterpret_cast<uint4*>(dst) = v;
          if (do_stats) {
            float f[8];
            if (p.out_f16) unpack8h(v, f); else unpack8(v, f);
            if (kFast) {
#pragma unroll
              for (int j = 0; j < 8; ++j) { cs[j] += f[j]; cq[j] = fmaf(f[j], f[j], cq[j]); }
            } else {
              double* d = p.stats + ((size_t)((m0 + r) / p.P) * p.Nc + n0 + cu * 8) * 2;
#pragma unroll
              for (int j = 0; j < 8; ++j) { atomicAdd(d + 2 * j, (double)f[j]); atomicAdd(d + 2 * j + 1, (double)f[j] * f[j]); }
            }
          }
        }
      }
      mbar_arrive(sempty_bar(sb));    // staging tile consumed (release)
      if (et == 0) TSTAMP(10);
      if (++sb == p.nbuf) { sb = 0; sphase ^= 1u; }
    }
    if (do_stats && kFast && cur_img >= 0) flush();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
  // pair mode: the peer may still be multicasting into this CTA's shared memory / arriving on its barriers
  if (p.bpair) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- host side: tensor maps ----------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

struct MapKey {
  const void* ptr; long long M; int K, ld;   // ld carries the element type in bit 30
  bool operator==(const MapKey& o) const { return ptr == o.ptr && M == o.M && K == o.K && ld == o.ld; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    return std::hash<const void*>()(k.ptr) ^ (std::hash<long long>()(k.M) * 1315423911u) ^ ((size_t)k.K << 20) ^ (size_t)k.ld;
  }
};
std::mutex g_map_mu;
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

// [M][K] bf16 activation slice with row stride ld: box = 64 channels x 128 rows, 128-byte swizzle, zero fill
bool activation_map(const void* ptr, long long M, int K, int ld, bool f16, CUtensorMap* out) {
  MapKey key{ptr, M, K, ld | (f16 ? (1 << 30) : 0)};
  std::lock_guard<std::mutex> lk(g_map_mu);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) { *out = it->second; return true; }
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)M};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, 128};
  CUtensorMap m;
  if (!encode_tmap(&m, f16 ? TMAP_F16 : TMAP_BF16, 2, ptr, gdim, gstride, box, true)) return false;
  if (g_maps.size() > 4096) g_maps.clear();
  g_maps[key] = m;
  *out = m;
  return true;
}

// NHWC image tensor [N][H][W][C] bf16: box = 64 channels x box_w pixels x box_h rows (box_w * box_h = 128), zero fill
// box = box_w x box_h pixels of 64 channels; stride 2: every second pixel / row of a (2 box_w) x (2 box_h) window
// (element strides: the box is given in traversal space and ceil(box / stride) elements are loaded)
bool image_map(const void* ptr, int N, int H, int W, int C, int box_w, int box_h, CUtensorMap* out, int stride = 1) {
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t gstride[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)(box_w * stride), (cuuint32_t)(box_h * stride), 1};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  return encode_tmap(out, TMAP_BF16, 4, ptr, gdim, gstride, box, true, estr);
}

}  // namespace

bool encode_tmap(CUtensorMap* out, int dtype, int rank, const void* ptr, const cuuint64_t* gdim,
                 const cuuint64_t* gstride_bytes, const cuuint32_t* box, bool swizzle128, const cuuint32_t* elem_strides) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  if (elem_strides)
    for (int i = 0; i < rank && i < 5; ++i) estr[i] = elem_strides[i];
  return enc(out, dtype == TMAP_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank,
             const_cast<void*>(ptr), gdim, gstride_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int gemm_tc_pick_block_n(int Nc) {
  static int forced = -1;   // experiment switch: LCM_BLOCK_N=<n> caps the N tile
  if (forced < 0) { const char* e = getenv("LCM_BLOCK_N"); forced = e ? atoi(e) : 0; }
  for (int bn = (forced >= 16 && forced <= 256) ? forced / 16 * 16 : 256; bn >= 16; bn -= 16)
    if (Nc % bn == 0) return bn;
  return 0;
}

int gemm_tc_read_timeline(long long* host, int n) {
  return cudaMemcpyFromSymbol(host, g_timeline, sizeof(long long) * (n < 1024 ? n : 1024)) == cudaSuccess ? 0 : -1;
}

int launch_gemm_tc(const GemmParams& g, const ConvGeom& cg, int block_n, int num_sms, cudaStream_t st) {
  Tc2Params p;
  memset(&p, 0, sizeof(p));
  if (block_n < 16 || block_n > 256 || block_n % 16 || g.Nc % block_n) return -1;
  if (g.M <= 0 || g.M > 0x7fffff00LL || g.P <= 0) return -1;
  p.nseg = g.nseg;
  p.W = reinterpret_cast<const bf16*>(g.W);
  p.out = reinterpret_cast<bf16*>(g.out);
  p.stats = g.stats;
  p.M = g.M; p.P = g.P; p.Nc = g.Nc; p.block_n = block_n;
  p.out_f16 = g.out_f16;
  p.n_tiles = g.Nc / block_n;
  p.m_tiles = (g.M + 127) / 128;
  p.fast = (g.P % 128 == 0) ? 1 : 0;
  p.conv_mode = cg.mode;
  int nch = 0, ncoef = 0;
  if (cg.mode < 0) {
    for (int s = 0; s < g.nseg; ++s) {
      p.seg[s] = g.seg[s];
      if (g.seg[s].K % 16 || g.seg[s].ld % 8) return -1;
      p.coef_base[s] = -1;
      if (g.seg[s].mode != XF_NONE) { p.coef_base[s] = ncoef; ncoef += g.seg[s].K; }
      if (!activation_map(g.seg[s].A, g.M, g.seg[s].K, g.seg[s].ld, g.seg[s].f16 != 0, &p.tmap[s])) return -3;
      for (int c0 = 0; c0 < g.seg[s].K; c0 += 64) {
        if (nch >= kMaxChunks2) return -1;
        const int kv = g.seg[s].K - c0 < 64 ? g.seg[s].K - c0 : 64;
        p.chunk[nch++] = (uint32_t)s | (g.seg[s].f16 ? 0x80u : 0u) | ((uint32_t)kv << 8) | ((uint32_t)c0 << 16);
      }
    }
  } else {
    p.bias = cg.bias;
    p.seg[0] = g.seg[0];
    p.Hin = cg.Hin; p.Win = cg.Win; p.Hout = cg.Hout; p.Wout = cg.Wout; p.Ci = cg.Ci;
    if (cg.Ci % 16) return -1;
    // stride-1 convs whose 128-pixel tiles are boxes of the image go through TMA (the zero fill of out-of-bounds
    // coordinates is the conv's zero padding): W >= 128 with 128 | W (tile = part of a row) or W | 128 (whole rows)
    {
      static int no_tma = -1;
      if (no_tma < 0) { const char* e = getenv("LCM_CONV_GATHER"); no_tma = (e && atoi(e)) ? 1 : 0; }
      // tiles are boxes of OUTPUT pixels; the stride-2 convs read every second input pixel / row through a tensor map
      // with element strides (they used to gather their taps with ordinary loads: 1.3 TB/s)
      const int W = cg.Wout, H = cg.Hout;
      const int cstride = cg.mode == CONV_S2 ? 2 : 1;
      const bool boxable = (W >= 128 ? W % 128 == 0 : 128 % W == 0) && ((long long)H * W) % 128 == 0 &&
                           cg.Win == W * cstride && cg.Hin == H * cstride;
      static int no_s2 = -1;
      if (no_s2 < 0) { const char* e = getenv("LCM_CONV_S2_GATHER"); no_s2 = (e && atoi(e)) ? 1 : 0; }
      if ((cg.mode == CONV_S1 || (cg.mode == CONV_S2 && !no_s2)) && boxable && !no_tma) {
        const int bw = W >= 128 ? 128 : W, bh = 128 / bw;
        const long long imgs = g.M / ((long long)H * W);
        if (!image_map(g.seg[0].A, (int)imgs, cg.Hin, cg.Win, cg.Ci, bw, bh, &p.tmap[0], cstride)) return -3;
        p.conv_tma = 1;
        p.conv_stride = cstride;
        p.box_w = bw;
        p.conv_mode = -1;   // runs the TMA-fed (non-gather) kernel variant
        // halo mode: tile = 128 pixels of one image row, weights of all 9 taps resident next to >= 2 halo stages
        static int no_halo = -1;
        if (no_halo < 0) { const char* e = getenv("LCM_CONV_NO_HALO"); no_halo = (e && atoi(e)) ? 1 : 0; }
        const int nchh = (cg.Ci + 63) / 64;
        const uint32_t wres = 9u * nchh * (uint32_t)block_n * 128u;
        const uint32_t stg1 = (128u * ((uint32_t)block_n * 2u + 16u) + 1023u) & ~1023u;
        if (!no_halo && cstride == 1 && bw == 128 && wres <= 131072 && kMisc2 + 2048 + stg1 + wres + 2 * kHaloStage <= kSmemLimit2) {
          if (!image_map(g.seg[0].A, (int)imgs, H, W, cg.Ci, 130, 3, &p.tmap[0])) return -3;
          p.conv_halo = 1;
        }
      }
    }
    for (int tap = 0; tap < 9; ++tap)
      for (int c0 = 0; c0 < cg.Ci; c0 += 64) {
        if (nch >= kMaxChunks2) return -1;
        const int kv = cg.Ci - c0 < 64 ? cg.Ci - c0 : 64;
        p.chunk[nch++] = (uint32_t)tap | ((uint32_t)kv << 8) | ((uint32_t)c0 << 16);
      }
  }
  p.nchunks = nch;
  p.achunks = p.conv_halo ? (cg.Ci + 63) / 64 : nch;
  p.ncoef = ncoef;
  bool has_gate = false;
  if (p.conv_mode < 0 && !p.conv_tma)
    for (int s2 = 0; s2 < g.nseg; ++s2) has_gate |= g.seg[s2].mode == XF_SCALE;
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("LCM_TC_DEBUG"); dbg = e ? atoi(e) : 0; } p.debug = dbg; }
  // shared-memory layout
  const uint32_t b_chunk = (uint32_t)block_n * 128u;
  const uint32_t stg_stride = (uint32_t)block_n * 2u + 16u;
  const uint32_t stg_bytes = (128u * stg_stride + 1023u) & ~1023u;
  uint32_t coef_bytes = ((uint32_t)ncoef * 8u + 1023u) & ~1023u;
  uint32_t base_fixed = coef_bytes + kMisc2 + 1024;
  int ngate = 0;
  for (int ci = 0; ci < nch; ++ci) {
    p.lo_slot[ci] = 0xff;
    if (has_gate && g.seg[p.chunk[ci] & 0x7f].mode == XF_SCALE && !g.seg[p.chunk[ci] & 0x7f].f16) p.lo_slot[ci] = (uint8_t)ngate++;
  }
  static int no_wgate = -1;
  if (no_wgate < 0) { const char* e = getenv("LCM_NO_WGATE"); no_wgate = (e && atoi(e)) ? 1 : 0; }
  uint32_t bres = (uint32_t)nch * b_chunk;
  static uint32_t res_limit = 0;   // largest weight image kept resident (LCM_TC_RES_LIMIT, bytes)
  if (!res_limit) { const char* e = getenv("LCM_TC_RES_LIMIT"); res_limit = e ? (uint32_t)atoi(e) : 131072u; }
  p.resident = (bres <= res_limit && base_fixed + stg_bytes + bres + 3 * kStageA2 <= kSmemLimit2) ? 1 : 0;
  if (has_gate && p.resident && p.fast && !no_wgate) {
    const uint32_t bres2 = (uint32_t)(nch + ngate) * b_chunk;   // + low-order images of the gated chunks
    if (bres2 <= res_limit && base_fixed + stg_bytes + bres2 + 3 * kStageA2 <= kSmemLimit2) { p.wgate = 1; bres = bres2; }
  }
  if (!p.resident && p.fast && p.conv_mode < 0 && !p.conv_tma) {
    // streamed weights: an fp16 SE-gated segment needs only its packed gate pairs (K * 2 bytes) in shared memory
    int nc2 = 0, base2[LCM_MAX_SEGS];
    bool any = false;
    for (int s2 = 0; s2 < g.nseg; ++s2) {
      base2[s2] = -1;
      if (g.seg[s2].mode == XF_NONE) continue;
      base2[s2] = nc2;
      const bool gate16 = g.seg[s2].f16 && g.seg[s2].mode == XF_SCALE && g.seg[s2].K % 8 == 0;
      any |= gate16;
      nc2 += ((gate16 ? g.seg[s2].K / 4 : g.seg[s2].K) + 1) & ~1;   // float2 units, 16-byte aligned bases
    }
    bool all16 = true;   // the flag is per launch: every gated fp16 segment must qualify
    for (int s2 = 0; s2 < g.nseg; ++s2)
      if (g.seg[s2].f16 && g.seg[s2].mode == XF_SCALE && g.seg[s2].K % 8) all16 = false;
    if (any && all16) {
      for (int s2 = 0; s2 < g.nseg; ++s2) p.coef_base[s2] = base2[s2];
      p.coef_compact = 1;
      p.ncoef = ncoef = nc2;
      coef_bytes = ((uint32_t)ncoef * 8u + 1023u) & ~1023u;
      base_fixed = coef_bytes + kMisc2 + 1024;
    }
  }
  p.all_raw = p.conv_tma ? 1 : 0;
  if (p.conv_mode < 0 && !p.conv_tma) {
    bool raw = true;
    for (int s2 = 0; s2 < g.nseg; ++s2)
      raw &= g.seg[s2].mode == XF_NONE || (p.wgate && g.seg[s2].mode == XF_SCALE);
    p.all_raw = raw ? 1 : 0;
  }
  p.stage_bytes = p.conv_halo ? kHaloStage : kStageA2 + (p.resident ? 0u : b_chunk);
  if (p.conv_halo && !p.resident) return -1;
  const uint32_t used1 = base_fixed + stg_bytes + (p.resident ? bres : 0u);
  static int nbuf_min_stages = -1;   // a second epilogue staging tile only if this many pipeline stages still fit
  if (nbuf_min_stages < 0) { const char* e = getenv("LCM_TC_NBUF_STAGES"); nbuf_min_stages = e ? atoi(e) : 6; }   // sweep 4..8 on the model: 6 is best (-0.16 ms per forward against 4)
  p.nbuf = (used1 + stg_bytes + (uint32_t)nbuf_min_stages * p.stage_bytes <= kSmemLimit2) ? 2 : 1;
  const uint32_t used = used1 + (p.nbuf == 2 ? stg_bytes : 0u);
  if (used >= kSmemLimit2) return -1;
  int stages = (int)((kSmemLimit2 - used) / p.stage_bytes);
  static int max_stages = -1;   // LCM_TC_MAX_STAGES (<= 16); sweep 8..16 on the model: no difference, the small-K GEMMs are not bound by bytes in flight
  if (max_stages < 0) { const char* e = getenv("LCM_TC_MAX_STAGES"); max_stages = e ? atoi(e) : 8; if (max_stages > 16) max_stages = 16; if (max_stages < 2) max_stages = 2; }
  if (stages > max_stages) stages = max_stages;
  if (stages < 2) return -1;
  p.stages = stages;
  {
    // LCM_TC_PREFETCH: chunks ahead (0 = off).  Same-box sweep on the model: 4 -> 2.97 ms over the 21 launches, 8 -> 3.00, 0 -> 3.03,
    // 16 -> 3.12; nearly all of it is the K = 768 + 192 project of the 128^2 level (0.398 -> 0.351 ms).  The same prefetch one tile
    // ahead in gemm_wide (whose activation ring already spans a tile) measured 1 % slower and is not in that kernel.
    static int pf = -1;
    if (pf < 0) { const char* e = getenv("LCM_TC_PREFETCH"); pf = e ? atoi(e) : 4; }
    p.prefetch = (p.conv_mode < 0 && !p.conv_tma && !p.conv_halo && p.achunks > stages) ? pf : 0;
  }
  uint32_t off = (uint32_t)stages * p.stage_bytes;
  p.bres_off = off; off += p.resident ? bres : 0u;
  p.stg_off = off; off += (uint32_t)p.nbuf * stg_bytes;
  p.stg_stride = stg_stride;
  p.stg_bytes = stg_bytes;
  p.coef_off = off; off += coef_bytes;
  p.misc_off = off; off += kMisc2;
  const uint32_t smem_bytes = off + 1024;
  if (smem_bytes > kSmemLimit2) return -1;
  const long long tiles = p.m_tiles * p.n_tiles;
  int grid = (int)(tiles < num_sms ? tiles : num_sms);
  // weights streamed (not resident) by a plain GEMM: CTA pairs can share the weight stream (LCM_PAIR=1).  OFF by default:
  // measured on the model it is a wash (32x32-level projects 0.102 -> 0.094 ms, but only 72 pairs = 144 CTAs fit and
  // the K = 768 expand / 64x64 concat project get slower; 912 vs 920 images/s same-box) — those GEMMs are bound by the
  // depth of the TMA -> prologue -> MMA pipeline (2-3 stages of 48 KB), not by L2 bandwidth.  Kept as the tested
  // building block (multicast halves, cluster-wide empty barriers) for a cta_group::2 version.
  static int no_pair = -1;
  if (no_pair < 0) { const char* e = getenv("LCM_PAIR"); no_pair = (e && atoi(e)) ? 0 : 1; }
  p.bpair = (!no_pair && p.conv_mode < 0 && !p.conv_tma && !p.resident && p.m_tiles % 2 == 0 && tiles >= 2 && !(p.debug & 1)) ? 1 : 0;
  typedef void (*KernelFn)(const Tc2Params);
  KernelFn fn;
  if (p.debug) fn = p.conv_mode >= 0 ? (p.fast ? gemm_tc2_kernel<true, true, true> : gemm_tc2_kernel<true, false, true>)
                                     : (p.fast ? gemm_tc2_kernel<false, true, true> : gemm_tc2_kernel<false, false, true>);
  else fn = p.conv_mode >= 0 ? (p.fast ? gemm_tc2_kernel<true, true, false> : gemm_tc2_kernel<true, false, false>)
                             : (p.fast ? gemm_tc2_kernel<false, true, false> : gemm_tc2_kernel<false, false, false>);
  if (ensure_dyn_smem((const void*)fn, kSmemLimit2)) return -2;
  if (p.bpair) {
    int mp;   // per (device, instantiation): 0 = not measured yet, -1 = no pair fits
    {
      int* slot = device_cache_slot((const void*)fn, 1);
      std::lock_guard<std::mutex> lk(g_map_mu);
      if (*slot == 0) {
        int n = 0;
        if (launch_pdl_cluster(fn, dim3(num_sms & ~1), dim3(kThreads2), kSmemLimit2, st, 2, &n, p) != cudaSuccess || n < 1) n = -1;
        *slot = n;
      }
      mp = *slot;
    }
    if (mp < 1) { p.bpair = 0; }
    else {
      long long pairs = tiles / 2;
      if (pairs > mp) pairs = mp;
      if (pairs > num_sms / 2) pairs = num_sms / 2;
      grid = (int)(2 * pairs);
      launch_pdl_cluster(fn, dim3(grid), dim3(kThreads2), smem_bytes, st, 2, nullptr, p);
      return 0;
    }
  }
  launch_pdl(fn, dim3(grid), dim3(kThreads2), smem_bytes, st, p);
  return 0;
}

}  // namespace lcm
