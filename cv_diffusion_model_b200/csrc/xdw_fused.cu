// Fused expand -> GroupNorm2 / FiLM / ReLU6 -> depthwise 3x3 (+ SE pool) of an inverted-residual block on the inference
// tensor-core plan (efficient_unet.py:207-220, :97): the hidden tensor h1 (4x the block width, written once and read once:
// 42 % of a block's HBM traffic) never leaves the SM.
//
//   h1[p][n] = sum_k t[p][k] W6[n][k],  t = relu6(a1 x + b1) / 6  tcgen05, exactly the MMAs of gemm_expand.cu; t is written by the
//                                                                 statistics pass (xstats.cu), so there is no prologue role here
//   t[p][n]  = sat(a2_n/6 fp16(h1) + b2_n/6)                      epilogue: TMEM -> registers -> fp16 -> shared-memory row ring
//   h2[q][n] = sum_tap 6 w[tap][n] t[q + off(tap)][n]             the consumer warps of dwconv_stream.cu, reading the ring
//
// The GroupNorm2 coefficients (a2, b2) need the statistics of ALL of h1 first; they come from the input side — the
// pass of xstats.cu over the block input (S, G of the transformed input t) — so this kernel can start with them in hand.
// Values are bit-identical to the unfused pair (same MMAs, same fp16 rounding of h1, same HFMA2 sequences); only the
// reduction order of the SE pool's fp32 partial sums differs.
//
// Work item = (128-channel block nb, image, band of 64 pixel columns, segment of `hseg` <= 62 rows).  Tiles of an item:
//   * one HALO tile: the two pixel columns left and right of the band (x0 - 1 and x0 + 64) for rows y0-1 .. y0+62 as the
//     128 rows of one A tile (two TMA boxes [64 ch][1 px][64 rows]) -> a halo buffer [row][side][128 ch];
//   * hseg/2 + 1 MAIN tiles: two band rows (y0-1+2j, y0+2j) x 64 pixels = 128 A rows (one TMA box [64 ch][64 px][2 rows]);
//     every in-image row becomes one slot of the row ring (XOR-swizzled 16-byte units: conflict-free for the row-per-thread
//     writer and the 4-channels-per-lane reader).
// Roles (448 threads, one persistent CTA per SM):
//   warps 0-7   CONV  strip of 8 pixels x 128 channels each, 3-row register window (dwconv_stream.cu), h2 stores, pool
//   warps 8-11  EPI   TMEM lane quadrants: accumulator -> fp16 -> GN2/FiLM/ReLU6 -> ring slot / halo buffer
//                     (tcgen05.ld of the next 32 columns in flight while the current 32 are converted and stored)
//   warp  12    MMA   M = 128, N = 128 per tile, two accumulators in TMEM
//   warp  13    TMA   weights of the n-block; t tiles through a ring of stages
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

// A conv warp is a strip of kPxS = 8 pixels x 128 channels (~110 registers: three window rows of 10 pixels x 4 channels, the taps).
// The conv role is latency-bound (ncu: its warps are busy ~90 % of the time at ~0.12 instructions per cycle each).  Measured
// alternatives: kPxS = 4 with 16 conv warps (22 warps -> 80 registers each, spills in the row loop and in the epilogue role):
// 0.63 vs 0.45 ms per level-0 launch; prefetching the next row into a fourth register row at 128 registers: 0.74 ms.
// The register file is per SM sub-partition (4 warps x 144 registers do not fit), and ptxas does not raise a role's budget
// on setmaxnreg.inc above the launch bound's, so 8 conv warps at 128 registers it is.
constexpr int kPxS = 8;
static_assert(kPxS == 8, "only the 8-pixel strip is validated");
constexpr int kConvWarps = 64 / kPxS, kEpiWarp0 = kConvWarps, kMmaWarpF = kConvWarps + 4, kTmaWarpF = kConvWarps + 5;
constexpr int kThreadsF = (kConvWarps + 6) * 32;   // 14 warps
constexpr uint32_t kChunkF = 16384;       // 128 rows x 64 16-bit elements
constexpr uint32_t kSlotF = 16384;        // ring slot: 64 px x 128 ch fp16
constexpr uint32_t kHaloRowF = 512;       // halo buffer: per row [2 sides][128 ch] fp16
constexpr int kMaxStagesF = 4, kMaxRingF = 6;
constexpr uint32_t kSmemLimitF = 232448;

struct FParams {
  CUtensorMap tmap_main, tmap_halo;   // t [N][H][W][Kt] bf16
  int nchunks;
  uint32_t chunk[2];          // kvalid << 8 (64-wide K chunks of t)
  const bf16* Wp;             // packed [NB][chunk][128 x 64] (x6), as for gemm_expand
  const float2* coef2;        // [N][Ch] (a2, b2)
  const float* wdw;           // [9][Ch]
  __half* out;                // h2 [N][H][W][Ch]
  double* pool;               // [N][Ch]
  int N, H, W, Ch, NB;
  int hseg, bandsX, segsY, items, stages, ring;
  uint32_t w_off, ring_off, halo_off, halo_bytes, misc_off;
  int spin;
  int dbg;   // LCM_XDW_DBG bottleneck experiments: 1 conv warps skip the arithmetic, 2 EPI skips conversion + stores
};

__device__ __forceinline__ void tma_load_4d_f(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}
__device__ __forceinline__ uint2 lds64f(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
  return v;
}
template <int OFF>
__device__ __forceinline__ uint2 lds64f_imm(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2+%3];" : "=r"(v.x), "=r"(v.y) : "r"(addr), "n"(OFF) : "memory");
  return v;
}
template <int I>
__device__ __forceinline__ uint2 lds_px(uint32_t rw, const uint32_t (&xo)[8]) {   // window pixel I of a strip: band pixel strip * 8 - 1 + I
  return lds64f_imm<(I - 1) * 256>(rw + xo[(I + 7) & 7]);
}
__device__ __forceinline__ __half2 as_h2f(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t as_u32f(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

// every wait of this kernel suspends with a time hint instead of re-polling: the polling loops of the waiting roles were 12 % of
// all executed instructions (ncu), issued on the sub-partitions the conv warps need (LCM_XDW_SPIN=1: plain try_wait loops)
__device__ __forceinline__ void waitf_impl(uint32_t bar, uint32_t parity, bool spin) {
  if (spin) mbar_wait(bar, parity); else mbar_wait_relaxed(bar, parity);
}
struct ItemF { int nb, n, bx, sy; };
__device__ __forceinline__ ItemF decode_f(int item, const FParams& p) {
  ItemF q;
  q.sy = item % p.segsY; item /= p.segsY;
  q.bx = item % p.bandsX; item /= p.bandsX;
  q.n = item % p.N;
  q.nb = item / p.N;          // slowest: the weights of an n-block are reloaded a couple of times per CTA at most
  return q;
}

typedef __half2 RowF[kPxS + 2][2];

// kPartial: the hidden width is not a multiple of 128 — lanes whose four channels lie beyond it (cvalid false) compute on zeros
// and skip their stores
template <bool kRagged, bool kPartial = false>
__device__ __forceinline__ void emit_row_f(const RowF& r0, const RowF& r1, const RowF& r2, const __half2 (&w6)[9][2], __half* orow, int C,
                                           int nvalid, float (&psum)[4], bool cvalid = true) {
  __half2 s0 = __float2half2_rn(0.f), s1 = s0;
#pragma unroll
  for (int px = 0; px < kPxS; ++px) {
    __half2 a0 = __hmul2(r0[px][0], w6[0][0]), a1 = __hmul2(r0[px][1], w6[0][1]);
    a0 = __hfma2(r0[px + 1][0], w6[1][0], a0); a1 = __hfma2(r0[px + 1][1], w6[1][1], a1);
    a0 = __hfma2(r0[px + 2][0], w6[2][0], a0); a1 = __hfma2(r0[px + 2][1], w6[2][1], a1);
    a0 = __hfma2(r1[px][0], w6[3][0], a0); a1 = __hfma2(r1[px][1], w6[3][1], a1);
    a0 = __hfma2(r1[px + 1][0], w6[4][0], a0); a1 = __hfma2(r1[px + 1][1], w6[4][1], a1);
    a0 = __hfma2(r1[px + 2][0], w6[5][0], a0); a1 = __hfma2(r1[px + 2][1], w6[5][1], a1);
    a0 = __hfma2(r2[px][0], w6[6][0], a0); a1 = __hfma2(r2[px][1], w6[6][1], a1);
    a0 = __hfma2(r2[px + 1][0], w6[7][0], a0); a1 = __hfma2(r2[px + 1][1], w6[7][1], a1);
    a0 = __hfma2(r2[px + 2][0], w6[8][0], a0); a1 = __hfma2(r2[px + 2][1], w6[8][1], a1);
    if (!kRagged || px < nvalid) {
      if (!kPartial || cvalid) *reinterpret_cast<uint2*>(orow + (size_t)px * C) = make_uint2(as_u32f(a0), as_u32f(a1));
      s0 = __hadd2(s0, a0); s1 = __hadd2(s1, a1);
    }
  }
  const float2 f0 = __half22float2(s0), f1 = __half22float2(s1);
  psum[0] += f0.x; psum[1] += f0.y; psum[2] += f1.x; psum[3] += f1.y;
}

// kDbg: the LCM_XDW_DBG / LCM_XDW_SPIN experiment switches are compiled into a second instantiation only (the row loop of the conv
// role pays for every instruction)
// kCh: the hidden width as a compile-time constant (0 = p.Ch): the eight h2 stores of a row get immediate offsets instead of an
// IMAD.WIDE each — on the pipe the HFMA2s need
template <bool kDbg, int kCh>
__global__ void __launch_bounds__(kThreadsF, 1) xdw_fused_kernel(const __grid_constant__ FParams p) {
  extern __shared__ uint8_t fsm_raw2[];
  const uint32_t sraw = smem_u32(fsm_raw2);
  const uint32_t sbase = (sraw + 1023u) & ~1023u;
  uint8_t* smem = fsm_raw2 + (sbase - sraw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool spin = kDbg && p.spin != 0;
  const int dbg = kDbg ? p.dbg : 0;
  // hidden widths that are not a multiple of the 128-channel n-block (Base variant: 192): the last block is half empty — its weight
  // rows and coefficients are zero, its stores and pool sums masked.  Compile-time false for the widths of the Small / Large variants.
  constexpr bool kPartial = kCh == 0 || (kCh % 128) != 0;
  auto waitf = [&](uint32_t bar, uint32_t parity) { if (kDbg) waitf_impl(bar, parity, spin); else mbar_wait_relaxed(bar, parity); };

  const uint32_t bar0 = sbase + p.misc_off;
  auto raw_bar = [&](int s) { return bar0 + 8u * s; };                 // TMA -> MMA          (stages)
  auto empty_bar = [&](int s) { return bar0 + 8u * (8 + s); };         // MMA -> TMA
  auto tfull_bar = [&](int a) { return bar0 + 8u * (12 + a); };        // MMA -> EPI          (accumulators)
  auto tempty_bar = [&](int a) { return bar0 + 8u * (14 + a); };       // EPI -> MMA
  auto sfull_bar = [&](int s) { return bar0 + 8u * (16 + s); };        // EPI -> CONV         (ring slots)
  auto sempty_bar = [&](int s) { return bar0 + 8u * (24 + s); };       // CONV -> EPI
  auto hfull_bar = [&](int b) { return bar0 + 8u * (32 + b); };        // EPI -> CONV         (halo buffers)
  auto hempty_bar = [&](int b) { return bar0 + 8u * (34 + b); };       // CONV -> EPI
  const uint32_t wres_bar = bar0 + 8u * 36;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + p.misc_off + 320);
  uint32_t* s_ab2 = reinterpret_cast<uint32_t*>(smem + p.misc_off + 1536);        // [2][64] half2 pairs: a2/6 then b2/6 (128 ch)
  float* s_red = reinterpret_cast<float*>(smem + p.misc_off + 2048);             // [2][8][128]

  if (warp == kTmaWarpF && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(raw_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 128); }
    for (int s = 0; s < p.ring; ++s) { mbar_init(sfull_bar(s), 64); mbar_init(sempty_bar(s), kConvWarps); }
    for (int b = 0; b < 2; ++b) { mbar_init(hfull_bar(b), 128); mbar_init(hempty_bar(b), kConvWarps); }
    mbar_init(wres_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmap_main)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmap_halo)) : "memory");
  }
  if (warp == kMmaWarpF) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  const int ntiles = p.hseg / 2 + 2;     // per item: 1 halo tile + hseg/2 + 1 main tiles (hseg is even)

  if (warp < kConvWarps) {
    // ================================ CONV: strip of kPxS pixels, lane = 4 channels ===================================
    const int strip = warp;
    int slot = 0; uint32_t sphase = 0;
    const __half2 hz = __float2half2_rn(0.f);
    const uint32_t ux = (uint32_t)(lane >> 1) << 4, hoff = (uint32_t)(lane & 1) * 8u;
    // unit (lane >> 1) of a pixel sits at unit ^ (pixel & 7): the eight XOR patterns of this lane, kept in registers so that a row's
    // ten loads are one add + an immediate pixel offset each
    uint32_t xo[8];
#pragma unroll
    for (int sx = 0; sx < 8; ++sx) xo[sx] = (ux ^ ((uint32_t)sx << 4)) + hoff + (uint32_t)(strip * kPxS) * 256u;
    int k = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++k) {
      const ItemF q = decode_f(item, p);
      const int c = q.nb * 128 + lane * 4;
      const int x0 = q.bx * 64, xs = x0 + strip * kPxS;
      const int y0 = q.sy * p.hseg, y1 = min(p.H, y0 + p.hseg);
      const int hb = k & 1;
      const uint32_t halo = sbase + p.halo_off + (uint32_t)hb * p.halo_bytes + (uint32_t)lane * 8u;

      const bool cvalid = !kPartial || c < p.Ch;
      __half2 w6[9][2];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float4 wv = cvalid ? *reinterpret_cast<const float4*>(p.wdw + (size_t)t * p.Ch + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        w6[t][0] = __floats2half2_rn(6.f * wv.x, 6.f * wv.y);
        w6[t][1] = __floats2half2_rn(6.f * wv.z, 6.f * wv.w);
      }
      // (no padding mask here: band pixels are always inside the image (W % 64 == 0) and the EPI warps write ZEROS into the halo
      //  buffer for a neighbour column outside it, so the zero padding after the activation costs this role nothing)

      waitf(hfull_bar(hb), (uint32_t)(k >> 1) & 1u);     // the item's halo columns are in place

      // (prefetching row y + 2 into a fourth register row while row y is computed was tried: at the 128 registers that four
      //  warps per sub-partition allow it spills inside the row loop and is 60 % slower)
      auto load_row = [&](RowF& r, int y) {
        if (y < 0 || y >= p.H) {
#pragma unroll
          for (int i = 0; i < kPxS + 2; ++i) r[i][0] = r[i][1] = hz;
          return;
        }
        waitf(sfull_bar(slot), sphase);
        const uint32_t rw = sbase + p.ring_off + (uint32_t)slot * kSlotF;
        const uint32_t hrow = halo + (uint32_t)(y - (y0 - 1)) * kHaloRowF;
        uint2 v[kPxS + 2];
        v[0] = strip == 0 ? lds64f(hrow) : lds_px<0>(rw, xo);                                 // left neighbour column from the halo buffer
        v[1] = lds_px<1>(rw, xo); v[2] = lds_px<2>(rw, xo); v[3] = lds_px<3>(rw, xo); v[4] = lds_px<4>(rw, xo);
        v[5] = lds_px<5>(rw, xo); v[6] = lds_px<6>(rw, xo); v[7] = lds_px<7>(rw, xo); v[8] = lds_px<8>(rw, xo);
        v[9] = strip == kConvWarps - 1 ? lds64f(hrow + 256u) : lds_px<9>(rw, xo);             // right neighbour column
#pragma unroll
        for (int i = 0; i < kPxS + 2; ++i) { r[i][0] = as_h2f(v[i].x); r[i][1] = as_h2f(v[i].y); }
        __syncwarp();
        if (lane == 0) mbar_arrive(sempty_bar(slot));
        if (++slot == p.ring) { slot = 0; sphase ^= 1u; }
      };

      float psum[4] = {0.f, 0.f, 0.f, 0.f};
      RowF ra, rb, rc;
      load_row(ra, y0 - 1);
      load_row(rb, y0);
      __half* orow = p.out + (((size_t)q.n * p.H + y0) * p.W + xs) * p.Ch + c;
      const size_t ostep = (size_t)p.W * p.Ch;
      auto emit = [&](const RowF& r0, const RowF& r1, const RowF& r2) {   // (W % 64 == 0: a strip is never ragged)
        if (!(dbg & 1)) emit_row_f<false, kPartial>(r0, r1, r2, w6, orow, kCh ? kCh : p.Ch, kPxS, psum, cvalid);
        orow += ostep;
      };
      for (int y = y0; y < y1; y += 3) {
        load_row(rc, y + 1);
        emit(ra, rb, rc);
        if (y + 1 >= y1) break;
        load_row(ra, y + 2);
        emit(rb, rc, ra);
        if (y + 2 >= y1) break;
        load_row(rb, y + 3);
        emit(rc, ra, rb);
      }
      // halo buffer of this item is free again
      __syncwarp();
      if (lane == 0) mbar_arrive(hempty_bar(hb));

      // ---- pooled sums: fixed-order reduction over the strips, one fp64 atomic per channel ---------------
      float* red = s_red + hb * (kConvWarps * 128);
      *reinterpret_cast<float4*>(red + strip * 128 + lane * 4) = make_float4(psum[0], psum[1], psum[2], psum[3]);
      bar_sync(1, kConvWarps * 32);
      if (tid < 128 && (!kPartial || q.nb * 128 + tid < p.Ch)) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < kConvWarps; ++i) s += red[i * 128 + tid];
        atomicAdd(&p.pool[(size_t)q.n * p.Ch + q.nb * 128 + tid], (double)s);
      }
    }
  } else if (warp < kMmaWarpF) {
    // ================================ EPI: accumulator -> fp16 -> GN2 / FiLM / ReLU6 -> ring / halo =================
    const int ew = warp - kEpiWarp0;                   // TMEM lane quadrant (warp % 4 == ew since kEpiWarp0 % 4 == 0)
    const int et = ew * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(ew * 32) << 16);
    long long nslot = 0;                               // ring slots produced so far
    int tcount = 0;                                    // tiles drained so far (accumulator = tcount & 1)
    int k = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++k) {
      const ItemF q = decode_f(item, p);
      const int y0 = q.sy * p.hseg;
      const int hb = k & 1;
      // a2/6, b2/6 of this (image, n-block) as half2 pairs
      bar_sync(2, 128);      // previous item's reads of s_ab2 are done
      if (et < 64) {
        const float4 cf = (!kPartial || q.nb * 128 + et * 2 < p.Ch)
                              ? *reinterpret_cast<const float4*>(p.coef2 + (size_t)q.n * p.Ch + q.nb * 128 + et * 2)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
        const __half2 a = __floats2half2_rn(cf.x * (1.f / 6.f), cf.z * (1.f / 6.f));
        const __half2 b = __floats2half2_rn(cf.y * (1.f / 6.f), cf.w * (1.f / 6.f));
        s_ab2[et] = as_u32f(a);
        s_ab2[64 + et] = as_u32f(b);
      }
      bar_sync(2, 128);
      for (int tl = 0; tl < ntiles; ++tl, ++tcount) {
        const int acc = tcount & 1;
        // destination of this thread's pixel (128 channels = 256 bytes)
        uint32_t dst; bool swz; int px = 0; bool valid; int sl = -1; bool hzero = false;
        if (tl == 0) {
          waitf(hempty_bar(hb), ((uint32_t)(k >> 1) & 1u) ^ 1u);
          const int side = et >> 6, hr = et & 63;
          dst = sbase + p.halo_off + (uint32_t)hb * p.halo_bytes + (uint32_t)hr * kHaloRowF + (uint32_t)side * 256u;
          swz = false;
          valid = hr < p.hseg + 2;
          hzero = side == 0 ? q.bx == 0 : q.bx == p.bandsX - 1;   // that neighbour column is outside the image: zero padding
        } else {
          const int j = tl - 1;
          const int ra = y0 - 1 + 2 * j, rbw = ra + 1;
          const bool va = ra >= 0 && ra < p.H, vb = rbw >= 0 && rbw < p.H;
          const int half = ew >> 1;                                   // 0: row ra (lanes 0-63), 1: row rb
          valid = half ? vb : va;
          const long long idx = nslot + (half && va ? 1 : 0);
          nslot += (va ? 1 : 0) + (vb ? 1 : 0);
          px = (ew & 1) * 32 + lane;
          dst = 0; swz = true;
          if (valid) {
            sl = (int)(idx % p.ring);
            const uint32_t ph = (uint32_t)((idx / p.ring) & 1);
            waitf(sempty_bar(sl), ph ^ 1u);
            dst = sbase + p.ring_off + (uint32_t)sl * kSlotF + (uint32_t)px * 256u;
          }
        }
        waitf(tfull_bar(acc), (uint32_t)(tcount >> 1) & 1u);
        tc_fence_after();
        const uint32_t taddr = lane_base + (uint32_t)acc * 128u;
        // 32 columns per step; the load of step i + 1 is in flight while step i is converted and stored
        auto emit32 = [&](const uint32_t (&r0)[16], const uint32_t (&r1)[16], int cb) {
          if (valid && !(dbg & 2)) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const uint32_t* r = u < 2 ? r0 : r1;
              const int o = (u & 1) * 8;
              const int unit = (cb >> 3) + u;                         // 8-channel unit 0..15
              const uint4 a4 = *reinterpret_cast<const uint4*>(s_ab2 + unit * 4);
              const uint4 b4 = *reinterpret_cast<const uint4*>(s_ab2 + 64 + unit * 4);
              uint4 v;
              v.x = as_u32f(__hfma2_sat(as_h2f(a4.x), as_h2f(pack_f16(__uint_as_float(r[o + 0]), __uint_as_float(r[o + 1]))), as_h2f(b4.x)));
              v.y = as_u32f(__hfma2_sat(as_h2f(a4.y), as_h2f(pack_f16(__uint_as_float(r[o + 2]), __uint_as_float(r[o + 3]))), as_h2f(b4.y)));
              v.z = as_u32f(__hfma2_sat(as_h2f(a4.z), as_h2f(pack_f16(__uint_as_float(r[o + 4]), __uint_as_float(r[o + 5]))), as_h2f(b4.z)));
              v.w = as_u32f(__hfma2_sat(as_h2f(a4.w), as_h2f(pack_f16(__uint_as_float(r[o + 6]), __uint_as_float(r[o + 7]))), as_h2f(b4.w)));
              if (hzero) v = make_uint4(0u, 0u, 0u, 0u);
              const uint32_t uu = swz ? (uint32_t)(unit ^ (px & 7)) : (uint32_t)unit;
              sts128(dst + (uu << 4), v);
            }
          }
        };
        {
          uint32_t ra0[16], ra1[16], rb0[16], rb1[16];
          tmem_ld16(taddr, ra0); tmem_ld16(taddr + 16, ra1);
          tmem_wait_ld();
          tmem_ld16(taddr + 32, rb0); tmem_ld16(taddr + 48, rb1);
          emit32(ra0, ra1, 0);
          tmem_wait_ld();
          tmem_ld16(taddr + 64, ra0); tmem_ld16(taddr + 80, ra1);
          emit32(rb0, rb1, 32);
          tmem_wait_ld();
          tmem_ld16(taddr + 96, rb0); tmem_ld16(taddr + 112, rb1);
          emit32(ra0, ra1, 64);
          tmem_wait_ld();
          emit32(rb0, rb1, 96);
        }
        tc_fence_before();
        mbar_arrive(tempty_bar(acc));
        if (tl == 0) mbar_arrive(hfull_bar(hb));
        else if (valid) mbar_arrive(sfull_bar(sl));
      }
    }
  } else {
  if (warp == kTmaWarpF) {
    // ================================ TMA: weights of the n-block, x tiles ==========================================
    int stage = 0; uint32_t phase = 0;
    int cur_nb = -1;
    int last_stage = -1; uint32_t last_phase = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const ItemF q = decode_f(item, p);
      const int x0 = q.bx * 64, y0 = q.sy * p.hseg;
      if (q.nb != cur_nb) {
        // every MMA that reads the old weights has completed when the last stage issued has been released
        if (last_stage >= 0) waitf(empty_bar(last_stage), last_phase);
        if (elect_one()) {
          const uint32_t wbytes = (uint32_t)p.nchunks * kChunkF;
          mbar_expect_tx(wres_bar, wbytes);
          for (int ci = 0; ci < p.nchunks; ++ci)
            bulk_g2s(sbase + p.w_off + (uint32_t)ci * kChunkF,
                     reinterpret_cast<const uint8_t*>(p.Wp) + ((size_t)q.nb * p.nchunks + ci) * kChunkF, kChunkF, wres_bar);
        }
        __syncwarp();
        cur_nb = q.nb;
      }
      for (int tl = 0; tl < ntiles; ++tl) {
        for (int ci = 0; ci < p.nchunks; ++ci) {
          const int c0 = ci * 64;
          mbar_wait_relaxed(empty_bar(stage), phase ^ 1u);
          if (elect_one()) {
            const uint32_t dst = sbase + (uint32_t)stage * kChunkF;
            mbar_expect_tx(raw_bar(stage), kChunkF);
            if (tl == 0) {
              tma_load_4d_f(dst, &p.tmap_halo, c0, x0 - 1, y0 - 1, q.n, raw_bar(stage));
              tma_load_4d_f(dst + 8192u, &p.tmap_halo, c0, x0 + 64, y0 - 1, q.n, raw_bar(stage));
            } else {
              tma_load_4d_f(dst, &p.tmap_main, c0, x0, y0 - 1 + 2 * (tl - 1), q.n, raw_bar(stage));
            }
          }
          __syncwarp();
          last_stage = stage; last_phase = phase;
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == kMmaWarpF) {
    // ================================ MMA issuer ======================================================================
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | (8u << 24);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const int ks0 = (int)((p.chunk[0] >> 8) & 0xff) >> 4;
    const int ks1 = p.nchunks > 1 ? (int)((p.chunk[1] >> 8) & 0xff) >> 4 : 0;
    const uint64_t wd0 = umma_desc(sbase + p.w_off);
    int stage = 0; uint32_t phase = 0;
    int tcount = 0, cur_nb = -1;
    uint32_t wphase = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const ItemF q = decode_f(item, p);
      if (q.nb != cur_nb) { waitf(wres_bar, wphase); wphase ^= 1u; cur_nb = q.nb; }
      for (int tl = 0; tl < ntiles; ++tl, ++tcount) {
        const int acc = tcount & 1;
        {
          int s2 = stage; uint32_t ph2 = phase;
          for (int ci = 0; ci < p.nchunks; ++ci) {
            waitf(raw_bar(s2), ph2);
            if (++s2 == p.stages) { s2 = 0; ph2 ^= 1u; }
          }
        }
        waitf(tempty_bar(acc), ((uint32_t)(tcount >> 1) & 1u) ^ 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ad0 = umma_desc(sbase + (uint32_t)stage * kChunkF);
          const uint32_t d = tmem_u + (uint32_t)acc * 128u;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            if (kk < ks0) umma_bf16(d, ad0 + (uint64_t)(2 * kk), wd0 + (uint64_t)(2 * kk), idesc, kk != 0 ? 1u : 0u);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            if (kk < ks1) umma_bf16(d, ad0 + (uint64_t)((kChunkF >> 4) + 2 * kk), wd0 + (uint64_t)((kChunkF >> 4) + 2 * kk), idesc, 1u);
          umma_commit(tfull_bar(acc));
          for (int ci = 0; ci < p.nchunks; ++ci) umma_commit(empty_bar((stage + ci) % p.stages));
        }
        __syncwarp();
        stage += p.nchunks;
        if (stage >= p.stages) { stage -= p.stages; phase ^= 1u; }
      }
    }
  }

  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarpF) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
  }
}

std::mutex g_f_mu;
struct FKey {
  const void* ptr; int N, H, W, C, kind;
  bool operator==(const FKey& o) const { return ptr == o.ptr && N == o.N && H == o.H && W == o.W && C == o.C && kind == o.kind; }
};
struct FKeyHash {
  size_t operator()(const FKey& k) const {
    return std::hash<const void*>()(k.ptr) ^ ((size_t)k.N * 1315423911u) ^ ((size_t)k.H << 40) ^ ((size_t)k.W << 24) ^ ((size_t)k.C << 8) ^ (size_t)k.kind;
  }
};
std::unordered_map<FKey, CUtensorMap, FKeyHash> g_f_maps;

// x [N][H][W][C] bf16, 128-byte swizzle; kind 0: box {64 ch, 64 px, 2 rows}, kind 1: box {64 ch, 1 px, 64 rows}
bool x_map(const void* ptr, int N, int H, int W, int C, int kind, CUtensorMap* out) {
  FKey key{ptr, N, H, W, C, kind};
  std::lock_guard<std::mutex> lk(g_f_mu);
  auto it = g_f_maps.find(key);
  if (it != g_f_maps.end()) { *out = it->second; return true; }
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t gstride[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, kind == 0 ? 64u : 1u, kind == 0 ? 2u : 64u, 1};
  CUtensorMap m;
  if (!encode_tmap(&m, TMAP_BF16, 4, ptr, gdim, gstride, box, true)) return false;
  if (g_f_maps.size() > 4096) g_f_maps.clear();
  g_f_maps[key] = m;
  *out = m;
  return true;
}

int pick_hseg(int H) {
  // even, <= 62 (the halo tile holds hseg + 2 rows), dividing H when possible; longer segments = less halo overhead
  for (int hs = 62; hs >= 2; hs -= 2)
    if (H % hs == 0 && hs <= 32) return hs;
  return H >= 32 ? 32 : (H & ~1);
}

}  // namespace

bool xdw_fused_supported(int nseg, const int* segK, int Nc, int H, int W) {
  // LCM_NO_XDW=1 falls back to the unfused kernel pair (A/B timing; tests/diag_xdw.py)
  static int off = -1;
  if (off < 0) { const char* e = getenv("LCM_NO_XDW"); off = (e && atoi(e)) ? 1 : 0; }
  if (off || nseg < 1 || nseg > 2 || Nc % 64 || Nc < 128 || W % 64 || H % 2 || H < 2) return false;
  int Kt = 0;
  for (int s = 0; s < nseg; ++s) {
    if (segK[s] % 16 || segK[s] < 16) return false;
    Kt += segK[s];
  }
  return Kt <= 128 && xstats_supported(Kt, H * W);
}

// t: relu6(GN1(x)) / 6 of the block input, bf16 [N][H][W][Kt] (xstats.cu); Wp: the expand weights packed as ONE K segment
// for gemm_expand (block_n 128, x6); coef2: GroupNorm2 + FiLM (a, b) per (image, channel); wdw [9][Nc];
// out: h2 fp16 [N][H][W][Nc]; pool [N][Nc] fp64 (+=).
int launch_xdw_fused(const void* t, int Kt, const void* Wp, int Nc, const float2* coef2, const float* wdw, void* out, double* pool,
                     int N, int H, int Wd, int num_sms, cudaStream_t st) {
  if (Kt % 16 || Kt < 16 || Kt > 128 || Nc % 64 || Nc < 128 || Wd % 64 || H % 2 || H < 2) return -1;
  FParams p;
  memset(&p, 0, sizeof(p));
  if (!x_map(t, N, H, Wd, Kt, 0, &p.tmap_main) || !x_map(t, N, H, Wd, Kt, 1, &p.tmap_halo)) return -3;
  int nch = 0;
  for (int c0 = 0; c0 < Kt; c0 += 64) p.chunk[nch++] = (uint32_t)(Kt - c0 < 64 ? Kt - c0 : 64) << 8;
  p.nchunks = nch;
  p.Wp = reinterpret_cast<const bf16*>(Wp);
  p.coef2 = coef2; p.wdw = wdw; p.out = reinterpret_cast<__half*>(out); p.pool = pool;
  p.N = N; p.H = H; p.W = Wd; p.Ch = Nc; p.NB = (Nc + 127) / 128;
  p.hseg = pick_hseg(H);
  p.bandsX = Wd / 64;
  p.segsY = (H + p.hseg - 1) / p.hseg;
  p.items = p.NB * N * p.bandsX * p.segsY;
  p.stages = 4;                                    // 4 x 16 KB (two tiles in flight when a tile has two chunks)
  p.ring = nch == 1 ? 5 : 4;
  { const char* e = getenv("LCM_XDW_STAGES"); if (e && atoi(e) >= nch && atoi(e) <= 4) p.stages = atoi(e) - atoi(e) % nch; }
  { const char* e = getenv("LCM_XDW_RING"); if (e && atoi(e) >= 3 && atoi(e) <= 8) p.ring = atoi(e); }
  { const char* e = getenv("LCM_XDW_HSEG"); if (e && atoi(e) >= 2 && atoi(e) <= 62 && atoi(e) % 2 == 0) p.hseg = atoi(e); }
  p.segsY = (H + p.hseg - 1) / p.hseg;
  p.items = p.NB * N * p.bandsX * p.segsY;
  p.halo_bytes = (uint32_t)(p.hseg + 2) * kHaloRowF;
  uint32_t off = (uint32_t)p.stages * kChunkF;
  p.w_off = off; off += (uint32_t)nch * kChunkF;
  p.ring_off = off; off += (uint32_t)p.ring * kSlotF;
  p.halo_off = off; off += 2 * p.halo_bytes;
  off = (off + 1023u) & ~1023u;
  p.misc_off = off; off += 2048 + 2 * kConvWarps * 128 * 4;
  const uint32_t total = off + 1024;
  if (total > kSmemLimitF) return -1;
  { static int d = -1; if (d < 0) { const char* e = getenv("LCM_XDW_DBG"); d = e ? atoi(e) : 0; } p.dbg = d; }
  { static int sp = -1; if (sp < 0) { const char* e = getenv("LCM_XDW_SPIN"); sp = (e && atoi(e)) ? 1 : 0; } p.spin = sp; }
  const int grid = p.items < num_sms ? p.items : num_sms;
  auto go = [&](auto kfn) -> int {
    if (ensure_dyn_smem_fn(kfn, kSmemLimitF)) return -2;
    launch_pdl(kfn, dim3(grid), dim3(kThreadsF), total, st, p);
    return 0;
  };
  if (p.dbg || p.spin) return go(xdw_fused_kernel<true, 0>);
  switch (Nc) {
    case 128: return go(xdw_fused_kernel<false, 128>);
    case 192: return go(xdw_fused_kernel<false, 192>);
    case 256: return go(xdw_fused_kernel<false, 256>);
    case 384: return go(xdw_fused_kernel<false, 384>);
    default: return go(xdw_fused_kernel<false, 0>);
  }
}

}  // namespace lcm
