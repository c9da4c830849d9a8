// Depthwise 3x3 (efficient_unet.py:177-180,220) of the tcgen05 product path — the largest HBM consumer of the model.
//
//   out = dw3x3( relu6(a*x + b) ),  pooled[n][c] += sum_{y,x} out       (a,b: GroupNorm2 + FiLM per (image, channel))
//
// Both hidden tensors of an inverted-residual block are stored as fp16 on this path (h1 is written by the expand
// GEMM's epilogue, h2 is read by the project GEMM as an f16 operand): values are O(1) after the expand, fp16 keeps
// 11 significant bits against bf16's 8, and — the point — the whole prologue becomes ONE packed instruction per
// two elements:   relu6(a*x + b) = 6 * sat(a/6 * x + b/6)   ->  fma.rn.sat.f16x2,  the 6 folded into the taps.
// That takes the kernel from ~22 thread instructions per output element (issue-bound at 2.7 TB/s) to ~7.
//
// Structure (memory-bound; nothing waits on global latency with registers):
//   * persistent CTAs; a work item is (image, 128-channel block, band of `pxw` pixel columns, segment of rows).
//   * thread 0 streams the band row by row with 4-D TMA boxes [128 ch][pxw+2 px][1 row] into a ring of
//     shared-memory row slots (full/empty mbarriers), running ahead of the consumers across item boundaries;
//     out-of-image columns arrive zero-filled.  (No dedicated producer warp: 8 warps per CTA keep the four
//     register-file partitions evenly loaded, which is what lets two CTAs x 128 registers fit an SM.)
//   * consumer warps: one warp = one strip of 8 pixels x 128 channels (lane = 4 channels, LDS.64 / STG.64: a warp
//     reads and writes 256 contiguous bytes per pixel).  Each thread keeps a 3-row window of TRANSFORMED inputs in
//     registers (10 pixels x 4 channels per row), marches down the band, and emits one output row per input row:
//     10 LDS.64 + 20 HFMA2.SAT + 144 HFMA2 + 8 STG.64 per 32 outputs.  No block-wide barrier inside an item.
//   * zero padding is applied after the activation (halo pixels/rows outside the image are forced to 0).
//   * SE pooled sums: fp16 over the 8 pixels of a row strip, fp32 down the band, fixed-order reduction over the
//     strips in shared memory, one fp64 atomic per (item, channel): reproducible to ~1e-16.
#include <cuda_fp16.h>

#include <mutex>
#include <unordered_map>

#include "kernels.h"
#include "tc_common.cuh"
#include "tmap.h"

namespace lcm {

namespace {

using namespace tc;

constexpr int kCB = 128;          // channels per box
constexpr int kMaxRing = 8;

struct DwParams {
  CUtensorMap tmap;
  const float2* coef;    // [N][C] (a, b)
  const float* w;        // [9][C]
  __half* out;
  double* pool;          // [N][C]
  int N, H, W, C;
  int pxw, hseg, bandsX, segsY, cblocks, items, ring;
  uint32_t row_bytes;
};

__device__ __forceinline__ void tma_row(uint32_t dst, const CUtensorMap* map, int c0, int x, int y, int n, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(x), "r"(y), "r"(n), "r"(bar)
      : "memory");
}
template <int OFF>
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2+%3];" : "=r"(v.x), "=r"(v.y) : "r"(addr), "n"(OFF) : "memory");
  return v;
}
__device__ __forceinline__ __half2 as_h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t as_u32(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

struct ItemPos { int n, cb, bx, sy; };
__device__ __forceinline__ ItemPos decode(int item, const DwParams& p) {
  ItemPos q;
  q.sy = item % p.segsY; item /= p.segsY;
  q.bx = item % p.bandsX; item /= p.bandsX;
  q.cb = item % p.cblocks;
  q.n = item / p.cblocks;
  return q;
}

typedef __half2 Row[10][2];

// one output row from the three transformed input rows; `nvalid` (< 8 only on ragged right edges) masks pixels
template <bool kRagged>
__device__ __forceinline__ void emit_row(const Row& r0, const Row& r1, const Row& r2, const __half2 (&w6)[9][2], __half* orow,
                                         int C, int nvalid, bool cvalid, float (&psum)[4]) {
  __half2 s0 = __float2half2_rn(0.f), s1 = s0;
#pragma unroll
  for (int px = 0; px < 8; ++px) {
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
      if (cvalid) *reinterpret_cast<uint2*>(orow + (size_t)px * C) = make_uint2(as_u32(a0), as_u32(a1));
      s0 = __hadd2(s0, a0); s1 = __hadd2(s1, a1);
    }
  }
  const float2 f0 = __half22float2(s0), f1 = __half22float2(s1);
  psum[0] += f0.x; psum[1] += f0.y; psum[2] += f1.x; psum[3] += f1.y;
}

// Producer state of a CTA (lives in registers of thread 0 only): the TMA row stream runs ahead of the consumers,
// across work-item boundaries, by up to `ring` rows.
struct RowStream {
  int item, y, yb, c0, x, n;   // next row to issue: item, row, end row (exclusive) and the box coordinates
  int slot; uint32_t phase;
  int issued;                   // rows issued so far (all items)
  __device__ __forceinline__ void open(const DwParams& p) {   // position on `item` (caller checked item < items)
    const ItemPos q = decode(item, p);
    const int y0 = q.sy * p.hseg, y1 = min(p.H, y0 + p.hseg);
    y = max(y0 - 1, 0); yb = min(y1 + 1, p.H);
    c0 = q.cb * kCB; x = q.bx * p.pxw - 1; n = q.n;
  }
};

template <int STRIPS>
__global__ void __launch_bounds__(STRIPS * 32, 512 / (STRIPS * 32)) dwconv_stream_kernel(const __grid_constant__ DwParams p) {
  extern __shared__ uint8_t dsm_raw[];
  const uint32_t sraw = smem_u32(dsm_raw);
  const uint32_t sbase = (sraw + 127u) & ~127u;
  uint8_t* smem = dsm_raw + (sbase - sraw);
  const uint32_t ring_bytes = (uint32_t)p.ring * p.row_bytes;
  const uint32_t bar0 = sbase + ring_bytes;                       // full[ring], empty[ring]
  float* s_red = reinterpret_cast<float*>(smem + ring_bytes + 128);   // [2][STRIPS][128]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kMaxRing + s); };

  if (tid == 0) {
    for (int s = 0; s < p.ring; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), STRIPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmap)) : "memory");
  }
  __syncthreads();
  pdl_wait();
  pdl_trigger();

  // ---- producer (thread 0, interleaved with its consumer work): keep the ring full -----------------------
  // Its state lives in shared memory so that it costs the 255 other threads (and the hot loop) no registers.
  RowStream* ps = reinterpret_cast<RowStream*>(smem + ring_bytes + 128 + 2 * STRIPS * kCB * sizeof(float));
  // issue rows until `ring` rows are out beyond the `consumed` ones; a row the consumer needs right now is waited
  // for, look-ahead rows are only issued when their slot has already been released by every warp
  auto produce = [&](int consumed) {
    RowStream r = *ps;
    if (r.item >= p.items) return;
    const int target = consumed + p.ring;
    bool dirty = false;
    while (r.issued < target) {
      if (r.issued > consumed) { if (!mbar_test(empty_bar(r.slot), r.phase ^ 1u)) break; }
      else mbar_wait(empty_bar(r.slot), r.phase ^ 1u);
      mbar_expect_tx(full_bar(r.slot), p.row_bytes);
      tma_row(sbase + (uint32_t)r.slot * p.row_bytes, &p.tmap, r.c0, r.x, r.y, r.n, full_bar(r.slot));
      dirty = true;
      ++r.issued;
      if (++r.slot == p.ring) { r.slot = 0; r.phase ^= 1u; }
      if (++r.y == r.yb) {
        r.item += gridDim.x;
        if (r.item >= p.items) break;
        r.open(p);
      }
    }
    if (dirty) *ps = r;
  };
  if (tid == 0) {
    RowStream r;
    r.item = blockIdx.x; r.slot = 0; r.phase = 0; r.issued = 0;
    r.y = r.yb = r.c0 = r.x = r.n = 0;
    if (r.item < p.items) r.open(p);
    *ps = r;
    produce(0);
  }
  int consumed = 0;   // rows this warp has taken from the ring

  // ================================== consumers: warp = strip, lane = 4 channels =========================
  const int strip = warp;
  const uint32_t lane_off = (uint32_t)strip * 8u * 256u + (uint32_t)lane * 8u;
  int slot = 0; uint32_t phase = 0;
  int par = 0;
  const __half2 hz = __float2half2_rn(0.f);

  for (int item = blockIdx.x; item < p.items; item += gridDim.x, par ^= 1) {
    const ItemPos q = decode(item, p);
    const int c = q.cb * kCB + lane * 4;
    const bool cvalid = c < p.C;
    const int xs = q.bx * p.pxw + strip * 8;          // first output pixel of this strip
    const int nvalid = min(8, p.W - xs);              // <= 0: the strip lies outside the image (ragged last band)
    const int y0 = q.sy * p.hseg, y1 = min(p.H, y0 + p.hseg);

    __half2 a6[2], b6[2], w6[9][2];
    if (cvalid) {
      const float4* cf = reinterpret_cast<const float4*>(p.coef + (size_t)q.n * p.C + c);
      const float4 c01 = cf[0], c23 = cf[1];
      const float k = 1.f / 6.f;
      a6[0] = __floats2half2_rn(c01.x * k, c01.z * k); b6[0] = __floats2half2_rn(c01.y * k, c01.w * k);
      a6[1] = __floats2half2_rn(c23.x * k, c23.z * k); b6[1] = __floats2half2_rn(c23.y * k, c23.w * k);
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float4 wv = *reinterpret_cast<const float4*>(p.w + (size_t)t * p.C + c);
        w6[t][0] = __floats2half2_rn(6.f * wv.x, 6.f * wv.y);
        w6[t][1] = __floats2half2_rn(6.f * wv.z, 6.f * wv.w);
      }
    } else {
      a6[0] = a6[1] = b6[0] = b6[1] = hz;
#pragma unroll
      for (int t = 0; t < 9; ++t) w6[t][0] = w6[t][1] = hz;
    }
    // halo / ragged-edge pixels that must read as zero AFTER the activation: bit i <-> window pixel i (x = xs-1+i)
    uint32_t zmask = 0;
#pragma unroll
    for (int i = 0; i < 10; ++i) { const int x = xs - 1 + i; if (x < 0 || x >= p.W) zmask |= 1u << i; }

    auto load_row = [&](Row& r, int y) {
      if (y < 0 || y >= p.H) {
#pragma unroll
        for (int i = 0; i < 10; ++i) r[i][0] = r[i][1] = hz;
        return;
      }
      if (tid == 0) produce(consumed);
      mbar_wait(full_bar(slot), phase);
      const uint32_t a = sbase + (uint32_t)slot * p.row_bytes + lane_off;
      uint2 v[10];
      v[0] = lds64<0>(a); v[1] = lds64<256>(a); v[2] = lds64<512>(a); v[3] = lds64<768>(a); v[4] = lds64<1024>(a);
      v[5] = lds64<1280>(a); v[6] = lds64<1536>(a); v[7] = lds64<1792>(a); v[8] = lds64<2048>(a); v[9] = lds64<2304>(a);
#pragma unroll
      for (int i = 0; i < 10; ++i) {
        r[i][0] = __hfma2_sat(a6[0], as_h2(v[i].x), b6[0]);
        r[i][1] = __hfma2_sat(a6[1], as_h2(v[i].y), b6[1]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_bar(slot));
      ++consumed;
      if (++slot == p.ring) { slot = 0; phase ^= 1u; }
      if (zmask) {
        if ((zmask & 0x1feu) == 0) {   // the usual case (W % 8 == 0): only a neighbour column can lie outside the image
          if (zmask & 1u) r[0][0] = r[0][1] = hz;
          if (zmask & 0x200u) r[9][0] = r[9][1] = hz;
        } else {
#pragma unroll
          for (int i = 0; i < 10; ++i)
            if (zmask & (1u << i)) r[i][0] = r[i][1] = hz;
        }
      }
    };

    float psum[4] = {0.f, 0.f, 0.f, 0.f};
    Row ra, rb, rc;
    load_row(ra, y0 - 1);
    load_row(rb, y0);
    __half* orow = p.out + (((size_t)q.n * p.H + y0) * p.W + xs) * p.C + c;
    const size_t ostep = (size_t)p.W * p.C;
    const bool ragged = nvalid < 8;
    auto emit = [&](const Row& r0, const Row& r1, const Row& r2) {
      if (!ragged) emit_row<false>(r0, r1, r2, w6, orow, p.C, 8, cvalid, psum);
      else emit_row<true>(r0, r1, r2, w6, orow, p.C, nvalid, cvalid, psum);
      orow += ostep;
    };
    // the 3-row window rotates through (ra, rb, rc) with period 3; every item starts with fresh loads of its first
    // two rows, so nothing carries over between items
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

    // ---- pooled sums: fixed-order reduction over the strips, one fp64 atomic per channel ---------------
    float* red = s_red + par * (STRIPS * kCB);
    *reinterpret_cast<float4*>(red + strip * kCB + lane * 4) = make_float4(psum[0], psum[1], psum[2], psum[3]);
    __syncthreads();
    for (int ch = tid; ch < kCB; ch += STRIPS * 32) {
      if (q.cb * kCB + ch < p.C) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < STRIPS; ++i) s += red[i * kCB + ch];
        atomicAdd(&p.pool[(size_t)q.n * p.C + q.cb * kCB + ch], (double)s);
      }
    }
  }
}

std::mutex g_dw_mu;
struct DwKey {
  const void* ptr; int N, H, W, C, pxw;
  bool operator==(const DwKey& o) const { return ptr == o.ptr && N == o.N && H == o.H && W == o.W && C == o.C && pxw == o.pxw; }
};
struct DwKeyHash {
  size_t operator()(const DwKey& k) const {
    return std::hash<const void*>()(k.ptr) ^ ((size_t)k.N * 1315423911u) ^ ((size_t)k.H << 40) ^ ((size_t)k.W << 24) ^ ((size_t)k.C << 8) ^ (size_t)k.pxw;
  }
};
std::unordered_map<DwKey, CUtensorMap, DwKeyHash> g_dw_maps;

template <int STRIPS>
int launch_strips(DwParams& p, int num_sms, cudaStream_t st) {
  auto fn = dwconv_stream_kernel<STRIPS>;
  const int threads = STRIPS * 32;
  const size_t fixed = 128 + 128 + (size_t)2 * STRIPS * kCB * sizeof(float) + 64;   // alignment slack, barriers, reduction scratch, producer state
  // ring depth: as deep as ~96 KB allows (two CTAs of the widest band per SM), at most kMaxRing rows
  int ring = (int)((96 * 1024 - fixed) / p.row_bytes);
  if (ring > kMaxRing) ring = kMaxRing;
  if (ring < 3) return -1;
  p.ring = ring;
  const size_t smem = fixed + (size_t)ring * p.row_bytes;
  int ctas_per_sm;   // per (device, instantiation); smem depends on STRIPS only
  {
    if (ensure_dyn_smem((const void*)fn, 100 * 1024)) return -2;
    int* slot = device_cache_slot((const void*)fn);
    std::lock_guard<std::mutex> lk(g_dw_mu);
    if (!*slot) {
      int n = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, threads, smem) != cudaSuccess || n < 1) return -2;
      *slot = n;
    }
    ctas_per_sm = *slot;
  }
  const int max_ctas = num_sms * ctas_per_sm;
  // row segments: as long as possible while every CTA still gets >= 4 items (tail balance), never below 8 rows
  int hseg = p.H;
  auto items_for = [&](int hs) { return p.N * p.cblocks * p.bandsX * ((p.H + hs - 1) / hs); };
  while (hseg > 8 && items_for(hseg) < 4 * max_ctas) hseg = (hseg + 1) / 2;
  p.hseg = hseg;
  p.segsY = (p.H + hseg - 1) / hseg;
  p.items = items_for(hseg);
  const int grid = p.items < max_ctas ? p.items : max_ctas;
  launch_pdl(fn, dim3(grid), dim3(threads), smem, st, p);
  return 0;
}

}  // namespace

// in/out: fp16 NHWC [N][H][W][C]; C % 8 == 0
int launch_dwconv_f16(const void* in, const float2* coef, const float* w, void* out, double* pool, int N, int H, int W, int C,
                      int num_sms, cudaStream_t st) {
  if (C % 8 || N < 1 || H < 1 || W < 1) return -1;
  DwParams p;
  memset(&p, 0, sizeof(p));
  p.coef = coef; p.w = w; p.out = reinterpret_cast<__half*>(out); p.pool = pool;
  p.N = N; p.H = H; p.W = W; p.C = C;
  int strips = W >= 64 ? 8 : (W > 16 ? 4 : (W > 8 ? 2 : 1));
  p.pxw = strips * 8;
  p.bandsX = (W + p.pxw - 1) / p.pxw;
  p.cblocks = (C + kCB - 1) / kCB;
  p.row_bytes = (uint32_t)(p.pxw + 2) * kCB * 2;
  {
    DwKey key{in, N, H, W, C, p.pxw};
    std::lock_guard<std::mutex> lk(g_dw_mu);
    auto it = g_dw_maps.find(key);
    if (it != g_dw_maps.end()) {
      p.tmap = it->second;
    } else {
      cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
      cuuint64_t gstride[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
      cuuint32_t box[4] = {(cuuint32_t)kCB, (cuuint32_t)(p.pxw + 2), 1, 1};
      if (!encode_tmap(&p.tmap, TMAP_F16, 4, in, gdim, gstride, box, false)) return -3;
      if (g_dw_maps.size() > 4096) g_dw_maps.clear();
      g_dw_maps[key] = p.tmap;
    }
  }
  switch (strips) {
    case 8: return launch_strips<8>(p, num_sms, st);
    case 4: return launch_strips<4>(p, num_sms, st);
    case 2: return launch_strips<2>(p, num_sms, st);
    default: return launch_strips<1>(p, num_sms, st);
  }
}

}  // namespace lcm
