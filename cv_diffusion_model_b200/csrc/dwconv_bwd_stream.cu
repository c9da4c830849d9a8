// Depthwise 3x3 backward of the tensor-core training plan (efficient_unet.py:212-223 reversed; trainer.py:303-310 runs it
// through autograd) — the streaming counterpart of dwconv_stream.cu.  The previous tile kernel (train_kernels.cu, kept
// for the fp32 plan) spent ~70 thread instructions per output element on fp32 halo tiles and ran at 1.2 TB/s.
//
//   dh2 = gate * dq + dpm/P                        (SE-scale backward, zero outside the image)
//   t   = sat(a2/6 h1 + b2/6) = relu6(a2 h1 + b2) / 6
//   dv[q]    = sum_tap w[tap] dh2[q - off(tap)]    du = dv [0 < t < 1]
//   dW[tap] += sum_q 6 t[q] dh2[q - off(tap)]      (= sum_p dh2[p] v[p + off(tap)]: the SAME operand as dv's, so one
//                                                   window of dh2 feeds both products; h1 is needed at the centre only)
//   T1 += sum_q du ;  T2 += sum_q du h1 = (sum du v - b2 sum du) / a2
//
// Arithmetic is packed fp16 (HFMA2: two channels per instruction).  Gradients are O(1/numel) and far below fp16's range,
// so se_bwd_vec hands in a power-of-two scale s per (image, channel) with (gate rms(dq) + |dpm/P|) s in (0.5, 1]:
// |dq| <= sqrt(P) rms(dq), hence |dh2 s| <= sqrt(P) <= 256 and no intermediate can overflow; du is STORED scaled, as fp16
// (11 significant bits instead of bf16's 8), and 1/s is folded into the A coefficient of the GroupNorm-backward pass that
// reads it (gn_bwd_coef).  fp16 partial sums never run over more than 8 pixels x kFold rows before they are folded into
// fp32 (dW: per-warp shared-memory accumulators; T1 / T2: registers).
//
// Structure: persistent CTAs, 2 per SM; a work item is (64-channel block, image, band of `pxw` pixel columns, segment of
// rows).  Thread 0 streams rows with two TMA boxes per row — dq [64 ch][pxw + 2 px] bf16 and, for centre rows, h1
// [64 ch][pxw px] fp16 — into a ring of row slots (full / empty mbarriers), running ahead across item boundaries.
// A warp is a strip of 8 pixels x 64 channels (lane = 2 channels: LDS.32 / STG.32, 128 contiguous bytes per pixel) and
// keeps a 3-row window of CONVERTED dh2 in registers.  A slot is released after its row has been the centre row.
#include <cuda_fp16.h>

#include <mutex>
#include <unordered_map>

#include "kernels.h"
#include "tc_common.cuh"
#include "tmap.h"

namespace lcm {

namespace {

using namespace tc;

constexpr int kCBb = 64;          // channels per box
constexpr int kMaxRingB = 8;
constexpr int kFold = 4;          // rows between folds of the fp16 weight-gradient partials

struct DwbParams {
  CUtensorMap tmap_g, tmap_h;
  const float4* cse;     // [N][C] (gate s, dpm/P s, 1/s, -)
  const float2* coef2;   // [N][C] (a2, b2)
  const float* w;        // [9][C]
  __half* du;            // [N][H][W][C], scaled by s
  double* t12;           // [N][C][2]
  float* dW;             // [C][9]
  int N, H, W, C;
  int pxw, hseg, bandsX, segsY, cblocks, items, ring;
  uint32_t g_bytes, h_bytes, stage_bytes;
};

__device__ __forceinline__ void tma_row4(uint32_t dst, const CUtensorMap* map, int c0, int x, int y, int n, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(x), "r"(y), "r"(n), "r"(bar)
      : "memory");
}
template <int OFF>
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF) : "memory");
  return v;
}
__device__ __forceinline__ __half2 u2h(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t h2u(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

struct ItemPosB { int n, cb, bx, sy; };
// the channel block is the SLOWEST index: a CTA's consecutive items share it, so its weight-gradient accumulators are
// flushed (global atomics) only a couple of times per launch
__device__ __forceinline__ ItemPosB decode_b(int item, const DwbParams& p) {
  ItemPosB q;
  q.sy = item % p.segsY; item /= p.segsY;
  q.bx = item % p.bandsX; item /= p.bandsX;
  q.n = item % p.N;
  q.cb = item / p.N;
  return q;
}

typedef __half2 RowB[10];

struct RowStreamB {
  int item, y, yb, y0, y1, c0, x, n;
  int slot; uint32_t phase;
  int issued;
  __device__ __forceinline__ void open(const DwbParams& p) {
    const ItemPosB q = decode_b(item, p);
    y0 = q.sy * p.hseg; y1 = min(p.H, y0 + p.hseg);
    y = max(y0 - 1, 0); yb = min(y1 + 1, p.H);
    c0 = q.cb * kCBb; x = q.bx * p.pxw; n = q.n;
  }
};

template <int STRIPS>
__global__ void __launch_bounds__(STRIPS * 32, 512 / (STRIPS * 32)) dwconv_bwd_stream_kernel(const __grid_constant__ DwbParams p) {
  extern __shared__ uint8_t bsm_raw[];
  const uint32_t sraw = smem_u32(bsm_raw);
  const uint32_t sbase = (sraw + 127u) & ~127u;
  uint8_t* smem = bsm_raw + (sbase - sraw);
  const uint32_t ring_bytes = (uint32_t)p.ring * p.stage_bytes;
  const uint32_t bar0 = sbase + ring_bytes;                                    // full[ring], empty[ring]
  float* s_dw = reinterpret_cast<float*>(smem + ring_bytes + 128);            // [STRIPS][9][64] fp32, private per warp
  float* s_red = s_dw + STRIPS * 9 * kCBb;                                     // [2][STRIPS][128] (T1 | T2 per channel)
  RowStreamB* ps = reinterpret_cast<RowStreamB*>(s_red + 2 * STRIPS * 2 * kCBb);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kMaxRingB + s); };

  if (tid == 0) {
    for (int s = 0; s < p.ring; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), STRIPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmap_g)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmap_h)) : "memory");
  }
  for (int i = lane; i < 9 * kCBb; i += 32) s_dw[warp * 9 * kCBb + i] = 0.f;
  __syncthreads();

  // ---- producer (thread 0, interleaved with its consumer work) -----------------------------------------------
  auto produce = [&](int consumed) {
    RowStreamB r = *ps;
    if (r.item >= p.items) return;
    const int target = consumed + p.ring;
    bool dirty = false;
    while (r.issued < target) {
      if (r.issued > consumed) { if (!mbar_test(empty_bar(r.slot), r.phase ^ 1u)) break; }
      else mbar_wait(empty_bar(r.slot), r.phase ^ 1u);
      const bool centre = r.y >= r.y0 && r.y < r.y1;
      const uint32_t dst = sbase + (uint32_t)r.slot * p.stage_bytes;
      mbar_expect_tx(full_bar(r.slot), p.g_bytes + (centre ? p.h_bytes : 0u));
      tma_row4(dst, &p.tmap_g, r.c0, r.x - 1, r.y, r.n, full_bar(r.slot));
      if (centre) tma_row4(dst + p.g_bytes, &p.tmap_h, r.c0, r.x, r.y, r.n, full_bar(r.slot));
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
    RowStreamB r;
    r.item = blockIdx.x; r.slot = 0; r.phase = 0; r.issued = 0;
    r.y = r.yb = r.y0 = r.y1 = r.c0 = r.x = r.n = 0;
    if (r.item < p.items) r.open(p);
    *ps = r;
    produce(0);
  }
  int consumed = 0;

  // ================================== consumers: warp = strip, lane = 2 channels ==============================
  const int strip = warp;
  const uint32_t lane_g = (uint32_t)strip * 8u * 128u + (uint32_t)lane * 4u;              // window pixel 0 of this strip
  int slot = 0; uint32_t phase = 0;
  int par = 0, cur_cb = -1;
  const __half2 hz = __float2half2_rn(0.f), h1v = __float2half2_rn(1.f);
  float* my_dw = s_dw + warp * 9 * kCBb + lane * 2;

  auto flush_dw = [&](int cb) {   // all warps: fixed-order sum over the strips, one global atomic per (channel, tap)
    __syncthreads();
    for (int i = tid; i < 9 * kCBb; i += STRIPS * 32) {
      float s = 0.f;
#pragma unroll
      for (int wv = 0; wv < STRIPS; ++wv) { s += s_dw[wv * 9 * kCBb + i]; s_dw[wv * 9 * kCBb + i] = 0.f; }
      const int t = i / kCBb, c = i - t * kCBb;
      atomicAdd(p.dW + (size_t)(cb * kCBb + c) * 9 + t, s);
    }
    __syncthreads();
  };

  for (int item = blockIdx.x; item < p.items; item += gridDim.x, par ^= 1) {
    const ItemPosB q = decode_b(item, p);
    if (q.cb != cur_cb) { if (cur_cb >= 0) flush_dw(cur_cb); cur_cb = q.cb; }
    const int c = q.cb * kCBb + lane * 2;
    const int xs = q.bx * p.pxw + strip * 8;
    const int nvalid = min(8, p.W - xs);
    const int y0 = q.sy * p.hseg, y1 = min(p.H, y0 + p.hseg);

    float gs0, gs1, ds0, ds1, k0, k1;          // gate s, dpm/P s, 6 / s
    __half2 a6, b6, w[9];
    {
      const float4 e0 = p.cse[(size_t)q.n * p.C + c], e1 = p.cse[(size_t)q.n * p.C + c + 1];
      gs0 = e0.x; ds0 = e0.y; k0 = 6.f * e0.z;
      gs1 = e1.x; ds1 = e1.y; k1 = 6.f * e1.z;
      const float4 cf = *reinterpret_cast<const float4*>(p.coef2 + (size_t)q.n * p.C + c);
      a6 = __floats2half2_rn(cf.x * (1.f / 6.f), cf.z * (1.f / 6.f));
      b6 = __floats2half2_rn(cf.y * (1.f / 6.f), cf.w * (1.f / 6.f));
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float2 wv = *reinterpret_cast<const float2*>(p.w + (size_t)t * p.C + c);
        w[t] = __floats2half2_rn(wv.x, wv.y);
      }
    }
    uint32_t zmask = 0;
#pragma unroll
    for (int i = 0; i < 10; ++i) { const int x = xs - 1 + i; if (x < 0 || x >= p.W) zmask |= 1u << i; }

    // converted dh2 row (window of 10 pixels); keep: the slot also holds the row's h1 and stays until it was the centre
    auto load_g = [&](RowB& r, int y, bool keep) -> int {
      if (y < 0 || y >= p.H) {
#pragma unroll
        for (int i = 0; i < 10; ++i) r[i] = hz;
        return -1;
      }
      if (tid == 0) produce(consumed);
      mbar_wait(full_bar(slot), phase);
      const uint32_t a = sbase + (uint32_t)slot * p.stage_bytes + lane_g;
      uint32_t v[10];
      v[0] = lds32<0>(a); v[1] = lds32<128>(a); v[2] = lds32<256>(a); v[3] = lds32<384>(a); v[4] = lds32<512>(a);
      v[5] = lds32<640>(a); v[6] = lds32<768>(a); v[7] = lds32<896>(a); v[8] = lds32<1024>(a); v[9] = lds32<1152>(a);
#pragma unroll
      for (int i = 0; i < 10; ++i) {
        const float f0 = __uint_as_float(v[i] << 16), f1 = __uint_as_float(v[i] & 0xffff0000u);
        r[i] = u2h(pack_f16(fmaf(gs0, f0, ds0), fmaf(gs1, f1, ds1)));
      }
      const int got = slot;
      if (!keep) {
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar(slot));
      }
      ++consumed;
      if (++slot == p.ring) { slot = 0; phase ^= 1u; }
      if (zmask) {
        if ((zmask & 0x1feu) == 0) {   // the usual case (W % 8 == 0): only a neighbour column can lie outside the image
          if (zmask & 1u) r[0] = hz;
          if (zmask & 0x200u) r[9] = hz;
        } else {
#pragma unroll
          for (int i = 0; i < 10; ++i)
            if (zmask & (1u << i)) r[i] = hz;
        }
      }
      return got;
    };

    __half2 dwp[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) dwp[t] = hz;
    float s1x = 0.f, s1y = 0.f, s2x = 0.f, s2y = 0.f;
    auto fold = [&]() {
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float2 f = __half22float2(dwp[t]);
        float2* d = reinterpret_cast<float2*>(my_dw + t * kCBb);
        float2 o = *d;
        o.x = fmaf(f.x, k0, o.x); o.y = fmaf(f.y, k1, o.y);
        *d = o;
        dwp[t] = hz;
      }
    };

    __half* orow = p.du + (((size_t)q.n * p.H + y0) * p.W + xs) * p.C + c;
    const size_t ostep = (size_t)p.W * p.C;
    int pend = -1;      // slot whose h1 row belongs to the next centre row
    int rows_since_fold = 0;

    // rp / rc / rn = dh2 rows y-1, y, y+1
    auto step = [&](const RowB& rp, const RowB& rc, RowB& rn, int y) {
      const bool next_centre = y + 1 < y1;
      const int ns = load_g(rn, y + 1, next_centre);
      // t of the centre row from h1 (slot `pend`), then the slot is free
      __half2 t[8];
      {
        const uint32_t a = sbase + (uint32_t)pend * p.stage_bytes + p.g_bytes + (uint32_t)strip * 8u * 128u + (uint32_t)lane * 4u;
        uint32_t hv[8];
        hv[0] = lds32<0>(a); hv[1] = lds32<128>(a); hv[2] = lds32<256>(a); hv[3] = lds32<384>(a);
        hv[4] = lds32<512>(a); hv[5] = lds32<640>(a); hv[6] = lds32<768>(a); hv[7] = lds32<896>(a);
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = (i < nvalid) ? __hfma2_sat(a6, u2h(hv[i]), b6) : hz;
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar(pend));
      }
      pend = next_centre ? ns : -1;
      __half2 dv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) dv[i] = hz;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const __half2* r = ky == 0 ? rn : (ky == 1 ? rc : rp);   // dh2 row y - ky + 1
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int tap = ky * 3 + kx;
#pragma unroll
          for (int px = 0; px < 8; ++px) {
            const __half2 g = r[px + 2 - kx];
            dv[px] = __hfma2(w[tap], g, dv[px]);
            dwp[tap] = __hfma2(t[px], g, dwp[tap]);
          }
        }
      }
      __half2 r1 = hz, r2 = hz;
#pragma unroll
      for (int px = 0; px < 8; ++px) {
        // ReLU6 backward: 0 < u < 6  <=>  0 < t < 1 (t = 0 also for pixels beyond a ragged right edge)
        const __half2 m = __hmul2(__hgt2(t[px], hz), __hlt2(t[px], h1v));
        const __half2 d = __hmul2(dv[px], m);
        if (px < nvalid) *reinterpret_cast<uint32_t*>(orow + (size_t)px * p.C) = h2u(d);
        r1 = __hadd2(r1, d);
        r2 = __hfma2(d, t[px], r2);
      }
      orow += ostep;
      const float2 f1 = __half22float2(r1), f2 = __half22float2(r2);
      s1x += f1.x; s1y += f1.y; s2x += f2.x; s2y += f2.y;
      if (++rows_since_fold == kFold) { fold(); rows_since_fold = 0; }
    };

    RowB ra, rb, rc;
    load_g(ra, y0 - 1, false);
    pend = load_g(rb, y0, true);
    for (int y = y0; y < y1; y += 3) {
      step(ra, rb, rc, y);
      if (y + 1 >= y1) break;
      step(rb, rc, ra, y + 1);
      if (y + 2 >= y1) break;
      step(rc, ra, rb, y + 2);
    }
    if (rows_since_fold) fold();

    // ---- T1 / T2 of this (image, channel block): fixed-order reduction over the strips, fp64 atomics --------
    {
      // sum du = s1 / s ; sum du v = 6 s2 / s ; sum du h1 = (sum du v - b2 sum du) / a2   (v = a2 h1 + b2 wherever du != 0)
      const float4 cf = *reinterpret_cast<const float4*>(p.coef2 + (size_t)q.n * p.C + c);
      const float i0 = k0 * (1.f / 6.f), i1 = k1 * (1.f / 6.f);
      const float t1x = s1x * i0, t1y = s1y * i1;
      const float t2x = fabsf(cf.x) > 1e-30f ? (s2x * k0 - cf.y * t1x) / cf.x : 0.f;
      const float t2y = fabsf(cf.z) > 1e-30f ? (s2y * k1 - cf.w * t1y) / cf.z : 0.f;
      float* red = s_red + par * (STRIPS * 2 * kCBb) + strip * 2 * kCBb;
      *reinterpret_cast<float4*>(red + lane * 4) = make_float4(t1x, t2x, t1y, t2y);     // [channel][T1, T2]
      __syncthreads();
      for (int i = tid; i < 2 * kCBb; i += STRIPS * 32) {
        float s = 0.f;
#pragma unroll
        for (int wv = 0; wv < STRIPS; ++wv) s += s_red[par * (STRIPS * 2 * kCBb) + wv * 2 * kCBb + i];
        atomicAdd(p.t12 + ((size_t)q.n * p.C + q.cb * kCBb) * 2 + i, (double)s);
      }
    }
  }
  if (cur_cb >= 0) flush_dw(cur_cb);
}

std::mutex g_dwb_mu;
struct DwbKey {
  const void* ptr; int N, H, W, C, box, dt;
  bool operator==(const DwbKey& o) const { return ptr == o.ptr && N == o.N && H == o.H && W == o.W && C == o.C && box == o.box && dt == o.dt; }
};
struct DwbKeyHash {
  size_t operator()(const DwbKey& k) const {
    return std::hash<const void*>()(k.ptr) ^ ((size_t)k.N * 1315423911u) ^ ((size_t)k.H << 40) ^ ((size_t)k.W << 24) ^ ((size_t)k.C << 8) ^
           ((size_t)k.box << 1) ^ (size_t)k.dt;
  }
};
std::unordered_map<DwbKey, CUtensorMap, DwbKeyHash> g_dwb_maps;

bool row_map(const void* ptr, int N, int H, int W, int C, int box_px, int dtype, CUtensorMap* out) {
  DwbKey key{ptr, N, H, W, C, box_px, dtype};
  std::lock_guard<std::mutex> lk(g_dwb_mu);
  auto it = g_dwb_maps.find(key);
  if (it != g_dwb_maps.end()) { *out = it->second; return true; }
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t gstride[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)kCBb, (cuuint32_t)box_px, 1, 1};
  CUtensorMap m;
  if (!encode_tmap(&m, dtype, 4, ptr, gdim, gstride, box, false)) return false;
  if (g_dwb_maps.size() > 4096) g_dwb_maps.clear();
  g_dwb_maps[key] = m;
  *out = m;
  return true;
}

template <int STRIPS>
int launch_strips_b(DwbParams& p, int num_sms, cudaStream_t st) {
  auto fn = dwconv_bwd_stream_kernel<STRIPS>;
  const int threads = STRIPS * 32;
  const size_t fixed = 128 + 128 + (size_t)STRIPS * 9 * kCBb * sizeof(float) + (size_t)2 * STRIPS * 2 * kCBb * sizeof(float) + 128;
  int ring = (int)((108 * 1024 - fixed) / p.stage_bytes);
  if (ring > kMaxRingB) ring = kMaxRingB;
  if (ring < 3) return -1;
  p.ring = ring;
  const size_t smem = fixed + (size_t)ring * p.stage_bytes;
  int ctas_per_sm;
  {
    if (ensure_dyn_smem((const void*)fn, 112 * 1024)) return -2;
    int* slot = device_cache_slot((const void*)fn);
    std::lock_guard<std::mutex> lk(g_dwb_mu);
    if (!*slot) {
      int n = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, threads, smem) != cudaSuccess || n < 1) return -2;
      *slot = n;
    }
    ctas_per_sm = *slot;
  }
  const int max_ctas = num_sms * ctas_per_sm;
  int hseg = p.H;
  auto items_for = [&](int hs) { return p.N * p.cblocks * p.bandsX * ((p.H + hs - 1) / hs); };
  while (hseg > 8 && items_for(hseg) < 4 * max_ctas) hseg = (hseg + 1) / 2;
  p.hseg = hseg;
  p.segsY = (p.H + hseg - 1) / hseg;
  p.items = items_for(hseg);
  const int grid = p.items < max_ctas ? p.items : max_ctas;
  fn<<<grid, threads, smem, st>>>(p);
  return 0;
}

}  // namespace

// dq: bf16 [N][H][W][C]; h1: fp16; du: fp16 (scaled, see the header); cse: [N][C] (gate s, dpm/P s, 1/s, -).
// Returns non-zero when the shape is not covered (C % 64).
int launch_dwconv_bwd_stream(const void* dq, const float4* cse, const void* h1, const float2* coef2, const float* w, void* du,
                             double* t12, float* dW, int N, int H, int W, int C, int num_sms, cudaStream_t st) {
  if (C % kCBb || N < 1 || H < 1 || W < 1) return -1;
  DwbParams p;
  memset(&p, 0, sizeof(p));
  p.cse = cse; p.coef2 = coef2; p.w = w; p.du = reinterpret_cast<__half*>(du); p.t12 = t12; p.dW = dW;
  p.N = N; p.H = H; p.W = W; p.C = C;
  const int strips = W >= 64 ? 8 : (W > 16 ? 4 : (W > 8 ? 2 : 1));
  p.pxw = strips * 8;
  p.bandsX = (W + p.pxw - 1) / p.pxw;
  p.cblocks = C / kCBb;
  p.g_bytes = (uint32_t)(p.pxw + 2) * kCBb * 2;
  p.h_bytes = (uint32_t)p.pxw * kCBb * 2;
  p.stage_bytes = p.g_bytes + p.h_bytes;
  if (!row_map(dq, N, H, W, C, p.pxw + 2, TMAP_BF16, &p.tmap_g)) return -3;
  if (!row_map(h1, N, H, W, C, p.pxw, TMAP_F16, &p.tmap_h)) return -3;
  switch (strips) {
    case 8: return launch_strips_b<8>(p, num_sms, st);
    case 4: return launch_strips_b<4>(p, num_sms, st);
    case 2: return launch_strips_b<2>(p, num_sms, st);
    default: return launch_strips_b<1>(p, num_sms, st);
  }
}

}  // namespace lcm
