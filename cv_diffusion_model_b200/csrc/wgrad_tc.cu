// Weight gradient of the 1x1 convolutions on the 5th-gen tensor cores (training step, bf16 plan):
//
//     dW[n][k] += sum_m dY[m][n] * xform(A[m][k])          m = pixels (millions), k <= ~5000, n <= 2048
//
// Reference semantics: the weight gradients autograd produces for expand / project / skip / to_qkv / to_out
// (src/models/efficient_unet.py:174,186,199,265,267) inside LowLightTrainer.train_epoch (src/training/trainer.py:303-310).
//
// The reduction dimension is the PIXEL index, so both operands are read "transposed": an NHWC chunk [128 pixels][64 channels]
// lands in shared memory exactly as for the forward GEMM (TMA box, 128-byte swizzle) and is handed to tcgen05.mma as an
// MN-major operand (the same trick as the Gram MMA of gemm_expand.cu): D[k][n] (TMEM, lane = k, column = n) accumulates
// A_chunk^T . dY_chunk, 16 pixels per instruction, over ALL pixel tiles a CTA owns — one epilogue per CTA.
//
//   warps 0-7   XF    in-place prologue on the landed A chunks (ReLU6(a x + b) for the expand, gate * h2 for the project,
//                     a x + b for to_qkv; fp16 chunks are re-packed as bf16) — warps 0-3 then run the epilogue at the very
//                     end: TMEM -> fp32 atomics into the flat gradient buffer (reference state_dict layout)
//   warp  8     TMA   A chunks (1-2) + dY chunks (1-2) per 128-pixel tile into a ring of 3-4 stages
//   warp  9     MMA   8 instructions (M = 128, N = 64 / 128, K = 16 pixels) per tile; owns the TMEM allocation
//
// Work split: output tiles (128 A-channels x 128 dY-channels) x pixel splits, all CTAs co-resident.  At the two
// high-resolution levels (K <= 192) the op streams (K + N) * 2 bytes per pixel once: HBM-bound.
#include <cstdlib>
#include <cstring>

#include "kernels.h"
#include "tc_common.cuh"
#include "tmap.h"

namespace lcm {

namespace {

using namespace tc;

constexpr int kThreadsWT = 320;
constexpr int kXfThreadsWT = 256;            // 8 prologue warps (the first four also run the epilogue: TMEM lane quadrants)
constexpr int kTmaWarpWT = 8, kMmaWarpWT = 9;
constexpr int kNChunksWT = 2;                // dY chunks per output tile (N = 128): stages of 48-64 KB, 3-4 of them in flight
constexpr uint32_t kChunkWT = 16384;         // 128 pixels x 64 16-bit channels
constexpr uint32_t kSmemLimitWT = 232448;
constexpr int kMaxAChunksWT = 96;            // K up to 6144 channels
constexpr int kMaxStagesWT = 6;

struct WTParams {
  CUtensorMap tmap_a[LCM_MAX_SEGS];
  CUtensorMap tmap_y;
  const float2* coef[LCM_MAX_SEGS];
  int coef_ld[LCM_MAX_SEGS], coef_off[LCM_MAX_SEGS];
  float* dst[LCM_MAX_SEGS];                   // gradient tensor of segment s ([Nc][K_s], row stride dst_ld) or null
  int dst_ld[LCM_MAX_SEGS];
  uint32_t achunk[kMaxAChunksWT];             // seg | kvalid << 8 | mode << 24 | f16 << 28
  uint16_t achunk_c0[kMaxAChunksWT];          // first channel of the chunk inside its segment
  uint8_t mb_first[kMaxAChunksWT], mb_count[kMaxAChunksWT];   // M-block -> its 1-2 A chunks
  // CTAs of M-block mb: [mb_cta0[mb], mb_cta0[mb + 1]) = n_nblocks x mb_splits[mb] (N-block major).  The pixel range is
  // split in proportion to the bytes an M-block streams per pixel, so every CTA pulls about the same amount from HBM
  // (a 32-channel residual segment next to a 128-channel hidden segment used to get the same 74 CTAs: half the SMs idle).
  uint16_t mb_cta0[kMaxAChunksWT / 2 + 2], mb_splits[kMaxAChunksWT / 2 + 2];
  int n_mblocks, n_nblocks, nychunks, Nc;
  int m_tiles, P, stages;
  // per-image mode: the product of every image goes to img_dst[img][k][n] (fp32 atomics, k = global A channel) instead of
  // the weight-gradient tensors; the accumulator (two of them in TMEM) is drained at every image boundary of the CTA's range
  float* img_dst;
  int img_ktot;
  uint16_t achunk_k0[kMaxAChunksWT];          // global A channel of the chunk's first channel
  uint32_t stage_bytes, y_off, coef_smem_off, misc_off;
  // dense 3x3 conv (stride 1 or 2): achunk's segment field is the TAP, the A chunk of tap (ky, kx) is the image box shifted
  // by (kx - 1, ky - 1) (out-of-image elements zero-filled by TMA = the conv's padding; stride 2: tensor map with element
  // strides 2, every second pixel / row of a twice as large window); dW layout [Co][Ci][3][3].  Wimg = OUTPUT width.
  int conv, Wimg, box_w, box_h, Ci, cstride;
  CUtensorMap tmap_img;
  float* conv_dst;
};

__device__ __forceinline__ void tma_load_2d_wt(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_wt(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}
// MN-major operand, 128-byte swizzle (see gemm_expand.cu): 64 channels contiguous, 8-pixel groups 1024 B apart along K,
// 64-channel blocks `lbo` bytes apart
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(kThreadsWT, 1) wgrad_tc_kernel(const __grid_constant__ WTParams p) {
  extern __shared__ uint8_t wsm_raw[];
  const uint32_t sraw = smem_u32(wsm_raw);
  const uint32_t sbase = (sraw + 1023u) & ~1023u;
  uint8_t* smem = wsm_raw + (sbase - sraw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const uint32_t bar0 = sbase + p.misc_off;
  auto raw_bar = [&](int s) { return bar0 + 8u * s; };
  auto xf_bar = [&](int s) { return bar0 + 8u * (8 + s); };
  auto empty_bar = [&](int s) { return bar0 + 8u * (16 + s); };
  auto done_bar = [&](int a) { return bar0 + 8u * (24 + a); };      // accumulator a holds a finished product
  auto drained_bar = [&](int a) { return bar0 + 8u * (26 + a); };   // ... has been read out
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + p.misc_off + 256);
  float2* s_coef = reinterpret_cast<float2*>(smem + p.coef_smem_off);     // [2 chunks][64]

  // this CTA: output tile (M-block mb, N-block nb) and a contiguous range of pixel tiles
  int mb = 0;
  while (mb + 1 < p.n_mblocks && (int)blockIdx.x >= (int)p.mb_cta0[mb + 1]) ++mb;
  const int nsplit = p.mb_splits[mb];
  const int local = (int)blockIdx.x - (int)p.mb_cta0[mb];
  const int nb = local / nsplit, split = local - nb * nsplit;
  const int nA = p.mb_count[mb], a0 = p.mb_first[mb];
  const int y0 = nb * kNChunksWT;
  const int nY = min(kNChunksWT, p.nychunks - y0);
  const int t_begin = (int)((long long)p.m_tiles * split / nsplit);
  const int t_end = (int)((long long)p.m_tiles * (split + 1) / nsplit);

  if (warp == kTmaWarpWT && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(raw_bar(s), 1); mbar_init(xf_bar(s), kXfThreadsWT); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(done_bar(a), 1); mbar_init(drained_bar(a), 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmap_y)) : "memory");
  }
  if (warp == kMmaWarpWT) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_img_all = p.P >> 7;
  // a "product" ends at the last tile of the range and, in per-image mode, at the last tile of every image
  auto ends_product = [&](int t, int tin) { return t == t_end - 1 || (p.img_dst != nullptr && tin == tiles_per_img_all - 1); };

  if (warp < 8) {
    // ================================ XF: prologue of the A chunks, in place ====================================
    const int xt = tid;
    const int cu = (xt & 7) ^ ((xt >> 3) & 7);     // the 8-channel unit this thread owns in every row it touches
    const int tiles_per_img = p.P >> 7;
    int stage = 0; uint32_t phase = 0;
    int img = t_begin / tiles_per_img, tin = t_begin - img * tiles_per_img;
    int coef_img = -1;
    bool any_xf = false;
    // fp16 chunks (a block's hidden tensor) always pass through here: tcgen05.mma kind::f16 wants A and B in the SAME
    // 16-bit format, and dY is bf16 — the prologue re-packs them as bf16
    for (int ci = 0; ci < nA; ++ci) any_xf |= ((p.achunk[a0 + ci] >> 24) & 0x1f) != XF_NONE;
    // ---- epilogue (warps 0-3 = the four TMEM lane quadrants): product number e -> fp32 atomics --------------------
    int nprod = 0;
    auto epilogue = [&](int e, int im) {
      const int acc = e & 1;
      mbar_wait(done_bar(acc), (uint32_t)(e >> 1) & 1u);
      tc_fence_after();
      const int m = warp * 32 + lane;                 // TMEM lane = A channel within the M-block
      const int ci = m >> 6, cc = m & 63;
      const bool lane_ok = ci < nA;
      const uint32_t cd = p.achunk[a0 + (lane_ok ? ci : 0)];
      const int s = cd & 0xff, kvalid = (cd >> 8) & 0xff, c0 = p.achunk_c0[a0 + (lane_ok ? ci : 0)];
      float* dst = p.conv ? p.conv_dst : p.dst[s];
      float* idst = p.img_dst ? p.img_dst + ((size_t)im * p.img_ktot + p.achunk_k0[a0 + (lane_ok ? ci : 0)] + cc) * p.Nc : nullptr;
      const bool ok = lane_ok && cc < kvalid && (dst != nullptr || idst != nullptr);
      const int ldd = p.conv ? 0 : p.dst_ld[s];
      const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)acc * 128u;
      const int ncols = nY * 64;
      for (int c = 0; c < ncols; c += 16) {
        uint32_t r[16];
        tmem_ld16(lane_base + (uint32_t)c, r);
        tmem_wait_ld();
        if (ok) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = y0 * 64 + c + j;
            if (n < p.Nc) {
              if (idst) atomicAdd(idst + n, __uint_as_float(r[j]));
              else if (p.conv) atomicAdd(dst + ((size_t)n * p.Ci + c0 + cc) * 9 + s, __uint_as_float(r[j]));   // s = tap
              else atomicAdd(dst + (size_t)n * ldd + c0 + cc, __uint_as_float(r[j]));
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(drained_bar(acc));
    };
    for (int t = t_begin; t < t_end; ++t) {
      if (any_xf && img != coef_img) {
        bar_sync(1, kXfThreadsWT);
        for (int ci = 0; ci < nA; ++ci) {
          const uint32_t cd = p.achunk[a0 + ci];
          const int s = cd & 0xff, kvalid = (cd >> 8) & 0xff, c0 = p.achunk_c0[a0 + ci];
          if (((cd >> 24) & 0xf) == XF_NONE) continue;   // (an fp16 chunk without transform needs no coefficients)
          const float2* src = p.coef[s] + (size_t)img * p.coef_ld[s] + p.coef_off[s] + c0;
          for (int k = xt; k < kvalid; k += kXfThreadsWT) s_coef[ci * 64 + k] = src[k];
        }
        bar_sync(1, kXfThreadsWT);
        coef_img = img;
      }
      mbar_wait(raw_bar(stage), phase);
      if (any_xf) {
        for (int ci = 0; ci < nA; ++ci) {
          const uint32_t cd = p.achunk[a0 + ci];
          const int kvalid = (cd >> 8) & 0xff, mode = (cd >> 24) & 0xf;
          const bool f16 = (cd >> 28) & 1u;
          if ((mode == XF_NONE && !f16) || cu * 8 >= kvalid) continue;
          const uint32_t a_smem = sbase + (uint32_t)stage * p.stage_bytes + (uint32_t)ci * kChunkWT;
          float2 ab[8];
          if (mode != XF_NONE) {
            const float4* c4 = reinterpret_cast<const float4*>(s_coef + ci * 64 + cu * 8);
#pragma unroll
            for (int j = 0; j < 4; ++j) { const float4 c = c4[j]; ab[2 * j] = make_float2(c.x, c.y); ab[2 * j + 1] = make_float2(c.z, c.w); }
          }
#pragma unroll
          for (int i = 0; i < 1024 / kXfThreadsWT; ++i) {
            const uint32_t addr = a_smem + (uint32_t)(xt + i * kXfThreadsWT) * 16u;
            const uint4 v = lds128(addr);
            float f[8];
            if (f16) unpack8h(v, f); else unpack8(v, f);
            if (mode == XF_SCALE) {
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] *= ab[j].x;
            } else if (mode == XF_AFFINE_RELU6) {
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = fminf(fmaxf(fmaf(ab[j].x, f[j], ab[j].y), 0.f), 6.f);
            } else if (mode != XF_NONE) {
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = fmaf(ab[j].x, f[j], ab[j].y);
            }
            sts128(addr, pack8(f));     // always bf16 for the tensor core
          }
        }
        fence_proxy_async();
      }
      mbar_arrive(xf_bar(stage));
      if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      if (warp < 4 && ends_product(t, tin)) { epilogue(nprod, img); ++nprod; }
      if (++tin == tiles_per_img) { tin = 0; ++img; }
    }
  } else if (warp == kTmaWarpWT) {
    // ================================ TMA producer ===============================================================
    int stage = 0; uint32_t phase = 0;
    const uint32_t bytes = (uint32_t)(nA + nY) * kChunkWT;
    for (int t = t_begin; t < t_end; ++t) {
      mbar_wait_relaxed(empty_bar(stage), phase ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(raw_bar(stage), bytes);
        const uint32_t sb = sbase + (uint32_t)stage * p.stage_bytes;
        if (p.conv) {
          const int tpi = p.P >> 7;
          const int n = t / tpi, tin = t - n * tpi;
          int x0, y0;
          if (p.box_h == 1) { const int tpr = p.Wimg >> 7; y0 = tin / tpr; x0 = (tin - y0 * tpr) << 7; }
          else { y0 = tin * p.box_h; x0 = 0; }
          for (int ci = 0; ci < nA; ++ci) {
            const uint32_t cd = p.achunk[a0 + ci];
            const int tap = cd & 0xff, ky = tap / 3, kx = tap - ky * 3;
            tma_load_4d_wt(sb + (uint32_t)ci * kChunkWT, &p.tmap_img, (int)p.achunk_c0[a0 + ci], p.cstride * x0 + kx - 1, p.cstride * y0 + ky - 1, n,
                           raw_bar(stage));
          }
        } else {
          for (int ci = 0; ci < nA; ++ci) {
            const uint32_t cd = p.achunk[a0 + ci];
            tma_load_2d_wt(sb + (uint32_t)ci * kChunkWT, &p.tmap_a[cd & 0xff], (int)p.achunk_c0[a0 + ci], t * 128, raw_bar(stage));
          }
        }
        for (int j = 0; j < nY; ++j)
          tma_load_2d_wt(sb + p.y_off + (uint32_t)j * kChunkWT, &p.tmap_y, (y0 + j) * 64, t * 128, raw_bar(stage));
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1u; }
    }
  } else {
    // ================================ MMA issuer =================================================================
    // D = f32; A (activations) and B (dY) both MN-major; M = 128 (one chunk: the second 64-lane block aliases the first),
    // N = 64 nY; K = 16 pixels per instruction, 8 instructions per 128-pixel tile
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)((nY * 64) >> 3) << 17) | (8u << 24);
    const uint32_t a_lbo = nA == 2 ? kChunkWT : 0u;
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    int stage = 0; uint32_t phase = 0;
    int tin = t_begin % tiles_per_img_all;
    int e = 0;                 // product number: accumulator e & 1
    bool first = true;
    for (int t = t_begin; t < t_end; ++t) {
      if (first && e >= 2) { mbar_wait(drained_bar(e & 1), (uint32_t)((e >> 1) - 1) & 1u); tc_fence_after(); }
      mbar_wait(xf_bar(stage), phase);
      tc_fence_after();
      const bool last = ends_product(t, tin);
      if (elect_one()) {
        const uint32_t sb = sbase + (uint32_t)stage * p.stage_bytes;
        const uint64_t ad = desc_mn(sb, a_lbo), bd = desc_mn(sb + p.y_off, kChunkWT);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16(tmem_u + (uint32_t)(e & 1) * 128u, ad + (uint64_t)(k * (2048 >> 4)), bd + (uint64_t)(k * (2048 >> 4)), idesc,
                    (first && k == 0) ? 0u : 1u);
        umma_commit(empty_bar(stage));
        if (last) umma_commit(done_bar(e & 1));
      }
      __syncwarp();
      first = last;
      if (last) ++e;
      if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      if (++tin == tiles_per_img_all) tin = 0;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarpWT) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
  }
}

int launch_common(WTParams& p, int num_sms, cudaStream_t st);

}  // namespace

// Returns 0 when launched, non-zero when the shape is not covered (the caller falls back to the CUDA-core kernel):
// 16-bit operands, P % 128 == 0 (a 128-pixel tile lies inside one image: per-image prologue coefficients), segment widths
// multiples of 16, Nc a multiple of 16.
// img_dst (optional): per-image mode — img_dst[img][k][n] += sum over the image's pixels of xform(A)[p][k] dY[p][n]
// (fp32 [M / P][Ktot][Nc], zeroed by the caller); dst / dst_ld are ignored.
int launch_wgrad_tc(const GemmParams& g, const int* seg_dt, const void* dY, int dty, float* const* dst, const int* dst_ld,
                    int num_sms, cudaStream_t st, float* img_dst) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("LCM_NO_WGRAD_TC"); off = (e && atoi(e)) ? 1 : 0; }
  if (off || dty != DT_BF16 || g.P % 128 || g.M % g.P || g.M <= 0 || g.M > 0x7fffff00LL || g.Nc % 16 || g.nseg < 1 || g.nseg > LCM_MAX_SEGS)
    return -1;
  WTParams p;
  memset(&p, 0, sizeof(p));
  int nch = 0, nmb = 0, k0 = 0;
  p.img_dst = img_dst;
  for (int s = 0; s < g.nseg; ++s) {
    if (seg_dt[s] == DT_F32 || g.seg[s].K % 16 || g.seg[s].ld % 8) return -1;
    if (g.seg[s].mode != XF_NONE && !g.seg[s].coef) return -1;
    if (g.seg[s].mode == XF_AFFINE_SILU) return -1;
    const bool f16 = seg_dt[s] == DT_F16;
    if (!tmap_rows128(g.seg[s].A, g.M, g.seg[s].K, g.seg[s].ld, f16 ? TMAP_F16 : TMAP_BF16, &p.tmap_a[s])) return -3;
    p.coef[s] = g.seg[s].coef; p.coef_ld[s] = g.seg[s].coef_ld; p.coef_off[s] = g.seg[s].coef_off;
    p.dst[s] = dst[s]; p.dst_ld[s] = dst_ld[s];
    for (int c0 = 0; c0 < g.seg[s].K; c0 += 64) {
      if (nch >= kMaxAChunksWT) return -1;
      const int kv = g.seg[s].K - c0 < 64 ? g.seg[s].K - c0 : 64;
      p.achunk_c0[nch] = (uint16_t)c0;
      p.achunk_k0[nch] = (uint16_t)(k0 + c0);
      p.achunk[nch++] = (uint32_t)s | ((uint32_t)kv << 8) | ((uint32_t)g.seg[s].mode << 24) | ((f16 ? 1u : 0u) << 28);
    }
    k0 += g.seg[s].K;
  }
  p.img_ktot = k0;
  // M-blocks = consecutive pairs of chunks (they may span segments: mode, destination and source map are per chunk, and the
  // prologue hands every chunk to the tensor core as bf16)
  for (int c = 0; c < nch; c += 2) {
    p.mb_first[nmb] = (uint8_t)c;
    p.mb_count[nmb] = (uint8_t)(c + 1 < nch ? 2 : 1);
    ++nmb;
  }
  if (!tmap_rows128(dY, g.M, g.Nc, g.Nc, TMAP_BF16, &p.tmap_y)) return -3;
  p.n_mblocks = nmb;
  p.Nc = g.Nc;
  p.m_tiles = (int)(g.M / 128); p.P = g.P;
  return launch_common(p, num_sms, st);
}

namespace {
int launch_common(WTParams& p, int num_sms, cudaStream_t st) {
  const int nmb = p.n_mblocks;
  p.nychunks = (p.Nc + 63) / 64;
  p.n_nblocks = (p.nychunks + kNChunksWT - 1) / kNChunksWT;
  int maxA = 1;
  for (int i = 0; i < nmb; ++i) if (p.mb_count[i] > maxA) maxA = p.mb_count[i];
  const int maxY = p.nychunks < kNChunksWT ? p.nychunks : kNChunksWT;
  p.stage_bytes = (uint32_t)(maxA + maxY) * kChunkWT;
  p.y_off = (uint32_t)maxA * kChunkWT;
  const uint32_t fixed = 1024 /* coef */ + 1024 /* misc */ + 1024 /* align */;
  int stages = (int)((kSmemLimitWT - fixed) / p.stage_bytes);
  if (stages > kMaxStagesWT) stages = kMaxStagesWT;
  if (stages < 2) return -1;
  p.stages = stages;
  p.coef_smem_off = (uint32_t)stages * p.stage_bytes;
  p.misc_off = p.coef_smem_off + 1024;
  const uint32_t smem = p.misc_off + 1024 + 1024;
  // pixel splits per M-block: greedy, one more split to the M-block with the most bytes per pixel and CTA, while CTAs are left
  int cost[kMaxAChunksWT / 2 + 2], splits[kMaxAChunksWT / 2 + 2];
  const int ych = (p.Nc < kNChunksWT * 64 ? p.Nc : kNChunksWT * 64);
  for (int i = 0; i < nmb; ++i) {
    int ach = 0;
    for (int c = 0; c < p.mb_count[i]; ++c) ach += (int)((p.achunk[p.mb_first[i] + c] >> 8) & 0xff);
    cost[i] = ach + ych;
    splits[i] = 1;
  }
  int left = num_sms - nmb * p.n_nblocks;
  while (left >= p.n_nblocks) {
    int best = -1;
    for (int i = 0; i < nmb; ++i)
      if (splits[i] < p.m_tiles && (best < 0 || (long long)cost[i] * splits[best] > (long long)cost[best] * splits[i])) best = i;
    if (best < 0) break;
    ++splits[best];
    left -= p.n_nblocks;
  }
  int ctas = 0;
  for (int i = 0; i < nmb; ++i) {
    p.mb_cta0[i] = (uint16_t)ctas;
    p.mb_splits[i] = (uint16_t)splits[i];
    ctas += splits[i] * p.n_nblocks;
  }
  p.mb_cta0[nmb] = (uint16_t)ctas;
  if (ensure_dyn_smem_fn(wgrad_tc_kernel, kSmemLimitWT)) return -2;
  wgrad_tc_kernel<<<ctas, kThreadsWT, smem, st>>>(p);
  return 0;
}
}  // namespace

// dense 3x3, pad 1, stride 1 (the up-convolutions on the materialised bilinear output) or 2 (the downsamplers):
// in [N][H][W][Ci] bf16, dY [N][H/stride][W/stride][Co] bf16 -> dW [Co][Ci][3][3] (+=).  Tiles of 128 output pixels must be
// image boxes (128 | Wout, or Wout | 128 with whole rows).  Returns non-zero when the shape is not covered.
int launch_wgrad_conv3_tc(const void* in, const void* dY, float* dW, int N, int H, int W, int Ci, int Co, int stride, int num_sms,
                          cudaStream_t st) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("LCM_NO_WGRAD_TC"); off = (e && atoi(e)) ? 1 : 0; }
  if (stride != 1 && stride != 2) return -1;
  if (H % stride || W % stride) return -1;
  const int Ho = H / stride, Wo = W / stride;
  const long long P = (long long)Ho * Wo, M = P * N;
  const bool boxable = (Wo >= 128 ? Wo % 128 == 0 : 128 % Wo == 0) && P % 128 == 0;
  if (off || !boxable || Ci % 16 || Co % 16 || M > 0x7fffff00LL || 9 * ((Ci + 63) / 64) > kMaxAChunksWT) return -1;
  WTParams p;
  memset(&p, 0, sizeof(p));
  p.conv = 1; p.Wimg = Wo; p.Ci = Ci; p.conv_dst = dW; p.cstride = stride;
  p.box_w = Wo >= 128 ? 128 : Wo; p.box_h = 128 / p.box_w;
  {
    cuuint64_t gdim[4] = {(cuuint64_t)Ci, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t gstride[3] = {(cuuint64_t)Ci * 2, (cuuint64_t)W * Ci * 2, (cuuint64_t)H * W * Ci * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)(p.box_w * stride), (cuuint32_t)(p.box_h * stride), 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    if (!encode_tmap(&p.tmap_img, TMAP_BF16, 4, in, gdim, gstride, box, true, estr)) return -3;
  }
  int nch = 0, nmb = 0;
  for (int tap = 0; tap < 9; ++tap)
    for (int c0 = 0; c0 < Ci; c0 += 64) {
      const int kv = Ci - c0 < 64 ? Ci - c0 : 64;
      p.achunk_c0[nch] = (uint16_t)c0;
      p.achunk[nch++] = (uint32_t)tap | ((uint32_t)kv << 8);     // mode XF_NONE, bf16
    }
  for (int c = 0; c < nch; c += 2) {
    p.mb_first[nmb] = (uint8_t)c;
    p.mb_count[nmb] = (uint8_t)(c + 1 < nch ? 2 : 1);
    ++nmb;
  }
  if (!tmap_rows128(dY, M, Co, Co, TMAP_BF16, &p.tmap_y)) return -3;
  p.n_mblocks = nmb;
  p.Nc = Co;
  p.m_tiles = (int)(M / 128); p.P = (int)P;
  return launch_common(p, num_sms, st);
}

}  // namespace lcm
