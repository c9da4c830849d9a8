// tcgen05 / TMEM / TMA GEMM for the 1x1 convolutions and (as implicit GEMM) the dense 3x3
// convolutions of the EfficientUNet (efficient_unet.py:174,186,199,265,267 and :367,380-384).
//
//   out[m][n] = sum_k xform(A[m][k]) * W[n][k]  (+ bias[n]),   bf16 operands, fp32 accumulation in TMEM
//
// One persistent CTA per SM, warp-specialised (448 threads):
//   warps 0-3   epilogue   : tcgen05.ld accumulator -> (+bias) -> bf16 -> padded smem staging -> coalesced
//                            16-byte global stores + per-(image, channel) sum / sum-of-squares for the next
//                            GroupNorm (kept in smem across tiles, flushed with fp64 atomics on image change)
//   warp  4     MMA issuer : one thread issues tcgen05.mma (M=128, N=block_n, K=16) per 32-byte K step,
//                            tcgen05.commit releases smem stages / publishes the accumulator
//   warp  5     weights    : TMA bulk copies (cp.async.bulk, mbarrier complete_tx) of pre-swizzled weight
//                            chunks — resident in smem for the whole N tile when they fit, else streamed
//   warps 6-13  A producers: two groups of 128 threads alternate over K chunks; each loads 16-byte vectors
//                            of NHWC activations, applies the fused prologue (GroupNorm/FiLM affine + ReLU6,
//                            SE gate, identity for residual / skip operands; for 3x3 convs: tap gather with
//                            zero padding, stride 2, or bilinear x2 upsampling on load), and stores them into
//                            the 128B-swizzled K-major layout the UMMA descriptor expects.
// The A operand cannot come straight from TMA because ReLU6 sits between the normalisation and the
// contraction; weights are static, so they are swizzled once at load time (PackJob / WL_UMMA).
//
// Tiles are ordered N-tile-major so that a CTA's contiguous tile range walks along M with a fixed weight
// tile (weights stay in smem, statistics stay in one image for many tiles).
#include <cstdio>
#include <cstdlib>

#include "kernels.h"
#include "tc_common.cuh"

namespace lcm {

namespace {

constexpr int kThreads = 448;
constexpr int kEpiThreads = 128;
constexpr int kProdThreads = 256;
constexpr int kProdBase = 192;       // first producer thread
constexpr uint32_t kStageA = 16384;  // 128 rows x 128 bytes
constexpr int kMaxChunks = 160;
constexpr uint32_t kSmemLimit = 232448;

struct TcParams {
  GemmSeg seg[LCM_MAX_SEGS];
  int coef_base[LCM_MAX_SEGS];  // float2 index of the segment's coefficients in the smem table
  int nseg, ncoef;
  const bf16* W;
  bf16* out;
  double* stats;
  const float* bias;
  long long M, m_tiles;
  int P, Nc, block_n, n_tiles, nchunks;
  int resident, stages, fast;
  int conv_mode, Hin, Win, Hout, Wout, Ci;
  uint32_t stage_bytes, bres_off, stg_off, stg_stride, coef_off, misc_off;
  uint32_t chunk[kMaxChunks];  // seg/tap | kvalid << 8 | c0 << 16
  int debug;  // LCM_TC_DEBUG bit mask (bottleneck experiments only): 1 no A loads, 2 no stores, 4 no copy-out, 8 no stats
};

using namespace tc;

__global__ void __launch_bounds__(kThreads, 1) gemm_tc_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sraw = smem_u32(smem_raw);
  const uint32_t sbase = (sraw + 1023u) & ~1023u;   // SWIZZLE_128B operand tiles need 1024-byte alignment
  uint8_t* smem = smem_raw + (sbase - sraw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // misc region: barriers, TMEM pointer, statistics accumulators, bias
  uint8_t* misc = smem + p.misc_off;
  const uint32_t bar0 = sbase + p.misc_off;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (16 + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (32 + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (34 + a); };
  const uint32_t bres_bar = bar0 + 8u * 36;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 8 * 40);
  float* s_sum = reinterpret_cast<float*>(misc + 512);
  float* s_sq = s_sum + 256;
  float* s_bias = s_sq + 256;
  float* s_psum = reinterpret_cast<float*>(misc + 4096);   // per-row-group partial column sums [RG][block_n]
  float* s_psq = s_psum + 1024;
  float2* s_coef = reinterpret_cast<float2*>(smem + p.coef_off);

  if (warp == 5 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 128 + (p.resident ? 0 : 1));
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), kEpiThreads); }
    mbar_init(bres_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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

  const int m_tiles = (int)p.m_tiles;
  const long long total_tiles = p.m_tiles * p.n_tiles;
  const long long t_begin = total_tiles * blockIdx.x / gridDim.x;
  const int my_tiles = (int)(total_tiles * (blockIdx.x + 1) / gridDim.x - t_begin);
  const uint32_t b_chunk_bytes = (uint32_t)p.block_n * 128u;
  const int M = (int)p.M;

  if (warp >= 6) {
    // ================================ A producers ================================================
    const int ptid = tid - kProdBase;
    const int group = ptid >> 7, gt = ptid & 127;
    int cur_img = -1;
    TileIter ti; ti.init(t_begin, m_tiles, p.P);
    Ring ring{0, 0u, p.stages};
    int par = 0;   // chunk parity: group g fills the chunks with par == g
    for (int it = 0; it < my_tiles; ++it, ti.next(m_tiles, p.P)) {
      const int m0 = ti.m0;
      if (p.ncoef > 0 && p.fast && ti.img != cur_img) {
        bar_sync(1, kProdThreads);
        for (int s = 0; s < p.nseg; ++s) {
          if (p.seg[s].mode == XF_NONE) continue;
          const float2* src = p.seg[s].coef + (size_t)ti.img * p.seg[s].coef_ld + p.seg[s].coef_off;
          for (int k = ptid; k < p.seg[s].K; k += kProdThreads) s_coef[p.coef_base[s] + k] = src[k];
        }
        bar_sync(1, kProdThreads);
        cur_img = ti.img;
      }
      // conv: this thread's output pixel
      int cy = 0, cx = 0, cn = 0;
      const int cm = m0 + gt;
      if (p.conv_mode >= 0) {
        const int q = cm / p.Wout;
        cx = cm - q * p.Wout;
        cn = q / p.Hout;
        cy = q - cn * p.Hout;
      }
      for (int ci = 0; ci < p.nchunks; ++ci, ring.advance(), par ^= 1) {
        if (par != group) continue;
        const int stage = ring.stage;
        const uint32_t cd = p.chunk[ci];
        const int sidx = cd & 0xff, kvalid = (cd >> 8) & 0xff, c0 = cd >> 16;
        const int upr = kvalid >> 3;
        const uint32_t a_smem = sbase + stage * p.stage_bytes;
        mbar_wait(empty_bar(stage), ring.phase ^ 1u);
        if (p.conv_mode < 0) {
          // ---- 1x1: 16-byte units interleaved over threads (8 lanes cover one 128-byte row) ----------
          const GemmSeg& sg = p.seg[sidx];
          const bf16* A = reinterpret_cast<const bf16*>(sg.A);
          const int total = 128 * upr;
          const int mode = sg.mode;
          uint4 v[8];
          int rows[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int u = gt + i * 128;
            v[i] = make_uint4(0u, 0u, 0u, 0u);
            rows[i] = div_upr(u, upr);
            if (u < total) {
              const int cu = u - rows[i] * upr;
              const int m = m0 + rows[i];
              if (m < M && !(p.debug & 1)) v[i] = ldg_stream(A + (long long)m * sg.ld + c0 + cu * 8);
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int u = gt + i * 128;
            if (u < total && !(p.debug & 16)) {
              const int row = rows[i], cu = u - row * upr;
              uint4 o = v[i];
              if (mode != XF_NONE && m0 + row < M) {
                if (p.fast) {
                  o = apply_xform(o, s_coef + p.coef_base[sidx] + c0 + cu * 8, mode);
                } else {
                  const int img = (m0 + row) / p.P;
                  __align__(16) float2 ab[8];
                  const float2* src = sg.coef + (size_t)img * sg.coef_ld + sg.coef_off + c0 + cu * 8;
#pragma unroll
                  for (int j = 0; j < 8; ++j) ab[j] = src[j];
                  o = apply_xform(o, ab, mode);
                }
              }
              sts128(a_smem + row * 128 + ((cu ^ (row & 7)) << 4), o);
            }
          }
        } else {
          // ---- 3x3 tap gather: one output pixel (row) per thread -----------------------------------
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
        }
        fence_proxy_async();           // generic-proxy smem writes -> visible to the tensor core (async proxy)
        mbar_arrive(full_bar(stage));
      }
    }
  } else if (warp == 5) {
    // ================================ weight loader (TMA bulk copies), one thread =====================
    if (lane == 0) {
      int cur_nt = -1;
      TileIter ti; ti.init(t_begin, m_tiles, p.P);
      Ring ring{0, 0u, p.stages};
      int acc = 0; uint32_t aphase = 0;   // accumulator stage / phase of the PREVIOUS tile
      for (int it = 0; it < my_tiles; ++it, ti.next(m_tiles, p.P)) {
        const bf16* wt = p.W + (size_t)ti.n_tile * p.nchunks * p.block_n * 64;
        if (p.resident) {
          if (ti.n_tile != cur_nt) {
            if (it > 0) mbar_wait(tfull_bar(acc), aphase);   // MMAs of the previous tile done: old weights dead
            mbar_expect_tx(bres_bar, (uint32_t)p.nchunks * b_chunk_bytes);
            for (int ci = 0; ci < p.nchunks; ++ci)
              bulk_g2s(sbase + p.bres_off + ci * b_chunk_bytes, wt + (size_t)ci * p.block_n * 64, b_chunk_bytes, bres_bar);
            cur_nt = ti.n_tile;
          }
          if (it > 0) { acc ^= 1; if (acc == 0) aphase ^= 1u; }   // now describes tile `it`
        } else {
          for (int ci = 0; ci < p.nchunks; ++ci, ring.advance()) {
            mbar_wait(empty_bar(ring.stage), ring.phase ^ 1u);
            mbar_expect_tx(full_bar(ring.stage), b_chunk_bytes);
            bulk_g2s(sbase + ring.stage * p.stage_bytes + kStageA, wt + (size_t)ci * p.block_n * 64, b_chunk_bytes,
                     full_bar(ring.stage));
          }
        }
      }
    }
  } else if (warp == 4) {
    // ================================ MMA issuer, one thread ==========================================
    if (lane == 0) {
      // instruction descriptor: D fp32, A/B bf16, both K-major, N = block_n, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) | (8u << 24);
      int cur_nt = -1;
      uint32_t bres_phase = 0;
      TileIter ti; ti.init(t_begin, m_tiles, p.P);
      Ring ring{0, 0u, p.stages};
      int acc = 0; uint32_t aphase = 0;
      for (int it = 0; it < my_tiles; ++it, ti.next(m_tiles, p.P)) {
        if (p.resident && ti.n_tile != cur_nt) {
          mbar_wait(bres_bar, bres_phase);
          bres_phase ^= 1u;
          cur_nt = ti.n_tile;
        }
        mbar_wait(tempty_bar(acc), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
        for (int ci = 0; ci < p.nchunks; ++ci, ring.advance()) {
          const int stage = ring.stage;
          mbar_wait(full_bar(stage), ring.phase);
          tc_fence_after();
          const int ksteps = (p.chunk[ci] >> 12) & 0xf;   // kvalid / 16
          const uint32_t a_addr = sbase + stage * p.stage_bytes;
          const uint32_t b_addr = p.resident ? sbase + p.bres_off + ci * b_chunk_bytes : a_addr + kStageA;
          const uint64_t ad = umma_desc(a_addr), bd = umma_desc(b_addr);
          for (int k = 0; k < ksteps; ++k)
            umma_bf16(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (ci | k) != 0 ? 1u : 0u);
          umma_commit(empty_bar(stage));
          if (ci == p.nchunks - 1) umma_commit(tfull_bar(acc));
        }
        acc ^= 1;
        if (acc == 0) aphase ^= 1u;
      }
    }
  } else {
    // ================================ epilogue (warps 0-3) ==========================================
    const int upr = p.block_n >> 3;
    const int RG = 128 / upr;
    const bool active = tid < upr * RG;
    const int cu = tid % upr, rg = tid / upr;
    const uint32_t stg = sbase + p.stg_off;
    const bool do_stats = p.stats != nullptr && !(p.debug & 8);
    int cur_img = -1, cur_nt = -1;
    auto flush = [&]() {   // all statistics of (cur_img, cur_nt) -> global, fp64 atomics
      for (int c = tid; c < p.block_n; c += kEpiThreads) {
        double* d = p.stats + ((size_t)cur_img * p.Nc + (size_t)cur_nt * p.block_n + c) * 2;
        atomicAdd(d, (double)s_sum[c]);
        atomicAdd(d + 1, (double)s_sq[c]);
        s_sum[c] = 0.f;
        s_sq[c] = 0.f;
      }
    };
    TileIter ti; ti.init(t_begin, m_tiles, p.P);
    int acc = 0; uint32_t aphase = 0;
    for (int it = 0; it < my_tiles; ++it, ti.next(m_tiles, p.P)) {
      const int n_tile = ti.n_tile, m0 = ti.m0, n0 = n_tile * p.block_n;
      if (do_stats && p.fast && (ti.img != cur_img || n_tile != cur_nt)) {
        if (cur_img >= 0) flush();   // columns are thread-owned: program order suffices
        cur_img = ti.img;
      }
      if (p.bias && n_tile != cur_nt) {
        for (int c = tid; c < p.block_n; c += kEpiThreads) s_bias[c] = p.bias[n0 + c];
        bar_sync(2, kEpiThreads);
      }
      cur_nt = n_tile;

      mbar_wait(tfull_bar(acc), aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)acc * 256u;
      const uint32_t my_row = stg + (uint32_t)tid * p.stg_stride;
      for (int cb = 0; cb < ((p.debug & 32) ? 0 : p.block_n); cb += 32) {
        // two 16-column TMEM loads in flight per wait
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
          sts128(dst, make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7])));
          sts128(dst + 16, make_uint4(pack_bf16(f[8], f[9]), pack_bf16(f[10], f[11]), pack_bf16(f[12], f[13]), pack_bf16(f[14], f[15])));
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));   // accumulator drained: the MMA warp may start the tile after next
      bar_sync(2, kEpiThreads);       // staging complete
      if (active && !(p.debug & 4)) {
        float cs[8], cq[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { cs[j] = 0.f; cq[j] = 0.f; }
        const int rows_valid = min(128, M - m0);
        uint32_t src = stg + (uint32_t)rg * p.stg_stride + cu * 16;
        bf16* dst = p.out + (long long)(m0 + rg) * p.Nc + n0 + cu * 8;
        const uint32_t src_step = (uint32_t)RG * p.stg_stride;
        const long long dst_step = (long long)RG * p.Nc;
#pragma unroll 4
        for (int r = rg; r < rows_valid; r += RG, src += src_step, dst += dst_step) {
          const uint4 v = lds128(src);
          if (!(p.debug & 2)) *reinterpret_cast<uint4*>(dst) = v;
          if (do_stats) {
            float f[8];
            unpack8(v, f);
            if (p.fast) {
#pragma unroll
              for (int j = 0; j < 8; ++j) { cs[j] += f[j]; cq[j] = fmaf(f[j], f[j], cq[j]); }
            } else {
              double* d = p.stats + ((size_t)((m0 + r) / p.P) * p.Nc + n0 + cu * 8) * 2;
#pragma unroll
              for (int j = 0; j < 8; ++j) { atomicAdd(d + 2 * j, (double)f[j]); atomicAdd(d + 2 * j + 1, (double)f[j] * f[j]); }
            }
          }
        }
        if (do_stats && p.fast) {   // partial sums of this thread's rows; folded (without atomics) after the barrier
          float4* ps = reinterpret_cast<float4*>(s_psum + rg * p.block_n + cu * 8);
          float4* pq = reinterpret_cast<float4*>(s_psq + rg * p.block_n + cu * 8);
          ps[0] = make_float4(cs[0], cs[1], cs[2], cs[3]); ps[1] = make_float4(cs[4], cs[5], cs[6], cs[7]);
          pq[0] = make_float4(cq[0], cq[1], cq[2], cq[3]); pq[1] = make_float4(cq[4], cq[5], cq[6], cq[7]);
        }
      }
      bar_sync(2, kEpiThreads);       // staging free, partial statistics visible
      if (do_stats && p.fast && !(p.debug & 4)) {
        for (int c = tid; c < p.block_n; c += kEpiThreads) {   // column c is owned by thread c % 128
          float a = 0.f, b = 0.f;
#pragma unroll 4
          for (int g = 0; g < RG; ++g) { a += s_psum[g * p.block_n + c]; b += s_psq[g * p.block_n + c]; }
          s_sum[c] += a;
          s_sq[c] += b;
        }
      }
      acc ^= 1;
      if (acc == 0) aphase ^= 1u;
    }
    if (do_stats && p.fast && cur_img >= 0) flush();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

bool g_attr_set = false;

}  // namespace

int gemm_tc_pick_block_n(int Nc) {
  static int forced = -1;   // experiment switch: LCM_BLOCK_N=<n> caps the N tile
  if (forced < 0) { const char* e = getenv("LCM_BLOCK_N"); forced = e ? atoi(e) : 0; }
  for (int bn = (forced >= 16 && forced <= 256) ? forced / 16 * 16 : 256; bn >= 16; bn -= 16)
    if (Nc % bn == 0) return bn;
  return 0;
}

int launch_gemm_tc_v1(const GemmParams& g, const ConvGeom& cg, int block_n, int num_sms, cudaStream_t st) {
  TcParams p{};
  if (block_n < 16 || block_n > 256 || block_n % 16 || g.Nc % block_n) return -1;
  if (g.M <= 0 || g.M > 0x7fffff00LL || g.P <= 0) return -1;   // 32-bit row indices inside the kernel
  p.nseg = g.nseg;
  p.W = reinterpret_cast<const bf16*>(g.W);
  p.out = reinterpret_cast<bf16*>(g.out);
  p.stats = g.stats;
  p.M = g.M; p.P = g.P; p.Nc = g.Nc; p.block_n = block_n;
  p.n_tiles = g.Nc / block_n;
  p.m_tiles = (g.M + 127) / 128;
  p.fast = (g.P % 128 == 0) ? 1 : 0;
  p.conv_mode = cg.mode;
  int nch = 0, ncoef = 0;
  if (cg.mode < 0) {
    p.bias = nullptr;
    for (int s = 0; s < g.nseg; ++s) {
      p.seg[s] = g.seg[s];
      if (g.seg[s].K % 16) return -1;
      p.coef_base[s] = -1;
      if (g.seg[s].mode != XF_NONE) { p.coef_base[s] = ncoef; ncoef += g.seg[s].K; }
      for (int c0 = 0; c0 < g.seg[s].K; c0 += 64) {
        if (nch >= kMaxChunks) return -1;
        const int kv = g.seg[s].K - c0 < 64 ? g.seg[s].K - c0 : 64;
        p.chunk[nch++] = (uint32_t)s | ((uint32_t)kv << 8) | ((uint32_t)c0 << 16);
      }
    }
  } else {
    p.bias = cg.bias;
    p.seg[0] = g.seg[0];
    p.Hin = cg.Hin; p.Win = cg.Win; p.Hout = cg.Hout; p.Wout = cg.Wout; p.Ci = cg.Ci;
    if (cg.Ci % 16) return -1;
    for (int tap = 0; tap < 9; ++tap)
      for (int c0 = 0; c0 < cg.Ci; c0 += 64) {
        if (nch >= kMaxChunks) return -1;
        const int kv = cg.Ci - c0 < 64 ? cg.Ci - c0 : 64;
        p.chunk[nch++] = (uint32_t)tap | ((uint32_t)kv << 8) | ((uint32_t)c0 << 16);
      }
  }
  p.nchunks = nch;
  p.ncoef = ncoef;
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("LCM_TC_DEBUG"); dbg = e ? atoi(e) : 0; } p.debug = dbg; }
  // shared-memory layout
  const uint32_t b_chunk = (uint32_t)block_n * 128u;
  const uint32_t stg_stride = (uint32_t)block_n * 2u + 16u;
  const uint32_t stg_bytes = (128u * stg_stride + 1023u) & ~1023u;
  const uint32_t coef_bytes = ((uint32_t)ncoef * 8u + 1023u) & ~1023u;
  const uint32_t misc_bytes = 4096 + 8192;
  const uint32_t fixed = stg_bytes + coef_bytes + misc_bytes + 1024;  // +1024: base alignment slack
  const uint32_t bres = (uint32_t)nch * b_chunk;
  p.resident = 0;
  if (bres <= 98304 && fixed + bres + 3 * kStageA <= kSmemLimit) p.resident = 1;
  p.stage_bytes = kStageA + (p.resident ? 0u : b_chunk);
  const uint32_t avail = kSmemLimit - fixed - (p.resident ? bres : 0u);
  int stages = (int)(avail / p.stage_bytes);
  if (stages > 8) stages = 8;
  if (stages < 2) return -1;
  p.stages = stages;
  uint32_t off = (uint32_t)stages * p.stage_bytes;
  p.bres_off = off; off += p.resident ? bres : 0u;
  p.stg_off = off; off += stg_bytes;
  p.stg_stride = stg_stride;
  p.coef_off = off; off += coef_bytes;
  p.misc_off = off; off += misc_bytes;
  const uint32_t smem_bytes = off + 1024;
  if (smem_bytes > kSmemLimit) return -1;
  if (!g_attr_set) {
    if (cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit) != cudaSuccess) return -2;
    g_attr_set = true;
  }
  const long long tiles = p.m_tiles * p.n_tiles;
  const int grid = (int)(tiles < num_sms ? tiles : num_sms);
  gemm_tc_kernel<<<grid, kThreads, smem_bytes, st>>>(p);
  return 0;
}

}  // namespace lcm
