// Project GEMM of the level-0 inverted-residual blocks (efficient_unet.py:100,226,230-234) as a STREAMING kernel:
//
//   out[p][n] = sum_{k < K0} (gate_k h2[p][k]) Wp[n][k] + sum_{k < K1} x[p][k] Ws[n][k]      K0 = 128 (fp16 hidden tensor, SE gate),
//                                                                                            K1 = 32 (bf16 block input: skip conv or identity), N = 32
//   + per-(image, channel) sum / sum^2 of the stored bf16 output (the next block's GroupNorm1)
//
// 5 bytes are read per byte written and the MMA work is tiny (40 FLOP/B), so the op should run at the rate the memory system
// gives a 5:1 stream — 6.7 TB/s with plain vector loads (tests/diag/write_bw.cu).  The general tcgen05 kernel (gemm_tc2.cu) does
// 4.8 TB/s here, and not because of memory: it runs at the same speed with its activation loads switched off; its ~3000 cycles
// per 128-pixel tile are the hand-offs of its five warp roles (profiles/r02_experiments_fused_path_and_gemm_issue.txt, sections 5, 8).
// This kernel has no roles and no block barriers: every warp walks its own list of 16-pixel tiles —
//   10 x cp.async (16 B) per lane into a warp-private, double-buffered shared-memory tile (the next tile's loads are in flight
//   while this one is multiplied; no registers) -> ldmatrix -> SE gate on the A fragments (4 HMUL2 per K step) -> 40 x mma.sync
//   m16n8k16 (fp16 operands for the h2 part, bf16 for the x part, one fp32 accumulator) -> bf16 -> statistics in registers ->
//   transposed through shared memory -> 2 x STG.128 per lane (1 KB contiguous per warp)
// at 16 warps per SM.  Weights [N][K] live in shared memory (ldmatrix B fragments).  Persistent CTAs over contiguous tile ranges;
// statistics are reduced over the lanes of a warp at image boundaries only and added with fp64 atomics (a few exact adds per entry).
#include <cuda_fp16.h>

#include "kernels.h"
#include "tc_common.cuh"

namespace lcm {

namespace {

using namespace tc;


struct PsParams {
  const __half* h2;        // [M][K0] fp16
  const bf16* x;           // [M][K1] bf16
  const float2* gate;      // [N][gate_ld] (gate, 0)
  int gate_ld, gate_off;
  const uint16_t* W;       // [Nc][K0 + K1] row-major 16-bit: fp16 for k < K0, bf16 beyond
  bf16* out;               // [M][Nc]
  double* stats;           // [N][Nc][2] (+=)
  int P;                   // pixels per image (multiple of 16)
  long long tiles;         // 16-pixel tiles in total
};

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t hmul2_u(uint32_t a, uint32_t b) {
  const __half2 r = __hmul2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// KS0 / KS1: K steps of 16 in the fp16 / bf16 segment; NT: n-tiles of 8 output channels (even); kPsWarps warps per CTA (one CTA per
// SM), NBUF activation tiles per warp: the loads of the next NBUF - 1 tiles are in flight while one is multiplied
// kBReg: the B fragments (all weights: 8 registers per K step) live in registers for the whole kernel instead of being re-read from
// shared memory for every 16-pixel tile (20 of the 52 load/store instructions and 10 of the 22 KB of shared-memory traffic per tile)
template <int KS0, int KS1, int NT, int kPsWarps, int NBUF, bool kBReg>
__global__ void __launch_bounds__(kPsWarps * 32, 1) proj_stream_kernel(const PsParams p) {
  constexpr int K0 = KS0 * 16, K1 = KS1 * 16, K = K0 + K1, Nc = NT * 8;
  constexpr int UPP = K / 8, U0 = K0 / 8;                 // 16-byte units per pixel (all / fp16 part)
  constexpr int U = 16 * UPP / 32;                        // units per lane and tile
  constexpr uint32_t PITCH = K * 2 + 16;                  // activation tile row (conflict-free ldmatrix: odd multiple of 16 B mod 128)
  constexpr uint32_t WPITCH = K * 2 + 16;
  constexpr uint32_t OPITCH = Nc * 2 + 16;
  constexpr uint32_t WARP_SMEM = NBUF * 16 * PITCH;       // ring of activation tiles; the output tile re-uses the one just consumed
  static_assert(16 * OPITCH <= 16 * PITCH, "output staging fits an activation tile");
  static_assert((16 * UPP) % 32 == 0 && NT % 2 == 0, "shape");
  extern __shared__ __align__(16) uint8_t ps_raw[];
  uint8_t* s_w = ps_raw;                                                    // [Nc][WPITCH]
  uint8_t* s_tiles = s_w + Nc * WPITCH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const uint32_t tile_s = smem_u32(s_tiles) + (uint32_t)warp * WARP_SMEM;
  const uint32_t w_s = smem_u32(s_w);

  // weights: do not depend on the previous kernel
  for (int i = tid; i < Nc * UPP; i += kPsWarps * 32) {
    const int n = i / UPP, u = i - n * UPP;
    *reinterpret_cast<uint4*>(s_w + n * WPITCH + u * 16) = *reinterpret_cast<const uint4*>(p.W + (size_t)n * K + u * 8);
  }
  __syncthreads();
  pdl_wait();
  pdl_trigger();

  const int tiles_per_img = p.P / 16;
  const long long tb = p.tiles * blockIdx.x / gridDim.x, te = p.tiles * (blockIdx.x + 1) / gridDim.x;
  // ldmatrix lane addresses.  A (x4): matrices (rows 0-7 | 8-15) x (k 0-7 | 8-15) in the order a0..a3 of the MMA.
  const uint32_t a_lane = tile_s + (uint32_t)((lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + (uint32_t)(lane >> 4) * 16u;
  // B (x4) for an n-tile pair: (n 0-7, k 0-7), (n 0-7, k 8-15), (n 8-15, k 0-7), (n 8-15, k 8-15)
  const uint32_t b_lane = w_s + (uint32_t)((lane & 7) + (lane >> 4) * 8) * WPITCH + (uint32_t)((lane >> 3) & 1) * 16u;

  // the tile's activations go global -> shared with cp.async (no registers): the NEXT tile's loads are in flight while this one is
  // multiplied.  The SE gate is applied to the A fragments (a0, a1: k = 16 ks + 2t, +1; a2, a3: k + 8): 4 HMUL2 per K step.
  auto issue = [&](long long tile, int buf) {
    const long long px0 = tile * 16;
    const uint32_t dst = tile_s + (uint32_t)buf * 16u * PITCH;
    if (tile < te) {
#pragma unroll
      for (int i = 0; i < U; ++i) {
        const int idx = lane + 32 * i, px = idx / UPP, u = idx - px * UPP;
        const void* src = u < U0 ? (const void*)(p.h2 + (size_t)(px0 + px) * K0 + u * 8) : (const void*)(p.x + (size_t)(px0 + px) * K1 + (u - U0) * 8);
        cp_async16(dst + (uint32_t)px * PITCH + (uint32_t)u * 16u, src);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  uint32_t bfr[kBReg ? KS0 + KS1 : 1][kBReg ? NT / 2 : 1][4];
  if (kBReg) {
#pragma unroll
    for (int ks = 0; ks < KS0 + KS1; ++ks)
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) ldsm_x4(b_lane + (uint32_t)np * 16u * WPITCH + (uint32_t)ks * 32u, bfr[kBReg ? ks : 0][kBReg ? np : 0]);
  }
  float s1[NT][2], s2[NT][2];     // per-thread column sums of the stored values: columns 8 nt + 2 t + {0, 1}, rows g and g + 8
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) { s1[nt][0] = s1[nt][1] = s2[nt][0] = s2[nt][1] = 0.f; }
  uint32_t glo[KS0], ghi[KS0];    // gate pairs of this lane's fragment columns (current image)

  // per-warp flush at image boundaries: lanes -> one fp64 atomic per (column, moment); a few exact adds per entry
  auto flush_image = [&](int img) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float a = s1[nt][j], b = s2[nt][j];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
        if (g == 0) {
          double* d = p.stats + ((size_t)img * Nc + nt * 8 + 2 * t + j) * 2;
          atomicAdd(d, (double)a); atomicAdd(d + 1, (double)b);
        }
        s1[nt][j] = 0.f; s2[nt][j] = 0.f;
      }
  };

  long long tile = tb + warp;
  int buf = 0, cur_img = -1;
#pragma unroll
  for (int d = 0; d < NBUF - 1; ++d) issue(tile + (long long)d * kPsWarps, d);   // (commits an empty group beyond the range)
  for (; tile < te; tile += kPsWarps, buf = (buf + 1 == NBUF ? 0 : buf + 1)) {
    issue(tile + (long long)(NBUF - 1) * kPsWarps, buf == 0 ? NBUF - 1 : buf - 1);   // into the tile consumed last iteration
    asm volatile("cp.async.wait_group %0;" ::"n"(NBUF - 1) : "memory");
    __syncwarp();
    const uint32_t otile_s = tile_s + (uint32_t)buf * 16u * PITCH;
    const int img = (int)(tile / tiles_per_img);
    if (img != cur_img) {
      if (cur_img >= 0) flush_image(cur_img);
      const float2* gsrc = p.gate + (size_t)img * p.gate_ld + p.gate_off;
#pragma unroll
      for (int ks = 0; ks < KS0; ++ks) {
        const __half2 lo = __floats2half2_rn(gsrc[ks * 16 + 2 * t].x, gsrc[ks * 16 + 2 * t + 1].x);
        const __half2 hi = __floats2half2_rn(gsrc[ks * 16 + 2 * t + 8].x, gsrc[ks * 16 + 2 * t + 9].x);
        glo[ks] = *reinterpret_cast<const uint32_t*>(&lo);
        ghi[ks] = *reinterpret_cast<const uint32_t*>(&hi);
      }
      cur_img = img;
    }
    const long long px0 = tile * 16;
    const uint32_t a_buf = a_lane + (uint32_t)buf * 16u * PITCH;
    float acc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS0 + KS1; ++ks) {
      uint32_t a[4];
      ldsm_x4(a_buf + (uint32_t)ks * 32u, a);
      if (ks < KS0) {
        a[0] = hmul2_u(a[0], glo[ks < KS0 ? ks : 0]); a[1] = hmul2_u(a[1], glo[ks < KS0 ? ks : 0]);
        a[2] = hmul2_u(a[2], ghi[ks < KS0 ? ks : 0]); a[3] = hmul2_u(a[3], ghi[ks < KS0 ? ks : 0]);
      }
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        uint32_t b[4];
        if (kBReg) { b[0] = bfr[kBReg ? ks : 0][kBReg ? np : 0][0]; b[1] = bfr[kBReg ? ks : 0][kBReg ? np : 0][1]; b[2] = bfr[kBReg ? ks : 0][kBReg ? np : 0][2]; b[3] = bfr[kBReg ? ks : 0][kBReg ? np : 0][3]; }
        else ldsm_x4(b_lane + (uint32_t)np * 16u * WPITCH + (uint32_t)ks * 32u, b);
        if (ks < KS0) { mma_f16(acc[2 * np], a, b[0], b[1]); mma_f16(acc[2 * np + 1], a, b[2], b[3]); }
        else { mma_bf16(acc[2 * np], a, b[0], b[1]); mma_bf16(acc[2 * np + 1], a, b[2], b[3]); }
      }
    }
    // bf16, statistics of the stored values, transpose through shared memory (the activation tile is consumed: re-use it)
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const uint32_t p0 = pack_bf16(acc[nt][0], acc[nt][1]), p1 = pack_bf16(acc[nt][2], acc[nt][3]);
      const float r00 = bf16lo(p0), r01 = bf16hi(p0), r10 = bf16lo(p1), r11 = bf16hi(p1);
      s1[nt][0] += r00 + r10; s1[nt][1] += r01 + r11;
      s2[nt][0] = fmaf(r00, r00, fmaf(r10, r10, s2[nt][0])); s2[nt][1] = fmaf(r01, r01, fmaf(r11, r11, s2[nt][1]));
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(otile_s + (uint32_t)g * OPITCH + (uint32_t)(nt * 16 + t * 4)), "r"(p0) : "memory");
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(otile_s + (uint32_t)(g + 8) * OPITCH + (uint32_t)(nt * 16 + t * 4)), "r"(p1) : "memory");
    }
    __syncwarp();
    constexpr int OU = Nc / 8;                          // 16-byte units per output pixel
#pragma unroll
    for (int i = lane; i < 16 * OU; i += 32) {
      const int px = i / OU, uu = i - px * OU;
      const uint4 o = lds128(otile_s + (uint32_t)px * OPITCH + (uint32_t)uu * 16u);
      *reinterpret_cast<uint4*>(p.out + (size_t)(px0 + px) * Nc + uu * 8) = o;
    }
    __syncwarp();   // buffer `buf` is free again (the next iteration's loads go there)
  }
  if (cur_img >= 0) flush_image(cur_img);
}

}  // namespace

bool proj_stream_supported(int nseg, const int* segK, const int* seg_f16, const int* seg_mode, int Nc, int P) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("LCM_NO_PROJ_STREAM"); off = (e && atoi(e)) ? 1 : 0; }
  return !off && nseg == 2 && segK[0] == 128 && segK[1] == 32 && seg_f16[0] == 1 && seg_f16[1] == 0 && seg_mode[0] == XF_SCALE &&
         seg_mode[1] == XF_NONE && Nc == 32 && P % 16 == 0;
}

// g: seg 0 = h2 (fp16, XF_SCALE with the SE gate), seg 1 = the block input (bf16, raw); g.W: [Nc][160] row-major 16-bit weights
// (fp16 for the h2 columns, bf16 for the x columns; PackJob WL_ROWMAJOR); g.out bf16; g.stats [N][Nc][2] (+=).
int launch_proj_stream(const GemmParams& g, int num_sms, cudaStream_t st) {
  int segK[2] = {g.seg[0].K, g.seg[1].K}, f16[2] = {g.seg[0].f16, g.seg[1].f16}, mode[2] = {g.seg[0].mode, g.seg[1].mode};
  if (!proj_stream_supported(g.nseg, segK, f16, mode, g.Nc, g.P) || !g.stats || g.out_f16 || g.M % g.P || g.seg[0].ld != 128 || g.seg[1].ld != 32 ||
      !g.seg[0].coef)
    return -1;
  PsParams p{};
  p.h2 = reinterpret_cast<const __half*>(g.seg[0].A);
  p.x = reinterpret_cast<const bf16*>(g.seg[1].A);
  p.gate = g.seg[0].coef; p.gate_ld = g.seg[0].coef_ld; p.gate_off = g.seg[0].coef_off;
  p.W = reinterpret_cast<const uint16_t*>(g.W);
  p.out = reinterpret_cast<bf16*>(g.out);
  p.stats = g.stats;
  p.P = g.P;
  p.tiles = g.M / 16;
  constexpr int K = 160, Nc = 32;
  static int cfg = -1;   // LCM_PS_CFG: warps x tile ring (smem = warps x ring x 5.25 KB + 10.5 KB): 0 = 8 x 5, 1 = 8 x 5 with the weight fragments in
  // registers, 2 = 12 x 3, 3 = 16 x 2.  Measured at 64 x 256^2: 282 - 294 / 295 - 297 / 286 / 284 - 288 us — a plateau at ~5.6 TB/s that
  // neither the load depth, nor the warp count, nor the shared-memory traffic moves (general kernel: 342 - 354 us).
  if (cfg < 0) { const char* e = getenv("LCM_PS_CFG"); cfg = e ? atoi(e) : 0; if (cfg < 0 || cfg > 3) cfg = 0; }
  auto go = [&](auto kfn, int kW, int kB) -> int {
    const size_t smem = (size_t)Nc * (K * 2 + 16) + (size_t)kW * kB * 16 * (K * 2 + 16);
    if (ensure_dyn_smem_fn(kfn, smem)) return -2;
    long long grid = num_sms;
    const long long rounds = (p.tiles + kW - 1) / kW;
    if (grid > rounds) grid = rounds;
    if (grid < 1) grid = 1;
    launch_pdl(kfn, dim3((unsigned)grid), dim3(kW * 32), smem, st, p);
    return 0;
  };
  switch (cfg) {
    case 1: return go(proj_stream_kernel<8, 2, 4, 8, 5, true>, 8, 5);
    case 2: return go(proj_stream_kernel<8, 2, 4, 12, 3, false>, 12, 3);
    case 3: return go(proj_stream_kernel<8, 2, 4, 16, 2, false>, 16, 2);
    default: return go(proj_stream_kernel<8, 2, 4, 8, 5, false>, 8, 5);
  }
}

}  // namespace lcm
