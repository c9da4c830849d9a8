// Input-side pass of the fused expand -> depthwise path (xdw_fused.cu; efficient_unet.py:207-212):
//
//   t[m][k]  = sat(a1_k/6 x[m][k] + b1_k/6)   bf16: GroupNorm1 + ReLU6 of the block input (1-2 concat parts), the A operand of the
//                                             expand MMAs, written ONCE so that the fused kernel needs no prologue warps
//   S[k]     = sum_m t[m][k],   G[k][k'] = sum_m t[m][k] t[m][k']      per image: the statistics of the expand output follow
//                                             from them (gemm_expand.cu: sum = W6 . S, sum^2 = W6^T G W6)
//
// The pass moves 2 x K x 2 bytes per pixel (K = 32 .. 96 against the 4K channels of the hidden tensor it saves) and is
// HBM-bound: plain vector loads / stores at full occupancy; the Gram matrix on mma.sync (m16n8k16, bf16, fp32 accumulate)
// with A = B = the transposed pixel tile (ldmatrix.trans from a per-team shared-memory tile), upper triangle only, the
// column sums as one extra n-tile against a register of ones.
//   * team = GROUP warps sharing one pixel tile (16 x KSW pixels per warp): every warp loads, transforms, stores its 16 x KSW
//     pixels, then multiplies the WHOLE team tile for its share of the m-tile rows — GROUP = 1 / 2 / 4 keeps the accumulators
//     of K = 32 / 64 / 96 within ~60 registers.
//   * persistent CTAs over contiguous tile ranges; a CTA is inside one image per round, its warps' fragments are reduced in
//     shared memory and added to the per-image scratch (fp64 atomics) at image boundaries only.
#include <cuda_fp16.h>

#include "kernels.h"
#include "tc_common.cuh"

namespace lcm {

namespace {

using namespace tc;

constexpr int kXsWarps = 8;
constexpr int kGramLdS = 128;   // scratch layout of gemm_expand.cu: G[img][128][128], S[img][128]

struct XsParams {
  const bf16* x[2];
  const float2* coef[2];
  int coef_ld[2], coef_off[2], segK[2];
  int nseg;
  bf16* t;
  double* gram;
  double* colsum;
  int P;            // pixels per image
  long long tiles;  // team tiles in total
  int prefetch;     // pull the warp's next pixel tile towards L2 while the Gram update of this one runs
};

__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// number of accumulator tiles of m-tile row r (n-tiles j0 >= 16 r, 8 wide) plus the ones column
template <int KT> __host__ __device__ constexpr int row_tiles(int r) { return 2 * (KT - r) + 1; }
template <int KT, uint32_t MASK> __host__ __device__ constexpr int mask_tiles() {
  int n = 0;
  for (int r = 0; r < KT; ++r) if (MASK & (1u << r)) n += row_tiles<KT>(r);
  return n;
}
template <uint32_t MASK> __host__ __device__ constexpr int mask_min() {
  for (int r = 0; r < 32; ++r) if (MASK & (1u << r)) return r;
  return 0;
}

// One warp's share of the Gram update: m-tile rows in MASK, all k-steps of the team tile.
template <int KT, uint32_t MASK, int KSTEPS>
__device__ __forceinline__ void gram_update(float (&acc)[mask_tiles<KT, MASK>()][4], uint32_t tile, uint32_t pitch, int lane) {
  constexpr int R0 = mask_min<MASK>();
  // ldmatrix.x4.trans row address of this lane: matrix q = lane >> 3 -> (k half = q >> 1, channel half = q & 1), row = lane & 7
  const uint32_t lrow = (uint32_t)(((lane >> 4) & 1) * 8 + (lane & 7)) * pitch + (uint32_t)((lane >> 3) & 1) * 16u;
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks) {
    uint32_t fr[KT - R0][4];
#pragma unroll
    for (int r = R0; r < KT; ++r) ldsm_x4_t(tile + (uint32_t)ks * 16u * pitch + lrow + (uint32_t)r * 32u, fr[r - R0]);
    int ti = 0;
#pragma unroll
    for (int r = R0; r < KT; ++r) {
      if (!(MASK & (1u << r))) continue;
      const uint32_t* a = fr[r - R0];
#pragma unroll
      for (int r2 = r; r2 < KT; ++r2) {
        const uint32_t* b = fr[r2 - R0];
        mma_bf16_16816(acc[ti++], a[0], a[1], a[2], a[3], b[0], b[2]);     // n-tile 16 r2
        mma_bf16_16816(acc[ti++], a[0], a[1], a[2], a[3], b[1], b[3]);     // n-tile 16 r2 + 8
      }
      mma_bf16_16816(acc[ti++], a[0], a[1], a[2], a[3], 0x3f803f80u, 0x3f803f80u);   // column sums: B = ones
    }
  }
}

// fragments -> the CTA's image in shared memory (upper-triangle tiles mirrored).  The image is fp64: a few thousand fp32 partial
// sums add exactly in fp64, so the result does not depend on the order in which the warps arrive (bitwise reproducible runs)
template <int KT, uint32_t MASK>
__device__ __forceinline__ void gram_flush(float (&acc)[mask_tiles<KT, MASK>()][4], double* gs, double* ss, int lane) {
  constexpr int K = KT * 16, LD = K + 1;
  const int g = lane >> 2, t = lane & 3;
  int ti = 0;
#pragma unroll
  for (int r = 0; r < KT; ++r) {
    if (!(MASK & (1u << r))) continue;
#pragma unroll
    for (int j = 2 * r; j < 2 * KT; ++j) {
      float (&c)[4] = acc[ti++];
      const int i0 = 16 * r + g, j0 = 8 * j + 2 * t;
      atomicAdd(&gs[i0 * LD + j0], (double)c[0]); atomicAdd(&gs[i0 * LD + j0 + 1], (double)c[1]);
      atomicAdd(&gs[(i0 + 8) * LD + j0], (double)c[2]); atomicAdd(&gs[(i0 + 8) * LD + j0 + 1], (double)c[3]);
      if (j >= 2 * r + 2) {   // strictly above the diagonal block: mirror
        atomicAdd(&gs[j0 * LD + i0], (double)c[0]); atomicAdd(&gs[(j0 + 1) * LD + i0], (double)c[1]);
        atomicAdd(&gs[j0 * LD + i0 + 8], (double)c[2]); atomicAdd(&gs[(j0 + 1) * LD + i0 + 8], (double)c[3]);
      }
      c[0] = c[1] = c[2] = c[3] = 0.f;
    }
    float (&c)[4] = acc[ti++];
    if (t == 0) { atomicAdd(&ss[16 * r + g], (double)c[0]); atomicAdd(&ss[16 * r + g + 8], (double)c[2]); }
    c[0] = c[1] = c[2] = c[3] = 0.f;
  }
}

template <int KT, int GROUP> struct Masks;
template <int KT> struct Masks<KT, 1> { static constexpr uint32_t m[4] = {(1u << KT) - 1u, 0, 0, 0}; };
template <> struct Masks<4, 2> { static constexpr uint32_t m[4] = {0x9, 0x6, 0, 0}; };              // 9 + 3 | 7 + 5 tiles
template <> struct Masks<5, 4> { static constexpr uint32_t m[4] = {0x1, 0x2, 0x4, 0x18}; };
template <> struct Masks<6, 4> { static constexpr uint32_t m[4] = {0x1, 0x2, 0x24, 0x18}; };         // 13 | 11 | 9 + 3 | 7 + 5

// the main loop of one warp; MASK = this warp's m-tile rows of the Gram matrix (compile time: its accumulators stay in registers)
template <int KT, int GROUP, int KSW, uint32_t MASK>
__device__ __forceinline__ void xstats_body(const XsParams& p, double* gs, double* ss, uint32_t tile_s, int team, int gw) {
  constexpr int K = KT * 16, UPP = K / 8;                  // channels, 16-byte units per pixel
  constexpr int WPX = 16 * KSW, TP = WPX * GROUP;          // pixels per warp / per team tile
  // a lane owns ONE unit column (its 8 GroupNorm coefficients live in registers): a pass covers PPP pixels with PPP * UPP lanes
  constexpr int PPP = UPP <= 2 ? 16 : UPP <= 4 ? 8 : UPP <= 8 ? 4 : UPP <= 16 ? 2 : 1;
  constexpr int U = WPX / PPP;                             // passes (units per lane) per iteration
  constexpr int TEAMS = kXsWarps / GROUP;
  constexpr uint32_t PITCH = K * 2 + 16;
  constexpr int LD = K + 1;
  constexpr int NT = mask_tiles<KT, MASK>();
  const int tid = threadIdx.x, lane = tid & 31;
  const int upp0 = p.segK[0] / 8;
  const bool lane_on = lane < PPP * UPP;
  const int lpx = lane / UPP, lu = lane - lpx * UPP;       // this lane's pixel within a pass and its unit column
  const int lseg = lu < upp0 ? 0 : 1, lch = (lseg ? lu - upp0 : lu) * 8;   // segment and first channel inside it
  const bf16* lsrc = p.x[lseg] + lch;
  const int lK = p.segK[lseg];
  float2 cf[8];
  const int tiles_per_img = p.P / TP;
  const long long tb = p.tiles * blockIdx.x / gridDim.x, te = p.tiles * (blockIdx.x + 1) / gridDim.x;

  float acc[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

  int cur_img = -1;
  // every warp of the CTA gets here in the same round (base / img are CTA-uniform), from its own instantiation of this body
  auto flush_image = [&](int img) {
    gram_flush<KT, MASK>(acc, gs, ss, lane);
    bar_sync(0, kXsWarps * 32);
    double* g = p.gram + (size_t)img * kGramLdS * kGramLdS;
    for (int i = tid; i < K * K; i += kXsWarps * 32) {
      const int r = i / K, c = i - r * K;
      atomicAdd(g + r * kGramLdS + c, gs[r * LD + c]);
      gs[r * LD + c] = 0.0;
    }
    for (int i = tid; i < K; i += kXsWarps * 32) { atomicAdd(p.colsum + (size_t)img * kGramLdS + i, ss[i]); ss[i] = 0.0; }
    bar_sync(0, kXsWarps * 32);
  };

  // loop invariants of the L2 prefetch (kernel parameters indexed by lane: read once, not per iteration)
  const bf16* pf_src = lane ? p.x[1] : p.x[0];
  const int pf_sk = lane ? p.segK[1] : p.segK[0];
  const bool pf_on = p.prefetch && lane < p.nseg;
  int img = (int)(tb / tiles_per_img);                         // image of `base`, tracked incrementally (no division per round)
  long long img_end = (long long)(img + 1) * tiles_per_img;
  for (long long base = tb; base < te;) {
    if (base >= img_end) { ++img; img_end += tiles_per_img; }   // a round never crosses an image boundary
    long long end = base + TEAMS;
    if (end > img_end) end = img_end;
    if (end > te) end = te;
    if (img != cur_img) {
      if (cur_img >= 0) flush_image(cur_img); else bar_sync(0, kXsWarps * 32);   // (first image: the zeroed shared image is visible)
      if (lane_on) {
        const float4* src = reinterpret_cast<const float4*>(p.coef[lseg] + (size_t)img * p.coef_ld[lseg] + p.coef_off[lseg] + lch);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 cc = src[j];
          cf[2 * j] = make_float2(cc.x * (1.f / 6.f), cc.y * (1.f / 6.f));
          cf[2 * j + 1] = make_float2(cc.z * (1.f / 6.f), cc.w * (1.f / 6.f));
        }
      }
      cur_img = img;
    }
    const long long tile = base + team;
    const bool active = tile < end;
    if (active) {
      const long long px0 = tile * TP + (long long)gw * WPX;       // first pixel of this warp's part of the team tile
      uint4 v[U];
      if (lane_on) {
#pragma unroll
        for (int i = 0; i < U; ++i) v[i] = *reinterpret_cast<const uint4*>(lsrc + (size_t)(px0 + i * PPP + lpx) * lK);
      }
      // The loads of a warp are in flight for only part of an iteration (transform, stores, team barriers and the Gram MMAs
      // follow), so the DRAM latency of its NEXT tile is started now: one bulk L2 prefetch per segment (contiguous WPX pixels)
      const long long nt = end + team + (long long)(p.prefetch - 1) * TEAMS;
      if (pf_on && nt < te) {
        const bf16* nx = pf_src + (size_t)(nt * TP + (long long)gw * WPX) * pf_sk;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nx), "r"((uint32_t)(WPX * 2) * (uint32_t)pf_sk) : "memory");
      }
      if (lane_on) {
#pragma unroll
        for (int i = 0; i < U; ++i) {
          const int px = i * PPP + lpx;
          float f[8];
          unpack8(v[i], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = __saturatef(fmaf(cf[j].x, f[j], cf[j].y));
          const uint4 o = pack8(f);
          *reinterpret_cast<uint4*>(p.t + (size_t)(px0 + px) * K + lu * 8) = o;
          sts128(tile_s + (uint32_t)(gw * WPX + px) * PITCH + (uint32_t)lu * 16u, o);
        }
      }
    }
    if (GROUP == 1) __syncwarp(); else bar_sync(1 + team, GROUP * 32);
    if (active) gram_update<KT, MASK, KSW * GROUP>(acc, tile_s, PITCH, lane);
    if (GROUP == 1) __syncwarp(); else bar_sync(1 + team, GROUP * 32);   // the tile is free again
    base = end;
  }
  if (cur_img >= 0) flush_image(cur_img);
}

template <int KT, int GROUP, int KSW>
__global__ void __launch_bounds__(kXsWarps * 32, KT <= 1 ? 3 : 2) xstats_kernel(const XsParams p) {
  constexpr int K = KT * 16, LD = K + 1, TP = 16 * KSW * GROUP;
  constexpr uint32_t PITCH = K * 2 + 16;
  extern __shared__ __align__(16) uint8_t xs_raw[];   // (K (K + 2) * 8 is a multiple of 16)
  double* gs = reinterpret_cast<double*>(xs_raw);          // [K][K + 1]
  double* ss = gs + K * LD;                                // [K]
  uint8_t* tiles = reinterpret_cast<uint8_t*>(ss + K);     // [TEAMS][TP][PITCH], 16-byte aligned (K is even)
  const int tid = threadIdx.x, warp = tid >> 5;
  const int team = warp / GROUP, gw = warp % GROUP;
  const uint32_t tile_s = smem_u32(tiles) + (uint32_t)team * TP * PITCH;
  for (int i = tid; i < K * LD + K; i += kXsWarps * 32) gs[i] = 0.0;
  pdl_wait();
  pdl_trigger();
  using M = Masks<KT, GROUP>;
  if constexpr (GROUP == 1) {
    xstats_body<KT, GROUP, KSW, M::m[0]>(p, gs, ss, tile_s, team, gw);
  } else if constexpr (GROUP == 2) {
    if (gw == 0) xstats_body<KT, GROUP, KSW, M::m[0]>(p, gs, ss, tile_s, team, gw);
    else xstats_body<KT, GROUP, KSW, M::m[1]>(p, gs, ss, tile_s, team, gw);
  } else {
    if (gw == 0) xstats_body<KT, GROUP, KSW, M::m[0]>(p, gs, ss, tile_s, team, gw);
    else if (gw == 1) xstats_body<KT, GROUP, KSW, M::m[1]>(p, gs, ss, tile_s, team, gw);
    else if (gw == 2) xstats_body<KT, GROUP, KSW, M::m[2]>(p, gs, ss, tile_s, team, gw);
    else xstats_body<KT, GROUP, KSW, M::m[3]>(p, gs, ss, tile_s, team, gw);
  }
}

template <int KT, int GROUP, int KSW>
int launch_one(const XsParams& p0, long long M, int num_sms, cudaStream_t st) {
  constexpr int K = KT * 16, TP = 16 * KSW * GROUP, TEAMS = kXsWarps / GROUP;
  XsParams p = p0;
  static int pf = -1;
  if (pf < 0) { const char* e = getenv("LCM_XS_PREFETCH"); pf = e ? atoi(e) : 1; }
  p.prefetch = pf;
  if (p.P % TP) return -1;
  p.tiles = M / TP;
  const size_t smem = (size_t)(K * (K + 1) + K) * 8 + (size_t)TEAMS * TP * (K * 2 + 16);
  auto kfn = xstats_kernel<KT, GROUP, KSW>;
  if (ensure_dyn_smem_fn(kfn, smem)) return -2;
  int* occ = device_cache_slot((const void*)kfn);
  if (*occ == 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kfn, kXsWarps * 32, smem) != cudaSuccess || n < 1) n = 1;
    *occ = n;
  }
  long long grid = (long long)num_sms * *occ;
  const long long rounds = (p.tiles + TEAMS - 1) / TEAMS;
  if (grid > rounds) grid = rounds;
  if (grid < 1) grid = 1;
  launch_pdl(kfn, dim3((unsigned)grid), dim3(kXsWarps * 32), smem, st, p);
  return 0;
}

}  // namespace

bool xstats_supported(int Ktot, int P) {
  return (Ktot == 16 || Ktot == 32 || Ktot == 48 || Ktot == 64 || Ktot == 80 || Ktot == 96) && P % 128 == 0;
}

// segs of g: the block input parts (bf16, XF_AFFINE_RELU6 with the GroupNorm1 coefficients, K_s % 16 == 0); t: [M][Ktot] bf16;
// scratch: gemm_expand_scratch_bytes(images), zero on entry (+=).  The caller finishes with launch_expand_stats_finalize.
int launch_xstats(const GemmParams& g, void* t, void* scratch, int num_sms, cudaStream_t st) {
  if (g.nseg < 1 || g.nseg > 2 || g.M % g.P || !scratch || !t) return -1;
  XsParams p{};
  int Kt = 0;
  for (int s = 0; s < g.nseg; ++s) {
    if (g.seg[s].mode != XF_AFFINE_RELU6 || g.seg[s].f16 || !g.seg[s].coef || g.seg[s].ld != g.seg[s].K || g.seg[s].K % 16) return -1;
    p.x[s] = reinterpret_cast<const bf16*>(g.seg[s].A); p.coef[s] = g.seg[s].coef;
    p.coef_ld[s] = g.seg[s].coef_ld; p.coef_off[s] = g.seg[s].coef_off; p.segK[s] = g.seg[s].K;
    Kt += g.seg[s].K;
  }
  if (g.nseg == 1) { p.x[1] = p.x[0]; p.segK[1] = 0; }
  if (!xstats_supported(Kt, g.P)) return -1;
  const int images = (int)(g.M / g.P);
  p.nseg = g.nseg;
  p.t = reinterpret_cast<bf16*>(t);
  p.gram = reinterpret_cast<double*>(scratch);
  p.colsum = p.gram + (size_t)images * kGramLdS * kGramLdS;
  p.P = g.P;
  switch (Kt / 16) {
    case 1: return launch_one<1, 1, 4>(p, g.M, num_sms, st);   // <KT, GROUP, KSW>: 16 KSW pixels per warp and iteration
    case 2: return launch_one<2, 1, 2>(p, g.M, num_sms, st);
    case 3: return launch_one<3, 1, 2>(p, g.M, num_sms, st);
    case 4: return launch_one<4, 2, 2>(p, g.M, num_sms, st);
    case 5: return launch_one<5, 4, 1>(p, g.M, num_sms, st);
    case 6: return launch_one<6, 4, 1>(p, g.M, num_sms, st);
  }
  return -1;
}

}  // namespace lcm
