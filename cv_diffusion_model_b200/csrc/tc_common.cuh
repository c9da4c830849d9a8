// PTX wrappers and small device helpers shared by the tcgen05 kernels.
#pragma once
#include <cuda_fp16.h>

#include <cstdio>

#include "common.cuh"

namespace lcm {
namespace tc {

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// try_wait suspends the thread until the phase completes or a (short, implementation-defined) time limit expires.
// Measured on B200 (round 1): with a suspend-time HINT ptxas emits TRYWAIT + NANOSLEEP.SYNCS <hint>, and a thread that
// went to sleep wakes up several hundred cycles after the arrive — fine for a producer that waits for a free slot,
// ruinous on a dependent chain (the expand kernel's MMA warp spent ~600 cycles per hop).  So: plain try_wait
// (spinning, ~60 cycles wake-up) on latency-critical waits, the hinted form only where nobody waits for the waiter.
constexpr uint32_t kSuspendHintNs = 20000;
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_try_relaxed(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(kSuspendHintNs)
      : "memory");
  return ok != 0;
}
// non-blocking probe (try_wait may suspend the thread for a hardware time slice; test_wait never does)
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped kernel, never as a hung GPU.
static __device__ __noinline__ void mbar_timeout(uint32_t bar, uint32_t parity) {
  printf("gemm_tc: mbarrier timeout block %d thread %d bar %u parity %u\n", blockIdx.x, threadIdx.x, bar, parity);
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) mbar_timeout(bar, parity);   // ~2 s: a protocol bug, not a slow producer
  }
}
// for waiters nobody waits for (producers waiting for a free slot): sleeps instead of spinning
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_relaxed(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) mbar_timeout(bar, parity);
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major, 128-byte swizzle, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor, version 1 = sm_100)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 ldg_cached(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ void unpack8(uint4 u, float (&f)[8]) {
  f[0] = bf16lo(u.x); f[1] = bf16hi(u.x); f[2] = bf16lo(u.y); f[3] = bf16hi(u.y);
  f[4] = bf16lo(u.z); f[5] = bf16hi(u.z); f[6] = bf16lo(u.w); f[7] = bf16hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

// fp16 flavours (hidden tensors of the inverted-residual block)
__device__ __forceinline__ void unpack8h(uint4 u, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) { const float2 v = __half22float2(h[j]); f[2 * j] = v.x; f[2 * j + 1] = v.y; }
}
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
  // satfinite: an out-of-range value becomes +-65504 instead of inf (GroupNorm statistics stay finite)
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint4 pack8h(const float (&f)[8]) {
  return make_uint4(pack_f16(f[0], f[1]), pack_f16(f[2], f[3]), pack_f16(f[4], f[5]), pack_f16(f[6], f[7]));
}

// 16 bytes of activations through the prologue: coef = 8 x (a, b).  XF_AFFINE and XF_AFFINE_RELU6 share one
// instruction stream (clamp bounds are +-inf for the plain affine) to keep the code small; SiLU is only
// used by the final conv, which has its own kernel.
// On the tcgen05 path relu6(a x + b) is computed as sat(a/6 x + b/6) in [0, 1] (one FFMA.SAT) and the factor 6 lives in
// the packed weights of that segment (PackJob::scale).  `prescaled`: the coefficients already carry the 1/6.
__device__ __forceinline__ uint4 apply_xform(uint4 raw, const float2* ab, int mode, bool f16 = false, bool prescaled = false) {
  float f[8];
  if (f16) unpack8h(raw, f); else unpack8(raw, f);
  const float4* c4 = reinterpret_cast<const float4*>(ab);
  if (mode == XF_AFFINE_RELU6) {
    const float k = prescaled ? 1.f : (1.f / 6.f);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 c = c4[j];
      f[2 * j] = __saturatef(fmaf(c.x * k, f[2 * j], c.y * k));
      f[2 * j + 1] = __saturatef(fmaf(c.z * k, f[2 * j + 1], c.w * k));
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 c = c4[j];
      f[2 * j] = fmaf(c.x, f[2 * j], c.y);
      f[2 * j + 1] = fmaf(c.z, f[2 * j + 1], c.w);
    }
  }
  return f16 ? pack8h(f) : pack8(f);
}

__device__ __forceinline__ int div_upr(int u, int upr) {
  return upr == 8 ? (u >> 3) : (upr == 4 ? (u >> 2) : (upr == 2 ? (u >> 1) : u / upr));
}

// =====================================================================================================
// Loop bookkeeping is strictly incremental and 32-bit: a 64-bit divide or modulo costs hundreds of
// dependent cycles and every role is a single warp (or thread) on the critical path of the pipeline.
struct TileIter {
  int n_tile, m_tile, m0, img, rem;   // m0 = m_tile*128 ; img = m0 / P ; rem = m0 % P
  int step;                           // tiles per next(): 2 when a CTA pair walks the tile list in lockstep
  __device__ __forceinline__ void init(long long t, int m_tiles, int P, int step_ = 1) {
    step = step_;
    n_tile = (int)(t / m_tiles);
    m_tile = (int)(t - (long long)n_tile * m_tiles);
    m0 = m_tile * 128;
    img = m0 / P;
    rem = m0 - img * P;
  }
  __device__ __forceinline__ void next(int m_tiles, int P) {
    m_tile += step;
    if (m_tile >= m_tiles) {   // step == 2 requires an even m_tiles: the pair stays inside one n row
      m_tile -= m_tiles; ++n_tile; m0 = m_tile * 128; img = m0 / P; rem = m0 - img * P;
      return;
    }
    m0 += 128 * step;
    rem += 128 * step;
    while (rem >= P) { rem -= P; ++img; }
  }
};
struct Ring {
  int stage; uint32_t phase; int stages;
  __device__ __forceinline__ void advance() { if (++stage == stages) { stage = 0; phase ^= 1u; } }
};


}  // namespace tc
}  // namespace lcm
