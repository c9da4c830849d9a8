// Diagnostic: how fast can ONE CTA per SM stream global memory into shared memory with TMA, and does a second CTA per SM help?
// Producer thread issues loads into a ring of S stages (mbarrier full / empty); a consumer thread frees each stage as soon as it lands.
//   mode 0: cp.async.bulk 16 KB contiguous     mode 1: 2-D tensor box [64 x 128 rows] of a [M][128] bf16 matrix (128-byte pieces of 256-byte rows,
//   alternating halves — the project GEMM's pattern), 128-byte swizzle
// nvcc -arch=sm_100a -O3 tma_stream_bw.cu -o tma_stream_bw && ./tma_stream_bw
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint32_t b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(parity) : "memory");
}

template <int MODE>
__global__ void stream(const uint8_t* src, const __grid_constant__ CUtensorMap map, long long chunks_per_cta, int stages) {
  extern __shared__ __align__(1024) uint8_t sm[];
  const uint32_t base = (s32(sm) + 1023u) & ~1023u;
  const uint32_t bars = base + stages * 16384u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(bars + 8 * s, 1); mbar_init(bars + 8 * (16 + s), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long c0 = (long long)blockIdx.x * chunks_per_cta;
  if (threadIdx.x == 0) {
    int st = 0; uint32_t ph = 0;
    for (long long c = 0; c < chunks_per_cta; ++c) {
      mbar_wait(bars + 8 * (16 + st), ph ^ 1u);
      mbar_expect(bars + 8 * st, 16384u);
      const uint32_t dst = base + st * 16384u;
      if (MODE == 0) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src + (c0 + c) * 16384), "r"(16384), "r"(bars + 8 * st) : "memory");
      } else {
        const long long cc = c0 + c;
        const int row0 = (int)(cc >> 1) * 128, col0 = (int)(cc & 1) * 64;
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&map)), "r"(col0), "r"(row0), "r"(bars + 8 * st) : "memory");
      }
      if (++st == stages) { st = 0; ph ^= 1u; }
    }
  } else if (threadIdx.x == 32) {
    int st = 0; uint32_t ph = 0;
    for (long long c = 0; c < chunks_per_cta; ++c) {
      mbar_wait(bars + 8 * st, ph);
      mbar_arrive(bars + 8 * (16 + st));
      if (++st == stages) { st = 0; ph ^= 1u; }
    }
  }
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                          CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const size_t bytes = 4ull << 30;
  uint8_t* x; cudaMalloc(&x, bytes); cudaMemset(x, 1, bytes);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  CUtensorMap map;
  const cuuint64_t gdim[2] = {128, bytes / 256}; const cuuint64_t gstr[1] = {256}; const cuuint32_t box[2] = {64, 128}; const cuuint32_t es[2] = {1, 1};
  ((EncFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, x, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int mode = 0; mode < 2; ++mode)
    for (int per_sm : {1, 2, 4})
      for (int stages : {2, 4, 8, 12}) {
        const size_t smem = stages * 16384 + 2048;
        if (smem * per_sm > 226 * 1024) continue;
        const int grid = 148 * per_sm;
        const long long cpc = (long long)(bytes / 16384) / grid;
        if (mode == 0) cudaFuncSetAttribute(stream<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        else cudaFuncSetAttribute(stream<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        float best = 1e9f;
        for (int r = 0; r < 3; ++r) {
          cudaEventRecord(a);
          if (mode == 0) stream<0><<<grid, 64, smem>>>(x, map, cpc, stages); else stream<1><<<grid, 64, smem>>>(x, map, cpc, stages);
          cudaEventRecord(b); cudaEventSynchronize(b);
          float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
        }
        const cudaError_t e = cudaGetLastError();
        printf("%s  CTAs/SM %d  stages %2d: %7.0f GB/s  (%s)\n", mode ? "tensor 2-D [64x128] of 256-B rows" : "bulk 16 KB contiguous          ", per_sm, stages,
               (double)cpc * grid * 16384 / best / 1e6, cudaGetErrorString(e));
      }
  return 0;
}
