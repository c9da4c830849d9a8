// Diagnostic: what is the HBM write ceiling on this B200?  (plain / streaming / bulk-TMA stores, memset, 1:4 mix)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__global__ void fill_v4(uint4* p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = make_uint4(1, 2, 3, 4);
}
__global__ void fill_cs(uint4* p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p + i), "r"(1), "r"(2), "r"(3), "r"(4) : "memory");
}
// each CTA owns a contiguous slab: consecutive 16 KB bulk stores from shared memory
__global__ void fill_bulk(uint8_t* p, size_t bytes_per_cta, int chunk) {
  extern __shared__ __align__(128) uint8_t sm[];
  for (int i = threadIdx.x; i < chunk / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 7;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    uint8_t* dst = p + (size_t)blockIdx.x * bytes_per_cta;
    uint32_t s = (uint32_t)__cvta_generic_to_shared(sm);
    for (size_t o = 0; o < bytes_per_cta; o += chunk) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + o), "r"(s), "r"(chunk) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}
// 1 read : 4 write (the expand GEMM's traffic shape)
__global__ void mix14(const uint4* in, uint4* out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint4 v = in[i];
    size_t r = i / 4, c = i % 4;   // in row = 64 B, out row = 256 B
    uint4* o = out + r * 16 + c * 4;
    o[0] = v; o[1] = v; o[2] = v; o[3] = v;
  }
}
// 1 read : 1 write copy
__global__ void copy_v4(const uint4* in, uint4* out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}

// read-only stream (sum into one word per thread) and the project GEMM's traffic shape, 5 reads : 1 write
__global__ void read_v4(const uint4* in, uint32_t* out, size_t n) {
  uint32_t acc = 0;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t st = (size_t)gridDim.x * blockDim.x;
  for (; i + 3 * st < n; i += 4 * st) {
    const uint4 a = in[i], b = in[i + st], c = in[i + 2 * st], d = in[i + 3 * st];
    acc += a.x ^ b.y ^ c.z ^ d.w;
  }
  for (; i < n; i += st) acc += in[i].x;
  if (acc == 0x12345678u) out[0] = acc;
}
__global__ void mix51(const uint4* in, uint4* out, size_t nout) {   // out row i <- 5 input vectors (in has 5 * nout)
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t st = (size_t)gridDim.x * blockDim.x;
  for (; i < nout; i += st) {
    const uint4 a = in[i], b = in[i + nout], c = in[i + 2 * nout], d = in[i + 3 * nout], e = in[i + 4 * nout];
    out[i] = make_uint4(a.x ^ b.x, c.y ^ d.y, e.z, a.w);
  }
}

template <typename F> float timeit(F f, int n = 10) {
  f(); cudaDeviceSynchronize();
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  for (int i = 0; i < n; ++i) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / n;
}

int main() {
  const size_t bytes = 2ull << 30;
  uint8_t *x, *y;
  cudaMalloc(&x, bytes); cudaMalloc(&y, bytes);
  cudaMemset(x, 1, bytes); cudaMemset(y, 1, bytes);
  const size_t n = bytes / 16;
  float ms;
  ms = timeit([&] { cudaMemsetAsync(x, 0, bytes); }); printf("cudaMemset        %7.0f GB/s\n", bytes / ms / 1e6);
  for (int g : {148 * 4, 148 * 8, 148 * 32}) {
    ms = timeit([&] { fill_v4<<<g, 512>>>((uint4*)x, n); }); printf("st.v4   grid %5d %7.0f GB/s\n", g, bytes / ms / 1e6);
    ms = timeit([&] { fill_cs<<<g, 512>>>((uint4*)x, n); }); printf("st.cs   grid %5d %7.0f GB/s\n", g, bytes / ms / 1e6);
  }
  for (int chunk : {16384, 65536}) {
    cudaFuncSetAttribute(fill_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, chunk);
    for (int g : {148, 296, 592}) {
      size_t per = bytes / g / chunk * chunk;
      ms = timeit([&] { fill_bulk<<<g, 128, chunk>>>(x, per, chunk); });
      printf("bulk %3dK grid %3d %7.0f GB/s\n", chunk >> 10, g, per * g / ms / 1e6);
    }
  }
  ms = timeit([&] { copy_v4<<<148 * 16, 512>>>((const uint4*)x, (uint4*)y, n); }); printf("copy 1:1          %7.0f GB/s (r+w)\n", 2.0 * bytes / ms / 1e6);
  ms = timeit([&] { mix14<<<148 * 16, 512>>>((const uint4*)x, (uint4*)y, n / 4); }); printf("mix 1r:4w         %7.0f GB/s (r+w)\n", (bytes / 4.0 * 5) / ms / 1e6);
  uint32_t* flag; cudaMalloc(&flag, 4);
  for (int g : {148 * 8, 148 * 16, 148 * 32}) {
    ms = timeit([&] { read_v4<<<g, 512>>>((const uint4*)x, flag, n); }); printf("read-only grid %5d %7.0f GB/s\n", g, bytes / ms / 1e6);
  }
  { const size_t nout = n / 5;
    ms = timeit([&] { mix51<<<148 * 16, 512>>>((const uint4*)x, (uint4*)y, nout); }); printf("mix 5r:1w         %7.0f GB/s (r+w)\n", nout * 16.0 * 6 / ms / 1e6); }
  return 0;
}
