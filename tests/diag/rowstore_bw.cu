// Diagnostic: bandwidth of "one thread = one 128-byte row segment" stores (the TMEM epilogue pattern: lane = row),
// with 16-byte and 32-byte vector stores, for row pitches of 256 B and 768 B.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

template <int VEC>   // 16 or 32 bytes per store
__global__ void rowstore(uint8_t* out, size_t rows, int pitch, int nblocks) {
  // thread t owns row (blockIdx.x * blockDim.x + t); writes nblocks segments of 128 B at columns j*128
  for (size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x; r < rows; r += (size_t)gridDim.x * blockDim.x) {
    uint8_t* p = out + r * pitch;
    for (int j = 0; j < nblocks; ++j) {
      if (VEC == 16) {
#pragma unroll
        for (int u = 0; u < 8; ++u) asm volatile("st.global.v4.u32 [%0], {%1,%1,%1,%1};" ::"l"(p + j * 128 + u * 16), "r"(u) : "memory");
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) asm volatile("st.global.v8.u32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p + j * 128 + u * 32), "r"(u) : "memory");
      }
    }
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
  uint8_t* x; cudaMalloc(&x, bytes); cudaMemset(x, 0, bytes);
  for (int pitch : {256, 768}) {
    const size_t rows = bytes / pitch; const int nb = pitch / 128;
    for (int g : {148 * 2, 148 * 8}) {
      float ms = timeit([&] { rowstore<16><<<g, 256>>>(x, rows, pitch, nb); });
      printf("pitch %4d grid %5d  v4 (16 B) %7.0f GB/s\n", pitch, g, rows * (double)pitch / ms / 1e6);
      ms = timeit([&] { rowstore<32><<<g, 256>>>(x, rows, pitch, nb); });
      printf("pitch %4d grid %5d  v8 (32 B) %7.0f GB/s\n", pitch, g, rows * (double)pitch / ms / 1e6);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
