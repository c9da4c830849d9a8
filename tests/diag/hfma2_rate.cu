// Diagnostic: issue rate of HFMA2 / FFMA / HADD2 per SM sub-partition (warp-instructions per cycle) on this part.
// One CTA per SM, W warps, 16 independent dependency chains per thread.  nvcc -arch=sm_100a -O3 hfma2_rate.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>

template <int MODE>
__global__ void rate(float* out, long long* cyc, int iters) {
  __half2 a[16];
  float f[16];
  const __half2 m = __floats2half2_rn(1.0001f + threadIdx.x * 1e-6f, 0.9999f), c = __floats2half2_rn(1e-3f, -1e-3f);
  const float fm = 1.0001f + threadIdx.x * 1e-6f, fc = 1e-3f;
#pragma unroll
  for (int i = 0; i < 16; ++i) { a[i] = __floats2half2_rn(i * 0.1f, threadIdx.x * 0.01f); f[i] = i * 0.1f + threadIdx.x; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) a[i] = __hfma2(a[i], m, c);
      if (MODE == 1) f[i] = fmaf(f[i], fm, fc);
      if (MODE == 2) a[i] = __hadd2(a[i], c);
      if (MODE == 3) { a[i] = __hfma2(a[i], m, a[(i + 1) & 15]); }   // three distinct register operands
      if (MODE == 4) { if (i & 1) a[i] = __hfma2(a[i], m, c); else f[i] = fmaf(f[i], fm, fc); }          // 1 : 1 mix
      if (MODE == 5) { if ((i & 3) == 3) f[i] = fmaf(f[i], fm, fc); else a[i] = __hfma2(a[i], m, c); }   // 3 HFMA2 : 1 FFMA
      if (MODE == 6) { if (i & 1) a[i] = __hfma2(a[i], m, c); else f[i] = __int_as_float(__float_as_int(f[i]) * 3 + i); }   // HFMA2 : IMAD
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += __low2float(a[i]) + __high2float(a[i]) + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 4096;
  rate<MODE><<<148, warps * 32>>>(out, cyc, iters);
  rate<MODE><<<148, warps * 32>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double per_smsp = (double)iters * 16 * warps / 4.0 / (double)h[0];
  printf("%-16s warps/SM %2d: %.3f warp-instr / cycle / SMSP\n", name, warps, per_smsp);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {4, 8, 16, 32}) {
    run<0>("HFMA2", w);
    run<3>("HFMA2.3r", w);
    run<1>("FFMA", w);
    run<2>("HADD2", w);
    run<4>("HFMA2+FFMA 1:1", w);
    run<5>("HFMA2+FFMA 3:1", w);
    run<6>("HFMA2+IMAD 1:1", w);
  }
  return 0;
}
