// Shared device/host helpers for the sm_100a LCM-UNet kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace lcm {

typedef __nv_bfloat16 bf16;

// ---- programmatic dependent launch ---------------------------------------------------------------
// A forward is ~190 short dependent launches; with plain stream order every boundary pays the launch latency, the
// block scheduling and the next kernel's prologue (barrier init, TMEM allocation, tensor-map fetch) after the previous
// grid has drained.  Every kernel of the path therefore (1) calls pdl_trigger() once its own prologue is done, which lets
// the NEXT grid's blocks become resident as this grid's blocks retire, and (2) calls pdl_wait() before its first access
// to global memory: it returns when the previous grid has completed and its writes are visible.  Both are no-ops for a
// launch without the attribute, so the same kernels run under ops_api / the SIMT plan unchanged.
// LCM_NO_PDL=1 launches everything with plain stream order (A/B timing, debugging).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();   // plan.cu

// Function attributes and occupancy are per DEVICE: a process that builds plans on several GPUs must opt every kernel in
// on each of them.  ensure_dyn_smem remembers (current device, function) -> bytes opted in so far and raises the limit
// when needed; device_cache_slot returns a zero-initialised int slot per (current device, key) for occupancy caches.
// Both are thread-safe (plan.cu).
int ensure_dyn_smem(const void* fn, size_t bytes);   // 0 = ok
int* device_cache_slot(const void* key, int sub = 0);
template <class F> inline int ensure_dyn_smem_fn(F fn, size_t bytes) { return ensure_dyn_smem((const void*)fn, bytes); }

template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(static_cast<Args&&>(args))...);
}

// the same with a (cluster_x, 1, 1) thread-block cluster; max_clusters (optional) receives how many such clusters the
// device can hold at once (persistent kernels size their grid with it: a GPC with an odd SM count leaves one SM out)
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x,
                                      int* max_clusters, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster_x;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 2;
  if (max_clusters) return cudaOccupancyMaxActiveClusters(max_clusters, (const void*)kernel, &cfg);
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(static_cast<Args&&>(args))...);
}

// ---- prologue transform applied to an activation element as it is read -------------------------
enum XformMode : int {
  XF_NONE = 0,         // y = x
  XF_AFFINE = 1,       // y = a*x + b                      (GroupNorm, or SE gate with b = 0)
  XF_AFFINE_RELU6 = 2, // y = min(max(a*x + b, 0), 6)      (GroupNorm [+FiLM] then ReLU6)
  XF_AFFINE_SILU = 3,  // y = silu(a*x + b)                (final_norm + SiLU)
  XF_SCALE = 4         // y = a*x                          (SE gate; linear, so the tensor-core GEMM may fold it
                       //                                   into the weights per image instead of touching A)
};

__device__ __forceinline__ float xform(float x, float2 ab, int mode) {
  if (mode == XF_NONE) return x;
  if (mode == XF_SCALE) return ab.x * x;
  float y = fmaf(ab.x, x, ab.y);
  if (mode == XF_AFFINE_RELU6) y = fminf(fmaxf(y, 0.f), 6.f);
  else if (mode == XF_AFFINE_SILU) y = y / (1.f + __expf(-y));
  return y;
}

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// round-trip through the storage type (statistics are taken of the value that is stored)
template <typename T> __device__ __forceinline__ float rt(float v) { return to_f<T>(from_f<T>(v)); }

__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// 8 consecutive channels (one 16-byte bf16 vector / two float4)
template <typename T> struct Vec8;
template <> struct Vec8<bf16> {
  static __device__ __forceinline__ void load(const bf16* p, float (&v)[8]) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    v[0] = bf16lo(u.x); v[1] = bf16hi(u.x); v[2] = bf16lo(u.y); v[3] = bf16hi(u.y);
    v[4] = bf16lo(u.z); v[5] = bf16hi(u.z); v[6] = bf16lo(u.w); v[7] = bf16hi(u.w);
  }
  static __device__ __forceinline__ void store(bf16* p, const float (&v)[8]) {
    uint4 u;
    u.x = pack_bf16(v[0], v[1]); u.y = pack_bf16(v[2], v[3]);
    u.z = pack_bf16(v[4], v[5]); u.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};
template <> struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};

// ---- GEMM operand description shared by the CUDA-core and the tcgen05 kernels -------------------
// out[m][n] = sum over segments s, k < K_s of xform_s(A_s[m][k]) * W[n][koff_s + k]
// A_s is an NHWC activation of channel width ld_s (= K_s unless the segment is a channel slice).
#define LCM_MAX_SEGS 4
struct GemmSeg {
  const void* A;       // activation base (storage type of the plan)
  const float2* coef;  // per-(image, channel) affine, indexed [img * coef_ld + coef_off + k]; may be null
  int K;               // channels taken from this segment
  int ld;              // channel stride of the tensor (elements per pixel)
  int coef_ld;
  int coef_off;
  int mode;            // XformMode
  int f16;             // tcgen05 path only: this segment is stored as fp16 (hidden tensors of a block), not bf16
};

struct GemmParams {
  GemmSeg seg[LCM_MAX_SEGS];
  int nseg;
  int Ktot;
  const void* W;   // packed weights (layout depends on the kernel)
  void* out;       // [M][Nc] storage type
  double* stats;   // [images][Nc][2] (sum, sum of squares) or null
  long long M;     // rows = images * P
  int P;           // pixels per image
  int Nc;          // output channels
  int out_f16;     // tcgen05 path only: store the output as fp16 (the expand GEMM's hidden tensor)
};

}  // namespace lcm
