// Host-side helper shared by the TMA-fed kernels: cuTensorMapEncodeTiled through the runtime's driver
// entry point (the library links cudart only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace lcm {

enum TmapDtype : int { TMAP_BF16 = 0, TMAP_F16 = 1 };

// rank-`rank` tiled map over 16-bit elements; gdim/box innermost first; gstride_bytes has rank-1 entries
// (strides of dims 1..rank-1).  Out-of-bounds elements of a box are zero-filled.
bool encode_tmap(CUtensorMap* out, int dtype, int rank, const void* ptr, const cuuint64_t* gdim,
                 const cuuint64_t* gstride_bytes, const cuuint32_t* box, bool swizzle128,
                 const cuuint32_t* elem_strides = nullptr);   // traversal strides per dimension (default 1)

// cached map of a [M][K] 16-bit matrix with row stride ld: box = 64 elements x 128 rows, 128-byte swizzle (gemm_expand.cu)
bool tmap_rows128(const void* ptr, long long M, int K, int ld, int dtype, CUtensorMap* out);

}  // namespace lcm
