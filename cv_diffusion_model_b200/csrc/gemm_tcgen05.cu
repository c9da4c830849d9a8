// placeholder, replaced by the tcgen05 kernel
#include "kernels.h"
namespace lcm {
int gemm_tc_pick_block_n(int Nc) { return Nc; }
int launch_gemm_tc(const GemmParams&, const ConvGeom&, int, int, cudaStream_t) { return -1; }
}
