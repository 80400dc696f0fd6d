// tcgen05 GEMM kernels for the 160-column tile (N = 320 joint projections / dgrad, N = 144, 288 encoder widths): see gemm_kernel.cuh
#include "gemm_kernel.cuh"

namespace tfasr {
TFASR_GEMM_INSTANTIATE_KMAJOR(160)
}
