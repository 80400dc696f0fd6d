// tcgen05 GEMM kernels for the 128-column tile (all epilogue feature classes): see gemm_kernel.cuh
#include "gemm_kernel.cuh"

namespace tfasr {
TFASR_GEMM_INSTANTIATE(128)
}
