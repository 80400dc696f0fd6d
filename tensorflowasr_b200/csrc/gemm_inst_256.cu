// tcgen05 GEMM kernels for the 256-column tile (all epilogue feature classes): see gemm_kernel.cuh
#include "gemm_kernel.cuh"

namespace tfasr {
TFASR_GEMM_INSTANTIATE(256)
// lean epilogue + per-row softmax statistics (joint logits -> RNN-T loss): only the 256-column tile carries it
template int launch_gemm<256, false, false, 8u>(const void*, const void*, const GemmEpi&, int, int, int, int, int, int, cudaStream_t);
}
