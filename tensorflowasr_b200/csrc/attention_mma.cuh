// Warp-level tensor-core helpers for the attention kernels (mma.sync m16n8k16, bf16 x bf16 -> f32; ldmatrix).
// tcgen05 is not used here on purpose: the per-(utterance, head) problems are 250 x 250 x 36 — a 128-row UMMA tile with TMEM round
// trips would be mostly padding; 16 x 8 register tiles fit them exactly (SASS: HMMA.16816, LDSM).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace tfasr {
namespace mma {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// A fragment (16 x 16, row-major source): lane supplies the address of row (lane % 16), column block (lane / 16) * 8.
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* row_ptr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_addr(row_ptr)));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* row_ptr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_addr(row_ptr)));
}
// B fragment (16 x 8): source stored [n][k] (k contiguous): lanes 0-7 -> rows n, k-block 0; lanes 8-15 -> rows n, k-block 1.
__device__ __forceinline__ void ldsm_x2(uint32_t (&r)[2], const void* row_ptr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(smem_addr(row_ptr)));
}
// B fragment from a source stored [k][n] (n contiguous): lanes 0-15 supply the addresses of rows k = 0..15 at column n0.
__device__ __forceinline__ void ldsm_x2_trans(uint32_t (&r)[2], const void* row_ptr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(smem_addr(row_ptr)));
}
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// asynchronous global -> shared copies (LDGSTS): no register staging, many in flight per thread
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

}  // namespace mma
}  // namespace tfasr
