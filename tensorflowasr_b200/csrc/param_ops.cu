// Parameter-side kernels of the training step for sm_100a: one flat AdamW launch and one batched weight-pack launch.
//
// The reference applies Keras Adam per variable (optimizers/accumulation.py, base_model.py:196-215 `optimizer.apply`) and casts
// weights on the fly inside each layer; at Conformer-S scale (~10 M parameters in ~500 tensors) that is several hundred tiny launches
// per step.  Here every parameter, gradient and moment lives in one flat fp32 buffer (dist.GradientReducer / FlatAdamW), so the
// update is ONE HBM-bound pass (28 B read + 16 B written per parameter), and all bf16 operand copies of the GEMM weights (plain,
// transposed, Q|K|V-concatenated, GLU-interleaved) are refreshed by ONE launch driven by a device-resident table.
#include "common.cuh"

#include <cuda_bf16.h>

namespace tfasr {

// p, g, m, v: flat fp32 [n].  AdamW with decoupled weight decay in the form of torch.optim.AdamW.  Keras Adam(weight_decay) differs in ONE
// place: it adds epsilon to sqrt(v) before the bias correction (p -= lr sqrt(1-b2^t)/(1-b1^t) m / (sqrt(v) + eps)), i.e. its effective epsilon
// is eps / sqrt(1 - b2^t) — with the reference's eps = 1e-9 (small.yml.j2:73-87) the two updates agree to ~1e-8 relative from the first step on:
//   g' = g * grad_scale + 2 * l2 * p * mask      (Keras l2 regularizers of conformer.py:27, applied through their gradient)
//   p *= 1 - lr * wd;  m = b1 m + (1 - b1) g';  v = b2 v + (1 - b2) g'^2
//   p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// lr and t are read from device memory (CUDA-graph replays pick up the schedule); g is zeroed for the next step when asked.
__global__ void __launch_bounds__(256) adamw_flat_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                         const uint8_t* __restrict__ l2_mask, long long n4, const float* __restrict__ lr_ptr,
                                                         const float* __restrict__ step_ptr, float b1, float b2, float eps, float wd, float l2,
                                                         float grad_scale, int zero_grad) {
  const float lr = *lr_ptr, t = *step_ptr;
  const float bc1 = 1.0f - powf(b1, t), bc2 = 1.0f - powf(b2, t);
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2), decay = 1.0f - lr * wd;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    uint32_t mk = l2_mask ? reinterpret_cast<const uint32_t*>(l2_mask)[i] : 0u;
    float* pe = &pp.x;
    const float* ge = &gg.x;
    float* me = &mm.x;
    float* ve = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gk = ge[k] * grad_scale;
      if ((mk >> (8 * k)) & 0xFFu) gk = fmaf(2.0f * l2, pe[k], gk);
      float pk = pe[k] * decay;
      const float mn = fmaf(b1, me[k] - gk, gk);           // b1 m + (1 - b1) g
      const float vn = fmaf(b2, ve[k] - gk * gk, gk * gk);  // b2 v + (1 - b2) g^2
      pk -= step_size * mn / (sqrtf(vn) * inv_sqrt_bc2 + eps);
      pe[k] = pk;
      me[k] = mn;
      ve[k] = vn;
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// One table entry = one fp32 source matrix [K, N] (row stride src_ld) scattered into bf16 operand buffers:
//   dst_w [K rows, column c0 + n * cs]   (row stride ldw)     -> K-major "B" operand of the data-gradient GEMM
//   dst_wt[row c0 + n * cs, column k]    (row stride ldwt)    -> K-major "B" operand of the forward GEMM
// c0 / cs express the Q|K|V concatenation (c0 = 0, NH, 2NH) and the GLU interleave (cs = 2, c0 = 0 / 1).  flags bit 0: dst_w is fp32
// (bias vectors re-laid out the same way); either destination may be null.
struct PackEntry {
  const float* src;
  void* dst_w;
  void* dst_wt;
  int K, N, src_ld, ldw, ldwt, c0, cs, flags;
};

__global__ void __launch_bounds__(256) pack_weights_kernel(const PackEntry* __restrict__ table) {
  const PackEntry e = table[blockIdx.x];
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int kt_n = (e.K + 31) / 32, nt_n = (e.N + 31) / 32;
  for (int t = blockIdx.y; t < kt_n * nt_n; t += gridDim.y) {  // tiles of one matrix are spread over gridDim.y CTAs
    const int k0 = (t / nt_n) * 32, n0 = (t % nt_n) * 32;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
      const int k = k0 + r, n = n0 + tx;
      float x = 0.f;
      if (k < e.K && n < e.N) {
        x = e.src[(size_t)k * e.src_ld + n];
        if (e.dst_w) {
          const size_t o = (size_t)k * e.ldw + e.c0 + (size_t)n * e.cs;
          if (e.flags & 1) static_cast<float*>(e.dst_w)[o] = x;
          else static_cast<__nv_bfloat16*>(e.dst_w)[o] = __float2bfloat16(x);
        }
      }
      tile[r][tx] = x;
    }
    __syncthreads();
    if (e.dst_wt) {
#pragma unroll
      for (int r = ty; r < 32; r += 8) {
        const int n = n0 + r, k = k0 + tx;
        if (n < e.N && k < e.K) static_cast<__nv_bfloat16*>(e.dst_wt)[((size_t)e.c0 + (size_t)n * e.cs) * e.ldwt + k] = __float2bfloat16(tile[tx][r]);
      }
    }
    __syncthreads();
  }
}

// Inverse of the pack for gradients: the weight-gradient GEMMs of the fused Q|K|V and GLU-interleaved projections accumulate into
// persistent packed fp32 buffers; at the end of the backward ONE launch adds every slice into its parameter's gradient
//   grad[k, n] += packed[k, c0 + n * cs]      and clears the packed element for the next step.
// Same table format: src = destination gradient (fp32 [K, N], row stride src_ld), dst_w = packed buffer (row stride ldw).
__global__ void __launch_bounds__(256) unpack_add_kernel(const PackEntry* __restrict__ table) {
  const PackEntry e = table[blockIdx.x];
  float* grad = const_cast<float*>(e.src);
  float* packed = static_cast<float*>(e.dst_w);
  const int total = e.K * e.N;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < total; i += gridDim.y * blockDim.x) {
    const int k = i / e.N, n = i - k * e.N;
    const size_t o = (size_t)k * e.ldw + e.c0 + (size_t)n * e.cs;
    grad[(size_t)k * e.src_ld + n] += packed[o];
    packed[o] = 0.f;
  }
}

}  // namespace tfasr

using namespace tfasr;

extern "C" int tfasr_adamw_flat_f32(float* p, float* g, float* m, float* v, const uint8_t* l2_mask, int64_t n, const float* lr, const float* step,
                                    float beta1, float beta2, float eps, float weight_decay, float l2, float grad_scale, int32_t zero_grad,
                                    tfasr_stream_t stream) {
  if (!p || !g || !m || !v || !lr || !step || n < 0 || (n & 3)) return TFASR_INVALID_VALUE;  // buffers are padded to 4 elements
  if ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15)
    return TFASR_INVALID_VALUE;
  if (l2_mask && (reinterpret_cast<uintptr_t>(l2_mask) & 3)) return TFASR_INVALID_VALUE;
  if (n == 0) return TFASR_OK;
  const long long n4 = n / 4;
  long long want = (n4 + 255) / 256;
  if (want > (long long)kNumSMs * 8) want = (long long)kNumSMs * 8;
  const int blocks = (int)want;
  adamw_flat_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(p, g, m, v, l2_mask, n4, lr, step, beta1, beta2, eps, weight_decay, l2,
                                                                           grad_scale, zero_grad);
  count_launch();
  return check_launch();
}

extern "C" int tfasr_pack_weights(const void* table, int32_t n_entries, tfasr_stream_t stream) {
  static_assert(sizeof(PackEntry) == sizeof(tfasr_pack_entry), "C ABI struct layout");
  if (!table || n_entries < 0) return TFASR_INVALID_VALUE;
  if (n_entries == 0) return TFASR_OK;
  pack_weights_kernel<<<dim3((unsigned)n_entries, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const PackEntry*>(table));
  count_launch();
  return check_launch();
}

extern "C" int tfasr_unpack_add(const void* table, int32_t n_entries, tfasr_stream_t stream) {
  if (!table || n_entries < 0) return TFASR_INVALID_VALUE;
  if (n_entries == 0) return TFASR_OK;
  unpack_add_kernel<<<dim3((unsigned)n_entries, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const PackEntry*>(table));
  count_launch();
  return check_launch();
}
