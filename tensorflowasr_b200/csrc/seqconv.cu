// C-ABI launchers of the ContextNet / DeepSpeech2 encoder-body kernels (seqconv_kernels.cuh).  Declared in include/tfasr_b200.h.
//
// The same file builds twice: with nvcc into libtfasr_b200.so, and with g++ -DTFASR_EMU_BUILD into tests/emu/libtfasr_emu.so, where
// the identical launch geometry runs on the CPU emulator with HOST pointers (test infrastructure, tests/test_emu_kernels.py).
#include "seqconv_kernels.cuh"

#ifdef TFASR_EMU_BUILD
#include "../../include/tfasr_b200.h"
typedef void* cudaStream_t;
#define TFASR_LAUNCH(kern, L, stream, ...) emu::launch((L).grid, (L).block, (L).smem, [&] { kern(__VA_ARGS__); })
#define TFASR_ZERO(ptr, bytes, stream) (memset((ptr), 0, (bytes)), 0)
#define TFASR_SMEM_OPT_IN(kern, bytes) 0
#define TFASR_LAUNCH_STATUS() TFASR_OK
#else
#define TFASR_LAUNCH(kern, L, stream, ...)                           \
  do {                                                               \
    kern<<<(L).grid, (L).block, (L).smem, (stream)>>>(__VA_ARGS__);  \
    count_launch();                                                  \
  } while (0)
#define TFASR_ZERO(ptr, bytes, stream) (cudaMemsetAsync((ptr), 0, (bytes), (stream)) != cudaSuccess)
#define TFASR_SMEM_OPT_IN(kern, bytes) \
  ((bytes) > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)) != cudaSuccess)
#define TFASR_LAUNCH_STATUS() check_launch()
#endif

using namespace tfasr;

namespace {
constexpr size_t kMaxSmem = 200 * 1024;
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
inline int out_frames(int T, int stride) { return (T + stride - 1) / stride; }
}  // namespace

extern "C" int tfasr_seq_dwconv_fwd(const void* x, const float* w, void* y, int32_t B, int32_t T, int32_t C, int32_t K, int32_t stride,
                                    tfasr_stream_t stream_) {
  if (!x || !w || !y || B < 0 || B > 65535 || T < 1 || C < 8 || (C & 7) || K < 1 || stride < 1 || !aligned16(x) || !aligned16(y))
    return TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int To = out_frames(T, stride);
  const SeqLaunch L = seq_dwconv_launch(B, To, C, K);
  if (L.smem > kMaxSmem) return TFASR_INVALID_VALUE;
  if (TFASR_SMEM_OPT_IN(seq_dwconv_kernel<false>, L.smem)) return TFASR_EXECUTION_FAILED;
  TFASR_LAUNCH(seq_dwconv_kernel<false>, L, stream, static_cast<const __nv_bfloat16*>(x), w, static_cast<__nv_bfloat16*>(y), T, To, C, K, stride);
  return TFASR_LAUNCH_STATUS();
}

extern "C" int tfasr_seq_dwconv_bwd(const void* dy, const void* x, const float* w, void* dx, float* dw, int32_t B, int32_t T, int32_t C, int32_t K,
                                    int32_t stride, tfasr_stream_t stream_) {
  if (!dy || !x || !w || !dw || B < 0 || B > 65535 || T < 1 || C < 8 || (C & 7) || K < 1 || K > 32 || stride < 1 || !aligned16(x) ||
      !aligned16(dy) || !aligned16(dx))
    return TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int To = out_frames(T, stride);
  if (dx) {  // data gradient over the INPUT frames (skipped for the first layer: the features carry no gradient)
    const SeqLaunch L = seq_dwconv_launch(B, T, C, K);
    if (L.smem > kMaxSmem) return TFASR_INVALID_VALUE;
    if (TFASR_SMEM_OPT_IN(seq_dwconv_kernel<true>, L.smem)) return TFASR_EXECUTION_FAILED;
    TFASR_LAUNCH(seq_dwconv_kernel<true>, L, stream, static_cast<const __nv_bfloat16*>(dy), w, static_cast<__nv_bfloat16*>(dx), To, T, C, K, stride);
  }
  int rpb = 0;
  const SeqLaunch L = seq_dwconv_wgrad_launch(B, To, C, K, &rpb);
  if (K <= 8) {
    TFASR_LAUNCH(seq_dwconv_wgrad_kernel<8>, L, stream, static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(x), dw, B, T, To, C, K,
                 stride, rpb);
  } else {
    TFASR_LAUNCH(seq_dwconv_wgrad_kernel<32>, L, stream, static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(x), dw, B, T, To, C, K,
                 stride, rpb);
  }
  return TFASR_LAUNCH_STATUS();
}

extern "C" int tfasr_seq_masked_mean(const void* y, const int32_t* length, float* m, int32_t B, int32_t T, int32_t C, tfasr_stream_t stream_) {
  if (!y || !length || !m || B < 0 || B > 65535 || T < 1 || C < 2 || (C & 1)) return TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (TFASR_ZERO(m, sizeof(float) * (size_t)B * C, stream)) return TFASR_EXECUTION_FAILED;
  const int gx = (int)ceil_div(C / 2, 64);
  const int fps = seq_split_frames(T, B, gx);
  const SeqLaunch L = {dim3((unsigned)gx, (unsigned)ceil_div(T, fps), (unsigned)B), dim3(64, 4), 0};
  if (L.grid.y > 65535) return TFASR_INVALID_VALUE;
  TFASR_LAUNCH(seq_masked_mean_kernel, L, stream, static_cast<const __nv_bfloat16*>(y), length, m, T, C, fps);
  return TFASR_LAUNCH_STATUS();
}

extern "C" int tfasr_se_fc_fwd(const float* m, const float* W1, const float* b1, const float* W2, const float* b2, float* z1, float* s, int32_t B,
                               int32_t C, int32_t H, int32_t act, tfasr_stream_t stream_) {
  if (!m || !W1 || !b1 || !W2 || !b2 || !z1 || !s || B < 0 || C < 1 || H < 1 || act < 0 || act > 2) return TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const SeqLaunch L = {dim3((unsigned)B), dim3(256), sizeof(float) * (size_t)(C + H)};
  if (L.smem > 48 * 1024) return TFASR_INVALID_VALUE;
  TFASR_LAUNCH(se_fc_fwd_kernel, L, stream, m, W1, b1, W2, b2, z1, s, C, H, act);
  return TFASR_LAUNCH_STATUS();
}

extern "C" int tfasr_se_fc_bwd(const float* ds, const float* s, const float* z1, const float* m, const float* W1, const float* W2,
                               const int32_t* length, float* dz2, float* dz1, float* dm, float* dW1, float* db1, float* dW2, float* db2, int32_t B,
                               int32_t T, int32_t C, int32_t H, int32_t act, tfasr_stream_t stream_) {
  if (!ds || !s || !z1 || !m || !W1 || !W2 || !length || !dz2 || !dz1 || !dm || !dW1 || !db1 || !dW2 || !db2 || B < 0 || T < 1 || C < 1 || H < 1 ||
      act < 0 || act > 2)
    return TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const SeqLaunch L = {dim3((unsigned)B), dim3(256), sizeof(float) * (size_t)(C + H)};
  if (L.smem > 48 * 1024) return TFASR_INVALID_VALUE;
  TFASR_LAUNCH(se_fc_bwd_data_kernel, L, stream, ds, s, z1, W1, W2, length, dz2, dz1, dm, T, C, H, act);
  const int64_t total = 2 * (int64_t)C * H + C + H;
  const SeqLaunch P = {dim3(seq_ew_grid(total)), dim3(256), 0};
  TFASR_LAUNCH(se_fc_bwd_params_kernel, P, stream, m, z1, dz1, dz2, dW1, db1, dW2, db2, B, C, H, act);
  return TFASR_LAUNCH_STATUS();
}

extern "C" int tfasr_se_apply_fwd(const void* y, const float* s, const void* res, void* out, int32_t B, int32_t T, int32_t C, int32_t act,
                                  tfasr_stream_t stream_) {
  if (!y || !s || !out || B < 0 || T < 1 || C < 8 || (C & 7) || act < 0 || act > 2 || !aligned16(y) || !aligned16(res) || !aligned16(out))
    return TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int64_t rows = (int64_t)B * T;
  const SeqLaunch L = {dim3(seq_ew_grid(rows * (C / 8))), dim3(256), 0};
  TFASR_LAUNCH(se_apply_fwd_kernel, L, stream, static_cast<const __nv_bfloat16*>(y), s, static_cast<const __nv_bfloat16*>(res),
               static_cast<__nv_bfloat16*>(out), rows, T, C, act);
  return TFASR_LAUNCH_STATUS();
}

extern "C" int tfasr_se_apply_bwd(const void* dout, const void* y, const float* s, const void* res, void* dres, float* ds, int32_t B, int32_t T,
                                  int32_t C, int32_t act, tfasr_stream_t stream_) {
  if (!dout || !y || !s || !dres || !ds || B < 0 || B > 65535 || T < 1 || C < 8 || (C & 7) || act < 0 || act > 2 || !aligned16(dout) ||
      !aligned16(y) || !aligned16(res) || !aligned16(dres))
    return TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (TFASR_ZERO(ds, sizeof(float) * (size_t)B * C, stream)) return TFASR_EXECUTION_FAILED;
  const int gx = (int)ceil_div(C / 8, 32);
  const int fps = seq_split_frames(T, B, gx);
  const SeqLaunch L = {dim3((unsigned)gx, (unsigned)ceil_div(T, fps), (unsigned)B), dim3(32, 8), 0};
  if (L.grid.y > 65535) return TFASR_INVALID_VALUE;
  TFASR_LAUNCH(se_apply_bwd_kernel, L, stream, static_cast<const __nv_bfloat16*>(dout), static_cast<const __nv_bfloat16*>(y), s,
               static_cast<const __nv_bfloat16*>(res), static_cast<__nv_bfloat16*>(dres), ds, T, C, act, fps);
  return TFASR_LAUNCH_STATUS();
}

extern "C" int tfasr_se_dy(const void* dres, const float* s, const float* dm, const int32_t* length, void* dy, int32_t B, int32_t T, int32_t C,
                           tfasr_stream_t stream_) {
  if (!dres || !s || !dm || !length || !dy || B < 0 || T < 1 || C < 8 || (C & 7) || !aligned16(dres) || !aligned16(dy)) return TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int64_t rows = (int64_t)B * T;
  const SeqLaunch L = {dim3(seq_ew_grid(rows * (C / 8))), dim3(256), 0};
  TFASR_LAUNCH(se_dy_kernel, L, stream, static_cast<const __nv_bfloat16*>(dres), s, dm, length, static_cast<__nv_bfloat16*>(dy), rows, T, C);
  return TFASR_LAUNCH_STATUS();
}

namespace {
// -> false when the geometry is inconsistent
bool conv_geom(const tfasr_conv_geom* p, ConvGeom* g) {
  if (!p || p->B < 0 || p->T < 1 || p->F < 1 || p->Cin < 1 || p->kh < 1 || p->kw < 1 || p->sh < 1 || p->sw < 1 || p->pt < 0 || p->pl < 0 ||
      p->To < 1 || p->Fo < 1)
    return false;
  const int kreal = p->kh * p->kw * p->Cin;
  if (p->Kp != (kreal + 7) / 8 * 8) return false;
  *g = ConvGeom{p->B, p->T, p->F, p->Cin, p->To, p->Fo, p->kh, p->kw, p->sh, p->sw, p->pt, p->pl, p->Kp};
  return true;
}
}  // namespace

extern "C" int tfasr_conv_im2col(const void* x, int32_t x_f32, void* col, const tfasr_conv_geom* geom, tfasr_stream_t stream_) {
  ConvGeom g;
  if (!x || !col || !conv_geom(geom, &g) || !aligned16(col) || (!x_f32 && !(g.Cin & 7) && !aligned16(x))) return TFASR_INVALID_VALUE;
  if (g.B == 0) return TFASR_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int64_t items = (int64_t)g.B * g.To * g.Fo * (g.Kp / 8);
  const SeqLaunch L = {dim3(seq_ew_grid(items)), dim3(256), 0};
  if (x_f32) {
    TFASR_LAUNCH(conv_im2col_kernel<true>, L, stream, x, static_cast<__nv_bfloat16*>(col), g);
  } else {
    TFASR_LAUNCH(conv_im2col_kernel<false>, L, stream, x, static_cast<__nv_bfloat16*>(col), g);
  }
  return TFASR_LAUNCH_STATUS();
}

extern "C" int tfasr_conv_col2im(const void* dcol, void* da, const tfasr_conv_geom* geom, tfasr_stream_t stream_) {
  ConvGeom g;
  if (!dcol || !da || !conv_geom(geom, &g) || (g.Cin & 7) || !aligned16(dcol) || !aligned16(da)) return TFASR_INVALID_VALUE;
  if (g.B == 0) return TFASR_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int64_t items = (int64_t)g.B * g.T * g.F * (g.Cin / 8);
  const SeqLaunch L = {dim3(seq_ew_grid(items)), dim3(256), 0};
  TFASR_LAUNCH(conv_col2im_kernel, L, stream, static_cast<const __nv_bfloat16*>(dcol), static_cast<__nv_bfloat16*>(da), g);
  return TFASR_LAUNCH_STATUS();
}
