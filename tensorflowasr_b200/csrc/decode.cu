// C-ABI launcher of the persistent greedy transducer decode kernel (decode_kernels.cuh).  Declared in include/tfasr_b200.h.
// Builds twice like seqconv.cu: nvcc -> libtfasr_b200.so, g++ -DTFASR_EMU_BUILD -> tests/emu/libtfasr_emu.so (host pointers, CPU emulator).
#include "decode_kernels.cuh"

#ifdef TFASR_EMU_BUILD
#include "../../include/tfasr_b200.h"
typedef void* cudaStream_t;
#endif

using namespace tfasr;

extern "C" int tfasr_greedy_decode_f32(const tfasr_greedy_params* q, int32_t B, int32_t threads, tfasr_stream_t stream_) {
  if (!q || B < 0 || threads < 32 || threads > 1024 || (threads & 31)) return TFASR_INVALID_VALUE;
  if (!q->e_proj || !q->nframes || !q->emb || !q->w_ih || !q->w_hh || !q->b || !q->wp || !q->bp || !q->wv || !q->bv || !q->prev_tok || !q->states ||
      !q->frame_idx || !q->tok_idx || !q->tokens || !q->iters)
    return TFASR_INVALID_VALUE;
  if ((q->ln_g == nullptr) != (q->ln_b == nullptr)) return TFASR_INVALID_VALUE;
  if (q->T < 1 || q->E < 1 || q->R < 1 || q->J < 1 || q->V < 1 || q->max_tokens < 1 || q->blank < 0 || q->blank >= q->V) return TFASR_INVALID_VALUE;
  if (q->mode < 0 || q->mode > 2 || (q->mode == 0 && (!q->cap_iter || q->max_iters < 0)) || (q->mode == 2 && q->per_frame_cap < 1))
    return TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  GreedyParams p;
  p.e_proj = q->e_proj;
  p.nframes = q->nframes;
  p.emb = q->emb;
  p.w_ih = q->w_ih;
  p.w_hh = q->w_hh;
  p.b = q->b;
  p.ln_g = q->ln_g;
  p.ln_b = q->ln_b;
  p.wp = q->wp;
  p.bp = q->bp;
  p.wv = q->wv;
  p.bv = q->bv;
  p.prev_tok = q->prev_tok;
  p.states = q->states;
  p.frame_idx = q->frame_idx;
  p.tok_idx = q->tok_idx;
  p.tokens = q->tokens;
  p.iters = q->iters;
  p.cap_iter = q->cap_iter;
  p.T = q->T;
  p.E = q->E;
  p.R = q->R;
  p.J = q->J;
  p.V = q->V;
  p.max_tokens = q->max_tokens;
  p.blank = q->blank;
  p.mode = q->mode;
  p.max_iters = q->max_iters;
  p.per_frame_cap = q->per_frame_cap;
  // split-K scratch: the largest ks * N any of the three matrix shapes can use with this CTA size
  int part = 0;
  const int Ns[3] = {4 * p.R, p.J, p.V};
  for (int N : Ns) {
    if (N & 3) continue;
    int ks = threads / (N / 4);
    if (ks > kMvMaxSplit) ks = kMvMaxSplit;
    if (ks > 1 && ks * N > part) part = ks * N;
  }
  const size_t kMax = 200 * 1024;
  while (part > 0 && greedy_smem_floats(p.E, p.R, p.J, p.V, part) * sizeof(float) > kMax) part /= 2;
  p.part_cap = part;
  const size_t smem = greedy_smem_floats(p.E, p.R, p.J, p.V, part) * sizeof(float);
  if (smem > kMax) return TFASR_INVALID_VALUE;
#ifdef TFASR_EMU_BUILD
  (void)stream_;
  emu::launch(dim3((unsigned)B), dim3((unsigned)threads), smem, [&] { greedy_decode_kernel(p); });
  return TFASR_OK;
#else
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (smem > 48 * 1024 && cudaFuncSetAttribute(greedy_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return TFASR_EXECUTION_FAILED;
  greedy_decode_kernel<<<dim3((unsigned)B), dim3((unsigned)threads), smem, stream>>>(p);
  count_launch();
  return check_launch();
#endif
}
