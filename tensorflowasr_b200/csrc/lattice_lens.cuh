// Lengths of one utterance's RNN-T lattice / CTC trellis, made safe for the kernels.
// The host raises logit_length to label_length like the reference (losses/base_loss.py:33-37), so a length can exceed the tensor's
// extent (more labels than encoder frames): the reference then returns an inf / garbage loss but stays inside its tensors.  Here such an
// utterance is declared infeasible: (Tb, Ub) = (0, 0) makes every kernel treat all of its rows as outside the lattice — cost = +inf, zero
// gradient, nothing read or written beyond the utterance's own slab.
#pragma once
__device__ __forceinline__ void lattice_lens(const int* __restrict__ logit_len, const int* __restrict__ label_len, int b, int T, int Umax,
                                             int& Tb, int& Ub) {
  Tb = logit_len[b];
  Ub = label_len[b];
  if (Tb > T || Ub > Umax || Ub < 0 || Tb < 0) {
    Tb = 0;
    Ub = 0;
  }
}
