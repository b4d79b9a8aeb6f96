// Arguments of one recurrent sweep (all time steps of one layer, both directions).
#pragma once
#include <stdint.h>

namespace ds2 {

struct SeqArgs {
  int T, B, H, D, G;
  const int32_t* len;
  float* gates;  // (T,B,D,G*H): input projections in, gate activations (fwd) / gate gradients (bwd) out
  float* hseq;   // (D,T,B,H) per-direction outputs, zero at masked steps
  float* aux;    // (D,T,B,H) LSTM cell states / GRU W_hn h + b_hn (-> dGh_n after bwd); null for tanh
  const float* w_hh[2];   // fwd: (G*H,H) ; bwd: transposed (H,G*H)
  const float* b_ih[2];
  const float* b_hh[2];
  const float* h0;        // (D,B,H) or null
  const float* c0;
  const float* dy;        // bwd: (T,B,H)
  float* carry;           // bwd: (D,B,H) dc (LSTM) / dh (GRU)
  int training;
  // bwd, optional: bias gradients accumulated inside the sweep (column sums of the gate gradients over (t,b)).
  // dbias[d]: (G*H) zeroed by the caller; dbias_hn[d]: GRU only, (H) sum of the h-side n-gate gradient.
  // The sweep sets *dbias_done = 1 when it filled them (otherwise the caller runs the column-sum kernels).
  float* dbias[2];
  float* dbias_hn[2];
  int* dbias_done;
  // bwd, optional (precision-16 GEMM operands): fp16 copies of the gate gradients, multiplied by the power of two
  // f16_scale[0] (device), written by the sweep as it produces them — f16_dg (T*B, D*G*H) row-major, f16_dgT
  // (D*G*H, T*B) transposed, f16_auxT (D*H, T*B) GRU h-side n-gate gradient transposed.  The sweep sets
  // *f16_done = 1 when it filled them (otherwise the caller converts from the fp32 gate gradients).
  void* f16_dg;
  void* f16_dgT;
  void* f16_auxT;
  const float* f16_scale;
  int* f16_done;
  // bwd, optional: fp16 copy of the transposed recurrent matrix (H, G*H) made by the forward pass of the same step.
  // The fp16-resident sweep takes it as it is; then w_hh[] points at buffers that are only filled (fp32 transposes)
  // when a path that needs them calls materialize_w_hh().
  const void* w_hhT16[2];
  int (*fill_w_hh)(void* ctx, void* stream);
  void* fill_w_hh_ctx;
};

inline int materialize_w_hh(const SeqArgs& a, void* stream) {
  return a.fill_w_hh ? a.fill_w_hh(a.fill_w_hh_ctx, stream) : 0;
}

}  // namespace ds2
