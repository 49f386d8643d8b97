// Types shared by the kernels of the fused verify step.
#pragma once

#include "common.cuh"

namespace lantern {

struct RowStats {
  float thr;   // top-k threshold on the tempered value (keep s >= thr); -inf = keep all
  float mx;    // max of the tempered row
  float sum;   // sum over kept columns of exp(s - mx)
  float vcut;  // top-p: columns with (s, idx) <= (vcut, icut) are removed; -inf / -1 = none
  int icut;
  int kind;    // LANTERN_ROW_*
  int pad0, pad1;
};
static_assert(sizeof(RowStats) == 32, "RowStats layout");

struct AcceptParams {
  lantern_accept_cfg cfg;
  lantern_accept_in in;
  lantern_accept_out out;
  RowStats* stats;
  float* p_spill;  // [n_items, ncols rounded up to 4] walk probability vectors in global memory when they do not fit
                   // in shared memory (more than ~50K live columns); NULL otherwise
  MixParams mix;
  int vec_ok;     // rows can be read with 4-element vector loads
  int do_topk;    // 0 < top_k < ncols
  int do_topp;
  int tail_raw;   // vanilla: tail row softmax without the processors
  int lumina;
  int static_zero_q;  // static + relaxed rejection zeroes neighbours in q (LlamaGen/Anole) instead of gtp
  int prefetch_rows;  // walk: pull the row of the child being tried towards L2 (off when the logits sit in host memory)
  float inv_ncols;    // 1 / ncols
  float z_guess;      // inverse normal CDF of 1 - top_k/ncols: first bracket of the top-k select
  float win_sd;       // half-width of the bracket in standard deviations once a CTA tracks the observed quantile
  float win_sd_first; // half-width for a CTA's first row (Gaussian prior only)
};


__device__ __forceinline__ bool kept_col(float s, int idx, const RowStats& st) {
  return (s >= st.thr) && ((s > st.vcut) || (s == st.vcut && idx > st.icut));
}

}  // namespace lantern
