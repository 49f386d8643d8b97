// Top-p (nucleus) cut of one row, HF TopPLogitsWarper semantics (transformers 4.45, used by
// prepare_logits_processor, drafters/utils.py:48-49): sort ascending (stable), softmax, remove every position whose
// cumulative probability is <= 1 - top_p, always keep the last one.  The removed set is a prefix of the
// (value, column) order, so it is described by one cut (vcut, icut): a column is removed iff
// (s, idx) <= (vcut, icut).  Runs after the row-statistics kernel (which provides the row max), before the walk;
// rewrites RowStats.{vcut, icut, thr, sum}.  Slow path by design: no BASELINE config enables top-p.
#pragma once

#include "accept_types.cuh"

namespace lantern {

constexpr int kToppThreads = 512;

template <int DT, bool VEC>
__global__ void __launch_bounds__(kToppThreads) row_topp_kernel(const AcceptParams P) {
  __shared__ double hist[256];
  __shared__ double dscr[34];
  __shared__ int iscr[40];
  __shared__ unsigned ucr[4];
  __shared__ double dcr[4];
  const lantern_accept_cfg& cfg = P.cfg;
  const int row = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
  RowStats st = P.stats[row];
  if (st.kind != LANTERN_ROW_IMAGE) return;
  const int b = row / cfg.n_rows, t = row % cfg.n_rows;
  const int64_t base = (int64_t)b * cfg.item_stride + (int64_t)t * cfg.row_stride + cfg.col0;
  const MixParams mix = P.mix;
  auto value = [&](int e) -> float {
    const float c = Elem<DT>::load1(P.in.logits_cond, base + e);
    const float u = mix.has_uncond ? Elem<DT>::load1(P.in.logits_uncond, base + e) : 0.f;
    return mix_temper(c, u, mix);
  };
  const ExpShift ex(st.mx);
  // softmax denominator over the whole row (top-p sees the unfiltered tempered row)
  float part = 0.f;
  for (int e = tid; e < cfg.ncols; e += NT) part += ex(value(e));
  const float sum_all = (float)block_reduce((double)part, OpSum(), 0.0, dscr);
  const float inv = __fdiv_rn(1.0f, sum_all);
  auto prob = [&](float s) -> float { return __fmul_rn(ex(s), inv); };
  const float T = 1.0f - cfg.top_p;   // fp32, like torch's `cumulative_probs <= (1 - top_p)`

  // ---- radix descent (ascending) on the ordered value key: find the key of the first kept column ----
  uint32_t prefix = 0, mask = 0;
  double below = 0.0;   // mass of keys strictly below the current prefix range
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = tid; i < 256; i += NT) hist[i] = 0.0;
    __syncthreads();
    for (int e = tid; e < cfg.ncols; e += NT) {
      const float s = value(e);
      const uint32_t key = float_key(s);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xffu], (double)prob(s));
    }
    __syncthreads();
    if (tid == 0) {
      double run = below;
      int d = 255;
      for (int i = 0; i < 256; ++i) {
        if ((float)(run + hist[i]) > T) { d = i; break; }   // the cumulative sum first exceeds the bound in this digit
        run += hist[i];
      }
      ucr[0] = (unsigned)d;
      dcr[0] = run;
    }
    __syncthreads();
    prefix |= ucr[0] << shift;
    mask |= 0xffu << shift;
    below = dcr[0];
    __syncthreads();
  }
  const float vstar = key_float(prefix);   // value of the first kept column (ties resolved by column below)
  const float pv = prob(vstar);
  // ---- ties at vstar are removed in column order while the running sum stays within the bound ----
  const int per = (cfg.ncols + NT - 1) / NT;
  const int i0 = min(tid * per, cfg.ncols), i1 = min(i0 + per, cfg.ncols);
  int ties = 0, smaller = 0;
  for (int e = i0; e < i1; ++e) {
    const float s = value(e);
    ties += s == vstar;
    smaller += s < vstar;
  }
  double tot_ties_d;
  const double ties_incl = block_scan_incl((double)ties, dscr, &tot_ties_d);
  const int n_smaller = block_reduce(smaller, OpSum(), 0, iscr);
  const int n_ties = (int)tot_ties_d;
  if (tid == 0) {
    int r = 0;
    double run = below;
    while (r < n_ties && (float)(run + (double)pv) <= T) { run += (double)pv; ++r; }
    if (n_smaller + r >= cfg.ncols) r = cfg.ncols - n_smaller - 1;   // min_tokens_to_keep = 1
    iscr[36] = r;
    iscr[37] = -1;
  }
  __syncthreads();
  const int r = iscr[36];
  if (r > 0) {   // column of the r-th tie
    const int before = (int)ties_incl - ties;
    if (before < r && r <= before + ties) {
      int seen = before;
      for (int e = i0; e < i1; ++e) {
        if (value(e) == vstar && ++seen == r) { iscr[37] = e; break; }
      }
    }
  }
  __syncthreads();
  float vcut;
  int icut;
  if (r > 0) { vcut = vstar; icut = iscr[37]; }
  else {
    // nothing at vstar is removed: the cut is "everything strictly below vstar"
    vcut = vstar; icut = -1;
  }
  // ---- top-k on what is left (HF order: top-p, then top-k) and the final softmax sum ----
  const int remaining = cfg.ncols - n_smaller - r;
  float thr = st.thr;
  if (!P.do_topk || cfg.top_k >= remaining) thr = -INFINITY;
  RowStats out = st;
  out.vcut = vcut; out.icut = icut; out.thr = thr;
  float ksum = 0.f;
  for (int e = tid; e < cfg.ncols; e += NT) {
    const float s = value(e);
    if (kept_col(s, e, out)) ksum += ex(s);
  }
  out.sum = (float)block_reduce((double)ksum, OpSum(), 0.0, dscr);
  if (tid == 0) P.stats[row] = out;
}

}  // namespace lantern
