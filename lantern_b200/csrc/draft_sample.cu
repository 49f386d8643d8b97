// Static-tree drafter sampling — SURVEY.md 8(f) row N3: Model.sample of the reference
// (models/drafters/cnets_llamagen.py:924-940, cnets_lumina_mgpt.py:936-955): warp the drafter logits, softmax,
// draw k tokens without replacement, return their conditional probabilities p_i / (1 - sum_{j<i} p_j) clamped to
// [0, 1] and the full distribution (`op`, which evaluate_posterior_v1 subtracts from the target residual).
//
// The warp + softmax statistics come from the same row-statistics kernels as the verification step (CFG mix, exact
// top-k, top-p), so this file only adds the draw: an exponential race (key = -log(u) / p, u from the device Philox
// stream, one uniform per column) whose k smallest keys are a sample without replacement with the same law as
// torch.multinomial(p, k, replacement=False).
#include "accept_types.cuh"

namespace lantern {

constexpr int kDraftThreads = 512;

__device__ __forceinline__ double race_key(float p, uint64_t seed, uint64_t step, uint32_t row, uint32_t col) {
  uint32_t c[4] = {col >> 2, static_cast<uint32_t>(step), row, static_cast<uint32_t>(step >> 32) ^ 0x5A17u};
  philox4x32_10(c, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  const double u = ((double)(c[col & 3] >> 8) + 1.0) * (1.0 / 16777216.0);   // (0, 1]
  return -log(u) / (double)p;
}

template <int DT>
__global__ void __launch_bounds__(kDraftThreads) draft_sample_kernel(const AcceptParams P, int k, float* probs_out,
                                                                     int32_t* idx_out, float* cond_out) {
  __shared__ double best_key[kDraftThreads / 32];
  __shared__ int best_idx[kDraftThreads / 32];
  __shared__ double last_key_s;
  __shared__ int last_idx_s;
  __shared__ float picked_p[64];
  const lantern_accept_cfg& cfg = P.cfg;
  const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, V = cfg.vocab;
  const RowStats st = P.stats[row];
  const int b = row / cfg.n_rows, t = row % cfg.n_rows;
  const int64_t base = (int64_t)b * cfg.item_stride + (int64_t)t * cfg.row_stride + cfg.col0;
  const MixParams mix = P.mix;
  const ExpShift ex(st.mx);
  const float inv = __fdiv_rn(1.0f, st.sum);
  auto prob = [&](int e) -> float {
    const float c = Elem<DT>::load1(P.in.logits_cond, base + e);
    const float u = mix.has_uncond ? Elem<DT>::load1(P.in.logits_uncond, base + e) : 0.f;
    const float s = mix_temper(c, u, mix);
    return kept_col(s, e, st) ? __fmul_rn(ex(s), inv) : 0.f;
  };
  float* prow = probs_out + (size_t)row * V;
  for (int v = tid; v < V; v += kDraftThreads) {
    const int e = v - cfg.col0;
    prow[v] = (e >= 0 && e < cfg.ncols) ? prob(e) : 0.f;
  }
  if (tid == 0) { last_key_s = -1.0; last_idx_s = -1; }
  __syncthreads();
  for (int r = 0; r < k; ++r) {
    const double lk = last_key_s;
    const int li = last_idx_s;
    double bk = INFINITY;
    int bi = 0x7fffffff;
    for (int e = tid; e < cfg.ncols; e += kDraftThreads) {
      const float p = prow[e + cfg.col0];
      if (p <= 0.f) continue;
      const double key = race_key(p, cfg.philox_seed, cfg.philox_step, (uint32_t)row, (uint32_t)e);
      const bool after = key > lk || (key == lk && e > li);   // strictly after the previous pick in (key, column) order
      if (after && (key < bk || (key == bk && e < bi))) { bk = key; bi = e; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ok = __shfl_xor_sync(0xffffffffu, bk, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ok < bk || (ok == bk && oi < bi)) { bk = ok; bi = oi; }
    }
    if (lane == 0) { best_key[warp] = bk; best_idx[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < kDraftThreads / 32; ++w)
        if (best_key[w] < bk || (best_key[w] == bk && best_idx[w] < bi)) { bk = best_key[w]; bi = best_idx[w]; }
      last_key_s = bk;
      last_idx_s = bi;
      const bool ok = bi != 0x7fffffff;   // fewer than k columns with positive probability: pad like torch would fail
      idx_out[(size_t)row * k + r] = ok ? bi + cfg.col0 : -1;
      picked_p[r] = ok ? prow[bi + cfg.col0] : 0.f;
    }
    __syncthreads();
  }
  if (tid == 0) {   // conditional probabilities (cnets_llamagen.py:930-938): cumsum in fp64, rounded per prefix
    double acc = 0.0;
    float excl = 0.f;
    for (int r = 0; r < k; ++r) {
      float cp = __fdiv_rn(picked_p[r], __fsub_rn(1.0f, excl));
      if (isinf(cp) || isnan(cp)) cp = -1.0f;
      cp = fminf(fmaxf(cp, 0.0f), 1.0f);
      cond_out[(size_t)row * k + r] = cp;
      acc += (double)picked_p[r];
      excl = (float)acc;
    }
  }
}

}  // namespace lantern

using namespace lantern;

int accept_row_stats_only(const lantern_accept_cfg* cfg, const lantern_accept_in* in, void* workspace_dev,
                          size_t workspace_bytes, void* stream, AcceptParams* params_out);

// logits: [n_rows, V] (cond, optional uncond -> CFG mix), warp knobs and Philox seed/step in cfg (n_items = 1).
extern "C" LANTERN_API int lantern_draft_sample(const lantern_accept_cfg* cfg, const lantern_accept_in* in, int32_t k,
                                                float* probs_dev, int32_t* idx_dev, float* cond_probs_dev,
                                                void* workspace_dev, size_t workspace_bytes, void* stream) {
  if (!cfg || !in || !probs_dev || !idx_dev || !cond_probs_dev || k < 1 || k > 64) {
    set_error("lantern_draft_sample: bad argument (1 <= k <= 64)");
    return LANTERN_E_INVALID;
  }
  AcceptParams P;
  const int rc = accept_row_stats_only(cfg, in, workspace_dev, workspace_bytes, stream, &P);
  if (rc) return rc;
  const unsigned rows = (unsigned)(cfg->n_items * cfg->n_rows);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (cfg->logits_dtype) {
    case LANTERN_F32: draft_sample_kernel<LANTERN_F32><<<rows, kDraftThreads, 0, s>>>(P, k, probs_dev, idx_dev, cond_probs_dev); break;
    case LANTERN_BF16: draft_sample_kernel<LANTERN_BF16><<<rows, kDraftThreads, 0, s>>>(P, k, probs_dev, idx_dev, cond_probs_dev); break;
    default: draft_sample_kernel<LANTERN_F16><<<rows, kDraftThreads, 0, s>>>(P, k, probs_dev, idx_dev, cond_probs_dev); break;
  }
  LANTERN_CUDA(cudaGetLastError());
  return LANTERN_OK;
}
