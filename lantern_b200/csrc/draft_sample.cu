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

constexpr int kDraftGroups = 8;   // 4-column groups per thread: a row of up to 32 * NT window columns stays in registers

// One CTA per drafter row.  Single pass over the row: probabilities (written out as `op`), one Philox call per four
// columns, fp64 race keys kept in registers.  The k draws are then k block-wide argmins over (key, column) in which
// only the thread that owned the previous winner rescans its registers.
template <int DT, int NT>
__global__ void __launch_bounds__(NT) draft_sample_kernel(const AcceptParams P, int k, float* probs_out,
                                                          int32_t* idx_out, float* cond_out) {
  constexpr int NG = kDraftGroups, NE = NG * 4, NW = NT / 32;
  __shared__ double warp_key[NW];
  __shared__ int warp_col[NW];
  __shared__ int win_col;
  __shared__ float picked_p[64];
  const lantern_accept_cfg& cfg = P.cfg;
  const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, V = cfg.vocab;
  const RowStats st = P.stats[row];
  const int b = row / cfg.n_rows, t = row % cfg.n_rows;
  const int64_t base = (int64_t)b * cfg.item_stride + (int64_t)t * cfg.row_stride + cfg.col0;
  const MixParams mix = P.mix;
  const ExpShift ex(st.mx);
  const float inv = __fdiv_rn(1.0f, st.sum);
  float* prow = probs_out + (size_t)row * V;
  // columns outside the image-token window carry no mass
  for (int v = tid; v < cfg.col0; v += NT) prow[v] = 0.f;
  for (int v = cfg.col0 + cfg.ncols + tid; v < V; v += NT) prow[v] = 0.f;
  const bool vec_out = P.vec_ok && (reinterpret_cast<uintptr_t>(prow + cfg.col0) % 16 == 0);

  double key[NE];
  const uint32_t k0 = static_cast<uint32_t>(cfg.philox_seed), k1 = static_cast<uint32_t>(cfg.philox_seed >> 32);
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    const int e0 = (g * NT + tid) * 4;
    float c4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, u4[4] = {0.f, 0.f, 0.f, 0.f};
    if (e0 < cfg.ncols) {
      if (P.vec_ok) {
        Elem<DT>::load4(P.in.logits_cond, base + e0, c4);
        if (mix.has_uncond) Elem<DT>::load4(P.in.logits_uncond, base + e0, u4);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (e0 + j < cfg.ncols) {
            c4[j] = Elem<DT>::load1(P.in.logits_cond, base + e0 + j);
            if (mix.has_uncond) u4[j] = Elem<DT>::load1(P.in.logits_uncond, base + e0 + j);
          }
        }
      }
    }
    float p4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float sv = mix_temper(c4[j], u4[j], mix);
      p4[j] = (e0 + j < cfg.ncols && kept_col(sv, e0 + j, st)) ? __fmul_rn(ex(sv), inv) : 0.f;
    }
    if (e0 < cfg.ncols) {
      if (vec_out) {
        *reinterpret_cast<float4*>(prow + cfg.col0 + e0) = make_float4(p4[0], p4[1], p4[2], p4[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (e0 + j < cfg.ncols) prow[cfg.col0 + e0 + j] = p4[j];
      }
    }
    // race keys: u = ((word >> 8) + 1) * 2^-24 in (0, 1], word (col & 3) of the Philox block of counter col / 4
    uint32_t c[4] = {(uint32_t)(e0 >> 2), static_cast<uint32_t>(cfg.philox_step), (uint32_t)row,
                     static_cast<uint32_t>(cfg.philox_step >> 32) ^ 0x5A17u};
    const bool any = p4[0] > 0.f || p4[1] > 0.f || p4[2] > 0.f || p4[3] > 0.f;
    if (any) philox4x32_10(c, k0, k1);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double kk = INFINITY;
      if (p4[j] > 0.f) {
        const double u = ((double)(c[j] >> 8) + 1.0) * (1.0 / 16777216.0);
        kk = -log(u) / (double)p4[j];
      }
      key[g * 4 + j] = kk;
    }
  }

  // thread-local argmin; slots are visited in increasing column order, so `<` keeps the smallest column among ties
  auto local_best = [&](double& bk, int& bc) {
    bk = INFINITY;
    bc = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < NE; ++i) {
      if (key[i] < bk) { bk = key[i]; bc = ((i >> 2) * NT + tid) * 4 + (i & 3); }
    }
  };
  double my_key;
  int my_col;
  local_best(my_key, my_col);
  for (int r = 0; r < k; ++r) {
    double bk = my_key;
    int bc = my_col;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ok = __shfl_xor_sync(0xffffffffu, bk, o);
      const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
      if (ok < bk || (ok == bk && oc < bc)) { bk = ok; bc = oc; }
    }
    if (lane == 0) { warp_key[warp] = bk; warp_col[warp] = bc; }
    __syncthreads();
    if (warp == 0) {
      bk = lane < NW ? warp_key[lane] : INFINITY;
      bc = lane < NW ? warp_col[lane] : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ok = __shfl_xor_sync(0xffffffffu, bk, o);
        const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
        if (ok < bk || (ok == bk && oc < bc)) { bk = ok; bc = oc; }
      }
      if (lane == 0) win_col = bc;
    }
    __syncthreads();
    const int wc = win_col;
    const bool ok = wc != 0x7fffffff;   // fewer than k columns with positive probability: pad like torch would fail
    if (ok && wc == my_col) {           // the owner retires the winner, reports it and rescans its registers
#pragma unroll
      for (int i = 0; i < NE; ++i)
        if (((i >> 2) * NT + tid) * 4 + (i & 3) == wc) key[i] = INFINITY;
      idx_out[(size_t)row * k + r] = wc + cfg.col0;
      picked_p[r] = prow[wc + cfg.col0];
      local_best(my_key, my_col);
    } else if (!ok && tid == 0) {
      idx_out[(size_t)row * k + r] = -1;
      picked_p[r] = 0.f;
    }
  }
  __syncthreads();
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

template <int DT>
static int launch_draft(const AcceptParams& P, int k, float* probs, int32_t* idx, float* cond, unsigned rows,
                        cudaStream_t s) {
  const int ncols = P.cfg.ncols;
  if (ncols <= 4 * kDraftGroups * 128) draft_sample_kernel<DT, 128><<<rows, 128, 0, s>>>(P, k, probs, idx, cond);
  else if (ncols <= 4 * kDraftGroups * 256) draft_sample_kernel<DT, 256><<<rows, 256, 0, s>>>(P, k, probs, idx, cond);
  else if (ncols <= 4 * kDraftGroups * 512) draft_sample_kernel<DT, 512><<<rows, 512, 0, s>>>(P, k, probs, idx, cond);
  else if (ncols <= 4 * kDraftGroups * 1024) draft_sample_kernel<DT, 1024><<<rows, 1024, 0, s>>>(P, k, probs, idx, cond);
  else {
    set_error("lantern_draft_sample: ncols=%d exceeds %d", ncols, 4 * kDraftGroups * 1024);
    return LANTERN_E_UNSUPPORTED;
  }
  LANTERN_CUDA(cudaGetLastError());
  return LANTERN_OK;
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
    case LANTERN_F32: return launch_draft<LANTERN_F32>(P, k, probs_dev, idx_dev, cond_probs_dev, rows, s);
    case LANTERN_BF16: return launch_draft<LANTERN_BF16>(P, k, probs_dev, idx_dev, cond_probs_dev, rows, s);
    default: return launch_draft<LANTERN_F16>(P, k, probs_dev, idx_dev, cond_probs_dev, rows, s);
  }
}

