// Greedy verification (`logits_processor is None`, temperature 0) in tree form.
//
//   plain    drafters/utils.py:356-369, ea_model_anole.py:812-817: a child is accepted iff its token is the argmax of
//            its parent's logits row; accept length = longest accepted prefix over the leaf paths.
//   relaxed  ea_model_anole.py:789-902 (LANTERN, "TVD" form): gtp = softmax(row); px = gtp[x]; neighbour masses
//            np = gtp[table[x - off, :k] + off], cs = cumsum(np); approx = px + cs;
//            tvd = 0.5*|px - approx| + cumsum(0.5*np); idx = LAST position with tvd <= delta (or <= (delta-1)*px);
//            gtp[x] = approx[idx]; accepted iff argmax(gtp) == x (first index on ties).
//
// The reference evaluates every (path, level) position of the gathered [L, D, V] tensor; positions that share a tree
// node see the same row and the same token, so one CTA per tree node decides each edge once (greedy_node_kernel) and
// one CTA per prompt turns the per-node flags into path prefixes and the outputs (greedy_finish_kernel).
// The third return value of the reference, logits[best, accept_length], is the CFG-mixed (and, for Anole, masked) row
// of the last accepted node; it is written to out.sample_p when requested, and out.token is its argmax.
#include "accept_types.cuh"

namespace lantern {

constexpr int kGreedyThreads = 512;

struct ArgMax {
  float v;
  int i;
};
__device__ __forceinline__ ArgMax better(ArgMax a, ArgMax b) {   // larger value, then smaller index
  return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}
__device__ __forceinline__ ArgMax block_argmax(ArgMax a, float* sv, int* si) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ArgMax b;
    b.v = __shfl_xor_sync(0xffffffffu, a.v, o);
    b.i = __shfl_xor_sync(0xffffffffu, a.i, o);
    a = better(a, b);
  }
  __syncthreads();
  if (lane == 0) { sv[warp] = a.v; si[warp] = a.i; }
  __syncthreads();
  ArgMax r;
  r.v = sv[0]; r.i = si[0];
  for (int w = 1; w < nw; ++w) {
    ArgMax b;
    b.v = sv[w]; b.i = si[w];
    r = better(r, b);
  }
  __syncthreads();
  return r;
}

template <int DT>
__global__ void __launch_bounds__(kGreedyThreads) greedy_node_kernel(const AcceptParams P, int32_t* __restrict__ acc) {
  extern __shared__ __align__(16) float row[];   // [ncols] CFG-mixed logits of the parent's row
  __shared__ float sv[32];
  __shared__ int si[32];
  __shared__ double dscr[34];
  __shared__ float fscr[34];
  __shared__ int first_pos;
  const lantern_accept_cfg& cfg = P.cfg;
  const int c = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, NT = blockDim.x;
  const int L = cfg.n_paths, D = cfg.depth, T = cfg.n_rows, ncols = cfg.ncols, col0 = cfg.col0, off = cfg.tok_offset;
  int32_t* acc_b = acc + (size_t)b * T;
  if (c == 0) {   // the root is the token already emitted
    if (tid == 0) acc_b[0] = 1;
    return;
  }
  const int* ri = P.in.retrieve + (cfg.retrieve_shared ? 0 : (size_t)b * L * D);
  // first (path, level) position that reaches node c: gives the parent
  if (tid == 0) first_pos = 0x7fffffff;
  __syncthreads();
  for (int p = tid; p < L * D; p += NT)
    if (ri[p] == c && (p % D) > 0) atomicMin(&first_pos, p);
  __syncthreads();
  const int pos = first_pos;
  if (pos == 0x7fffffff) {   // no path reaches this node
    if (tid == 0) acc_b[c] = 0;
    return;
  }
  int parent = ri[pos - 1];
  if (parent < 0) parent += T;
  const int x = P.in.tree_tokens[(size_t)b * T + c];
  const int xe = x - col0;   // window-relative column of the child token

  MixParams mix = P.mix;
  mix.do_temp = 0;
  const int64_t base = (int64_t)b * cfg.item_stride + (int64_t)parent * cfg.row_stride + col0;
  ArgMax top;
  top.v = -INFINITY; top.i = 0x7fffffff;
  for (int e = tid; e < ncols; e += NT) {
    const float cv = Elem<DT>::load1(P.in.logits_cond, base + e);
    const float uv = mix.has_uncond ? Elem<DT>::load1(P.in.logits_uncond, base + e) : 0.f;
    const float s = mix_temper(cv, uv, mix);
    row[e] = s;
    if (s > top.v) { top.v = s; top.i = e; }   // increasing e: keeps the first index of the thread's maximum
  }
  top = block_argmax(top, sv, si);
  if (!cfg.lantern) {
    if (tid == 0) acc_b[c] = (xe == top.i) ? 1 : 0;
    return;
  }
  // ---- relaxed: probabilities, second maximum, neighbour aggregation ----
  const ExpShift ex(top.v);
  float part = 0.f;
  ArgMax second;
  second.v = -INFINITY; second.i = 0x7fffffff;
  for (int e = tid; e < ncols; e += NT) {
    const float s = row[e];
    part += ex(s);
    if (e != top.i && s > second.v) { second.v = s; second.i = e; }
  }
  const float tot = block_reduce(part, OpSum(), 0.f, fscr);
  second = block_argmax(second, sv, si);
  const float inv = __fdiv_rn(1.0f, tot);
  auto prob = [&](int e) -> float { return (e >= 0 && e < ncols) ? __fmul_rn(ex(row[e]), inv) : 0.f; };
  if (xe < 0 || xe >= ncols) {   // non-image token: probability 0 after the mask, no table row
    if (tid == 0) acc_b[c] = 0;
    return;
  }
  const float px = prob(xe);
  const int kk = min(cfg.lantern_k, cfg.table_cols);
  const int* nb_row = P.in.nbr_table + (size_t)(x - off) * cfg.table_cols;
  const float bound = cfg.lantern_delta > 1.0f ? __fmul_rn(cfg.lantern_delta_m1, px) : cfg.lantern_delta;
  int best_m = -1;          // last position whose tvd is within the bound, and approx there (per thread, then block)
  float best_approx = 0.f;
  double carry = 0.0;
  for (int base_t = 0; base_t < kk; base_t += NT) {
    const int m = base_t + tid;
    double v = 0.0;
    if (m < kk) v = (double)prob(__ldg(nb_row + m) + off - col0);
    double total;
    const double incl = carry + block_scan_incl(v, dscr, &total);
    if (m < kk) {
      const float cs = (float)incl;
      const float approx = __fadd_rn(px, cs);
      const float tvd = __fadd_rn(__fmul_rn(0.5f, fabsf(__fsub_rn(px, approx))), __fmul_rn(0.5f, cs));
      if (tvd <= bound) { best_m = m; best_approx = approx; }   // m grows with the chunk: the thread keeps its last hit
    }
    carry += total;
    __syncthreads();
  }
  ArgMax last;   // block-wide maximum of m (value field carries m, index field carries the approx bits)
  last.v = (float)best_m; last.i = __float_as_int(best_approx);
  {
    const int lane = tid & 31, warp = tid >> 5, nw = NT >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, last.v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, last.i, o);
      if (ov > last.v) { last.v = ov; last.i = oi; }
    }
    if (lane == 0) { sv[warp] = last.v; si[warp] = last.i; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < nw; ++w)
        if (sv[w] > last.v) { last.v = sv[w]; last.i = si[w]; }
      const float px_adj = last.v >= 0.f ? __int_as_float(last.i) : px;
      // argmax of the row with gtp[x] replaced: compare with the best other column (first index wins ties)
      const int oi = (top.i != xe) ? top.i : second.i;
      const float o = prob(oi);
      acc_b[c] = (px_adj > o || (px_adj == o && xe < oi)) ? 1 : 0;
    }
  }
}

template <int DT>
__global__ void __launch_bounds__(kGreedyThreads) greedy_finish_kernel(const AcceptParams P,
                                                                       const int32_t* __restrict__ acc) {
  __shared__ float sv[32];
  __shared__ int si[32];
  const lantern_accept_cfg& cfg = P.cfg;
  const int b = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
  const int L = cfg.n_paths, D = cfg.depth, T = cfg.n_rows, V = cfg.vocab;
  const int* ri = P.in.retrieve + (cfg.retrieve_shared ? 0 : (size_t)b * L * D);
  const int* tok = P.in.tree_tokens + (size_t)b * T;
  const int32_t* acc_b = acc + (size_t)b * T;
  // longest accepted prefix per path; first path on ties (torch.argmax), path 0 when nothing is accepted
  ArgMax best;
  best.v = -1.f; best.i = 0x7fffffff;
  for (int j = tid; j < L; j += NT) {
    int a = 0;
    for (int i = 1; i < D; ++i) {
      const int n = ri[j * D + i];
      if (n < 0 || !acc_b[n]) break;
      ++a;
    }
    if ((float)a > best.v) { best.v = (float)a; best.i = j; }
  }
  best = block_argmax(best, sv, si);
  const int a = (int)best.v;
  const int bj = a == 0 ? 0 : best.i;
  int node = ri[bj * D + a];
  if (node < 0) node += T;
  // the returned row: CFG-mixed logits of that node over the whole vocabulary (Anole: finfo.min outside the window)
  MixParams mix = P.mix;
  mix.do_temp = 0;
  const int64_t base = (int64_t)b * cfg.item_stride + (int64_t)node * cfg.row_stride;
  const bool masked = cfg.family == LANTERN_FAMILY_ANOLE;
  float* out_row = P.out.sample_p ? P.out.sample_p + (size_t)b * V : nullptr;
  ArgMax top;
  top.v = -INFINITY; top.i = 0x7fffffff;
  for (int v = tid; v < V; v += NT) {
    float s;
    if (masked && (v < cfg.col0 || v >= cfg.col0 + cfg.ncols)) {
      s = -3.4028234663852886e38f;   // torch.finfo(float32).min (ea_model_anole.py:931)
    } else {
      const float cv = Elem<DT>::load1(P.in.logits_cond, base + v);
      const float uv = mix.has_uncond ? Elem<DT>::load1(P.in.logits_uncond, base + v) : 0.f;
      s = mix_temper(cv, uv, mix);
    }
    if (out_row) out_row[v] = s;
    if (s > top.v) { top.v = s; top.i = v; }
  }
  top = block_argmax(top, sv, si);
  if (tid == 0) {
    P.out.accept_length[b] = a;
    P.out.best_candidate[b] = bj;
    P.out.token[b] = top.i == 0x7fffffff ? 0 : top.i;
    if (P.out.n_draws) P.out.n_draws[b] = 0;
    if (P.out.flags) P.out.flags[b] = 0;
  }
  for (int i = tid; i < D; i += NT) {
    const int n = ri[bj * D + i];
    if (P.out.path_tokens) P.out.path_tokens[(size_t)b * D + i] = (i <= a && n >= 0) ? tok[n] : -1;
    if (P.out.select_indices) P.out.select_indices[(size_t)b * D + i] = i <= a ? n : -1;
  }
}

template <int DT>
static int launch_greedy(const AcceptParams& P, int32_t* acc, cudaStream_t s) {
  const lantern_accept_cfg& c = P.cfg;
  const size_t smem = (size_t)c.ncols * 4;
  if (smem > 200 * 1024) {
    set_error("lantern_accept_greedy: window of %d columns exceeds shared memory", c.ncols);
    return LANTERN_E_UNSUPPORTED;
  }
  auto k1 = greedy_node_kernel<DT>;
  if (smem > 48 * 1024) LANTERN_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k1<<<dim3((unsigned)c.n_rows, (unsigned)c.n_items), kGreedyThreads, smem, s>>>(P, acc);
  LANTERN_CUDA(cudaGetLastError());
  greedy_finish_kernel<DT><<<(unsigned)c.n_items, kGreedyThreads, 0, s>>>(P, acc);
  LANTERN_CUDA(cudaGetLastError());
  return LANTERN_OK;
}

}  // namespace lantern

using namespace lantern;

extern "C" LANTERN_API size_t lantern_accept_greedy_workspace_bytes(const lantern_accept_cfg* cfg) {
  return cfg ? (size_t)cfg->n_items * (size_t)cfg->n_rows * sizeof(int32_t) : 0;
}

extern "C" LANTERN_API int lantern_accept_greedy(const lantern_accept_cfg* cfg, const lantern_accept_in* in,
                                                 const lantern_accept_out* out, void* workspace_dev,
                                                 size_t workspace_bytes, void* stream) {
  if (!cfg || !in || !out) {
    set_error("lantern_accept_greedy: null argument");
    return LANTERN_E_INVALID;
  }
  const lantern_accept_cfg& c = *cfg;
  if (c.n_items <= 0 || c.n_rows <= 0 || c.n_paths <= 0 || c.depth <= 0 || c.vocab <= 0 || c.ncols <= 0 || c.col0 < 0 ||
      c.col0 + c.ncols > c.vocab || c.row_stride < c.vocab || c.logits_dtype < LANTERN_F32 ||
      c.logits_dtype > LANTERN_F16 || !in->logits_cond || !in->tree_tokens || !in->retrieve || !out->accept_length ||
      !out->best_candidate || !out->token) {
    set_error("lantern_accept_greedy: bad shape / window / dtype, or a required pointer is null (rows must span the vocabulary)");
    return LANTERN_E_INVALID;
  }
  if (c.family == LANTERN_FAMILY_LUMINA) {   // the reference raises NotImplementedError (ea_model_lumina_mgpt.py:728-729)
    set_error("lantern_accept_greedy: greedy decoding is not defined for the Lumina-mGPT family");
    return LANTERN_E_UNSUPPORTED;
  }
  if (c.lantern && (!in->nbr_table || c.lantern_k < 1 || c.lantern_k > c.table_cols)) {
    set_error("lantern_accept_greedy: lantern=1 needs nbr_table and 1 <= lantern_k <= table_cols");
    return LANTERN_E_INVALID;
  }
  if (!workspace_dev || workspace_bytes < lantern_accept_greedy_workspace_bytes(cfg)) {
    set_error("lantern_accept_greedy: workspace too small");
    return LANTERN_E_WORKSPACE;
  }
  AcceptParams P;
  memset(&P, 0, sizeof(P));
  P.cfg = *cfg;
  P.in = *in;
  P.out = *out;
  P.mix.cfg_scale = cfg->cfg_scale;
  P.mix.temperature = 1.0f;
  P.mix.has_uncond = in->logits_uncond != nullptr;
  P.mix.do_temp = 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int32_t* acc = static_cast<int32_t*>(workspace_dev);
  switch (c.logits_dtype) {
    case LANTERN_F32: return launch_greedy<LANTERN_F32>(P, acc, s);
    case LANTERN_BF16: return launch_greedy<LANTERN_BF16>(P, acc, s);
    default: return launch_greedy<LANTERN_F16>(P, acc, s);
  }
}
