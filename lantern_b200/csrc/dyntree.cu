// Dynamic (EAGLE-2) draft-tree post-processing on the device — the host-side tail of the drafter's topK_genrate
// (models/drafters/cnets_llamagen.py:831-908, cnets_lumina_mgpt.py:1330-1393): global top `total_tokens` of the
// cumulative scores, flat-index order, parent map (searchsorted), ancestor mask, depths, and the leaf-path table
// `retrieve_indices`.  The reference does this with Python loops and .tolist() round trips every verify step; here it
// is one CTA per prompt and its outputs feed lantern_accept_fused directly (parent pointers, int32, -1 padded).
// Tie rule (torch.topk leaves it unspecified): higher score first, then lower flat index.
#include "common.cuh"

namespace lantern {

constexpr int kTreeThreads = 256;

struct DynTreeParams {
  const float* scores;      // [B, n_cand] cumulative log-prob of every drafted candidate, flat order
  const int32_t* tokens;    // [B, n_cand]
  const int32_t* parents;   // [B, n_groups] parent flat id + 1 of every sibling group (0 = root)
  const int32_t* root_tok;  // [B] sample_token
  int32_t* tree_tokens;     // [B, T]
  int32_t* parent;          // [B, T]
  int32_t* depth;           // [B, T]  (tree_position_ids)
  float* mask;              // [B, T, T] ancestor-or-self mask (tree_mask) or NULL
  int32_t* retrieve;        // [B, T, D_max]  -1 padded, rows beyond n_leaves all -1
  int32_t* counts;          // [B, 2] n_leaves, max_depth + 1
  int n_items, n_cand, n_groups, top_k, T, d_max, sort_rows;
};

__global__ void __launch_bounds__(kTreeThreads) dyntree_kernel(const DynTreeParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x, tid = threadIdx.x, n = P.n_cand, T = P.T, keep = P.T - 1, DM = P.d_max;
  float* sc = reinterpret_cast<float*>(smem_raw);                // [n]
  int* pos = reinterpret_cast<int*>(sc + n);                     // [n + 1] kept-count prefix (exclusive)
  int* par = pos + n + 1;                                        // [T]
  int* dep = par + T;                                            // [T]
  int* haschild = dep + T;                                       // [T]
  int* leaf_of = haschild + T;                                   // [T] row id of each leaf (or -1)
  int* rows = leaf_of + T;                                       // [T * DM]
  int* rank = rows + T * DM;                                     // [T] sorted position of each row
  __shared__ int iscr[40];
  __shared__ int n_leaves_s, max_depth_s;
  const float* s_g = P.scores + (size_t)b * n;
  for (int i = tid; i < n; i += kTreeThreads) sc[i] = s_g[i];
  __syncthreads();
  // ---- global top-(T-1): rank by (score desc, flat index asc) ----
  for (int i = tid; i < n; i += kTreeThreads) {
    const float si = sc[i];
    int r = 0;
    for (int j = 0; j < n; ++j) {
      const float sj = sc[j];
      r += (sj > si) || (sj == si && j < i);
    }
    pos[i] = r < keep ? 1 : 0;
  }
  __syncthreads();
  // exclusive prefix of the kept flags (n <= a few thousand: chunked serial scan + block scan of chunk sums)
  {
    const int per = (n + kTreeThreads - 1) / kTreeThreads;
    const int i0 = min(tid * per, n), i1 = min(i0 + per, n);
    int loc = 0;
    for (int i = i0; i < i1; ++i) loc += pos[i];
    // block exclusive scan of loc
    const int lane = tid & 31, warp = tid >> 5;
    int v = loc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (lane == 31) iscr[warp] = v;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < warp; ++w) base += iscr[w];
    int run = base + v - loc;
    for (int i = i0; i < i1; ++i) {
      const int f = pos[i];
      pos[i] = f ? run : -(run + 1);   // kept: its 0-based rank among kept; dropped: -(kept before it) - 1
      run += f;
    }
    if (tid == kTreeThreads - 1) pos[n] = run;
  }
  __syncthreads();
  auto kept_before = [&](int flat) -> int {   // number of kept candidates with flat index < flat (searchsorted left)
    if (flat >= n) return pos[n];
    const int p = pos[flat];
    return p >= 0 ? p : -(p + 1);
  };
  // ---- nodes: token, parent ----
  int32_t* tok_o = P.tree_tokens + (size_t)b * T;
  if (tid == 0) { tok_o[0] = P.root_tok[b]; par[0] = -1; dep[0] = 0; }
  for (int i = tid; i < T; i += kTreeThreads) { haschild[i] = 0; leaf_of[i] = -1; if (i > 0) dep[i] = -1; }
  __syncthreads();
  for (int i = tid; i < n; i += kTreeThreads) {
    if (pos[i] < 0) continue;
    const int node = pos[i] + 1;
    tok_o[node] = P.tokens[(size_t)b * n + i];
    const int pf = P.parents[(size_t)b * P.n_groups + i / P.top_k];   // parent flat id + 1, 0 = root
    const int pn = pf == 0 ? 0 : kept_before(pf - 1) + 1;
    par[node] = pn;
    haschild[pn] = 1;   // benign race: all writers store 1
  }
  __syncthreads();
  // ---- depths: every node walks its own parent chain (at most d_max steps; no shared writes but its own slot) ----
  for (int i = tid + 1; i < T; i += kTreeThreads) {
    int d = 0, a = i;
    while (a > 0 && d <= T) { a = par[a]; ++d; }
    dep[i] = d;
  }
  __syncthreads();
  for (int i = tid; i < T; i += kTreeThreads) {
    P.parent[(size_t)b * T + i] = par[i];
    P.depth[(size_t)b * T + i] = dep[i];
  }
  // ---- ancestor-or-self mask ----
  if (P.mask) {
    float* m = P.mask + (size_t)b * T * T;
    for (int i = tid; i < T * T; i += kTreeThreads) m[i] = (i % T == 0) ? 1.f : 0.f;
    __syncthreads();
    for (int i = tid; i < T; i += kTreeThreads) {
      int a = i;
      while (a > 0) { m[(size_t)i * T + a] = 1.f; a = par[a]; }
    }
  }
  // ---- leaves in node order -> rows ----
  if (tid == 0) {
    int nl = 0, md = 0;
    for (int i = 0; i < T; ++i) {
      if (!haschild[i] && (i > 0 || T == 1)) leaf_of[i] = nl++;
      md = max(md, dep[i]);
    }
    n_leaves_s = nl;
    max_depth_s = md;
  }
  __syncthreads();
  const int nl = n_leaves_s, D = max_depth_s + 1;
  for (int i = tid; i < T * DM; i += kTreeThreads) rows[i] = -1;
  __syncthreads();
  for (int i = tid; i < T; i += kTreeThreads) {
    const int rid = leaf_of[i];
    if (rid < 0) continue;
    int c = i;
    for (int j = dep[i]; j >= 0; --j) {   // a tree deeper than d_max is reported through counts below, never written
      if (j < DM) rows[rid * DM + j] = c;
      c = par[c];
    }
  }
  __syncthreads();
  // ---- lexicographic row order with padding last (cnets_llamagen.py:896-906) ----
  for (int r = tid; r < nl; r += kTreeThreads) {
    int rk = 0;
    if (P.sort_rows) {
      for (int q = 0; q < nl; ++q) {
        if (q == r) continue;
        int cmp = 0;   // -1: q < r
        for (int j = 0; j < min(D, DM) && cmp == 0; ++j) {
          const int a = rows[q * DM + j] >= 0 ? rows[q * DM + j] : T + 5;
          const int bb = rows[r * DM + j] >= 0 ? rows[r * DM + j] : T + 5;
          cmp = a < bb ? -1 : (a > bb ? 1 : 0);
        }
        rk += (cmp < 0) || (cmp == 0 && q < r);
      }
    } else {
      rk = r;
    }
    rank[r] = rk;
  }
  __syncthreads();
  int32_t* ri = P.retrieve + (size_t)b * T * DM;
  for (int i = tid; i < T * DM; i += kTreeThreads) ri[i] = -1;
  __syncthreads();
  for (int i = tid; i < nl * DM; i += kTreeThreads) {
    const int r = i / DM, j = i % DM;
    ri[rank[r] * DM + j] = rows[r * DM + j];
  }
  // max_depth + 1 > d_max: the path table does not fit; n_leaves = -1 tells the caller (the rows above are truncated)
  if (tid == 0) { P.counts[b * 2] = D <= DM ? nl : -1; P.counts[b * 2 + 1] = D; }
}

}  // namespace lantern

extern "C" LANTERN_API int lantern_build_dynamic_tree(const float* scores_dev, const int32_t* tokens_dev,
                                                      const int32_t* parents_dev, const int32_t* root_tokens_dev,
                                                      int32_t n_items, int32_t n_cand, int32_t n_groups, int32_t top_k,
                                                      int32_t total_tokens, int32_t d_max, int32_t sort_rows,
                                                      int32_t* tree_tokens_dev, int32_t* parent_dev, int32_t* depth_dev,
                                                      float* mask_dev, int32_t* retrieve_dev, int32_t* counts_dev,
                                                      void* stream) {
  using namespace lantern;
  if (!scores_dev || !tokens_dev || !parents_dev || !root_tokens_dev || !tree_tokens_dev || !parent_dev || !depth_dev ||
      !retrieve_dev || !counts_dev || n_items <= 0 || n_cand <= 0 || top_k <= 0 || total_tokens < 1 || d_max < 1 ||
      total_tokens - 1 > n_cand || n_groups * top_k < n_cand) {
    set_error("lantern_build_dynamic_tree: bad argument");
    return LANTERN_E_INVALID;
  }
  DynTreeParams P{scores_dev, tokens_dev, parents_dev, root_tokens_dev, tree_tokens_dev, parent_dev, depth_dev, mask_dev,
                  retrieve_dev, counts_dev, n_items, n_cand, n_groups, top_k, total_tokens, d_max, sort_rows};
  const size_t smem = (size_t)n_cand * 4 + (size_t)(n_cand + 1) * 4 + (size_t)total_tokens * 4 * 5 +
                      (size_t)total_tokens * d_max * 4 + 64;
  if (smem > 200 * 1024) {
    set_error("lantern_build_dynamic_tree: tree too large for shared memory");
    return LANTERN_E_UNSUPPORTED;
  }
  if (smem > 48 * 1024)
    LANTERN_CUDA(cudaFuncSetAttribute(dyntree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dyntree_kernel<<<n_items, kTreeThreads, smem, static_cast<cudaStream_t>(stream)>>>(P);
  LANTERN_CUDA(cudaGetLastError());
  return LANTERN_OK;
}
