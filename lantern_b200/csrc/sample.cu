// Bonus-token draw from caller-supplied probability rows: token = min{i : cdf_i > u * total},
// fp64 accumulation in index order.  Replaces torch.multinomial(prob, 1) of update_inference_inputs
// (ea_model_llamagen.py:978, ea_model_lumina_mgpt.py:781) when sample_p did not come from the fused step.
#include "common.cuh"

namespace lantern {

constexpr int kSampleThreads = 512;

__global__ void __launch_bounds__(kSampleThreads) sample_tokens_kernel(const float* __restrict__ probs,
                                                                       int64_t row_stride, int vocab,
                                                                       const float* __restrict__ uniforms,
                                                                       int32_t* __restrict__ tokens) {
  __shared__ double dscr[34];
  __shared__ int found, last_nz;
  const float* p = probs + (int64_t)blockIdx.x * row_stride;
  const int tid = threadIdx.x, NT = blockDim.x;
  const int per = (vocab + NT - 1) / NT;
  const int i0 = min(tid * per, vocab), i1 = min(i0 + per, vocab);
  double loc = 0.0;
  for (int i = i0; i < i1; ++i) loc += (double)p[i];
  double total;
  const double incl = block_scan_incl(loc, dscr, &total);
  const double target = (double)uniforms[blockIdx.x] * total;
  if (tid == 0) { found = 0x7fffffff; last_nz = -1; }
  __syncthreads();
  double run = incl - loc;
  int lnz = -1, hit = -1;
  const bool mine = target >= run && target < run + loc;
  for (int i = i0; i < i1; ++i) {
    const float v = p[i];
    if (v > 0.f) lnz = i;
    run += (double)v;
    if (mine && hit < 0 && run > target) hit = i;
  }
  if (hit >= 0) atomicMin(&found, hit);
  if (lnz >= 0) atomicMax(&last_nz, lnz);
  __syncthreads();
  if (tid == 0) tokens[blockIdx.x] = found != 0x7fffffff ? found : (last_nz >= 0 ? last_nz : 0);
}

// The reference's evaluate_posterior receives `candidates [L, D]` (token of node retrieve_indices[j, i], -1 padded,
// int64) rather than the tree's token vector; this rebuilds the kernel-side inputs in one launch:
// tree_tokens[node] (int32, 0 for nodes no path reaches) and the int32 copy of retrieve_indices.
__global__ void tree_from_candidates_kernel(const int64_t* __restrict__ cand, const int64_t* __restrict__ ri, int n,
                                            int T, int32_t* __restrict__ tokens, int32_t* __restrict__ ri32) {
  // tokens was zeroed by the memset the host enqueued before this launch
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int64_t node = ri[i];
    ri32[i] = (node >= 0 && node < T) ? (int32_t)node : -1;   // out-of-range entries become padding (never indexed)
    if (node >= 0 && node < T) tokens[node] = (int32_t)cand[i];   // every path through a node carries the same token
  }
}

}  // namespace lantern

extern "C" int lantern_tree_from_candidates(const int64_t* cand_dev, const int64_t* retrieve_dev, int32_t n_paths,
                                            int32_t depth, int32_t n_rows, int32_t* tokens_dev, int32_t* retrieve32_dev,
                                            void* stream) {
  using namespace lantern;
  if (!cand_dev || !retrieve_dev || !tokens_dev || !retrieve32_dev || n_paths <= 0 || depth <= 0 || n_rows <= 0) {
    set_error("lantern_tree_from_candidates: bad argument");
    return LANTERN_E_INVALID;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  LANTERN_CUDA(cudaMemsetAsync(tokens_dev, 0, (size_t)n_rows * 4, s));
  const int n = n_paths * depth;
  tree_from_candidates_kernel<<<(n + 255) / 256, 256, 0, s>>>(cand_dev, retrieve_dev, n, n_rows, tokens_dev,
                                                            retrieve32_dev);
  LANTERN_CUDA(cudaGetLastError());
  return LANTERN_OK;
}

extern "C" int lantern_sample_tokens(const float* probs_dev, int64_t row_stride, int32_t n_rows, int32_t vocab,
                                     const float* uniforms_dev, int32_t* tokens_dev, void* stream) {
  using namespace lantern;
  if (!probs_dev || !uniforms_dev || !tokens_dev || n_rows <= 0 || vocab <= 0 || row_stride < vocab) {
    set_error("lantern_sample_tokens: bad argument");
    return LANTERN_E_INVALID;
  }
  sample_tokens_kernel<<<n_rows, kSampleThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      probs_dev, row_stride, vocab, uniforms_dev, tokens_dev);
  LANTERN_CUDA(cudaGetLastError());
  return LANTERN_OK;
}
