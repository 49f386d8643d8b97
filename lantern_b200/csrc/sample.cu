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

}  // namespace lantern

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
