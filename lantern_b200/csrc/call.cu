// Single-prompt drop-in call: everything the reference's evaluate_posterior(logits, candidates, ...) needs around
// the fused kernels, behind ONE entry point, so that the Python shim pays one FFI call per verify step instead of a
// dozen tensor operations: upload of the step's uniforms, tree inputs rebuilt from `candidates` / `retrieve_indices`
// (int64 device tensors, as the reference holds them), the fused verify step (automatic schedule), read-back of the
// five result integers + the accepted path, one stream synchronisation.
//
// The context owns its small staging buffers (pinned host + device; allocated in lantern_call_create, the only place
// this file allocates) and is tied to the device that was current at creation.  One context per host thread.
#include <cstring>
#include <new>

#include "common.cuh"

struct lantern_call {
  int32_t max_rows = 0, max_cells = 0;
  float* h_uni = nullptr;      // pinned [max_rows + 1]
  int32_t* h_out = nullptr;    // pinned [5 + 2 * max_cells]
  float* d_uni = nullptr;
  int32_t *d_tokens = nullptr, *d_retrieve = nullptr, *d_out = nullptr;
};

using namespace lantern;

extern "C" void lantern_call_destroy(lantern_call* h) {
  if (!h) return;
  if (h->h_uni) cudaFreeHost(h->h_uni);
  if (h->h_out) cudaFreeHost(h->h_out);
  cudaFree(h->d_uni); cudaFree(h->d_tokens); cudaFree(h->d_retrieve); cudaFree(h->d_out);
  delete h;
}

extern "C" int lantern_call_create(int32_t max_rows, int32_t max_cells, lantern_call** out) {
  if (!out || max_rows < 1 || max_cells < 1) {
    set_error("lantern_call_create: bad argument");
    return LANTERN_E_INVALID;
  }
  lantern_call* h = new (std::nothrow) lantern_call();
  if (!h) return LANTERN_E_INVALID;
  h->max_rows = max_rows;
  h->max_cells = max_cells;
  cudaError_t e = cudaHostAlloc(&h->h_uni, (size_t)(max_rows + 1) * 4, cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaHostAlloc(&h->h_out, (size_t)(5 + 2 * max_cells) * 4, cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_uni, (size_t)(max_rows + 1) * 4);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_tokens, (size_t)max_rows * 4);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_retrieve, (size_t)max_cells * 4);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_out, (size_t)(5 + 2 * max_cells) * 4);
  if (e != cudaSuccess) {
    lantern_call_destroy(h);
    return cuda_fail(e, "lantern_call_create");
  }
  *out = h;
  return LANTERN_OK;
}

extern "C" float* lantern_call_uniforms(lantern_call* h) { return h ? h->h_uni : nullptr; }

extern "C" int lantern_posterior_call(lantern_call* h, const lantern_accept_cfg* cfg, const lantern_accept_in* in,
                                      const int64_t* cand_dev, const int64_t* retrieve_dev, int32_t n_uniforms,
                                      float* sample_p_dev, void* workspace_dev, size_t workspace_bytes,
                                      int32_t* out_host, void* stream) {
  if (!h || !cfg || !in || !cand_dev || !retrieve_dev || !out_host) {
    set_error("lantern_posterior_call: null argument");
    return LANTERN_E_INVALID;
  }
  const int T = cfg->n_rows, cells = cfg->n_paths * cfg->depth, D = cfg->depth;
  if (cfg->n_items != 1 || T > h->max_rows || cells > h->max_cells || n_uniforms < 0 || n_uniforms > h->max_rows + 1) {
    set_error("lantern_posterior_call: one prompt per call, tree of at most %d nodes / %d path cells (got %d / %d)",
              h->max_rows, h->max_cells, T, cells);
    return LANTERN_E_INVALID;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  lantern_accept_cfg c = *cfg;
  lantern_accept_in a = *in;
  if (n_uniforms > 0) {   // the caller filled lantern_call_uniforms(h)[0 .. n_uniforms)
    LANTERN_CUDA(cudaMemcpyAsync(h->d_uni, h->h_uni, (size_t)n_uniforms * 4, cudaMemcpyHostToDevice, s));
    a.uniforms = h->d_uni;
    c.n_uniforms = n_uniforms;
  } else {
    a.uniforms = nullptr;
  }
  int rc = lantern_tree_from_candidates(cand_dev, retrieve_dev, c.n_paths, c.depth, T, h->d_tokens, h->d_retrieve, s);
  if (rc != LANTERN_OK) return rc;
  a.tree_tokens = h->d_tokens;
  a.retrieve = h->d_retrieve;
  c.retrieve_shared = 0;
  lantern_accept_out o;
  memset(&o, 0, sizeof(o));
  o.accept_length = h->d_out;
  o.best_candidate = h->d_out + 1;
  o.token = h->d_out + 2;
  o.n_draws = h->d_out + 3;
  o.flags = h->d_out + 4;
  o.path_tokens = h->d_out + 5;
  o.select_indices = h->d_out + 5 + D;
  o.sample_p = sample_p_dev;
  rc = lantern_accept_phases(&c, &a, &o, workspace_dev, workspace_bytes, stream, 8);
  if (rc != LANTERN_OK) return rc;
  LANTERN_CUDA(cudaMemcpyAsync(h->h_out, h->d_out, (size_t)(5 + 2 * D) * 4, cudaMemcpyDeviceToHost, s));
  LANTERN_CUDA(cudaStreamSynchronize(s));
  memcpy(out_host, h->h_out, (size_t)(5 + 2 * D) * 4);
  return LANTERN_OK;
}
