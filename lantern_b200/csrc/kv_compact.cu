// KV-cache compaction of update_inference_inputs (ea_model_llamagen.py:962-970): for every
// (slab, outer) pair move the accepted positions select[b, 0:n_keep[b]] to prev_len[b]...  One launch for all
// slabs, layers, heads; per-item lengths live on the device so ragged batches need no host sync.
// Source and destination ranges may overlap (select[i] >= prev_len + i), so a CTA reads all its rows
// before it writes any (the reference gathers into a temporary first, :963-967).
#include "common.cuh"

namespace lantern {

constexpr int kKvThreads = 128;
constexpr int kKvMaxPerThread = 16;

__global__ void __launch_bounds__(kKvThreads) kv_compact_kernel(const lantern_kv_cfg cfg, void* const* slabs,
                                                                const void* __restrict__ select_raw,
                                                                const int32_t* __restrict__ prev_len,
                                                                const int32_t* __restrict__ n_keep) {
  const int64_t outer = blockIdx.x;
  char* slab = static_cast<char*>(slabs ? slabs[blockIdx.y] : cfg.slab0);
  const int b = (int)((outer / cfg.outer_per_batch) % cfg.n_batch);
  const int keep = n_keep ? n_keep[b] : cfg.n_keep0, prev = prev_len ? prev_len[b] : cfg.prev_len0;
  const int32_t* select = static_cast<const int32_t*>(select_raw);
  const int64_t* select64 = static_cast<const int64_t*>(select_raw);
  const int row_bytes = cfg.head_dim * cfg.elem_bytes;
  const int vpr = row_bytes / 16;   // 16-byte vectors per position
  const int work = keep * vpr;
  char* base = slab + outer * (int64_t)cfg.s_max * row_bytes;
  uint4 regs[kKvMaxPerThread];
  int n = 0;
  for (int w = threadIdx.x; w < work && n < kKvMaxPerThread; w += kKvThreads, ++n) {
    const int i = w / vpr, v = w % vpr;
    const int src = cfg.select_i64 ? (int)select64[b * cfg.max_keep + i] : select[b * cfg.max_keep + i];
    regs[n] = *reinterpret_cast<const uint4*>(base + (int64_t)src * row_bytes + v * 16);
  }
  __syncthreads();
  n = 0;
  for (int w = threadIdx.x; w < work && n < kKvMaxPerThread; w += kKvThreads, ++n) {
    const int i = w / vpr, v = w % vpr;
    *reinterpret_cast<uint4*>(base + (int64_t)(prev + i) * row_bytes + v * 16) = regs[n];
  }
}

}  // namespace lantern

extern "C" int lantern_kv_compact(const lantern_kv_cfg* cfg, void* const* slab_ptrs_dev, const void* select_dev,
                                  const int32_t* prev_len_dev, const int32_t* n_keep_dev, void* stream) {
  using namespace lantern;
  if (!cfg || !select_dev || (!slab_ptrs_dev && (!cfg->slab0 || cfg->n_slabs != 1))) {
    set_error("lantern_kv_compact: null argument (slab_ptrs_dev may be NULL only with n_slabs == 1 and cfg.slab0 set)");
    return LANTERN_E_INVALID;
  }
  if (!n_keep_dev && (cfg->n_keep0 < 0 || cfg->n_keep0 > cfg->max_keep)) {
    set_error("lantern_kv_compact: n_keep0 outside [0, max_keep]");
    return LANTERN_E_INVALID;
  }
  const int row_bytes = cfg->head_dim * cfg->elem_bytes;
  if (cfg->n_slabs <= 0 || cfg->n_outer <= 0 || cfg->n_batch <= 0 || cfg->outer_per_batch <= 0 || row_bytes % 16 != 0 ||
      cfg->max_keep <= 0) {
    set_error("lantern_kv_compact: bad config (head_dim * elem_bytes must be a multiple of 16)");
    return LANTERN_E_INVALID;
  }
  if ((int64_t)cfg->max_keep * (row_bytes / 16) > (int64_t)kKvThreads * kKvMaxPerThread) {
    set_error("lantern_kv_compact: max_keep * head_dim too large for one CTA");
    return LANTERN_E_UNSUPPORTED;
  }
  if (cfg->n_outer > 0x7fffffffLL || cfg->n_slabs > 65535) {
    set_error("lantern_kv_compact: grid too large");
    return LANTERN_E_UNSUPPORTED;
  }
  dim3 grid((unsigned)cfg->n_outer, (unsigned)cfg->n_slabs);
  kv_compact_kernel<<<grid, kKvThreads, 0, static_cast<cudaStream_t>(stream)>>>(*cfg, slab_ptrs_dev, select_dev,
                                                                               prev_len_dev, n_keep_dev);
  LANTERN_CUDA(cudaGetLastError());
  return LANTERN_OK;
}
