// Host-buffer session: the end-to-end call for a caller whose logits live in host memory (the reference's
// CPU path, or a reference-side integration that has not moved its tensors to the new device API yet).
// Owns device buffers, pinned result staging and a stream.  Each step copies only the live column window of
// the logits rows host->device (strided 2-D DMA), runs lantern_accept_fused and copies the per-item results back.
#include <cstdlib>
#include <cstring>
#include <new>

#include "common.cuh"

struct lantern_session {
  lantern_accept_cfg cfg;      // creation-time sizes (upper bounds for every step)
  cudaStream_t stream = nullptr;
  int c_lo = 0, width = 0;     // device rows hold columns [c_lo, c_lo + width)
  int table_rows = 0;
  // device inputs
  char *d_cond = nullptr, *d_uncond = nullptr;
  int32_t *d_tokens = nullptr, *d_retrieve = nullptr, *d_table = nullptr;
  uint8_t* d_kinds = nullptr;
  float* d_uniforms = nullptr;
  float *d_node_q = nullptr, *d_op = nullptr;
  int32_t *d_qrow = nullptr, *d_sib_off = nullptr, *d_sib_idx = nullptr, *d_sib_tokens = nullptr;
  size_t sib_idx_cap = 0, sib_tok_cap = 0;
  // device outputs (one contiguous int block + optional sample_p)
  int32_t* d_out = nullptr;
  int32_t* h_out = nullptr;    // pinned mirror
  float* d_sample_p = nullptr;
  void* d_work = nullptr;
  size_t work_bytes = 0;
  int last_in_place = 0;       // route of the last step (lantern_session_last_route)
};

namespace lantern {

static size_t elem_bytes(int dt) { return dt == LANTERN_F32 ? 4 : 2; }
static size_t out_ints(const lantern_accept_cfg& c) { return (size_t)c.n_items * (5 + 2 * (size_t)c.depth); }

}  // namespace lantern

using namespace lantern;

extern "C" void lantern_session_destroy(lantern_session* s) {
  if (!s) return;
  cudaFree(s->d_cond); cudaFree(s->d_uncond); cudaFree(s->d_tokens); cudaFree(s->d_retrieve);
  cudaFree(s->d_table); cudaFree(s->d_kinds); cudaFree(s->d_uniforms); cudaFree(s->d_node_q);
  cudaFree(s->d_op); cudaFree(s->d_qrow); cudaFree(s->d_sib_off); cudaFree(s->d_sib_idx);
  cudaFree(s->d_sib_tokens); cudaFree(s->d_out); cudaFree(s->d_sample_p); cudaFree(s->d_work);
  if (s->h_out) cudaFreeHost(s->h_out);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
}

extern "C" int lantern_session_create(const lantern_accept_cfg* cfg, const int32_t* nbr_table_host,
                                      int32_t table_rows, lantern_session** out) {
  if (!cfg || !out) {
    set_error("lantern_session_create: null argument");
    return LANTERN_E_INVALID;
  }
  if (cfg->lantern && (!nbr_table_host || table_rows <= 0)) {
    set_error("lantern_session_create: lantern=1 needs the neighbour table");
    return LANTERN_E_INVALID;
  }
  lantern_session* s = new (std::nothrow) lantern_session();
  if (!s) return LANTERN_E_INVALID;
  s->cfg = *cfg;
  const lantern_accept_cfg& c = s->cfg;
  const size_t eb = elem_bytes(c.logits_dtype);
  s->c_lo = c.col0 & ~7;
  int c_hi = (c.col0 + c.ncols + 7) & ~7;
  if (c_hi > c.vocab) c_hi = c.vocab;
  s->width = (c_hi - s->c_lo + 3) & ~3;
  if (s->c_lo + s->width > c.row_stride) s->width = (int)c.row_stride - s->c_lo;
  s->table_rows = table_rows;
  const size_t rows = (size_t)c.n_items * c.n_rows;
  int rc = LANTERN_OK;
#define TRY(expr)                                   \
  do {                                              \
    cudaError_t _e = (expr);                        \
    if (_e != cudaSuccess) {                        \
      rc = cuda_fail(_e, #expr);                    \
      lantern_session_destroy(s);                   \
      return rc;                                    \
    }                                               \
  } while (0)
  TRY(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  TRY(cudaMalloc(&s->d_cond, rows * s->width * eb));
  TRY(cudaMalloc(&s->d_uncond, rows * s->width * eb));
  TRY(cudaMalloc(&s->d_tokens, rows * 4));
  TRY(cudaMalloc(&s->d_retrieve, (size_t)c.n_items * c.n_paths * c.depth * 4));
  TRY(cudaMalloc(&s->d_kinds, rows));
  TRY(cudaMalloc(&s->d_uniforms, (size_t)c.n_items * (c.n_uniforms > 0 ? c.n_uniforms : 1) * 4));
  if (c.lantern) {
    TRY(cudaMalloc(&s->d_table, (size_t)table_rows * c.table_cols * 4));
    TRY(cudaMemcpy(s->d_table, nbr_table_host, (size_t)table_rows * c.table_cols * 4, cudaMemcpyHostToDevice));
  }
  if (c.static_tree) {
    TRY(cudaMalloc(&s->d_node_q, rows * 4));
    TRY(cudaMalloc(&s->d_op, (size_t)c.n_items * c.n_q_rows * c.vocab * 4));
    TRY(cudaMalloc(&s->d_qrow, (size_t)c.n_rows * 4));
    TRY(cudaMalloc(&s->d_sib_off, ((size_t)c.n_rows + 1) * 4));
  }
  TRY(cudaMalloc(&s->d_out, out_ints(c) * 4));
  TRY(cudaMallocHost(&s->h_out, out_ints(c) * 4));
  s->work_bytes = lantern_accept_workspace_bytes(&c);
  TRY(cudaMalloc(&s->d_work, s->work_bytes));
#undef TRY
  *out = s;
  return LANTERN_OK;
}

// 1 if the last step read the logits in place from page-locked host memory, 0 if it staged them on the device.
extern "C" int lantern_session_last_route(const lantern_session* s) { return s ? s->last_in_place : -1; }

extern "C" int lantern_session_step(lantern_session* s, const lantern_accept_cfg* cfg, const lantern_accept_in* in,
                                    const lantern_accept_out* out) {
  if (!s || !cfg || !in || !out) {
    set_error("lantern_session_step: null argument");
    return LANTERN_E_INVALID;
  }
  const lantern_accept_cfg& m = s->cfg;
  if (cfg->n_items > m.n_items || cfg->n_rows > m.n_rows || cfg->n_paths * cfg->depth > m.n_paths * m.depth ||
      cfg->depth > m.depth || cfg->vocab != m.vocab || cfg->col0 != m.col0 || cfg->ncols != m.ncols ||
      cfg->logits_dtype != m.logits_dtype || cfg->table_cols != m.table_cols || cfg->static_tree != m.static_tree ||
      cfg->n_q_rows > m.n_q_rows || cfg->n_uniforms > (m.n_uniforms > 0 ? m.n_uniforms : 1) ||
      (cfg->lantern && !m.lantern)) {
    set_error("lantern_session_step: step config exceeds the sizes the session was created with");
    return LANTERN_E_INVALID;
  }
  if (!in->logits_cond || !in->tree_tokens || !in->retrieve || !out->accept_length || !out->best_candidate ||
      !out->token) {
    set_error("lantern_session_step: logits_cond/tree_tokens/retrieve and the three scalar outputs are required");
    return LANTERN_E_INVALID;
  }
  cudaStream_t st = s->stream;
  const size_t eb = elem_bytes(cfg->logits_dtype);
  const int B = cfg->n_items, T = cfg->n_rows, L = cfg->n_paths, D = cfg->depth;
  const size_t rows = (size_t)B * T;
  const size_t dpitch = (size_t)s->width * eb, spitch = (size_t)cfg->row_stride * eb;
  auto copy_rows = [&](char* dst, const void* src) -> cudaError_t {
    const char* sp = static_cast<const char*>(src) + (size_t)s->c_lo * eb;
    if (cfg->item_stride == (int64_t)T * cfg->row_stride)
      return cudaMemcpy2DAsync(dst, dpitch, sp, spitch, dpitch, rows, cudaMemcpyHostToDevice, st);
    for (int b = 0; b < B; ++b) {
      cudaError_t e = cudaMemcpy2DAsync(dst + (size_t)b * T * dpitch, dpitch, sp + (size_t)b * cfg->item_stride * eb,
                                        spitch, dpitch, T, cudaMemcpyHostToDevice, st);
      if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
  };
  // In-place route: page-locked logits the device can address are read by the lazy walk directly (zero-copy).
  const void *dev_cond = nullptr, *dev_uncond = nullptr;
  bool in_place = getenv("LANTERN_SESSION_STAGED") == nullptr;
  if (in_place) {
    auto mapped = [](const void* host, const void** dev) -> bool {
      cudaPointerAttributes a;
      if (cudaPointerGetAttributes(&a, host) != cudaSuccess) { cudaGetLastError(); return false; }
      if (a.type != cudaMemoryTypeHost || !a.devicePointer) return false;
      *dev = a.devicePointer;
      return true;
    };
    in_place = mapped(in->logits_cond, &dev_cond) && (!in->logits_uncond || mapped(in->logits_uncond, &dev_uncond));
  }
  s->last_in_place = 0;
  if (!in_place) {
    LANTERN_CUDA(copy_rows(s->d_cond, in->logits_cond));
    if (in->logits_uncond) LANTERN_CUDA(copy_rows(s->d_uncond, in->logits_uncond));
  }
  LANTERN_CUDA(cudaMemcpyAsync(s->d_tokens, in->tree_tokens, rows * 4, cudaMemcpyHostToDevice, st));
  const size_t n_ri = (size_t)(cfg->retrieve_shared ? 1 : B) * L * D;
  LANTERN_CUDA(cudaMemcpyAsync(s->d_retrieve, in->retrieve, n_ri * 4, cudaMemcpyHostToDevice, st));
  if (in->row_kinds) LANTERN_CUDA(cudaMemcpyAsync(s->d_kinds, in->row_kinds, rows, cudaMemcpyHostToDevice, st));
  if (in->uniforms)
    LANTERN_CUDA(cudaMemcpyAsync(s->d_uniforms, in->uniforms, (size_t)B * cfg->n_uniforms * 4, cudaMemcpyHostToDevice, st));

  lantern_accept_cfg dc = *cfg;
  dc.row_stride = s->width;
  dc.item_stride = (int64_t)T * s->width;
  lantern_accept_in di;
  memset(&di, 0, sizeof(di));
  // device rows start at column c_lo: bias the base pointers so that absolute column numbers still address them
  di.logits_cond = s->d_cond - (size_t)s->c_lo * eb;
  di.logits_uncond = in->logits_uncond ? s->d_uncond - (size_t)s->c_lo * eb : nullptr;
  di.tree_tokens = s->d_tokens;
  di.retrieve = s->d_retrieve;
  di.row_kinds = in->row_kinds ? s->d_kinds : nullptr;
  di.nbr_table = s->d_table;
  di.uniforms = in->uniforms ? s->d_uniforms : nullptr;
  if (cfg->static_tree) {
    if (!in->node_q || !in->draft_op || !in->node_qrow || !in->sib_off || !in->sib_idx || !in->sib_tokens) {
      set_error("lantern_session_step: static_tree=1 needs the draft inputs");
      return LANTERN_E_INVALID;
    }
    const size_t n_sib = (size_t)in->sib_off[T];
    const size_t n_sib_tok = (size_t)B * in->sib_tokens_stride;
    if (n_sib > s->sib_idx_cap) {
      cudaFree(s->d_sib_idx);
      s->d_sib_idx = nullptr;
      LANTERN_CUDA(cudaMalloc(&s->d_sib_idx, (n_sib + 1) * 4));
      s->sib_idx_cap = n_sib;
    }
    if (n_sib_tok > s->sib_tok_cap) {
      cudaFree(s->d_sib_tokens);
      s->d_sib_tokens = nullptr;
      LANTERN_CUDA(cudaMalloc(&s->d_sib_tokens, (n_sib_tok + 1) * 4));
      s->sib_tok_cap = n_sib_tok;
    }
    LANTERN_CUDA(cudaMemcpyAsync(s->d_node_q, in->node_q, rows * 4, cudaMemcpyHostToDevice, st));
    LANTERN_CUDA(cudaMemcpyAsync(s->d_op, in->draft_op, (size_t)B * cfg->n_q_rows * cfg->vocab * 4,
                                 cudaMemcpyHostToDevice, st));
    LANTERN_CUDA(cudaMemcpyAsync(s->d_qrow, in->node_qrow, (size_t)T * 4, cudaMemcpyHostToDevice, st));
    LANTERN_CUDA(cudaMemcpyAsync(s->d_sib_off, in->sib_off, ((size_t)T + 1) * 4, cudaMemcpyHostToDevice, st));
    if (n_sib) LANTERN_CUDA(cudaMemcpyAsync(s->d_sib_idx, in->sib_idx, n_sib * 4, cudaMemcpyHostToDevice, st));
    LANTERN_CUDA(cudaMemcpyAsync(s->d_sib_tokens, in->sib_tokens, n_sib_tok * 4, cudaMemcpyHostToDevice, st));
    di.node_q = s->d_node_q;
    di.draft_op = s->d_op;
    di.node_qrow = s->d_qrow;
    di.sib_off = s->d_sib_off;
    di.sib_idx = s->d_sib_idx;
    di.sib_tokens = s->d_sib_tokens;
    di.sib_tokens_stride = in->sib_tokens_stride;
  }
  lantern_accept_out dout;
  memset(&dout, 0, sizeof(dout));
  int32_t* o = s->d_out;
  dout.accept_length = o;            o += B;
  dout.best_candidate = o;           o += B;
  dout.token = o;                    o += B;
  dout.n_draws = o;                  o += B;
  dout.flags = o;                    o += B;
  dout.path_tokens = o;              o += (size_t)B * D;
  dout.select_indices = o;           o += (size_t)B * D;
  if (out->sample_p) {
    if (!s->d_sample_p) LANTERN_CUDA(cudaMalloc(&s->d_sample_p, (size_t)m.n_items * m.vocab * 4));
    dout.sample_p = s->d_sample_p;
  }
  int rc = LANTERN_E_UNSUPPORTED;
  if (in_place) {
    lantern_accept_cfg hc = *cfg;           // host strides, absolute columns
    lantern_accept_in hi = di;
    hi.logits_cond = dev_cond;
    hi.logits_uncond = in->logits_uncond ? dev_uncond : nullptr;
    rc = lantern_accept_phases(&hc, &hi, &dout, s->d_work, s->work_bytes, st, 6 | 16);   // lazy, no speculative row reads
    if (rc == LANTERN_OK) s->last_in_place = 1;
    else if (rc == LANTERN_E_UNSUPPORTED) {   // not lazy-eligible: stage the window after all
      LANTERN_CUDA(copy_rows(s->d_cond, in->logits_cond));
      if (in->logits_uncond) LANTERN_CUDA(copy_rows(s->d_uncond, in->logits_uncond));
    }
  }
  if (rc == LANTERN_E_UNSUPPORTED) rc = lantern_accept_fused(&dc, &di, &dout, s->d_work, s->work_bytes, st);
  if (rc) return rc;
  const size_t n_out = (size_t)B * (5 + 2 * (size_t)D);
  LANTERN_CUDA(cudaMemcpyAsync(s->h_out, s->d_out, n_out * 4, cudaMemcpyDeviceToHost, st));
  if (out->sample_p)
    LANTERN_CUDA(cudaMemcpyAsync(out->sample_p, s->d_sample_p, (size_t)B * cfg->vocab * 4, cudaMemcpyDeviceToHost, st));
  LANTERN_CUDA(cudaStreamSynchronize(st));
  const int32_t* h = s->h_out;
  memcpy(out->accept_length, h, (size_t)B * 4);                  h += B;
  memcpy(out->best_candidate, h, (size_t)B * 4);                 h += B;
  memcpy(out->token, h, (size_t)B * 4);                          h += B;
  if (out->n_draws) memcpy(out->n_draws, h, (size_t)B * 4);      h += B;
  if (out->flags) memcpy(out->flags, h, (size_t)B * 4);          h += B;
  if (out->path_tokens) memcpy(out->path_tokens, h, (size_t)B * D * 4);       h += (size_t)B * D;
  if (out->select_indices) memcpy(out->select_indices, h, (size_t)B * D * 4);
  return LANTERN_OK;
}
