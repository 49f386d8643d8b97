/*
 * lantern_b200 — C ABI of the B200-native LANTERN verification hot path.
 *
 * The reference (jadohu/LANTERN) is pure Python/PyTorch and has no FFI, plugin or operator
 * registry; its boundary for this path is a set of Python call signatures.
 * Each entry point below states which reference code it replaces.  The Python shims in
 * lantern_b200/ keep those signatures and call this library through ctypes
 * (see INTEGRATION.md for the binding a reference maintainer would add).
 *
 * Conventions
 *   - plain pointers and sizes only; all `*_dev` pointers are CUDA device pointers owned by
 *     the caller; nothing is allocated, cached or freed inside (the host-buffer session and the
 *     lantern_debug_* test hook say so where they do);
 *   - every launch goes to the caller's `stream` (a cudaStream_t passed as void*; NULL = legacy
 *     default stream); calls are asynchronous unless documented otherwise;
 *   - return value: 0 = LANTERN_OK, negative = LANTERN_E_*, positive = cudaError_t;
 *     `lantern_last_error()` returns a thread-local message for the last failure;
 *   - re-entrant, no global mutable state except the sessions the caller creates.
 */
#ifndef LANTERN_B200_H_
#define LANTERN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LANTERN_ABI_VERSION 2

#if defined(__GNUC__)
#define LANTERN_API __attribute__((visibility("default")))
#else
#define LANTERN_API
#endif

enum {
  LANTERN_OK = 0,
  LANTERN_E_INVALID = -1,      /* bad argument (message says which) */
  LANTERN_E_UNSUPPORTED = -2,  /* valid but not implemented for this size / mode */
  LANTERN_E_WORKSPACE = -3,    /* workspace too small */
  LANTERN_E_NO_DEVICE = -4     /* no sm_100 device */
};

enum { LANTERN_F32 = 0, LANTERN_BF16 = 1, LANTERN_F16 = 2 };

/* Model families: the per-family differences of evaluate_posterior. */
enum {
  LANTERN_FAMILY_VANILLA = 0,  /* models/drafters/utils.py:333-410 (tail row not re-warped, :408) */
  LANTERN_FAMILY_LLAMAGEN = 1, /* models/ea_model_llamagen.py:463-669, :709-787 */
  LANTERN_FAMILY_ANOLE = 2,    /* models/ea_model_anole.py:464-669, :709-788 (offset 4, :931 mask) */
  LANTERN_FAMILY_LUMINA = 3    /* models/ea_model_lumina_mgpt.py:556-729 */
};

/* Lumina row classes of MultiModalLogitsProcessor (ea_model_lumina_mgpt.py:45-86). */
enum { LANTERN_ROW_IMAGE = 0, LANTERN_ROW_NEWLINE = 1, LANTERN_ROW_EOI = 2 };

/* Bits of lantern_accept_out.flags[b]. */
enum {
  LANTERN_OUT_RESIDUAL_TAIL = 1,    /* sample_p is the residual distribution (adjustflag path) */
  LANTERN_OUT_UNIFORM_FALLBACK = 2, /* a residual summed to 0 and was reset to ones (:775-776) */
  LANTERN_OUT_ROWS_READ_SHIFT = 8   /* bits 8..15: logits rows the walk turned into probabilities for this item */
};

#define LANTERN_MAX_SYNTAX_TOKENS 8

/*
 * One verify step over `n_items` independent prompts.  Shapes follow the reference:
 * T = n_rows tree nodes (root included), candidates/retrieve_indices are [n_paths, depth].
 */
typedef struct lantern_accept_cfg {
  int32_t n_items;   /* B: prompts in this launch (reference: 1) */
  int32_t n_rows;    /* T: logits rows per item */
  int32_t n_paths;   /* L */
  int32_t depth;     /* D */
  int32_t vocab;     /* V: elements per logits row */
  int32_t col0;      /* first live (image-token) column */
  int32_t ncols;     /* live columns; everything else has probability 0 */
  int32_t logits_dtype; /* LANTERN_F32 | LANTERN_BF16 | LANTERN_F16; arithmetic is fp32 */
  int64_t item_stride;  /* elements between consecutive items in logits_cond / logits_uncond */
  int64_t row_stride;   /* elements between consecutive rows (>= vocab) */
  int32_t family;       /* LANTERN_FAMILY_* */
  int32_t static_tree;  /* 0: q(x)=1 (EAGLE-2 dynamic tree); 1: LANTERN++ static tree with drafter q */
  float cfg_scale;      /* used when logits_uncond != NULL: uncond + (cond - uncond) * scale */
  float temperature;    /* HF TemperatureLogitsWarper (skipped when == 1) */
  float top_p;          /* HF TopPLogitsWarper when 1e-8 <= top_p < 1 */
  int32_t top_k;        /* HF TopKLogitsWarper / InterleavedTopKLogitsWarper when > 0 (ties kept) */
  int32_t lantern;      /* relaxed acceptance on/off */
  int32_t lantern_k;    /* neighbours aggregated (k+1 are zeroed on rejection) */
  float lantern_delta;  /* <= 1: additive bound delta; > 1: multiplicative bound (delta-1)*p(x) */
  float lantern_delta_m1; /* (float)(delta - 1.0) computed in double by the caller */
  int32_t table_cols;   /* row stride of nbr_table (>= min(lantern_k + 1, N - 1)) */
  int32_t tok_offset;   /* image_token_offset: table row = token - tok_offset, entries + tok_offset */
  int32_t n_syntax;     /* Lumina: tokens accepted with p = 1 (ea_model_lumina_mgpt.py:654-656) */
  int32_t syntax_tokens[LANTERN_MAX_SYNTAX_TOKENS];
  int32_t newline_token; /* Lumina: the only finite column of a NEWLINE row (8803) */
  int32_t eoi_token;     /* Lumina: the only finite column of an EOI row (8196) */
  int32_t retrieve_shared; /* 1: one retrieve_indices [L,D] for all items (static tree) */
  int32_t n_uniforms;   /* row stride of `uniforms` (>= tried candidates + 1 bonus draw) */
  int32_t n_q_rows;     /* static: rows per item in draft_op (sum over levels of parent groups) */
  int32_t bonus_uniform_last; /* 1: the bonus-token draw reads uniforms[b, n_uniforms-1] instead of the next unread one */
  uint64_t philox_seed; /* used when uniforms == NULL */
  uint64_t philox_step; /* verify-step counter of the device Philox stream */
} lantern_accept_cfg;

typedef struct lantern_accept_in {
  const void* logits_cond;      /* [B, T, V] (strides above), dtype logits_dtype */
  const void* logits_uncond;    /* same layout, or NULL: logits_cond is already CFG-mixed */
  const int32_t* tree_tokens;   /* [B, T] token of every tree node (tree_candidates) */
  const int32_t* retrieve;      /* [B or 1, L, D] node index, -1 padded (retrieve_indices) */
  const uint8_t* row_kinds;     /* [B, T] LANTERN_ROW_* or NULL (all IMAGE) */
  const int32_t* nbr_table;     /* [N, table_cols] nearest-first codebook ids (nearest_latents) */
  const float* uniforms;        /* [B, n_uniforms] or NULL -> device Philox stream */
  /* static tree only (evaluate_posterior_v1 extras) */
  const float* node_q;          /* [B, T] drafter probability of each node's token (cart_candidates_prob) */
  const float* draft_op;        /* [B, n_q_rows, V] fp32 drafter distributions (op / original_prob) */
  const int32_t* node_qrow;     /* [T] row of draft_op holding the node's sibling-group distribution */
  const int32_t* sib_off;       /* [T + 1] CSR offsets into sib_idx */
  const int32_t* sib_idx;       /* earlier siblings of each node, as indices into sib_tokens rows */
  const int32_t* sib_tokens;    /* [B, sib_tokens_stride] tokens addressed by sib_idx (tree_candidates) */
  int64_t sib_tokens_stride;
} lantern_accept_in;

typedef struct lantern_accept_out {
  int32_t* accept_length;  /* [B] 0 .. D-1 */
  int32_t* best_candidate; /* [B] row of candidates */
  int32_t* token;          /* [B] bonus token drawn from sample_p (inverse CDF, one uniform) */
  int32_t* path_tokens;    /* [B, D] candidates[best, :a+1], rest -1 */
  int32_t* select_indices; /* [B, D] retrieve_indices[best, :a+1], rest -1 */
  int32_t* n_draws;        /* [B] uniforms consumed (bonus draw included) */
  int32_t* flags;          /* [B] LANTERN_OUT_* */
  float* sample_p;         /* [B, V] or NULL */
} lantern_accept_out;

/* Library / ABI version: (LANTERN_ABI_VERSION << 16) | patch. */
LANTERN_API int lantern_version(void);
LANTERN_API const char* lantern_last_error(void);

/* Scratch the fused step needs for `cfg`: 32 bytes of softmax statistics per logits row, plus one fp32 probability
 * vector per prompt when the live window is too wide for shared memory (more than ~50K columns, e.g. plain EAGLE
 * verification on a 65536-entry vocabulary; such rows take a multi-pass statistics kernel - a parity path, not tuned). */
LANTERN_API size_t lantern_accept_workspace_bytes(const lantern_accept_cfg* cfg);

/*
 * The fused verify step.  Replaces, for all items at once and without host round trips:
 *   cfg_logit_process                ea_model_llamagen.py:26-29, ea_model_lumina_mgpt.py:597
 *   tree_decoding post-processing    ea_model_llamagen.py:930-931, ea_model_anole.py:930-932,
 *                                    ea_model_lumina_mgpt.py:597-607 (the [L,D,V] gather is never built)
 *   prepare_logits_processor warpers drafters/utils.py:36-52 (+ HF Temperature/TopP/TopK)
 *   evaluate_posterior[_v1]          drafters/utils.py:371-410, ea_model_llamagen.py:597-669, :709-787,
 *                                    ea_model_anole.py:598-669, :709-788, ea_model_lumina_mgpt.py:610-726
 *   bonus-token draw                 ea_model_llamagen.py:976-979 (torch.multinomial -> inverse CDF)
 */
LANTERN_API int lantern_accept_fused(const lantern_accept_cfg* cfg, const lantern_accept_in* in,
                         const lantern_accept_out* out, void* workspace_dev, size_t workspace_bytes,
                         void* stream);

/* The same step with an explicit schedule.  phases bit 0: per-row statistics kernel over every tree row (streamed,
 * HBM-bound); bit 1: walk kernel; bit 2: lazy - the walk computes the statistics of the rows it visits itself (no
 * streamed kernel; needs a 2048/4096/8192/16384-column window and no top-p); bit 3: automatic - lazy when eligible
 * and the batch holds at least 2048 tree rows or the trees at least 48 rows each, streamed otherwise; bit 4: no speculative row prefetch in the walk (used
 * when the logits are read over PCIe).  3 is lantern_accept_fused, 6 lazy, 8 automatic,
 * 1 times the HBM-bound kernel on its own (bench.py).  Results are identical in every schedule. */
LANTERN_API int lantern_accept_phases(const lantern_accept_cfg* cfg, const lantern_accept_in* in,
                                      const lantern_accept_out* out, void* workspace_dev, size_t workspace_bytes,
                                      void* stream, int phases);

/*
 * Greedy verification (logits_processor is None / temperature 0), tree form, LlamaGen / Anole / vanilla families:
 * drafters/utils.py:356-369 and ea_model_anole.py:789-902.  cfg->lantern = 0: a child is accepted iff its token is the
 * argmax of its parent's (CFG-mixed, masked) logits row.  cfg->lantern = 1: the "TVD" relaxation of the reference -
 * gtp = softmax(row), px = gtp[x], approx = px + cumsum(gtp[neighbours]), tvd = 0.5|px - approx| + cumsum(0.5 gtp[nb]),
 * last position within lantern_delta (or (delta-1) px) replaces gtp[x]; accepted iff x is then the row's argmax.
 * accept_length = longest accepted prefix over the leaf paths (first path on ties, path 0 if none).  out->token is
 * the argmax of the last accepted node's row, which is also written to out->sample_p ([B, vocab], the reference's
 * third return value `logits[best, accept_length]`) when that pointer is set.  Rows must span the vocabulary
 * (row_stride >= vocab); warp knobs, uniforms and the static-tree inputs are ignored.  Workspace: 4 bytes per tree row.
 */
LANTERN_API size_t lantern_accept_greedy_workspace_bytes(const lantern_accept_cfg* cfg);
LANTERN_API int lantern_accept_greedy(const lantern_accept_cfg* cfg, const lantern_accept_in* in,
                                      const lantern_accept_out* out, void* workspace_dev, size_t workspace_bytes,
                                      void* stream);

/*
 * Glue for the reference's evaluate_posterior signature (drafters/utils.py:333, ea_model_llamagen.py:709), which
 * receives `candidates [L, D]` (int64, -1 padded; token of node retrieve_indices[j, i]) instead of the tree's token
 * vector: writes tree_tokens [n_rows] int32 (0 for unreachable nodes) and the int32 copy of retrieve_indices
 * [n_paths, depth] that lantern_accept_fused consumes.  One memset + one launch, no host synchronisation.
 */
LANTERN_API int lantern_tree_from_candidates(const int64_t* cand_dev, const int64_t* retrieve_dev, int32_t n_paths,
                                             int32_t depth, int32_t n_rows, int32_t* tokens_dev,
                                             int32_t* retrieve32_dev, void* stream);

/*
 * Single-prompt drop-in call (the reference's batch-1 evaluate_posterior(logits, candidates, ...),
 * ea_model_llamagen.py:709 / :464, ea_model_lumina_mgpt.py:610, drafters/utils.py:333): one entry that uploads the
 * step's uniforms, rebuilds the tree inputs from `candidates` / `retrieve_indices` (lantern_tree_from_candidates),
 * runs the fused step with the automatic schedule, reads the results back and synchronises `stream` once.
 * The context owns small staging buffers (pinned host + device, allocated by lantern_call_create for trees of up to
 * max_rows nodes and max_cells = n_paths * depth path cells); one context per host thread and device.
 * Uniforms: write n_uniforms values into lantern_call_uniforms(h) before the call (n_uniforms = 0: device Philox
 * stream of cfg).  cfg->n_items must be 1; in->tree_tokens / retrieve / uniforms are ignored.
 * out_host[5 + 2*depth] = {accept_length, best_candidate, token, n_draws, flags, path_tokens[depth],
 * select_indices[depth]}.  sample_p_dev: [vocab] fp32 or NULL.
 */
typedef struct lantern_call lantern_call;
LANTERN_API int lantern_call_create(int32_t max_rows, int32_t max_cells, lantern_call** out);
LANTERN_API void lantern_call_destroy(lantern_call* h);
LANTERN_API float* lantern_call_uniforms(lantern_call* h);
LANTERN_API int lantern_posterior_call(lantern_call* h, const lantern_accept_cfg* cfg, const lantern_accept_in* in,
                                       const int64_t* cand_dev, const int64_t* retrieve_dev, int32_t n_uniforms,
                                       float* sample_p_dev, void* workspace_dev, size_t workspace_bytes,
                                       int32_t* out_host, void* stream);

/*
 * Bonus-token draw from caller-supplied probability rows (update_inference_inputs given a
 * sample_p that did not come from lantern_accept_fused): token[b] = min{i : cdf_i > u[b] * total}.
 * Replaces torch.multinomial(prob, 1) at ea_model_llamagen.py:978, ea_model_lumina_mgpt.py:781.
 */
LANTERN_API int lantern_sample_tokens(const float* probs_dev, int64_t row_stride, int32_t n_rows, int32_t vocab,
                          const float* uniforms_dev, int32_t* tokens_dev, void* stream);

/*
 * KV-cache compaction of update_inference_inputs (ea_model_llamagen.py:962-970,
 * ea_model_lumina_mgpt.py:741-746, drafters/kv_cache.py:38-52): for every slab
 * [n_outer, S_max, head_dim] (n_outer = 2*layers*batch*heads flattened) copy positions
 * select[b, 0:n_keep[b]] to prev_len[b] .. prev_len[b]+n_keep[b].  `batch_of_outer` maps the flattened
 * outer index to its item: b = (outer / outer_per_batch) % n_batch.
 */
typedef struct lantern_kv_cfg {
  int32_t n_slabs;
  int32_t elem_bytes;       /* 2 (bf16/fp16) or 4 */
  int64_t n_outer;          /* product of the leading dims of one slab */
  int64_t outer_per_batch;  /* heads (the dims after batch, before S) */
  int32_t n_batch;          /* batch dim of the slab (2 for parallel CFG) */
  int32_t s_max;            /* max_position_embeddings */
  int32_t head_dim;
  int32_t max_keep;         /* row stride of `select` (D) */
  /* Single-prompt shortcuts (the reference's batch-1 call): host scalars used where the matching device array
   * argument is NULL, so the drop-in shim needs no host-to-device copy at all. */
  void* slab0;              /* used when slab_ptrs_dev == NULL (n_slabs must be 1) */
  int32_t prev_len0;        /* used for every item when prev_len_dev == NULL */
  int32_t n_keep0;          /* used for every item when n_keep_dev == NULL */
  int32_t select_i64;       /* != 0: `select_dev` holds int64 (torch.long) indices */
  int32_t reserved0;
} lantern_kv_cfg;
LANTERN_API int lantern_kv_compact(const lantern_kv_cfg* cfg, void* const* slab_ptrs_dev, const void* select_dev,
                       const int32_t* prev_len_dev, const int32_t* n_keep_dev, void* stream);

/*
 * Neighbour-table build (entrypoints/generate_codebook.py:53-60): for every codebook row the ids of
 * its K nearest other rows, nearest first; order = (squared L2 distance in fp64, id).
 * E_dev: [N, d] fp32 row-major.  out_dev: [N, K] int32.  Caller-owned scratch of
 * lantern_build_neighbors_workspace_bytes(N, d, K) bytes; nothing is allocated, no host synchronisation, every launch
 * goes to `stream`.  For K <= N/2 (N <= 16384, d <= 256) the candidates come from a tcgen05 distance GEMM and are
 * re-ranked exactly; rows that route cannot finish (candidate overflow on massive ties, a violated error bound) and all
 * other shapes are computed entirely in fp64 by a second kernel.  Both routes give identical tables.
 * route_dev (optional, device int32[2]): [0] = 1 when every row came from the tensor-core route, else 2;
 * [1] = rows computed by the all-fp64 kernel.
 */
LANTERN_API size_t lantern_build_neighbors_workspace_bytes(int32_t N, int32_t d, int32_t K);
LANTERN_API int lantern_build_neighbors(const float* E_dev, int32_t N, int32_t d, int32_t K, int32_t* out_dev,
                                        void* workspace_dev, size_t workspace_bytes, int32_t* route_dev, void* stream);

/*
 * Dynamic (EAGLE-2) draft-tree post-processing: the host loops at the end of the drafter's topK_genrate
 * (models/drafters/cnets_llamagen.py:831-908, cnets_lumina_mgpt.py:1330-1393).  Per prompt: scores [n_cand] =
 * cat(scores_list), tokens [n_cand] = cat(ss_token), parents [n_groups] = cat(parents_list) (parent flat id + 1 of
 * each group of `top_k` siblings, 0 = root).  Outputs: tree_tokens [B,T] (draft_tokens, root = sample token),
 * parent [B,T], depth [B,T] (tree_position_ids), mask [B,T,T] (tree_mask, optional), retrieve [B,T,d_max]
 * (retrieve_indices, -1 padded; valid block is counts[b] = {n_leaves, max_depth+1}; a tree deeper than d_max gives
 * counts[b] = {-1, max_depth+1} and a truncated table); rows sorted like the reference
 * does when a logits processor is present if sort_rows != 0.  top-k tie rule: higher score, then lower flat index.
 */
LANTERN_API int lantern_build_dynamic_tree(const float* scores_dev, const int32_t* tokens_dev, const int32_t* parents_dev,
                                           const int32_t* root_tokens_dev, int32_t n_items, int32_t n_cand,
                                           int32_t n_groups, int32_t top_k, int32_t total_tokens, int32_t d_max,
                                           int32_t sort_rows, int32_t* tree_tokens_dev, int32_t* parent_dev,
                                           int32_t* depth_dev, float* mask_dev, int32_t* retrieve_dev,
                                           int32_t* counts_dev, void* stream);

/*
 * Static-tree drafter sampling (Model.sample, models/drafters/cnets_llamagen.py:924-940): for each of the
 * n_items * n_rows logits rows (cond / optional uncond, warped with cfg's temperature / top_p / top_k) write the
 * full distribution probs [rows, V] (`op`), k tokens drawn without replacement idx [rows, k] (exponential race on
 * the device Philox stream cfg.philox_seed / philox_step; same law as torch.multinomial(p, k, False)) and their
 * conditional probabilities p_i / (1 - sum_{j<i} p_j) clamped to [0,1] (cond_probs [rows, k]).
 * Only the shape / window / dtype / warp / philox fields of cfg and in->logits_* are used.
 */
LANTERN_API int lantern_draft_sample(const lantern_accept_cfg* cfg, const lantern_accept_in* in, int32_t k,
                                     float* probs_dev, int32_t* idx_dev, float* cond_probs_dev, void* workspace_dev,
                                     size_t workspace_bytes, void* stream);

/* Test hook of the tensor-core path of lantern_build_neighbors: the approximate squared-distance matrix
 * D~[i][j] = |e_i|^2 + |e_j|^2 - 2 e_i.e_j (tcgen05, TF32 cross term), fp32 [N, ld]. The diagonal holds
 * the computed value (~0); the select stage excludes self by index. */
LANTERN_API int lantern_debug_dist_gemm(const float* E_dev, int32_t N, int32_t d, float* D_dev, int32_t ld, void* stream);

/* Host-side copy of the device Philox stream (for tests and for seeding the CPU oracle):
 * draw i of item `item` at step `step`, i in [0, n). */
LANTERN_API void lantern_philox_uniforms(uint64_t seed, uint64_t step, uint32_t item, int32_t n, float* out_host);

/*
 * Host-buffer session: the call a reference-side caller makes when its logits live in host
 * memory.  Owns device buffers, pinned staging and a stream sized for `cfg`; `lantern_session_step` is synchronous.
 * Two routes, same results:
 *   - in place: if the logits are page-locked host memory the device can address (cudaHostAlloc /
 *     cudaHostRegister, e.g. a torch pin_memory() tensor) and the window is lazy-eligible (see
 *     lantern_accept_phases), the walk reads the rows it visits straight from host memory over PCIe: only
 *     ~accept_length + 2 of the T tree rows of each prompt cross the bus (flags bits 8..15 report how many);
 *   - staged: otherwise (pageable memory, top-p, other window widths, or LANTERN_SESSION_STAGED=1 in the
 *     environment) the live column window of every row is copied to the device and lantern_accept_fused runs on it.
 */
typedef struct lantern_session lantern_session;
LANTERN_API int lantern_session_create(const lantern_accept_cfg* cfg, const int32_t* nbr_table_host, int32_t table_rows,
                           lantern_session** out);
LANTERN_API int lantern_session_step(lantern_session* s, const lantern_accept_cfg* cfg, const lantern_accept_in* in_host,
                         const lantern_accept_out* out_host);
LANTERN_API void lantern_session_destroy(lantern_session* s);
/* Route of the last lantern_session_step: 1 = logits read in place from page-locked host memory, 0 = staged. */
LANTERN_API int lantern_session_last_route(const lantern_session* s);

#ifdef __cplusplus
}
#endif
#endif /* LANTERN_B200_H_ */
