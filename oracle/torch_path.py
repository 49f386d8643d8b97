"""PyTorch restatement of the reference's verification step, op for op, device-agnostic.

TEST / BASELINE INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): this is the "GPU-PyTorch path" of SURVEY.md 8(d) —
the reference's own sequence of small ATen calls and host synchronisations (``.item()``, ``bool(tensor)``,
``nonzero``, a numpy slice of the neighbour table copied to the device per candidate) — restated so that it can be
timed on the GPU box, where ``/root/reference`` does not exist.  ``bench.py`` times it as ``torch_gpu_baseline``;
``tests/test_oracle_golden.py`` checks its decisions against the numpy oracle (and therefore against the golden
vectors produced by the live reference).  Nothing under ``lantern_b200/`` imports it.

Restated code (dynamic / EAGLE-2 tree, sampling branch):
  * tree_decoding post-processing   ea_model_llamagen.py:26-29, 930-931; ea_model_anole.py:930-932;
                                    ea_model_lumina_mgpt.py:45-86, 106-112, 597-607
  * evaluate_posterior              ea_model_llamagen.py:709-787; ea_model_anole.py:709-788;
                                    ea_model_lumina_mgpt.py:610-726 (eagle_version 2)
  * HF warpers                      transformers 4.45 TemperatureLogitsWarper / TopPLogitsWarper / TopKLogitsWarper
  * bonus token                     the build's inverse-CDF replacement of torch.multinomial (ea_model_llamagen.py:976-979)
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import lantern_oracle as O


def hf_warp(scores: torch.Tensor, warp: O.Warp) -> torch.Tensor:
    """``logits_processor(None, scores)`` for a [1, V] tensor (drafters/utils.py:36-52 order)."""
    if not warp.enabled:
        return scores
    if warp.temperature != 1.0:
        scores = scores / warp.temperature
    if 1e-8 <= warp.top_p < 1.0:
        sorted_logits, sorted_indices = torch.sort(scores, descending=False)
        cumulative_probs = sorted_logits.softmax(dim=-1).cumsum(dim=-1)
        sorted_indices_to_remove = cumulative_probs <= (1 - warp.top_p)
        sorted_indices_to_remove[..., -1:] = 0
        indices_to_remove = sorted_indices_to_remove.scatter(1, sorted_indices, sorted_indices_to_remove)
        scores = scores.masked_fill(indices_to_remove, -float("inf"))
    if warp.top_k > 0:
        top_k = min(warp.top_k, scores.size(-1))
        indices_to_remove = scores < torch.topk(scores, top_k)[0][..., -1, None]
        scores = scores.masked_fill(indices_to_remove, -float("inf"))
    return scores


def tree_decoding_post(cond: torch.Tensor, uncond: Optional[torch.Tensor], cfg_scale: float, fam: O.Family,
                       top_k: int, retrieve: torch.Tensor, row_kinds: Optional[np.ndarray] = None) -> torch.Tensor:
    """CFG mix, family masks, (Lumina) top-k, then the ``[L, D, V]`` leaf-path gather."""
    logits = cond.float()
    if uncond is not None:
        u = uncond.float()
        logits = u + cfg_scale * (logits - u)                       # three elementwise kernels, as cfg_logit_process
    if fam.lumina:
        T = logits.shape[0]
        kinds = np.zeros(T, dtype=np.int64) if row_kinds is None else np.asarray(row_kinds, dtype=np.int64)
        new = torch.full_like(logits, -float("inf"))
        for t in range(T):                                          # the processor works row by row on positions
            if kinds[t] == O.ROW_IMAGE:
                new[t, fam.col0:fam.col1] = logits[t, fam.col0:fam.col1]
            elif kinds[t] == O.ROW_NEWLINE:
                new[t, O.LUMINA_NEWLINE_TOKEN] = 0.0
            else:
                new[t, O.LUMINA_EOI_TOKEN] = 0.0
        logits = new
        if top_k > 0:
            k = min(top_k, logits.size(-1))
            indices_to_remove = logits < torch.topk(logits, k)[0][..., -1, None]
            logits = logits.masked_fill(indices_to_remove, -float("inf"))
    elif fam.mask_non_image:
        masked = torch.full_like(logits, torch.finfo(torch.float32).min)
        masked[:, fam.col0:fam.col1] = logits[:, fam.col0:fam.col1]
        logits = masked
    return logits[retrieve]                                         # -1 gathers the last node, like the reference


def evaluate_posterior(logits: torch.Tensor, candidates: torch.Tensor, uniforms: Sequence[float], fam: O.Family,
                       warp: Optional[O.Warp], lantern: bool, lantern_k: int, lantern_delta: float,
                       table: Optional[np.ndarray]):
    """Sampling branch, dynamic tree.  Returns (best_candidate, accept_length, sample_p, n_draws)."""
    dev = logits.device
    L, D = candidates.shape
    draws = 0
    accept_length = 1
    accept_cand = candidates[0][:1]
    best_candidate = 0
    adjustflag = False
    gtp = None
    off = fam.offset
    for i in range(1, D):
        if i != accept_length:
            break
        adjustflag = False
        is_eq = (candidates[:, :accept_length] == accept_cand).all(dim=1)
        fi = torch.nonzero(is_eq, as_tuple=True)[0][0]
        gt_logits = logits[fi, i - 1][None]
        if warp is not None:
            gt_logits = hf_warp(gt_logits, warp)
        gtp = torch.softmax(gt_logits[0], dim=0)
        candidates_set = []
        for j in range(L):
            if is_eq[j]:                                            # host sync
                x = candidates[j, i]
                xi = x.item()                                       # host sync
                if xi in candidates_set or xi == -1:
                    continue
                candidates_set.append(xi)
                r = float(uniforms[draws])
                draws += 1
                px = gtp[xi]
                idx = -1
                relaxable = True
                if fam.lumina:
                    if xi in fam.syntax_tokens:
                        px = torch.ones((), device=dev)
                        relaxable = False
                    elif not (fam.col0 <= xi < fam.col1):
                        px = torch.zeros((), device=dev)
                        relaxable = False
                if lantern and relaxable:
                    # the table lives in host memory as numpy; its row slice is copied per candidate (:744)
                    nearest = torch.tensor(table[xi - off][:lantern_k].astype(np.int64), device=dev) + off
                    nearest_probs = gtp[nearest]
                    cumsum = torch.cumsum(nearest_probs, dim=0)
                    bound = (lantern_delta - 1) * px if lantern_delta > 1 else lantern_delta
                    valid = torch.nonzero(cumsum <= bound, as_tuple=True)[0]   # host sync
                    if valid.numel() > 0:
                        idx = int(valid[-1].item())
                        px = px + cumsum[idx]
                acp = px / 1.0
                if r <= acp:                                        # host sync
                    accept_cand = torch.cat((accept_cand, x[None]), dim=0)
                    accept_length += 1
                    best_candidate = j
                    break
                gtp[xi] = 0
                if lantern and relaxable and idx != -1:
                    nb = torch.tensor(table[xi - off][:lantern_k + 1].astype(np.int64), device=dev) + off
                    gtp[nb] = 0
                if gtp.sum() == 0:                                  # host sync
                    gtp = torch.ones_like(gtp)
                gtp = gtp / gtp.sum()
                adjustflag = True
    if adjustflag and accept_length != D:
        sample_p = gtp
    else:
        gt_logits = logits[best_candidate, accept_length - 1][None]
        if warp is not None and fam.tail_rewarp:
            gt_logits = hf_warp(gt_logits, warp)
        sample_p = torch.softmax(gt_logits[0], dim=0)
    return best_candidate, accept_length - 1, sample_p, draws


def sample_token(sample_p: torch.Tensor, u: float) -> int:
    c = torch.cumsum(sample_p.double(), dim=0)
    t = float(np.float32(u)) * c[-1]
    tok = int(torch.searchsorted(c, t, right=True).item())
    if tok >= c.numel():
        tok = int(torch.nonzero(sample_p > 0)[-1].item())
    return tok


def verify_step(cond: torch.Tensor, uncond: Optional[torch.Tensor], cfg_scale: float, tree_tokens: torch.Tensor,
                retrieve: torch.Tensor, uniforms: Sequence[float], fam: O.Family, warp: O.Warp, lantern: bool,
                lantern_k: int, lantern_delta: float, table: Optional[np.ndarray],
                row_kinds: Optional[np.ndarray] = None):
    """One prompt, one step: returns (best_candidate, accept_length, token, sample_p, n_uniforms)."""
    ri = retrieve.long()
    logits = tree_decoding_post(cond, uncond, cfg_scale, fam, warp.top_k, ri, row_kinds)
    ext = torch.cat([tree_tokens.long(), torch.full((1,), -1, dtype=torch.long, device=tree_tokens.device)])
    candidates = ext[ri]
    best, a, sample_p, draws = evaluate_posterior(logits, candidates, uniforms, fam, None if fam.lumina else warp,
                                                  lantern, lantern_k, lantern_delta, table)
    tok = sample_token(sample_p, float(uniforms[draws]))
    return best, a, tok, sample_p, draws + 1


# --------------------------------------------------------------------------------------------------------------------
# Next rows: torch op sequences of the drafter-side helpers, for GPU-PyTorch baseline timing
# --------------------------------------------------------------------------------------------------------------------
def dynamic_tree_tail(scores: torch.Tensor, tokens: torch.Tensor, parents: torch.Tensor, sample_token: torch.Tensor,
                      total_tokens: int, top_k: int, sort_rows: bool = True):
    """Tail of ``topK_genrate`` (cnets_llamagen.py:831-912) as the reference runs it: device topk / sort /
    searchsorted, then ``.tolist()`` round trips, a CPU ancestor-mask loop and Python list work for the leaf paths.
    Inputs are the flattened lists the expansion loop leaves behind.  Returns (draft_tokens [1,T], retrieve_indices
    [L,D], tree_mask [1,1,T,T], tree_position_ids [T])."""
    picked = torch.sort(torch.topk(scores, total_tokens, dim=-1).indices).values
    draft_tokens = torch.cat((sample_token[:1], tokens[picked]), dim=0)
    draft_parents = parents[picked // top_k].long()
    mask_index = torch.searchsorted(picked, draft_parents - 1, right=False)
    mask_index[draft_parents == 0] = -1
    mask_index = mask_index + 1
    par = mask_index.tolist()                                   # host sync
    tree_mask = torch.eye(total_tokens + 1).bool()               # CPU, like the reference
    tree_mask[:, 0] = True
    for i in range(total_tokens):
        tree_mask[i + 1].add_(tree_mask[par[i]])
    position_ids = torch.sum(tree_mask, dim=1) - 1
    max_depth = int(torch.max(position_ids).item()) + 1
    inner = set(torch.unique(mask_index).tolist())              # host sync
    depth_of = position_ids.tolist()
    paths = []
    for node in range(total_tokens + 1):
        if node in inner:
            continue
        row = [-1] * max_depth
        cur = node
        for j in range(depth_of[node], -1, -1):
            row[j] = cur
            cur = par[cur - 1]
        paths.append(row)
    if sort_rows:
        big = total_tokens + 5
        paths.sort(key=lambda r: [x if x >= 0 else big for x in r])
    retrieve = torch.tensor(paths, dtype=torch.long)
    return draft_tokens[None], retrieve, tree_mask.float()[None, None], position_ids.to(scores.device)


def drafter_sample(logits: torch.Tensor, warp: Optional[O.Warp], k: int):
    """``Model.sample`` (cnets_llamagen.py:924-940): warp, softmax, ``multinomial`` without replacement, gathered
    probabilities rescaled by the mass still available before each draw."""
    if warp is not None:
        logits = hf_warp(logits, warp)
    probs = torch.softmax(logits, dim=-1)
    idx = torch.multinomial(probs, k, replacement=False)
    picked = probs.gather(-1, idx)
    drawn_before = torch.nn.functional.pad(picked.cumsum(-1)[:, :-1], (1, 0))    # mass removed by the earlier draws
    cond = picked / (1 - drawn_before)
    cond = torch.where(torch.isfinite(cond), cond, torch.full_like(cond, -1.0)).clamp(0.0, 1.0)
    return idx, cond, probs
