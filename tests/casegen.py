"""Shared case builder: one parameter dict -> identical inputs for the live reference
(tests/golden/gen_golden.py), the NumPy oracle and the CUDA path."""
from __future__ import annotations

import os
import sys
from dataclasses import dataclass
from typing import Optional

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from lantern_b200 import synth  # noqa: E402
from oracle import lantern_oracle as O  # noqa: E402

from lantern_b200 import choices as CH  # noqa: E402


def family_of(name: str, ncols: Optional[int] = None) -> O.Family:
    base = {"llamagen": O.LLAMAGEN, "anole": O.ANOLE, "lumina_mgpt": O.LUMINA, "vanilla": O.VANILLA}[name]
    if ncols is None or ncols == base.ncols:
        return base
    return O.small_family(base, ncols)


@dataclass
class Built:
    params: dict
    fam: O.Family
    warp: O.Warp
    cond: np.ndarray                 # [T, V]
    uncond: Optional[np.ndarray]     # [T, V] or None
    tree: synth.Tree
    candidates: np.ndarray           # [L, D]
    table: Optional[np.ndarray]      # [N, k+1] int32
    uniforms: np.ndarray             # [T + 1] fp32
    row_kinds: Optional[np.ndarray]  # [T] int8 (Lumina)
    static: Optional[O.StaticDraft]
    static_synth: Optional[synth.StaticDraftSynth] = None
    tree_buffers: Optional[dict] = None


def default_params(**kw) -> dict:
    p = dict(family="llamagen", ncols=None, tree="eagle2", total_tokens=59, depth=4, seed=0,
             lantern=True, lantern_k=1000, lantern_delta=0.1, temperature=1.0, top_k=2000, top_p=1.0,
             cfg_scale=3.0, cfg=True, boost=13.0, static_tree=None, newline_depth=-1, sharp=1.0, table_seed=0,
             newline_junk=False, dup_siblings=0)
    p.update(kw)
    return p


def build(params: dict) -> Built:
    p = default_params(**params)
    fam = family_of(p["family"], p["ncols"])
    seed = int(p["seed"])
    static_name = p["static_tree"]
    tbuf = None
    if static_name is not None:
        choices = CH.tree(static_name)
        tbuf = O.generate_tree_buffers(choices)
        ri = tbuf["retrieve_indices"]
        T = int(tbuf["tree_indices"].shape[0])
        # parent pointers from the attention mask: deepest ancestor other than self
        depth = tbuf["tree_position_ids"]
        parent = np.full(T, -1, dtype=np.int64)
        for row in ri:
            for a, b in zip(row[:-1], row[1:]):
                if b >= 0:
                    parent[b] = a
        tree = synth.Tree(parent, depth, ri)
    elif p["tree"] == "eagle2":
        tree = synth.eagle2_tree(seed, p["total_tokens"], p["depth"])
    else:
        tree = synth.random_tree(seed, p["total_tokens"], max_depth=p["depth"] + 2)
    T = tree.T
    lo, hi = fam.col0, fam.col1

    row_kinds = None
    static = None
    ssyn = None
    if static_name is not None:
        counts = synth.static_group_counts(choices)
        ssyn = synth.static_draft(seed, counts, fam.vocab, lo, hi, sharp=p["sharp"])
        root_tok = lo + int(synth.hash_u64(seed, 1, 55)[0] % np.uint64(hi - lo))
        cart, cart_prob, tree_cand = O.generate_candidates(ssyn.ss_token, ssyn.ss_prob, tbuf["tree_indices"],
                                                           ri, root_tok)
        tree.tokens = tree_cand.astype(np.int64)
        static = O.StaticDraft(cart_prob, ssyn.op, tbuf["p_indices"], tbuf["b_indices"], tree.tokens)
    else:
        synth.assign_tokens(seed, tree, lo, hi)
        if p["dup_siblings"]:
            # Duplicate tokens among siblings (never produced by a top-k drafter, but the reference dedups the children
            # of a node by TOKEN, ea_model_llamagen.py:728-739, so two subtrees can sit behind one accepted token): the
            # later sibling takes an earlier sibling's token.
            r = synth.hash_u64(seed, 4 * T, stream=91)
            done = 0
            for t in range(4 * T):
                i = 1 + int(r[t] % np.uint64(T - 1))
                sibs = [j for j in range(1, i) if tree.parent[j] == tree.parent[i] and tree.tokens[j] != tree.tokens[i]]
                if sibs:
                    tree.tokens[i] = tree.tokens[sibs[int(r[(t + 1) % (4 * T)] % np.uint64(len(sibs)))]]
                    done += 1
                    if done >= p["dup_siblings"]:
                        break

    if fam.lumina:
        row_kinds = np.zeros(T, dtype=np.int8)
        nd = p["newline_depth"]
        if nd >= 0:
            row_kinds[tree.depth == nd] = O.ROW_NEWLINE
            if static_name is None:
                # children of newline rows: first child is the newline token, the rest are junk ids
                # (newline_junk: every child is junk, so the walk rejects them all and ends on the one-hot row)
                seen = set(range(T)) if p["newline_junk"] else set()
                for i in range(1, T):
                    par = int(tree.parent[i])
                    if row_kinds[par] == O.ROW_NEWLINE:
                        if par not in seen:
                            tree.tokens[i] = O.LUMINA_NEWLINE_TOKEN
                            seen.add(par)
                        else:
                            tree.tokens[i] = fam.col1 + 700 + (i % 20)     # non-image, non-syntax

    cond, uncond = synth.tree_logits(seed, tree, fam.vocab, cfg=p["cfg"], boost=p["boost"])
    cand = O.candidates_from_tree(tree.tokens, tree.retrieve_indices)
    table = None
    if p["lantern"]:
        k = min(int(p["lantern_k"]), fam.ncols - 1)
        table = synth.neighbor_table(p["table_seed"], fam.ncols, min(k + 1, fam.ncols - 1))
    warp = O.Warp(p["temperature"], p["top_p"], p["top_k"])
    u = synth.uniforms(seed, T + 1, stream=77)
    return Built(p, fam, warp, cond, uncond, tree, cand, table, u, row_kinds, static, ssyn, tbuf)


def oracle_step(b: Built, keep_trace: bool = False) -> O.StepResult:
    p = b.params
    k = min(int(p["lantern_k"]), b.fam.ncols - 1)
    return O.verify_step(b.cond, b.uncond, p["cfg_scale"], b.tree.tokens, b.tree.retrieve_indices,
                         b.uniforms, b.fam, b.warp, p["lantern"], k, p["lantern_delta"], b.table,
                         static=b.static, row_kinds=b.row_kinds, keep_trace=keep_trace)


def oracle_greedy(b: Built):
    """Greedy branches (``logits_processor is None``): tree_decoding post-processing + the [L, D, V] gather, then
    ``evaluate_posterior_greedy`` / ``evaluate_posterior_greedy_lantern``.  Returns (best, accept_length, row, margin)."""
    p = b.params
    T = b.cond.shape[0]
    rows = np.stack([O.cfg_mix(b.cond[n], b.uncond[n], p["cfg_scale"]) if b.uncond is not None
                     else b.cond[n].astype(np.float32) for n in range(T)])
    if b.fam.mask_non_image:
        rows = np.stack([O.anole_mask_row(r, b.fam) for r in rows])
    gathered = rows[np.asarray(b.tree.retrieve_indices, dtype=np.int64)]
    if p["lantern"]:
        k = min(int(p["lantern_k"]), b.fam.ncols - 1)
        return O.evaluate_posterior_greedy_lantern(gathered, b.candidates, b.fam, b.table, k, p["lantern_delta"])
    best, a, row = O.evaluate_posterior_greedy(gathered, b.candidates)
    return best, a, row, 1.0


def sample_p_probe(sample_p: np.ndarray, n: int = 48):
    """Compact fingerprint of a probability vector: its n largest entries + n hashed probes."""
    V = sample_p.shape[0]
    top = np.argsort(-sample_p, kind="stable")[:n]
    probe = (synth.hash_u64(12345, n, 3) % np.uint64(V)).astype(np.int64)
    idx = np.concatenate([top, probe]).astype(np.int32)
    return idx, sample_p[idx].astype(np.float32), int((sample_p > 0).sum()), float(sample_p.astype(np.float64).sum())
