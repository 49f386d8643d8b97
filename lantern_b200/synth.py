"""Deterministic synthetic inputs for the verification hot path.

Everything here is integer hashing plus exact power-of-two scaling, so the same
(seed, shape) produces bit-identical arrays on any machine / numpy version.  The
golden fixtures under ``tests/golden`` store only seeds and reference outputs and
rely on that.  Shapes follow SURVEY.md Appendix D:

* logits: sum of four 16-bit uniforms (Irwin-Hall), sigma ~= 2.31 after scaling; cond and
  uncond share a common base (real CFG branches are strongly correlated), and every drafted
  token is boosted in its parent's row so walks go several levels deep;
* EAGLE-2 dynamic trees built by the drafter's own rule (cnets_llamagen.py:731-912):
  top-k children per frontier node, global top ``total_tokens``, flat-index order;
* static-tree drafts shaped like ``topK_genrate_v1`` output (cnets_llamagen.py:943-1023).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def hash_u64(seed: int, n: int, stream: int = 0) -> np.ndarray:
    """n pseudo-random uint64 words, a pure function of (seed, stream, index)."""
    with np.errstate(over="ignore"):
        base = _splitmix64(np.asarray([(seed * 0x632BE59BD9B4E019 + stream * 0xD1342543DE82EF95)
                                       & 0xFFFFFFFFFFFFFFFF], dtype=np.uint64))[0]
        idx = np.arange(n, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
        return _splitmix64(idx + base)


def uniforms(seed: int, n: int, stream: int = 0) -> np.ndarray:
    """fp32 uniforms in [0,1) with 24 random bits (exactly representable)."""
    return ((hash_u64(seed, n, stream) >> np.uint64(40)).astype(np.float32)
            * np.float32(1.0 / 16777216.0))


def gauss(seed: int, shape, stream: int = 0, shift: int = 14) -> np.ndarray:
    """Approximately normal fp32 values, exact multiples of 2**-shift (sigma ~= 2.31 at 14)."""
    n = int(np.prod(shape))
    w = hash_u64(seed, n, stream)
    s = ((w & np.uint64(0xFFFF)).astype(np.int64) + ((w >> np.uint64(16)) & np.uint64(0xFFFF)).astype(np.int64)
         + ((w >> np.uint64(32)) & np.uint64(0xFFFF)).astype(np.int64)
         + ((w >> np.uint64(48)) & np.uint64(0xFFFF)).astype(np.int64) - 2 * 65535)
    return (s.astype(np.float32) * np.float32(2.0 ** -shift)).reshape(shape)


def permutation(seed: int, n: int, stream: int = 0) -> np.ndarray:
    return np.argsort(hash_u64(seed, n, stream), kind="stable")


def neighbor_table(seed: int, N: int, K: int, dtype=np.int32) -> np.ndarray:
    """A synthetic [N, K] table with the structural properties of the real one: every row
    holds K distinct ids, never the row's own id.  (Real tables come from
    ``lantern_b200.codebook``; this one is for acceptance-path tests and benches.)"""
    K = min(K, N - 1)
    perm = permutation(seed, N - 1, stream=7).astype(np.int64)
    rot = (hash_u64(seed, N, stream=8) % np.uint64(N - 1)).astype(np.int64)
    j = np.arange(K, dtype=np.int64)[None, :]
    x = np.arange(N, dtype=np.int64)[:, None]
    tbl = (x + 1 + perm[(j + rot[:, None]) % (N - 1)]) % N
    return tbl.astype(dtype)


# --------------------------------------------------------------------------- #
# Trees
# --------------------------------------------------------------------------- #
@dataclass
class Tree:
    parent: np.ndarray              # [T] int64, parent[0] = -1
    depth: np.ndarray               # [T] int64 (tree_position_ids)
    retrieve_indices: np.ndarray    # [L, D] int64, -1 padded, rows sorted (-1 last)
    tokens: Optional[np.ndarray] = None   # [T] int64

    @property
    def T(self) -> int:
        return int(self.parent.shape[0])


def retrieve_from_parents(parent: np.ndarray, sort_rows: bool = True):
    """Leaf paths, as cnets_llamagen.py:876-908 builds them from ``mask_index``."""
    T = parent.shape[0]
    depth = np.zeros(T, dtype=np.int64)
    for i in range(1, T):
        depth[i] = depth[parent[i]] + 1
    has_child = np.zeros(T, dtype=bool)
    has_child[parent[1:]] = True
    D = int(depth.max()) + 1
    rows = []
    for i in range(T):
        if has_child[i] and T > 1:
            continue
        row = [-1] * D
        c = i
        for j in range(int(depth[i]), -1, -1):
            row[j] = c
            c = int(parent[c])
        rows.append(row)
    if sort_rows:
        big = T + 5
        rows.sort(key=lambda r: [v if v >= 0 else big for v in r])
    return depth, np.asarray(rows, dtype=np.int64)


def eagle2_tree(seed: int, total_tokens: int = 59, depth: int = 4, top_k: int = 10) -> Tree:
    """Shape of the EAGLE-2 dynamic tree: ``depth`` expansion rounds of a ``top_k`` frontier,
    global top ``total_tokens - 1`` by cumulative score, kept in flat-candidate order."""
    n_keep = total_tokens - 1
    # per-child log-probs: sorted descending within a parent, like topk(log_softmax)
    def child_scores(stream, n_par):
        u = uniforms(seed, n_par * 2, stream).reshape(n_par, 2).astype(np.float64)
        c0 = 0.05 + 0.95 * u[:, :1]
        a = 0.15 + 1.05 * u[:, 1:]
        return -(c0 + a * np.arange(top_k, dtype=np.float64)[None, :])

    scores = [child_scores(100, 1)[0]]
    parents = [np.zeros(top_k, dtype=np.int64)]       # flat parent id (0 = root)
    frontier_flat = np.arange(top_k, dtype=np.int64) + 1
    frontier_score = scores[0].copy()
    base = 1 + top_k
    for d in range(depth):
        cs = child_scores(101 + d, top_k) + frontier_score[:, None]
        scores.append(cs.reshape(-1))
        parents.append(np.repeat(frontier_flat, top_k))
        pick = np.argsort(-cs.reshape(-1), kind="stable")[:top_k]
        frontier_flat = base + pick
        frontier_score = cs.reshape(-1)[pick]
        base += top_k * top_k
    all_scores = np.concatenate(scores)
    all_parents = np.concatenate(parents)
    n_keep = min(n_keep, all_scores.shape[0])
    keep = np.sort(np.argsort(-all_scores, kind="stable")[:n_keep]) + 1     # flat ids, ascending
    # ancestor-closure holds because a child's cumulative score never exceeds its parent's
    flat_to_node = {0: 0}
    for i, f in enumerate(keep):
        flat_to_node[int(f)] = i + 1
    parent = np.full(n_keep + 1, -1, dtype=np.int64)
    kept = []
    for i, f in enumerate(keep):
        p = int(all_parents[f - 1])
        if p not in flat_to_node:            # defensive: drop orphans (cannot happen with sorted scores)
            continue
        parent[i + 1] = flat_to_node[p]
        kept.append(i + 1)
    assert len(kept) == n_keep
    d_arr, ri = retrieve_from_parents(parent)
    return Tree(parent, d_arr, ri)


def random_tree(seed: int, T: int, max_depth: int = 6, max_children: int = 10) -> Tree:
    """Arbitrary tree with T nodes (sweeps: T from 1 to 256), BFS/level order."""
    parent = np.full(T, -1, dtype=np.int64)
    depth = np.zeros(T, dtype=np.int64)
    nchild = np.zeros(T, dtype=np.int64)
    r = hash_u64(seed, T, stream=21)
    for i in range(1, T):
        ok = [p for p in range(i) if depth[p] < max_depth - 1 and nchild[p] < max_children
              and (p == i - 1 or depth[p] >= depth[i - 1])]
        if not ok:
            ok = [p for p in range(i) if nchild[p] < max_children] or [0]
        p = ok[int(r[i] % np.uint64(len(ok)))]
        parent[i] = p
        depth[i] = depth[p] + 1
        nchild[p] += 1
    order = np.argsort(depth, kind="stable")          # relabel so parents precede children by level
    relabel = np.empty(T, dtype=np.int64)
    relabel[order] = np.arange(T)
    new_parent = np.full(T, -1, dtype=np.int64)
    for i in range(1, T):
        new_parent[relabel[i]] = relabel[parent[i]]
    d_arr, ri = retrieve_from_parents(new_parent)
    return Tree(new_parent, d_arr, ri)


def assign_tokens(seed: int, tree: Tree, lo: int, hi: int, root_token: Optional[int] = None) -> np.ndarray:
    """Random tokens in [lo, hi), distinct among siblings (top-k / sampling w/o replacement)."""
    T = tree.T
    r = hash_u64(seed, T * 4, stream=31)
    tok = np.zeros(T, dtype=np.int64)
    tok[0] = lo + int(r[0] % np.uint64(hi - lo)) if root_token is None else root_token
    used = {}
    for i in range(1, T):
        p = int(tree.parent[i])
        s = used.setdefault(p, set())
        t = 0
        while True:
            v = lo + int(r[(i * 4 + t) % (T * 4)] % np.uint64(hi - lo)) + (t // 4)
            v = lo + (v - lo) % (hi - lo)
            if v not in s:
                break
            t += 1
        s.add(v)
        tok[i] = v
    tree.tokens = tok
    return tok


def tree_logits(seed: int, tree: Tree, V: int, cfg: bool = True, boost: float = 13.0,
                dtype=np.float32):
    """cond/uncond logits [T, V]; every node's token is boosted in its parent's row."""
    T = tree.T
    base = gauss(seed, (T, V), stream=40)
    quarter = np.float32(0.25)
    cond = base + gauss(seed, (T, V), stream=41) * quarter
    uncond = base + gauss(seed, (T, V), stream=42) * quarter if cfg else None
    if tree.tokens is not None and boost:
        for i in range(1, T):
            p = int(tree.parent[i])
            cond[p, tree.tokens[i]] += np.float32(boost)
            if uncond is not None:
                uncond[p, tree.tokens[i]] += np.float32(boost)
    return cond.astype(dtype), (uncond.astype(dtype) if uncond is not None else None)


# --------------------------------------------------------------------------- #
# Static-tree drafts
# --------------------------------------------------------------------------- #
@dataclass
class StaticDraftSynth:
    ss_token: np.ndarray          # [n_groups_total, 10]
    ss_prob: np.ndarray           # [n_groups_total, 10] conditional probs (sample())
    op: List[np.ndarray]          # per level [n_groups_level, V]


def static_draft(seed: int, group_counts: Sequence[int], V: int, lo: int, hi: int,
                 top_k: int = 10, sharp: float = 1.0) -> StaticDraftSynth:
    """Per level, per expanded parent: a drafter distribution over [lo,hi) and ``top_k`` tokens
    drawn without replacement with conditional probabilities p_i / (1 - sum_{j<i} p_j)
    clamped to [0,1] (cnets_llamagen.py:924-940).  ``group_counts`` = parents per level,
    starting with the root level ([1, 4, 4, 1, 1] for mc_sim_7b_63)."""
    toks, probs, ops = [], [], []
    g = 0
    for lvl, n in enumerate(group_counts):
        lvl_op = np.zeros((n, V), dtype=np.float32)
        for r in range(n):
            z = gauss(seed, (hi - lo,), stream=1000 + g) * np.float32(sharp)
            z = z - z.max()
            e = np.exp(z.astype(np.float64))
            p = (e / e.sum()).astype(np.float32)
            lvl_op[r, lo:hi] = p
            # Gumbel-free sampling without replacement: exponential race with hashed uniforms
            u = uniforms(seed, hi - lo, stream=2000 + g).astype(np.float64)
            keys = -np.log(np.maximum(u, 1e-12)) / np.maximum(p.astype(np.float64), 1e-30)
            pick = np.argsort(keys, kind="stable")[:top_k]
            sp = p[pick]
            cum = np.concatenate([[0.0], np.cumsum(sp.astype(np.float32))[:-1]]).astype(np.float32)
            with np.errstate(divide="ignore", invalid="ignore"):
                cp = sp / (np.float32(1.0) - cum)
            cp[~np.isfinite(cp)] = -1
            cp = np.clip(cp, 0.0, 1.0).astype(np.float32)
            toks.append(pick + lo)
            probs.append(cp)
            g += 1
        ops.append(lvl_op)
    return StaticDraftSynth(np.asarray(toks, dtype=np.int64), np.asarray(probs, dtype=np.float32), ops)


def static_group_counts(tree_choices) -> List[int]:
    """Parents expanded per drafter level for a static tree (root level first)."""
    nodes = sorted((tuple(c) for c in tree_choices), key=lambda x: (len(x), x))
    maxd = max(len(p) for p in nodes)
    counts = [1]
    for lvl in range(1, maxd):
        counts.append(len({p[:-1] for p in nodes if len(p) == lvl + 1}))
    return counts


# --------------------------------------------------------------------------- #
# Drafter expansion outputs (inputs of the dynamic-tree post-processing)
# --------------------------------------------------------------------------- #
@dataclass
class Expansion:
    scores: np.ndarray     # [n_cand] fp32  cat(scores_list)           (cnets_llamagen.py:831)
    tokens: np.ndarray     # [n_cand] int64 cat(ss_token)              (:832)
    parents: np.ndarray    # [1 + depth * top_k] int64 cat(parents_list) (:840)
    sample_token: int
    top_k: int
    depth: int


def eagle2_expansion(seed: int, depth: int = 4, top_k: int = 10, lo: int = 0, hi: int = 16384) -> Expansion:
    """What the drafter's expansion loop (cnets_llamagen.py:766-821) leaves behind: per level the top_k children of
    each of the top_k frontier nodes with cumulative log-prob scores, the frontier chosen by cumulative score, and
    the `parents` bookkeeping (parent flat id + 1 per sibling group, bias rule of :788-791)."""
    rng = hash_u64(seed, 4 * top_k * top_k * (depth + 2), stream=61)
    cursor = [0]

    def logprobs(n_par):
        k = n_par * top_k
        u = ((rng[cursor[0]:cursor[0] + k] >> np.uint64(40)).astype(np.float64) + 0.5) / 16777216.0
        cursor[0] += k
        lp = np.log(u).reshape(n_par, top_k) * 0.8 - 0.05
        return -np.sort(-lp, axis=1).astype(np.float32)        # descending like topk(log_softmax)

    def toks(n_par):
        k = n_par * top_k
        t = lo + (rng[cursor[0]:cursor[0] + k] % np.uint64(hi - lo)).astype(np.int64)
        cursor[0] += k
        return t.reshape(n_par, top_k)

    scores_list = [logprobs(1)]
    tokens_list = [toks(1)]
    parents_list = [np.zeros(1, dtype=np.int64)]
    scores = scores_list[0][0].copy()
    topk_cs_index = np.arange(top_k)
    for i in range(depth):
        bias1 = top_k if i > 0 else 0
        bias2 = max(0, i - 1)
        bias = 1 + top_k ** 2 * bias2 + bias1
        parents_list.append((topk_cs_index + bias).astype(np.int64))
        cu = (logprobs(top_k) + scores[:, None]).astype(np.float32)
        tokens_list.append(toks(top_k))
        order = np.argsort(-cu.reshape(-1), kind="stable")[:top_k]
        topk_cs_index = order
        scores = cu.reshape(-1)[order]
        scores_list.append(cu)
    return Expansion(np.concatenate([s.reshape(-1) for s in scores_list]).astype(np.float32),
                     np.concatenate([t.reshape(-1) for t in tokens_list]).astype(np.int64),
                     np.concatenate(parents_list).astype(np.int64),
                     lo + int(rng[-1] % np.uint64(hi - lo)), top_k, depth)
