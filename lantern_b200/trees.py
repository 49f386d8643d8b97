"""Host-side tree logic: static-tree buffers and the kernel-ready tree description.

``generate_tree_buffers`` returns the same dictionary as the reference's
``models/drafters/utils.py:80-217`` (``tree_attn_mask``, ``tree_indices``, ``tree_position_ids``,
``retrieve_indices``, ``p_indices``, ``b_indices``) so existing ``generate()`` loops keep working, plus
``static_tree``: the parent-pointer / CSR form the CUDA walk consumes.  Built from a trie over the
sorted paths rather than the reference's repeated list scans.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch

from .verify import StaticTree

TOPK = 10  # children drafted per expanded node (models/drafters/utils.py:13)


class _Node:
    __slots__ = ("path", "pos", "parent", "children", "group", "level_group", "flat")

    def __init__(self, path, pos, parent):
        self.path, self.pos, self.parent = path, pos, parent
        self.children: List["_Node"] = []
        self.group = -1         # global index of the sibling group (order of expansion)
        self.level_group = -1   # index of the sibling group inside its level (p_indices)
        self.flat = 0           # index into cat(sample_token, ss_token.view(-1)) (tree_indices)


def _build(tree_choices: Sequence[Sequence[int]]):
    paths = sorted((tuple(p) for p in tree_choices), key=lambda p: (len(p), p))
    root = _Node((), 0, None)
    by_path = {(): root}
    nodes = [root]
    for i, p in enumerate(paths):
        par = by_path[p[:-1]]
        n = _Node(p, i + 1, par)
        par.children.append(n)
        by_path[p] = n
        nodes.append(n)
    # sibling groups are numbered in node order: a new group starts whenever the parent changes
    g = -1
    last_parent, last_depth, lg = None, 0, -1
    for n in nodes[1:]:
        d = len(n.path)
        if d != last_depth:
            lg = -1
            last_depth = d
            last_parent = None
        if n.parent is not last_parent:
            g += 1
            lg += 1
            last_parent = n.parent
        n.group, n.level_group = g, lg
        n.flat = n.path[-1] + TOPK * g + 1
    return nodes


def generate_tree_buffers(tree_choices: Sequence[Sequence[int]], device="cuda") -> Dict:
    nodes = _build(tree_choices)
    T = len(nodes)
    depth = [len(n.path) for n in nodes]
    mask = torch.eye(T)
    mask[:, 0] = 1
    for n in nodes[1:]:
        a = n.parent
        while a is not None and a.pos != 0:
            mask[n.pos, a.pos] = 1
            a = a.parent
    # leaf paths; rows ordered lexicographically by node position with padding last
    rows = []
    for n in nodes:
        if n.children or (T > 1 and n.pos == 0):
            continue
        chain = []
        a = n
        while a is not None:
            chain.append(a.pos)
            a = a.parent
        rows.append(chain[::-1])
    D = max(len(r) for r in rows)
    rows.sort(key=lambda r: r + [T + 5] * (D - len(r)))
    ri = [r + [-1] * (D - len(r)) for r in rows]

    level_group = [-1] + [n.level_group for n in nodes[1:]]
    earlier = [[]] + [[s.pos for s in n.parent.children if s.pos < n.pos] for n in nodes[1:]]
    p_indices = [[level_group[v] for v in row] for row in ri]          # -1 pads read the last node
    b_indices = [[(torch.tensor(earlier[v], device=device) if (v != -1 and earlier[v]) else [])
                  for v in row] for row in ri]

    # kernel-ready form
    n_groups = max(n.group for n in nodes[1:]) + 1 if T > 1 else 0
    level_first_group: Dict[int, int] = {}
    for n in nodes[1:]:
        level_first_group.setdefault(len(n.path), n.group)
    qrow = [0] + [n.group for n in nodes[1:]]     # op rows are stored level by level == group order
    sib_off, sib_idx = [0], []
    for v in range(T):
        sib_idx.extend(earlier[v])
        sib_off.append(len(sib_idx))
    st = StaticTree(
        retrieve=torch.tensor(ri, dtype=torch.int32, device=device),
        node_qrow=torch.tensor(qrow, dtype=torch.int32, device=device),
        sib_off=torch.tensor(sib_off, dtype=torch.int32, device=device),
        sib_idx=torch.tensor(sib_idx if sib_idx else [0], dtype=torch.int32, device=device),
        n_q_rows=n_groups,
    )
    return {
        "tree_attn_mask": mask[None, None].to(device),
        "tree_indices": torch.tensor([n.flat for n in nodes], dtype=torch.long, device=device),
        "tree_position_ids": torch.tensor(depth, dtype=torch.long, device=device),
        "retrieve_indices": torch.tensor(ri, dtype=torch.long, device=device),
        "p_indices": p_indices,
        "b_indices": b_indices,
        "static_tree": st,
        "parents": [-1] + [n.parent.pos for n in nodes[1:]],
        "group_counts": _group_counts(nodes),
    }


def _group_counts(nodes) -> List[int]:
    counts: Dict[int, set] = {}
    for n in nodes[1:]:
        counts.setdefault(len(n.path), set()).add(n.group)
    return [len(counts[d]) for d in sorted(counts)]


def generate_candidates(tree_logits, tree_indices, retrieve_indices, sample_token, logits_processor=None):
    """ea_model_llamagen.py:676-706 / ea_model_lumina_mgpt.py:525-554: map the drafter's
    ``(ss_token, ss_prob, ss_op)`` onto tree order.  Pure index gathers (torch, any device)."""
    sample_token = sample_token.to(tree_indices.device)
    flat = torch.cat([sample_token[0].view(-1)[:1], tree_logits[0].reshape(-1)], dim=-1)
    tree_candidates = flat[tree_indices]
    ext = torch.cat([tree_candidates, tree_candidates.new_full((1,), -1)])
    cart_candidates = ext[retrieve_indices]
    cart_prob = None
    if len(tree_logits) > 1 and tree_logits[1] is not None:
        probs = torch.cat([tree_logits[1].new_ones(1, dtype=torch.float32), tree_logits[1].reshape(-1).float()])
        tp = probs[tree_indices]
        cart_prob = torch.cat([tp, tp.new_ones(1)])[retrieve_indices]
    return cart_candidates, cart_prob, tree_candidates.unsqueeze(0)
