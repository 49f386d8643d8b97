"""Device-side dynamic (EAGLE-2) draft-tree construction — SURVEY.md 8(f) row N1.

``build_dynamic_tree`` replaces the tail of the reference drafter's ``topK_genrate``
(``models/drafters/cnets_llamagen.py:831-912``; same code in ``cnets_anole.py`` / ``cnets_lumina_mgpt.py:1330-1393``):
it takes the lists the expansion loop leaves behind and returns the same four objects, without Python loops or
``.tolist()`` round trips.  ``DynamicTree`` also carries the int32 / parent-pointer form that
``lantern_accept_fused`` consumes directly.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence

import torch

from . import _abi


@dataclass
class DynamicTree:
    tree_tokens: torch.Tensor     # [B, T] int32 (draft_tokens; root = sample token)
    parent: torch.Tensor          # [B, T] int32
    depth: torch.Tensor           # [B, T] int32 (tree_position_ids)
    mask: torch.Tensor            # [B, T, T] fp32 (tree_mask)
    retrieve: torch.Tensor        # [B, T, d_max] int32, -1 padded (rows beyond counts[b,0] are all -1)
    counts: torch.Tensor          # [B, 2] int32: n_leaves, max_depth + 1

    def reference_outputs(self, b: int = 0):
        """(draft_tokens [1,T] int64, retrieve_indices [L,D] int64, tree_mask [1,1,T,T] fp32, tree_position_ids [T])
        exactly as ``topK_genrate`` returns them; reads the two counts back (one small device->host copy)."""
        nl, D = (int(x) for x in self.counts[b].tolist())
        if nl < 0:
            raise ValueError(f"draft tree depth {D} exceeds d_max={self.retrieve.shape[2]}: rebuild with a larger d_max")
        return (self.tree_tokens[b:b + 1].long(), self.retrieve[b, :nl, :D].long(), self.mask[b][None, None],
                self.depth[b].long())


def build_dynamic_tree(scores: torch.Tensor, tokens: torch.Tensor, parents: torch.Tensor, sample_token: torch.Tensor,
                       total_tokens: int, top_k: int = 10, d_max: int = 8, sort_rows: bool = True,
                       want_mask: bool = True) -> DynamicTree:
    """scores / tokens: [B, n_cand] (``cat(scores_list).view(-1)``, ``cat(ss_token).view(-1)``), parents:
    [B, n_groups] (``cat(parents_list)``), sample_token: [B].  ``total_tokens`` = the drafter's ``self.total_tokens``
    (draft nodes, T - 1)."""
    lib = _abi.load()
    if scores.dim() == 1:
        scores, tokens, parents = scores[None], tokens[None], parents[None]
        sample_token = sample_token.reshape(1)
    dev = scores.device
    B, n = scores.shape
    T = total_tokens + 1
    sc = scores.to(torch.float32).contiguous()
    tk = tokens.to(device=dev, dtype=torch.int32).contiguous()
    pr = parents.to(device=dev, dtype=torch.int32).contiguous()
    rt = sample_token.to(device=dev, dtype=torch.int32).contiguous()
    out = DynamicTree(torch.empty(B, T, dtype=torch.int32, device=dev), torch.empty(B, T, dtype=torch.int32, device=dev),
                      torch.empty(B, T, dtype=torch.int32, device=dev),
                      torch.empty(B, T, T, dtype=torch.float32, device=dev) if want_mask else None,
                      torch.empty(B, T, d_max, dtype=torch.int32, device=dev),
                      torch.empty(B, 2, dtype=torch.int32, device=dev))
    _abi.check(lib.lantern_build_dynamic_tree(
        sc.data_ptr(), tk.data_ptr(), pr.data_ptr(), rt.data_ptr(), B, n, pr.shape[1], top_k, T, d_max, int(sort_rows),
        out.tree_tokens.data_ptr(), out.parent.data_ptr(), out.depth.data_ptr(),
        out.mask.data_ptr() if want_mask else None, out.retrieve.data_ptr(), out.counts.data_ptr(),
        torch.cuda.current_stream(dev).cuda_stream))
    return out


def from_drafter_lists(scores_list: Sequence[torch.Tensor], ss_token: Sequence[torch.Tensor],
                       parents_list: Sequence[torch.Tensor], sample_token: torch.Tensor, total_tokens: int,
                       top_k: int = 10, sort_rows: bool = True) -> DynamicTree:
    """Convenience for a patched ``topK_genrate``: takes the three python lists as the reference builds them.
    ``scores_list`` holds one entry per drafted level, so the deepest leaf path has ``len(scores_list) + 1`` nodes."""
    scores = torch.cat([s.reshape(-1) for s in scores_list])
    tokens = torch.cat([t.reshape(-1) for t in ss_token])
    parents = torch.cat([p.reshape(-1) for p in parents_list])
    return build_dynamic_tree(scores, tokens, parents, sample_token.reshape(-1)[:1], total_tokens, top_k,
                              d_max=max(8, len(scores_list) + 1), sort_rows=sort_rows)
