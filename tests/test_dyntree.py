"""Dynamic (EAGLE-2) tree post-processing: oracle vs the live reference's own code (tests/golden/dynamic_trees.json),
and the CUDA builder vs both."""
import json
import os

import numpy as np
import pytest
import torch

from lantern_b200 import synth
from oracle import lantern_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "dynamic_trees.json")) as f:
    GOLD = json.load(f)


def _id(c):
    return f"s{c['seed']}-d{c['depth']}-t{c['total_tokens']}-{'sorted' if c['sort_rows'] else 'raw'}"


@pytest.mark.parametrize("case", GOLD, ids=_id)
def test_oracle_dynamic_tree_matches_reference(case):
    ex = synth.eagle2_expansion(case["seed"], depth=case["depth"], top_k=10)
    toks, parent, depth, mask, ri = O.dynamic_tree(ex.scores, ex.tokens, ex.parents, ex.sample_token,
                                                   case["total_tokens"], ex.top_k, case["sort_rows"])
    assert toks.tolist() == case["draft_tokens"]
    assert depth.tolist() == case["tree_position_ids"]
    assert mask.astype(np.int64).tolist() == case["tree_mask"]
    assert ri.tolist() == case["retrieve_indices"]


@pytest.mark.gpu
def test_cuda_dynamic_tree_matches_reference_batched():
    from lantern_b200 import dyntree
    by_shape = {}
    for c in GOLD:
        by_shape.setdefault((c["depth"], c["total_tokens"], c["sort_rows"]), []).append(c)
    for (depth, total, sort_rows), cases in by_shape.items():
        exs = [synth.eagle2_expansion(c["seed"], depth=depth, top_k=10) for c in cases]
        sc = torch.from_numpy(np.stack([e.scores for e in exs])).cuda()
        tk = torch.from_numpy(np.stack([e.tokens for e in exs])).cuda()
        pr = torch.from_numpy(np.stack([e.parents for e in exs])).cuda()
        st = torch.tensor([e.sample_token for e in exs]).cuda()
        tree = dyntree.build_dynamic_tree(sc, tk, pr, st, total, top_k=10, d_max=depth + 3, sort_rows=sort_rows)
        torch.cuda.synchronize()
        for b, c in enumerate(cases):
            toks, ri, mask, pos = tree.reference_outputs(b)
            assert toks[0].tolist() == c["draft_tokens"]
            assert pos.tolist() == c["tree_position_ids"]
            assert ri.tolist() == c["retrieve_indices"]
            assert mask[0, 0].long().tolist() == c["tree_mask"]
            # parent pointers agree with the mask: parent = deepest proper ancestor
            par = tree.parent[b].tolist()
            assert par[0] == -1 and all(0 <= par[i] < i for i in range(1, total + 1))


@pytest.mark.gpu
def test_dynamic_tree_feeds_the_fused_step():
    """Builder output (int32, -1 padded [B, T, d_max]) goes straight into lantern_accept_fused."""
    import casegen as C
    import cuda_runner as R
    from lantern_b200 import dyntree, verify
    from oracle import lantern_oracle as OO
    fam = verify.LLAMAGEN.resized(4096)
    B, total = 4, 58
    exs = [synth.eagle2_expansion(700 + i, depth=4, top_k=10, lo=0, hi=4096) for i in range(B)]
    tree = dyntree.build_dynamic_tree(torch.from_numpy(np.stack([e.scores for e in exs])).cuda(),
                                      torch.from_numpy(np.stack([e.tokens for e in exs])).cuda(),
                                      torch.from_numpy(np.stack([e.parents for e in exs])).cuda(),
                                      torch.tensor([e.sample_token for e in exs]).cuda(), total, d_max=7)
    T = total + 1
    cond = synth.gauss(5, (B, T, 4096), stream=1)
    uncond = cond + synth.gauss(5, (B, T, 4096), stream=2) * np.float32(0.25)
    table = synth.neighbor_table(0, 4096, 101)
    uni = synth.uniforms(9, B * (T + 1), stream=3).reshape(B, T + 1)
    ver = verify.Verifier(fam, top_k=500, cfg_scale=3.0, lantern=True, lantern_k=100, lantern_delta=0.1,
                          nbr_table=torch.from_numpy(table).cuda())
    res = ver.step(torch.from_numpy(cond).cuda(), torch.from_numpy(uncond).cuda(), tree.tree_tokens, tree.retrieve,
                   uniforms=torch.from_numpy(uni).cuda(), want_sample_p=True)
    torch.cuda.synchronize()
    ofam = OO.small_family(OO.LLAMAGEN, 4096)
    for b in range(B):
        toks, ri, _, _ = tree.reference_outputs(b)
        o = OO.verify_step(cond[b], uncond[b], 3.0, toks[0].cpu().numpy(), ri.cpu().numpy(), uni[b], ofam,
                           OO.Warp(1.0, 1.0, 500), True, 100, 0.1, table)
        if o.margin >= 1e-5:
            R.compare(res, b, o)
