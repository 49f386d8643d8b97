"""The torch restatement of the reference path (oracle/torch_path.py, timed by bench.py as the GPU-PyTorch baseline)
makes the same decisions as the numpy oracle, which is pinned to the live reference by tests/golden."""
import numpy as np
import pytest
import torch

import casegen as CG
from oracle import torch_path as TP


@pytest.mark.parametrize("params", [
    dict(family="llamagen", ncols=2048, top_k=300, lantern_k=100, seed=3),
    dict(family="llamagen", ncols=2048, top_k=300, lantern_k=100, lantern_delta=1.5, seed=4),
    dict(family="llamagen", ncols=2048, top_k=0, top_p=0.9, temperature=0.8, lantern_k=50, seed=5),
    dict(family="anole", ncols=1024, top_k=200, lantern_k=100, seed=6),
    dict(family="lumina_mgpt", ncols=1024, top_k=200, lantern_k=100, depth=5, seed=7),
    dict(family="lumina_mgpt", ncols=1024, top_k=200, lantern_k=100, depth=5, newline_depth=1, seed=8),
    dict(family="llamagen", ncols=2048, lantern=False, top_k=300, seed=9),
    dict(family="vanilla", ncols=2048, lantern=False, top_k=300, cfg=False, seed=10),
])
def test_torch_path_matches_oracle(params):
    tried = 0
    seed = params["seed"] * 100
    while tried < 3:
        b = CG.build(dict(params, seed=seed))
        seed += 1
        o = CG.oracle_step(b)
        if o.margin < 1e-5:
            continue
        tried += 1
        p = b.params
        k = min(int(p["lantern_k"]), b.fam.ncols - 1)
        best, a, tok, sp, n = TP.verify_step(
            torch.from_numpy(b.cond), None if b.uncond is None else torch.from_numpy(b.uncond), p["cfg_scale"],
            torch.from_numpy(np.asarray(b.tree.tokens)), torch.from_numpy(np.asarray(b.tree.retrieve_indices)),
            b.uniforms, b.fam, b.warp, p["lantern"], k, p["lantern_delta"], b.table, row_kinds=b.row_kinds)
        assert (best, a, tok, n) == (o.best_candidate, o.accept_length, o.token, o.n_uniforms)
        np.testing.assert_allclose(sp.numpy(), o.sample_p, rtol=2e-5, atol=1e-9)


@pytest.mark.parametrize("seed,depth", [(70, 5), (71, 4), (72, 6)])
def test_torch_dynamic_tree_tail_matches_oracle(seed, depth):
    from lantern_b200 import synth
    from oracle import lantern_oracle as O
    e = synth.eagle2_expansion(seed, depth=depth, top_k=10, lo=4, hi=8196)
    n_draft = 58
    toks, ri, mask, pos = TP.dynamic_tree_tail(torch.from_numpy(e.scores), torch.from_numpy(e.tokens),
                                               torch.from_numpy(e.parents), torch.tensor([e.sample_token]), n_draft, 10)
    o_tok, _, o_depth, o_mask, o_ri = O.dynamic_tree(e.scores, e.tokens, e.parents, e.sample_token, n_draft, 10)
    assert np.array_equal(toks[0].numpy(), o_tok) and np.array_equal(ri.numpy(), o_ri)
    assert np.array_equal(mask[0, 0].numpy(), o_mask) and np.array_equal(pos.numpy(), o_depth)


def test_torch_drafter_sample_law():
    """Conditional probabilities of the restated Model.sample: p_i / (1 - mass drawn before), clamped to [0, 1]."""
    from oracle import lantern_oracle as O
    torch.manual_seed(0)
    idx, cond, probs = TP.drafter_sample(torch.randn(4, 256) * 2, O.Warp(1.0, 1.0, 50), 6)
    p = probs.gather(-1, idx).double()
    before = torch.cumsum(p, -1) - p
    assert torch.allclose(cond.double(), (p / (1 - before)).clamp(0, 1), rtol=1e-5)
    assert (probs > 0).sum(-1).eq(50).all() and idx.shape == (4, 6)
