"""Greedy verification (temperature 0) on the GPU: lantern_accept_greedy against the golden outputs of the live
reference (tests/golden/greedy_cases.json: ea_model_anole.py:789-902, with and without the relaxation), against the
NumPy oracle on batches, and through the drop-in methods."""
import json
import os

import numpy as np
import pytest
import torch

import casegen as C
import cuda_runner as R
from lantern_b200 import posterior as PO
from lantern_b200 import verify

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "greedy_cases.json")) as f:
    GREEDY = json.load(f)


def _run(cases, dtype=torch.float32):
    b0 = cases[0]
    p = b0.params
    dev = torch.device("cuda")
    fam = R.family_spec(b0)
    table = torch.from_numpy(b0.table.astype(np.int32)).to(dev) if p["lantern"] else None
    k = min(int(p["lantern_k"]), b0.fam.ncols - 1)
    v = verify.Verifier(fam, cfg_scale=p["cfg_scale"], lantern=p["lantern"], lantern_k=k,
                        lantern_delta=p["lantern_delta"], nbr_table=table, device=dev)
    cond = torch.from_numpy(np.stack([c.cond for c in cases])).to(dev).to(dtype)
    uncond = torch.from_numpy(np.stack([c.uncond for c in cases])).to(dev).to(dtype)
    tokens = torch.from_numpy(np.stack([c.tree.tokens for c in cases]).astype(np.int32)).to(dev)
    retrieve = torch.from_numpy(R.pad_retrieve([c.tree.retrieve_indices for c in cases])).to(dev)
    res = v.greedy(cond, uncond, tokens, retrieve)
    torch.cuda.synchronize()
    return res


@pytest.mark.parametrize("case", GREEDY["cases"],
                         ids=lambda c: f"{'lantern' if c['params']['lantern'] else 'plain'}-s{c['params']['seed']}")
def test_greedy_matches_reference_golden(case):
    b = C.build(case["params"])
    res = _run([b])
    a = int(res.accept_length[0])
    assert (int(res.best_candidate[0]), a) == (case["best_candidate"], case["accept_length"])
    assert int(res.token[0]) == case["token"]
    best, oa, row, _ = C.oracle_greedy(b)
    got = res.sample_p[0].cpu().numpy()
    assert np.array_equal(got, row), "returned logits row differs from the reference's logits[best, accept_length]"
    ri = b.tree.retrieve_indices
    assert res.select_indices[0, :a + 1].tolist() == ri[best, :a + 1].tolist()
    assert res.path_tokens[0, :a + 1].tolist() == b.candidates[best, :a + 1].tolist()


@pytest.mark.parametrize("family,kw", [("anole", dict(ncols=2048, lantern_k=200, lantern_delta=0.2, boost=9.5)),
                                       ("llamagen", dict(ncols=4096, lantern_k=100, lantern_delta=0.1, boost=10.0)),
                                       ("anole", dict(ncols=1024, lantern_k=10, lantern_delta=5.0, boost=9.0)),
                                       ("llamagen", dict(ncols=2048, lantern=False, boost=10.0))])
def test_greedy_batched_matches_oracle(family, kw):
    built, orcs, seed = [], [], 83000
    while len(built) < 5:
        b = C.build(dict(family=family, top_k=0, temperature=0.0, seed=seed, **kw))
        seed += 1
        o = C.oracle_greedy(b)
        if o[3] >= 1e-5:
            built.append(b)
            orcs.append(o)
    res = _run(built)
    for i, (best, a, row, _) in enumerate(orcs):
        assert (int(res.best_candidate[i]), int(res.accept_length[i])) == (best, a)
        assert int(res.token[i]) == int(row.argmax())
        assert np.array_equal(res.sample_p[i].cpu().numpy(), row)
    assert len({int(x) for x in res.accept_length}) > 1     # the batch exercises different depths


def test_greedy_dropin_methods():
    """VerifyMixin.evaluate_posterior(logits_processor=None): fused handle and gathered logits, plain and relaxed."""
    class M(PO.VerifyMixin):
        lantern_family = "anole"
        lantern_image_tokens = 1024
    dev = torch.device("cuda")
    for lantern in (False, True):
        seed = 84000
        while True:
            b = C.build(dict(family="anole", ncols=1024, top_k=0, temperature=0.0, lantern=lantern, lantern_k=100,
                             lantern_delta=0.2, boost=9.0, seed=seed))
            seed += 1
            best, a, row, margin = C.oracle_greedy(b)
            if margin >= 1e-5 and a >= 1:
                break
        m = M()
        m.nearest_latents = None if b.table is None else b.table.astype(np.int64)
        m.image_token_offset = 4
        tl = torch.from_numpy(np.stack([b.cond, b.uncond])).to(dev)
        ri = torch.from_numpy(b.tree.retrieve_indices).to(dev)
        cand = torch.from_numpy(b.candidates).to(dev)
        handle = PO.TreeLogits(tl[:1], tl[1:2], float(b.params["cfg_scale"]), ri)
        gb, ga, grow = m.evaluate_posterior(handle, cand, None, lantern=lantern, lantern_k=100, lantern_delta=0.2)
        assert gb.is_cuda and gb.dtype == torch.int64 and (int(gb), int(ga)) == (best, a)
        assert np.array_equal(grow.cpu().numpy(), row)
        assert int(PO.sample_bonus_token(grow, do_sample=False)) == int(row.argmax())
        if lantern:     # gathered [L, D, V] logits, as the unfused tree_decoding returns them
            mixed = tl[1] + (tl[0] - tl[1]) * float(b.params["cfg_scale"])
            masked = torch.full_like(mixed, torch.finfo(torch.float32).min)
            masked[:, 4:4 + 1024] = mixed[:, 4:4 + 1024]
            gb2, ga2, grow2 = m.evaluate_posterior(masked[ri], cand, None, lantern=True, lantern_k=100, lantern_delta=0.2)
            assert (int(gb2), int(ga2)) == (best, a)


def test_greedy_static_tree_v1_method():
    """evaluate_posterior_v1(logits_processor=None) runs the same greedy code on a static tree (ea_model_anole.py:478-595)."""
    class M(PO.VerifyMixin):
        lantern_family = "anole"
        lantern_image_tokens = 1024
    dev = torch.device("cuda")
    seed = 85000
    while True:
        b = C.build(dict(family="anole", ncols=1024, top_k=0, temperature=0.0, lantern=True, lantern_k=50,
                         lantern_delta=0.3, boost=8.0, static_tree="mc_sim_7b_63", seed=seed))
        seed += 1
        best, a, row, margin = C.oracle_greedy(b)
        if margin >= 1e-5 and a >= 1:
            break
    m = M()
    m.nearest_latents = b.table.astype(np.int64)
    m.image_token_offset = 4
    tl = torch.from_numpy(np.stack([b.cond, b.uncond])).to(dev)
    ri = torch.from_numpy(np.asarray(b.tree.retrieve_indices)).to(dev)
    handle = PO.TreeLogits(tl[:1], tl[1:2], float(b.params["cfg_scale"]), ri)
    gb, ga, grow = m.evaluate_posterior_v1(handle, torch.from_numpy(b.candidates).to(dev), None, None, None, None, None,
                                           None, lantern=True, lantern_k=50, lantern_delta=0.3)
    assert (int(gb), int(ga)) == (best, a) and np.array_equal(grow.cpu().numpy(), row)
