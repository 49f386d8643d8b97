"""Model.sample (models/drafters/cnets_llamagen.py:924-940) - SURVEY.md 8(f) row N3.  tests/golden/sample_cases.json
holds outputs of the LIVE reference with torch.multinomial replaced by the build's documented draw (exponential race
on the Philox stream), so the gather, the exclusive cumsum, p_i / (1 - sum_{j<i} p_j), the inf / nan patch-up, the
clamp and the full distribution are the reference's own arithmetic.  Here: the NumPy oracle against those outputs."""
import json
import os

import numpy as np
import pytest

from lantern_b200 import synth
from oracle import lantern_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "sample_cases.json")) as f:
    GOLD = json.load(f)


def sample_inputs(p):
    logits = (synth.gauss(p["seed"], (p["rows"], p["V"]), stream=9) * np.float32(2.5)).astype(np.float32)
    return logits, O.Warp(p["temperature"], p["top_p"], p["top_k"])


def cond_prob_tolerance(picked_probs: np.ndarray) -> np.ndarray:
    """1e-5 relative, plus the cancellation in 1 - cumsum: the denominator carries an absolute error of a few fp32 ulps
    of 1, which is relative error ulp / (1 - cumsum) in the quotient (it reaches 3e-6 when top_k == k)."""
    excl = np.concatenate([np.zeros((picked_probs.shape[0], 1)), np.cumsum(picked_probs.astype(np.float64), axis=1)[:, :-1]], axis=1)
    return 1e-5 + 4 * 1.2e-7 / np.maximum(1.0 - excl, 1e-7)


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: f"V{c['params']['V']}-k{c['params']['k']}-s{c['params']['seed']}")
def test_oracle_draft_sample_matches_reference(case):
    p = case["params"]
    logits, warp = sample_inputs(p)
    probe = np.asarray(case["probe_cols"])
    for r in range(p["rows"]):
        idx, cp, probs = O.draft_sample(logits[r], warp, p["k"], seed=p["seed"], step=3, row=r)
        assert idx.tolist() == case["indices"][r]
        want_cp = np.asarray(case["cond_probs"][r], dtype=np.float32)
        tol = cond_prob_tolerance(np.asarray([case["picked_probs"][r]]))[0]
        assert np.all(np.abs(cp - want_cp) <= tol * np.abs(want_cp) + 1e-30)
        assert np.all((cp >= 0) & (cp <= 1))
        want_pr = np.asarray(case["probe_probs"][r], dtype=np.float32)
        got_pr = probs[probe]
        assert np.array_equal(got_pr > 0, want_pr > 0)
        nz = want_pr > 0
        assert np.all(np.abs(got_pr[nz] - want_pr[nz]) <= 1e-5 * want_pr[nz])
        assert int((probs > 0).sum()) == case["nnz"][r]
        got_pk = probs[idx]
        want_pk = np.asarray(case["picked_probs"][r], dtype=np.float32)
        assert np.all(np.abs(got_pk - want_pk) <= 1e-5 * want_pk)


def test_sample_fixture_meta():
    assert GOLD["meta"]["n_cases"] == len(GOLD["cases"]) >= 8
    assert all(c["oracle_rel_err"]["cond_probs"] <= 1e-5 and c["oracle_rel_err"]["probs"] <= 1e-5 for c in GOLD["cases"])
