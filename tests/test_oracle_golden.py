"""The NumPy oracle against outputs of the live reference (tests/golden/gen_golden.py)."""
import json
import os

import numpy as np
import pytest

import casegen as C
from lantern_b200 import choices as CH
from oracle import lantern_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))

with open(os.path.join(HERE, "golden", "posterior_cases.json")) as f:
    GOLD = json.load(f)
with open(os.path.join(HERE, "golden", "tree_buffers.json")) as f:
    TREES = json.load(f)


with open(os.path.join(HERE, "golden", "greedy_cases.json")) as f:
    GREEDY = json.load(f)


def _id(c):
    p = c["params"]
    return f"{p['family']}-{p['static_tree'] or p['tree']}-s{p['seed']}"


def test_golden_meta():
    m = GOLD["meta"]
    assert m["n_cases"] == len(GOLD["cases"]) >= 90
    assert m["oracle_mismatches_at_generation"] == 0


@pytest.mark.parametrize("case", GOLD["cases"], ids=_id)
def test_oracle_matches_reference(case):
    b = C.build(case["params"])
    assert float(b.cond.astype(np.float64).sum()) == case["input_checksum"], "synthetic inputs drifted"
    r = C.oracle_step(b)
    assert r.best_candidate == case["best_candidate"]
    assert r.accept_length == case["accept_length"]
    assert r.n_uniforms - 1 == case["n_uniforms"]          # the oracle adds one bonus-token draw
    idx = np.asarray(case["sp_idx"])
    val = np.asarray(case["sp_val"], dtype=np.float32)
    got = r.sample_p[idx]
    assert int((r.sample_p > 0).sum()) == case["sp_nnz"]
    nz = val > 0
    assert np.all(got[~nz] == 0)
    assert np.max(np.abs(got[nz] - val[nz]) / val[nz]) <= 1e-5     # north_star: 1e-5 relative in fp32
    assert abs(float(r.sample_p.astype(np.float64).sum()) - case["sp_sum"]) <= 1e-5


@pytest.mark.parametrize("case", GREEDY["cases"], ids=lambda c: f"greedy-{'lantern' if c['params']['lantern'] else 'plain'}-s{c['params']['seed']}")
def test_oracle_greedy_matches_reference(case):
    """Greedy branches of the live reference (ea_model_anole.py:789-902), with and without the relaxation."""
    b = C.build(case["params"])
    assert float(b.cond.astype(np.float64).sum()) == case["input_checksum"], "synthetic inputs drifted"
    best, a, row, _ = C.oracle_greedy(b)
    assert (best, a) == (case["best_candidate"], case["accept_length"])
    assert int(row.argmax()) == case["token"]
    assert float(np.maximum(row.astype(np.float64), -1e30).sum()) == pytest.approx(case["row_sum"], rel=1e-12)


@pytest.mark.parametrize("name", CH.NAMES + CH.SYNTH_NAMES)
def test_tree_buffers_match_reference(name):
    ref = TREES[name]
    tb = O.generate_tree_buffers(CH.tree(name))
    assert tb["tree_indices"].tolist() == ref["tree_indices"]
    assert tb["tree_position_ids"].tolist() == ref["tree_position_ids"]
    assert tb["retrieve_indices"].tolist() == ref["retrieve_indices"]
    assert tb["tree_attn_mask"][0, 0].astype(np.int64).tolist() == ref["tree_attn_mask"]
    assert tb["p_indices"] == ref["p_indices"]
    assert tb["b_indices"] == ref["b_indices"]


def test_tree_sizes_table():
    # SURVEY.md section 4 table (probed from the reference)
    want = {"mc_sim_7b_63": (26, (15, 6)), "mc_sim_7b_63_balanced": (26, (16, 6)),
            "naive_extend_57": (58, (33, 6)), "medusa_2_7b_63": (64, (42, 5)),
            "reverse_balanced_25": (26, (15, 6)), "chain": (6, (1, 6))}
    for name, (T, shp) in want.items():
        tb = O.generate_tree_buffers(CH.tree(name))
        assert tb["tree_indices"].shape[0] == T and tb["retrieve_indices"].shape == shp


def test_philox_known_answer():
    # Random123 known-answer vectors for philox4x32-10
    assert O.philox4x32_10((0, 0, 0, 0), (0, 0)) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert O.philox4x32_10((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert O.philox4x32_10((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    u = O.philox_uniforms(1234, 7, 3, 9)
    assert u.dtype == np.float32 and np.all((u >= 0) & (u < 1))


def test_sample_token_rule():
    p = np.array([0.0, 0.25, 0.0, 0.5, 0.25], dtype=np.float32)
    assert O.sample_token(p, 0.0)[0] == 1
    assert O.sample_token(p, 0.2499)[0] == 1
    assert O.sample_token(p, 0.25)[0] == 3
    assert O.sample_token(p, 0.7499)[0] == 3
    assert O.sample_token(p, 0.75)[0] == 4
    assert O.sample_token(p, 0.9999999)[0] == 4


def test_kv_compact_overlap():
    slab = np.arange(2 * 3 * 10 * 4, dtype=np.float32).reshape(2, 3, 10, 4)
    want = slab.copy()
    sel = np.array([4, 6, 7])
    want[..., 4:7, :] = slab[..., sel, :]
    n = O.kv_compact(slab, sel, 4)
    assert n == 7 and np.array_equal(slab, want)


def test_neighbor_table_small():
    rng = np.random.default_rng(0)
    E = rng.standard_normal((64, 8)).astype(np.float32)
    t = O.neighbor_table(E)
    assert t.shape == (64, 63)
    d = ((E[:, None, :].astype(np.float64) - E[None].astype(np.float64)) ** 2).sum(-1)
    for i in range(64):
        assert i not in t[i]
        assert sorted(t[i].tolist()) == [j for j in range(64) if j != i]
        dd = d[i, t[i]]
        assert np.all(np.diff(dd) >= -1e-12)
