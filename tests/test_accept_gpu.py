"""Parity of the fused CUDA verify step (through the C ABI) with the reference-pinned oracle."""
import json
import os

import numpy as np
import pytest
import torch

import casegen as C
import cuda_runner as R
from lantern_b200 import verify
from oracle import lantern_oracle as O

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "posterior_cases.json")) as f:
    GOLD = json.load(f)

MARGIN = 1e-5


def _id(c):
    p = c["params"]
    return f"{p['family']}-{p['static_tree'] or p['tree']}-s{p['seed']}"


def _supported(p):
    return True


@pytest.mark.parametrize("case", [c for c in GOLD["cases"] if _supported(c["params"])], ids=_id)
def test_golden_case_matches_reference(case):
    """CUDA path == live-reference outputs recorded in tests/golden (and == the oracle for the token)."""
    b = C.build(case["params"])
    orc = C.oracle_step(b)
    res = R.run_cases([b])
    assert int(res.accept_length[0]) == case["accept_length"]
    assert int(res.best_candidate[0]) == case["best_candidate"]
    assert int(res.n_draws[0]) - 1 == case["n_uniforms"]
    R.compare(res, 0, orc)
    idx = np.asarray(case["sp_idx"])
    val = np.asarray(case["sp_val"], dtype=np.float32)
    got = res.sample_p[0].cpu().numpy()[idx]
    atol = 1e-6 * float(val.max())
    assert np.all(np.abs(got - val) <= 1e-5 * val + atol)


@pytest.mark.parametrize("family,kw", [
    ("llamagen", dict()),
    ("llamagen", dict(lantern_k=10, lantern_delta=5.0, ncols=4096, top_k=500, boost=11.0)),
    ("anole", dict(ncols=2048, top_k=400, boost=11.0)),
    ("lumina_mgpt", dict(ncols=2048, top_k=400, depth=5, boost=11.0, newline_depth=1)),
])
def test_batched_dynamic(family, kw):
    """Many prompts in one launch (ragged trees padded with -1 rows) == per-item oracle."""
    built, orcs = [], []
    seed = 5000
    skipped = 0
    while len(built) < 12:
        b = C.build(dict(family=family, seed=seed, **kw))
        seed += 1
        o = C.oracle_step(b)
        if o.margin < MARGIN:
            skipped += 1
            continue
        built.append(b)
        orcs.append(o)
    res = R.run_cases(built)
    for i, o in enumerate(orcs):
        R.compare(res, i, o)
    assert skipped < 12


@pytest.mark.parametrize("tree", ["mc_sim_7b_63", "medusa_2_7b_63", "chain"])
@pytest.mark.parametrize("family", ["llamagen", "anole", "lumina_mgpt"])
def test_batched_static(family, tree):
    built, orcs = [], []
    seed = 7000
    while len(built) < 6:
        b = C.build(dict(family=family, seed=seed, ncols=2048, static_tree=tree, lantern_k=10, lantern_delta=10.0,
                         top_k=400, boost=8.0))
        seed += 1
        o = C.oracle_step(b)
        if o.margin >= MARGIN:
            built.append(b)
            orcs.append(o)
    res = R.run_cases(built)
    for i, o in enumerate(orcs):
        R.compare(res, i, o)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_half_precision_inputs(dtype):
    """bf16/fp16 logits are widened exactly and then follow the fp32 contract: feed the oracle the rounded values."""
    built, orcs = [], []
    seed = 9000
    while len(built) < 6:
        b = C.build(dict(family="llamagen", seed=seed, ncols=4096, top_k=500, boost=11.0))
        seed += 1
        b.cond = torch.from_numpy(b.cond).to(dtype).float().numpy()
        b.uncond = torch.from_numpy(b.uncond).to(dtype).float().numpy()
        o = C.oracle_step(b)
        if o.margin >= MARGIN:
            built.append(b)
            orcs.append(o)
    res = R.run_cases(built, dtype=dtype)
    for i, o in enumerate(orcs):
        R.compare(res, i, o)


@pytest.mark.parametrize("T", [1, 2, 3, 9, 64, 200])
def test_tree_size_edges(T):
    """Root-only tree (reference raises UnboundLocalError; defined here as accept 0 + fresh root row) up to 200 nodes."""
    b = C.build(dict(family="llamagen", ncols=1024, tree="random", total_tokens=T, top_k=100, lantern_k=64,
                     boost=9.0, seed=300 + T))
    o = C.oracle_step(b)
    res = R.run_cases([b])
    if o.margin >= MARGIN:
        R.compare(res, 0, o)
    else:
        assert 0 <= int(res.accept_length[0]) < b.tree.retrieve_indices.shape[1]


def test_unaligned_window_scalar_path():
    """ncols not a multiple of 4 exercises the scalar-load path."""
    b = C.build(dict(family="llamagen", ncols=1001, top_k=100, lantern_k=64, boost=9.0, seed=77))
    o = C.oracle_step(b)
    res = R.run_cases([b])
    if o.margin >= MARGIN:
        R.compare(res, 0, o)


def test_philox_stream_matches_host_copy():
    """Device Philox draws == lantern_philox_uniforms == the oracle's philox_uniforms."""
    from lantern_b200 import verify
    from oracle import lantern_oracle as O
    b = C.build(dict(family="llamagen", ncols=2048, top_k=300, lantern_k=100, boost=11.0, seed=4242))
    T = b.tree.T
    seed, step = 0x1234_5678_9ABC, 17
    host = verify.philox_uniforms(seed, step, 0, T + 1)
    assert np.array_equal(host, O.philox_uniforms(seed, step, 0, T + 1))
    b.uniforms = host
    o = C.oracle_step(b)
    res = R.run_cases([b], philox=(seed, step))
    if o.margin >= MARGIN:
        R.compare(res, 0, o)


def test_topk_ties_and_constant_rows():
    """Adversarial rows for the top-k select: heavy ties, a constant row, outliers."""
    b = C.build(dict(family="llamagen", ncols=2048, top_k=300, lantern_k=100, boost=11.0, seed=31337))
    rng = np.random.default_rng(0)
    b.cond = np.round(b.cond * 2) / 2          # quantised -> many exact ties at the threshold
    b.uncond = np.round(b.uncond * 2) / 2
    b.cond[3] = 1.0
    b.uncond[3] = 1.0                          # constant row
    b.cond[5, :10] += 1e4                      # outliers
    o = C.oracle_step(b)
    res = R.run_cases([b])
    if o.margin >= MARGIN:
        R.compare(res, 0, o)


@pytest.mark.parametrize("top_k,top_p", [(50, 1.0), (0, 1.0), (2000, 0.9)])
def test_vanilla_vocab_65536(top_k, top_p):
    """Plain EAGLE verification (drafters/utils.py:333-410) on a 65536-entry vocabulary: wider than the
    register-resident row limit, so the multi-pass statistics kernel runs and the walk keeps its probability vector in
    global memory."""
    built, orcs, seed = [], [], 56000
    while len(built) < 3:
        b = C.build(dict(family="vanilla", ncols=65536, cfg=False, lantern=False, top_k=top_k, top_p=top_p, boost=13.0,
                         total_tokens=20, seed=seed))
        seed += 1
        o = C.oracle_step(b)
        if o.margin >= MARGIN:
            built.append(b)
            orcs.append(o)
    res = R.run_cases(built)
    for i, o in enumerate(orcs):
        R.compare(res, i, o)


def test_error_paths():
    from lantern_b200 import _abi, verify
    with pytest.raises(ValueError):
        verify.Verifier(verify.LLAMAGEN, lantern=True)      # no table
    with pytest.raises(ValueError):
        verify.Verifier(verify.LLAMAGEN, temperature=0.0)   # greedy has no sampling walk


@pytest.mark.parametrize("top_p,top_k,temp", [(0.9, 0, 1.0), (0.5, 300, 1.0), (0.95, 2000, 0.8), (0.3, 0, 1.3)])
def test_top_p(top_p, top_k, temp):
    """HF order Temperature -> TopP -> TopK (drafters/utils.py:43-51), ties at the nucleus boundary by column."""
    built, orcs, seed = [], [], 31000
    while len(built) < 6:
        b = C.build(dict(family="llamagen", seed=seed, ncols=4096, top_p=top_p, top_k=top_k, temperature=temp,
                         lantern_k=50, boost=11.0))
        seed += 1
        if seed % 2:
            b.cond = np.round(b.cond * 4) / 4                # exact ties inside the row
            b.uncond = np.round(b.uncond * 4) / 4
        o = C.oracle_step(b)
        if o.margin >= MARGIN:
            built.append(b)
            orcs.append(o)
    res = R.run_cases(built)
    for i, o in enumerate(orcs):
        R.compare(res, i, o)


@pytest.mark.parametrize("family,kw", [
    ("llamagen", dict()),
    ("llamagen", dict(ncols=4096, top_k=500, lantern_k=10, lantern_delta=5.0, boost=11.0)),
    ("anole", dict()),
    ("lumina_mgpt", dict(depth=5, newline_depth=1)),
    ("llamagen", dict(ncols=4096, static_tree="mc_sim_7b_63", lantern_k=10, lantern_delta=10.0, top_k=500, boost=8.5)),
    ("lumina_mgpt", dict(ncols=2048, depth=5, top_k=400, static_tree="mc_sim_7b_63", lantern_k=10, lantern_delta=5.0,
                         boost=8.0)),
    ("anole", dict(ncols=2048, static_tree="mc_sim_7b_63_balanced", top_k=400, lantern_k=10, lantern_delta=10.0,
                   boost=8.0)),
    ("llamagen", dict(ncols=2048, lantern=False, top_k=300, boost=10.0)),
])
def test_lazy_statistics_mode(family, kw):
    """phases = 6: no streamed statistics kernel; the walk computes the statistics of the rows it visits."""
    built, orcs, seed = [], [], 41000
    while len(built) < 6:
        b = C.build(dict(family=family, seed=seed, **kw))
        seed += 1
        o = C.oracle_step(b)
        if o.margin >= MARGIN:
            built.append(b)
            orcs.append(o)
    res = R.run_cases(built, phases=6)
    for i, o in enumerate(orcs):
        R.compare(res, i, o)
    eager = R.run_cases(built, phases=3)
    assert torch.equal(res.accept_length, eager.accept_length) and torch.equal(res.token, eager.token)


def test_automatic_schedule_matches_both():
    """phases = 8: the library picks streamed or lazy per batch; either way the results are the oracle's.  Covers a
    generic window, small trees in a small batch (streamed), 59-row trees in a small batch (lazy since round 2: from 48
    rows per prompt on) and a batch past the 2048-row switch."""
    for kw, n in ((dict(family="llamagen", ncols=2048, top_k=300, lantern_k=100), 3),
                  (dict(family="lumina_mgpt", ncols=4096, top_k=500, lantern_k=100, depth=5, total_tokens=26), 3),
                  (dict(family="lumina_mgpt", ncols=4096, top_k=500, lantern_k=100, depth=5), 3),
                  (dict(family="lumina_mgpt", ncols=4096, top_k=500, lantern_k=100, depth=5), 40)):
        built, orcs, seed = [], [], 47000
        while len(built) < min(n, 6):
            b = C.build(dict(seed=seed, **kw))
            seed += 1
            o = C.oracle_step(b)
            if o.margin >= MARGIN:
                built.append(b)
                orcs.append(o)
        cases = [built[i % len(built)] for i in range(n)]      # 40 x 59 rows > 2048: the lazy side of the switch
        res = R.run_cases(cases, phases=8)
        for i in range(n):
            R.compare(res, i, orcs[i % len(built)])


def test_vanilla_llm_vocab_32000():
    """Plain EAGLE verification on an LLM-sized vocabulary (non power of two -> generic statistics kernel)."""
    built, orcs, seed = [], [], 52000
    while len(built) < 3:
        b = C.build(dict(family="vanilla", ncols=32000, cfg=False, lantern=False, top_k=50, boost=12.0, seed=seed))
        seed += 1
        o = C.oracle_step(b)
        if o.margin >= MARGIN:
            built.append(b)
            orcs.append(o)
    res = R.run_cases(built)
    for i, o in enumerate(orcs):
        R.compare(res, i, o)


@pytest.mark.parametrize("dtype", [torch.bfloat16])
def test_lumina_bf16_window_misaligned_for_16_bytes(dtype):
    """bf16 logits with the image window starting at column 4: 8-byte aligned only (TMA copies start 8 bytes early)."""
    built, orcs, seed = [], [], 53000
    while len(built) < 4:
        b = C.build(dict(family="lumina_mgpt", depth=5, seed=seed))
        seed += 1
        b.cond = torch.from_numpy(b.cond).to(dtype).float().numpy()
        b.uncond = torch.from_numpy(b.uncond).to(dtype).float().numpy()
        o = C.oracle_step(b)
        if o.margin >= MARGIN:
            built.append(b)
            orcs.append(o)
    res = R.run_cases(built, dtype=dtype)
    for i, o in enumerate(orcs):
        R.compare(res, i, o)


@pytest.mark.parametrize("dist", ["gauss", "student_t2", "bimodal", "lognormal", "quantised", "mixed_rows"])
@pytest.mark.parametrize("ncols,top_k", [(8192, 2000), (16384, 2000), (4096, 50)])
def test_row_statistics_exact_on_non_gaussian_rows(dist, ncols, top_k):
    """The tracked-quantile bracket is a speed heuristic; the threshold must stay the exact k-th largest (ties kept)
    and the softmax statistics right for any row shape.  Reads the RowStats workspace of a phases=1 launch."""
    _check_row_statistics(dist, ncols, top_k, B=5)


@pytest.mark.parametrize("dist", ["gauss", "quantised", "mixed_rows"])
@pytest.mark.parametrize("ncols,top_k", [(8192, 2000), (16384, 2000)])
def test_row_statistics_exact_with_several_rows_per_cta(dist, ncols, top_k):
    """Same check with 1320 rows on the 296 (or 148) persistent CTAs: every CTA streams four to nine rows, so the
    pipelined bulk copies, the row advance and the bracket that tracks the previous row's quantile all take part
    (rows of different shapes follow each other in `mixed_rows`)."""
    _check_row_statistics(dist, ncols, top_k, B=40)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_row_statistics_bitwise_reproducible(dtype):
    """The streaming kernel hands rows from the main warps to the select warps through mbarriers (double-buffered);
    whatever the timing, every launch must produce bit-identical records (threshold, maximum AND sum)."""
    rng = np.random.default_rng(11)
    B, T, ncols, top_k = 40, 33, 8192, 2000
    x = rng.standard_normal((B, T, ncols)) * 2.5
    x[:, ::7] = rng.standard_t(2, (B, len(range(0, T, 7)), ncols))          # some rows miss the bracket: redo path
    cond = torch.from_numpy(x.astype(np.float32)).cuda().to(dtype)
    uncond = torch.from_numpy((x + rng.standard_normal((B, T, ncols)) * 0.7).astype(np.float32)).cuda().to(dtype)
    fam = verify.LLAMAGEN.resized(ncols)
    v = verify.Verifier(fam, temperature=1.0, top_k=top_k, cfg_scale=3.0, lantern=False, device=torch.device("cuda"))
    tokens = torch.zeros(B, T, dtype=torch.int32, device="cuda")
    retrieve = torch.zeros(B, 1, 1, dtype=torch.int32, device="cuda")
    uni = torch.rand(B, 2, device="cuda")
    first = None
    junk = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    for rep in range(6):
        if rep % 2:
            junk.random_(0, 255)                     # perturb cache state / timing between launches
        v.step(cond, uncond, tokens, retrieve, uniforms=uni, phases=1)
        torch.cuda.synchronize()
        stats = v._work[:B * T * 32].view(torch.int32).view(B * T, 8)[:, :3].clone()
        if first is None:
            first = stats
        else:
            assert torch.equal(stats, first), f"launch {rep}: {int((stats != first).any(1).sum())} rows differ"


def _check_row_statistics(dist, ncols, top_k, B):
    rng = np.random.default_rng(hash((dist, ncols)) % 2**32)
    T = 33
    shape = (B, T, ncols)
    if dist == "gauss":
        x = rng.standard_normal(shape) * 2.5
    elif dist == "student_t2":
        x = rng.standard_t(2, shape)
    elif dist == "bimodal":
        x = np.where(rng.random(shape) < 0.3, rng.standard_normal(shape) + 6, rng.standard_normal(shape) * 0.3 - 4)
    elif dist == "lognormal":
        x = rng.lognormal(0.0, 1.5, shape)
    elif dist == "quantised":
        x = np.round(rng.standard_normal(shape) * 3) / 2.0                  # ~30 distinct values: massive ties
    else:                                                                    # the shape changes from row to row
        x = rng.standard_normal(shape) * rng.uniform(0.1, 8.0, (B, T, 1)) + rng.uniform(-20, 20, (B, T, 1))
        x[:, ::5] = rng.standard_t(1.5, (B, len(range(0, T, 5)), ncols))
    cond = x.astype(np.float32)
    uncond = (x + rng.standard_normal(shape) * 0.7).astype(np.float32)
    fam = verify.LLAMAGEN.resized(ncols)
    v = verify.Verifier(fam, temperature=1.0, top_k=top_k, cfg_scale=3.0, lantern=False, device=torch.device("cuda"))
    tokens = torch.zeros(B, T, dtype=torch.int32, device="cuda")
    retrieve = torch.zeros(B, 1, 1, dtype=torch.int32, device="cuda")
    v.step(torch.from_numpy(cond).cuda(), torch.from_numpy(uncond).cuda(), tokens, retrieve,
           uniforms=torch.rand(B, 2, device="cuda"), phases=1)
    torch.cuda.synchronize()
    stats = v._work[:B * T * 32].view(torch.float32).view(B * T, 8).cpu().numpy()
    s = O.cfg_mix(cond.reshape(-1, ncols), uncond.reshape(-1, ncols), 3.0)
    kth = np.partition(s, ncols - top_k, axis=1)[:, ncols - top_k]
    assert np.array_equal(stats[:, 0], kth), f"{int((stats[:, 0] != kth).sum())} thresholds differ"
    assert np.array_equal(stats[:, 1], s.max(axis=1))
    kept = s >= kth[:, None]
    want = (np.exp((s - s.max(axis=1, keepdims=True)).astype(np.float64)) * kept).sum(axis=1)
    # the stored sum carries the rounding of max*log2(e) as a common factor (it cancels in every probability, which is
    # formed with the same constant): allow 2^-24 * |max| * log2(e) on top of the 2e-6 of the exp itself
    tol = 2e-6 + 1.0e-7 * 1.45 * np.abs(s.max(axis=1))
    assert np.all(np.abs(stats[:, 2] - want) / want < tol)


def _random_params(rng):
    family = rng.choice(["llamagen", "anole", "lumina_mgpt"])
    ncols = int(rng.choice([1024, 2048, 4096]))
    p = dict(family=family, ncols=ncols, depth=5 if family == "lumina_mgpt" else int(rng.choice([3, 4])),
             total_tokens=int(rng.choice([12, 26, 59, 90])), tree=str(rng.choice(["eagle2", "random"])),
             top_k=int(rng.choice([0, 50, 300, ncols // 4])), temperature=float(rng.choice([1.0, 0.7, 1.4])),
             lantern=bool(rng.random() < 0.8), lantern_k=int(rng.choice([5, 64, 1000])),
             lantern_delta=float(rng.choice([0.05, 0.1, 0.4, 1.5, 5.0])), cfg_scale=float(rng.choice([1.5, 3.0, 7.5])),
             boost=float(rng.choice([8.0, 10.0, 12.0])))
    if family == "lumina_mgpt":
        p["temperature"] = 1.0                      # the reference applies no HF warper inside Lumina's walk
        p["top_k"] = max(p["top_k"], 50)
        p["newline_depth"] = int(rng.choice([-1, -1, 1, 2]))
    elif rng.random() < 0.25:
        p["top_p"] = float(rng.choice([0.9, 0.6]))
    if rng.random() < 0.3:
        p["static_tree"] = str(rng.choice(["mc_sim_7b_63", "medusa_2_7b_63", "reverse_balanced_25", "chain"]))
        p["lantern_delta"] = float(rng.choice([5.0, 10.0, 0.1]))
        p["lantern_k"] = int(rng.choice([5, 10, 100]))
    return p


@pytest.mark.parametrize("chunk", range(6))
def test_random_configurations_match_oracle(chunk):
    """Differential fuzz: random family / window / warp knobs / relaxation / tree shape / CFG scale, three prompts per
    configuration, every schedule the configuration is eligible for; each prompt must equal its own oracle."""
    rng = np.random.default_rng(9000 + chunk)
    done = 0
    while done < 7:
        p = _random_params(rng)
        built, orcs, seed, tries = [], [], int(rng.integers(1, 10**6)), 0
        while len(built) < 3 and tries < 12:
            tries += 1
            b = C.build(dict(p, seed=seed + tries))
            o = C.oracle_step(b)
            if o.margin >= MARGIN:
                built.append(b)
                orcs.append(o)
        if len(built) < 3:
            continue
        done += 1
        lazy_ok = p["ncols"] in (2048, 4096) and not (1e-8 <= p.get("top_p", 1.0) < 1.0)
        for phases in (3, 8) + ((6,) if lazy_ok else ()):
            res = R.run_cases(built, phases=phases)
            for i, o in enumerate(orcs):
                try:
                    R.compare(res, i, o)
                except AssertionError as e:
                    raise AssertionError(f"params={p} seed={built[i].params['seed']} phases={phases}: {e}") from e


@pytest.mark.parametrize("family,kw", [("llamagen", dict(ncols=2048, top_k=300, lantern_k=100, boost=10.0)),
                                       ("lumina_mgpt", dict(ncols=4096, top_k=500, lantern_k=100, depth=5, boost=10.0)),
                                       ("anole", dict(ncols=2048, top_k=400, lantern_k=64, boost=9.0, tree="random"))])
def test_duplicate_sibling_tokens(family, kw):
    """The reference lists the children of a node by TOKEN (`candidates_set`, ea_model_llamagen.py:728-739): siblings
    that carry the same token are one candidate, and after its acceptance the rows of both subtrees stay in play.  The
    walk's dedup is a static table (longest token prefix a row shares with an earlier row); synthetic trees never
    repeat a token among siblings, so this case builds them on purpose.  Both schedules, every prompt against its oracle."""
    built, orcs, seed = [], [], 31000
    while len(built) < 8:
        b = C.build(dict(family=family, seed=seed, total_tokens=40, dup_siblings=6, **kw))
        seed += 1
        toks, par = b.tree.tokens, b.tree.parent
        assert any(toks[i] == toks[j] for i in range(1, b.tree.T) for j in range(1, i) if par[i] == par[j])
        o = C.oracle_step(b)
        if o.margin >= MARGIN:
            built.append(b)
            orcs.append(o)
    def through_duplicate(b, o):   # the accepted path passes a token that a sibling carries too
        toks, par = b.tree.tokens, b.tree.parent
        return any(j != int(n) and par[j] == par[int(n)] and toks[j] == toks[int(n)]
                   for n in o.select_indices[1:] for j in range(1, b.tree.T))
    assert sum(through_duplicate(b, o) for b, o in zip(built, orcs)) >= 2
    for phases in (3, 6):
        res = R.run_cases(built, phases=phases)
        for i, o in enumerate(orcs):
            R.compare(res, i, o)
