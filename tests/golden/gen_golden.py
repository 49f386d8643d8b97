#!/usr/bin/env python
"""Generate golden vectors by running the LIVE reference (/root/reference) on CPU.

Run in the build container only (the reference checkout does not travel to the GPU box):

    python tests/golden/gen_golden.py

Writes ``tests/golden/posterior_cases.json`` (case parameters + reference outputs; inputs are
regenerated from the seeds by ``tests/casegen.py``) and ``tests/golden/tree_buffers.json``.

Import recipe: bypass ``models/__init__.py``, stub ``ftfy``/``bs4``,
alias the removed ``LogitsWarper``.  The reference methods are called unbound with a
``SimpleNamespace`` for ``self``; ``random.random`` is patched to replay supplied uniforms.
Cases whose smallest decision margin is below ``MARGIN`` are dropped (their outcome
legitimately depends on the last ulp of ``exp``) and counted in the metadata.
"""
from __future__ import annotations

import importlib
import json
import os
import random
import sys
import types
from types import SimpleNamespace as NS

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import casegen as C  # noqa: E402
from oracle import lantern_oracle as O  # noqa: E402
from lantern_b200 import choices as CH  # noqa: E402

REF = os.environ.get("LANTERN_REFERENCE", "/root/reference")
MARGIN = 1e-5


def import_reference():
    m = types.ModuleType("models")
    m.__path__ = [os.path.join(REF, "models")]
    sys.modules["models"] = m
    for n in ("ftfy", "bs4"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["bs4"].BeautifulSoup = object
    from transformers.generation import logits_process as lp
    if not hasattr(lp, "LogitsWarper"):
        lp.LogitsWarper = lp.LogitsProcessor
    import warnings
    warnings.simplefilter("ignore")
    mods = NS()
    mods.llamagen = importlib.import_module("models.ea_model_llamagen")
    mods.anole = importlib.import_module("models.ea_model_anole")
    mods.lumina = importlib.import_module("models.ea_model_lumina_mgpt")
    mods.utils = importlib.import_module("models.drafters.utils")
    return mods


class Replay:
    """Context manager: random.random() pops from a list and counts the draws."""

    def __init__(self, values):
        self.values = [float(v) for v in values]
        self.n = 0

    def __enter__(self):
        self._orig = random.random
        random.random = self._next
        return self

    def _next(self):
        v = self.values[self.n]
        self.n += 1
        return v

    def __exit__(self, *a):
        random.random = self._orig


def reference_tree_logits(R, b: C.Built) -> torch.Tensor:
    """What the reference's tree_decoding hands to evaluate_posterior (gathered [L,D,V])."""
    p = b.params
    fam = b.fam
    cond = torch.from_numpy(b.cond)
    if b.uncond is not None:
        comb = torch.stack([cond, torch.from_numpy(b.uncond)])[:, None]       # [2,1,T,V]
        if fam.lumina:
            # ea_model_lumina_mgpt.py:597
            tl = comb[1] + p["cfg_scale"] * (comb[0] - comb[1])
        else:
            tl = R.llamagen.cfg_logit_process(comb, p["cfg_scale"])[0]        # ea_model_llamagen.py:26-29
    else:
        tl = cond[None]
    ri = torch.from_numpy(b.tree.retrieve_indices)
    if fam.lumina:
        V = fam.vocab
        vocab = torch.arange(V)
        suppress = ~((vocab >= fam.col0) & (vocab < fam.col1))
        mm = NS(suppress_token_mask=suppress, image_next_line_token_id=O.LUMINA_NEWLINE_TOKEN,
                image_end_token_id=O.LUMINA_EOI_TOKEN)
        # position ids chosen so the row classes equal b.row_kinds: feed them through the real processor
        isi = 10
        w = 48
        # n+1 == 0 mod 49 for newline rows; pick n = 48 (newline) or n = 5 (image)
        n = np.where(b.row_kinds == O.ROW_NEWLINE, 48, 5)
        n = np.where(b.row_kinds == O.ROW_EOI, (w + 1) * 48, n)
        pos_plus1 = torch.from_numpy(n + (isi + 1 + 2))
        assert (O.lumina_row_kinds(pos_plus1.numpy(), isi) == b.row_kinds).all()
        scores = R.lumina.MultiModalLogitsProcessor.__call__(mm, tl[0], image_start_token_id_index=isi,
                                                             position_ids=pos_plus1)
        if p["top_k"] > 0:
            scores = R.lumina.InterleavedTopKLogitsWarper(image_top_k=p["top_k"])(scores)
        return scores[ri]
    if fam.mask_non_image:
        non_image = torch.tensor([i for i in range(fam.vocab) if not (fam.col0 <= i < fam.col1)])
        tl[:, :, non_image] = torch.finfo(tl.dtype).min                        # ea_model_anole.py:931
    return tl[0, ri]


def run_reference(R, b: C.Built):
    p = b.params
    fam = b.fam
    logits = reference_tree_logits(R, b)
    cand = torch.from_numpy(b.candidates)
    k = min(int(p["lantern_k"]), fam.ncols - 1)
    proc = R.utils.prepare_logits_processor(temperature=p["temperature"], top_p=p["top_p"], top_k=p["top_k"])
    tbl = None
    if b.table is not None:
        tbl = b.table.astype(np.uint16) if p["family"] == "llamagen" else b.table.astype(np.int64)
    with Replay(b.uniforms.astype(np.float64)) as rp:
        if p["family"] == "vanilla":
            out = R.utils.evaluate_posterior(logits, cand, proc)
        elif fam.lumina:
            me = NS(nearest_latents=tbl, image_token_offset=fam.offset,
                    image_tokens=torch.arange(fam.col0, fam.col1),
                    image_syntax_tokens=torch.tensor(fam.syntax_tokens),
                    eagle_version=1 if b.static is not None else 2)
            kw = {}
            if b.static is not None:
                st = b.static
                kw = dict(cart_candidates_prob=torch.from_numpy(st.cart_prob),
                          original_prob=[torch.from_numpy(o) for o in st.op], p_indices=st.p_indices,
                          tree_candidates=torch.from_numpy(st.tree_candidates)[None],
                          b_indices=[[torch.tensor(x, dtype=torch.long) if len(x) else [] for x in row]
                                     for row in st.b_indices])
            out = R.lumina.EaLumina_mGPT.evaluate_posterior(me, logits, cand, do_sample=True, lantern=p["lantern"],
                                                            lantern_k=k, lantern_delta=p["lantern_delta"], **kw)
        else:
            mod = R.llamagen if p["family"] == "llamagen" else R.anole
            me = NS(nearest_latents=tbl, image_token_offset=fam.offset)
            if b.static is not None:
                st = b.static
                out = mod.EaModel.evaluate_posterior_v1(
                    me, logits, cand, proc, torch.from_numpy(st.cart_prob), [torch.from_numpy(o) for o in st.op],
                    st.p_indices, torch.from_numpy(st.tree_candidates)[None],
                    [[torch.tensor(x, dtype=torch.long) if len(x) else [] for x in row] for row in st.b_indices],
                    lantern=p["lantern"], lantern_k=k, lantern_delta=p["lantern_delta"])
            else:
                out = mod.EaModel.evaluate_posterior(me, logits, cand, proc, lantern=p["lantern"], lantern_k=k,
                                                     lantern_delta=p["lantern_delta"])
    best, alen, sp = out
    return int(best), int(alen), sp.detach().numpy().astype(np.float32), rp.n


def case_list():
    cases = []
    s = 0

    def add(**kw):
        nonlocal s
        kw.setdefault("seed", 1000 + s)
        s += 1
        cases.append(C.default_params(**kw))

    # Config 1 (BASELINE.json configs[0]): LlamaGen 16384x8, k=1000, delta=0.1, EAGLE-2 tree, batch 1
    for i in range(12):
        add(family="llamagen")
    for i in range(4):
        add(family="llamagen", lantern=False)
    # smaller vocabularies, knob sweep
    for k, d in [(5, 5.0), (10, 10.0), (10, 20.0), (100, 0.05), (1000, 0.3), (300, 1.5)]:
        for tk in (0, 200):
            for temp in (1.0, 0.8, 1.3):
                add(family="llamagen", ncols=2048, lantern_k=k, lantern_delta=d, top_k=tk, temperature=temp,
                    boost=11.0)
    for tp in (0.9, 0.5):
        for tk in (0, 300):
            add(family="llamagen", ncols=2048, top_p=tp, top_k=tk, lantern_k=50, boost=11.0)
    for T in (2, 5, 17, 120):
        add(family="llamagen", ncols=1024, tree="random", total_tokens=T, top_k=100, lantern_k=64, boost=9.0)
    add(family="llamagen", ncols=1024, cfg=False, top_k=100, lantern_k=64, boost=9.0)
    # vanilla EAGLE (drafters/utils.py:333-410): no CFG, no lantern
    for i in range(4):
        add(family="vanilla", ncols=4096, cfg=False, lantern=False, top_k=(0 if i % 2 else 50), boost=10.0)
    # static trees (LANTERN++), LlamaGen
    for name in CH.NAMES:
        for k, d in [(5, 5.0), (10, 10.0), (10, 20.0), (1000, 0.1)]:
            add(family="llamagen", ncols=4096, static_tree=name, lantern_k=k, lantern_delta=d, top_k=500,
                boost=8.5 if d > 1 else 7.0)
    for i in range(4):
        add(family="llamagen", ncols=4096, static_tree="mc_sim_7b_63", lantern=False, top_k=500, boost=8.5)
    # Anole: offset 4, finfo.min mask
    for i in range(6):
        add(family="anole", ncols=2048, top_k=400, lantern_k=(1000 if i < 3 else 10),
            lantern_delta=(0.1 if i < 3 else 5.0), boost=11.0)
    for i in range(3):
        add(family="anole", ncols=2048, static_tree="mc_sim_7b_63_balanced", top_k=400, lantern_k=10,
            lantern_delta=10.0, boost=8.0)
    add(family="anole")                                    # full 65536 vocab, 8192 image tokens
    # Lumina-mGPT: dynamic (eagle_version 2) and static (eagle_version 1), newline rows
    for i in range(6):
        add(family="lumina_mgpt", ncols=2048, depth=5, top_k=400, newline_depth=(i % 3) if i < 4 else -1,
            lantern_k=(1000 if i % 2 else 100), boost=11.0)
    for i in range(4):
        add(family="lumina_mgpt", ncols=2048, depth=5, top_k=400, static_tree="mc_sim_7b_63", lantern_k=10,
            lantern_delta=(5.0 if i % 2 else 20.0), boost=8.0)
    add(family="lumina_mgpt", depth=5)                      # full size, top_k 2000
    add(family="lumina_mgpt", depth=5, lantern=False)
    # ---- round 2 (appended: the seeds of the cases above do not move) ----
    # walks that END on a one-hot newline row: every child of a newline row is a junk token (all rejected, residual
    # tail), or the newline rows are the deepest level (leaf rows: fresh tail); the gathered [L, D, V] form carries no
    # row classes, so these pin the kernel's own detection of pre-masked one-hot rows
    for nd in (0, 1, 2):
        add(family="lumina_mgpt", ncols=2048, depth=5, top_k=400, newline_depth=nd, newline_junk=True, lantern_k=100,
            boost=11.0)
    for seed, nd in ((2001, 2), (2010, 1), (2018, 1), (2003, 2)):    # seeds whose walk stops on a childless newline node
        add(family="lumina_mgpt", ncols=2048, depth=5, top_k=400, newline_depth=nd, lantern_k=100, boost=13.0, seed=seed)
    # generate(top_k=...) other than the default 2000 (ea_model_lumina_mgpt.py:822-823)
    add(family="lumina_mgpt", depth=5, top_k=500)
    add(family="lumina_mgpt", depth=5, top_k=4000, lantern_k=300)
    # BASELINE configs[2]: LANTERN++ static trees on Lumina-mGPT at full size, k in {5, 10} x lambda in {5, 10, 20}
    for name in ("mc_sim_7b_63", "naive_extend_57", "medusa_2_7b_63", "chain"):
        for k in (5, 10):
            for d in (5.0, 10.0, 20.0):
                add(family="lumina_mgpt", depth=5, static_tree=name, lantern_k=k, lantern_delta=d, boost=8.0)
    # ... and synthetic static trees of 10 / 24 / 40 / 80 nodes (lantern_b200.choices.synth_tree)
    for i, name in enumerate(CH.SYNTH_NAMES):
        for k, d in ((5, 10.0), (10, 5.0), (10, 20.0)):
            add(family="lumina_mgpt", depth=5, static_tree=name, lantern_k=k, lantern_delta=d, boost=8.0)
        add(family="llamagen", ncols=4096, static_tree=name, lantern_k=10, lantern_delta=10.0, top_k=500, boost=8.5)
    return cases


def tree_buffer_fixtures(R):
    out = {}
    for name in CH.NAMES + CH.SYNTH_NAMES:
        tb = R.utils.generate_tree_buffers(CH.tree(name), device="cpu")
        out[name] = {
            "tree_attn_mask": tb["tree_attn_mask"][0, 0].to(torch.int64).tolist(),
            "tree_indices": tb["tree_indices"].tolist(),
            "tree_position_ids": tb["tree_position_ids"].tolist(),
            "retrieve_indices": tb["retrieve_indices"].tolist(),
            "p_indices": tb["p_indices"],
            "b_indices": [[(x.tolist() if isinstance(x, torch.Tensor) else list(x)) for x in row]
                          for row in tb["b_indices"]],
        }
    return out


def dynamic_tree_fixtures():
    """Execute the reference's own dynamic-tree post-processing (the tail of Model.topK_genrate,
    models/drafters/cnets_llamagen.py:831-912) on synthetic expansion outputs.  The code is taken from the mounted
    reference at generation time (inspect.getsource) and never written to this repository."""
    import inspect
    import textwrap
    from lantern_b200 import synth
    cn = importlib.import_module("models.drafters.cnets_llamagen")
    src = inspect.getsource(cn.Model.topK_genrate)
    start = src.index("scores_list = torch.cat(scores_list, dim=0).view(-1)")
    end = src.index("return draft_tokens, retrieve_indices, tree_mask, tree_position_ids")
    body = textwrap.dedent(" " * 8 + src[start:end])
    out = []
    for seed, depth, total in [(1, 4, 58), (2, 4, 58), (3, 5, 58), (4, 5, 58), (5, 4, 25), (6, 3, 99), (7, 6, 120),
                               (8, 4, 9), (9, 5, 58), (10, 4, 58)]:
        ex = synth.eagle2_expansion(seed, depth=depth, top_k=10)
        for sort_rows in (True, False):
            k = ex.top_k
            sizes = [k] + [k * k] * depth
            sl = [t.view(-1, k) for t in torch.split(torch.from_numpy(ex.scores), sizes)]
            tl = [t.view(-1, k) for t in torch.split(torch.from_numpy(ex.tokens), sizes)]
            pl = list(torch.split(torch.from_numpy(ex.parents), [1] + [k] * depth))
            ns = {"torch": torch, "scores_list": sl, "ss_token": tl, "parents_list": pl, "total_tokens": total,
                  "top_k": k, "sample_token": torch.tensor([ex.sample_token]),
                  "logits_processor": (object() if sort_rows else None), "hidden_states": torch.zeros(1)}
            exec(body, ns)
            out.append({"seed": seed, "depth": depth, "total_tokens": total, "sort_rows": sort_rows,
                        "draft_tokens": ns["draft_tokens"][0].tolist(),
                        "retrieve_indices": ns["retrieve_indices"].tolist(),
                        "tree_position_ids": ns["tree_position_ids"].tolist(),
                        "tree_mask": ns["tree_mask"][0, 0].to(torch.int64).tolist()})
    return out


GREEDY_CASES = [dict(family="anole", ncols=nc, top_k=0, temperature=0.0, lantern=lan, lantern_k=k, lantern_delta=d,
                     seed=7000 + 13 * i + j, boost=bst, tree=tree, total_tokens=tt)
                for i, (nc, k, d, bst, tree, tt) in enumerate([
                    (1024, 100, 0.1, 9.0, "eagle2", 59), (1024, 100, 0.3, 9.0, "eagle2", 59),
                    (2048, 1000, 0.1, 9.5, "eagle2", 59), (2048, 10, 5.0, 9.5, "eagle2", 59),
                    (1024, 64, 2.0, 9.0, "random", 24), (4096, 300, 0.2, 10.0, "eagle2", 40)])
                for j in range(4) for lan in (True, False)] + [
    dict(family="anole", top_k=0, temperature=0.0, lantern=True, lantern_k=1000, lantern_delta=0.1, seed=7900, boost=11.0)]


def greedy_fixtures(R):
    """Greedy branches (logits_processor is None) of the live reference: ea_model_anole.py:789-902 with and without
    the LANTERN relaxation, on Anole-shaped inputs (int64 table, offset 4, finfo.min mask)."""
    kept, dropped = [], 0
    for p in GREEDY_CASES:
        b = C.build(C.default_params(**p))
        logits = reference_tree_logits(R, b)
        cand = torch.from_numpy(b.candidates)
        k = min(int(p["lantern_k"]), b.fam.ncols - 1)
        tbl = b.table.astype(np.int64) if b.table is not None else None
        me = NS(nearest_latents=tbl, image_token_offset=b.fam.offset)
        best, alen, row = R.anole.EaModel.evaluate_posterior(me, logits.clone(), cand, None, lantern=p["lantern"],
                                                             lantern_k=k, lantern_delta=p["lantern_delta"])
        if p["lantern"]:
            ob, oa, orow, margin = O.evaluate_posterior_greedy_lantern(logits.numpy(), b.candidates, b.fam, b.table, k,
                                                                       p["lantern_delta"])
        else:
            ob, oa, orow = O.evaluate_posterior_greedy(logits.numpy(), b.candidates)
            margin = 1.0
        if margin < MARGIN:
            dropped += 1
            continue
        assert (int(best), int(alen)) == (ob, oa) and np.array_equal(row.numpy(), orow), ("ORACLE MISMATCH", p)
        kept.append({"params": C.default_params(**p), "best_candidate": int(best), "accept_length": int(alen),
                     "token": int(row.argmax()), "row_sum": float(row.double().clamp(min=-1e30).sum()),
                     "input_checksum": float(b.cond.astype(np.float64).sum())})
    return {"meta": {"generator": "tests/golden/gen_golden.py greedy_fixtures", "dropped_fragile": dropped,
                     "n_cases": len(kept)}, "cases": kept}


SAMPLE_CASES = [dict(seed=8000 + i, V=V, rows=rows, k=k, temperature=t, top_k=tk, top_p=tp)
                for i, (V, rows, k, t, tk, tp) in enumerate([
                    (16384, 4, 10, 1.0, 2000, 1.0), (16384, 1, 10, 1.0, 0, 1.0), (4096, 11, 10, 0.8, 500, 1.0),
                    (4096, 3, 10, 1.3, 0, 0.9), (2048, 5, 4, 1.0, 12, 1.0), (2048, 2, 10, 1.0, 10, 1.0),
                    (8192, 4, 10, 1.0, 2000, 1.0), (1024, 6, 8, 1.0, 100, 0.98)])]


def sample_fixtures(R):
    """Model.sample (models/drafters/cnets_llamagen.py:924-940, same body in cnets_anole / cnets_lumina_mgpt.py:936-955)
    of the live reference with torch.multinomial replaced by the build's documented draw (oracle draft_sample: exponential
    race on the Philox stream), so that everything AFTER the draw - gather, exclusive cumsum, the conditional
    probability p_i / (1 - sum_{j<i} p_j), the inf / nan -> -1 patch-up, the clamp - and the full distribution `op` are
    the reference's own arithmetic.  top_k == k exhausts the kept mass, so 1 - cumsum reaches ~0 on the last draws (the inf / nan / clamp branch)."""
    from lantern_b200 import synth
    cn = importlib.import_module("models.drafters.cnets_llamagen")
    out = []
    for p in SAMPLE_CASES:
        V, rows, k = p["V"], p["rows"], p["k"]
        logits = (synth.gauss(p["seed"], (rows, V), stream=9) * np.float32(2.5)).astype(np.float32)
        proc = R.utils.prepare_logits_processor(temperature=p["temperature"], top_p=p["top_p"], top_k=p["top_k"])
        warp = O.Warp(p["temperature"], p["top_p"], p["top_k"])
        picks = [O.draft_sample(logits[r], warp, k, seed=p["seed"], step=3, row=r) for r in range(rows)]
        idx = torch.from_numpy(np.stack([pk[0] for pk in picks]))
        orig = torch.multinomial
        torch.multinomial = lambda probs, n, replacement=False: idx        # the draw itself is the build's definition
        try:
            si, sp, probs = cn.Model.sample(None, torch.from_numpy(logits), proc, k=k)
        finally:
            torch.multinomial = orig
        sp, probs = sp.numpy(), probs.numpy()
        o_cp = np.stack([pk[1] for pk in picks])
        o_pr = np.stack([pk[2] for pk in picks])
        # the oracle restatement must agree with the reference before the fixture is worth anything
        assert np.array_equal(si.numpy(), idx.numpy())
        err_cp = float(np.max(np.abs(o_cp - sp) / np.maximum(np.abs(sp), 1e-30) * (sp > 0)))
        err_pr = float(np.max(np.abs(o_pr - probs) / np.maximum(probs, 1e-30) * (probs > 0)))
        assert err_cp <= 1e-5 and err_pr <= 1e-5 and np.array_equal(o_pr > 0, probs > 0), ("ORACLE MISMATCH", p, err_cp, err_pr)
        probe = (synth.hash_u64(4242, 32, 3) % np.uint64(V)).astype(np.int64)
        out.append({"params": p, "indices": idx.tolist(), "cond_probs": [[float(v) for v in r] for r in sp],
                    "probe_cols": probe.tolist(), "probe_probs": [[float(v) for v in probs[r, probe]] for r in range(rows)],
                    "picked_probs": [[float(v) for v in probs[r, idx[r].numpy()]] for r in range(rows)],
                    "nnz": [int((probs[r] > 0).sum()) for r in range(rows)],
                    "oracle_rel_err": {"cond_probs": err_cp, "probs": err_pr}})
    return {"meta": {"generator": "tests/golden/gen_golden.py sample_fixtures",
                     "reference": "Model.sample, models/drafters/cnets_llamagen.py:924-940", "n_cases": len(out)},
            "cases": out}


def signature_fixture(R):
    """Parameter lists of the reference's call surface for tests/test_host_logic.py."""
    import inspect

    def sig(f):
        return [[n, (None if p.default is inspect._empty else repr(p.default))]
                for n, p in inspect.signature(f).parameters.items()]
    out = {f"utils.{n}": sig(getattr(R.utils, n)) for n in
           ("tree_decoding", "evaluate_posterior", "update_inference_inputs", "prepare_logits_processor",
            "generate_tree_buffers")}
    for fam, mod in (("llamagen", R.llamagen), ("anole", R.anole)):
        for m in ("tree_decoding", "evaluate_posterior", "evaluate_posterior_v1", "update_inference_inputs"):
            out[f"{fam}.EaModel.{m}"] = sig(getattr(mod.EaModel, m))
    for m in ("tree_decoding", "evaluate_posterior", "update_inference_inputs"):
        out[f"lumina.EaLumina_mGPT.{m}"] = sig(getattr(R.lumina.EaLumina_mGPT, m))
    return out


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    R = import_reference()
    with open(os.path.join(HERE, "signatures.json"), "w") as f:
        json.dump(signature_fixture(R), f, indent=1)
    with open(os.path.join(HERE, "greedy_cases.json"), "w") as f:
        json.dump(greedy_fixtures(R), f)
    with open(os.path.join(HERE, "sample_cases.json"), "w") as f:
        json.dump(sample_fixtures(R), f)
    if "--greedy-only" in sys.argv:
        return
    with open(os.path.join(HERE, "dynamic_trees.json"), "w") as f:
        json.dump(dynamic_tree_fixtures(), f)
    kept, fragile, dropped, mismatched, worst_sp = [], [], 0, 0, 0.0
    for p in case_list():
        b = C.build(p)
        best, alen, sp, ndraw = run_reference(R, b)
        orc = C.oracle_step(b)
        if orc.margin < MARGIN:
            # A decision sits within MARGIN (relative) of its threshold: the outcome depends on the last ulp of exp, so
            # the case cannot gate parity - but it is kept, with the reference's outcome, so that the tests can REPORT
            # how many such decisions the oracle and the CUDA path take the other way (tests/test_fragile_band*.py).
            dropped += 1
            fragile.append({"params": p, "best_candidate": best, "accept_length": alen, "n_uniforms": ndraw,
                            "oracle_margin": float(orc.margin),
                            "oracle_agrees": bool(best == orc.best_candidate and alen == orc.accept_length
                                                  and ndraw == orc.n_uniforms - 1)})
            continue
        idx, val, nnz, tot = C.sample_p_probe(sp)
        ok = (best == orc.best_candidate and alen == orc.accept_length and ndraw == orc.n_uniforms - 1)
        if not ok:
            mismatched += 1
            print("ORACLE MISMATCH", p, (best, alen, ndraw), (orc.best_candidate, orc.accept_length,
                                                               orc.n_uniforms - 1))
        sp_err = float(np.max(np.abs(orc.sample_p[idx] - val) / np.maximum(np.abs(val), 1e-30)
                              * (val > 0))) if ok else -1.0
        worst_sp = max(worst_sp, sp_err)
        if sp_err > 1e-5 or int((orc.sample_p > 0).sum()) != nnz:
            mismatched += 1
            print("ORACLE sample_p MISMATCH", p, sp_err, int((orc.sample_p > 0).sum()), nnz)
        kept.append({"params": p, "best_candidate": best, "accept_length": alen, "n_uniforms": ndraw,
                     "sp_idx": idx.tolist(), "sp_val": [float(v) for v in val], "sp_nnz": nnz, "sp_sum": tot,
                     "input_checksum": float(b.cond.astype(np.float64).sum()),
                     "oracle_margin": float(orc.margin)})
    meta = {"generator": "tests/golden/gen_golden.py", "reference": "jadohu/LANTERN @ /root/reference",
            "torch": torch.__version__, "margin_filter": MARGIN, "dropped_fragile": dropped,
            "oracle_mismatches_at_generation": mismatched, "n_cases": len(kept),
            "fragile_oracle_disagreements": sum(not c["oracle_agrees"] for c in fragile),
            "oracle_sample_p_max_rel_err": worst_sp}
    with open(os.path.join(HERE, "posterior_cases.json"), "w") as f:
        json.dump({"meta": meta, "cases": kept, "fragile_cases": fragile}, f)
    with open(os.path.join(HERE, "tree_buffers.json"), "w") as f:
        json.dump(tree_buffer_fixtures(R), f)
    print(json.dumps(meta, indent=1))
    hist = {}
    for c in kept:
        hist[c["accept_length"]] = hist.get(c["accept_length"], 0) + 1
    print("accept-length histogram:", dict(sorted(hist.items())))


if __name__ == "__main__":
    main()
