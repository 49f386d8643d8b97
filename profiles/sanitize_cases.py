#!/usr/bin/env python
"""Small end-to-end invocations of every kernel family, meant to be run under compute-sanitizer
(memcheck / racecheck on the kernels):

    compute-sanitizer --tool memcheck  python profiles/sanitize_cases.py
    compute-sanitizer --tool racecheck python profiles/sanitize_cases.py

Every result is still checked against the oracle, so the run also proves the sanitised path computes the same."""
import os, sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import casegen as C          # noqa: E402
import cuda_runner as R      # noqa: E402
from lantern_b200 import codebook, posterior as PO  # noqa: E402
from oracle import lantern_oracle as O  # noqa: E402


def accept(params, phases=3):
    built, orcs, seed = [], [], params.pop("seed", 300)
    while len(built) < 2:
        b = C.build(dict(params, seed=seed))
        seed += 1
        o = C.oracle_step(b)
        if o.margin >= 1e-5:
            built.append(b)
            orcs.append(o)
    res = R.run_cases(built, phases=phases)
    for i, o in enumerate(orcs):
        R.compare(res, i, o)


def stream_rows():
    from lantern_b200 import verify
    rng = np.random.default_rng(3)
    for ncols, dt, shape in ((2048, torch.float32, "gauss"), (4096, torch.bfloat16, "gauss"), (2048, torch.float32, "t2")):
        B, T, top_k = 14, 59, 300                               # 826 rows on 296 CTAs
        x = rng.standard_normal((B, T, ncols)) * 2.5 if shape == "gauss" else rng.standard_t(2, (B, T, ncols))
        cond = torch.from_numpy(x.astype(np.float32)).cuda().to(dt)
        uncond = torch.from_numpy((x + rng.standard_normal((B, T, ncols)) * 0.7).astype(np.float32)).cuda().to(dt)
        fam = verify.LLAMAGEN.resized(ncols)
        v = verify.Verifier(fam, temperature=1.0, top_k=top_k, cfg_scale=3.0, lantern=False, device=torch.device("cuda"))
        tokens = torch.zeros(B, T, dtype=torch.int32, device="cuda")
        retrieve = torch.zeros(B, 1, 1, dtype=torch.int32, device="cuda")
        v.step(cond, uncond, tokens, retrieve, uniforms=torch.rand(B, 2, device="cuda"), phases=1)
        torch.cuda.synchronize()
        stats = v._work[:B * T * 32].view(torch.float32).view(B * T, 8).cpu().numpy()
        s = O.cfg_mix(cond.float().cpu().numpy().reshape(-1, ncols), uncond.float().cpu().numpy().reshape(-1, ncols), 3.0)
        kth = np.partition(s, ncols - top_k, axis=1)[:, ncols - top_k]
        assert np.array_equal(stats[:, 0], kth) and np.array_equal(stats[:, 1], s.max(axis=1)), (ncols, dt, shape)


def one_call():
    import random
    from lantern_b200 import synth, verify

    class _M(PO.VerifyMixin):
        lantern_family = "llamagen"
        lantern_image_tokens = 2048
    fam = O.small_family(O.LLAMAGEN, 2048)
    tree = synth.eagle2_tree(5, 30, 4)
    synth.assign_tokens(5, tree, 0, 2048)
    cond, uncond = synth.tree_logits(5, tree, 2048, cfg=True, boost=11.0)
    m = _M()
    m.nearest_latents = synth.neighbor_table(0, 2048, 101)
    ri = torch.from_numpy(tree.retrieve_indices).cuda()
    toks = torch.from_numpy(tree.tokens).cuda()
    cand = torch.cat([toks, torch.full((1,), -1, device="cuda", dtype=toks.dtype)])[ri]
    handle = PO.TreeLogits(torch.from_numpy(cond).cuda()[None], torch.from_numpy(uncond).cuda()[None], 3.0, ri)
    proc = PO.prepare_logits_processor(temperature=1.0, top_p=1.0, top_k=300)
    random.seed(9)
    st = random.getstate()
    best, a, sp = m.evaluate_posterior(handle, cand, proc, lantern=True, lantern_k=100, lantern_delta=0.1)
    random.setstate(st)
    u = [random.random() for _ in range(tree.T)]
    o = O.verify_step(cond, uncond, 3.0, tree.tokens, tree.retrieve_indices, np.asarray(u + [0.5]), fam,
                      O.Warp(1.0, 1.0, 300), True, 100, 0.1, m.nearest_latents)
    assert o.margin < 1e-5 or (int(best), a) == (o.best_candidate, o.accept_length)


def main():
    torch.cuda.set_device(0)
    which = sys.argv[1:] or ["accept", "neighbors", "kv", "greedy", "drafter", "session"]
    if "accept" in which:
        accept(dict(family="llamagen", ncols=2048, top_k=300, lantern_k=100, boost=11.0))            # generic statistics kernel
        accept(dict(family="lumina_mgpt", ncols=2048, top_k=500, lantern_k=100, depth=5))            # fast (TMA) statistics kernel
        accept(dict(family="lumina_mgpt", ncols=4096, top_k=500, lantern_k=100, depth=5), phases=6)  # lazy walk
        accept(dict(family="anole", ncols=1024, top_k=0, top_p=0.9, lantern_k=50))                   # top-p kernel
        accept(dict(family="llamagen", ncols=2048, top_k=300, lantern_k=20, lantern_delta=5.0,
                    static_tree="mc_sim_7b_63", seed=500))                                           # static tree / LANTERN++
        # round 2: streaming statistics kernel with several rows per CTA (select warps two rows behind), bf16 rows,
        # non-Gaussian rows that miss the bracket (redo path), the 65536-column parity path, the one-call drop-in entry
        stream_rows()
        accept(dict(family="vanilla", ncols=65536, cfg=False, lantern=False, top_k=50, boost=13.0, total_tokens=8, seed=700))
        one_call()
        # round 2, lazy walk rewritten for latency (256 threads, children listed / tree staged inside the row's load
        # shadow, atomic-free bracket select): every row width it is instantiated for, a deep tree, a static tree,
        # rows that miss the bracket (heavy ties -> slow selectors), no top-k at all
        accept(dict(family="llamagen", ncols=2048, top_k=300, lantern_k=100, boost=11.0, total_tokens=59), phases=6)
        accept(dict(family="anole", ncols=8192, top_k=2000, lantern_k=1000, depth=5, total_tokens=59, boost=13.0, seed=900), phases=6)
        accept(dict(family="llamagen", ncols=16384, top_k=1000, lantern_k=1000, total_tokens=26, boost=13.0, seed=950), phases=6)
        accept(dict(family="llamagen", ncols=2048, top_k=300, lantern_k=20, lantern_delta=5.0,
                    static_tree="mc_sim_7b_63", seed=500), phases=6)
        accept(dict(family="llamagen", ncols=2048, top_k=0, lantern_k=100, boost=11.0, seed=980), phases=6)
        accept(dict(family="lumina_mgpt", ncols=4096, top_k=500, lantern_k=100, depth=5, newline_depth=2, seed=990), phases=6)
        print("accept ok", flush=True)
    if "neighbors" in which:
        for N, d, K in ((1024, 8, 65), (512, 256, 33), (300, 8, 299), (2304, 8, 1001)):
            rng = np.random.default_rng(N + d)
            E = rng.standard_normal((N, d)).astype(np.float32)
            got = codebook.build_neighbor_table(torch.from_numpy(E).cuda(), K).cpu().numpy()
            want = O.neighbor_table(E, K)
            assert np.array_equal(got, want), (N, d, K)
        print("neighbors ok", flush=True)
    if "kv" in which:
        slab = torch.arange(2 * 2 * 2 * 64 * 8, dtype=torch.float32, device="cuda").view(2, 2, 2, 64, 8).to(torch.bfloat16)
        ref = slab.clone()
        sel = torch.tensor([40, 43, 47], device="cuda")
        n = PO.kv_compact([slab], sel, 40)
        torch.cuda.synchronize()
        assert n == 43 and torch.equal(slab[..., 40:43, :], ref[..., [40, 43, 47], :])
        print("kv ok", flush=True)


def extra(which):
    from lantern_b200 import draft_sample, dyntree, synth, verify
    if "greedy" in which:
        for lantern in (False, True):
            b = C.build(dict(family="anole", ncols=1024, top_k=0, temperature=0.0, lantern=lantern, lantern_k=100,
                             lantern_delta=0.2, boost=9.0, seed=86000))
            best, a, row, _ = C.oracle_greedy(b)
            fam = R.family_spec(b)
            v = verify.Verifier(fam, cfg_scale=b.params["cfg_scale"], lantern=lantern, lantern_k=100, lantern_delta=0.2,
                                nbr_table=None if b.table is None else torch.from_numpy(b.table.astype(np.int32)).cuda())
            res = v.greedy(torch.from_numpy(b.cond)[None].cuda(), torch.from_numpy(b.uncond)[None].cuda(),
                           torch.from_numpy(b.tree.tokens.astype(np.int32))[None].cuda(),
                           torch.from_numpy(R.pad_retrieve([b.tree.retrieve_indices])).cuda())
            torch.cuda.synchronize()
            assert int(res.accept_length[0]) == a and int(res.best_candidate[0]) == best
        print("greedy ok", flush=True)
    if "drafter" in which:
        e = synth.eagle2_expansion(70, depth=5, top_k=10, lo=4, hi=8196)
        t = dyntree.build_dynamic_tree(torch.from_numpy(e.scores).cuda(), torch.from_numpy(e.tokens).cuda(),
                                       torch.from_numpy(e.parents).cuda(), torch.tensor([e.sample_token]).cuda(), 58, top_k=10)
        o = O.dynamic_tree(e.scores, e.tokens, e.parents, e.sample_token, 58, 10)
        assert np.array_equal(t.reference_outputs(0)[1].cpu().numpy(), o[4])
        logits = torch.randn(6, 4096, device="cuda") * 2.5
        proc = PO.prepare_logits_processor(temperature=1.0, top_p=1.0, top_k=300)
        idx, cond, probs = draft_sample.sample(logits, proc, 10, seed=3, step=4)
        torch.cuda.synchronize()
        oi, oc, op = O.draft_sample(logits[2].cpu().numpy(), O.Warp(1.0, 1.0, 300), 10, 3, 4, 2)
        assert np.array_equal(idx[2].cpu().numpy(), oi)
        print("drafter ok", flush=True)
    if "session" in which:
        from lantern_b200.session import HostSession
        b = C.build(dict(family="lumina_mgpt", ncols=4096, top_k=500, lantern_k=100, depth=5, seed=87000))
        o = C.oracle_step(b)
        fam = R.family_spec(b)
        v = verify.Verifier(fam, top_k=500, cfg_scale=b.params["cfg_scale"], lantern=True, lantern_k=100,
                            lantern_delta=b.params["lantern_delta"], nbr_table=torch.from_numpy(b.table.astype(np.int32)).cuda())
        ri = R.pad_retrieve([b.tree.retrieve_indices])
        tok = b.tree.tokens.astype(np.int32)[None]
        uni = b.uniforms.astype(np.float32)[None]
        with HostSession(v, 1, b.tree.T, ri.shape[1], ri.shape[2], n_uniforms=uni.shape[1]) as sess:
            for pin in (False, True):
                cnd, unc = torch.from_numpy(b.cond)[None], torch.from_numpy(b.uncond)[None]
                if pin:
                    cnd, unc = cnd.pin_memory(), unc.pin_memory()
                r = sess.step(cnd, unc, tok, ri, uniforms=uni)
                assert r.in_place == pin
                if o.margin >= 1e-5:
                    assert int(r.accept_length[0]) == o.accept_length and int(r.token[0]) == o.token
        print("session ok", flush=True)


if __name__ == "__main__":
    main()
    extra(sys.argv[1:] or ["greedy", "drafter", "session"])
