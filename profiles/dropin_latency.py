#!/usr/bin/env python
"""Batch-1 latency of the drop-in call a reference user makes (evaluate_posterior on one prompt, one step), the way the
reference's generate() loop calls it, against the reference's own op sequence on the same GPU (oracle/torch_path.py).

Measured per family at the reference's default tree (59 nodes): host wall time per call with a synchronize on both
sides (the call returns host values: best_candidate, accept_length), median of N calls.

Round 2: measured at realistic depth - for every family one case per accept length 0 .. 5 is searched (tree seed, RNG
seed) and timed; "typical" is the mean over the cases with accept length 2 .. 5 (the bench workload's mean is ~3).

usage: python profiles/dropin_latency.py [--out profiles/r2/dropin_latency_r2.json]
"""
import argparse, json, os, random, statistics, sys, time
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lantern_b200 import posterior as PO, synth, verify  # noqa: E402
from oracle import lantern_oracle as O, torch_path as TP  # noqa: E402  (baseline leg only)


class _M(PO.VerifyMixin):
    pass


class _L(PO.LuminaVerifyMixin):
    pass


def med(f, n=50, warm=5):
    for _ in range(warm):
        f()
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        f()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return statistics.median(ts) * 1e6


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="profiles/r2/dropin_latency_r2.json")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    out = {}
    for name, depth in (("llamagen", 4), ("anole", 4), ("lumina_mgpt", 5)):
        fam = verify.FAMILIES[name]
        ofam = {"llamagen": O.LLAMAGEN, "anole": O.ANOLE, "lumina_mgpt": O.LUMINA}[name]
        k = 1000
        table = synth.neighbor_table(0, fam.ncols, k + 1)
        proc = PO.prepare_logits_processor(temperature=1.0, top_p=1.0, top_k=2000)
        warp = O.Warp(1.0, 1.0, 2000)
        cases = {}
        for tseed in range(11, 60):
            if len(cases) >= 6:
                break
            tree = synth.eagle2_tree(tseed, 59, depth)
            synth.assign_tokens(tseed, tree, fam.col0, fam.col0 + fam.ncols)
            cond, uncond = synth.tree_logits(tseed, tree, fam.vocab, cfg=True, boost=13.0)
            tl = torch.from_numpy(np.stack([cond, uncond])).to(dev)                  # [2, T, V] like the target's output
            ri = torch.from_numpy(tree.retrieve_indices).to(dev)
            toks = torch.from_numpy(tree.tokens).to(dev)
            padded = torch.cat([toks, torch.full((1,), -1, device=dev, dtype=toks.dtype)])
            candidates = padded[ri]
            if name == "lumina_mgpt":
                m = _L()
                m.nearest_latents = table
                m.image_token_offset = 4
                m.eagle_version = 2
                handle = PO.TreeLogits(tl[:1], tl[1:2], 3.0, ri, None, 2000)
                call = lambda: m.evaluate_posterior(handle, candidates, lantern=True, lantern_k=k, lantern_delta=0.1)
            else:
                m = _M()
                m.nearest_latents = table
                m.lantern_family = name
                handle = PO.TreeLogits(tl[:1], tl[1:2], 3.0, ri)
                call = lambda: m.evaluate_posterior(handle, candidates, proc, lantern=True, lantern_k=k, lantern_delta=0.1)
            for rseed in range(5, 9):
                def ours(call=call, rseed=rseed):
                    random.seed(rseed)
                    return call()

                def ref(tl=tl, toks=toks, ri=ri, rseed=rseed):
                    random.seed(rseed)
                    u = [random.random() for _ in range(60)]
                    return TP.verify_step(tl[0], tl[1], 3.0, toks, ri, u, ofam, warp, True, k, 0.1, table)
                b, a, sp = ours()
                if a in cases:
                    continue
                r = ref()
                cases[a] = {"dropin_evaluate_posterior_us": round(med(ours), 1),
                            "reference_op_sequence_us": round(med(ref, n=6, warm=1), 1),
                            "same_decision": bool(int(b) == r[0] and a == r[1]), "tree_seed": tseed, "rng_seed": rseed}
        typ = [v for a, v in cases.items() if 2 <= a <= 5]
        out[name] = {"by_accept_length": {str(a): cases[a] for a in sorted(cases)},
                     "typical_accept_2_to_5": {
                         "dropin_evaluate_posterior_us": round(sum(v["dropin_evaluate_posterior_us"] for v in typ) / max(1, len(typ)), 1),
                         "reference_op_sequence_us": round(sum(v["reference_op_sequence_us"] for v in typ) / max(1, len(typ)), 1),
                         "cases": len(typ)}}
        print(name, json.dumps(out[name]), flush=True)
    json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
