#!/usr/bin/env python
"""Device time of the greedy branches (lantern_accept_greedy) at Anole shapes (V = 65536, 8192 image tokens, T = 59),
plain and LANTERN-relaxed, CUDA events around 50 calls.  usage: python profiles/time_greedy.py [--out file.json]"""
import argparse, json, os, sys
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lantern_b200 import synth, verify  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="profiles/greedy_r1.json")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    fam = verify.FAMILIES["anole"]
    table = torch.from_numpy(synth.neighbor_table(0, fam.ncols, 1001)).to(dev)
    out = {}
    for B in (1, 64):
        trees = []
        for i in range(min(B, 8)):
            t = synth.eagle2_tree(40 + i, 59, 4)
            synth.assign_tokens(40 + i, t, fam.col0, fam.col0 + fam.ncols)
            trees.append(t)
        trees = [trees[i % len(trees)] for i in range(B)]
        g = torch.Generator(device=dev)
        g.manual_seed(B)
        cond = torch.empty(B, 59, fam.vocab, device=dev).normal_(0, 2.3, generator=g)
        uncond = cond + torch.empty_like(cond).normal_(0, 0.8, generator=g)
        tokens = torch.from_numpy(np.stack([t.tokens for t in trees]).astype(np.int32)).to(dev)
        L = max(t.retrieve_indices.shape[0] for t in trees)
        D = max(t.retrieve_indices.shape[1] for t in trees)
        ri = np.full((B, L, D), -1, dtype=np.int32)
        for i, t in enumerate(trees):
            r = t.retrieve_indices
            ri[i, :r.shape[0], :r.shape[1]] = r
        retrieve = torch.from_numpy(ri).to(dev)
        for lantern in (False, True):
            v = verify.Verifier(fam, cfg_scale=3.0, lantern=lantern, lantern_k=1000, lantern_delta=0.1,
                                nbr_table=table if lantern else None, device=dev)
            for _ in range(3):
                v.greedy(cond, uncond, tokens, retrieve)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                r = v.greedy(cond, uncond, tokens, retrieve)
            e1.record()
            torch.cuda.synchronize()
            key = f"anole_B{B}_{'relaxed' if lantern else 'plain'}"
            out[key] = {"us_per_step": round(e0.elapsed_time(e1) / 50 * 1e3, 1),
                        "mean_accept": round(float(r.accept_length.float().mean()), 2)}
            print(key, out[key], flush=True)
    json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
