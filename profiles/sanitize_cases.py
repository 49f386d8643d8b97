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


def main():
    torch.cuda.set_device(0)
    which = sys.argv[1:] or ["accept", "neighbors", "kv"]
    if "accept" in which:
        accept(dict(family="llamagen", ncols=2048, top_k=300, lantern_k=100, boost=11.0))            # generic statistics kernel
        accept(dict(family="lumina_mgpt", ncols=2048, top_k=500, lantern_k=100, depth=5))            # fast (TMA) statistics kernel
        accept(dict(family="lumina_mgpt", ncols=4096, top_k=500, lantern_k=100, depth=5), phases=6)  # lazy walk
        accept(dict(family="anole", ncols=1024, top_k=0, top_p=0.9, lantern_k=50))                   # top-p kernel
        accept(dict(family="llamagen", ncols=2048, top_k=300, lantern_k=20, lantern_delta=5.0,
                    static_tree="mc_sim_7b_63", seed=500))                                           # static tree / LANTERN++
        print("accept ok", flush=True)
    if "neighbors" in which:
        for N, d, K in ((1024, 8, 65), (512, 256, 33), (300, 8, 299)):
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


if __name__ == "__main__":
    main()
