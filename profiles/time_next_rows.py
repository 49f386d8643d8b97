#!/usr/bin/env python
"""Device time of the SURVEY 8(f) "next rows" kernels against the reference's own op sequence on the same GPU
(oracle/torch_path.py restatements; baseline legs only).

  N1  lantern_build_dynamic_tree   vs  tail of topK_genrate (cnets_llamagen.py:831-912)
  N2  lantern_kv_compact           vs  per-slab gather + copy_ (ea_model_llamagen.py:962-970)
  N3  lantern_draft_sample         vs  Model.sample (cnets_llamagen.py:924-940)

usage: python profiles/time_next_rows.py [--out profiles/next_rows_r1.json]
"""
import argparse, json, os, statistics, sys, time
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lantern_b200 import draft_sample, dyntree, posterior as PO, synth  # noqa: E402
from oracle import lantern_oracle as O, torch_path as TP  # noqa: E402


def gpu_us(f, n=50, warm=5):
    for _ in range(warm):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def wall_us(f, n=20, warm=3):
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
    ap.add_argument("--out", default="profiles/next_rows_r1.json")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    out = {}

    # ---- N1: dynamic tree, Lumina defaults (depth 5, top_k 10, 59 draft nodes... total_tokens = 58 + root) ----
    for B in (1, 64):
        exps = [synth.eagle2_expansion(70 + i, depth=5, top_k=10, lo=4, hi=8196) for i in range(B)]
        sc = torch.from_numpy(np.stack([e.scores for e in exps])).to(dev)
        tk = torch.from_numpy(np.stack([e.tokens for e in exps])).to(dev)
        pr = torch.from_numpy(np.stack([e.parents for e in exps])).to(dev)
        st = torch.tensor([e.sample_token for e in exps], device=dev)
        ours = lambda: dyntree.build_dynamic_tree(sc, tk, pr, st, 58, top_k=10)
        ref = lambda: [TP.dynamic_tree_tail(sc[b], tk[b], pr[b], st[b:b + 1], 58, 10) for b in range(B)]
        t = ours()
        r = ref()
        same = all(torch.equal(t.reference_outputs(b)[1].cpu(), r[b][1]) and
                   torch.equal(t.reference_outputs(b)[0].cpu(), r[b][0].cpu()) for b in range(B))
        out[f"N1_dynamic_tree_B{B}"] = {"b200_device_us": round(gpu_us(ours), 1), "b200_call_us": round(wall_us(ours), 1),
                                        "reference_op_sequence_us": round(wall_us(ref, n=5, warm=1), 1),
                                        "same_tokens_and_paths": bool(same)}
        print(f"N1 B={B}", out[f"N1_dynamic_tree_B{B}"], flush=True)

    # ---- N2: KV compaction, Chameleon-7B slab geometry (32 layers x K/V, CFG batch 2, 32 heads, 4096 x 128 bf16) ----
    slabs = [torch.zeros(64, 2, 32, 4096, 128, device=dev, dtype=torch.bfloat16)]   # one slab per device, 4 GiB
    for a in (0, 3, 5):
        sel = torch.tensor(sorted({1000, 1003, 1010, 1024, 1031, 1047}.__iter__())[:a + 1], device=dev)

        def ours():
            return PO.kv_compact(slabs, sel, 1000)

        def ref():
            for s in slabs:                                  # the reference loop: gather, then copy, per device slab
                tgt = s[..., sel, :]
                s[..., 1000:1000 + tgt.shape[-2], :].copy_(tgt, non_blocking=True)
        moved = 64 * 2 * 32 * (a + 1) * 128 * 2 * 2
        u, r = gpu_us(ours), gpu_us(ref, n=10, warm=2)
        out[f"N2_kv_compact_accept{a}"] = {"b200_device_us": round(u, 1), "reference_loop_device_us": round(r, 1),
                                            "bytes_moved": moved, "b200_gbs": round(moved / u / 1e3, 1)}
        print(f"N2 a={a}", out[f"N2_kv_compact_accept{a}"], flush=True)
    del slabs

    # ---- N3: drafter sampling, k = 10 of V = 16384 (LlamaGen) / 65536 with an 8192 window (Lumina) ----
    for name, R in (("llamagen", 10), ("llamagen", 160)):
        g = torch.Generator(device=dev)
        g.manual_seed(7)
        logits = torch.randn(R, 16384, device=dev, generator=g) * 2.5
        proc = PO.prepare_logits_processor(temperature=1.0, top_p=1.0, top_k=2000)
        ours = lambda: draft_sample.sample(logits, proc, 10, seed=1, step=2)
        ref = lambda: TP.drafter_sample(logits, O.Warp(1.0, 1.0, 2000), 10)
        out[f"N3_draft_sample_R{R}"] = {"b200_device_us": round(gpu_us(ours), 1), "b200_call_us": round(wall_us(ours), 1),
                                        "reference_op_sequence_device_us": round(gpu_us(ref, n=10, warm=2), 1),
                                        "reference_op_sequence_call_us": round(wall_us(ref, n=10, warm=2), 1)}
        print(f"N3 R={R}", out[f"N3_draft_sample_R{R}"], flush=True)
    json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
