#!/usr/bin/env python
"""Accept-kernel sweep of BASELINE configs[4] / SURVEY 8(d) config 5: tree nodes T in {1..256} x prompts B in {1..256},
V = 16384 (LlamaGen), fp32 and bf16 logits, k = 1000, delta = 0.1, top-k 2000, cfg 3.

Per point: device time of the row-statistics kernel alone, of the whole step (statistics + walk) and of the lazy
variant, each a CUDA-graph replay timed with CUDA events on the launching stream, L2 flushed (256 MB memset) before
every timed replay.  achieved GB/s = algorithmic bytes (B * 2*T*V*s) / row-statistics time.

usage: python profiles/sweep_accept.py [--out profiles/sweep_r1.json] [--quick]
"""
import argparse, json, os, sys
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lantern_b200 import synth, verify  # noqa: E402


def pad_ri(trees):
    L = max(t.retrieve_indices.shape[0] for t in trees)
    D = max(t.retrieve_indices.shape[1] for t in trees)
    ri = np.full((len(trees), L, D), -1, dtype=np.int32)
    for i, t in enumerate(trees):
        r = t.retrieve_indices
        ri[i, :r.shape[0], :r.shape[1]] = r
    return ri


def make_batch(fam, B, T, dt, dev, seed):
    pool = []
    for i in range(min(B, 8)):
        t = synth.random_tree(seed + i, T, max_depth=6, max_children=10)
        synth.assign_tokens(seed + i, t, fam.col0, fam.col0 + fam.ncols)
        pool.append(t)
    trees = [pool[i % len(pool)] for i in range(B)]
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    V = fam.vocab
    cond = torch.empty(B, T, V, device=dev, dtype=torch.float32).normal_(0.0, 2.31, generator=g)
    uncond = torch.empty(B, T, V, device=dev, dtype=torch.float32).normal_(0.0, 0.8, generator=g)
    uncond += cond
    if T > 1:
        tok = np.stack([t.tokens for t in trees]).astype(np.int64)
        par = np.stack([t.parent for t in trees]).astype(np.int64)
        bi = np.repeat(np.arange(B)[:, None], T - 1, axis=1).reshape(-1)
        idx = tuple(torch.from_numpy(a).to(dev) for a in (bi, par[:, 1:].reshape(-1), tok[:, 1:].reshape(-1)))
        boost = torch.full((bi.shape[0],), 13.0, device=dev)
        cond.index_put_(idx, boost, accumulate=True)
        uncond.index_put_(idx, boost, accumulate=True)
    tokens = torch.from_numpy(np.stack([t.tokens for t in trees]).astype(np.int32)).to(dev)
    retrieve = torch.from_numpy(pad_ri(trees)).to(dev)
    uni = torch.rand(B, T + 1, device=dev, generator=g)
    return dict(cond=cond.to(dt), uncond=uncond.to(dt), tokens=tokens, retrieve=retrieve, uniforms=uni)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="profiles/sweep_r1.json")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    fam = verify.FAMILIES["llamagen"]
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) \
        if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {}
    peak = float(peaks.get("hbm_gbs", 6547.8))
    k = 1000
    table = torch.from_numpy(synth.neighbor_table(0, fam.ncols, k + 1)).to(dev)
    ver = verify.Verifier(fam, temperature=1.0, top_k=2000, cfg_scale=3.0, lantern=True, lantern_k=k,
                          lantern_delta=0.1, nbr_table=table, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    Ts = [1, 4, 16, 64, 256] if args.quick else [1, 2, 4, 8, 16, 32, 64, 128, 256]
    Bs = [1, 16, 256] if args.quick else [1, 2, 4, 8, 16, 32, 64, 128, 256]
    rows = []
    for dtn, dt, eb in (("fp32", torch.float32, 4), ("bf16", torch.bfloat16, 2)):
        for T in Ts:
            for B in Bs:
                bt = make_batch(fam, B, T, dt, dev, 4242 + T * 7 + B)

                def step(ph):
                    return ver.step(bt["cond"], bt["uncond"], bt["tokens"], bt["retrieve"], uniforms=bt["uniforms"],
                                    phases=ph)
                res = {}
                side = torch.cuda.Stream()
                for name, ph in (("stats", 1), ("step", 3), ("lazy", 6)):
                    for _ in range(3):
                        step(ph)
                    torch.cuda.synchronize()
                    side.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(side):
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g, stream=side):
                            out = step(ph)
                    torch.cuda.current_stream().wait_stream(side)
                    torch.cuda.synchronize()
                    g.replay()
                    ms = 0.0
                    for _ in range(args.iters):
                        flush.zero_()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        g.replay()
                        e1.record()
                        e1.synchronize()
                        ms += e0.elapsed_time(e1)
                    res[name] = ms / args.iters * 1e3   # us
                    if name == "step":
                        res["mean_accept"] = float(out.accept_length.float().mean().item())
                    del g
                alg = B * 2 * T * fam.ncols * eb
                gbs = alg / (res["stats"] * 1e-6) / 1e9
                rows.append(dict(dtype=dtn, T=T, B=B, stats_us=round(res["stats"], 2), step_us=round(res["step"], 2),
                                 lazy_us=round(res["lazy"], 2), algorithmic_bytes=alg, stats_gbs=round(gbs, 1),
                                 frac_of_measured_peak=round(gbs / peak, 4), mean_accept=round(res["mean_accept"], 3)))
                print(rows[-1], flush=True)
                del bt
                torch.cuda.empty_cache()
    out = dict(workload="llamagen V=16384, k=1000, delta=0.1, top_k=2000, cfg=3, random trees (depth<=6, <=10 children)",
               peak_gbs=peak, timing="CUDA-graph replay, CUDA events, L2 flushed (256 MB memset) before each replay",
               points=rows)
    json.dump(out, open(args.out, "w"), indent=0)
    md = [f"# Accept-kernel sweep (round 1)\n\n{out['workload']}; {out['timing']}.\n"
          f"Cells: row-statistics kernel GB/s (fraction of the measured {peak:.0f} GB/s) / whole step us / lazy step us.\n"]
    for dtn in ("fp32", "bf16"):
        md.append(f"\n## {dtn} logits\n\n| T \\\\ B | " + " | ".join(str(b) for b in Bs) + " |\n|---|" + "---|" * len(Bs))
        for T in Ts:
            cells = []
            for B in Bs:
                r = next(x for x in rows if x["dtype"] == dtn and x["T"] == T and x["B"] == B)
                cells.append(f"{r['stats_gbs']:.0f} ({100 * r['frac_of_measured_peak']:.0f}%) / {r['step_us']:.0f} / {r['lazy_us']:.0f}")
            md.append(f"| {T} | " + " | ".join(cells) + " |")
    open(os.path.splitext(args.out)[0] + ".md", "w").write("\n".join(md) + "\n")


if __name__ == "__main__":
    main()
