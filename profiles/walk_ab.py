"""A/B of the lazy walk variants on the bench workload: speculative group on/off, no speculation (phases bit 4)."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
args = bench.parse(); 
from lantern_b200 import verify, synth
fam = verify.FAMILIES[args.family]
dev = torch.device("cuda")
B, T = args.items, args.total_tokens
k = min(args.lantern_k, fam.ncols - 1)
table = torch.from_numpy(synth.neighbor_table(0, fam.ncols, k + 1)).to(dev)
ver = verify.Verifier(fam, temperature=1.0, top_k=args.top_k, cfg_scale=args.cfg, lantern=True, lantern_k=k, lantern_delta=args.lantern_delta, nbr_table=table, device=dev)
pool = bench.host_trees(args.family, T, 16, 1000)
batches = [bench.device_batch(args, fam, [pool[(i + 5 * p) % 16] for i in range(B)], 1234 + p, dev) for p in range(3)]
def run(phases, n=60):
    for i in range(5):
        bt = batches[i % 3]; r = ver.step(bt["cond"], bt["uncond"], bt["tokens"], bt["retrieve"], uniforms=bt["uniforms"], phases=phases)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(3_000_000); e0.record()
    for i in range(n):
        bt = batches[i % 3]; r = ver.step(bt["cond"], bt["uncond"], bt["tokens"], bt["retrieve"], uniforms=bt["uniforms"], phases=phases)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3, float(r.accept_length.float().mean()), float(((r.flags >> 8) & 255).float().mean())
for name, ph in (("lazy", 6), ("lazy, no speculation", 22), ("streamed", 3)):
    us, acc, rows = run(ph)
    print(f"B={B} {name:24s} {us:8.1f} us/step  accept {acc:.2f} rows_read {rows:.2f}")
