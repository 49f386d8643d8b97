"""Phase trace of the walk kernel on the bench workload (profiling build: `python -m lantern_b200.build --variant trace
-DLANTERN_WALK_TRACE`, run with LANTERN_B200_LIB=lantern_b200/variants/lib_trace.so).  Prints, for the chosen schedule,
the distribution of per-prompt walk times (cycles of thread 0 from first to last mark), the draws per prompt, and the
phase-by-phase trace of the slowest and of the median prompt.  Tags: accept.cu TR(...)."""
import ctypes, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
args = bench.parse()
from lantern_b200 import verify, synth, _abi
fam = verify.FAMILIES[args.family]
dev = torch.device("cuda")
B, T = args.items, args.total_tokens
k = min(args.lantern_k, fam.ncols - 1)
table = torch.from_numpy(synth.neighbor_table(0, fam.ncols, k + 1)).to(dev)
ver = verify.Verifier(fam, temperature=1.0, top_k=args.top_k, cfg_scale=args.cfg, lantern=True, lantern_k=k,
                      lantern_delta=args.lantern_delta, nbr_table=table, device=dev)
pool = bench.host_trees(args.family, T, 16, 1000)
batches = [bench.device_batch(args, fam, [pool[(i + 5 * p) % 16] for i in range(B)], 1234 + p, dev) for p in range(3)]
lib = _abi.load()
KT, KB = 120, 256
NAMES = {1: "start", 2: "prologue done", 3: "levels done", 4: "tail dist done", 5: "bonus done", 6: "outputs done"}
def name(t):
    if t in NAMES: return NAMES[t]
    if 21 <= t <= 25: return "  prologue: " + {21: "loads issued / uniforms", 22: "staged (barrier)", 23: "token table", 24: "dedup table", 25: "barrier"}[t]
    if t in (35, 36, 37): return "  " + {35: "node known", 36: "children listed", 37: "prefetches issued"}[t]
    if t == 39: return "  row loads issued + hook done"
    if 10 <= t < 20: return f"L{t-10} begin"
    if 30 <= t < 40: return f"L{t-30} distribution set"
    if 40 <= t < 50: return "  row " + {40: "lifted + local moments (children listed in the shadow)", 41: "moments exchanged", 42: "exp sweep + bracket select + kept mass", 43: "slow select (if any)", 44: "kept sum (slow path)", 45: "exp sweep + parking", 46: "counts exchanged", 47: "histogram scanned", 48: "candidates listed", 49: "ranked"}[t]
    for base, nm in ((100, "try begin"), (700, "own-prob test"), (800, "scan done"), (200, "rejected"), (300, "pre-zero sync"),
                     (400, "zeroed"), (500, "mass reduced"), (600, "rescaled")):
        if base <= t < base + 100: return f"  c{t-base} {nm}"
    return str(t)
for label, ph in (("lazy", 6), ("streamed", 3)):
    for i in range(4):
        bt = batches[i % 3]
        r = ver.step(bt["cond"], bt["uncond"], bt["tokens"], bt["retrieve"], uniforms=bt["uniforms"], phases=ph)
    buf = np.zeros((KB, 2 + 2 * KT), dtype=np.int64)
    rc = lib.lantern_debug_walk_trace(buf.ctypes.data_as(ctypes.c_void_p), KB)
    assert rc == 0, rc
    n = min(B, KB)
    tot = np.array([buf[b, 2 + 2 * (buf[b, 0] - 1) + 1] - buf[b, 3] for b in range(n)])
    draws = buf[:n, 1]
    order = np.argsort(tot)
    print(f"== {label} B={B}: walk cycles min {tot.min()} median {int(np.median(tot))} max {tot.max()}  (1.9 GHz: {tot.max()/1900:.1f} us);"
          f" draws mean {draws.mean():.1f} max {draws.max()}; accept mean {float(r.accept_length.float().mean()):.2f}")
    print("   per-prompt (cycles, draws, accept):", [(int(tot[b]), int(draws[b]), int(r.accept_length[b])) for b in order[::max(1, n // 16)]])
    for which, b in (("slowest", order[-1]), ("median", order[n // 2])):
        print(f"-- {which} prompt {b}: {tot[b]} cycles, {draws[b]} draws")
        prev = buf[b, 3]
        for i in range(buf[b, 0]):
            tag, clk = buf[b, 2 + 2 * i], buf[b, 3 + 2 * i]
            print(f"   {name(int(tag)):28s} {clk - buf[b, 3]:8d}  +{clk - prev}")
            prev = clk
    # phase totals over all prompts
    agg = {}
    for b in range(n):
        prev_tag, prev = None, None
        for i in range(buf[b, 0]):
            tag, clk = int(buf[b, 2 + 2 * i]), buf[b, 3 + 2 * i]
            if prev is not None:
                key = name(tag).strip()
                key = key.split(" ", 1)[1] if key[0] in "Lc" and " " in key and not key.startswith("row") else key
                a = agg.setdefault(key, [0, 0]); a[0] += clk - prev; a[1] += 1
            prev = clk
    print("-- phase totals over all prompts (segment ending at the mark): mean cycles x count per prompt")
    for kname, (c, cnt) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"   {kname:24s} {c / cnt:8.0f} x {cnt / n:5.2f} = {c / n:8.0f} per prompt")
