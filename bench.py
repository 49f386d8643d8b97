#!/usr/bin/env python
"""Benchmark of the LANTERN verification hot path (see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W            # B200 arm (prints ONE JSON line on rank 0)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the reference algorithm on host cores

A "step" is one verify step (tree_decoding post-processing -> evaluate_posterior -> bonus token) over one batch
of ``--items`` prompts of synthetic logits of the named family's shapes.  Default workload = BASELINE.json
configs[1]: Lumina-mGPT 768px shapes (V=65536, 8192 image tokens, CFG-doubled logits, top-k 2000, cfg 3.0),
EAGLE-2 tree of 59 nodes / depth 5, LANTERN k=1000 delta=0.1.  The 7B target / drafter forward passes are not
part of this path and are not run.

value  = images/s of the verification path with inputs resident in HBM:
         tokens emitted per second (sum over prompts of accept_length+1) / tokens per image.
e2e    = the same through the host-buffer session call (lantern_session_step): logits in pinned host memory,
         host->device copy of the live logits window and device->host copy of the results inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

TOKENS_PER_IMAGE = {"lumina_mgpt": 2354, "anole": 1024, "llamagen": 256}   # generate_images.py:209,215-218
DEPTH = {"lumina_mgpt": 5, "anole": 4, "llamagen": 4}                       # ea_model_*.from_pretrained defaults


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--family", default="lumina_mgpt", choices=list(TOKENS_PER_IMAGE))
    ap.add_argument("--items", type=int, default=64, help="prompts per GPU per step")
    ap.add_argument("--total-tokens", type=int, default=59)
    ap.add_argument("--logits-dtype", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--lantern-k", type=int, default=1000)
    ap.add_argument("--lantern-delta", type=float, default=0.1)
    ap.add_argument("--top-k", type=int, default=2000)
    ap.add_argument("--cfg", type=float, default=3.0)
    ap.add_argument("--pool", type=int, default=3, help="distinct input batches cycled through (each > L2)")
    ap.add_argument("--cpu-items", type=int, default=0, help="items in the CPU sample (0 = auto)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-torch", action="store_true", help="skip the GPU-PyTorch baseline (reference op sequence)")
    ap.add_argument("--torch-items", type=int, default=8, help="prompts in the GPU-PyTorch baseline sample")
    ap.add_argument("--no-lazy", action="store_true", help="skip the lazy-statistics side measurement")
    ap.add_argument("--no-graph", action="store_true", help="plain launches instead of CUDA-graph replay")
    ap.add_argument("--schedule", default="auto", choices=["streamed", "lazy", "auto"],
                    help="timed step: the library's own choice (default: lantern_accept_phases(..., 8), what the Python "
                         "Verifier uses), every tree row through the HBM-bound statistics kernel (the kernel the "
                         "roofline is always quoted on), or statistics computed inside the walk")
    ap.add_argument("--no-extra", action="store_true", help="skip the bf16 / streamed / neighbour-build / config-4 side measurements")
    ap.add_argument("--e2e-steps", type=int, default=50)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# Synthetic workload
# ------------------------------------------------------------------------------------------------
def host_trees(family, total_tokens, n, seed0):
    """n EAGLE-2-shaped trees with sibling-distinct image tokens (host, deterministic)."""
    from lantern_b200 import synth
    from lantern_b200.verify import FAMILIES
    fam = FAMILIES[family]
    out = []
    for i in range(n):
        t = synth.eagle2_tree(seed0 + i, total_tokens, DEPTH[family])
        synth.assign_tokens(seed0 + i, t, fam.col0, fam.col0 + fam.ncols)
        out.append(t)
    return out


def pad_ri(trees):
    L = max(t.retrieve_indices.shape[0] for t in trees)
    D = max(t.retrieve_indices.shape[1] for t in trees)
    ri = np.full((len(trees), L, D), -1, dtype=np.int32)
    for i, t in enumerate(trees):
        r = t.retrieve_indices
        ri[i, :r.shape[0], :r.shape[1]] = r
    return ri


def device_batch(args, fam, trees, seed, device):
    """One batch of device-resident inputs: cond/uncond [B,T,V], tokens [B,T], retrieve [B,L,D], uniforms."""
    import torch
    B, T, V = len(trees), trees[0].T, fam.vocab
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    dt = torch.float32 if args.logits_dtype == "fp32" else torch.bfloat16
    cond = torch.empty(B, T, V, device=device, dtype=torch.float32)
    cond.normal_(0.0, 2.31, generator=g)
    uncond = torch.empty(B, T, V, device=device, dtype=torch.float32)
    uncond.normal_(0.0, 0.8, generator=g)
    uncond += cond
    tok = np.stack([t.tokens for t in trees]).astype(np.int64)
    par = np.stack([t.parent for t in trees]).astype(np.int64)
    bi = np.repeat(np.arange(B)[:, None], T - 1, axis=1).reshape(-1)
    pi, ti = par[:, 1:].reshape(-1), tok[:, 1:].reshape(-1)
    idx = (torch.from_numpy(bi).to(device), torch.from_numpy(pi).to(device), torch.from_numpy(ti).to(device))
    boost = torch.full((bi.shape[0],), 13.0, device=device)
    cond.index_put_(idx, boost, accumulate=True)
    uncond.index_put_(idx, boost, accumulate=True)
    cond, uncond = cond.to(dt), uncond.to(dt)
    tokens = torch.from_numpy(tok.astype(np.int32)).to(device)
    retrieve = torch.from_numpy(pad_ri(trees)).to(device)
    uni = torch.rand(B, T + 1, device=device, generator=g)
    return dict(cond=cond, uncond=uncond, tokens=tokens, retrieve=retrieve, uniforms=uni)


# ------------------------------------------------------------------------------------------------
# Clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port, eager like the reference) on the host cores
# ------------------------------------------------------------------------------------------------
def _cpu_item(job):
    """One prompt, one verify step, reference cost structure (all T rows post-processed + [L,D,V] gather)."""
    family, total_tokens, seed, k, delta, top_k, cfg, reps = job
    import casegen as C
    b = C.build(dict(family=family, seed=seed, total_tokens=total_tokens, depth=DEPTH[family], lantern_k=k,
                     lantern_delta=delta, top_k=top_k, cfg_scale=cfg))
    from oracle import lantern_oracle as O
    kk = min(k, b.fam.ncols - 1)
    t0 = time.perf_counter()
    tokens = 0
    for _ in range(reps):
        r = O.verify_step(b.cond, b.uncond, cfg, b.tree.tokens, b.tree.retrieve_indices, b.uniforms, b.fam, b.warp,
                          True, kk, delta, b.table, row_kinds=b.row_kinds, eager=True)
        tokens += r.accept_length + 1
    return time.perf_counter() - t0, tokens


def cpu_reference(args, n_items, reps, cores):
    """Returns (images/s, seconds of wall time, tokens) for n_items x reps verify steps on `cores` processes."""
    import multiprocessing as mp
    jobs = [(args.family, args.total_tokens, 90000 + i, args.lantern_k, args.lantern_delta, args.top_k, args.cfg, reps)
            for i in range(n_items)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_item, jobs[:cores])                 # build caches / warm numpy
        t0 = time.perf_counter()
        res = pool.map(_cpu_item, jobs, chunksize=1)
        wall = time.perf_counter() - t0
    # the timed region should be the verify steps, not input synthesis: use the summed in-step time / cores
    step_time = sum(r[0] for r in res) / cores
    tokens = sum(r[1] for r in res)
    return tokens / step_time / TOKENS_PER_IMAGE[args.family], step_time, tokens, wall


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_items = args.cpu_items or max(cores, 16)
    vals = []
    for _ in range(args.warmup > 0 and 1 or 0):
        cpu_reference(args, min(n_items, cores), 1, cores)
    t_all, tok_all = 0.0, 0
    for _ in range(args.steps):
        v, st, tok, _ = cpu_reference(args, n_items, 1, cores)
        t_all += st
        tok_all += tok
    value = tok_all / t_all / TOKENS_PER_IMAGE[args.family]
    line = {
        "impl": "reference", "metric": "images/sec (verification hot path)", "value": value, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_all / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(args, n_items), launch="host processes, one per core"),
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{n_items} prompts x {args.steps} verify steps, oracle/lantern_oracle.py eager "
                                   f"(all T rows post-processed + [L,D,V] gather, as the reference does), one process per core"},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, items):
    from lantern_b200.verify import FAMILIES
    fam = FAMILIES[args.family]
    return {"workload": f"{args.family} verify step (BASELINE configs[1] shapes)" if args.family == "lumina_mgpt"
            else f"{args.family} verify step", "family": args.family, "vocab": fam.vocab, "image_tokens": fam.ncols,
            "prompts_per_gpu": items, "tree": f"EAGLE-2 dynamic, {args.total_tokens} nodes, depth {DEPTH[args.family]}",
            "cfg_scale": args.cfg, "top_k": args.top_k, "temperature": 1.0, "lantern_k": args.lantern_k,
            "lantern_delta": args.lantern_delta, "logits": args.logits_dtype,
            "tokens_per_image": TOKENS_PER_IMAGE[args.family], "launch": "plain" if getattr(args, "no_graph", False) else "cuda-graph replay",
            "schedule": getattr(args, "schedule", "auto"),
            "synthetic_logits": "cond ~ N(0, 2.31^2), uncond = cond + N(0, 0.8^2), +13 on every drafted token in its "
                                "parent's row (lantern_b200.synth shapes; SURVEY 8(d) names N(0, 2.5^2) + 6 - the mean "
                                "accept length is a property of this choice, not of a model)",
            "l2_policy": "inputs larger than L2: a pool of distinct batches, each > 126 MB of live logits"}


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa(local):
    """Pin this rank's host threads (and therefore its first-touch pinned buffers) to the NUMA node of its GPU, so that
    the in-place PCIe reads of the host-buffer route do not cross sockets when several ranks run on one box."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) > 4:
            bus = bus[-12:]                               # nvml pads the PCI domain to 8 hex digits
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.extend(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


def run_b200(args):
    import torch
    import torch.distributed as dist
    from lantern_b200 import _abi, verify, synth
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa(local) if world > 1 else None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    fam = verify.FAMILIES[args.family]
    B, T = args.items, args.total_tokens
    k = min(args.lantern_k, fam.ncols - 1)
    table_np = synth.neighbor_table(0, fam.ncols, k + 1)
    table = torch.from_numpy(table_np).to(dev)
    ver = verify.Verifier(fam, temperature=1.0, top_k=args.top_k, cfg_scale=args.cfg, lantern=True, lantern_k=k,
                          lantern_delta=args.lantern_delta, nbr_table=table, device=dev)
    tree_pool = host_trees(args.family, T, 16, 1000 + 97 * rank)
    batches = []
    for p in range(args.pool):
        trees = [tree_pool[(i + 5 * p) % len(tree_pool)] for i in range(B)]
        batches.append(device_batch(args, fam, trees, 1234 + 1000 * rank + p, dev))
    eb = 4 if args.logits_dtype == "fp32" else 2

    default_phases = {"streamed": 3, "lazy": 6, "auto": 8}[args.schedule]

    def step(i, phases=None):
        phases = default_phases if phases is None else phases
        bt = batches[i % len(batches)]
        return ver.step(bt["cond"], bt["uncond"], bt["tokens"], bt["retrieve"], uniforms=bt["uniforms"], phases=phases)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    torch.cuda.synchronize()
    # One CUDA graph per input batch (both kernels of the step): replay removes the Python/ctypes launch overhead,
    # which is comparable to the GPU time of a step.  --no-graph times plain launches instead.
    graphs, graph_out, kgraphs = [], [], []
    if not args.no_graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(len(batches)):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    r = step(i)
                graphs.append(g)
                graph_out.append(r)
                g1 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g1, stream=side):
                    step(i, phases=1)
                kgraphs.append(g1)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()

    def run(i):
        if graphs:
            graphs[i % len(graphs)].replay()
            return graph_out[i % len(graphs)]
        return step(i)

    def run_stats(i):
        if kgraphs:
            kgraphs[i % len(kgraphs)].replay()
        else:
            step(i, phases=1)

    # the other schedule of the same step, measured beside the headline: "streamed_stats" (phases = 3: every tree row
    # through the HBM-bound kernel, then the walk) when the headline resolves to the lazy schedule, else "lazy_stats"
    lazy_eligible = fam.ncols in (2048, 4096, 8192, 16384)
    headline_lazy = default_phases == 6 or (default_phases == 8 and B * T >= 2048 and lazy_eligible)
    side_phases, side_name = (3, "streamed_stats") if headline_lazy else (6, "lazy_stats")
    lazy_graphs = []
    if not args.no_graph and not args.no_lazy and (side_phases == 3 or lazy_eligible):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(len(batches)):
                step(i, phases=side_phases)
            for i in range(len(batches)):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    r = step(i, phases=side_phases)
                lazy_graphs.append((g, r))
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()

    for i in range(3):
        run(i)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    results = []
    with ClockSampler(local) as clocks:
        # A step is ~0.1 ms of GPU work, the same order as one host launch: park the GPU on a ~1 ms spin kernel before
        # the first event so that the K launches are already queued when the timed region starts and host jitter (the
        # clock sampler starting, a busy host after the CPU arm) cannot leak into a device-side measurement.  The
        # spin itself is outside the two events.
        torch.cuda._sleep(2_000_000)
        ev0.record()
        for i in range(args.steps):
            results.append(run(i))
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        # dominant kernel alone (row statistics), same inputs, events on the launching stream; long enough for
        # the clock sampler to see the GPU under load
        for i in range(3):
            run_stats(i)
        torch.cuda.synchronize()
        n_k = max(args.steps, 200)
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for i in range(n_k):
            run_stats(i)
        k1.record()
        torch.cuda.synchronize()
        ms_stats = k0.elapsed_time(k1) / n_k
        ms_lazy, tokens_lazy = None, 0
        if lazy_graphs:
            for g, _ in lazy_graphs:
                g.replay()
            torch.cuda.synchronize()
            l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0.record()
            for i in range(args.steps):
                lazy_graphs[i % len(lazy_graphs)][0].replay()
            l1.record()
            torch.cuda.synchronize()
            ms_lazy = l0.elapsed_time(l1) / args.steps
            tokens_lazy = sum(int((lazy_graphs[i % len(lazy_graphs)][1].accept_length.sum() + B).item())
                              for i in range(args.steps))
        t_end = time.time() + 1.0            # keep the device busy for ~1 s so the sampler gets >= 5 readings
        while time.time() < t_end:
            for i in range(50):
                run_stats(i)
            torch.cuda.synchronize()
    tokens = sum(int((r.accept_length.sum() + B).item()) for r in results)
    accept_mean = tokens / (args.steps * B)

    # ---- side measurements (never part of `value`) ----
    extra = None
    if not args.no_extra:
        extra = run_extras(args, fam, ver, batches, dev, world, rank, eb, peak_hbm())

    # ---- end to end through the host-buffer session (pinned host logits) ----
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, fam, ver, table_np, tree_pool, rank, dev, world)

    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    tk = torch.tensor([float(tokens)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tk, op=dist.ReduceOp.SUM)
        if e2e is not None:
            mx = torch.tensor([e2e["ms"]], device=dev, dtype=torch.float64)
            sm = torch.tensor([float(e2e["tokens"])], device=dev, dtype=torch.float64)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            e2e["ms"], e2e["tokens"] = float(mx[0]), float(sm[0])
    ms_all, tokens_all = float(t[0]), float(tk[0])
    tpi = TOKENS_PER_IMAGE[args.family]
    value = tokens_all / (ms_all * 1e-3) / tpi

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        stat_bytes = B * 2 * T * fam.ncols * eb           # algorithmic bytes of the row-statistics kernel
        achieved = stat_bytes / (ms_stats * 1e-3) / 1e9
        n_try = float(np.mean([float(r.n_draws.float().mean().item()) - 1 for r in results]))
        step_bytes = B * (2 * T * fam.ncols * eb + n_try * (k + 1) * 4 + 16)
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get(f"{args.family}_{args.logits_dtype}_B{B}_T{T}")
        except Exception:
            pass
        line = {
            "metric": "images/sec (verification hot path)", "value": value, "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_all / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, B),
            "mean_accept_length": accept_mean,
            "verify_steps_per_s": world * B * args.steps / (ms_all * 1e-3),
            "accept_step_gbs": step_bytes / (ms_all / args.steps * 1e-3) / 1e9,
            "clocks": clocks.summary(),
            "host_numa_node_rank0": numa,
            "gpu_launches": (1 if headline_lazy else 2) * args.steps,
            "schedule_resolved": "lazy (one kernel: the walk computes the statistics of the rows it visits)" if headline_lazy
                                 else "streamed (row statistics kernel over every tree row + walk kernel)",
            "roofline": {"bound": "hbm", "kernel": "row_stats_stream_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst; kernel timed alone)" if peaks else "fallback 6650 GB/s",
                         "bytes_per_launch": stat_bytes, "ms_per_launch": ms_stats,
                         "frac_of_nominal_8TBs": achieved / 8000.0,
                         "note": "timed alone (phases=1) whatever schedule the headline uses"},
        }
        if ms_lazy is not None:
            line[side_name] = {"ms_per_step": ms_lazy, "value": tokens_lazy / (ms_lazy * args.steps * 1e-3) / tpi,
                               "unit": "images/s (this rank)", "gpu_launches_per_step": 1 if side_phases == 6 else 2,
                               "note": "phases=6: statistics only for the rows the walk visits (no streamed kernel)"
                                       if side_phases == 6 else
                                       "phases=3: the HBM-bound row-statistics kernel over every tree row, then the walk"}
        if extra:
            line["extra"] = extra
        if e2e is not None:
            line["e2e"] = {"value": e2e["tokens"] / (e2e["ms"] * 1e-3) / tpi, "unit": "images/s",
                           "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                           "ms_per_step": e2e["ms"] / e2e["steps"], "steps": e2e["steps"], "route": e2e["route"], "l2_policy": e2e["l2_policy"]}
        if not args.no_cpu and world == 1:
            cores = os.cpu_count() or 1
            n_items = args.cpu_items or max(cores, 16)
            v, st, tok, wall = cpu_reference(args, n_items, 2, cores)
            line["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": cores, "kind": "port",
                                    "sample": f"{n_items} prompts x 2 verify steps of the same workload, "
                                              f"oracle/lantern_oracle.py eager, one process per core ({wall:.1f}s wall)"}
        if not args.no_torch and world == 1:
            line["torch_gpu_baseline"] = torch_gpu_baseline(args, fam, batches[0], results and run(0), table_np, k, tpi)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f).get("hbm_gbs", 6650.0))
    except Exception:
        return 6650.0


def run_extras(args, fam, ver, batches, dev, world, rank, eb, peak):
    """Side measurements reported under `extra`: the HBM-bound kernel on bf16 logits (the dtype LlamaGen / Anole run,
    generate_images.py:126-127), the neighbour-table build (N = 1 only), and BASELINE configs[3] as a strong-scaling
    leg (Anole shapes, 64 prompts in total sharded i::world over the ranks, 1024 tokens each)."""
    import torch
    import torch.distributed as dist
    out = {}
    B, T = args.items, args.total_tokens
    # ---- row statistics on bf16 logits ----
    if args.logits_dtype == "fp32":
        b16 = [(bt["cond"].to(torch.bfloat16), bt["uncond"].to(torch.bfloat16), bt) for bt in batches]

        def launch16(i):
            c16, u16, bt = b16[i % len(b16)]
            ver.step(c16, u16, bt["tokens"], bt["retrieve"], uniforms=bt["uniforms"], phases=1)
        for i in range(6):
            launch16(i)
        torch.cuda.synchronize()
        n16 = 120
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(2_000_000)          # the launches below are queued while the GPU is parked (see run_b200)
        e0.record()
        for i in range(n16):
            launch16(i)
        e1.record()
        torch.cuda.synchronize()
        ms16 = e0.elapsed_time(e1) / n16
        by = B * 2 * T * fam.ncols * 2
        out["bf16_row_stats"] = {"ms_per_launch": ms16, "achieved": by / (ms16 * 1e-3) / 1e9, "unit": "GB/s",
                                 "frac": by / (ms16 * 1e-3) / 1e9 / peak, "bytes_per_launch": by,
                                 "note": f"same kernel on bf16 logits (the dtype LlamaGen / Anole run), {n16} back-to-back "
                                         f"launches cycling {len(b16)} batches (each batch is re-read after "
                                         f"{(len(b16) - 1) * by >> 20} MB of other rows)"}
        del b16
    # ---- neighbour-table build ----
    if world == 1 and not args.no_cpu:
        out["neighbor_build"] = neighbor_build_extra(dev)
    # ---- BASELINE configs[3]: Anole, 64 prompts in total over the ranks, 1024 tokens each ----
    from lantern_b200 import generate as G, shard
    gargs = G.parse_args().parse_args(["--model", "anole", "--num_images", "64", "--lantern", "--lantern_k", "1000",
                                       "--lantern_delta", "0.1", "--precision", "bf16", "--set_seed"])
    torch.cuda.synchronize()
    eng = G.StandInEngine(gargs, device=dev)
    mine = shard.shard_indices(64, rank, world)
    eng.run(mine[:1])                                    # warm-up (compiles nothing; first-launch costs)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    recs = eng.run(mine)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    merged = recs
    if world > 1:
        tw = torch.tensor([wall], device=dev, dtype=torch.float64)
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        wall = float(tw[0])
        bucket = [None] * world
        dist.all_gather_object(bucket, recs)
        merged = sorted((r for part in bucket for r in part), key=lambda r: r["index"])
    steps = sum(r["steps"] for r in merged)
    out["strong_config4"] = {
        "workload": "Anole shapes (V=65536, 8192 image tokens, bf16 logits), 64 prompts in total, prompt i on rank i mod W, "
                    "1024 tokens per prompt, stand-in target/drafter (lantern_b200.generate.StandInEngine)",
        "n_gpus": world, "prompts_per_gpu": len(mine), "images_per_s": 64.0 / wall, "wall_s": wall,
        "mean_accept_length": sum(r["tokens"] for r in merged) / max(1, steps),
        "verify_steps_total": steps,
        "accept_checksum": sum((r["index"] + 1) * r["steps"] for r in merged),
        "note": "scaling = strong (total work fixed); accept_checksum is identical at every N because a prompt's inputs "
                "depend on (prompt, step) only; wall = max over ranks, includes the Python loop of the stand-in"}
    del eng
    return out


def neighbor_build_extra(dev):
    """lantern_build_neighbors at the BASELINE sizes beside a WARMED torch cdist + topk on the same GPU, and how many of
    the first K columns the reference's fp32 arithmetic orders differently from the exact fp64 definition."""
    import torch
    from lantern_b200 import codebook
    res = {}
    for N, d in ((8192, 256), (16384, 8)):
        g = torch.Generator(device=dev)
        g.manual_seed(N + d)
        E = torch.randn(N, d, device=dev, generator=g)
        if d == 8:
            E = E / E.norm(dim=1, keepdim=True)
        K = 1001
        for _ in range(2):
            t = codebook.build_neighbor_table(E, k=K)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            t = codebook.build_neighbor_table(E, k=K)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5

        def ref():
            dist_ = torch.cdist(E, E, p=2)
            dist_.fill_diagonal_(float("inf"))
            return torch.topk(dist_, K, largest=False)[1]
        for _ in range(2):
            r = ref()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            r = ref()
        e1.record()
        torch.cuda.synchronize()
        ms_ref = e0.elapsed_time(e1) / 3
        diff_torch = int((r.to(torch.int32) != t).sum())
        entry = {"ms": ms, "K": K, "torch_cdist_topk_ms_warmed": ms_ref,
                 "positions_differing_from_torch_fp32_cdist_topk": diff_torch, "positions": N * K,
                 "tf32_equiv_tflops_of_whole_build": 2.0 * N * N * d / (ms * 1e-3) / 1e12}
        try:
            from oracle import c_oracle as CO      # bench CPU leg: the checker, timed nowhere
            en = E.cpu().numpy()
            want = CO.neighbor_table(en, K)
            entry["mismatches_vs_fp64_oracle"] = int((want != t.cpu().numpy()).sum())
        except Exception as ex:                    # pragma: no cover
            entry["mismatches_vs_fp64_oracle"] = f"oracle unavailable: {ex}"
        res[f"{N}x{d}"] = entry
        del E, t, r
    return res


def torch_gpu_baseline(args, fam, bt, ours, table_np, k, tpi):
    """SURVEY 8(d) "GPU-PyTorch path": the reference's own op sequence (small ATen kernels + host syncs, table rows
    copied from numpy per candidate) on the same device inputs, one prompt at a time like the reference (batch 1)."""
    import torch
    from oracle import lantern_oracle as O, torch_path as TP
    ofam = {"llamagen": O.LLAMAGEN, "anole": O.ANOLE, "lumina_mgpt": O.LUMINA}[args.family]
    warp = O.Warp(1.0, 1.0, args.top_k)
    n = min(args.torch_items, bt["cond"].shape[0])
    uni = bt["uniforms"].cpu().numpy()

    def one(b):
        return TP.verify_step(bt["cond"][b], bt["uncond"][b], args.cfg, bt["tokens"][b], bt["retrieve"][b], uni[b],
                              ofam, warp, True, k, args.lantern_delta, table_np)
    one(0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    outs = [one(b) for b in range(n)]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tokens = sum(o[1] + 1 for o in outs)
    agree = None
    if ours:
        al = ours.accept_length.cpu().numpy()
        tk = ours.token.cpu().numpy()
        agree = int(sum(int(al[b]) == outs[b][1] and int(tk[b]) == outs[b][2] for b in range(n)))
    return {"value": tokens / dt / tpi, "unit": "images/s", "ms_per_prompt_step": 1e3 * dt / n,
            "sample": f"{n} prompts x 1 verify step of batch 0, oracle/torch_path.py on cuda (reference op sequence, batch 1)",
            "same_accept_and_token_as_b200": None if agree is None else f"{agree}/{n}"}


def run_e2e(args, fam, ver, table_np, tree_pool, rank, dev, world):
    """Host-buffer path through the public API (lantern_b200.session.HostSession -> lantern_session_step): pinned host
    logits in, host results out, synchronous."""
    import torch
    from lantern_b200.session import HostSession
    B, T = args.items, args.total_tokens
    dt = torch.float32 if args.logits_dtype == "fp32" else torch.bfloat16
    trees = [tree_pool[i % len(tree_pool)] for i in range(B)]
    hb = device_batch(args, fam, trees, 777 + rank, dev)
    host = {}
    for name in ("cond", "uncond"):
        h = torch.empty(hb[name].shape, dtype=dt, pin_memory=True)
        h.copy_(hb[name])
        host[name] = h
    tokens = hb["tokens"].cpu().pin_memory()
    retrieve = hb["retrieve"].cpu().pin_memory()
    uni = hb["uniforms"].cpu().pin_memory()
    L, D = retrieve.shape[1:]
    del hb
    torch.cuda.synchronize()
    sess = HostSession(ver, B, T, L, D, logits_dtype=dt, n_uniforms=T + 1)

    def step():
        return sess.step(host["cond"], host["uncond"], tokens, retrieve, uniforms=uni)
    steps = max(3, args.e2e_steps)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    # the rows the in-place route reads fit in L2 many times over, so L2 is flushed (256 MB memset, outside the
    # timed region) before every step; each step is timed on its own, host clock around the synchronous call
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tok, rows_read, ms, in_place = 0, 0, 0.0, False
    for _ in range(steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = step()
        ms += (time.perf_counter() - t0) * 1e3
        tok += int(r.accept_length.sum()) + B
        rows_read += int(r.rows_read.sum())
        in_place = r.in_place
    sess.close()
    eb = 4 if args.logits_dtype == "fp32" else 2
    small = tokens.numel() * 4 + retrieve.numel() * 4 + uni.numel() * 4
    if in_place:     # the walk read the visited rows (cond + uncond windows) straight from pinned host memory
        h2d = int(round(rows_read / steps * 2 * fam.ncols * eb)) + small
    else:            # staged: the live window of every tree row is copied
        width = ((fam.col0 + fam.ncols + 7) & ~7) - (fam.col0 & ~7)
        h2d = 2 * B * T * width * eb + small
    d2h = B * (5 + 2 * D) * 4
    return {"ms": ms, "tokens": tok, "steps": steps, "h2d": h2d, "d2h": d2h,
            "route": "in place: pinned host logits read over PCIe by the lazy walk, visited rows only" if in_place
                     else "staged: live window of every row copied to the device, streamed schedule",
            "l2_policy": "L2 flushed (256 MB memset) before every step, outside the timed region"}


_JSON_OUT = None


def emit(line):
    """The one JSON line of the contract, written to the process's original stdout."""
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


if __name__ == "__main__":
    # Libraries chat on stdout (NCCL prints its version there at NCCL_DEBUG=VERSION/WARN): keep the original stdout
    # for the JSON line only and send everything else to stderr.
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
