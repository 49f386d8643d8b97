"""Prompt-sharded generation launcher around the verification hot path — SURVEY.md 8(f) row N4.

Mirrors the reference's ``entrypoints/generate_images.py`` for the part this package owns: the flags that configure
the verification step (:20-71), ``--slice start-end`` prompt selection (:39, :185-192), one process per GPU with disjoint
prompts (``run.sh:3-16``), and the per-prompt statistics file ``global_statistics_{start}_{end}.json`` +
``generation_configs.json`` (:297-309: ``{"prompt", "step_compression" (= mean accepted tokens per step), "latency"}``).

    python -m lantern_b200.generate --model anole --num_images 64 --slice 0-64 --lantern --output_dir out
    python -m torch.distributed.run --nproc-per-node 8 -m lantern_b200.generate --model anole --num_images 64 ...

Under ``torch.distributed`` (WORLD_SIZE > 1) rank r takes prompts r, r + W, ... of the slice (no collective on the hot
path), writes ``global_statistics_{start}_{end}.rank{r}.json``, and rank 0 writes the merged file and ``summary.json``
after one ``all_gather_object``.

The 7B target and the drafter are out of scope: the loop is driven by a stand-in "engine" that
serves, per (prompt, step), synthetic target logits of the model's shapes and an EAGLE-2 draft tree, and runs the real
verification step on them.  Every input of prompt i at step s is a function of (i, s) only - not of the rank, the world
size or the position in the batch - so the accepted lengths of a prompt are bit-identical at 1, 2, 4 and 8 GPUs.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import time
from typing import Callable, Dict, List, Optional, Sequence

from . import shard

TOKENS_PER_IMAGE = {"lumina_mgpt": 2354, "anole": 1024, "llamagen": 256, "llamagen2": 1024}   # generate_images.py:209-218
DEPTH = {"lumina_mgpt": 5, "anole": 4, "llamagen": 4, "llamagen2": 4}                          # ea_model_*.from_pretrained
FAMILY = {"lumina_mgpt": "lumina_mgpt", "anole": "anole", "llamagen": "llamagen", "llamagen2": "llamagen"}


def parse_args() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="LANTERN verification loop on stand-in models, sharded by prompt")
    # reference flags (entrypoints/generate_images.py:20-71); model / checkpoint paths have no meaning for stand-ins
    p.add_argument("--model", type=str, default="lumina_mgpt", choices=list(TOKENS_PER_IMAGE))
    p.add_argument("--model_type", type=str, default="eagle", choices=["eagle"])
    p.add_argument("--precision", type=str, default="bf16", choices=["bf16", "fp32", "fp16"])
    p.add_argument("--prompt", type=str, default="PartiPrompts", help="prompt text, or a file with one prompt per line")
    p.add_argument("--num_images", type=int, default=10)
    p.add_argument("--slice", type=str, default=None, help="slice of prompts to use; format 'start-end'")
    p.add_argument("--output_dir", type=str, default="generated_images")
    p.add_argument("--set_seed", action="store_true")
    p.add_argument("--random_seed", type=int, default=42)
    p.add_argument("--temperature", type=float, default=1.0)
    p.add_argument("--top_k", type=int, default=2000)
    p.add_argument("--top_p", type=float, default=1.0)
    p.add_argument("--cfg", type=float, default=3.0)
    p.add_argument("--lantern", action="store_true")
    p.add_argument("--lantern_k", type=int, default=1000)
    p.add_argument("--lantern_delta", type=float, default=0.1)
    p.add_argument("--start_idx", type=int, default=0)
    p.add_argument("--end_idx", type=int, default=10000)
    # stand-in engine
    p.add_argument("--total_token", type=int, default=59, help="draft tree size (EaModel total_token, default 59)")
    p.add_argument("--tokens", type=int, default=0, help="image tokens per prompt (0 = the model's: 2354 / 1024 / 256)")
    p.add_argument("--pool", type=int, default=64, help="distinct (logits, tree) rows served by the stand-in target")
    p.add_argument("--vq_distances", type=str, default=None,
                   help="top_{K}_indices.npy neighbour table (reference format); default: a synthetic table")
    p.add_argument("--sync_every", type=int, default=32, help="verify steps between host checks for finished prompts")
    return p


def load_prompts(args) -> List[str]:
    """generate_images.py:150-201 for the cases that need no dataset download: a prompt file or one repeated prompt,
    then ``--slice`` (:185-192)."""
    if os.path.isfile(args.prompt):
        with open(args.prompt) as f:
            prompts = [ln.rstrip("\n") for ln in f if ln.strip()]
    else:
        prompts = [args.prompt] * args.num_images
    if args.slice is not None:
        assert re.match(r"^\d+-\d+$", args.slice), f"Invalid format: '{args.slice}'. Expected format is 'start-end'."
        start, end = map(int, args.slice.split("-"))
        assert start < end, f"Invalid range: '{args.slice}'. Start value must be less than end value."
        prompts = prompts[start:end]
    return prompts[:args.num_images]


class StandInEngine:
    """Random-init stand-in for (target, drafter) + the real verification step on the GPU.  ``run(indices)`` advances
    the given prompts in lockstep, one fused verify launch per step for all of them."""

    def __init__(self, args, device=None):
        import numpy as np
        import torch
        from . import synth, verify
        if not torch.cuda.is_available():
            raise RuntimeError("lantern_b200.generate needs a CUDA device (there is no CPU path)")
        self.torch = torch
        self.args = args
        self.dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        fam = verify.FAMILIES[FAMILY[args.model]]
        self.fam = fam
        self.T = args.total_token
        self.tokens_per_image = args.tokens or TOKENS_PER_IMAGE[args.model]
        dt = {"bf16": torch.bfloat16, "fp32": torch.float32, "fp16": torch.float16}[args.precision]
        P = self.P = args.pool
        k = min(args.lantern_k, fam.ncols - 1)
        if args.lantern:
            if args.vq_distances:
                from . import codebook
                table = codebook.load_neighbor_table(args.vq_distances, cols=k + 1, device=self.dev)
            else:
                table = torch.from_numpy(synth.neighbor_table(0, fam.ncols, k + 1)).to(self.dev)
        else:
            table = None
        self.ver = verify.Verifier(fam, temperature=args.temperature, top_k=args.top_k, top_p=args.top_p,
                                   cfg_scale=args.cfg, lantern=args.lantern, lantern_k=k,
                                   lantern_delta=args.lantern_delta, nbr_table=table, device=self.dev)
        # ---- the pool: P rows of (cond, uncond, tree), stored twice back to back so that any cyclic window of rows
        # taken with a constant stride is one strided view (no gather of logits, ever)
        seed = args.random_seed if args.set_seed else 42
        g = torch.Generator(device=self.dev)
        g.manual_seed(seed)
        trees = []
        for i in range(P):
            t = synth.eagle2_tree(seed * 1000 + i, self.T, DEPTH[args.model])
            synth.assign_tokens(seed * 1000 + i, t, fam.col0, fam.col0 + fam.ncols)
            trees.append(t)
        L = max(t.retrieve_indices.shape[0] for t in trees)
        D = max(t.retrieve_indices.shape[1] for t in trees)
        ri = np.full((P, L, D), -1, dtype=np.int32)
        for i, t in enumerate(trees):
            r = t.retrieve_indices
            ri[i, :r.shape[0], :r.shape[1]] = r
        tok = np.stack([t.tokens for t in trees]).astype(np.int64)
        par = np.stack([t.parent for t in trees]).astype(np.int64)
        cond = torch.empty(2 * P, self.T, fam.vocab, device=self.dev, dtype=dt)
        uncond = torch.empty(2 * P, self.T, fam.vocab, device=self.dev, dtype=dt)
        for i in range(P):                      # row by row: the fp32 temporaries stay small
            c = torch.empty(self.T, fam.vocab, device=self.dev).normal_(0.0, 2.31, generator=g)
            u = torch.empty(self.T, fam.vocab, device=self.dev).normal_(0.0, 0.8, generator=g) + c
            pi = torch.from_numpy(par[i, 1:]).to(self.dev)
            ti = torch.from_numpy(tok[i, 1:]).to(self.dev)
            c.index_put_((pi, ti), torch.full((self.T - 1,), 13.0, device=self.dev), accumulate=True)
            u.index_put_((pi, ti), torch.full((self.T - 1,), 13.0, device=self.dev), accumulate=True)
            cond[i] = cond[i + P] = c.to(dt)
            uncond[i] = uncond[i + P] = u.to(dt)
        self.cond, self.uncond = cond, uncond
        self.tokens = torch.from_numpy(np.concatenate([tok, tok]).astype(np.int32)).to(self.dev)
        self.retrieve = torch.from_numpy(np.concatenate([ri, ri])).to(self.dev)
        self.uni_pool = torch.rand(4099, self.T + 1, device=self.dev, generator=g)     # prime row count
        self.L, self.D = L, D

    def run(self, indices: Sequence[int], on_step: Optional[Callable] = None) -> List[Dict]:
        """Generate ``tokens_per_image`` tokens for every prompt index; returns one record per prompt."""
        torch = self.torch
        idx = list(indices)
        if not idx:
            return []
        stride = idx[1] - idx[0] if len(idx) > 1 else 1
        if any(b - a != stride for a, b in zip(idx, idx[1:])) or stride <= 0:
            raise ValueError("prompt indices must be an arithmetic progression (rank, rank + W, ...)")
        out: List[Dict] = []
        # prompts are served in chunks whose pool rows do not collide: at most P / stride consecutive local prompts
        per = max(1, self.P // stride)
        for c0 in range(0, len(idx), per):
            out.extend(self._run_chunk(idx[c0:c0 + per], stride))
        return out

    def _run_chunk(self, idx: List[int], stride: int) -> List[Dict]:
        """Lockstep loop of one chunk.  Per step the host issues the fused verify launch and one tiny copy; everything
        else (which row serves which prompt, the uniforms of a block of steps) is prepared per block of ``sync_every``
        steps, and the per-prompt bookkeeping is done afterwards from the [steps, B] matrix of accepted lengths -
        prompts that are already complete keep stepping and are ignored."""
        torch = self.torch
        B, P, T = len(idx), self.P, self.T
        dev = self.dev
        need = self.tokens_per_image
        pid = torch.tensor(idx, device=dev, dtype=torch.long)
        ar = torch.arange(B, device=dev)
        every = self.args.sync_every
        tok_cache, ret_cache = {}, {}
        blocks, times = [], []
        produced = torch.zeros(B, device=dev, dtype=torch.int64)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        s = 0
        while True:
            acc = torch.empty(every, B, device=dev, dtype=torch.int32)
            srange = torch.arange(s, s + every, device=dev)
            uni_blk = self.uni_pool[(pid[None, :] * 31 + srange[:, None] * 17) % self.uni_pool.shape[0]]   # [every, B, T+1]
            for j in range(every):
                start = (idx[0] + s) % P                         # row of prompt i at step s: (i + s) mod P
                if start not in tok_cache:                       # at most P distinct windows: gathered once each
                    rows = start + stride * ar                   # < 2P: the pool is stored twice
                    tok_cache[start], ret_cache[start] = self.tokens[rows], self.retrieve[rows]
                hi = start + stride * (B - 1) + 1
                res = self.ver.step(self.cond[start:hi:stride], self.uncond[start:hi:stride], tok_cache[start],
                                    ret_cache[start], uniforms=uni_blk[j])
                acc[j] = res.accept_length
                s += 1
            produced += (acc.long() + 1).sum(0)
            blocks.append(acc)
            done = bool((produced >= need).all())                # one host sync per `sync_every` steps
            times.append(time.perf_counter() - t0)
            if done:
                break
        wall = time.perf_counter() - t0
        a = torch.cat(blocks).cpu().numpy().astype("int64")      # [steps, B]
        import numpy as np
        cum = np.cumsum(a + 1, axis=0)
        out = []
        for b in range(B):
            n_steps = int(np.argmax(cum[:, b] >= need)) + 1      # steps until the prompt had all its tokens
            hist = np.bincount(a[:n_steps, b], minlength=self.D + 1)
            out.append({"index": idx[b], "tokens": int(cum[n_steps - 1, b]), "steps": n_steps,
                        "step_compression": float(cum[n_steps - 1, b]) / n_steps,
                        "latency": times[(n_steps - 1) // every], "accept_histogram": hist.tolist(),
                        "rank_wall_s": wall, "batch": B})
        return out


def run_generate_image(args, engine_factory: Optional[Callable] = None) -> Dict:
    """generate_images.py:250-309 on the stand-in engine.  ``engine_factory(args)`` may supply another engine (the CPU
    tests drive the sharding / statistics logic with a fake one); the default is the CUDA engine."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    own_group = False
    if world > 1 and not dist.is_initialized():
        backend = "nccl" if engine_factory is None else "gloo"
        if backend == "nccl":
            import torch
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend)
        own_group = True
    prompts = load_prompts(args)
    sel = [i for i in range(len(prompts)) if args.start_idx <= i < args.end_idx]
    mine = [sel[j] for j in shard.shard_indices(len(sel), rank, world)]
    os.makedirs(args.output_dir, exist_ok=True)
    engine = (engine_factory or StandInEngine)(args)
    t0 = time.perf_counter()
    records = engine.run(mine)
    wall = time.perf_counter() - t0
    for r in records:
        r["prompt"] = prompts[r["index"]]
        r["rank"] = rank
        r["rank_wall_s"] = wall
    stats = {f"prompt_{r['index']}": {"prompt": r["prompt"], "step_compression": r["step_compression"],
                                      "latency": r["latency"]} for r in records}
    tag = f"{args.start_idx}_{args.end_idx}"
    name = f"global_statistics_{tag}.json" if world == 1 else f"global_statistics_{tag}.rank{rank}.json"
    with open(os.path.join(args.output_dir, name), "w") as f:
        json.dump(stats, f, indent=4)
    merged = shard.merge_statistics(records)
    summary = shard.summarize(merged)
    summary.update({"world_size": world, "tokens_per_image": getattr(engine, "tokens_per_image", None),
                    "accept_checksum": sum((r["index"] + 1) * r["steps"] for r in merged)})
    if rank == 0:
        if world > 1:
            with open(os.path.join(args.output_dir, f"global_statistics_{tag}.json"), "w") as f:
                json.dump({f"prompt_{r['index']}": {"prompt": r["prompt"], "step_compression": r["step_compression"],
                                                    "latency": r["latency"]} for r in merged}, f, indent=4)
        with open(os.path.join(args.output_dir, "generation_configs.json"), "w") as f:
            json.dump(vars(args), f, indent=4)
        with open(os.path.join(args.output_dir, "summary.json"), "w") as f:
            json.dump(summary, f, indent=4)
    if own_group:
        dist.barrier()
        dist.destroy_process_group()
    return {"records": merged, "summary": summary}


def main(argv=None):
    args = parse_args().parse_args(argv)
    if args.set_seed:
        import random
        random.seed(args.random_seed)
    out = run_generate_image(args)
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps(out["summary"]))


if __name__ == "__main__":
    main()
