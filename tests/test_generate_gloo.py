"""The prompt-sharding launcher (lantern_b200.generate) at world_size 2 over gloo on CPU: `--slice`
selection, rank r takes prompts r, r + W, ... (run.sh:3-16 / generate_images.py:185-192), per-rank and merged
`global_statistics_{start}_{end}.json` in the reference's record format (:297-309), `generation_configs.json`.  The
verification engine is replaced by a deterministic fake (the real one needs a GPU; tests/test_generate_gpu.py)."""
import json
import os
import socket

import torch.multiprocessing as mp

from lantern_b200 import generate as G


class FakeEngine:
    """accept lengths are a pure function of (prompt index, step), like the real engine's inputs"""
    tokens_per_image = 40

    def __init__(self, args):
        self.args = args

    def run(self, indices):
        out = []
        for i in indices:
            tokens = steps = 0
            while tokens < self.tokens_per_image:
                tokens += 1 + (i * 7 + steps * 3) % 5
                steps += 1
            out.append({"index": i, "tokens": tokens, "steps": steps, "step_compression": tokens / steps,
                        "latency": 0.01 * steps, "rank_wall_s": 0.5})
        return out


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, outdir, prompt_file):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    args = G.parse_args().parse_args(["--model", "anole", "--prompt", prompt_file, "--num_images", "100", "--slice", "3-14",
                                      "--output_dir", outdir, "--lantern"])
    G.run_generate_image(args, engine_factory=FakeEngine)


def test_launcher_world2_gloo(tmp_path):
    prompts = tmp_path / "prompts.txt"
    prompts.write_text("".join(f"a photo of object {i}\n" for i in range(20)))
    out = tmp_path / "out"
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, str(out), str(prompts))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    merged = json.load(open(out / "global_statistics_0_10000.json"))
    r0 = json.load(open(out / "global_statistics_0_10000.rank0.json"))
    r1 = json.load(open(out / "global_statistics_0_10000.rank1.json"))
    n = 11                                                   # --slice 3-14 of 20 prompts
    assert sorted(merged) == sorted(f"prompt_{i}" for i in range(n))
    assert sorted(r0) == sorted(f"prompt_{i}" for i in range(0, n, 2))
    assert sorted(r1) == sorted(f"prompt_{i}" for i in range(1, n, 2))
    assert merged["prompt_0"]["prompt"] == "a photo of object 3"          # the slice is applied before sharding
    assert set(merged["prompt_4"]) == {"prompt", "step_compression", "latency"}
    # same numbers as a single-process run: sharding changes who computes a prompt, not what is computed
    single = {r["index"]: r for r in FakeEngine(None).run(range(n))}
    for i in range(n):
        assert merged[f"prompt_{i}"]["step_compression"] == single[i]["step_compression"]
    summ = json.load(open(out / "summary.json"))
    assert summ["n"] == n and summ["world_size"] == 2
    assert abs(summ["mean_accept_length"] - sum(r["step_compression"] for r in single.values()) / n) < 1e-12
    cfg = json.load(open(out / "generation_configs.json"))
    assert cfg["lantern"] is True and cfg["slice"] == "3-14" and cfg["lantern_k"] == 1000


def test_launcher_single_process_and_idx_window(tmp_path):
    args = G.parse_args().parse_args(["--model", "llamagen", "--prompt", "a red cube", "--num_images", "7",
                                      "--start_idx", "2", "--end_idx", "6", "--output_dir", str(tmp_path)])
    os.environ.pop("WORLD_SIZE", None)
    os.environ.pop("RANK", None)
    out = G.run_generate_image(args, engine_factory=FakeEngine)
    assert [r["index"] for r in out["records"]] == [2, 3, 4, 5]           # legacy --start_idx / --end_idx window
    stats = json.load(open(tmp_path / "global_statistics_2_6.json"))
    assert sorted(stats) == ["prompt_2", "prompt_3", "prompt_4", "prompt_5"] and stats["prompt_2"]["prompt"] == "a red cube"


def test_reference_flag_names_and_defaults():
    a = G.parse_args().parse_args([])
    # entrypoints/generate_images.py:47-55 defaults of the hot-path knobs
    assert (a.temperature, a.top_k, a.top_p, a.cfg, a.lantern, a.lantern_k, a.lantern_delta) == (1.0, 2000, 1.0, 3.0, False, 1000, 0.1)
    assert a.model == "lumina_mgpt" and a.slice is None and a.output_dir == "generated_images"
