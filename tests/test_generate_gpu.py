"""lantern_b200.generate on the GPU: the stand-in loop runs the real verification step, and a prompt's accepted
lengths do not depend on how the prompts are sharded (what makes the 1/2/4/8-GPU runs comparable)."""
import json

import pytest

from lantern_b200 import generate as G

pytestmark = pytest.mark.gpu


def _args(tmp_path, *extra):
    return G.parse_args().parse_args(["--model", "llamagen", "--num_images", "6", "--tokens", "48", "--pool", "8",
                                      "--lantern", "--lantern_k", "100", "--top_k", "500", "--precision", "fp32",
                                      "--sync_every", "4", "--output_dir", str(tmp_path), *extra])


def test_generate_records_and_files(tmp_path):
    out = G.run_generate_image(_args(tmp_path))
    recs = out["records"]
    assert [r["index"] for r in recs] == list(range(6))
    for r in recs:
        assert r["tokens"] >= 48 and r["steps"] >= 1 and 1.0 <= r["step_compression"] <= 6.0
        assert sum(r["accept_histogram"]) == r["steps"]
    stats = json.load(open(tmp_path / "global_statistics_0_10000.json"))
    assert set(stats["prompt_3"]) == {"prompt", "step_compression", "latency"}
    assert json.load(open(tmp_path / "summary.json"))["n"] == 6


def test_accept_lengths_do_not_depend_on_sharding(tmp_path):
    eng = G.StandInEngine(_args(tmp_path))
    whole = {r["index"]: r for r in eng.run(range(6))}
    for world in (2, 3):
        for rank in range(world):
            for r in eng.run(range(rank, 6, world)):
                w = whole[r["index"]]
                assert (r["steps"], r["tokens"], r["accept_histogram"]) == (w["steps"], w["tokens"], w["accept_histogram"])
