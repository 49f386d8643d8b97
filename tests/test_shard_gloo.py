"""The N>1 path on CPU: world_size 2, gloo backend — prompt sharding covers every prompt once, the statistics
merge is ordered and complete.  (The hot path itself has no collective)"""
import os
import socket

import torch.distributed as dist
import torch.multiprocessing as mp

from lantern_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_prompts, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.shard_indices(n_prompts, rank, world)
    local = [{"index": i, "prompt": f"p{i}", "step_compression": 2.0 + (i % 3), "latency": 0.1 * (i + 1),
              "rank_wall_s": 1.0 + rank} for i in mine]
    merged = shard.merge_statistics(local)
    q.put((rank, mine, [r["index"] for r in merged], shard.summarize(merged)))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_merge_world2():
    world, n = 2, 11
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    seen = sorted(i for _, mine, _, _ in out for i in mine)
    assert seen == list(range(n))                                   # every prompt exactly once
    for _, _, merged_idx, summ in out:
        assert merged_idx == list(range(n))                         # every rank sees the full ordered merge
        assert summ["n"] == n and abs(summ["mean_accept_length"] - (2.0 + sum(i % 3 for i in range(n)) / n)) < 1e-12
        assert abs(summ["images_per_s"] - n / 2.0) < 1e-12          # max over ranks of the wall time


def test_slice_indices_match_run_sh_split():
    assert shard.slice_indices(5000, 0, 3) == list(range(0, 1666))
    assert shard.slice_indices(5000, 1, 3) == list(range(1666, 3332))
    assert shard.slice_indices(5000, 2, 3)[0] == 3332 and shard.slice_indices(5000, 2, 3)[-1] == 4999
    assert shard.shard_indices(10, 1, 4) == [1, 5, 9]
