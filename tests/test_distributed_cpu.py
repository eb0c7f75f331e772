"""CPU, world_size 2, gloo: the host-side sharding / gather logic of the multi-GPU path (no GPU compute)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from easy_vitpose_b200.distributed import gather_keypoints, shard_counts, shard_range


def test_shard_ranges_cover_everything():
    for n in (0, 1, 2, 7, 64, 513):
        for world in (1, 2, 3, 8):
            counts = shard_counts(n, world)
            assert sum(counts) == n and max(counts) - min(counts) <= 1
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[r][1] == spans[r + 1][0] for r in range(world - 1))


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        K = 5
        full = torch.arange(n_total * K * 3, dtype=torch.float32).reshape(n_total, K, 3)
        lo, hi = shard_range(n_total, rank, world)
        out = gather_keypoints(full[lo:hi].clone(), n_total)
        q.put((rank, bool(torch.equal(out, full))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 7, 1])
def test_gather_keypoints_world2_gloo(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n_total) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(0, True), (1, True)]
