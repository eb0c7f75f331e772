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


class _FakeFrameModel:
    """Stands in for the engine on CPU: keypoints that encode (box index given by x0, keypoint, frame checksum), so the
    sharded call can be checked for order and coverage without a GPU."""
    max_batch = 2
    num_keypoints = 3

    def infer_frame(self, frame, bboxes):
        n = bboxes.shape[0]
        tag = float(frame.to(torch.int64).sum() % 97)
        kp = torch.stack([bboxes[:, 0:1].float().expand(n, 3), torch.arange(3.0).expand(n, 3), torch.full((n, 3), tag)], -1)
        return kp, (bboxes[:, 0:1] * 10 + torch.arange(3)).to(torch.int32)


def _frame_worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from easy_vitpose_b200.distributed import infer_frame_sharded
        frame = (torch.arange(6 * 8 * 3) % 251).to(torch.uint8).reshape(6, 8, 3)
        boxes = torch.stack([torch.arange(n_total), torch.zeros(n_total, dtype=torch.int64), torch.arange(n_total) + 5,
                             torch.full((n_total,), 9)], 1).to(torch.int32)
        kp, idx = infer_frame_sharded(_FakeFrameModel(), frame, boxes)
        want_kp, want_idx = _FakeFrameModel().infer_frame(frame, boxes)
        q.put((rank, bool(torch.equal(kp, want_kp)) and bool(torch.equal(idx, want_idx))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [7, 1])
def test_infer_frame_sharded_world2_gloo(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() + n_total) % 2000
    procs = [ctx.Process(target=_frame_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(0, True), (1, True)]
