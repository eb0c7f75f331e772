"""Crop sharding across GPUs: one process per GPU, weights replicated, crops split by index, the only exchange is
the gather of the final keypoint tensors.  The reference has no multi-GPU inference path
(its torch.distributed use is training-only, vit_utils/dist_util.py); this is the data-parallel form its README
lists as future work ("parallel batched inference", README.md:323).

Backend-agnostic host logic (nccl on GPUs, gloo on CPU in the tests): torch.distributed does the plumbing.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

__all__ = ["shard_counts", "shard_range", "gather_keypoints", "infer_sharded", "infer_frame_sharded"]


def shard_counts(n: int, world: int) -> list[int]:
    """Crops per rank: contiguous blocks, the remainder goes to the lowest ranks."""
    base, rem = divmod(n, world)
    return [base + (1 if r < rem else 0) for r in range(world)]


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    counts = shard_counts(n, world)
    lo = sum(counts[:rank])
    return lo, lo + counts[rank]


def gather_keypoints(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """local [n_local, K, C] on every rank (n_local = shard_counts(n_total, world)[rank]) -> [n_total, K, C] on every
    rank, in crop order.  Equal shards take one all_gather_into_tensor; ragged ones are padded to the largest shard."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = shard_counts(n_total, world)
    if local.shape[0] != counts[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} crops, expected {counts[rank]}")
    tail = tuple(local.shape[1:])
    if len(set(counts)) == 1:
        out = torch.empty((n_total,) + tail, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    cmax = max(counts)
    padded = torch.zeros((cmax,) + tail, dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    buf = torch.empty((world * cmax,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, padded, group=group)
    return torch.cat([buf[r * cmax: r * cmax + counts[r]] for r in range(world)], 0)


@torch.no_grad()
def infer_sharded(model, crops: torch.Tensor, org_wh: torch.Tensor, group=None):
    """Every rank passes the SAME global batch (crops [N,3,256,192], org_wh [N,2]); each runs its own slice through its
    engine and all ranks return the full keypoints [N,K,3] and argmax [N,K]."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = crops.shape[0]
    lo, hi = shard_range(n, rank, world)
    outs_kp, outs_idx = [], []
    for s in range(lo, hi, model.max_batch):
        e = min(hi, s + model.max_batch)
        kp, idx = model.infer_crops(crops[s:e], org_wh[s:e])
        outs_kp.append(kp)
        outs_idx.append(idx)
    if outs_kp:
        kp, idx = torch.cat(outs_kp, 0), torch.cat(outs_idx, 0)
    else:   # more ranks than crops
        dev = torch.device("cuda", torch.cuda.current_device())
        kp = torch.empty((0, model.num_keypoints, 3), dtype=torch.float32, device=dev)
        idx = torch.empty((0, model.num_keypoints), dtype=torch.int32, device=dev)
    return gather_keypoints(kp, n, group), gather_keypoints(idx.unsqueeze(-1), n, group).squeeze(-1)


@torch.no_grad()
def infer_frame_sharded(model, frame: torch.Tensor, bboxes: torch.Tensor, group=None):
    """Frame-level form (SURVEY.md section 8 rows f1/f2 + e): every rank holds the SAME uint8 frame [H,W,3] and the SAME boxes
    [n,4]; the people of the frame are sharded by index, each rank runs its slice through `model.infer_frame` (crop
    pre-processing, model, decode, offsets back to frame pixels on its GPU) and all ranks return the full frame-space
    keypoints [n,K,3] and argmax [n,K] in box order.  The only exchange is the keypoint gather."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = bboxes.shape[0]
    lo, hi = shard_range(n, rank, world)
    outs_kp, outs_idx = [], []
    for s in range(lo, hi, model.max_batch):
        kp, idx = model.infer_frame(frame, bboxes[s:min(hi, s + model.max_batch)])
        outs_kp.append(kp)
        outs_idx.append(idx)
    if outs_kp:
        kp, idx = torch.cat(outs_kp, 0), torch.cat(outs_idx, 0)
    else:   # more ranks than people
        kp = torch.empty((0, model.num_keypoints, 3), dtype=torch.float32, device=frame.device)
        idx = torch.empty((0, model.num_keypoints), dtype=torch.int32, device=frame.device)
    return gather_keypoints(kp, n, group), gather_keypoints(idx.unsqueeze(-1), n, group).squeeze(-1)
