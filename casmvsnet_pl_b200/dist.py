"""Data-parallel depth inference over the GPUs of one box.

Reference views are independent (eval.py:213-222 has no cross-iteration state and
eval-mode BN has no cross-batch ops), so the batch of reference views is sharded
over ranks with NO collective on the data path; the only exchange is ONE all_gather
of the per-view results at the end (depth_0 and the confidence_2 map eval.py:224-226
consumes).  One process per GPU (torchrun), NCCL on GPUs, gloo for the CPU tests of
this host logic.  Results are bit-identical to a single-rank run (no reduction).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world: int):
    """Contiguous, balanced split: the first n % world ranks get one extra item."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_shard(n_items: int, world: int) -> int:
    return -(-n_items // world)


def sharded_depth_inference(engine, imgs, proj_mats, init_depth_min, depth_interval,
                            keys=("depth_0", "confidence_2"), group=None):
    """Run `engine(imgs, proj_mats, init_depth_min, depth_interval) -> dict` on this
    rank's shard of the reference views and all_gather the requested outputs.

    imgs (B,V,3,H,W), proj_mats (B,V-1,3,3,4) hold ALL B reference views on every rank
    (or at least this rank's shard rows; only those are touched).  init_depth_min /
    depth_interval: float or (B,1) tensors.  Returns {key: (B,h,w)} on every rank.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = imgs.shape[0]
    lo, hi = shard_bounds(B, rank, world)

    def pick(x):
        return x[lo:hi] if torch.is_tensor(x) else x

    if hi > lo:
        local = engine(imgs[lo:hi], proj_mats[lo:hi], pick(init_depth_min), pick(depth_interval))
    else:
        local = None
    if world == 1:
        if local is None:                      # empty batch: nothing to infer, nothing to gather
            return {k: torch.zeros(0, 0, 0, dtype=torch.float32, device=imgs.device) for k in keys}
        return {k: local[k] for k in keys}

    cap = max_shard(B, world)
    if B >= world:
        # every rank has views: each knows the output shapes from its own results
        shapes = [tuple(local[k].shape[-2:]) for k in keys]
    else:
        # some shard is empty: that rank learns the shapes from the others (control message)
        t = torch.zeros(len(keys), 2, dtype=torch.int64, device=imgs.device)
        if local is not None:
            for i, k in enumerate(keys):
                t[i, 0], t[i, 1] = local[k].shape[-2:]
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        shapes = [(int(t[i, 0]), int(t[i, 1])) for i in range(len(keys))]
    out = {}
    for i, k in enumerate(keys):
        h, w = shapes[i]
        send = torch.zeros(cap, h, w, dtype=torch.float32, device=imgs.device)
        if local is not None:
            send[: hi - lo] = local[k]
        recv = torch.empty(world * cap, h, w, dtype=torch.float32, device=imgs.device)
        dist.all_gather_into_tensor(recv, send, group=group)   # the path's single collective
        if B == world * cap:                   # balanced shards: the gathered tensor IS the result
            out[k] = recv
        else:
            parts = []
            for r in range(world):
                rlo, rhi = shard_bounds(B, r, world)
                parts.append(recv[r * cap: r * cap + rhi - rlo])
            out[k] = torch.cat(parts, 0)
    return out
