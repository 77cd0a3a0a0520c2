"""N>1 host logic on CPU: world_size-2 (and 3) gloo runs of the view sharding +
final all_gather, against the single-process result.  The engine here is a
deterministic stand-in (the CUDA engine has no CPU path by design)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from casmvsnet_pl_b200.dist import max_shard, shard_bounds, sharded_depth_inference


def fake_engine(imgs, proj_mats, init_depth_min, depth_interval):
    """depth maps that depend on every input of the view (and only on that view)."""
    B, V, _, H, W = imgs.shape
    base = imgs.mean(dim=(1, 2)) + proj_mats.reshape(B, -1).sum(1).reshape(B, 1, 1)
    dmin = init_depth_min if not torch.is_tensor(init_depth_min) else init_depth_min.reshape(B, 1, 1)
    return {"depth_0": base + dmin, "confidence_2": base[:, ::4, ::4] * depth_interval,
            "depth_2": base[:, ::4, ::4]}


def make_inputs(B):
    g = torch.Generator().manual_seed(7)
    imgs = torch.randn(B, 3, 3, 16, 24, generator=g)
    pm = torch.randn(B, 2, 3, 3, 4, generator=g)
    dmin = 425.0 + torch.arange(B, dtype=torch.float32).reshape(B, 1)
    return imgs, pm, dmin, 2.65


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        imgs, pm, dmin, dint = make_inputs(B)
        out = sharded_depth_inference(fake_engine, imgs, pm, dmin, dint)
        want = fake_engine(imgs, pm, dmin, dint)
        # no reduction on the path => bit-identical to one process, on every rank
        q.put((rank, bool(torch.equal(out["depth_0"], want["depth_0"])),
               bool(torch.equal(out["confidence_2"], want["confidence_2"]))))
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 32, 33):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) == (max_shard(n, world) if n else 0)


@pytest.mark.parametrize("world,B", [(2, 4), (2, 5), (3, 2)])
def test_sharded_inference_equals_single_process(world, B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r for r, _, _ in got) == list(range(world))
    assert all(d and c for _, d, c in got)


def test_single_process_passthrough():
    imgs, pm, dmin, dint = make_inputs(3)
    out = sharded_depth_inference(fake_engine, imgs, pm, dmin, dint)
    assert torch.equal(out["depth_0"], fake_engine(imgs, pm, dmin, dint)["depth_0"])


def test_single_process_empty_batch():
    """world == 1 and B == 0 used to raise TypeError (`local` was None)."""
    imgs, pm, dmin, dint = make_inputs(2)
    out = sharded_depth_inference(fake_engine, imgs[:0], pm[:0], dmin[:0], dint)
    assert out["depth_0"].shape[0] == 0 and out["confidence_2"].shape[0] == 0
