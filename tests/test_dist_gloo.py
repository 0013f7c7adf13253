"""CPU, world_size 2 over gloo: the host-side multi-GPU logic (env sharding and the
episode-statistics all-reduce).  The step path itself has no collective."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from walker_gym_b200.dist import all_reduce_stats, finalize_stats, shard_range


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 8, 1000, 1 << 20, (1 << 20) + 3):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0
            for (o0, c0), (o1, _) in zip(spans, spans[1:]):
                assert o0 + c0 == o1
            assert spans[-1][0] + spans[-1][1] == total
            counts = [c for _, c in spans]
            assert max(counts) - min(counts) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_finalize_stats():
    s = finalize_stats([10.0, 60.0, 30.0, 2.0, 0, 0, 0, 0])
    assert s["episodes"] == 2 and s["return_mean"] == 5.0 and s["length_mean"] == 15.0
    assert abs(s["return_std"] - (30.0 - 25.0) ** 0.5) < 1e-12
    assert finalize_stats([0.0] * 8)["episodes"] == 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        off, cnt = shard_range(total, rank, world)
        # per-env finished-episode accumulators of the *global* job; each rank reduces its shard
        rng = np.random.default_rng(0)
        ret = rng.normal(size=total)
        length = rng.integers(1, 100, size=total).astype(np.float64)
        sl = slice(off, off + cnt)
        vec = torch.tensor([ret[sl].sum(), (ret[sl] ** 2).sum(), length[sl].sum(), float(cnt), 0, 0, 0, 0],
                           dtype=torch.float64)
        out = all_reduce_stats(vec)
        q.put((rank, off, cnt, out.tolist()))
    finally:
        dist.destroy_process_group()


def test_stats_all_reduce_world2_gloo():
    world, total = 2, 1001
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    ret = rng.normal(size=total)
    length = rng.integers(1, 100, size=total).astype(np.float64)
    assert sorted(c for _, _, c, _ in results) == [500, 501]
    for _, _, _, vec in results:
        np.testing.assert_allclose(vec[:4], [ret.sum(), (ret ** 2).sum(), length.sum(), total], rtol=1e-12)
        s = finalize_stats(vec)
        np.testing.assert_allclose(s["return_mean"], ret.mean(), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(s["return_std"], ret.std(), rtol=1e-9)
