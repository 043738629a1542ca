"""Ensemble sharding + score gather on CPU: world size 2, gloo backend, the per-member run stubbed.

The reference evaluates a population with Pool.map (wmf.py:872-878); wmf_b200.ensemble shards members round-robin over
ranks and exchanges only the (members, 2) score table.  The GPU run itself is covered by tests/test_gpu_shia.py."""
import os
import socket

import numpy as np
import pytest


def test_shard_members_partition():
    from wmf_b200.ensemble import shard_members
    for P in (0, 1, 7, 64, 4096):
        for world in (1, 2, 3, 8):
            seen = np.concatenate([shard_members(P, r, world) for r in range(world)])
            assert sorted(seen.tolist()) == list(range(P))
            sizes = [shard_members(P, r, world).size for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_members(4, 2, 2)


def _fake_scores(cb):
    # deterministic "objective" of a calibration set, so every rank can predict every member's score
    cb = np.asarray(cb, np.float64)
    return np.stack([cb.sum(1), (cb * np.arange(1, 12)).sum(1)], 1).astype(np.float32)


def _worker(rank, world, port, P, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from wmf_b200.ensemble import evaluate_population, shard_members
        rng = np.random.default_rng(0)
        calibs = rng.uniform(0.1, 2.0, (P, 11)).astype(np.float32)
        calls = []

        def run_batch(cb):
            calls.append(cb.shape[0])
            return torch.from_numpy(_fake_scores(cb))

        got = evaluate_population(None, calibs, None, 3, 10, None, None, members_per_launch=3, run_batch=run_batch)
        want = _fake_scores(calibs)
        ok = bool(np.array_equal(got.numpy(), want)) and sum(calls) == shard_members(P, rank, world).size and max(calls) <= 3
        # the shared rain upload: every rank contributes one block of records, all ranks end up with the whole table
        from wmf_b200.ensemble import share_rain
        for R in (1, 2, P, 2 * P + 1):
            table = rng.integers(0, 90000, (R, 37)).astype(np.int32)
            ok = ok and bool(np.array_equal(share_rain(table, "cpu").numpy(), table))
        q.put((rank, ok, got.shape))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("P", [7, 8])
def test_evaluate_population_gloo_world2(P):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, P, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert all(tuple(shape) == (P, 2) for _, _, shape in res)


def test_single_process_no_group():
    import torch
    from wmf_b200.ensemble import evaluate_population
    calibs = np.ones((5, 11), np.float32) * np.arange(1, 6, dtype=np.float32)[:, None]
    got = evaluate_population(None, calibs, None, 1, 4, None, None, members_per_launch=2,
                              run_batch=lambda cb: torch.from_numpy(_fake_scores(cb)))
    assert np.array_equal(got.numpy(), _fake_scores(calibs))
