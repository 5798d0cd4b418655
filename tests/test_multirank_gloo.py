"""world_size-2 gloo tests (CPU) of the sample-parallel host logic used by bench.py at N>1."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from framedipt_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total_b, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from framedipt_b200 import SE3Diffuser, synthetic
        from framedipt_b200.config import default_conf

        # weights: only rank 0 holds the real values before the broadcast
        g = torch.Generator().manual_seed(7)
        sd = {"a.weight": torch.randn(5, 3, generator=g), "b.bias": torch.randn(4, generator=g), "c": torch.randn(2, 2, 2, generator=g)}
        mine = {k: (v.clone() if rank == 0 else torch.full_like(v, float("nan"))) for k, v in sd.items()}
        got = sharding.broadcast_state_dict(mine, dist, 0)
        ok_w = all(torch.equal(got[k], sd[k]) for k in sd)
        # samples: each rank builds its own slice of the batch with a rank-offset seed, like bench.py
        lo, hi = sharding.shard_range(total_b, rank, world)
        wl = synthetic.Workload("t", hi - lo, (10, 6), ((3, 6),), 4)
        diffuser = SE3Diffuser(default_conf().diffuser)  # its ctor seeds the global numpy RNG like the reference's
        np.random.seed(123 + rank)
        feats = synthetic.make_features(wl, diffuser, seed=0)
        local = feats["rigids_t"].float() + 1000.0 * rank
        full = sharding.gather_samples(local, dist, 0)
        q.put((rank, ok_w, (lo, hi), None if full is None else full.numpy(), local.numpy()))
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_batch():
    for total in (1, 7, 8, 32, 33):
        for world in (1, 2, 4, 8):
            spans = [sharding.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1


def test_broadcast_and_gather_world2():
    world, total_b = 2, 5  # ragged: 3 + 2 samples
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total_b, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=180) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), "weights differ after broadcast"
    assert [r[2] for r in res] == [(0, 3), (3, 5)]
    full = res[0][3]
    assert res[1][3] is None and full.shape[0] == total_b
    np.testing.assert_array_equal(full[:3], res[0][4])
    np.testing.assert_array_equal(full[3:], res[1][4])
    # independent chains: different ranks drew different x_T
    assert np.abs(res[0][4][0][3:6] - (res[1][4][0][3:6] - 1000.0)).max() > 1e-2  # the diffused span
