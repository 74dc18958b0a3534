"""The N > 1 host path on CPU: two gloo ranks shard an ensemble by replica blocks, run a
stand-in for the device launch, and gather -- exactly the code path bench.py / run_last_metrics
take under torchrun (with nccl) on the GPU box."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as td  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from jaxabm_b200 import dist, ensemble
        from jaxabm_b200.rules import growth
        import jaxabm_b200 as jx
        R = 11
        lo, hi = dist.shard_range(R)
        local = np.zeros((hi - lo, 8))
        local[:, 0] = np.arange(lo, hi) * 10.0
        local[:, 1] = rank
        full = dist.gather_rows(local, R)
        t = dist.max_over_ranks(1.0 + rank)
        s = dist.sum_over_ranks(1.0 + rank)
        # the real sharded entry point with the device launch replaced by a host stand-in
        models = [growth.create_test_model(initial_value=1.0, num_agents=10, params={"growth_rate": 0.01 * i},
                                           config=jx.ModelConfig(seed=i)) for i in range(R)]
        calls = []

        def fake_launch(desc, slots, params, seeds, steps, env0=None):
            calls.append((len(seeds), int(seeds[0])))
            out = np.zeros((len(seeds), 3))
            out[:, 0] = params[:, 0] * 2
            return out, 0.5 + rank
        ensemble.ensemble_run = fake_launch
        last, secs = ensemble.run_last_metrics(models, steps=5)
        # host side of a row-band sharded Grid (sharding.DistGroup): handle table + the combines of the
        # ranks' views of the per-agent columns (position max, satisfied min, moves sum)
        from jaxabm_b200 import sharding
        from jaxabm_b200.device import DeviceModel
        g = sharding.DistGroup()
        table = g.all_gather_bytes(np.full(64, rank + 1, dtype=np.uint8))
        assert table.shape == (world, 64) and [int(r[0]) for r in table] == [1, 2]
        pos = np.full((5, 2), -1, dtype=np.int32)
        pos[rank::2] = rank + 10                               # every agent sits in exactly one band
        sat = np.ones(5, dtype=np.bool_)
        sat[rank] = False
        moves = np.arange(5, dtype=np.int32) * (rank + 1)
        assert DeviceModel.GRID_COMBINE == {"position": "max", "satisfied": "min", "moves": "sum"}
        combined = (g.all_reduce(pos, "max").tolist(), g.all_reduce(sat, "min").tolist(), g.all_reduce(moves, "sum").tolist())
        assert g.all_reduce(sat, "min").dtype == np.bool_
        assert dist.shard_bounds(4096, rank, world) == ((0, 2048) if rank == 0 else (2048, 4096))
        dist.barrier()
        q.put((rank, (lo, hi), full[:, 0].tolist(), full[:, 1].tolist(), t, s, calls, last["avg_value"].tolist(), secs,
               combined))
    finally:
        td.destroy_process_group()


def test_two_rank_sharding_and_gather():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, b0, f0, o0, t0, s0, c0, l0, e0, g0), (r1, b1, f1, o1, t1, s1, c1, l1, e1, g1) = out
    assert g0 == g1 == ([[10, 10], [11, 11], [10, 10], [11, 11], [10, 10]], [False, False, True, True, True],
                        [0, 3, 6, 9, 12])
    assert b0 == (0, 6) and b1 == (6, 11)
    assert f0 == f1 == [10.0 * i for i in range(11)]
    assert o0 == [0.0] * 6 + [1.0] * 5
    assert t0 == t1 == 2.0 and s0 == s1 == 3.0
    assert c0 == [(6, 0)] and c1 == [(5, 6)]                     # each rank launched only its block
    want = [float(np.float32(2 * float(np.float32(1.0 + 0.01 * i)))) for i in range(11)]
    assert np.allclose(l0, want) and l0 == l1                    # final gather gives every rank the whole result
    assert e0 == e1 == 1.5                                       # device time = max over ranks
