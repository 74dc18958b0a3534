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
        dist.barrier()
        q.put((rank, (lo, hi), full[:, 0].tolist(), full[:, 1].tolist(), t, s, calls, last["avg_value"].tolist(), secs))
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
    (r0, b0, f0, o0, t0, s0, c0, l0, e0), (r1, b1, f1, o1, t1, s1, c1, l1, e1) = out
    assert b0 == (0, 6) and b1 == (6, 11)
    assert f0 == f1 == [10.0 * i for i in range(11)]
    assert o0 == [0.0] * 6 + [1.0] * 5
    assert t0 == t1 == 2.0 and s0 == s1 == 3.0
    assert c0 == [(6, 0)] and c1 == [(5, 6)]                     # each rank launched only its block
    want = [float(np.float32(2 * float(np.float32(1.0 + 0.01 * i)))) for i in range(11)]
    assert np.allclose(l0, want) and l0 == l1                    # final gather gives every rank the whole result
    assert e0 == e1 == 1.5                                       # device time = max over ranks
