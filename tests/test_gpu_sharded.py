"""Population sharding on real GPUs (one process per GPU, torchrun): the sharded market model
must reproduce the single-GPU model -- per-agent state bit for bit (keys use the global agent
index), env/metric trajectories to float32 rounding (partials are folded in a different order).

Run directly under torchrun on a box with >= 2 GPUs (``gpurun --gpus 2``):
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
      tests/test_gpu_sharded.py
Also collected by pytest (-m gpu): the test spawns the 2-rank job itself when >= 2 GPUs are visible.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker():
    import torch
    import torch.distributed as td
    one_gpu = os.environ.get("JXB_SHARD_TEST_ONE_GPU") == "1"
    if one_gpu:
        # all ranks time-share GPU 0: CUDA IPC peer mappings and the in-kernel spin waits work across processes on
        # one device, so the whole exchange path is covered wherever a single GPU is visible
        os.environ["JXB_DEVICE"] = "0"
        td.init_process_group("gloo")
    else:
        local = int(os.environ["LOCAL_RANK"])
        torch.cuda.set_device(local)
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = td.get_rank(), td.get_world_size()
    import jaxabm_b200 as jx
    from jaxabm_b200 import sharding, dist
    from jaxabm_b200.rules import market

    nc, npr, steps = 200_003, 50_001, 25
    for mode in (0, 1):
        ref = market.create_economy_model(num_consumers=nc, num_producers=npr, config=jx.ModelConfig(seed=11, rng_mode=mode))
        r0 = ref.run(steps=steps)
        sh = market.create_economy_model(num_consumers=nc, num_producers=npr, config=jx.ModelConfig(seed=11, rng_mode=mode))
        sharding.shard_model(sh)
        r1 = sh.run(steps=steps)
        for k in ("gdp", "price_level", "unemployment", "avg_utility", "avg_profit"):
            a, b = np.array(r0[k], dtype=np.float64), np.array(r1[k], dtype=np.float64)
            assert np.allclose(a, b, rtol=2e-6, atol=1e-9), (mode, k, a[-1], b[-1])
        # all ranks hold identical env trajectories (folded in rank order from identical rows)
        mine = np.array([float(v) for v in r1["price_level"]])
        t = torch.from_numpy(mine)
        t = t if one_gpu else t.cuda()
        parts = [torch.zeros_like(t) for _ in range(world)]
        td.all_gather(parts, t)
        for p in parts:
            assert torch.equal(p, parts[0]), "ranks disagree on the env trajectory"
        # per-agent init is bit-identical to the unsharded model's index range (global keys)
        lo, hi = dist.shard_bounds(nc, rank, world)
        inc_ref = ref.agent_collections["consumers"].states["income"][lo:hi]
        inc_sh = sh.agent_collections["consumers"].states["income"]
        assert inc_sh.shape[0] == hi - lo and np.array_equal(inc_ref, inc_sh), "sharded init differs"
        lo, hi = dist.shard_bounds(npr, rank, world)
        cap_ref = ref.agent_collections["producers"].states["capital"][lo:hi]
        cap_sh = sh.agent_collections["producers"].states["capital"]
        assert np.allclose(cap_ref, cap_sh, rtol=1e-4), "capital trajectories diverged"
    # ---- C4-B economy: 15 env partial sums through the in-kernel exchange + NCCL all-reduce of the
    # Gini histogram; per-agent draws use global indices, so booleans are exact and floats agree to
    # the fold-order rounding
    if os.environ.get("JXB_EXCHANGE", "p2p") != "nccl":
        from jaxabm_b200.rules import economy
        nh, nf = 60_001, 1_503
        for mode in (0, 1):
            ref = economy.create_economy_model(nh, nf, config=jx.ModelConfig(seed=5, rng_mode=mode))
            r0 = ref.run(steps=3)
            sh = economy.create_economy_model(nh, nf, config=jx.ModelConfig(seed=5, rng_mode=mode))
            sharding.shard_model(sh)
            r1 = sh.run(steps=3)
            for k in ("gdp", "wage_rate", "interest_rate", "unemployment", "inequality"):
                a, b = np.array(r0[k], dtype=np.float64), np.array(r1[k], dtype=np.float64)
                assert np.allclose(a, b, rtol=1e-5, atol=1e-6), ("economy", mode, k, a, b)
            lo, hi = dist.shard_bounds(nh, rank, world)
            assert np.array_equal(ref.agent_collections["households"].states["employed"][lo:hi],
                                  sh.agent_collections["households"].states["employed"])
            assert np.allclose(ref.agent_collections["households"].states["cash"][lo:hi],
                               sh.agent_collections["households"].states["cash"], rtol=1e-4)
    td.barrier()
    if rank == 0:
        print(f"sharded market OK on {world} {'ranks sharing one GPU' if one_gpu else 'GPUs'} "
              f"(exchange={os.environ.get('JXB_EXCHANGE', 'p2p')})")
    td.destroy_process_group()


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_sharded_market_two_gpus(exchange):
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    env = dict(os.environ, JXB_EXCHANGE=exchange)
    port = 29533 if exchange == "p2p" else 29534
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)],
                         env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "sharded market OK" in out.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_market_and_economy_ranks_share_one_gpu(world):
    """The product exchange path (in-kernel partial-sum exchange of the market and the economy, peer-memory
    all-reduce of the Gini histogram) with every rank on GPU 0 -- runs wherever one GPU is visible."""
    env = dict(os.environ, JXB_SHARD_TEST_ONE_GPU="1")
    env.pop("JXB_EXCHANGE", None)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                          "--master-addr", "127.0.0.1", "--master-port", str(29535 + world), os.path.abspath(__file__)],
                         env=env, capture_output=True, text=True, timeout=900)
    if out.returncode != 0:
        err = out.stderr
        cut = err.find("Traceback")
        raise AssertionError(out.stdout[-1500:] + (err[cut:cut + 4000] if cut >= 0 else err[-4000:]))
    assert "sharded market OK" in out.stdout


if __name__ == "__main__":
    worker()
