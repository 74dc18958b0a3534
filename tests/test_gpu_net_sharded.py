"""ONE SIR network split by node ranges (SURVEY.md 8(e) "Network", ``csrc/sir.cuh``): every rank's
``state`` slice and the S/I/R metric rows must equal the single-GPU run bit for bit (per-agent draws
use the global agent index, the counts are exact integers), for every direction the single-GPU
engine may pick, and the CPU oracle.

One process per rank under torchrun (CUDA IPC receive areas, spin waits): all ranks on ONE GPU
(``JXB_DEVICE=0``, gloo) wherever one GPU is visible; one rank per GPU (nccl) when >= 2 are.
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
    one_gpu = os.environ.get("JXB_NET_TEST_ONE_GPU") == "1"
    if one_gpu:
        os.environ["JXB_DEVICE"] = "0"
        td.init_process_group("gloo")
    else:
        local = int(os.environ["LOCAL_RANK"])
        torch.cuda.set_device(local)
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = td.get_rank(), td.get_world_size()
    import jaxabm_b200 as jx
    from jaxabm_b200 import sharding, synthetic
    from jaxabm_b200.rules import sir
    from oracle import rules as orules, runtime as ort

    cases = [(5_003, 3, 30), (200_000, 5, 60)] + ([] if one_gpu else [(2_000_000, 5, 60)])
    for mode in (0, 1):
        for n, deg, steps in cases:
            edges = synthetic.scale_free_edges(n, deg, 11)

            def build():
                return sir.create_sir_model(n, edges, beta=0.08, gamma=0.1, initial_infected=0.01, seed=3,
                                            config=jx.ModelConfig(seed=3, rng_mode=mode))
            ref = build()
            r0 = ref.run(steps=steps)
            r0b = ref.run(steps=7)
            st0 = np.array(ref.agent_collections["agents"].states["state"])
            sh = build()
            sharding.shard_model(sh)
            r1 = sh.run(steps=steps)
            r1b = sh.run(steps=7)                      # _time_step / bitmap parity persist across run() calls
            lo, hi = sh._node_range
            st1 = np.array(sh.agent_collections["agents"].states["state"])
            what = f"rank {rank}/{world} n {n} mode {mode}"
            for k in ("count_S", "count_I", "count_R"):
                assert [int(v) for v in r0[k]] == [int(v) for v in r1[k]], (what, k)
                assert [int(v) for v in r0b[k]] == [int(v) for v in r1b[k]], (what, k, "second run")
            assert st1.shape[0] == hi - lo and np.array_equal(st0[lo:hi], st1), (what, "state slice")
            assert max(int(v) for v in r0["count_I"]) > int(0.02 * n), (what, "the epidemic must take off")
            if n < 10_000:
                om = orules.create_sir_model(n, edges, beta=0.08, gamma=0.1, initial_infected=0.01, seed=3,
                                             config=ort.ModelConfig(seed=3, rng_mode=mode))
                ores = om.run(steps=steps)
                for k in ("count_S", "count_I", "count_R"):
                    assert [int(v) for v in ores[k]] == [int(v) for v in r1[k]], (what, "oracle", k)
            del ref, sh
    td.barrier()
    if rank == 0:
        print(f"sharded network OK on {world} ranks ({'one GPU' if one_gpu else 'one GPU per rank'})")
    td.destroy_process_group()


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _torchrun(world, port, env):
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                          "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)],
                         env=env, capture_output=True, text=True, timeout=900)
    if out.returncode != 0:        # the workers' own tracebacks precede torchrun's summary
        err = out.stderr
        cut = err.find("Traceback")
        raise AssertionError(out.stdout[-1500:] + (err[cut:cut + 4000] if cut >= 0 else err[-4000:]))
    assert "sharded network OK" in out.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
def test_network_node_ranges_ranks_share_one_gpu(world):
    _torchrun(world, 29550 + world, dict(os.environ, JXB_NET_TEST_ONE_GPU="1"))


@pytest.mark.gpu
def test_network_node_ranges_two_gpus():
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    _torchrun(2, 29556, dict(os.environ))


if __name__ == "__main__":
    worker()
