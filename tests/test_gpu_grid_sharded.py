"""ONE Schelling grid split into row bands (SURVEY.md 8(e) "Grid: row blocks + halo",
``csrc/grid_shard.cuh``): every sharded run must reproduce the single-GPU run -- and the CPU oracle
-- bit for bit (types, positions, satisfied, moves, env grid, empty-cell slots, metric rows).

One process per rank under torchrun, the product path (CUDA IPC receive areas, spin waits):
  * all ranks on ONE GPU (``JXB_DEVICE=0``, gloo process group) -- IPC and the waits work across
    processes that time-share a device, so the whole path is covered wherever one GPU is visible;
  * one rank per GPU (nccl process group, NVLink peer memory) when >= 2 GPUs are visible:
      python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
          tests/test_gpu_grid_sharded.py
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

FIELDS = ("type", "position", "satisfied", "moves")


def _snapshot(model, res):
    st = model.agent_collections["agents"].states
    return {"res": {k: [float(v) for v in vals] for k, vals in res.items()},
            "state": {k: np.array(st[k]) for k in FIELDS},
            "grid": np.array(model._dev.download_grid()),
            "empty": np.array(model._dev.download_empty_cells())}


def _assert_same(a, b, what):
    assert a["res"] == b["res"], (what, "metric rows differ")
    for k in FIELDS:
        assert np.array_equal(a["state"][k], b["state"][k]), (what, k)
    assert np.array_equal(a["grid"], b["grid"]), (what, "env grid")
    assert np.array_equal(a["empty"], b["empty"]), (what, "empty_cells")


def worker():
    import torch
    import torch.distributed as td
    one_gpu = os.environ.get("JXB_GRID_TEST_ONE_GPU") == "1"
    if one_gpu:
        os.environ["JXB_DEVICE"] = "0"
        td.init_process_group("gloo")
    else:
        local = int(os.environ["LOCAL_RANK"])
        torch.cuda.set_device(local)
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = td.get_rank(), td.get_world_size()
    import jaxabm_b200 as jx
    from jaxabm_b200.rules import schelling
    from oracle import rules as orules, runtime as ort

    cases = [(64, 3300, False, 0.5, 8), (96, 7000, True, 0.6, 12), (96, 7000, False, 0.6, 12), (1024, 800_000, False, 0.5, 30)]
    if not one_gpu:
        cases += [(2048, 3_200_000, True, 0.5, 40), (4096, 13_000_000, False, 0.5, 60)]
    for mode in (0, 1):
        for grid, n, periodic, thr, steps in cases:
            if grid < world:
                continue

            def build(**kw):
                return schelling.create_schelling_model(grid, n, seed=5, periodic=periodic, similarity_threshold=thr,
                                                        config=jx.ModelConfig(seed=9, rng_mode=mode), **kw)
            ref = build()
            a1 = _snapshot(ref, ref.run(steps=steps))
            a2 = _snapshot(ref, ref.run(steps=5))       # _time_step and state persist across run() calls
            small = grid <= 96
            if small:
                # a position upload between runs rebuilds the cell binning (slot order restarts): on the bands every
                # rank must rebuild from WHOLE columns although its reads only ever returned the band's view
                ref.agent_collections["agents"].states["position"] = a2["state"]["position"]
                a3 = _snapshot(ref, ref.run(steps=4))
            del ref
            sh = build(shard=True)
            s1 = _snapshot(sh, sh.run(steps=steps))
            s2 = _snapshot(sh, sh.run(steps=5))
            what = f"rank {rank}/{world} grid {grid} periodic {periodic} mode {mode}"
            _assert_same(a1, s1, what + " first run")
            _assert_same(a2, s2, what + " second run")
            if small:
                sh.agent_collections["agents"].states["position"] = s2["state"]["position"]
                s3 = _snapshot(sh, sh.run(steps=4))
                _assert_same(a3, s3, what + " after a position upload")
                assert s3["state"]["moves"].sum() > s2["state"]["moves"].sum()
            assert a1["res"]["total_moves"][-1] > 0
            del sh
            if grid == 64:
                # the sharded path against the CPU restatement directly, not only via the single-GPU kernel
                om = orules.create_schelling_model(grid, n, seed=5, periodic=periodic, similarity_threshold=thr,
                                                   config=ort.ModelConfig(seed=9, rng_mode=mode))
                ores = om.run(steps=steps)
                ost = om.agent_collections["agents"].states
                for k in FIELDS:
                    assert np.array_equal(s1["state"][k], np.asarray(ost[k])), ("oracle", k)
                assert [int(v) for v in s1["res"]["total_moves"]] == [int(v) for v in ores["total_moves"]]
                # float32 mean of same/occupied: the oracle sums floats, the engine exact integers (rtol as in
                # tests/test_gpu_parity.py)
                np.testing.assert_allclose(s1["res"]["segregation_index"],
                                           [float(v) for v in ores["segregation_index"]], rtol=1e-6)
                np.testing.assert_allclose(s1["res"]["percent_satisfied"],
                                           [float(v) for v in ores["percent_satisfied"]], rtol=1e-6)
    td.barrier()
    if rank == 0:
        print(f"sharded grid OK on {world} ranks ({'one GPU' if one_gpu else 'one GPU per rank'})")
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
    assert "sharded grid OK" in out.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
def test_grid_row_bands_ranks_share_one_gpu(world):
    _torchrun(world, 29540 + world, dict(os.environ, JXB_GRID_TEST_ONE_GPU="1"))


@pytest.mark.gpu
def test_grid_row_bands_two_gpus():
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    _torchrun(2, 29546, dict(os.environ))


@pytest.mark.gpu
@pytest.mark.parametrize("grid,n,periodic,steps", [(64, 3300, False, 10), (96, 7000, True, 12), (1024, 800_000, False, 30),
                                                   (4128, 13_000_000, False, 25)])
def test_single_gpu_band_mode_matches_persistent_kernels(grid, n, periodic, steps, mode, monkeypatch):
    """On ONE GPU the band kernels (whole grid = one band, 4 launches per step) serve the large grids outside the
    persistent bit-sliced kernel's shapes; they must reproduce the persistent kernels bit for bit.  4128 x 4128
    (>= 2^24 cells, rows not a multiple of 1024) selects them by default."""
    import jaxabm_b200 as jx
    from jaxabm_b200.rules import schelling

    def run(bands):
        monkeypatch.setenv("JXB_GRID_BANDS", bands)
        m = schelling.create_schelling_model(grid, n, seed=5, periodic=periodic, similarity_threshold=0.6,
                                             config=jx.ModelConfig(seed=9, rng_mode=mode))
        snap = _snapshot(m, m.run(steps=steps)), _snapshot(m, m.run(steps=4)), m._dev.profile()[2]
        del m
        return snap
    a1, a2, ka = run("0")
    b1, b2, kb = run("1")
    assert ka in ("schelling_bits_kernel", "schelling_run_kernel") and kb == "grid_shard_sweep_kernel"
    _assert_same(a1, b1, "first run")
    _assert_same(a2, b2, "second run")
    if grid == 4128:
        monkeypatch.delenv("JXB_GRID_BANDS")
        m = schelling.create_schelling_model(grid, n, seed=5, periodic=periodic, similarity_threshold=0.6,
                                             config=jx.ModelConfig(seed=9, rng_mode=mode))
        assert m._dev.profile()[2] == "grid_shard_sweep_kernel"       # the default for this shape


@pytest.mark.gpu
def test_grid_shard_rejects_bad_shapes():
    from jaxabm_b200 import _native as nat
    from jaxabm_b200.device import DeviceModel, TypeSpec, make_desc
    d = make_desc("schelling", [TypeSpec("schelling", 100)], [0.5], grid=(40, 40, False), world_size=2, rank=0)
    with pytest.raises(nat.JxbError, match="multiple of 32"):
        DeviceModel(d)
    d = make_desc("schelling", [TypeSpec("schelling", 100)], [0.5], grid=(64, 64, False), world_size=2, rank=0)
    dev = DeviceModel(d)
    h = np.zeros(80, dtype=np.uint8)
    with pytest.raises(nat.JxbError, match="consecutive rows"):
        nat.check(nat.lib().jxb_model_grid_shard_export(dev.handle, 0, 40, nat.ptr(h), h.nbytes))
    with pytest.raises(nat.JxbError, match="export"):
        nat.check(nat.lib().jxb_model_grid_rebuild(dev.handle))


if __name__ == "__main__":
    worker()
