"""Small shapes of every kernel family with cross-thread / cross-CTA writes, for compute-sanitizer
(memcheck / racecheck / synccheck): run as `compute-sanitizer --tool <tool> python scripts/sanitize_small.py`.
Each case is also checked against the CPU oracle, so a sanitizer-clean run is a correct run."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import jaxabm_b200 as jx
from jaxabm_b200 import ensemble, synthetic
from jaxabm_b200.rules import growth, market, schelling, sir
from oracle import rules as orules, runtime as ort

which = sys.argv[1:] or ["schelling_bits", "schelling_byte", "schelling_bands", "sir", "ensemble", "market"]
mode = 1

if "schelling_bits" in which or "schelling_byte" in which or "schelling_bands" in which:
    for name, g, n in (("schelling_bits", 1024, 800_000), ("schelling_byte", 96, 7000), ("schelling_bands", 96, 7000)):
        if name not in which:
            continue
        if name == "schelling_bands":
            os.environ["JXB_GRID_BANDS"] = "1"
        m = schelling.create_schelling_model(g, n, seed=5, config=jx.ModelConfig(seed=9, rng_mode=mode))
        r = m.run(steps=6)
        print(name, m._dev.profile()[2], "moves", int(r["total_moves"][-1]))
        if g <= 128:
            om = orules.create_schelling_model(g, n, seed=5, config=ort.ModelConfig(seed=9, rng_mode=mode))
            ores = om.run(steps=6)
            for k in ("type", "position", "satisfied", "moves"):
                assert np.array_equal(m.agent_collections["agents"].states[k], om.agent_collections["agents"].states[k]), k
        os.environ.pop("JXB_GRID_BANDS", None)
        del m

if "sir" in which:
    n = 60_000
    edges = synthetic.scale_free_edges(n, 5, 42)
    for sm in ("auto", "pull", "push", "pull_s"):
        os.environ["JXB_SIR_MODE"] = sm
        m = sir.create_sir_model(n, edges, beta=0.2, gamma=0.1, initial_infected=0.02, seed=3, config=jx.ModelConfig(seed=3, rng_mode=mode))
        r = m.run(steps=10)
        om = orules.create_sir_model(n, edges, beta=0.2, gamma=0.1, initial_infected=0.02, seed=3, config=ort.ModelConfig(seed=3, rng_mode=mode))
        orr = om.run(steps=10)
        assert [int(v) for v in r["count_I"]] == [int(v) for v in orr["count_I"]], sm
        print("sir", sm, "I", int(r["count_I"][-1]))
        del m
    os.environ.pop("JXB_SIR_MODE", None)

if "ensemble" in which:
    for shape, env in (("smem", {}), ("cluster2", {"JXB_ENS_MIN_CLUSTER": "2"}), ("cluster8", {"JXB_ENS_MIN_CLUSTER": "8"}),
                       ("scratch", {"JXB_ENS_FORCE_SCRATCH": "1"})):
        os.environ.update(env)
        models = [growth.create_test_model(params={"growth_rate": 0.05 + 0.01 * i, "adjustment_rate": 0.1}, initial_value=1.0,
                                           config=jx.ModelConfig(seed=i + 1000, steps=20, rng_mode=mode), num_agents=20_000)
                  for i in range(4)]
        last, _ = ensemble.run_last_metrics(models, steps=20)
        print("ensemble", shape, [float(v) for v in last["avg_value"]])
        for k in env:
            os.environ.pop(k)

if "market" in which:
    mm = market.create_economy_model(num_consumers=20_000, num_producers=5_000, config=jx.ModelConfig(seed=3, rng_mode=mode))
    r = mm.run(steps=5)
    print("market gdp", float(r["gdp"][-1]))
print("sanitize_small OK")
