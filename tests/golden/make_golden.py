"""Regenerates the oracle-made fixtures in this directory (``python tests/golden/make_golden.py``).

JAX cannot be installed in this image, so these are outputs of the CPU oracle (``oracle/``),
NOT of the reference itself: they pin the oracle against silent drift and give the CUDA path a
second, frozen target.  ``threefry_kat.json`` is different: it holds published third-party
known-answer vectors and is hand-written.
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import economy as oeco, jaxlike as jl, rules as orules, runtime as ort  # noqa: E402
from jaxabm_b200.synthetic import ring_lattice_edges  # noqa: E402  (host-side generator, no device needed)


def series(d):
    return {k: np.array([float(x) for x in v], dtype=np.float64) for k, v in d.items()}


def main():
    out = {}
    for mode in (0, 1):
        tag = "legacy" if mode == 0 else "part"
        m = orules.RandomWalkModelNamed({"n_agents": 1000, "steps": 100, "seed": 42}, rng_mode=mode)
        for k, v in series(m.run()).items():
            out[f"walk_{tag}_{k}"] = v
        mk = orules.create_economy_model(num_consumers=2000, num_producers=500,
                                         config=ort.ModelConfig(seed=42, rng_mode=mode))
        for k, v in series(mk.run(steps=40)).items():
            out[f"market_{tag}_{k}"] = v
        out[f"market_{tag}_income"] = mk.agent_collections["consumers"].states["income"]
        sc = orules.create_schelling_model(48, 1800, seed=5, config=ort.ModelConfig(seed=5, rng_mode=mode))
        for k, v in series(sc.run(steps=15)).items():
            out[f"schelling_{tag}_{k}"] = v
        st = sc.agent_collections["agents"].states
        out[f"schelling_{tag}_position"] = st["position"]
        out[f"schelling_{tag}_moves"] = st["moves"]
        out[f"schelling_{tag}_satisfied"] = st["satisfied"]
        edges = ring_lattice_edges(3000, 2)
        sr = orules.create_sir_model(3000, edges, beta=0.3, gamma=0.1, initial_infected=0.02, seed=9,
                                     config=ort.ModelConfig(seed=9, rng_mode=mode))
        for k, v in series(sr.run(steps=25)).items():
            out[f"sir_{tag}_{k}"] = v
        out[f"sir_{tag}_state"] = sr.agent_collections["agents"].states["state"]
        g = orules.create_test_model(growth_rate=0.07, adjustment_rate=0.13, initial_value=2.5, num_agents=64,
                                     config=ort.ModelConfig(seed=0, rng_mode=mode))
        for k, v in series(g.run(steps=30)).items():
            out[f"growth_{tag}_{k}"] = v
        ec = oeco.create_economy_model(1200, 30, config=ort.ModelConfig(seed=42, rng_mode=mode))
        ec.initialize()
        out[f"economy_{tag}_init_income"] = ec.agent_collections["households"].states["income"].copy()
        out[f"economy_{tag}_init_ptc"] = ec.agent_collections["households"].states["propensity_to_consume"].copy()
        out[f"economy_{tag}_init_capital"] = ec.agent_collections["consumer_firms"].states["capital_stock"].copy()
        for k, v in series(ec.run(steps=8)).items():
            out[f"economy_{tag}_{k}"] = v
        out[f"economy_{tag}_employed"] = ec.agent_collections["households"].states["employed"]
        key = jl.PRNGKey(42)
        out[f"prng_{tag}_split5"] = jl.split(key, 5, mode)
        out[f"prng_{tag}_uniform9"] = jl.uniform(key, (9,), -1.0, 3.0, mode)
        out[f"prng_{tag}_randint11"] = jl.randint(key, (11,), 0, 1000000, mode)
        out[f"prng_{tag}_normal7"] = jl.normal(key, (7,), mode)
        out[f"prng_{tag}_perm50"] = jl.permutation(key, np.arange(50), mode)
    try:
        rev = subprocess.check_output(["git", "rev-parse", "--short", "HEAD"], cwd=HERE).decode().strip()
    except Exception:
        rev = "unknown"
    out["_oracle_git_rev"] = np.array(rev)
    np.savez_compressed(os.path.join(HERE, "oracle_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
