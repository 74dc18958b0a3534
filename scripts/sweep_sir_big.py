"""Sweep of the pull kernel's lane-serial / warp-cooperative row-length threshold (JXB_SIR_BIG) at C3."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jaxabm_b200 as jx
from jaxabm_b200 import synthetic
from jaxabm_b200.rules import sir

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
edges = synthetic.scale_free_edges(n, 5, 42)
ref = None
for mode in ("pull_s", "auto"):
    for big in (256, 128, 64, 32, 16):
        os.environ["JXB_SIR_MODE"] = mode
        os.environ["JXB_SIR_BIG"] = str(big)
        m = sir.create_sir_model(n, edges, beta=0.05, gamma=0.1, initial_infected=0.01, seed=42, config=jx.ModelConfig(seed=42))
        m.initialize()
        r = m.run(steps=100)
        sig = (tuple(int(v) for v in r["count_I"]))
        ref = ref or sig
        print(f"mode {mode:7s} big {big:4d}: {m.last_device_seconds / 100 * 1e6:8.1f} us/step  same_result={sig == ref}", flush=True)
        del m
