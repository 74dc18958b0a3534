"""ncu target: the converged ("frozen") phase of the Schelling run -- launch #2 of schelling_run_kernel.
  ncu --set full -k regex:schelling_run --launch-skip 1 --launch-count 1 python scripts/profile_schelling_frozen.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402

g.build()
import jaxabm_b200 as jx  # noqa: E402
from jaxabm_b200.rules import schelling  # noqa: E402

G, N = 4096, 13_000_000
t, p = schelling.initial_layout(G, N, 0.5, 42)
m = schelling.create_schelling_model(G, N, seed=42, types=t, positions=p, config=jx.ModelConfig(seed=42))
m.run(steps=int(os.environ.get("WARM", "200")))
print("warm device s", m.last_device_seconds)
r = m.run(steps=int(os.environ.get("STEPS", "100")))
print("frozen: us/step", m.last_device_seconds / int(os.environ.get("STEPS", "100")) * 1e6, "moves", int(r["total_moves"][-1]) - int(r["total_moves"][0]))
