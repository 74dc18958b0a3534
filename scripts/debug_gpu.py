import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import jaxabm_b200 as jx
from jaxabm_b200.rules import random_walk, market
m = random_walk.RandomWalkModel({"n_agents": 1000, "steps": 3, "seed": 42, "name": "walkers"})
res = m.run()
print({k: v for k, v in res._data.items()})
d = m._jax_model._dev
print("env", [d.get_env(i) for i in range(len(d.env_slots))])
print("metric slots", d.metric_slots)
st, rec, secs = d.run(2, 1)
print(st, rec, secs)
mm = market.create_economy_model(num_consumers=20, num_producers=5, config=jx.ModelConfig(seed=42))
print(mm.run(steps=3))
