"""Where the end-to-end time of the default bench workload goes (wall clock, host side)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
g.build()
import jaxabm_b200 as jx
from jaxabm_b200.rules import schelling
from jaxabm_b200.agent import AgentCollection
from jaxabm_b200.model import Model

G, N, K = 4096, 13_000_000, 1000
types, pos = schelling.initial_layout(G, N, 0.5, 42)
tp, pp = torch.from_numpy(types).pin_memory(), torch.from_numpy(pos).pin_memory()
outs = {"position": torch.empty((N, 2), dtype=torch.int32).pin_memory(), "moves": torch.empty(N, dtype=torch.int32).pin_memory(),
        "satisfied": torch.empty(N, dtype=torch.bool).pin_memory()}
for it in range(3):
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    coll = AgentCollection(schelling.SchellingAgentType(0.5), N)
    m = Model(params={"similarity_threshold": 0.5}, config=jx.ModelConfig(seed=42), update_state_fn=schelling.schelling_update_state,
              metrics_fn=schelling.schelling_metrics)
    m.add_agent_collection("agents", coll)
    m.add_env_state("grid_shape", (G, G)); m.add_env_state("grid_periodic", False)
    m.add_env_state("segregation_index", 0.0); m.add_env_state("percent_satisfied", 0.0); m.add_env_state("total_moves", 0)
    m.initialize(); t.append(time.perf_counter())
    coll.states["type"] = tp.numpy(); t.append(time.perf_counter())
    coll.states["position"] = pp.numpy(); t.append(time.perf_counter())
    m._dev.grid_rebuild(); t.append(time.perf_counter())
    res = m.run(steps=K); t.append(time.perf_counter())
    dev = m._dev
    for k, buf in outs.items():
        dev.download(0, dev.field_index(0, k), out=buf.numpy())
        t.append(time.perf_counter())
    names = ["create+initialize", "upload type", "upload position", "grid_rebuild", "run (device %.2f ms)" % (m.last_device_seconds * 1e3),
             "dl position", "dl moves", "dl satisfied"]
    print(it, " | ".join("%s %.2f" % (n, (b - a) * 1e3) for n, a, b in zip(names, t[:-1], t[1:])), "| total %.2f ms" % ((t[-1] - t[0]) * 1e3))
    del m, coll, dev
