"""Per-step device time of the economy workload over the first steps (finite regime -> NaN regime)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
wl = bench.EconomyWorkload(0, nh=int(os.environ.get("NH", 49_000_000)), nf=int(os.environ.get("NF", 1_000_000)))
m = wl.fresh()
for t in range(1, 13):
    m.run(steps=1)
    inc = m.agent_collections["households"].states["income"]
    print(t, "device us %.1f" % (m.last_device_seconds * 1e6), "nan frac %.3f" % float(np.isnan(inc).mean()),
          "income min/max", float(np.nanmin(inc)) if not np.isnan(inc).all() else None, float(np.nanmax(inc)) if not np.isnan(inc).all() else None, flush=True)
