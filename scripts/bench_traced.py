"""Traced (user-Python) market vs the registered hand-written kernel at C4-A size."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
g.build()
import jaxabm_b200 as jx
from jaxabm_b200.rules import market
from traced_models import build

nc, npr, K = 45_000_000, 5_000_000, 50
for name, mk in (("registered", lambda: market.create_economy_model(num_consumers=nc, num_producers=npr, config=jx.ModelConfig(seed=42))),
                 ("traced", lambda: build.device_market(nc, npr, 42, 1))):
    m = mk()
    m.run(steps=5)
    r = m.run(steps=K)
    print(name, "us/step %.1f" % (m.last_device_seconds / K * 1e6), "agent-steps/s %.3g" % ((nc + npr) * K / m.last_device_seconds),
          "gdp", float(r["gdp"][-1]))
