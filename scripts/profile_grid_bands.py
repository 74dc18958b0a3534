"""Profiling driver for the band kernels on ONE GPU (whole grid = one band, JXB_GRID_BANDS=1):
python scripts/profile_grid_bands.py GRID STEPS  -- runs STEPS steps from the seeded layout."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["JXB_GRID_BANDS"] = "1"
os.environ["JXB_NO_GRAPH"] = "1"          # plain launches: ncu names every kernel
import jaxabm_b200 as jx
from jaxabm_b200.rules import schelling

grid, steps = int(sys.argv[1]), int(sys.argv[2])
n = 13_000_000 if grid == 4096 else int(grid * grid * 0.775)
m = schelling.create_schelling_model(grid, n, seed=42, config=jx.ModelConfig(seed=42))
m._dev.grid_rebuild()
m.run(steps=steps)
print("device us/step", m.last_device_seconds / steps * 1e6)
