"""ncu / sanitizer target: ONE fresh model of a bench workload stepped `steps` times (no warm-up model, no
CPU baseline), so `ncu -k regex:<kernel> -s <skip> -c <count>` picks launches by step number.
  python scripts/prof_target.py <schelling|market|economy|walk|sir|ensemble> [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

g.build()
import bench  # noqa: E402

name = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
if name == "ensemble":
    from jaxabm_b200.device import ensemble_run
    wl = bench.EnsembleWorkload(0, 1, samples=int(os.environ.get("SAMPLES", "1184")))
    desc, slots, params, seeds, env0 = wl.plan(steps)
    vals, secs = ensemble_run(desc, slots, params, seeds, steps, env0)
    print("ensemble", wl.samples_total, "replicas x", steps, "steps:", secs, "s")
else:
    wl = bench.WORKLOADS[name](0)
    m = wl.fresh()
    m.run(steps=steps)
    print(name, steps, "steps:", m.last_device_seconds, "s ->", m.last_device_seconds / steps * 1e6, "us/step")
