#!/usr/bin/env python
"""bench.py -- throughput of the JaxABM hot path on B200 (contract in the task statement).

  python bench.py --gpus N --steps K --warmup W [--workload schelling|market|walk|sir]
  python bench.py --impl reference ...      # the CPU restatement timed on the host cores

Metric (BASELINE.json): agent-steps/sec, device-timed.  A "step" is one Model.step() of the
workload (jaxabm/model.py:146-216).  Default workload = BASELINE.json configs[1]: Schelling
segregation on a 4096x4096 Grid with 13 M agents.  With N > 1 every rank runs an independent
replica of the workload on its own GPU (ensemble sharding, no data-path collective):
scaling = "weak", value = sum of agent-steps over ranks / max device time over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "agent_steps_per_sec"
UNIT = "agent-steps/s"


# ------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------
class SchellingWorkload:
    """C2: Schelling 4096x4096, 13,000,000 agents (77.5 % fill), threshold 0.5, Moore-8."""
    name = "schelling_4096x4096_13M"
    dtype = "i8/i32"

    def __init__(self, rank, grid=4096, n=13_000_000):
        from jaxabm_b200.rules import schelling
        self.grid, self.n, self.seed = grid, n, 42 + rank
        self.types, self.positions = schelling.initial_layout(grid, n, 0.5, self.seed)
        self.agents = n

    def make(self):
        import jaxabm_b200 as jx
        from jaxabm_b200.rules import schelling
        return schelling.create_schelling_model(self.grid, self.n, seed=self.seed, types=self.types,
                                                positions=self.positions,
                                                config=jx.ModelConfig(seed=self.seed))

    def pinned_inputs(self):
        import torch
        t = torch.from_numpy(self.types).pin_memory()
        p = torch.from_numpy(self.positions).pin_memory()
        outs = {"position": torch.empty((self.n, 2), dtype=torch.int32).pin_memory(),
                "moves": torch.empty(self.n, dtype=torch.int32).pin_memory(),
                "satisfied": torch.empty(self.n, dtype=torch.bool).pin_memory()}
        return {"type": t, "position": p}, outs

    def e2e_run(self, model, ins, outs, steps):
        """Public-API call with host buffers: upload state -> run -> read results back."""
        st = model.agent_collections["agents"].states
        st["type"] = ins["type"].numpy()
        st["position"] = ins["position"].numpy()
        res = model.run(steps=steps)
        dev = model._dev
        for k, buf in outs.items():
            dev.download(0, dev.field_index(0, k), out=buf.numpy())
        h2d = ins["type"].numel() * 4 + ins["position"].numel() * 4
        d2h = sum(b.numel() * b.element_size() for b in outs.values()) + steps * (3 * 8 + 4)
        return res, h2d, d2h

    def kernel_bytes(self, res):
        """Algorithmic bytes of ONE launch of the dominant kernel (stencil_compact_kernel) in the
        engine's layout: packed grid read once (1 B/cell) + ordered lists written
        (U, UA: 4 B each + 4 B cell_agent read per unsatisfied agent; E: 4 B per empty cell)."""
        cells = self.grid * self.grid
        ps = np.array([float(v) for v in res["percent_satisfied"]])
        u = float(np.mean((1.0 - ps) * self.n))
        e = cells - self.n
        return cells * 1 + u * 12 + e * 4

    def api_bytes_per_step(self):
        """SURVEY.md 8(d) figure in the reference's API-visible dtypes: 29 B/agent + 8 B/cell."""
        return 29 * self.n + 8 * self.grid * self.grid

    def cpu_run(self, steps):
        from oracle import cfast
        f = cfast.SchellingFast(self.grid, self.types, self.positions, seed=self.seed, mode=1)
        t0 = time.perf_counter()
        f.run(steps)
        return time.perf_counter() - t0, cfast.num_threads()


class MarketWorkload:
    """C4-A: 45 M consumers + 5 M producers, well-mixed, env-level reductions fused in the step."""
    name = "market_45M_consumers_5M_producers"
    dtype = "f32"

    def __init__(self, rank, nc=45_000_000, npr=5_000_000):
        self.nc, self.npr, self.seed = nc, npr, 42 + rank
        self.agents = nc + npr

    def make(self):
        import jaxabm_b200 as jx
        from jaxabm_b200.rules import market
        m = market.create_economy_model(num_consumers=self.nc, num_producers=self.npr,
                                        config=jx.ModelConfig(seed=self.seed))
        m.initialize()
        return m

    def pinned_inputs(self):
        return None, None

    def kernel_bytes(self, res):
        # consumers: read savings+income, write savings+consumption+utility = 20 B; producers:
        # read capital, write capital+production+profit = 16 B (in-place; 'income' is not rewritten)
        return 20 * self.nc + 16 * self.npr

    def api_bytes_per_step(self):
        return 32 * self.nc + 24 * self.npr

    def cpu_run(self, steps):
        raise NotImplementedError


class WalkWorkload:
    """C1 scaled: 2^26 random walkers (48 B/agent-step), fused distance reductions."""
    name = "random_walk_2^26"
    dtype = "f32"

    def __init__(self, rank, n=1 << 26):
        self.n, self.seed = n, 42 + rank
        self.agents = n

    def make(self):
        import jaxabm_b200 as jx
        from jaxabm_b200.rules import random_walk
        m = random_walk.create_scaled_walk_model(self.n, config=jx.ModelConfig(seed=self.seed))
        m.initialize()
        return m

    def pinned_inputs(self):
        return None, None

    def kernel_bytes(self, res):
        return 48 * self.n

    def api_bytes_per_step(self):
        return 48 * self.n

    def cpu_run(self, steps):
        raise NotImplementedError


WORKLOADS = {"schelling": SchellingWorkload, "market": MarketWorkload, "walk": WalkWorkload}


# ------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def load_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def load_traffic(workload):
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(workload)
        except Exception:
            pass
    return None


def reference_arm(args, rank, world):
    """CPU restatement of the reference's path (oracle/c, OpenMP over the host cores) on the same
    workload; under torchrun only rank 0 works."""
    if rank != 0:
        return
    wl = WORKLOADS[args.workload](0)
    for _ in range(min(args.warmup, 1)):
        wl.cpu_run(1)
    secs, threads = wl.cpu_run(args.steps)
    value = wl.agents * args.steps / secs
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": wl.dtype,
            "data": "synthetic", "config": {"workload": wl.name},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{args.steps} full-size steps of {wl.name} on the C/OpenMP oracle "
                                       "(JAX is not installable here, so the reference itself cannot run)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="schelling", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-steps", type=int, default=40, help="steps of the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return reference_arm(args, rank, world)

    import torch
    import torch.distributed as td
    torch.cuda.set_device(local)
    if world > 1:
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as g
    if rank == 0:
        g.build()
    if world > 1:
        td.barrier()
    import jaxabm_b200 as jx
    from jaxabm_b200 import _native as nat

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
            torch.cuda.synchronize()

    wl = WORKLOADS[args.workload](rank)
    eng = nat.engine()
    model = wl.make()
    model.run(steps=args.warmup)                       # untimed warm-up (also builds the CUDA graphs)

    # ---- timed region: exactly K steps, device-timed on the engine's stream ------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    sync_all()
    l0 = eng.launch_count
    res = model.run(steps=args.steps)
    dev_s = model.last_device_seconds
    sync_all()
    launches = eng.launch_count - l0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([dev_s], dtype=torch.float64, device="cuda")
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX)
    max_s = float(t.item())
    total_agents = wl.agents * world
    value = total_agents * args.steps / max_s

    # ---- dominant kernel: CUDA events around every launch of it over K more steps ---------------
    model._dev.set_profile(True)
    res_p = model.run(steps=args.steps)
    ksecs, klaunches, kname = model._dev.profile()
    model._dev.set_profile(False)
    peak, peak_kind = load_peak()
    kbytes = wl.kernel_bytes(res_p)
    achieved = kbytes / (ksecs / max(klaunches, 1)) / 1e9
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_kind": f"of {peak_kind}", "traffic": load_traffic(args.workload),
                "bytes_per_launch": kbytes, "us_per_launch": ksecs / max(klaunches, 1) * 1e6,
                "kernel_share_of_step": (ksecs / max(klaunches, 1)) / (model.last_device_seconds / args.steps),
                "api_layout_gbs": wl.api_bytes_per_step() / (max_s / args.steps) / 1e9}

    # ---- end to end through the public API with pinned host buffers --------------------------------
    e2e = None
    ins, outs = wl.pinned_inputs()
    if ins is not None:
        m2 = wl.make()
        wl.e2e_run(m2, ins, outs, 3)                    # warm-up of the same call
        sync_all()
        t0 = time.perf_counter()
        _, h2d, d2h = wl.e2e_run(m2, ins, outs, args.steps)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        tt = torch.tensor([wall], dtype=torch.float64, device="cuda")
        if world > 1:
            td.all_reduce(tt, op=td.ReduceOp.MAX)
        e2e = {"value": total_agents * args.steps / float(tt.item()), "unit": UNIT,
               "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
               "what": "upload type+position from pinned host memory, Model.run(K), read back "
                       "position/moves/satisfied + the metrics history"}
    if rank != 0:
        return
    cpu = None
    if world == 1 and not args.no_cpu:
        try:
            secs, threads = wl.cpu_run(args.cpu_steps)
            cpu = {"value": wl.agents * args.cpu_steps / secs, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{args.cpu_steps} full-size steps of {wl.name} on the C/OpenMP oracle ({secs:.1f} s)"}
        except NotImplementedError:
            cpu = None
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": max_s / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
            "config": {"workload": wl.name, "agents_per_gpu": wl.agents,
                       "parallelism": f"replica-per-gpu x{world}" if world > 1 else "single-gpu",
                       "l2": "no flush: resident state (agents SoA + cell arrays + lists, >400 MB) exceeds the 126 MB L2"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
