#!/usr/bin/env python
"""bench.py -- throughput of the JaxABM hot path on B200 (contract in the task statement).

  python bench.py --gpus N --steps K --warmup W [--workload schelling|market|economy|walk|sir|ensemble]
  python bench.py --impl reference ...      # the CPU restatement timed on the host cores

Metric (BASELINE.json): agent-steps/sec, device-timed.  A "step" is one Model.step() of the
workload (jaxabm/model.py:146-216).  Default workload = BASELINE.json configs[1]: Schelling
segregation on a 4096x4096 Grid with 13 M agents; its timed region is K steps of the run that
STARTS FROM THE SEEDED INITIAL LAYOUT (so it covers the active phase in which millions of agents
move as well as the converged tail), default K = 1000 = the configured run length.  With N > 1
every rank runs an independent replica of the workload on its own GPU (ensemble sharding, no
data-path collective): scaling = "weak", value = sum of agent-steps over ranks / max device time
over ranks.  `--workload market --shard` instead splits ONE population across the ranks (strong
scaling, per-step cross-rank reduction of the env partial sums); `--workload schelling --shard
[--grid G --agents N]` splits ONE grid into row bands over the ranks (strong scaling, the
unsatisfied agents' records exchanged over NVLink peer memory every step); `--workload sir --shard`
splits ONE network by node ranges (the infected bitmap's new words stored into every rank's copy).
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


def _pin(a):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t.pin_memory() if torch.cuda.is_available() else t      # the reference arm also runs on a box without a GPU


def _pinned_like(shape, dtype):
    import torch
    return torch.empty(shape, dtype=dtype).pin_memory()


# ------------------------------------------------------------------------------------------
# workloads.  Protocol:
#   fresh()            -> model whose state is resident in HBM, ready for step 1 (untimed)
#   stationary         -> True: per-step cost does not depend on t (warm up and time one model);
#                         False: the timed run starts from a fresh model
#   api_bytes(res,K)   -> SURVEY.md 8(d) algorithmic bytes of K steps (reference API dtypes)
#   engine_bytes(res,K)-> same in the engine's packed layout (what the kernels must move)
#   e2e(K)             -> (wall_s, h2d_bytes, d2h_bytes) through the public API with host buffers
#   cpu_run(steps)     -> (seconds, threads, what) on the CPU oracle
# ------------------------------------------------------------------------------------------
class SchellingWorkload:
    """C2: Schelling 4096x4096, 13,000,000 agents (77.5 % fill), threshold 0.5, Moore-8."""
    name = "schelling_4096x4096_13M"
    dtype = "i8/i32"
    default_steps = 1000
    stationary = False
    kernel = "schelling_bits_kernel"
    bound = "latency"
    bound_note = ("NOT an HBM-bound kernel: grid barriers + integer issue in the converged phase (bit planes are "
                  "L2-resident), dependent random 4-8 B accesses in the mover phase; `frac` on SURVEY 8(d) API bytes is "
                  "reported per the contract, `engine_layout.frac` is the bandwidth the packed layout really needs")
    l2_note = ("no flush (one persistent launch runs all K steps); active-phase working set "
               "(agents SoA + cell_agent + U/E lists, ~350 MB) exceeds the 126 MB L2; the bit-plane "
               "grid (4 MB) is L2-resident by design")

    def __init__(self, rank, grid=4096, n=13_000_000, shard=False):
        from jaxabm_b200.rules import schelling
        self.grid, self.n, self.seed, self.shard = grid, n, 42 + (0 if shard else rank), shard
        self.name = f"schelling_{grid}x{grid}_{n / 1e6:.3g}M" if (grid, n) != (4096, 13_000_000) else self.name
        if shard:
            self.kernel = "grid_shard_sweep_kernel"
            self.l2_note = ("no flush; 5 launches per step and rank (band sweep, counts + flag wait, movers out as requests "
                            "to the slot owners, slot owners forward to the cell owners, apply); the band's bit planes "
                            "are L2-resident, the active-phase working set is not")
        self.types, self.positions = schelling.initial_layout(grid, n, 0.5, self.seed)
        self.agents = n
        self._pins = None

    def fresh(self):
        import jaxabm_b200 as jx
        from jaxabm_b200.rules import schelling
        m = schelling.create_schelling_model(self.grid, self.n, seed=self.seed, types=self.types,
                                             positions=self.positions, config=jx.ModelConfig(seed=self.seed),
                                             shard=self.shard)
        m._dev.grid_rebuild()                       # cell binning of the uploaded layout (untimed setup)
        return m

    def movers(self, res):
        """agents moved during the (fresh) run: total_moves is cumulative since model creation"""
        return int(res["total_moves"][-1]) if len(res.get("total_moves", [])) else 0

    def unsat_sum(self, res):
        ps = np.array([float(v) for v in res["percent_satisfied"]], dtype=np.float64)
        return float(np.sum((1.0 - ps) * self.n))

    def api_bytes(self, res, K):
        """29 B/agent + 8 B/cell per step + 16 B per mover."""
        return (29 * self.n + 8 * self.grid * self.grid) * K + 16 * self.movers(res)

    def engine_bytes(self, res, K):
        """Bit-plane layout: 2 bits/cell sweep + 1 bit/cell mask per step; 4 B per unsatisfied agent
        (U list); per mover U/E/cell_agent/plane words/position/moves accesses = 52 B."""
        cells = self.grid * self.grid
        return (cells // 4 + cells // 8) * K + 4 * self.unsat_sum(res) + 52 * self.movers(res)

    def e2e(self, K):
        """Upload type+position from pinned host memory -> Model.run(K) -> read back
        position/moves/satisfied + the metrics history, all through the public API."""
        import torch
        if self._pins is None:
            self._pins = ({"type": _pin(self.types), "position": _pin(self.positions)},
                          {"position": _pinned_like((self.n, 2), torch.int32),
                           "moves": _pinned_like((self.n,), torch.int32),
                           "satisfied": _pinned_like((self.n,), torch.bool)})
        ins, outs = self._pins
        import jaxabm_b200 as jx
        from jaxabm_b200.rules import schelling
        t0 = time.perf_counter()
        m = schelling.create_schelling_model(self.grid, self.n, seed=self.seed, types=ins["type"].numpy(),
                                             positions=ins["position"].numpy(),
                                             config=jx.ModelConfig(seed=self.seed))
        res = m.run(steps=K)
        dev = m._dev
        for k, buf in outs.items():
            dev.download(0, dev.field_index(0, k), out=buf.numpy())
        wall = time.perf_counter() - t0
        h2d = sum(b.numel() * b.element_size() for b in ins.values())
        d2h = sum(b.numel() * b.element_size() for b in outs.values()) + K * (3 * 8 + 4)
        del m
        return wall, h2d, d2h, ("create model + upload type/position from pinned host memory, Model.run(K), "
                                "read back position/moves/satisfied + the metrics history")

    def cpu_run(self, steps):
        from oracle import cfast
        f = cfast.SchellingFast(self.grid, self.types, self.positions, seed=self.seed, mode=1)
        t0 = time.perf_counter()
        f.run(steps)
        return (time.perf_counter() - t0, cfast.num_threads(),
                f"first {steps} full-size steps of {self.name} (from the same seeded layout) on the C/OpenMP oracle")


class MarketWorkload:
    """C4-A: 45 M consumers + 5 M producers, well-mixed, env-level reductions fused in the step."""
    name = "market_45M_consumers_5M_producers"
    bound = "hbm"
    bound_note = "streaming SoA update"
    dtype = "f32"
    default_steps = 100
    stationary = True
    kernel = "step_kernel"
    l2_note = "no flush: per-step state traffic (0.98 GB) exceeds the 126 MB L2"

    def __init__(self, rank, nc=45_000_000, npr=5_000_000, world=1, shard=False):
        self.nc, self.npr, self.seed = nc, npr, 42 + (0 if shard else rank)
        self.rank, self.world, self.shard = rank, world, shard
        self.agents = nc + npr

    def fresh(self):
        import jaxabm_b200 as jx
        from jaxabm_b200.rules import market
        m = market.create_economy_model(num_consumers=self.nc, num_producers=self.npr,
                                        config=jx.ModelConfig(seed=self.seed))
        if self.shard:
            from jaxabm_b200 import sharding
            sharding.shard_model(m)
        m.initialize()
        return m

    def api_bytes(self, res, K):
        return (32 * self.nc + 24 * self.npr) * K

    def engine_bytes(self, res, K):
        # consumers: read savings+income, write savings+consumption+utility = 20 B; producers:
        # read capital, write capital+production+profit = 16 B (in place; 'income' is not rewritten)
        return (20 * self.nc + 16 * self.npr) * K

    def e2e(self, K):
        """initialize() on the device from the seed (the reference's init_state draws), run(K),
        read every state column back to pinned host memory."""
        import torch
        t0 = time.perf_counter()
        m = self.fresh()
        m.run(steps=K)
        d2h = 0
        for name, c in m.agent_collections.items():
            st = c.states
            for k in st:
                a = st[k]
                d2h += a.nbytes
        wall = time.perf_counter() - t0
        del m
        return wall, 0, d2h + K * 5 * 8, ("create + initialize on device (keys from the seed), Model.run(K), "
                                          "download all 7 state columns + the metrics history")

    def cpu_run(self, steps):
        from oracle import cfast
        n_c, n_p = self.nc // 10, self.npr // 10
        f = cfast.MarketFast(n_c, n_p, seed=self.seed, mode=1)
        t0 = time.perf_counter()
        f.run(steps)
        secs = (time.perf_counter() - t0) * 10.0
        return secs, cfast.num_threads(), (f"{steps} steps on a 1/10 population sample ({n_c}+{n_p} agents, "
                                            "time scaled x10) on the C/OpenMP oracle")


class EconomyWorkload:
    """C4-B: 49 M households + 1 M consumer-goods firms (advanced_economic_model.py), per-agent threefry
    draws every step, 14 fused env reductions, Gini by histogram rank."""
    name = "economy_49M_households_1M_firms"
    bound = "hbm"
    bound_note = "streaming SoA update with a heavy integer side (3 threefry blocks per household)"
    dtype = "f32"
    default_steps = 50
    stationary = True
    kernel = "economy_step_kernel"
    l2_note = "no flush: per-step state traffic (3.3 GB) exceeds the 126 MB L2"

    def __init__(self, rank, nh=49_000_000, nf=1_000_000, world=1, shard=False):
        self.nh, self.nf, self.seed = nh, nf, 42 + (0 if shard else rank)
        self.shard = shard
        self.agents = nh + nf

    def fresh(self):
        import jaxabm_b200 as jx
        from jaxabm_b200.rules import economy
        m = economy.create_economy_model(self.nh, self.nf, config=jx.ModelConfig(seed=self.seed))
        if self.shard:
            from jaxabm_b200 import sharding
            sharding.shard_model(m)
        m.initialize()
        return m

    def api_bytes(self, res, K):
        """households 57 B read + 57 B write, firms 73 B + 73 B."""
        return (114 * self.nh + 146 * self.nf) * K

    def engine_bytes(self, res, K):
        # households: read 7 f32 + employed, write 9 f32 + employed = 66 B (constant columns are not
        # rewritten); firms: read 11 f32 + age, write 11 f32 + age + is_active = 97 B
        return (66 * self.nh + 97 * self.nf) * K

    def e2e(self, K):
        t0 = time.perf_counter()
        m = self.fresh()
        m.run(steps=K)
        d2h = 0
        for name in ("income", "savings", "employed"):
            d2h += m.agent_collections["households"].states[name].nbytes
        wall = time.perf_counter() - t0
        del m
        return wall, 0, d2h + K * 29 * 8, ("create + initialize on device (keys from the seed), Model.run(K), "
                                           "download households' income/savings/employed + the metrics history")

    def cpu_run(self, steps):
        """1/49 population sample on the NumPy restatement (single thread; time scaled x49): the C/OpenMP oracle
        does not carry the economy rules."""
        from oracle import economy as oe, runtime as ort
        steps = min(steps, 5)
        nh, nf = self.nh // 49, self.nf // 49
        m = oe.create_economy_model(nh, nf, config=ort.ModelConfig(seed=self.seed, rng_mode=1))
        m.initialize()
        t0 = time.perf_counter()
        m.run(steps=steps)
        secs = (time.perf_counter() - t0) * 49.0
        self._cpu_steps = steps
        return secs, 1, (f"{steps} steps on a 1/49 population sample ({nh} households + {nf} firms, time scaled x49) "
                         "on the single-threaded NumPy restatement")


class WalkWorkload:
    """C1 scaled: 2^26 random walkers (48 B/agent-step), fused distance reductions."""
    name = "random_walk_2^26"
    bound = "hbm"
    bound_note = "streaming SoA update"
    dtype = "f32"
    default_steps = 50
    stationary = True
    kernel = "step_kernel"
    l2_note = "no flush: per-step state traffic (3.2 GB) exceeds the 126 MB L2"

    def __init__(self, rank, n=1 << 26):
        self.n, self.seed = n, 42 + rank
        self.agents = n

    def fresh(self):
        import jaxabm_b200 as jx
        from jaxabm_b200.rules import random_walk
        m = random_walk.create_scaled_walk_model(self.n, config=jx.ModelConfig(seed=self.seed))
        m.initialize()
        return m

    def api_bytes(self, res, K):
        return 48 * self.n * K

    engine_bytes = api_bytes

    def e2e(self, K):
        t0 = time.perf_counter()
        m = self.fresh()
        m.run(steps=K)
        d2h = 0
        st = m.agent_collections["walkers"].states
        for k in st:
            d2h += st[k].nbytes
        wall = time.perf_counter() - t0
        del m
        return wall, 0, d2h + K * 7 * 8, "create + initialize on device, Model.run(K), download all 4 state columns"

    def cpu_run(self, steps):
        """1/4 population sample on the C/OpenMP oracle (time scaled x4); uniform starts and velocities of the
        scaled variant's ranges (the walker update draws nothing, so the stream is irrelevant to its cost)."""
        from oracle import cfast
        n = self.n // 4
        rng = np.random.RandomState(self.seed)
        f = cfast.WalkFast(rng.uniform(0, 1, (n, 2)).astype(np.float32), rng.uniform(-0.01, 0.01, (n, 2)).astype(np.float32))
        f.run(1)
        t0 = time.perf_counter()
        f.run(steps)
        secs = (time.perf_counter() - t0) * 4.0
        return secs, cfast.num_threads(), (f"{steps} steps on a 1/4 population sample ({n} walkers, time scaled x4) "
                                            "on the C/OpenMP oracle")


class SirWorkload:
    """C3: SIR on a synthetic scale-free Network, 10 M agents / ~100 M adjacency entries."""
    name = "sir_10M_agents_100M_adjacency"
    bound = "hbm"
    bound_note = "CSR stream + random bitmap gathers / L2 reductions (L1tex-wavefront and L2-atomic bound around the epidemic's peak)"
    dtype = "i8/u32"
    default_steps = 100
    stationary = False
    kernel = "sir_step_kernel"
    l2_note = "no flush: adjacency (400 MB) is streamed every step and exceeds the 126 MB L2"

    def __init__(self, rank, n=10_000_000, m=5, shard=False):
        from jaxabm_b200 import synthetic
        self.n, self.seed, self.shard = n, 42 + (0 if shard else rank), shard
        if shard:
            self.kernel = "sir_pull_s_kernel"
        # the host copy of env['network_edges'] lives in pinned memory (the e2e call uploads it from there)
        self._edges_pin = _pin(synthetic.scale_free_edges(n, m, self.seed))
        self.edges = self._edges_pin.numpy()
        self.nnz = int(self.edges.shape[0])
        self.agents = n

    def fresh(self):
        import jaxabm_b200 as jx
        from jaxabm_b200.rules import sir
        m = sir.create_sir_model(self.n, self.edges, beta=0.05, gamma=0.1, initial_infected=0.01,
                                 seed=self.seed, config=jx.ModelConfig(seed=self.seed))
        if self.shard:
            from jaxabm_b200 import sharding
            sharding.shard_model(m)
        m.initialize()
        return m

    def api_bytes(self, res, K):
        """row_ptr 4(N+1) + col 4 nnz + state read 4N + state write 4N."""
        return (4 * (self.n + 1) + 4 * self.nnz + 8 * self.n) * K

    def engine_bytes(self, res, K):
        """row_ptr + col streamed; int8 state read+write; infected bitmap read + write."""
        return (4 * (self.n + 1) + 4 * self.nnz + 2 * self.n + 2 * (self.n // 8)) * K

    def e2e(self, K):
        t0 = time.perf_counter()
        m = self.fresh()
        m.run(steps=K)
        st = m.agent_collections["agents"].states["state"]
        wall = time.perf_counter() - t0
        del m
        return wall, self.edges.nbytes, st.nbytes + K * 3 * 8, ("create model, upload network_edges from pinned host memory and bin "
                                                               "them into CSR on the device, initialize, Model.run(K), download 'state'")

    def cpu_run(self, steps):
        """The first `steps` full-size steps of the same seeded epidemic on the C/OpenMP oracle."""
        from oracle import cfast
        f = cfast.SirFast(self.n, self.edges, beta=0.05, gamma=0.1, initial_infected=0.01, seed=self.seed, mode=1)
        t0 = time.perf_counter()
        f.run(steps)
        return (time.perf_counter() - t0, cfast.num_threads(),
                f"first {steps} full-size steps of {self.name} (same seeded network and start) on the C/OpenMP oracle")


class EnsembleWorkload:
    """C5: 8,192 parameter samples x 100k agents x K steps (K = 200 in the config) of
    tests/unit/test_analysis.py's create_test_model, one ensemble launch."""
    name = "sensitivity_sweep_8192x100k"
    bound = "smem"
    bound_note = "replica state lives in shared memory / L2: HBM sees only the result rows; reported against the HBM figure for comparability"
    dtype = "f32"
    default_steps = 200
    stationary = True
    kernel = "ensemble_kernel"
    l2_note = "replica state (400 KB) lives in one CTA's L2-resident slot for its whole run; HBM sees only results"

    def __init__(self, rank, world=1, samples=8192, n=100_000):
        self.samples_total, self.n = samples, n
        lo = samples * rank // world
        hi = samples * (rank + 1) // world
        self.lo, self.hi = lo, hi
        self.agents = (hi - lo) * n

    def plan(self, K):
        import jaxabm_b200 as jx
        from jaxabm_b200 import ensemble
        from jaxabm_b200.rules import growth
        rng = np.random.RandomState(0)
        g = rng.uniform(0.05, 0.2, self.samples_total)
        a = rng.uniform(0.05, 0.3, self.samples_total)
        models = [growth.create_test_model(params={"growth_rate": float(g[i]), "adjustment_rate": float(a[i])},
                                           config=jx.ModelConfig(seed=i + 1000, steps=K), num_agents=self.n)
                  for i in range(self.lo, self.hi)]
        return ensemble.plan(models)

    def api_bytes(self, res, K):
        return 8 * self.agents * K

    engine_bytes = api_bytes

    def cpu_run(self, steps):
        """16 of the replicas, each a whole serial run of the NumPy restatement (what the reference's sweep loop does,
        analysis.py:113-157), time scaled to all replicas of this rank."""
        from oracle import rules as orules, runtime as ort
        rng = np.random.RandomState(0)
        g = rng.uniform(0.05, 0.2, self.samples_total)
        a = rng.uniform(0.05, 0.3, self.samples_total)
        sample = 16
        t0 = time.perf_counter()
        for i in range(sample):
            m = orules.create_test_model(params={"growth_rate": float(g[i]), "adjustment_rate": float(a[i])},
                                         config=ort.ModelConfig(seed=i + 1000, steps=steps), num_agents=self.n)
            m.run()
        secs = (time.perf_counter() - t0) * ((self.hi - self.lo) / sample)
        return secs, 1, (f"{sample} of the {self.hi - self.lo} replicas x {steps} steps as serial runs of the single-threaded "
                         "NumPy restatement (time scaled to all replicas)")


WORKLOADS = {"schelling": SchellingWorkload, "market": MarketWorkload, "walk": WalkWorkload, "sir": SirWorkload,
             "ensemble": EnsembleWorkload, "economy": EconomyWorkload}


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
                                          "-lms", "20", "-i", str(self.device)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            time.sleep(0.25)                       # first sample lands before the timed region
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
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(workload, K):
    """profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from an
    `ncu --set full` capture, per launch.  Step-dependent workloads (one persistent launch = the whole timed
    window) store one figure per captured window length; a window that was never captured reports null."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(p)).get(workload)
    except Exception:
        return None
    if not t:
        return None
    if "by_steps" in t:
        hit = t["by_steps"].get(str(K))
        return None if hit is None else {"dram_bytes_per_launch": hit["dram_bytes"], "source": hit.get("source", t.get("source"))}
    return {"dram_bytes_per_launch": t.get("dram_bytes_per_launch"), "source": t.get("source")}


def cpu_record(wl, steps, warm=False):
    """CPU restatement of the workload (oracle/, all host cores where the restatement is threaded) on a bounded
    sample; None when the workload has none."""
    from oracle import cfast
    cfast.use_all_cores()                 # torchrun exports OMP_NUM_THREADS=1
    try:
        if warm:
            wl.cpu_run(1)
        secs, threads, what = wl.cpu_run(steps)
    except NotImplementedError:
        return None
    cpu_steps = getattr(wl, "_cpu_steps", steps)
    return {"value": wl.agents * cpu_steps / secs, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{what} ({secs:.1f} s)", "_steps": cpu_steps, "_secs": secs}


def reference_arm(args, rank, world):
    """CPU restatement of the reference's path (oracle/c, OpenMP over the host cores) on the same
    workload; under torchrun only rank 0 works.  JAX is not installable in this image, so the
    reference itself cannot run (DESIGN.md "Reference install")."""
    if rank != 0:
        return
    wl = make_workload(args, 0, 1, False)
    steps = args.steps if args.steps is not None else wl.default_steps
    cpu_steps = max(1, min(steps, args.cpu_steps * 5))        # bounded sample: at most 200 full-size steps
    cpu = cpu_record(wl, cpu_steps, warm=bool(args.warmup))
    if cpu is None:
        print(json.dumps({"impl": "reference", "unavailable": f"no CPU restatement of the {args.workload} workload's whole-run "
                                                               "launch is wired into bench.py"}), flush=True)
        return
    cpu_steps, secs = cpu.pop("_steps"), cpu.pop("_secs")
    cpu["sample"] += " (JAX is not installable here, so the reference itself cannot run)"
    value = cpu["value"]
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": cpu_steps, "warmup": min(args.warmup, 1), "ms_per_step": secs / cpu_steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": wl.dtype,
            "data": "synthetic", "config": {"workload": wl.name}, "cpu_baseline": cpu,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def make_workload(args, rank, world, shard, name=None):
    name = name or args.workload
    if name == "ensemble":
        return EnsembleWorkload(rank, world)
    if name == "market":
        return MarketWorkload(rank, world=world, shard=shard)
    if name == "economy":
        return EconomyWorkload(rank, world=world, shard=shard)
    if name == "schelling":
        n_ag = args.agents if args.agents is not None else (13_000_000 if args.grid == 4096 else int(args.grid * args.grid * 0.775))
        return SchellingWorkload(rank, grid=args.grid, n=n_ag, shard=shard and world > 1)
    if name == "sir":
        return SirWorkload(rank, shard=shard and world > 1)
    return WORKLOADS[name](rank)


class Ctx:
    """torch / torch.distributed plumbing of one bench process."""

    def __init__(self, rank, world, local):
        import torch
        import torch.distributed as td
        self.torch, self.td, self.rank, self.world, self.local = torch, td, rank, world, local

    def sync_all(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.td.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.td.all_reduce(t, op=self.td.ReduceOp.MAX)
        return float(t.item())


def measure(ctx, eng, args, name, K, shard=False, want_e2e=True):
    """One workload: W warm-up steps, exactly K device-timed steps (max over ranks), the dominant kernel's own
    CUDA-event time, and the end-to-end figure through the public API with host buffers."""
    rank, world = ctx.rank, ctx.world
    wl = make_workload(args, rank, world, shard, name)
    sharded = shard and world > 1 and name in ("market", "economy", "schelling", "sir")
    total_agents = wl.agents * (1 if sharded else world)
    if name == "ensemble":
        total_agents = wl.samples_total * wl.n
    extra = {}
    if name == "ensemble":
        from jaxabm_b200.device import ensemble_run
        desc, slots, params, seeds, env0 = wl.plan(K)
        for _ in range(args.warmup):
            ensemble_run(desc, slots, params[:64], seeds[:64], min(K, 20), env0)
        ctx.sync_all()
        l0 = eng.launch_count
        t0 = time.perf_counter()
        vals, dev_s = ensemble_run(desc, slots, params, seeds, K, env0)     # host params in, last metrics out
        wall = time.perf_counter() - t0
        ctx.sync_all()
        launches = eng.launch_count - l0
        max_s = ctx.max_over_ranks(dev_s)
        max_wall = ctx.max_over_ranks(wall)
        res = None
        ksecs, klaunches = dev_s, 1
        extra["runs_per_sec"] = wl.samples_total / max_s
        e2e = {"value": total_agents * K / max_wall, "unit": UNIT,
               "h2d_bytes_per_step": (params.nbytes + seeds.nbytes) / K, "d2h_bytes_per_step": vals.nbytes / K,
               "runs_per_sec": wl.samples_total / max_wall,
               "what": "jxb_ensemble_run with host parameter/seed tables in, last-metric rows out (wall clock)"}
    else:
        # ---- warm-up (W untimed steps) then the timed region: exactly K steps, device-timed ----------
        model = wl.fresh()
        model.run(steps=args.warmup)
        if not wl.stationary:
            del model
            model = wl.fresh()                     # the timed run starts from the seeded initial state
        ctx.sync_all()
        l0 = eng.launch_count
        res = model.run(steps=K)
        dev_s = model.last_device_seconds
        ctx.sync_all()
        launches = eng.launch_count - l0
        max_s = ctx.max_over_ranks(dev_s)
        # ---- dominant kernel: its own CUDA-event time -------------------------------------------------
        if name == "schelling" and not sharded and model._dev.profile()[2] != "grid_shard_sweep_kernel":
            ksecs, klaunches = dev_s, 1            # the persistent kernel IS the timed region (one launch)
            wl.kernel = model._dev.profile()[2]
        else:
            if not wl.stationary:
                del model
                model = None
            pm = model if wl.stationary else wl.fresh()
            if name == "schelling":
                wl.kernel = pm._dev.profile()[2]   # band kernels: whole-grid band on one GPU, or one band per rank
                if wl.kernel == "grid_shard_sweep_kernel":
                    wl.kernel = "grid_shard_{sweep,counts,moveout,forward,apply}_kernel (the 5 launches of a step, timed as one unit)"
            pm._dev.set_profile(True)              # events around every launch, no graph
            res = pm.run(steps=K)
            ksecs, klaunches, _ = pm._dev.profile()
            pm._dev.set_profile(False)
            model = pm
        del model
        # ---- end to end through the public API with host buffers ------------------------------------------
        e2e = None
        if want_e2e and not args.no_e2e and not sharded:
            wl.e2e(min(K, 3))                      # warm-up of the same call
            ctx.sync_all()
            wall, h2d, d2h, what = wl.e2e(K)
            ctx.torch.cuda.synchronize()
            e2e = {"value": total_agents * K / ctx.max_over_ranks(wall), "unit": UNIT, "h2d_bytes_per_step": h2d / K,
                   "d2h_bytes_per_step": d2h / K, "what": what}

    peak, peak_kind = load_peak()
    band = world if (sharded and name in ("schelling", "sir")) else 1     # a rank's launch covers its band / node range only
    api_b = wl.api_bytes(res, K) / max(klaunches, 1) / band
    eng_b = wl.engine_bytes(res, K) / max(klaunches, 1) / band
    per_launch = ksecs / max(klaunches, 1)
    traffic = load_traffic(name, K) if not sharded else None
    roofline = {"bound": wl.bound, "bound_note": wl.bound_note, "kernel": wl.kernel,
                "achieved": api_b / per_launch / 1e9, "peak": peak,
                "unit": "GB/s", "frac": api_b / per_launch / 1e9 / peak, "peak_kind": peak_kind,
                "traffic": traffic and traffic["dram_bytes_per_launch"], "traffic_source": traffic and traffic["source"],
                "bytes_per_launch": api_b, "us_per_launch": per_launch * 1e6, "launches_timed": int(klaunches),
                "kernel_share_of_step": min(1.0, ksecs / dev_s) if dev_s > 0 else None,
                "engine_layout": {"bytes_per_launch": eng_b, "achieved": eng_b / per_launch / 1e9,
                                  "frac": eng_b / per_launch / 1e9 / peak},
                "note": "achieved/frac use SURVEY.md 8(d) algorithmic bytes in the reference's API dtypes; "
                        "engine_layout uses the bytes the packed HBM layout actually has to move "
                        "(frac > 1 on the API figure means the engine moves fewer bytes than the reference layout implies)"}
    par = "single-gpu"
    if world > 1:
        par = (f"one grid in {world} row bands, one per gpu (movers travel source band -> slot owner -> target band as posted stores over NVLink peer memory)"
               if (sharded and name == "schelling") else
               f"one network in {world} node ranges, one per gpu (new infected-bitmap words stored into every rank's copy over NVLink peer memory)"
               if (sharded and name == "sir") else
               f"one population sharded over {world} gpus (per-step env partial-sum exchange over NVLink peer memory)" if sharded else
               (f"replica blocks over {world} gpus (no data-path collective, final gather)" if name == "ensemble" else f"replica-per-gpu x{world}"))
    rec = {"metric": METRIC, "value": total_agents * K / max_s, "unit": UNIT, "n_gpus": world, "steps": K,
           "warmup": args.warmup, "ms_per_step": max_s / K * 1e3, "higher_is_better": True,
           "scaling": "strong" if (name == "ensemble" or sharded) else "weak",
           "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
           "config": {"workload": wl.name, "agents_per_gpu": wl.agents, "parallelism": par, "l2": wl.l2_note,
                      "timed_region": ("K steps from the seeded initial state (fresh model after the warm-up model)"
                                       if not wl.stationary else "K steps after W warm-up steps on the same model")},
           "roofline": roofline, "cpu_baseline": None, "e2e": e2e, "gpu_launches": int(launches)}
    rec.update(extra)
    return rec, wl


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps K (default: the workload's run length)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="schelling", choices=sorted(WORKLOADS))
    ap.add_argument("--shard", action="store_true",
                    help="market / economy: split ONE population over the ranks; schelling: ONE grid in row bands")
    ap.add_argument("--grid", type=int, default=4096, help="schelling: grid side")
    ap.add_argument("--agents", type=int, default=None, help="schelling: agents (default 77.5 %% fill)")
    ap.add_argument("--cpu-steps", type=int, default=40, help="steps of the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-also", action="store_true",
                    help="default run only: skip the other BASELINE configs (`also`, and `sharded` at N > 1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return reference_arm(args, rank, world)

    ctx = Ctx(rank, world, local)
    torch, td = ctx.torch, ctx.td
    torch.cuda.set_device(local)
    if world > 1:
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as g
    if rank == 0:
        g.build()
    if world > 1:
        td.barrier()
    import jaxabm_b200 as jx  # noqa: F401
    from jaxabm_b200 import _native as nat

    eng = nat.engine()
    default_run = (args.workload == "schelling" and not args.shard and args.grid == 4096 and args.agents is None
                   and not args.no_also)
    # the clocks line covers the whole measurement (warm-ups, every timed region, the profiling passes and the
    # end-to-end calls): a single timed region of a few milliseconds is shorter than one nvidia-smi sample
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    K = args.steps if args.steps is not None else WORKLOADS[args.workload].default_steps
    line, wl = measure(ctx, eng, args, args.workload, K, shard=args.shard)
    if world == 1 and not args.no_cpu and rank == 0:
        cpu = cpu_record(wl, args.cpu_steps)
        if cpu:
            cpu.pop("_steps"), cpu.pop("_secs")
        line["cpu_baseline"] = cpu
    del wl
    if default_run and world == 1:
        # every other BASELINE.json config in the same run, same K / W (the driver runs the defaults only)
        also = []
        for name in ("walk", "sir", "market", "economy", "ensemble"):
            Kn = args.steps if args.steps is not None else WORKLOADS[name].default_steps
            rec, w2 = measure(ctx, eng, args, name, Kn)
            if not args.no_cpu:
                cpu = cpu_record(w2, min(args.cpu_steps, 20))
                if cpu:
                    cpu.pop("_steps"), cpu.pop("_secs")
                rec["cpu_baseline"] = cpu
            del w2
            for k in ("n_gpus", "warmup", "higher_is_better", "vs_baseline", "data"):
                rec.pop(k, None)
            also.append(rec)
        line["also"] = also
    if default_run and world > 1:
        # the two partitionings north_star names, over the same N GPUs: ONE C2 grid in row bands (strong scaling,
        # per-step exchange over NVLink peer memory) and the C5 sweep in replica blocks (no data-path collective)
        shd = []
        for name, sh in (("schelling", True), ("ensemble", False)):
            Kn = args.steps if args.steps is not None else WORKLOADS[name].default_steps
            rec, w2 = measure(ctx, eng, args, name, Kn, shard=sh, want_e2e=False)
            del w2
            for k in ("warmup", "higher_is_better", "vs_baseline", "data", "cpu_baseline", "e2e"):
                rec.pop(k, None)
            shd.append(rec)
        line["sharded"] = shd
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        td.barrier()
        td.destroy_process_group()
    if rank != 0:
        return
    line["clocks"] = clocks
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
