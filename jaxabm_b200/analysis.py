"""Ensemble drivers: ``SensitivityAnalysis`` and ``ModelCalibrator`` with the entry points of
``jaxabm/analysis.py`` -- the per-sample model runs go to the device as ONE ensemble launch.

What is kept verbatim (observable behaviour):
* LHS sampling schedule ``analysis.py:67-95`` -- same ``jax.random`` calls (``split``,
  ``uniform``, sort-based ``permutation``) through the bit-compatible host key algebra, so
  the sample matrix is the one the reference would draw;
* per-sample seed ``i + 1000`` and default ``ModelConfig`` (``analysis.py:124-128``), "last
  value" extraction (``:146-157``), squared-correlation "sobol" proxy (``:167-203``);
* ``_evaluate_params_robust`` seeds: ``randint(split(key)[1], (), 0, 1_000_000)`` per run
  (``:438-441``), mean / 95 % CI / normalised loss (``:457-487``).

What differs (documented in DESIGN.md): the optimiser loops are plain NumPy host code with
the reference's hyper-parameters (their FLOPs are microscopic); population methods
(es / pso / cem) evaluate their whole population x evaluation runs in one launch; adam / sgd
use central finite differences through the ensemble instead of ``jax.grad`` through the
Python-unrolled simulation; the RL variants are not provided.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Tuple

import numpy as np

from . import ensemble
from . import random as jrandom
from .core import ModelConfig

f32 = np.float32


class SensitivityAnalysis:
    """``jaxabm/analysis.py:23-288``."""

    def __init__(self, model_factory: Callable, param_ranges: Dict[str, Tuple[float, float]],
                 metrics_of_interest: List[str], num_samples: int = 100, seed: int = 0):
        self.model_factory = model_factory
        self.param_ranges = param_ranges
        self.metrics_of_interest = metrics_of_interest
        self.num_samples = num_samples
        self.key = jrandom.PRNGKey(seed)
        self.samples = self._generate_lhs_samples()
        self.results: Optional[Dict[str, np.ndarray]] = None
        self.last_device_seconds = 0.0

    def _generate_lhs_samples(self) -> np.ndarray:                       # analysis.py:67-95
        keys = jrandom.split(self.key)
        self.key, subkey = keys[0], keys[1]
        n = self.num_samples
        points = np.linspace(0, 1, n + 1, dtype=f32)[:-1]
        points = (points + jrandom.uniform(subkey, (n,)) / f32(n)).astype(f32)
        samples = np.zeros((n, len(self.param_ranges)), dtype=f32)
        for i, _ in enumerate(self.param_ranges):
            keys = jrandom.split(self.key)
            self.key, subkey = keys[0], keys[1]
            samples[:, i] = jrandom.permutation(subkey, points)
        for i, (_, (lo, hi)) in enumerate(self.param_ranges.items()):
            samples[:, i] = (samples[:, i] * f32(hi - lo)).astype(f32)
            samples[:, i] = (samples[:, i] + f32(lo)).astype(f32)
        return samples

    def run(self, verbose: bool = True) -> Dict[str, np.ndarray]:        # analysis.py:97-165
        if verbose:
            print(f"Running sensitivity analysis with {self.num_samples} samples...")
        names = list(self.param_ranges.keys())
        results = {m: np.zeros(self.num_samples, dtype=f32) for m in self.metrics_of_interest}
        models = []
        for i in range(self.num_samples):
            params = {p: float(self.samples[i, j]) for j, p in enumerate(names)}
            config = ModelConfig(seed=i + 1000)                            # analysis.py:124-125
            models.append(self.model_factory(params=params, config=config))
        if ensemble.batchable(models):
            last, secs = ensemble.run_last_metrics(models)                # one launch, sharded over ranks
            self.last_device_seconds = secs
            for m in self.metrics_of_interest:
                if m in last:
                    results[m] = np.asarray(last[m], dtype=f32)
        else:
            for i, model in enumerate(models):
                res = model.run()
                data = res._data if hasattr(res, "_data") else res         # analysis.py:141-144
                for m in self.metrics_of_interest:
                    if m in data and data[m] is not None:
                        v = data[m]
                        v = v[-1] if hasattr(v, "__len__") and not isinstance(v, str) else v
                        results[m][i] = v
        if verbose:
            for i in range(self.num_samples):
                ps = ", ".join(f"{p}={float(self.samples[i, j]):.4f}" for j, p in enumerate(names))
                ms = ", ".join(f"{m}: {float(results[m][i]):.4f}" for m in self.metrics_of_interest)
                print(f"Sample {i + 1}/{self.num_samples}  {ps}  ->  {ms}")
            print("\nSensitivity analysis complete!")
        self.results = results
        return results

    def sobol_indices(self) -> Dict[str, Dict[str, float]]:             # analysis.py:167-203
        if self.results is None:
            raise ValueError("Must run sensitivity analysis before calculating indices")
        names = list(self.param_ranges.keys())
        out = {}
        for metric, values in self.results.items():
            values = np.asarray(values, dtype=f32)
            vn = (values - np.mean(values, dtype=f32)) / (np.std(values, dtype=f32) + f32(1e-8))
            idx = {}
            for i, p in enumerate(names):
                pv = self.samples[:, i]
                pn = (pv - np.mean(pv, dtype=f32)) / (np.std(pv, dtype=f32) + f32(1e-8))
                corr = np.mean((pn * vn).astype(f32), dtype=f32)
                idx[p] = float(corr ** 2)
            out[metric] = idx
        return out

    # ---- true Sobol indices (addition) -------------------------------------
    def run_saltelli(self, num_base: Optional[int] = None, steps: Optional[int] = None,
                     seed_base: int = 1000) -> Dict[str, Dict[str, Dict[str, float]]]:
        """Variance-based first-order (S1) and total-order (ST) Sobol indices with the Saltelli /
        Jansen estimators -- the reference's ``sobol_indices`` is a squared-correlation proxy
        (``analysis.py:167-203``) and stays the default; this is an additional method.

        Design: two independent ``num_base x P`` Latin-hypercube matrices A and B drawn with the same
        ``jax.random`` schedule as ``_generate_lhs_samples``, plus the P "radial" matrices AB_i (A with
        column i taken from B): ``num_base * (P + 2)`` model runs, evaluated as ONE ensemble launch
        (sharded over the ranks of the process group).  Every run of one base row shares the model
        seed ``seed_base + row`` (common random numbers), so the index measures the parameters,
        not the run-to-run noise.

        Returns ``{metric: {"S1": {param: v}, "ST": {param: v}}}`` and keeps the evaluations in
        ``self.saltelli_results``.
        """
        names = list(self.param_ranges.keys())
        P = len(names)
        n = int(num_base if num_base is not None else max(2, self.num_samples // (P + 2)))
        saved_n, saved_key = self.num_samples, self.key
        self.num_samples = n
        try:
            A = self._generate_lhs_samples()
            B = self._generate_lhs_samples()
        finally:
            self.num_samples = saved_n
        blocks = [A, B]
        for i in range(P):
            AB = A.copy()
            AB[:, i] = B[:, i]
            blocks.append(AB)
        design = np.concatenate(blocks, axis=0)                                # [(P+2) n, P]
        models = []
        for r in range(design.shape[0]):
            params = {p: float(design[r, j]) for j, p in enumerate(names)}
            models.append(self.model_factory(params=params, config=ModelConfig(seed=seed_base + r % n)))
        vals: Dict[str, np.ndarray] = {}
        if ensemble.batchable(models):
            last, secs = ensemble.run_last_metrics(models, steps)
            self.last_device_seconds = secs
            for m in self.metrics_of_interest:
                if m in last:
                    vals[m] = np.asarray(last[m], dtype=np.float64)
        else:
            cols = {m: np.zeros(len(models)) for m in self.metrics_of_interest}
            for r, model in enumerate(models):
                res = model.run(steps) if steps is not None else model.run()
                data = res._data if hasattr(res, "_data") else res
                for m in self.metrics_of_interest:
                    if m in data and data[m] is not None:
                        v = data[m]
                        cols[m][r] = v[-1] if hasattr(v, "__len__") and not isinstance(v, str) else v
            vals = cols
        out: Dict[str, Dict[str, Dict[str, float]]] = {}
        for m, y in vals.items():
            yA, yB = y[:n], y[n:2 * n]
            var = float(np.var(np.concatenate([yA, yB])))
            s1, st = {}, {}
            for i, p in enumerate(names):
                yAB = y[(2 + i) * n:(3 + i) * n]
                if var <= 0.0:
                    s1[p], st[p] = 0.0, 0.0
                    continue
                s1[p] = float(np.mean(yB * (yAB - yA)) / var)                    # Saltelli et al. 2010
                st[p] = float(0.5 * np.mean((yA - yAB) ** 2) / var)              # Jansen 1999
            out[m] = {"S1": s1, "ST": st}
        self.saltelli_design = design
        self.saltelli_results = vals
        return out

    def plot(self, metric=None, ax=None, **kwargs):                       # analysis.py:205-246
        import matplotlib.pyplot as plt
        if ax is None:
            _, ax = plt.subplots()
        indices = self.sobol_indices()
        if metric is None and self.metrics_of_interest:
            metric = self.metrics_of_interest[0]
        if metric in indices:
            items = sorted(indices[metric].items(), key=lambda x: x[1], reverse=True)
            ax.bar([p for p, _ in items], [v for _, v in items], **kwargs)
            ax.set_xlabel("Parameter")
            ax.set_ylabel("Sensitivity Index")
            ax.set_title(f"Sensitivity Indices for {metric}")
        return ax

    def plot_indices(self, figsize: Tuple[int, int] = (10, 6)):          # analysis.py:248-288
        try:
            import matplotlib.pyplot as plt
        except ImportError:
            raise ImportError("Matplotlib is required for plotting. Install it with 'pip install matplotlib'")
        indices = self.sobol_indices()
        metrics = list(indices.keys())
        params = list(indices[metrics[0]].keys())
        fig, ax = plt.subplots(figsize=figsize)
        x = np.arange(len(metrics))
        width = 0.8 / len(params)
        for i, p in enumerate(params):
            ax.bar(x + width * i - width * len(params) / 2 + width / 2,
                   [indices[m][p] for m in metrics], width, label=p)
        ax.set_xticks(x)
        ax.set_xticklabels(metrics)
        ax.legend(loc="best")
        return fig, ax


class ModelCalibrator:
    """``jaxabm/analysis.py:290-2752`` -- entry points and the evaluation path."""

    SUPPORTED = ("adam", "sgd", "es", "pso", "cem", "bayesian")
    RL_METHODS = ("q_learning", "policy_gradient", "actor_critic", "multi_agent_rl", "dqn")

    def __init__(self, model_factory: Callable, initial_params: Dict[str, float], target_metrics: Dict[str, float],
                 param_bounds: Optional[Dict[str, Tuple[float, float]]] = None,
                 metrics_weights: Optional[Dict[str, float]] = None, learning_rate: float = 0.01,
                 max_iterations: int = 100, method: str = "adam", loss_type: str = "mse",
                 evaluation_steps: int = 50, num_evaluation_runs: int = 3, tolerance: float = 1e-6,
                 patience: int = 10, seed: int = 0):
        self.model_factory = model_factory
        self.params = initial_params.copy()
        self.target_metrics = target_metrics
        self.param_bounds = param_bounds or {k: (0.01, 10.0) for k in initial_params}
        self.metrics_weights = metrics_weights or {k: 1.0 for k in target_metrics}
        self.learning_rate = learning_rate
        self.max_iterations = max_iterations
        self.method = method
        self.loss_type = loss_type
        self.evaluation_steps = evaluation_steps
        self.num_evaluation_runs = num_evaluation_runs
        self.tolerance = tolerance
        self.patience = patience
        self.key = jrandom.PRNGKey(seed)
        self.loss_history: List[float] = []
        self.param_history: List[Dict[str, float]] = []
        self.confidence_intervals: List[Any] = []
        self.best_params = initial_params.copy()
        self.best_loss = float("inf")
        self.device_seconds = 0.0
        if method in self.RL_METHODS:
            raise NotImplementedError(
                f"calibration method {method!r}: the RL optimisers of the reference are host-side "
                "optimisation logic outside this engine's scope (SURVEY.md section 2); use one of "
                f"{self.SUPPORTED}")
        if method not in self.SUPPORTED:
            raise ValueError(f"Unknown calibration method: {self.method}")    # analysis.py:398-399
        self._np_rng = np.random.RandomState(int(jrandom.bits(self.key)) & 0x7FFFFFFF)
        self._names = list(self.params.keys())
        self._lo = np.array([self.param_bounds[n][0] for n in self._names], dtype=np.float64)
        self._hi = np.array([self.param_bounds[n][1] for n in self._names], dtype=np.float64)

    # ---- losses (analysis.py:401-432, 478-487) ------------------------------------------------
    def _compute_loss(self, metrics: Dict[str, float]) -> float:
        loss = 0.0
        for metric, target in self.target_metrics.items():
            if metric not in metrics:
                continue
            value, w = metrics[metric], self.metrics_weights[metric]
            if self.loss_type == "mse":
                ml = (value - target) ** 2
            elif self.loss_type == "mae":
                ml = abs(value - target)
            elif self.loss_type == "huber":
                r = abs(value - target)
                ml = 0.5 * r ** 2 if r <= 1.0 else (r - 0.5)
            elif self.loss_type == "relative":
                ml = abs(value - target) / (abs(target) + 1e-8)
            else:
                raise ValueError(f"Unknown loss type: {self.loss_type}")
            loss += w * ml
        return loss

    def _compute_normalized_loss(self, metrics: Dict[str, float]) -> float:
        total = 0.0
        for metric, target in self.target_metrics.items():
            if metric in metrics:
                total += (abs(metrics[metric] - target) / (abs(target) + 1e-8)) ** 2
        return float(total)

    # ---- evaluation -------------------------------------------------------------------------
    def _draw_seeds(self, n: int) -> List[int]:
        """``analysis.py:438-441``: one ``split`` + ``randint(sub, (), 0, 1_000_000)`` per run."""
        seeds = []
        for _ in range(n):
            ks = jrandom.split(self.key)
            self.key, sub = ks[0], ks[1]
            seeds.append(int(jrandom.randint(sub, (), 0, 1_000_000)))
        return seeds

    def _evaluate_population(self, population: List[Dict[str, float]]):
        """All ``len(population) * num_evaluation_runs`` runs as one ensemble launch.
        Returns ``[(loss, confidence_intervals)]`` in population order; seeds are drawn in the
        order the reference's serial loop would draw them."""
        R = self.num_evaluation_runs
        models, seeds = [], []
        for params in population:
            for s in self._draw_seeds(R):
                seeds.append(s)
                models.append(self.model_factory(params=params, config=ModelConfig(seed=s)))
        per_model: List[Dict[str, float]] = []
        if ensemble.batchable(models):
            last, secs = ensemble.run_last_metrics(models, steps=self.evaluation_steps)
            self.device_seconds += secs
            for i in range(len(models)):
                per_model.append({m: float(last[m][i]) if m in last else 0.0 for m in self.target_metrics})
        else:
            for model in models:
                res = model.run(steps=self.evaluation_steps)
                row = {}
                for m in self.target_metrics:                               # analysis.py:447-455
                    if m in res and hasattr(res[m], "__len__") and len(res[m]) > 0:
                        row[m] = float(res[m][-1])
                    else:
                        row[m] = 0.0
                per_model.append(row)
        out = []
        for p in range(len(population)):
            rows = per_model[p * R:(p + 1) * R]
            mean_metrics, ci = {}, {}
            for m in self.target_metrics:
                vals = np.array([r[m] for r in rows], dtype=f32)
                mean_val, std_val = float(np.mean(vals, dtype=f32)), float(np.std(vals, dtype=f32))
                mean_metrics[m] = mean_val
                half = 1.96 * std_val / np.sqrt(len(vals))
                ci[m] = (mean_val - half, mean_val + half)
            out.append((self._compute_normalized_loss(mean_metrics), ci))
        return out

    def _evaluate_params_robust(self, params: Dict[str, float]):        # analysis.py:434-476
        return self._evaluate_population([params])[0]

    # ---- optimisers -------------------------------------------------------------------------
    def _as_dict(self, x) -> Dict[str, float]:
        return {n: float(x[j]) for j, n in enumerate(self._names)}

    def _track(self, loss: float, params: Dict[str, float]) -> None:
        if loss < self.best_loss:
            self.best_loss, self.best_params = float(loss), dict(params)
        self.loss_history.append(float(self.best_loss if self.method != "es" else loss))
        self.param_history.append(self.best_params.copy())

    def calibrate(self, verbose: bool = True) -> Dict[str, float]:      # analysis.py:1281-1307
        if verbose:
            print(f"Starting calibration with {self.method} method...")
            print(f"Target metrics: {self.target_metrics}")
            print(f"Parameter bounds: {self.param_bounds}")
        return getattr(self, "_calibrate_" + ("gradient" if self.method in ("adam", "sgd") else self.method))(verbose)

    def _calibrate_es(self, verbose):                                     # analysis.py:1395-1458
        P, sigma, n_elite = 20, 0.1, int(20 * 0.2)
        x0 = np.array([self.params[n] for n in self._names])
        pop = np.clip(x0[None, :] + self._np_rng.normal(size=(P, len(x0))) * sigma, self._lo, self._hi)
        for it in range(self.max_iterations):
            scores = np.array([l for l, _ in self._evaluate_population([self._as_dict(p) for p in pop])])
            order = np.argsort(scores)[:n_elite]
            self._track(scores[order[0]], self._as_dict(pop[order[0]]))
            mean = pop[order].mean(axis=0)
            pop = np.clip(mean[None, :] + self._np_rng.normal(size=pop.shape) * sigma, self._lo, self._hi)
            sigma *= 0.995
            if verbose:
                print(f"Iteration {it + 1}/{self.max_iterations}  best loss {scores[order[0]]:.6f}")
            if scores[order[0]] < self.tolerance:
                break
        return self.best_params

    def _calibrate_pso(self, verbose):                                    # analysis.py:1460-1521
        P, w, c1, c2 = 20, 0.7, 1.5, 1.5
        d = len(self._names)
        pos = self._lo + self._np_rng.uniform(size=(P, d)) * (self._hi - self._lo)
        vel = self._np_rng.normal(size=(P, d)) * 0.1
        pbest, pbest_s = pos.copy(), np.full(P, np.inf)
        gbest, gbest_s = pos[0].copy(), np.inf
        for it in range(self.max_iterations):
            scores = np.array([l for l, _ in self._evaluate_population([self._as_dict(p) for p in pos])])
            better = scores < pbest_s
            pbest[better], pbest_s[better] = pos[better], scores[better]
            if scores.min() < gbest_s:
                gbest_s, gbest = float(scores.min()), pos[int(scores.argmin())].copy()
            self._track(gbest_s, self._as_dict(gbest))
            r1, r2 = self._np_rng.uniform(size=vel.shape), self._np_rng.uniform(size=vel.shape)
            vel = w * vel + c1 * r1 * (pbest - pos) + c2 * r2 * (gbest[None, :] - pos)
            pos = np.clip(pos + vel, self._lo, self._hi)
            if verbose:
                print(f"Iteration {it + 1}/{self.max_iterations}  best loss {gbest_s:.6f}")
            if gbest_s < self.tolerance:
                break
        return self.best_params

    def _calibrate_cem(self, verbose):                                    # analysis.py:1523-1587
        P, n_elite = 50, int(50 * 0.2)
        mean = np.array([self.params[n] for n in self._names], dtype=np.float64)
        std = np.ones_like(mean) * 0.5
        for it in range(self.max_iterations):
            pop = np.clip(self._np_rng.normal(size=(P, len(mean))) * std + mean, self._lo, self._hi)
            scores = np.array([l for l, _ in self._evaluate_population([self._as_dict(p) for p in pop])])
            order = np.argsort(scores)[:n_elite]
            mean, std = pop[order].mean(axis=0), (pop[order].std(axis=0) + 1e-6) * 0.99
            self._track(scores[order[0]], self._as_dict(pop[order[0]]))
            if verbose:
                print(f"Iteration {it + 1}/{self.max_iterations}  best loss {scores[order[0]]:.6f}")
            if scores[order[0]] < self.tolerance:
                break
        return self.best_params

    def _calibrate_bayesian(self, verbose):                               # analysis.py:1589-1660 (its "simplified" search)
        n_init = 10
        X = self._lo + self._np_rng.uniform(size=(n_init, len(self._names))) * (self._hi - self._lo)
        y = np.array([l for l, _ in self._evaluate_population([self._as_dict(x) for x in X])])
        for x, l in zip(X, y):
            self._track(l, self._as_dict(x))
        for it in range(max(self.max_iterations - n_init, 0)):
            scale = 0.1 * (1.0 - it / self.max_iterations)
            cand = np.clip(X[int(np.argmin(y))] + self._np_rng.normal(size=len(self._names)) * scale, self._lo, self._hi)
            l, _ = self._evaluate_params_robust(self._as_dict(cand))
            X, y = np.vstack([X, cand]), np.append(y, l)
            self._track(l, self._as_dict(cand))
            if l < self.tolerance:
                break
        return self.best_params

    def _calibrate_gradient(self, verbose):
        """adam / sgd.  The reference differentiates through ``model.run`` with ``jax.grad``
        (``analysis.py:489-531,1309-1394``); this engine has no tape, so the gradient is a central
        finite difference with a fixed seed (42, as the reference's ``loss_fn``), evaluated as
        one ensemble launch of 2P replicas."""
        x = np.array([self.params[n] for n in self._names], dtype=np.float64)
        m, v, t = np.zeros_like(x), np.zeros_like(x), 0
        stall = 0

        def losses(points):
            models = [self.model_factory(params=self._as_dict(p), config=ModelConfig(seed=42)) for p in points]
            if ensemble.batchable(models):
                last, secs = ensemble.run_last_metrics(models, steps=self.evaluation_steps)
                self.device_seconds += secs
                return [self._compute_loss({k: float(last[k][i]) for k in self.target_metrics if k in last})
                        for i in range(len(models))]
            out = []
            for mod in models:
                res = mod.run(steps=self.evaluation_steps)
                out.append(self._compute_loss({k: float(res[k][-1]) for k in self.target_metrics if k in res}))
            return out

        for it in range(self.max_iterations):
            h = 1e-3 * np.maximum(np.abs(x), 1e-2)
            pts = [x]
            for j in range(len(x)):
                e = np.zeros_like(x); e[j] = h[j]
                pts += [x + e, x - e]
            ls = losses(pts)
            loss = ls[0]
            g = np.array([(ls[1 + 2 * j] - ls[2 + 2 * j]) / (2 * h[j]) for j in range(len(x))])
            prev_best = self.best_loss
            self._track(loss, self._as_dict(x))
            if self.method == "adam":
                t += 1
                m = 0.9 * m + 0.1 * g
                v = 0.999 * v + 0.001 * g * g
                x = x - self.learning_rate * (m / (1 - 0.9 ** t)) / (np.sqrt(v / (1 - 0.999 ** t)) + 1e-8)
            else:
                x = x - self.learning_rate * g
            x = np.clip(x, self._lo, self._hi)
            stall = stall + 1 if loss >= prev_best - self.tolerance else 0
            if verbose:
                print(f"Iteration {it + 1}/{self.max_iterations}  loss {loss:.6f}")
            if loss < self.tolerance or stall >= self.patience:
                break
        self.params = self._as_dict(x)
        return self.best_params

    def get_calibration_history(self) -> Dict[str, Any]:
        return {"loss": self.loss_history, "params": self.param_history,
                "confidence_intervals": self.confidence_intervals}
