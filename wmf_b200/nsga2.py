"""NSGA-II calibration loop -- the host-side driver around the GPU ensemble evaluation.

Reference: ``SimuBasin.Calib_NSGAII`` (wmf.py:4662-4733) with ``class nsgaii_element`` (wmf.py:4736-4874): a population of
calibration vectors is evaluated generation after generation (``__ejec_parallel__``: one ``run_shia`` per Pool worker,
wmf.py:872-878), scored with ``__eval_nash__`` / ``__eval_q_pico__`` (wmf.py:749-754, 775-781), and evolved with DEAP's
``selTournamentDCD`` + ``selNSGA2``.  DEAP is not a dependency here: the non-dominated sort, the crowding distance and
the two selections are restated below (Deb et al. 2002, as DEAP implements them); the evaluation of a generation is ONE
call of ``wmf_b200.ensemble.evaluate_population`` -- every member of the generation in a single wavefront launch
sequence per GPU, members sharded over the ranks, scores all-gathered over NCCL.

Differences from the reference, on purpose:
* calibration vectors have the 11 entries ``shia_v1`` requires (modelosv2.f90:201); the reference's element still builds
  10 (wmf.py:4833) and cannot run against its own Fortran any more.  The eleventh (aquifer capacity factor) defaults to 1.
* ``__mutacion__`` draws from ``uniform(lo, lo)`` (wmf.py:4858), i.e. it always returns the lower bound; here a mutated
  gene is drawn from ``uniform(lo, hi)``.
* every rank of a multi-GPU run draws the same random numbers (one seeded generator), so the populations stay identical
  across ranks without any exchange besides the score gather.
"""
from __future__ import annotations

import numpy as np

GENES = ("evp", "infil", "perco", "losses", "velRun", "velSub", "velSup", "velStream", "Hu", "Hg", "Haq")


class NsgaElement:
    """Rules for creating, crossing and mutating calibration vectors (``nsgaii_element``, wmf.py:4736-4874)."""

    def __init__(self, evp=(0, 1), infil=(1, 200), perco=(1, 40), losses=(0, 1), velRun=(0.1, 1), velSub=(0.1, 1),
                 velSup=(0.1, 1), velStream=(0.1, 1), Hu=(0.1, 1), Hg=(0.1, 1), Haq=(1, 1), probCruce=None,
                 probMutacion=None, rangosMutacion=None, MaxMinOptima=(1.0, -1.0)):
        self.ranges = np.array([evp, infil, perco, losses, velRun, velSub, velSup, velStream, Hu, Hg, Haq], np.float64)
        self.prob_cruce = np.full(11, 0.5) if probCruce is None else np.resize(np.asarray(probCruce, np.float64), 11)
        self.prob_mutacion = np.full(11, 0.5) if probMutacion is None else np.resize(np.asarray(probMutacion, np.float64), 11)
        self.rangos_mutacion = self.ranges.copy() if rangosMutacion is None else np.asarray(rangosMutacion, np.float64)
        if self.rangos_mutacion.shape[0] < 11:               # the reference's 10-row table: the last gene keeps its range
            self.rangos_mutacion = np.vstack([self.rangos_mutacion, self.ranges[self.rangos_mutacion.shape[0]:]])
        self.weights = np.asarray(MaxMinOptima, np.float64)  # (+1 Nash: maximise, -1 peak error: minimise)

    def crea_calibracion(self, rng) -> np.ndarray:
        lo, hi = self.ranges[:, 0], self.ranges[:, 1]
        return np.where(lo != hi, rng.uniform(lo, np.where(lo != hi, hi, lo + 1.0)), lo)

    def population(self, n: int, rng) -> np.ndarray:
        return np.stack([self.crea_calibracion(rng) for _ in range(n)])

    def cruce(self, a: np.ndarray, b: np.ndarray, rng) -> None:
        """Uniform gene swap where the draw exceeds the gene's crossing probability (wmf.py:4843-4850); in place."""
        swap = rng.uniform(0, 1, 11) > self.prob_cruce
        a[swap], b[swap] = b[swap].copy(), a[swap].copy()

    def mutacion(self, a: np.ndarray, rng) -> None:
        mut = rng.uniform(0, 1, 11) > self.prob_mutacion
        new = rng.uniform(self.rangos_mutacion[:, 0], np.maximum(self.rangos_mutacion[:, 1], self.rangos_mutacion[:, 0] + 1e-300))
        a[mut] = new[mut]


# ---- NSGA-II machinery (Deb et al. 2002): all objectives are maximised after multiplying by the weights ----
def dominates(a: np.ndarray, b: np.ndarray) -> bool:
    return bool(np.all(a >= b) and np.any(a > b))


def sort_nondominated(w: np.ndarray) -> list:
    """Fronts (lists of indices) of the weighted fitness table ``w`` (P, n_obj), best first."""
    P = w.shape[0]
    ge = np.all(w[:, None, :] >= w[None, :, :], axis=2)
    gt = np.any(w[:, None, :] > w[None, :, :], axis=2)
    dom = ge & gt                                            # dom[i, j]: i dominates j
    n_dominating = dom.sum(0).astype(np.int64)               # how many dominate j
    fronts, current = [], [int(i) for i in np.nonzero(n_dominating == 0)[0]]
    assigned = np.zeros(P, bool)
    while current:
        fronts.append(current)
        assigned[current] = True
        n_dominating = n_dominating - dom[current].sum(0)
        current = [int(i) for i in np.nonzero((n_dominating == 0) & ~assigned)[0]]
    return fronts


def crowding_distance(w: np.ndarray, front: list) -> np.ndarray:
    d = np.zeros(len(front))
    if len(front) <= 2:
        d[:] = np.inf
        return d
    f = w[front]
    for k in range(w.shape[1]):
        order = np.argsort(f[:, k], kind="stable")
        d[order[0]] = d[order[-1]] = np.inf
        span = f[order[-1], k] - f[order[0], k]
        if span > 0:
            d[order[1:-1]] += (f[order[2:], k] - f[order[:-2], k]) / (span * w.shape[1])
    return d


def sel_nsga2(w: np.ndarray, k: int):
    """Indices of the ``k`` individuals kept by NSGA-II and every individual's crowding distance."""
    chosen, crowd = [], np.zeros(w.shape[0])
    for front in sort_nondominated(w):
        cd = crowding_distance(w, front)
        crowd[front] = cd
        if len(chosen) + len(front) <= k:
            chosen.extend(front)
        else:
            order = np.argsort(-cd, kind="stable")
            chosen.extend([front[i] for i in order[: k - len(chosen)]])
        if len(chosen) >= k:
            break
    return chosen, crowd


def sel_tournament_dcd(w: np.ndarray, crowd: np.ndarray, k: int, rng) -> list:
    """Binary tournaments on dominance, then crowding distance (DEAP's selTournamentDCD); ``k`` multiple of 4."""
    def tourn(a, b):
        if dominates(w[a], w[b]):
            return a
        if dominates(w[b], w[a]):
            return b
        if crowd[a] != crowd[b]:
            return a if crowd[a] > crowd[b] else b
        return a if rng.uniform() <= 0.5 else b
    n = w.shape[0]
    chosen = []
    for i in range(0, k, 4):
        p1, p2 = rng.permutation(n), rng.permutation(n)
        j = i % (n - n % 4 if n >= 4 else n)
        for p in (p1, p2):
            q = [int(p[(j + t) % n]) for t in range(4)]
            chosen.append(tourn(q[0], q[1]))
            chosen.append(tourn(q[2], q[3]))
    return chosen[:k]


def calib_nsga2(evaluate, element: NsgaElement, pop_size: int = 40, ngen: int = 6, mutpb: float = 0.5, cxpb: float = 0.5,
                seed: int = 0):
    """The loop of ``Calib_NSGAII`` (wmf.py:4684-4733).  ``evaluate(calibs (P,11) float32) -> (P,2)`` returns
    ``[nash, qpico]`` per member; use ``make_gpu_evaluator``.  Returns ``(population (P,11), fitness (P,2), history)``."""
    pop_size += (-pop_size) % 4                               # multiple of 4 (wmf.py:4685-4690)
    rng = np.random.default_rng(seed)
    pop = element.population(pop_size, rng)
    fit = np.asarray(evaluate(pop.astype(np.float32)), np.float64)
    history = [fit.copy()]
    for _ in range(ngen):
        w = np.nan_to_num(fit * element.weights, nan=-np.inf, neginf=-1e300, posinf=1e300)
        _, crowd = sel_nsga2(w, pop_size)
        parents = sel_tournament_dcd(w, crowd, pop_size, rng)
        off = pop[parents].copy()
        off_fit = fit[parents].copy()
        changed = np.zeros(pop_size, bool)
        for i in range(0, pop_size - 1, 2):
            if rng.uniform() < cxpb:
                element.cruce(off[i], off[i + 1], rng)
                changed[i] = changed[i + 1] = True
        for i in range(pop_size):
            if rng.uniform() < mutpb:
                element.mutacion(off[i], rng)
                changed[i] = True
        # the reference re-runs the whole offspring generation (wmf.py:4712-4713) and re-scores the changed ones
        new_fit = np.asarray(evaluate(off.astype(np.float32)), np.float64)
        off_fit[changed] = new_fit[changed]
        allpop, allfit = np.vstack([pop, off]), np.vstack([fit, off_fit])
        keep, _ = sel_nsga2(np.nan_to_num(allfit * element.weights, nan=-np.inf, neginf=-1e300, posinf=1e300), pop_size)
        pop, fit = allpop[keep], allfit[keep]
        history.append(fit.copy())
    return pop, fit, history


def make_gpu_evaluator(models, qobs, n_cont, n_reg, rain, pos_evento, row: int = 0, members_per_launch: int = 256):
    """``evaluate`` for ``calib_nsga2`` on top of ``evaluate_population`` (all GPUs of the process group)."""
    from .ensemble import evaluate_population

    def evaluate(calibs):
        return evaluate_population(models, calibs, qobs, n_cont, n_reg, rain, pos_evento, row=row,
                                   members_per_launch=members_per_launch).cpu().numpy()
    return evaluate
