"""The NSGA-II loop of wmf_b200.nsga2 (replacement of SimuBasin.Calib_NSGAII, wmf.py:4662-4733) on the CPU with an
analytic evaluator: selection machinery against brute force, convergence on a two-objective toy problem."""
import numpy as np

from wmf_b200 import nsga2


def test_fronts_match_brute_force():
    rng = np.random.default_rng(0)
    w = rng.normal(size=(60, 2))
    fronts = nsga2.sort_nondominated(w)
    assert sorted(i for f in fronts for i in f) == list(range(60))
    first = {i for i in range(60) if not any(nsga2.dominates(w[j], w[i]) for j in range(60))}
    assert set(fronts[0]) == first
    for a, b in zip(fronts[:-1], fronts[1:]):                 # every member of a later front is dominated by an earlier one
        for j in b:
            assert any(nsga2.dominates(w[i], w[j]) for i in a)


def test_crowding_and_selection():
    w = np.array([[0.0, 1.0], [0.25, 0.75], [0.3, 0.7], [0.5, 0.5], [1.0, 0.0]])
    cd = nsga2.crowding_distance(w, [0, 1, 2, 3, 4])
    assert np.isinf(cd[0]) and np.isinf(cd[4]) and cd[2] < cd[3]
    keep, _ = nsga2.sel_nsga2(w, 4)
    assert 2 not in keep and set(keep) == {0, 1, 3, 4}        # the most crowded point goes first


def test_calibration_loop_converges_on_a_toy_problem():
    """Objectives shaped like the reference's: "Nash" = 1 - distance to a target calibration (maximised), "peak error" =
    |first gene - its target| in percent (minimised).  Six generations must improve both."""
    target = np.array([0.5, 100.0, 20.0, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 1.0])
    el = nsga2.NsgaElement()
    span = el.ranges[:, 1] - el.ranges[:, 0]
    span[span == 0] = 1.0

    def evaluate(c):
        z = (np.asarray(c, np.float64) - target) / span
        return np.stack([1.0 - (z ** 2).sum(1), 100.0 * np.abs(z[:, 1])], 1)

    pop, fit, hist = nsga2.calib_nsga2(evaluate, el, pop_size=42, ngen=8, seed=1)
    assert pop.shape == (44, 11) and fit.shape == (44, 2)
    assert np.all(pop >= el.ranges[:, 0] - 1e-9) and np.all(pop <= el.ranges[:, 1] + 1e-9)
    assert hist[-1][:, 0].max() >= hist[0][:, 0].max() and np.median(hist[-1][:, 0]) > np.median(hist[0][:, 0])
    assert hist[-1][:, 1].min() <= hist[0][:, 1].min()
    # same seed, same result (every rank of a multi-GPU run must evolve the same population)
    pop2, fit2, _ = nsga2.calib_nsga2(evaluate, el, pop_size=42, ngen=8, seed=1)
    np.testing.assert_array_equal(pop, pop2)
