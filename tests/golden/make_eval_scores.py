#!/usr/bin/env python
"""Golden objective scores from the reference's OWN functions: `__eval_nash__` and `__eval_q_pico__` of
/root/reference/wmf/wmf.py (wmf.py:749-754, 775-781), executed here from the unmodified source on seeded series (with NaN
gaps in the observations, as field records have).  Writes tests/golden/eval_scores.npz; tests/test_gpu_shia.py compares
`wmf_eval_scores` (the device-side scoring of the ensemble path) with it.

Run here (CPU, reference tree present):  python tests/golden/make_eval_scores.py"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_simubasin_trace as mst  # noqa: E402


def main():
    w = mst.load_wmf(mst.make_cu_backend(), mst.OracleModels())   # wmf.py reads cu.dxp at import time
    nash, qpico = getattr(w, "__eval_nash__"), getattr(w, "__eval_q_pico__")
    rng = np.random.default_rng(0)
    T, M = 300, 24
    t = np.arange(T)
    qobs = (5 + 40 * np.exp(-((t - 120) / 25.0) ** 2) + 15 * np.exp(-((t - 210) / 12.0) ** 2) + rng.normal(0, 0.5, T)).astype(np.float32)
    qobs[[7, 8, 9, 150, 299]] = np.nan                       # gaps in the record
    qsim = np.empty((M, T), np.float32)
    for m in range(M):
        a, s, d = rng.uniform(0.4, 1.6), rng.integers(-10, 11), rng.uniform(-3, 3)
        qsim[m] = np.clip(a * np.roll(np.nan_to_num(qobs, nan=5.0), s) + d + rng.normal(0, 0.3, T), 0, None)
    qsim[0] = np.nan_to_num(qobs, nan=5.0)                   # a perfect member
    want = np.array([[nash(qobs.astype(np.float64), qsim[m].astype(np.float64)),
                      qpico(qobs.astype(np.float64), qsim[m].astype(np.float64))] for m in range(M)], np.float64)
    np.savez(os.path.join(HERE, "eval_scores.npz"), qobs=qobs, qsim=qsim, scores=want)
    print("nash range", want[:, 0].min(), want[:, 0].max(), "qpico range", want[:, 1].min(), want[:, 1].max())


if __name__ == "__main__":
    main()
