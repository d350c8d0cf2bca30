"""Multiple-histogram reweighting for NVT runs: the consumer of the energy histograms the sweep produces and the
third parity criterion of the north star ("heat-capacity curves from multihist must agree within statistical
error").  Host-side numpy restatement of src/multihist_NVT.jl:95-221, including its conventions that differ from
the sampler's: dE = (Emax - Emin) / n_bin (the sampler bins with n_bin - 1) and
k = 3.166811429e-6 (InputParams uses 3.16681196e-6).  Pinned to the reference's known-answer vector
test/multihist_test.jl:4-14 by tests/test_multihist.py."""
from __future__ import annotations

import numpy as np

K_MULTIHIST = 3.166811429 / 10 ** 6     # src/multihist_NVT.jl:18


def _quasiprob(betat, energy, hist, ncycles, beta, free_energy):
    """quasiprob (src/multihist_NVT.jl:59-85) for every bin m (vectorised over m): energy [m], hist [T][m]."""
    expo = -beta[:, None] * energy[None, :] - free_energy[:, None]                  # [T][m]
    offset = np.maximum(expo.max(axis=0), -10.0 ** 6) + np.log(10 ** 3)
    denom = (ncycles[:, None] * np.exp(expo - offset[None, :])).sum(axis=0)
    return hist.sum(axis=0) / denom * np.exp(-betat * energy - offset)


def multihistogram_NVT(t_grid, beta_grid, en_histogram, en_min, en_max, n_bin, conv_threshold=1e-3, max_iter=1000):
    """-> (temperatures [10 T], cv [10 T]).  en_histogram [T][n_bin + 2] as in Output.en_histogram (bin 0 and n_bin + 1 are
    the under / overflow bins and are ignored, src/multihist_NVT.jl:51)."""
    from .host import TempGrid
    t_grid = np.asarray(t_grid, dtype=np.float64)
    beta = np.asarray(beta_grid, dtype=np.float64)
    T = len(t_grid)
    hist = np.asarray(en_histogram, dtype=np.float64)[:, 1:n_bin + 1]
    d_e = (en_max - en_min) / n_bin
    energy = en_min + (np.arange(1, n_bin + 1) - 0.5) * d_e
    ncycles = hist.sum(axis=1)
    free_energy = np.zeros(T)
    for _ in range(max_iter):
        new = np.array([np.log(_quasiprob(beta[i], energy, hist, ncycles, beta, free_energy).sum()) for i in range(T)])
        delta = (np.abs(new - free_energy) ** 2).sum()
        free_energy = new
        if delta < conv_threshold:
            break
    res = TempGrid(t_grid[0], t_grid[-1], 10 * T)
    cv = np.zeros(10 * T)
    for i, (tt, bt) in enumerate(zip(res.t_grid, res.beta_grid)):
        q = _quasiprob(bt, energy, hist, ncycles, beta, free_energy)
        norm = q.sum()
        e1 = (q / norm * energy).sum()
        e2 = (q / norm * energy ** 2).sum()
        cv[i] = (e2 - e1 ** 2) / (K_MULTIHIST * tt ** 2)
    return res.t_grid, cv


def multihistogram_NPT(t_grid, beta_grid, ev_histogram, en_min, en_max, v_min, v_max, n_bin, pressure, conv_threshold=1e-3,
                       max_iter=1000):
    """src/multihist_NPT.jl:113-338 -> (temperatures [10 T], cp [10 T]).  ev_histogram [T][n_bin + 2][n_bin + 2] as in
    Output.ev_histogram (E index first); the under / overflow rows and columns are ignored (:56-60).  The weight of a bin
    is exp(-beta (E + p V)); dE = (Emax - Emin) / n_bin, dV = (Vmax - Vmin) / n_bin (:27-28)."""
    from .host import TempGrid
    t_grid = np.asarray(t_grid, dtype=np.float64)
    beta = np.asarray(beta_grid, dtype=np.float64)
    T = len(t_grid)
    hist = np.asarray(ev_histogram, dtype=np.float64)[:, 1:n_bin + 1, 1:n_bin + 1].reshape(T, -1)
    d_e, d_v = (en_max - en_min) / n_bin, (v_max - v_min) / n_bin
    energy = en_min + (np.arange(1, n_bin + 1) - 0.5) * d_e
    volume = v_min + (np.arange(1, n_bin + 1) - 0.5) * d_v
    enthalpy = (energy[:, None] + pressure * volume[None, :]).reshape(-1)        # E + pV per (m, n) bin
    e_flat = np.repeat(energy, n_bin)
    ncycles = hist.sum(axis=1)
    free_energy = np.zeros(T)
    for _ in range(max_iter):
        new = np.array([np.log(_quasiprob(beta[i], enthalpy, hist, ncycles, beta, free_energy).sum()) for i in range(T)])
        delta = (np.abs(new - free_energy) ** 2).sum()
        free_energy = new
        if delta < conv_threshold:
            break
    res = TempGrid(t_grid[0], t_grid[-1], 10 * T)
    cp = np.zeros(10 * T)
    for i, (tt, bt) in enumerate(zip(res.t_grid, res.beta_grid)):
        q = _quasiprob(bt, enthalpy, hist, ncycles, beta, free_energy)
        w = q / q.sum()
        h1, h2 = (w * enthalpy).sum(), (w * enthalpy ** 2).sum()
        cp[i] = (h2 - h1 ** 2) / (K_MULTIHIST * tt ** 2)
    return res.t_grid, cp
