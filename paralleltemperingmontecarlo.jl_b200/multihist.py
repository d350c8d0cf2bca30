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


# ------------------------------------------------------------------------------------------------------------------------
# src/multihist.jl: the direct (linear-solve) multiple-histogram analysis behind `multihistogram(output, Tvec)` and
# `postprocess()` (:236-485).  Unlike multihist_NVT.jl it bins with dE = (Emax - Emin) / (n_bin - 1) and
# kB = 3.16681196e-6, like the sampler (readfile :52-73).
K_B = 3.16681196e-6


def _initialise(en_histogram, en_min, en_max, n_bin):
    """readfile(output, Tvals) :52-73 + processhist! :83-107: drop the overflow bins, divide the counts by n_bin, remove
    energy bins without counts."""
    hist = np.asarray(en_histogram, dtype=np.float64)
    if hist.shape[1] != n_bin:
        hist = hist[:, 1:-1]
    de = (en_max - en_min) / (n_bin - 1)
    energy = np.array([(j - 1) * de + en_min for j in range(1, n_bin + 1)])
    hist = hist / n_bin
    nsum = hist.sum(axis=0)
    keep = nsum != 0
    return hist[:, keep], energy[keep], nsum[keep]


def _systemsolver(hist, energy, beta, nsum):
    """bvector :159-199, amatrix :204-219, systemsolver :224-252, Entropycalc :257-264."""
    with np.errstate(divide="ignore", invalid="ignore"):
        logmat = np.log(hist) + beta[:, None] * energy[None, :]
        bmat = np.where(hist == 0.0, 0.0, hist * logmat)
    rhvec = bmat.sum(axis=0)
    b = (bmat - hist * rhvec[None, :] / nsum[None, :]).sum(axis=1)
    a = -(hist[:, None, :] * hist[None, :, :] / nsum[None, None, :]).sum(axis=2)
    a[np.diag_indices_from(a)] += hist.sum(axis=1)
    if np.isnan(a).any() or np.isnan(b).any():
        raise FloatingPointError("NaN in the multihistogram system")
    alpha = np.linalg.solve(a, b)
    entropy = (bmat - hist * alpha[:, None]).sum(axis=0) / nsum
    return alpha, entropy


def _analysis(energy, entropy, beta, k_b, n_points):
    """analysis :272-340 (self-scaling Boltzmann factors: the running offset nexp keeps 1 <= Z <= 100)."""
    tvec = 1.0 / (k_b * beta)
    d_t = (tvec[-1] - tvec[0]) / n_points
    t = np.array([(i - 1) * d_t + tvec[0] for i in range(1, n_points + 1)])
    z, u, cv, dcv, s_t = (np.zeros(n_points) for _ in range(5))
    nexp = 0.0
    for i in range(n_points):
        y = entropy - energy / (t[i] * k_b)
        while True:
            xp = np.exp(y - nexp)
            z[i] = xp.sum()
            if z[i] < 1.0:
                nexp -= 1.0
            elif z[i] > 100.0:
                nexp += 0.7
            else:
                break
        u[i] = (xp * energy).sum() / z[i]
        u2 = (xp * energy * energy).sum() / z[i]
        r2 = (xp * (energy - u[i]) ** 2).sum() / z[i]
        r3 = (xp * (energy - u[i]) ** 3).sum() / z[i]
        cv[i] = (u2 - u[i] * u[i]) / k_b / t[i] ** 2
        s_t[i] = u[i] / t[i] + k_b * np.log(z[i])
        dcv[i] = r3 / k_b ** 2 / t[i] ** 4 - 2 * r2 / k_b / t[i] ** 3
    return z, u, cv, dcv, s_t, t


def multihistogram(en_histogram, beta_grid, en_min, en_max, n_bin, outdir=None, n_points=1000):
    """multihistogram(output::Output, Tvec::TempGrid; outdir, NPoints) src/multihist.jl:456-461 -> dict with T, Z, Cv, dCv, U, S_T,
    alpha, energy, entropy.  With `outdir` the four files of run_multihistogram (:353-375) are written: histograms.data, Sol.X,
    S.data, analysis.NVT (columns T Z Cv dCv U(T))."""
    beta = np.asarray(beta_grid, dtype=np.float64)
    hist, energy, nsum = _initialise(en_histogram, en_min, en_max, n_bin)
    alpha, entropy = _systemsolver(hist, energy, beta, nsum)
    z, u, cv, dcv, s_t, t = _analysis(energy, entropy, beta, K_B, n_points)
    if outdir is not None:
        import pathlib
        from .readsave import _row
        d = pathlib.Path(outdir)
        d.mkdir(parents=True, exist_ok=True)
        with open(d / "histograms.data", "w") as f:
            f.write(_row([float(v) for v in energy], "\t"))
            for row in hist:
                f.write(_row([float(v) for v in row], "\t"))
        (d / "Sol.X").write_text("alpha\n" + _row([float(v) for v in alpha], "\t"))
        with open(d / "S.data", "w") as f:
            f.write("E\tEntropy\n")
            for e, s in zip(energy, entropy):
                f.write(_row([float(e), float(s)], "\t"))
        with open(d / "analysis.NVT", "w") as f:
            f.write("T\tZ\tCv\tdCv\tU(T)\n")
            for row in zip(t, z, cv, dcv, u):
                f.write(_row([float(v) for v in row], "\t"))
    return {"T": t, "Z": z, "Cv": cv, "dCv": dcv, "U": u, "S_T": s_t, "alpha": alpha, "energy": energy, "entropy": entropy}
