"""Shared helpers for the test-suite: synthetic inputs."""
import csv
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def cosmic():
    """COSMIC v3.3.1 SBS GRCh37 signatures, 96 x 79, columns sum to 1 (fixture copied
    from the reference's inst/extdata by tests/golden/make_golden.py)."""
    rows = list(csv.reader(open(os.path.join(GOLDEN, "COSMIC_v3.3.1_SBS_GRCh37.csv"))))
    names = rows[0][1:]
    ctx = [r[0] for r in rows[1:]]
    C = np.array([[float(x) for x in r[1:]] for r in rows[1:]])
    return C, names, ctx


def synth_counts(K, G, N, mu_T=4000.0, seed=0):
    """M ~ Poisson(P_true E_true); P_true = first N COSMIC columns (K = 96) or
    Dirichlet(0.1) columns; E_true ~ Gamma(1, mu_T / N)."""
    rng = np.random.default_rng(seed)
    if K == 96:
        P = cosmic()[0][:, :N]
    else:
        P = rng.dirichlet(np.full(K, 0.1), size=N).T
    E = rng.gamma(1.0, mu_T / N, size=(N, G))
    M = rng.poisson(P @ E).astype(np.float64)
    return M, P, E


def example_data():
    """inst/extdata/example_data.rds of the reference (tests/golden/make_golden.py):
    M 96 x 64 counts and the planted P 96 x 4 (COSMIC SBS58, SBS40, SBS26, SBS2)."""
    d = np.load(os.path.join(GOLDEN, "example_data.npz"))
    return d["M"].astype(np.float64), d["P"].astype(np.float64)
