"""Minimal driver for ncu captures of the sweep kernels at C5 / C4 shape (dev tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bayesnmf_b200 import Handle
from tests.util import synth_counts
which = sys.argv[1] if len(sys.argv) > 1 else "c5"
if which == "c5":
    K, G, N, lik, prior, MH = 1536, 50000, 40, "poisson", "exponential", True
else:
    K, G, N, lik, prior, MH = 96, 20000, 15, "normal", "truncnormal", False
M, _, _ = synth_counts(K, G, N, 4000.0, seed=0)
if lik == "normal":
    M = M + np.random.default_rng(1).normal(0, 0.05 * M.mean(axis=0, keepdims=True) + 1e-3, M.shape)
h = Handle(M, N, likelihood=lik, prior=prior, MH=MH, seed=1)
h.init_from_prior()
h.step(3)
t = h.timing()
print(which, t["iter_ms"] / 3)
