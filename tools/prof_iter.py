"""Minimal driver for ncu captures of a whole Poisson-Gamma iteration at the C3 shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bayesnmf_b200 import Handle
from tests.util import synth_counts
from bayesnmf_b200.hyperpriors import fill_hyperprior_params

mu = float(sys.argv[1]) if len(sys.argv) > 1 else 4000.0
G = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
M, _, _ = synth_counts(96, G, 20, mu, seed=0)
h = Handle(M, 20, likelihood="poisson", prior="gamma", MH=False, seed=1)
for k, v in fill_hyperprior_params(None, "gamma", float(M.mean()), 20).items():
    h.set_hyper(k, v)
h.init_from_prior()
h.step(4)
print(h.timing())
